// TEST INFRASTRUCTURE ONLY: stand-in for base/Pose.hpp (base/types): Pose, Pose2D, getYaw, removeYaw, unknown<T>
#ifndef ESLAM_REF_SHIM_BASE_POSE
#define ESLAM_REF_SHIM_BASE_POSE
#include <limits>
#include "Eigen.hpp"
namespace base {
typedef Vector3d Position;
typedef Quaterniond Orientation;
// yaw = first angle of the (Z, Y, X) Euler decomposition of the rotation
inline double getYaw(const Orientation& q) { const Matrix3d R = q.toRotationMatrix(); return std::atan2(R(1, 0), R(0, 0)); }
inline Orientation removeYaw(const Orientation& q) { return Eigen::AngleAxisd(-getYaw(q), Eigen::Vector3d::UnitZ()) * q; }
template <class T> inline T unknown() { return std::numeric_limits<T>::quiet_NaN(); }
struct Pose {
    Position position; Orientation orientation;
    Pose() {}
    Pose(const Position& p, const Orientation& o) : position(p), orientation(o) {}
    Affine3d toTransform() const { Affine3d t; t.linear() = orientation.toRotationMatrix(); t.translation() = position; return t; }
};
struct Pose2D {
    Vector2d position; double orientation;
    Pose2D() : orientation(0) {}
    Pose2D(const Vector2d& p, double o) : position(p), orientation(o) {}
};
}
#endif
