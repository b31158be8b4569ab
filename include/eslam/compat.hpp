// compat.hpp -- minimal stand-ins for the un-vendored types that appear in the reference's kept signatures
// (base/types, Eigen, odometry, envire, terrain_estimator: SURVEY.md §7.1 step 0).  Only what the particle-filter update
// path touches.  With the real packages on the include path, drop this header and include theirs instead; the facade
// classes only use the members defined here.
#ifndef ESLAM_COMPAT_HPP
#define ESLAM_COMPAT_HPP
#include <cmath>
#include <limits>
#include <vector>

namespace base {

struct Vector2d {
    double v[2];
    Vector2d() { v[0] = v[1] = 0; }
    Vector2d(double x, double y) { v[0] = x; v[1] = y; }
    double& x() { return v[0]; } double& y() { return v[1]; }
    double x() const { return v[0]; } double y() const { return v[1]; }
};

struct Vector3d {
    double v[3];
    Vector3d() { v[0] = v[1] = v[2] = 0; }
    Vector3d(double x, double y, double z) { v[0] = x; v[1] = y; v[2] = z; }
    double& x() { return v[0]; } double& y() { return v[1]; } double& z() { return v[2]; }
    double x() const { return v[0]; } double y() const { return v[1]; } double z() const { return v[2]; }
    double operator[](int i) const { return v[i]; }
    double& operator[](int i) { return v[i]; }
    Vector3d operator-(const Vector3d& o) const { return Vector3d(v[0] - o.v[0], v[1] - o.v[1], v[2] - o.v[2]); }
    Vector3d operator+(const Vector3d& o) const { return Vector3d(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
    bool operator==(const Vector3d& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2]; }
    double norm() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
    static Vector3d UnitZ() { return Vector3d(0, 0, 1); }
};

struct Quaterniond {
    double w, x, y, z;
    Quaterniond() : w(1), x(0), y(0), z(0) {}
    Quaterniond(double w_, double x_, double y_, double z_) : w(w_), x(x_), y(y_), z(z_) {}
    static Quaterniond Identity() { return Quaterniond(); }
    static Quaterniond fromYaw(double a) { return Quaterniond(std::cos(a / 2), 0, 0, std::sin(a / 2)); }
    Quaterniond operator*(const Quaterniond& o) const {
        return Quaterniond(w * o.w - x * o.x - y * o.y - z * o.z, w * o.x + x * o.w + y * o.z - z * o.y,
                           w * o.y - x * o.z + y * o.w + z * o.x, w * o.z + x * o.y - y * o.x + z * o.w);
    }
    void toRotationMatrix(double R[9]) const {
        R[0] = 1 - 2 * (y * y + z * z); R[1] = 2 * (x * y - w * z); R[2] = 2 * (x * z + w * y);
        R[3] = 2 * (x * y + w * z); R[4] = 1 - 2 * (x * x + z * z); R[5] = 2 * (y * z - w * x);
        R[6] = 2 * (x * z - w * y); R[7] = 2 * (y * z + w * x); R[8] = 1 - 2 * (x * x + y * y);
    }
    Vector3d operator*(const Vector3d& p) const {
        double R[9]; toRotationMatrix(R);
        return Vector3d(R[0] * p[0] + R[1] * p[1] + R[2] * p[2], R[3] * p[0] + R[4] * p[1] + R[5] * p[2],
                        R[6] * p[0] + R[7] * p[1] + R[8] * p[2]);
    }
};

// base::getYaw / base::removeYaw (used at src/PoseEstimator.cpp:186-187, src/ContactModel.cpp:35): yaw is the first
// angle of the (Z, Y, X) Euler decomposition; removeYaw = AngleAxis(-yaw, Z) * q.
inline double getYaw(const Quaterniond& q) {
    double R[9]; q.toRotationMatrix(R);
    return std::atan2(R[3], R[0]);
}
inline Quaterniond removeYaw(const Quaterniond& q) { return Quaterniond::fromYaw(-getYaw(q)) * q; }

struct Affine3d {                       // Eigen::Affine3d: linear part (row major) + translation
    double R[9]; Vector3d t;
    Affine3d() { for (int i = 0; i < 9; ++i) R[i] = (i % 4 == 0) ? 1.0 : 0.0; }
    static Affine3d Identity() { return Affine3d(); }
    static Affine3d fromTranslationYaw(const Vector3d& tr, double yaw) {
        Affine3d a; a.t = tr;
        const double c = std::cos(yaw), s = std::sin(yaw);
        a.R[0] = c; a.R[1] = -s; a.R[3] = s; a.R[4] = c;
        return a;
    }
    const Vector3d& translation() const { return t; }
    Vector3d& translation() { return t; }
    Quaterniond rotation() const {      // Quaterniond(affine.linear())
        const double tr = R[0] + R[4] + R[8];
        Quaterniond q;
        if (tr > 0) { double s = std::sqrt(tr + 1.0) * 2; q.w = 0.25 * s; q.x = (R[7] - R[5]) / s; q.y = (R[2] - R[6]) / s; q.z = (R[3] - R[1]) / s; }
        else if (R[0] > R[4] && R[0] > R[8]) { double s = std::sqrt(1.0 + R[0] - R[4] - R[8]) * 2; q.w = (R[7] - R[5]) / s; q.x = 0.25 * s; q.y = (R[1] + R[3]) / s; q.z = (R[2] + R[6]) / s; }
        else if (R[4] > R[8]) { double s = std::sqrt(1.0 + R[4] - R[0] - R[8]) * 2; q.w = (R[2] - R[6]) / s; q.x = (R[1] + R[3]) / s; q.y = 0.25 * s; q.z = (R[5] + R[7]) / s; }
        else { double s = std::sqrt(1.0 + R[8] - R[0] - R[4]) * 2; q.w = (R[3] - R[1]) / s; q.x = (R[2] + R[6]) / s; q.y = (R[5] + R[7]) / s; q.z = 0.25 * s; }
        return q;
    }
    Affine3d inverse() const {
        Affine3d o;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) o.R[3 * i + j] = R[3 * j + i];
        o.t = Vector3d(-(o.R[0] * t[0] + o.R[1] * t[1] + o.R[2] * t[2]), -(o.R[3] * t[0] + o.R[4] * t[1] + o.R[5] * t[2]),
                       -(o.R[6] * t[0] + o.R[7] * t[1] + o.R[8] * t[2]));
        return o;
    }
    Affine3d operator*(const Affine3d& b) const {
        Affine3d o;
        for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j)
            o.R[3 * i + j] = R[3 * i] * b.R[j] + R[3 * i + 1] * b.R[3 + j] + R[3 * i + 2] * b.R[6 + j];
        o.t = (*this) * b.t;
        return o;
    }
    Vector3d operator*(const Vector3d& p) const {
        return Vector3d(R[0] * p[0] + R[1] * p[1] + R[2] * p[2] + t[0], R[3] * p[0] + R[4] * p[1] + R[5] * p[2] + t[1],
                        R[6] * p[0] + R[7] * p[1] + R[8] * p[2] + t[2]);
    }
    double rotationAngle() const {      // Eigen::AngleAxisd(linear()).angle()
        double c = (R[0] + R[4] + R[8] - 1.0) / 2.0;
        if (c > 1) c = 1;
        if (c < -1) c = -1;
        return std::acos(c);
    }
};

struct Pose2D {
    Vector2d position; double orientation;
    Pose2D() : orientation(0) {}
    Pose2D(const Vector2d& p, double o) : position(p), orientation(o) {}
};

struct Pose {
    Vector3d position; Quaterniond orientation;
    Pose() {}
    Pose(const Vector3d& p, const Quaterniond& q) : position(p), orientation(q) {}
    Affine3d toTransform() const {
        Affine3d a; orientation.toRotationMatrix(a.R); a.t = position; return a;
    }
};

template <class T> inline T unknown() { return std::numeric_limits<T>::quiet_NaN(); }

struct Time {                           // base/Time.hpp: microseconds since the epoch
    long long microseconds;
    Time() : microseconds(0) {}
    static Time fromMicroseconds(long long us) { Time t; t.microseconds = us; return t; }
    bool isNull() const { return microseconds == 0; }
};

struct Matrix2d { double m[2][2]; Matrix2d() { m[0][0] = m[0][1] = m[1][0] = m[1][1] = 0; } };

}  // namespace base

namespace envire {
// envire/tools/GaussianMixture.hpp as PoseDistribution names it (src/PoseParticle.hpp:88-114): a list of weighted Gaussians.
// The reference never fills it inside this library (the Rock task that logs the distribution does).
template <class S, int D, class Adapter> struct GaussianMixture {
    struct Parameter {
        S weight; typename Adapter::Vector mean; typename Adapter::Matrix cov;
        Parameter() : weight(0) {}
    };
    std::vector<Parameter> params;
};
}  // namespace envire

namespace odometry {

struct BodyContactPoint {               // odometry/ContactState.hpp (fields as used at test/testContactModel.cpp:72-77)
    base::Vector3d position;
    float contact;
    float slip;
    int groupId;
    BodyContactPoint() : contact(base::unknown<float>()), slip(0), groupId(-1) {}
};

struct BodyContactState {
    std::vector<BodyContactPoint> points;
};

struct Configuration {};                // odometry::Configuration: error-model parameters live upstream

// The slice of odometry::FootContact the filter calls (src/PoseEstimator.cpp:188,192,198; src/EmbodiedSlamFilter.cpp:357).
// The upstream error model is not vendored; this stand-in carries the per-step mean delta and its standard deviation
// (getPoseDeltaSample2D() == mean + sigma o n) and derives the mean delta from successive body poses.
class FootContact {
public:
    explicit FootContact(const Configuration& = Configuration())
        : z_var(2e-6), have_last(false) { sigma[0] = sigma[1] = 0.002; sigma[2] = 0.001; }
    void setSampleSigma(double sx, double sy, double stheta) { sigma[0] = sx; sigma[1] = sy; sigma[2] = stheta; }
    void setPositionErrorZ(double var) { z_var = var; }
    void update(const base::Affine3d& body2odometry) {          // mean delta in the previous body frame
        if (have_last) delta = last.inverse() * body2odometry; else delta = base::Affine3d::Identity();
        last = body2odometry; have_last = true;
    }
    void setPoseDelta(const base::Affine3d& d) { delta = d; }
    base::Pose getPoseDelta() const { return base::Pose(delta.translation(), delta.rotation()); }
    double getPositionErrorZZ() const { return z_var; }             // getPositionError()(2,2)
    const double* getSampleSigma() const { return sigma; }
private:
    base::Affine3d delta, last; double sigma[3]; double z_var; bool have_last;
};

}  // namespace odometry

namespace terrain_estimator { struct TerrainClassification { int wheel_idx; }; }

namespace eslam {
// Dense stand-in for envire::MLSMap as the filter reads it through GridAccess::get (src/PoseEstimator.hpp:97-105):
// one surface patch {mean, stdev} per cell, NaN mean = no patch; cell (ix, iy) covers
// [offset + i*scale, offset + (i+1)*scale).  cells is row major: cells[2*(iy*width+ix) + {0,1}].
struct DenseGrid {
    std::vector<float> cells; int width, height; double scaleX, scaleY, offsetX, offsetY;
    DenseGrid() : width(0), height(0), scaleX(1), scaleY(1), offsetX(0), offsetY(0) {}
    DenseGrid(int w, int h, double sx, double sy, double ox, double oy)
        : cells((size_t)w * h * 2, std::numeric_limits<float>::quiet_NaN()), width(w), height(h), scaleX(sx), scaleY(sy),
          offsetX(ox), offsetY(oy) {}
    void set(int ix, int iy, float mean, float stdev) { cells[2 * ((size_t)iy * width + ix)] = mean; cells[2 * ((size_t)iy * width + ix) + 1] = stdev; }
};
}  // namespace eslam

#endif
