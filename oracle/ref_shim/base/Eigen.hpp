// TEST INFRASTRUCTURE ONLY: stand-in for base/Eigen.hpp (base/types): the unaligned Eigen typedefs
#ifndef ESLAM_REF_SHIM_BASE_EIGEN
#define ESLAM_REF_SHIM_BASE_EIGEN
#include <Eigen/Core>
#include <Eigen/Geometry>
namespace base {
typedef Eigen::Vector2d Vector2d;
typedef Eigen::Vector3d Vector3d;
typedef Eigen::Matrix2d Matrix2d;
typedef Eigen::Matrix3d Matrix3d;
typedef Eigen::Quaterniond Quaterniond;
typedef Eigen::Affine3d Affine3d;
typedef Eigen::AngleAxisd AngleAxisd;
}
#endif
