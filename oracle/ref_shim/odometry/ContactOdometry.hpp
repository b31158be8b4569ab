// TEST INFRASTRUCTURE ONLY: stand-in for odometry/ContactOdometry.hpp (slam/odometry).  Only the calls the filter makes
// (src/PoseEstimator.cpp:188,192,198).  The upstream error model is not vendored; the sampler is DEFINED as
// mean + sigma o n with three normals per call, drawn through the same noise hook as the filter's own
// generators (site == this).
#ifndef ESLAM_REF_SHIM_ODOMETRY_CONTACTODOMETRY
#define ESLAM_REF_SHIM_ODOMETRY_CONTACTODOMETRY
#include <boost/random/linear_congruential.hpp>
#include <boost/random/normal_distribution.hpp>
#include "ContactState.hpp"
namespace odometry {
struct Configuration {};
class FootContact {
public:
    base::Pose delta;                    // mean pose delta of this step (body frame)
    double mean2d[3], sigma2d[3];        // x, y, theta
    base::Matrix3d positionError;
    boost::minstd_rand eng;
    explicit FootContact(const Configuration& = Configuration()) : eng(1u) {
        for (int i = 0; i < 3; ++i) { mean2d[i] = 0; sigma2d[i] = 0; }
    }
    base::Pose getPoseDelta() const { return delta; }
    base::Matrix3d getPositionError() const { return positionError; }
    base::Pose2D getPoseDeltaSample2D() {
        double n[3];
        for (int i = 0; i < 3; ++i)
            n[i] = eslam_ref_shim::noise_hook() ? eslam_ref_shim::noise_hook()(eslam_ref_shim::DRAW_NORMAL, this)
                                                : boost::normal_distribution<>(0, 1).sample(eng);
        return base::Pose2D(base::Vector2d(mean2d[0] + sigma2d[0] * n[0], mean2d[1] + sigma2d[1] * n[1]), mean2d[2] + sigma2d[2] * n[2]);
    }
};
}
#endif
