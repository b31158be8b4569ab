// TEST INFRASTRUCTURE ONLY: stand-in for odometry/ContactState.hpp (slam/odometry); fields as the reference uses them
// (test/testContactModel.cpp:72-77, src/ContactModel.cpp:21-41,140-155)
#ifndef ESLAM_REF_SHIM_ODOMETRY_CONTACTSTATE
#define ESLAM_REF_SHIM_ODOMETRY_CONTACTSTATE
#include <vector>
#include <base/Eigen.hpp>
#include <base/Pose.hpp>
#include <base/Time.hpp>
namespace odometry {
struct BodyContactPoint {
    base::Vector3d position;
    float contact;
    float slip;
    int groupId;
    BodyContactPoint() : contact(base::unknown<float>()), slip(base::unknown<float>()), groupId(-1) {}
};
struct BodyContactState {
    base::Time time;
    std::vector<BodyContactPoint> points;
};
}
#endif
