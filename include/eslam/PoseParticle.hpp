// PoseParticle.hpp -- particle and contact-point types of the reference (src/PoseParticle.hpp:20-86), hot fields only:
// the debug vectors cpoints/spoints (logDebug) stay on the host side of the reference and are not mirrored.
#ifndef ESLAM_FACADE_POSEPARTICLE_HPP
#define ESLAM_FACADE_POSEPARTICLE_HPP
#include <limits>
#include <vector>
#include "compat.hpp"

namespace eslam {

struct ContactPoint {                                   // src/PoseParticle.hpp:20-43
    ContactPoint() : point(0, 0, 0), zdiff(std::numeric_limits<double>::infinity()),
                     zvar(std::numeric_limits<double>::infinity()), prob(1.0) {}
    ContactPoint(const base::Vector3d& point, double zdiff, double zvar) : point(point), zdiff(zdiff), zvar(zvar), prob(1.0) {}
    bool operator<(const ContactPoint& other) const { return zdiff < other.zdiff; }
    base::Vector3d point; double zdiff, zvar, prob;
};

struct PoseParticle {                                   // src/PoseParticle.hpp:52-86
    PoseParticle() : orientation(0), zPos(0), zSigma(0), mprob(0), floating(true), n_contacts(0), weight(0) {}
    PoseParticle(const base::Vector2d& position, double orientation, double zpos = 0, double zsigma = 0, bool floating = true)
        : position(position), orientation(orientation), zPos(zpos), zSigma(zsigma), mprob(0), floating(floating),
          n_contacts(0), weight(0) {}
    base::Affine3d getPose(const base::Quaterniond& _orientation) const {   // :58-67
        base::Affine3d t = base::Affine3d::fromTranslationYaw(base::Vector3d(position.x(), position.y(), zPos), orientation);
        base::Affine3d r; base::removeYaw(_orientation).toRotationMatrix(r.R);
        return t * r;
    }
    base::Vector2d position; double orientation;
    double zPos, zSigma;
    double mprob; bool floating;
    unsigned n_contacts;                                // cpoints.size() of the last update
    double weight;
};

typedef PoseParticle PoseParticleGA;                    // src/PoseEstimator.hpp:108-117: the per-particle GridAccess lives on the device

// The logging / transport snapshot of the filter (src/PoseParticle.hpp:88-114): what the task around the library writes
// out every logParticlePeriod steps.  particles = getParticles() (the host mirror of the device state).
struct PoseDistribution {
    struct BaseAdapter {
        enum { Dimension = 2 };
        typedef double Scalar;
        typedef base::Vector2d Vector;
        typedef base::Matrix2d Matrix;
    };
    typedef envire::GaussianMixture<double, 2, BaseAdapter> GMM;
    base::Time time;
    std::vector<PoseParticle> particles;
    GMM gmm;
    base::Quaterniond orientation;
    odometry::BodyContactState bodyState;
};

}  // namespace eslam
#endif
