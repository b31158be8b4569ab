// EmbodiedSlamFilter.hpp -- eslam::EmbodiedSlamFilter (src/EmbodiedSlamFilter.hpp:27-75), the contact-update slice:
// init, update(body2odometry, BodyContactState, ltc), getParticles, getBestParticleIndex, getCentroid.  The laser /
// distance-image overloads feed envire's scan pipeline, which is out of scope.
#ifndef ESLAM_FACADE_EMBODIEDSLAMFILTER_HPP
#define ESLAM_FACADE_EMBODIEDSLAMFILTER_HPP
#include "PoseEstimator.hpp"

namespace eslam {

class EmbodiedSlamFilter {
    eslam::Configuration eslamConfig;
    odometry::Configuration odometryConfig;
    odometry::FootContact odometry;
    eslam::PoseEstimator filter;
    base::Affine3d udPose;                                   // pose of the last measurement update

public:
    EmbodiedSlamFilter(const odometry::Configuration& odometryConfig, const eslam::Configuration& eslamConfig)   // src/EmbodiedSlamFilter.cpp:13-23
        : eslamConfig(eslamConfig), odometryConfig(odometryConfig), odometry(odometryConfig), filter(odometry, eslamConfig) {}

    // init(env, pose, useSharedMap): the environment must hold a map (:77-104), particles are drawn around the pose (:118-125)
    void init(const DenseGrid* map, const base::Pose& pose, bool useSharedMap = true,
              const SurfaceHashConfig& = SurfaceHashConfig()) {
        if (useSharedMap && (!map || map->width == 0))
            throw std::runtime_error("The provided environment does not contain an mls grid.");    // :104
        const double angle = base::getYaw(pose.orientation);
        filter.init((int)eslamConfig.particleCount,
                    base::Pose2D(base::Vector2d(pose.position.x(), pose.position.y()), angle),
                    base::Pose2D(base::Vector2d(eslamConfig.initialTranslationError.x(), eslamConfig.initialTranslationError.y()),
                                 eslamConfig.initialRotationError.z()),
                    pose.position.z(), eslamConfig.initialTranslationError.z() + 1e-3);              // :118-125
        udPose = base::Affine3d::Identity(); udPose.translation() = base::Vector3d(1000, 0, 0);      // :128
        if (map) filter.setEnvironment(*map);                                                        // :131-134
    }

    bool update(const base::Affine3d& body2odometry, const odometry::BodyContactState& bs,
                const std::vector<terrain_estimator::TerrainClassification>& ltc) {                  // :353-369
        base::Quaterniond orientation = body2odometry.rotation();
        odometry.update(body2odometry);                                                              // :357
        filter.project(bs, orientation);                                                             // :358
        if (eslamConfig.measurementThreshold.test(udPose.inverse() * body2odometry) || ltc.size() > 0) {   // :360
            filter.update(bs, orientation, ltc);
            udPose = body2odometry;
            return true;
        }
        return false;
    }

    std::vector<eslam::PoseEstimator::Particle>& getParticles() { return filter.getParticles(); }   // :371-374
    size_t getBestParticleIndex() { return filter.getBestParticleIndex(); }                         // :376-379
    base::Affine3d getCentroid() { return filter.getCentroid().toTransform(); }                      // :381-384
    odometry::FootContact& getOdometry() { return odometry; }
    eslam::PoseEstimator& getFilter() { return filter; }
};

}  // namespace eslam
#endif
