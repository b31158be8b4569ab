// EmbodiedSlamFilter.hpp -- eslam::EmbodiedSlamFilter (src/EmbodiedSlamFilter.hpp:27-75): init (shared map or per-particle
// maps), update(body2odometry, BodyContactState, ltc), the mapping update with an already projected scan, processMap,
// getParticles, getBestParticleIndex, getCentroid, the PoseDistribution snapshot.  The laser / distance-image overloads
// start with envire's scan pipeline (LaserScan -> ScanMeshing -> MLSProjection, src/EmbodiedSlamFilter.cpp:137-176), which is
// out of scope: here the update takes that pipeline's OUTPUT, the scan as a grid.
#ifndef ESLAM_FACADE_EMBODIEDSLAMFILTER_HPP
#define ESLAM_FACADE_EMBODIEDSLAMFILTER_HPP
#include "PoseEstimator.hpp"

namespace eslam {

class EmbodiedSlamFilter {
    eslam::Configuration eslamConfig;
    odometry::Configuration odometryConfig;
    odometry::FootContact odometry;
    eslam::PoseEstimator filter;
    base::Affine3d udPose, mapPose;                          // poses of the last measurement / mapping update
    bool sharedMap;
    SurfaceHash hash;                                        // :70
    base::Quaterniond lastOrientation; odometry::BodyContactState lastBodyState;

public:
    EmbodiedSlamFilter(const odometry::Configuration& odometryConfig, const eslam::Configuration& eslamConfig)   // src/EmbodiedSlamFilter.cpp:13-23
        : eslamConfig(eslamConfig), odometryConfig(odometryConfig), odometry(odometryConfig), filter(odometry, eslamConfig), sharedMap(true) {}

    // init(env, pose, useSharedMap): the environment must hold a map (:77-104), particles are drawn around the pose (:118-125)
    void init(const DenseGrid* map, const base::Pose& pose, bool useSharedMap = true,
              const SurfaceHashConfig& hashConfig = SurfaceHashConfig()) {
        if (hashConfig.useHash && useSharedMap && map && map->width > 0) {                             // :93-98,:109-114
            hash.setConfiguration(hashConfig); hash.create();
            filter.setEnvironment(*map, true, patchCells);
            filter.init((int)eslamConfig.particleCount, &hash);
            udPose = base::Affine3d::Identity(); udPose.translation() = base::Vector3d(1000, 0, 0); mapPose = udPose;
            sharedMap = true;
            return;
        }
        if (!map || map->width == 0)
            throw std::runtime_error("The provided environment does not contain an mls grid.");    // :104
        const double angle = base::getYaw(pose.orientation);
        filter.init((int)eslamConfig.particleCount,
                    base::Pose2D(base::Vector2d(pose.position.x(), pose.position.y()), angle),
                    base::Pose2D(base::Vector2d(eslamConfig.initialTranslationError.x(), eslamConfig.initialTranslationError.y()),
                                 eslamConfig.initialRotationError.z()),
                    pose.position.z(), eslamConfig.initialTranslationError.z() + 1e-3);              // :118-125
        udPose = base::Affine3d::Identity(); udPose.translation() = base::Vector3d(1000, 0, 0);      // :128
        mapPose = udPose;
        sharedMap = useSharedMap;
        if (!map || map->width == 0) throw std::runtime_error("The provided environment does not contain an mls grid.");
        filter.setEnvironment(*map, useSharedMap, patchCells);                                       // :131-134 (per-particle maps: cloneMaps)
    }
    int patchCells = 16;                                     // side of a particle's own map (cells) in per-particle-map mode

    bool update(const base::Affine3d& body2odometry, const odometry::BodyContactState& bs,
                const std::vector<terrain_estimator::TerrainClassification>& ltc) {                  // :353-369
        base::Quaterniond orientation = body2odometry.rotation();
        lastBodyState = bs;
        odometry.update(body2odometry);                                                              // :357
        filter.project(bs, orientation);                                                             // :358
        if (eslamConfig.measurementThreshold.test(udPose.inverse() * body2odometry) || ltc.size() > 0) {   // :360
            filter.update(bs, orientation, ltc);
            udPose = body2odometry;
            return true;
        }
        return false;
    }

    // processMap(scanMap, match, update) (:179-232) on every particle's own map
    void processMap(const DenseGrid& scanMap, bool match, bool update) { filter.processMap(scanMap, match, update); }

    // The mapping update (:311-351) from the point where envire hands over: `scanMap` is MLSProjection's output for this scan
    // in the yaw-compensated body frame.  project, then -- once the body has moved past mappingThreshold -- processMap with
    // match = useVisualUpdate and update = !sharedMap (:337-345).
    bool update(const base::Affine3d& body2odometry, const DenseGrid& scanMap) {
        base::Quaterniond orientation = body2odometry.rotation();
        if (eslamConfig.mappingThreshold.test(mapPose.inverse() * body2odometry)) {                  // :325
            if (eslamConfig.useVisualUpdate || !sharedMap) processMap(scanMap, eslamConfig.useVisualUpdate, !sharedMap);   // :343
            mapPose = body2odometry;
            lastOrientation = orientation;
            return true;
        }
        return false;
    }

    // the logging snapshot of the filter (src/PoseParticle.hpp:88-114)
    PoseDistribution getPoseDistribution(const base::Time& time = base::Time()) {
        PoseDistribution pd;
        pd.time = time; pd.bodyState = lastBodyState;
        filter.fillPoseDistribution(pd);
        return pd;
    }

    std::vector<eslam::PoseEstimator::Particle>& getParticles() { return filter.getParticles(); }   // :371-374
    size_t getBestParticleIndex() { return filter.getBestParticleIndex(); }                         // :376-379
    base::Affine3d getCentroid() { return filter.getCentroid().toTransform(); }                      // :381-384
    odometry::FootContact& getOdometry() { return odometry; }
    eslam::PoseEstimator& getFilter() { return filter; }
};

}  // namespace eslam
#endif
