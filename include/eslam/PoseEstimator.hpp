// PoseEstimator.hpp -- eslam::PoseEstimator (src/PoseEstimator.hpp:119-155) over libeslam_gpu.so: same constructor,
// init / project / update / setEnvironment / getCentroid, same exceptions.  Every step is a device call.
#ifndef ESLAM_FACADE_POSEESTIMATOR_HPP
#define ESLAM_FACADE_POSEESTIMATOR_HPP
#include "ContactModel.hpp"
#include "SurfaceHash.hpp"

namespace eslam {

class PoseEstimator : public ParticleFilter<PoseParticle> {
public:
    PoseEstimator(odometry::FootContact& odometry, const eslam::Configuration& config)      // src/PoseEstimator.cpp:13-25
        : ParticleFilter<Particle>(toGpuConfig(config)), config(config), odometry(odometry), max_weight(0), haveEnv(false), useSharedMap(true),
          hash(0), hashCalls(0) {}
    ~PoseEstimator() {}

    void init(int numParticles, const base::Pose2D& mu, const base::Pose2D& sigma, double zpos = 0, double zsigma = 0) {   // :88-102
        const double m[3] = {mu.position.x(), mu.position.y(), mu.orientation};
        const double s[3] = {sigma.position.x(), sigma.position.y(), sigma.orientation};
        gpu_check(gpu, eslam_gpu_init_gaussian(gpu, (uint64_t)numParticles, m, s, zpos, zsigma));
        deviceChanged();
    }

    // setEnvironment(env, map, useShared) (:47-62).  useShared == false is the per-particle-map mode: every particle owns
    // a map that contact heights and scans are merged into and that travels with it through resampling (cloneMaps :31-45,
    // :252-253).  On the device a particle's own map is a patchCells x patchCells overlay of the shared grid.
    // init(numParticles, hash) (:75-86): every particle a candidate pose drawn from the surface hash; project() then re-seeds the
    // lowest-weight particles from it every hash period (:239-241, sampleFromHash :130-182) and stops spreading (:224)
    void init(int numParticles, SurfaceHash* h) {
        if (!h || !h->created) throw std::runtime_error("could not sample from pose hash.");           // :84
        if (!gpuConfig.use_hash) { eslam_gpu_config c = gpuConfig; c.use_hash = 1; recreate(c); if (haveEnv) attachGrid(gpu, envGrid); }
        gpu_check(gpu, eslam_gpu_hash_create(gpu, (int)h->config.slopeBins, (int)h->config.angularSteps));
        const int st = eslam_gpu_init_from_hash(gpu, (uint64_t)numParticles);
        if (st == ESLAM_ERR_EMPTY) throw std::runtime_error("could not sample from pose hash.");         // :84
        gpu_check(gpu, st);
        hash = h; hashCalls = 0;
        deviceChanged();
    }

    void setEnvironment(const DenseGrid& map, bool useShared = true, int patchCells = 16, const double* C_global2local = 0) {
        envGrid = map;
        const int want = useShared ? 0 : patchCells;
        if (gpuConfig.patch_size != want) { eslam_gpu_config c = gpuConfig; c.patch_size = want; recreate(c); }
        useSharedMap = useShared;
        attachGrid(gpu, map, C_global2local); haveEnv = true;
    }
    bool usesSharedMap() const { return useSharedMap; }

    // the per-particle half of EmbodiedSlamFilter::processMap (src/EmbodiedSlamFilter.cpp:183-228): scan merge and match weight
    void processMap(const DenseGrid& scanMap, bool match, bool update) {
        flush();
        eslam_gpu_scan s;
        s.cells_mean_stdev = scanMap.cells.data(); s.width = scanMap.width; s.height = scanMap.height;
        s.scale_x = scanMap.scaleX; s.scale_y = scanMap.scaleY; s.offset_x = scanMap.offsetX; s.offset_y = scanMap.offsetY;
        gpu_check(gpu, eslam_gpu_process_map(gpu, &s, match ? 1 : 0, update ? 1 : 0));
        deviceChanged();
    }

    // what a PoseDistribution carries besides the particles (src/PoseParticle.hpp:88-114): the poses as a Gaussian mixture
    // (one component: weighted mean and covariance, reduced on the device)
    void fillPoseDistribution(PoseDistribution& pd) {
        flush();
        eslam_gpu_gaussian2d g;
        gpu_check(gpu, eslam_gpu_pose_moments(gpu, &g));
        pd.particles.assign(getParticles().begin(), getParticles().end());
        PoseDistribution::GMM::Parameter par;
        par.weight = 1.0; par.mean = base::Vector2d(g.mean[0], g.mean[1]);
        par.cov.m[0][0] = g.cov[0]; par.cov.m[0][1] = g.cov[1]; par.cov.m[1][0] = g.cov[2]; par.cov.m[1][1] = g.cov[3];
        pd.gmm.params.assign(1, par);
        pd.orientation = zCompensatedOrientation;
    }

    void project(const odometry::BodyContactState& state, const base::Quaterniond& orientation) {   // :184-242
        flush();
        eslam_gpu_step s;
        s.yaw = base::getYaw(orientation);                                                   // :186
        zCompensatedOrientation = base::removeYaw(orientation);                              // :187
        const base::Pose d = odometry.getPoseDelta();
        s.z_delta = (orientation * d.position).z();                                          // :188-189
        s.z_var = odometry.getPositionErrorZZ() * 2.0;                                       // :192
        s.delta_mean[0] = d.position.x(); s.delta_mean[1] = d.position.y(); s.delta_mean[2] = base::getYaw(d.orientation);
        for (int i = 0; i < 3; i++) s.delta_sigma[i] = odometry.getSampleSigma()[i];        // :198
        gpu_check(gpu, eslam_gpu_project(gpu, &s));
        deviceChanged();
        if (hash && ((hashCalls++) % hash->config.period) == 0) {                              // :239-241
            contactModel.setContactPoints(state, orientation);                               // :137
            std::vector<eslam_gpu_contact> pts = toAbi(contactModel.getContactState());
            gpu_check(gpu, eslam_gpu_sample_from_hash(gpu, hash->config.percentage, hash->config.avgFactor, pts.data(), (int)pts.size(), NULL));
        }
    }

    void update(const odometry::BodyContactState& state, const base::Quaterniond& orientation,
                const std::vector<terrain_estimator::TerrainClassification>& ltc) {          // :244-255
        flush();
        if (!haveEnv) throw std::runtime_error("No environment attached.");                 // :259-260
        contactModel.setTerrainClassification(ltc);
        contactModel.setContactPoints(state, orientation);                                   // :262
        std::vector<eslam_gpu_contact> pts = toAbi(contactModel.getContactState());
        gpu_check(gpu, eslam_gpu_update(gpu, pts.data(), (int)pts.size(), &lastResult));
        max_weight = lastResult.max_weight;
        deviceChanged();
    }

    base::Pose getCentroid() {                                                               // :354-383
        flush();
        double c[4]; gpu_check(gpu, eslam_gpu_centroid(gpu, c));
        deviceChanged();                                                                     // getCentroid re-normalises (:356)
        return base::Pose(base::Vector3d(c[0], c[1], c[3]), base::Quaterniond::fromYaw(c[2]) * zCompensatedOrientation);
    }

    const eslam_gpu_update_result& getLastUpdate() const { return lastResult; }

private:
    // the reference's ContactModel member (src/PoseEstimator.hpp:145) is only used for its host bookkeeping here
    struct HostContactModel {
        odometry::BodyContactState contactState;
        void setTerrainClassification(const std::vector<terrain_estimator::TerrainClassification>&) {}
        void setContactPoints(const odometry::BodyContactState& state, const base::Quaterniond& orientation) {
            contactState = state;
            const base::Quaterniond yc = base::removeYaw(orientation);
            for (size_t i = 0; i < contactState.points.size(); i++) contactState.points[i].position = yc * contactState.points[i].position;
        }
        const odometry::BodyContactState& getContactState() const { return contactState; }
    };

    eslam::Configuration config;
    HostContactModel contactModel;
    odometry::FootContact& odometry;
    base::Quaterniond zCompensatedOrientation;
    double max_weight;
    bool haveEnv, useSharedMap;
    DenseGrid envGrid;                                   // kept to re-attach the map when the context is recreated
    SurfaceHash* hash; size_t hashCalls;                 // :148, the call counter of :239
    eslam_gpu_update_result lastResult;
};

}  // namespace eslam
#endif
