// ContactModel.hpp -- eslam::ContactModel (src/ContactModel.hpp:32-166) over libeslam_gpu.so.
// setContactPoints / lowest-point heuristic are host bookkeeping exactly as in the reference (src/ContactModel.cpp:21-92);
// evaluatePose / evaluateWeight / updateZPositionEstimate run on the device (eslam_gpu_evaluate_poses).  The reference's
// map callback cannot run on a device: the map is a DenseGrid attached with setMap (the seam GridAccess::get implements,
// src/PoseEstimator.hpp:97-105).
#ifndef ESLAM_FACADE_CONTACTMODEL_HPP
#define ESLAM_FACADE_CONTACTMODEL_HPP
#include <algorithm>
#include <utility>
#include <vector>
#include "Configuration.hpp"
#include "ParticleFilter.hpp"

namespace eslam {

inline eslam_gpu_config toGpuConfig(const Configuration& c, const SurfaceHashConfig* hash = 0) {
    eslam_gpu_config g; eslam_gpu_default_config(&g);
    g.seed = c.seed; g.min_effective = c.minEffective; g.measurement_error = c.measurementError;
    g.discount_factor = c.discountFactor; g.spread_threshold = c.spreadThreshold;
    g.spread_translation_factor = c.spreadTranslationFactor; g.spread_rotation_factor = c.spreadRotationFactor;
    g.slip_factor = c.slipFactor; g.max_yaw_deviation = c.maxYawDeviation;
    g.min_contacts = c.contactModel.minContacts; g.contact_likelihood_correction = c.contactModel.contactLikelihoodCorrection;
    g.contact_point_radius = c.contactModel.contactPointRadius;
    g.use_slip_update = c.contactModel.useSlipUpdate; g.use_shape_update = c.contactModel.useShapeUpdate;
    g.use_hash = hash && hash->useHash;
    return g;
}

inline std::vector<eslam_gpu_contact> toAbi(const odometry::BodyContactState& state) {
    std::vector<eslam_gpu_contact> pts(state.points.size());
    for (size_t i = 0; i < pts.size(); i++) {
        const odometry::BodyContactPoint& p = state.points[i];
        pts[i].position[0] = p.position.x(); pts[i].position[1] = p.position.y(); pts[i].position[2] = p.position.z();
        pts[i].contact = p.contact; pts[i].slip = p.slip; pts[i].group_id = p.groupId; pts[i].pad = 0;
    }
    return pts;
}

inline void attachGrid(eslam_gpu_ctx* gpu, const DenseGrid& g, const double* g2l = 0) {
    gpu_check(gpu, eslam_gpu_set_grid(gpu, g.cells.data(), g.width, g.height, g.scaleX, g.scaleY, g.offsetX, g.offsetY, g2l));
}

class ContactModel {
protected:
    odometry::BodyContactState contactState;
    std::vector<base::Vector3d> lowestPointsPerGroup;
    std::vector<ContactPoint> contact_points;
    double m_zDelta, m_zVar, m_weight, m_poseVar;
    double last[5]; bool haveLast;                                    // x, y, yaw, z, measVar of the last evaluatePose
    ContactModelConfiguration config;
    eslam_gpu_ctx* gpu; bool ownsGpu;

    void recreate() {
        if (gpu && ownsGpu) eslam_gpu_destroy(gpu);
        Configuration c; c.contactModel = config;
        eslam_gpu_config g = toGpuConfig(c);
        if (eslam_gpu_create(&g, 0, NULL, &gpu) != ESLAM_OK) throw std::runtime_error("eslam_gpu_create failed: no CUDA device");
        ownsGpu = true;
    }

    // getLowestPointPerGroup / updateContactStateUsingLowestPointHeuristic (src/ContactModel.cpp:48-92): ungrouped points
    // (groupId < 0) stand for themselves; a run of consecutive points sharing a group id is represented by its lowest member
    // (first one on ties), which alone keeps contact probability 1 when the probabilities are rewritten.
    void lowestPointHeuristic(bool rewrite) {
        std::vector<odometry::BodyContactPoint>& pts = contactState.points;
        lowestPointsPerGroup.clear();
        const size_t n = pts.size();
        for (size_t first = 0; first < n;) {
            if (pts[first].groupId < 0) { lowestPointsPerGroup.push_back(pts[first].position); ++first; continue; }
            size_t last = first, low = first;                              // run [first, last] of one group id
            while (last + 1 < n && pts[last + 1].groupId == pts[first].groupId && pts[last + 1].groupId >= 0) ++last;
            for (size_t k = first + 1; k <= last; ++k)
                if (pts[k].position.z() < pts[low].position.z()) low = k;
            if (rewrite) for (size_t k = first; k <= last; ++k) pts[k].contact = (k == low) ? 1.0f : 0.0f;
            lowestPointsPerGroup.push_back(pts[low].position);
            first = last + 1;
        }
    }

public:
    static const int GROUP_SIZE = 4;

    ContactModel() : m_zDelta(0), m_zVar(0), m_weight(0), m_poseVar(0), haveLast(false), gpu(NULL), ownsGpu(false) { recreate(); }   // :12-14
    ~ContactModel() { if (gpu && ownsGpu) eslam_gpu_destroy(gpu); }

    void setConfiguration(const ContactModelConfiguration& c) { config = c; recreate(); }     // :16-19
    void setMap(const DenseGrid& grid) { attachGrid(gpu, grid); }

    // :21-41 -- the body's contact points, turned into the yaw-compensated frame (roll and pitch applied, heading left to the particles)
    void setContactPoints(const odometry::BodyContactState& state, const base::Quaterniond& orientation) {
        const base::Quaterniond rollPitch = base::removeYaw(orientation);
        contactState.points.resize(state.points.size());
        for (size_t i = 0; i < state.points.size(); ++i) {
            contactState.points[i] = state.points[i];
            contactState.points[i].position = rollPitch * state.points[i].position;
        }
        lowestPointsPerGroup.clear();
    }
    void updateContactStateUsingLowestPointHeuristic() { lowestPointHeuristic(true); }        // :81-84
    odometry::BodyContactState const& getContactState() const { return contactState; }        // :43-46
    const std::vector<base::Vector3d>& getLowestPointPerGroup() {                               // :86-92
        if (lowestPointsPerGroup.empty()) lowestPointHeuristic(false);
        return lowestPointsPerGroup;
    }
    void setTerrainClassification(const std::vector<terrain_estimator::TerrainClassification>&) {}   // inert (quirk Q5)

    // evaluatePose for a pose of the form Translation(x, y, z) * AngleAxis(yaw, Z), which is what the filter passes
    // (src/PoseEstimator.cpp:279-282) and what the reference's tests build (test/testContactModel.cpp:89,109,156).
    bool evaluatePose(const base::Affine3d& pos_and_heading, double measVar) {               // :117-224
        const double yaw = std::atan2(pos_and_heading.R[3], pos_and_heading.R[0]);
        const base::Vector3d& t = pos_and_heading.translation();
        std::vector<eslam_gpu_contact> pts = toAbi(contactState);
        uint8_t ok = 0; int32_t ncont = 0; double cp[ESLAM_MAX_CP * 5];
        double x = t.x(), y = t.y(), z = t.z();
        last[0] = x; last[1] = y; last[2] = yaw; last[3] = z; last[4] = measVar; haveLast = true;
        gpu_check(gpu, eslam_gpu_evaluate_poses(gpu, 1, &x, &y, &yaw, &z, &measVar, NULL, pts.data(), (int)pts.size(), &ok,
                                                &m_weight, &m_zDelta, &m_zVar, &m_poseVar, &ncont, NULL, NULL, NULL, cp));
        contact_points.clear();
        for (int k = 0; k < ncont && k < ESLAM_MAX_CP; k++)
            contact_points.push_back(ContactPoint(base::Vector3d(cp[5 * k], cp[5 * k + 1], cp[5 * k + 2]), cp[5 * k + 3], cp[5 * k + 4]));
        return ok != 0;
    }
    double getWeight() const { return m_weight; }
    double getZDelta() const { return m_zDelta; }
    double getZVar() const { return m_zVar; }
    std::vector<ContactPoint>& getContactPoints() { return contact_points; }

    // :319-340 -- the scalar Kalman step on the particle's height.  Evaluated on the device with the pose of the last
    // evaluatePose (the same device function the filter's weight kernels run), not restated on the host.
    bool updateZPositionEstimate(double& z_pos, double& z_var) {
        if (!haveLast) return false;
        std::vector<eslam_gpu_contact> pts = toAbi(contactState);
        double zo = z_pos, zvo = z_var; uint8_t ok = 0;
        // z_pos enters only as the state to correct: the measurement (m_zDelta, m_zVar, m_poseVar) is that of the last evaluated pose
        gpu_check(gpu, eslam_gpu_evaluate_poses(gpu, 1, &last[0], &last[1], &last[2], &last[3], &last[4], &z_var, pts.data(), (int)pts.size(),
                                                &ok, NULL, NULL, NULL, NULL, NULL, &zo, &zvo, NULL, NULL));
        if (!ok) return false;
        const bool changed = zvo != z_var || zo != last[3];
        if (z_pos == last[3]) z_pos = zo; else z_pos += zo - last[3];     // the device applied the gain to the evaluated height
        z_var = zvo;
        return changed;
    }
};

}  // namespace eslam
#endif
