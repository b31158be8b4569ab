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
    ContactModelConfiguration config;
    eslam_gpu_ctx* gpu; bool ownsGpu;

    void recreate() {
        if (gpu && ownsGpu) eslam_gpu_destroy(gpu);
        Configuration c; c.contactModel = config;
        eslam_gpu_config g = toGpuConfig(c);
        if (eslam_gpu_create(&g, 0, NULL, &gpu) != ESLAM_OK) throw std::runtime_error("eslam_gpu_create failed: no CUDA device");
        ownsGpu = true;
    }

    void lowestPointHeuristic(bool update_probabilities) {               // src/ContactModel.cpp:48-79
        lowestPointsPerGroup.clear();
        std::vector<odometry::BodyContactPoint>& contactPoints(contactState.points);
        std::vector<std::pair<double, int> > group;
        for (size_t i = 0; i < contactPoints.size(); i++) {
            if (contactPoints[i].groupId >= 0) {
                group.push_back(std::make_pair(contactPoints[i].position.z(), (int)i));
                if (update_probabilities) contactPoints[i].contact = 0;
            }
            if (group.empty())
                lowestPointsPerGroup.push_back(contactPoints[i].position);
            else if (i + 1 == contactPoints.size() || contactPoints[i + 1].groupId != contactPoints[i].groupId) {
                std::sort(group.begin(), group.end());
                odometry::BodyContactPoint& contact = contactPoints[group[0].second];
                lowestPointsPerGroup.push_back(contact.position);
                if (update_probabilities) contact.contact = 1;
                group.clear();
            }
        }
    }

public:
    static const int GROUP_SIZE = 4;

    ContactModel() : m_zDelta(0), m_zVar(0), m_weight(0), m_poseVar(0), gpu(NULL), ownsGpu(false) { recreate(); }   // :12-14
    ~ContactModel() { if (gpu && ownsGpu) eslam_gpu_destroy(gpu); }

    void setConfiguration(const ContactModelConfiguration& c) { config = c; recreate(); }     // :16-19
    void setMap(const DenseGrid& grid) { attachGrid(gpu, grid); }

    void setContactPoints(const odometry::BodyContactState& state, const base::Quaterniond& orientation) {   // :21-41
        contactState = state;
        lowestPointsPerGroup.clear();
        base::Quaterniond yawCompensatedOrientation = base::removeYaw(orientation);
        std::vector<odometry::BodyContactPoint>& contactPoints(contactState.points);
        for (size_t i = 0; i < contactPoints.size(); i++)
            contactPoints[i].position = yawCompensatedOrientation * contactPoints[i].position;
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

    bool updateZPositionEstimate(double& z_pos, double& z_var) {                               // :319-340 (scalar Kalman step)
        const double pose_var = m_poseVar / contact_points.size();
        double delta_var = std::max(z_var - pose_var, 1e-9);
        if (std::fabs(m_zDelta / std::sqrt(delta_var)) > 1.0) return false;
        const double gain = z_var / (z_var + m_zVar);
        z_pos += gain * m_zDelta;
        const double var_gain = delta_var / (delta_var + m_zVar);
        delta_var = (1.0 - var_gain) * delta_var;
        z_var = pose_var + delta_var;
        return true;
    }
};

}  // namespace eslam
#endif
