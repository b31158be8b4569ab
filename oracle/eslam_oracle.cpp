// eslam_oracle.cpp -- CPU oracle: TEST INFRASTRUCTURE ONLY (see eslam_oracle.h for the rules and the
// parity-pinning statement).  Single-threaded restatement of the reference's particle-filter update
// path; every function names the reference lines it follows (paths relative to liyangSKD/slam-eslam).
//
// Build: oracle/Makefile (g++ -O2 -ffp-contract=off -mfma).
#include "eslam_oracle.h"
#include "../slam-eslam_b200/csrc/eslam_math.h"   // arithmetic of the SHARED policy (product header; the oracle
                                                  // depends on the product, never the other way round)
#include <algorithm>
#include <cmath>
#include <cstring>
#include <limits>
#include <string>
#include <utility>
#include <vector>

namespace em = eslam_math;

static const em::dbl2 H_TABLE[ESLAM_H_TABLE_ENTRIES] = {
#include "../slam-eslam_b200/csrc/eslam_h_table.inc"
};

namespace {

// ------------------------------------------------------------------------------------------------
// Arithmetic policies
// ------------------------------------------------------------------------------------------------
struct Literal {
    static const bool shared = false;
    static void sincos(double t, double& s, double& c) { s = std::sin(t); c = std::cos(t); }
    static double exp(double x) { return std::exp(x); }
    static double sqrt(double x) { return std::sqrt(x); }
    // two N(0,1) from two open uniforms (the reference's boost::normal_distribution is Box-Muller too,
    // but its stream is not reproducible here; SURVEY.md §8c "parity unpinned")
    static void normal_pair(double u1, double u2, double& n0, double& n1) {
        double r = std::sqrt(-2.0 * std::log(u1));
        n0 = r * std::cos(2.0 * M_PI * u2);
        n1 = r * std::sin(2.0 * M_PI * u2);
    }
    // boost::math::normal n(0, sigma*corr); pdf(n,z)/cdf(n,z)   src/ContactModel.cpp:104-115
    static double likelihood_ratio(double z, double zvar, double corr) {
        const double sigma = std::sqrt(zvar);                 // :170 sqrt(zvar)
        const double sd = sigma * corr;                       // :109
        double exponent = z - 0.0;                            // boost normal pdf
        exponent *= -exponent;
        exponent /= 2 * sd * sd;
        double pdf = std::exp(exponent);
        pdf /= sd * std::sqrt(2 * M_PI);
        const double diff = (z - 0.0) / (sd * std::sqrt(2.0)); // boost normal cdf
        const double cdf = std::erfc(-diff) / 2;
        return pdf / cdf;                                      // :111-112
    }
    static double root_n(double w, size_t n) { return std::pow(w, 1.0 / n); }      // src/PoseEstimator.cpp:309
    static double pow_u64(double x, size_t e) { return std::pow(x, (double)e); }   // :336 (size_t exponent)
    static double sum(const std::vector<double>& v) {                              // left-to-right, src/ParticleFilter.hpp:36-37
        double s = 0;
        for (size_t n = 0; n < v.size(); s += v[n++]);
        return s;
    }
};

struct Shared {
    static const bool shared = true;
    static void sincos(double t, double& s, double& c) { em::sincos_(t, s, c); }
    static double exp(double x) { return em::exp_(x); }
    static double sqrt(double x) { return em::sqrt_(x); }
    static void normal_pair(double u1, double u2, double& n0, double& n1) { em::box_muller(u1, u2, n0, n1); }
    static double likelihood_ratio(double z, double zvar, double corr) {
        return em::likelihood_ratio(z, zvar, 1.0 / corr, H_TABLE);
    }
    static double root_n(double w, size_t n) { return em::root_n(w, n); }
    static double pow_u64(double x, size_t e) { return em::pow_u64(x, e); }
    // pairwise sum over the index-aligned binary tree (zero padded): the order the GPU reduces in
    static double sum(const std::vector<double>& v) {
        std::vector<double> a(v);
        size_t n = a.size();
        if (n == 0) return 0.0;
        while (n > 1) {
            size_t h = (n + 1) / 2;
            for (size_t i = 0; i < h; ++i) a[i] = (2 * i + 1 < n) ? a[2 * i] + a[2 * i + 1] : a[2 * i];
            n = h;
        }
        return a[0];
    }
};

// ------------------------------------------------------------------------------------------------
// Types (src/PoseParticle.hpp:20-86)
// ------------------------------------------------------------------------------------------------
struct Vec3 { double x, y, z; };

struct ContactPoint {          // src/PoseParticle.hpp:20-43
    Vec3 point; double zdiff, zvar, prob;
    int cell_ix, cell_iy;      // map cell of `point` (used by the per-particle patch write)
};

struct PoseParticle {          // src/PoseParticle.hpp:52-86 (hot fields)
    double px, py;             // position
    double orientation;
    double zPos, zSigma;
    double mprob;
    bool floating;
    size_t n_cpoints;          // cpoints.size() after the last update
    double weight;
};

struct SurfacePatch { double mean, stdev; };   // envire::MLSGrid::SurfacePatch (mean/stdev are what the path reads)

struct Grid {                  // dense single-patch-per-cell stand-in for envire::MLSMap (SURVEY.md §8c definition)
    std::vector<float> cells;  // {mean, stdev} per cell, NaN mean = empty
    int w = 0, h = 0;
    double sx = 1, sy = 1, ox = 0, oy = 0;
    double g2l[12] = {1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0};
    bool identity = true;
    bool attached = false;
};

struct FakeMap {               // FakeMLSAccess, test/testContactModel.cpp:8-38
    bool active = false;
    double z[4], stddev[4]; bool has_res = false; bool res[4];
    std::vector<Vec3> points;
};

// Philox slots: all slots are defined for every (step, particle), consumed conditionally
struct Noise {
    uint32_t k0, k1;
    explicit Noise(uint64_t seed) : k0((uint32_t)seed), k1((uint32_t)(seed >> 32)) {}
    em::Philox4 draw(uint64_t index, uint32_t step, uint32_t slot, uint32_t stream) const {
        return em::philox4x32_10((uint32_t)index, step, slot, stream ^ ((uint32_t)(index >> 32) << 8), k0, k1);
    }
};

}  // namespace

extern "C" int eslam_oracle_lowest_points(const eslam_gpu_contact* pts, int n_pts, int update_probabilities, double* lowest_xyz, float* contact_out);

// ------------------------------------------------------------------------------------------------
// The filter
// ------------------------------------------------------------------------------------------------
struct eslam_oracle_ctx {
    eslam_gpu_config cfg;
    int policy;
    eslam_gpu_shard shard;
    Noise noise;
    Grid grid;
    FakeMap fake;
    std::vector<PoseParticle> xi_k;              // ParticleFilter::xi_k  src/ParticleFilter.hpp:176
    std::vector<uint32_t> ancestors;
    std::vector<double> cdf;
    double max_weight = 0;                       // src/PoseEstimator.hpp:154
    uint32_t project_step = 0, resample_step = 0;
    int migrate = 1;                             // LOCAL_MIGRATE: 1 = the next resampling is exact-global (set by shard_decision)
    // contact model state (src/ContactModel.hpp:35-47)
    std::vector<eslam_gpu_contact> contactPoints;
    std::vector<ContactPoint> contact_points;
    std::vector<int32_t> cell_trace;             // parity tap: cell index per evaluated contact
    double m_zDelta = 0, m_zVar = 0, m_weight = 0, m_poseVar = 0;
    // shared-policy accumulators for evaluateWeight
    double sh_d1 = 0, sh_d2 = 0, sh_A = 0;
    // per-particle patches
    std::vector<float> patches; std::vector<int32_t> patch_origin;
    // SurfaceHash (src/SurfaceHash.hpp:113-232)
    struct HashPose { uint16_t m, n; uint8_t a; uint16_t bucket; double z; };
    std::vector<HashPose> hash_poses; std::vector<uint32_t> hash_by_bucket, hash_offs; int hash_bins = 0, hash_steps = 0; uint32_t hash_step = 0;

    eslam_oracle_ctx(const eslam_gpu_config& c, int pol) : cfg(c), policy(pol), noise(c.seed) {
        shard.rank = 0; shard.world = 1; shard.global_count = 0;
    }

    // --- map seam: GridAccess::get -> MLSMap::getPatch(C_global2local*pos, patch, 3.0)  src/PoseEstimator.hpp:97-105
    template <class A>
    bool map_get(const Vec3& pos, SurfacePatch& patch, double measVar, int32_t* cell_out, long particle, int* cix, int* ciy) {
        if (cell_out) *cell_out = -1;
        *cix = *ciy = -1;
        if (fake.active) {                                      // test/testContactModel.cpp:18-37
            fake.points.push_back(pos);
            int index = 0;
            if (pos.y > 0) index += 2;
            if (pos.x > 0) index++;
            patch.mean = fake.z[index];
            patch.stdev = fake.stddev[index];
            if (cell_out) *cell_out = index;
            if (fake.has_res) return fake.res[index];
            return true;
        }
        if (!grid.attached) return false;
        double gx = pos.x, gy = pos.y;
        if (!grid.identity) {
            const double* g = grid.g2l;
            if (A::shared) {
                gx = em::fma_(g[0], pos.x, em::fma_(g[1], pos.y, em::fma_(g[2], pos.z, g[3])));
                gy = em::fma_(g[4], pos.x, em::fma_(g[5], pos.y, em::fma_(g[6], pos.z, g[7])));
            } else {
                gx = (g[0] * pos.x + g[1] * pos.y + g[2] * pos.z) + g[3];
                gy = (g[4] * pos.x + g[5] * pos.y + g[6] * pos.z) + g[7];
            }
        }
        // envire GridBase::toGrid: floor((x - offset)/scale), in-range test
        int ix, iy;
        if (!em::cell_index_exact(gx, grid.ox, grid.sx, grid.w, ix)) return false;
        if (!em::cell_index_exact(gy, grid.oy, grid.sy, grid.h, iy)) return false;
        const size_t cell = (size_t)iy * grid.w + ix;
        if (cell_out) *cell_out = (int32_t)cell;
        *cix = ix; *ciy = iy;
        float mean = grid.cells[2 * cell], stdev = grid.cells[2 * cell + 1];
        // per-particle map first (MLSMap::getPatch walks the newest grid first): the particle-local patch overlays
        // the shared grid; its cell is used when it holds a surface that passes the 3-sigma gate
        if (particle >= 0 && cfg.patch_size > 0 && !patches.empty()) {
            const int P = cfg.patch_size;
            const int32_t* org = &patch_origin[2 * particle];
            if (org[0] != INT32_MIN) {
                const int lx = ix - org[0], ly = iy - org[1];
                if (lx >= 0 && lx < P && ly >= 0 && ly < P) {
                    const float* pc = &patches[((size_t)particle * P * P + (size_t)ly * P + lx) * 2];
                    if (!std::isnan(pc[0]) && !std::isnan(pc[1])) {
                        bool pass;
                        if (A::shared) {
                            const double d = patch.mean - (double)pc[0];
                            pass = d * d < 9.0 * ((double)pc[1] * (double)pc[1] + measVar);
                        } else {
                            const double interval = std::sqrt(patch.stdev * patch.stdev + (double)pc[1] * (double)pc[1]) * 3.0;
                            pass = std::fabs(patch.mean - (double)pc[0]) < interval;
                        }
                        if (pass) { patch.mean = pc[0]; patch.stdev = pc[1]; return true; }
                    }
                }
            }
        }
        if (std::isnan(mean) || std::isnan(stdev)) return false;   // empty cell
        // MLSGrid::get(pos, patch, sigma_threshold = 3.0): |mean - query| < 3 sqrt(stdev_q^2 + stdev_p^2)
        if (A::shared) {
            const double d = patch.mean - (double)mean;
            const double zvar = (double)stdev * (double)stdev + measVar;
            if (!(d * d < 9.0 * zvar)) return false;
        } else {
            const double interval = std::sqrt(patch.stdev * patch.stdev + (double)stdev * (double)stdev) * 3.0;
            if (!(std::fabs(patch.mean - (double)mean) < interval)) return false;
        }
        patch.mean = mean; patch.stdev = stdev;
        return true;
    }

    // --- ContactModel::evaluatePose  src/ContactModel.cpp:117-224
    template <class A>
    bool evaluatePose(double tx, double ty, double tz, double theta, double measVar, int32_t* cells, long particle = -1) {
        if (measVar == 0) throw std::string("using a zero measurement variance leads to singularities");  // :122-123
        contact_points.clear();                                 // :126
        double s, c; A::sincos(theta, s, c);
        // Eigen::AngleAxisd(theta, UnitZ).toRotationMatrix(): R = [c -s 0; s c 0; 0 0 (1-c)+c]
        const double r22 = (1.0 - c) * 1.0 * 1.0 + c;
        ContactPoint p = ContactPoint();
        bool valid = false;                                     // :134
        bool group_valid = true;                                // :135
        const double contact_threshold = 0.2;                   // :136
        double contact_ratio = 0;
        double pose_var_avg = 0;
        m_poseVar = 0;                                          // :139
        sh_d1 = sh_d2 = sh_A = 0;
        const double radius = cfg.contact_point_radius;
        for (size_t i = 0; i < contactPoints.size(); i++) {     // :140
            int groupId = contactPoints[i].group_id;
            const double* cp = contactPoints[i].position;
            Vec3 w;                                             // :147 pos_and_heading * contact_point - (0,0,radius)
            if (A::shared) {
                w.x = em::fma_(c, cp[0], em::fma_(-s, cp[1], tx));
                w.y = em::fma_(s, cp[0], em::fma_(c, cp[1], ty));
                w.z = tz + (cp[2] - radius);
            } else {
                w.x = (c * cp[0] + (-s) * cp[1] + 0.0 * cp[2]) + tx;
                w.y = (s * cp[0] + c * cp[1] + 0.0 * cp[2]) + ty;
                w.z = ((0.0 * cp[0] + 0.0 * cp[1] + r22 * cp[2]) + tz) - radius;
            }
            SurfacePatch patch = {w.z, std::sqrt(measVar)};     // :151
            const float contactProbability = contactPoints[i].contact;   // :153
            if (cells) cells[i] = -1;
            if (group_valid && !(contactProbability < contact_threshold)) {   // :154
                int cix, ciy;
                if (map_get<A>(w, patch, measVar, cells ? &cells[i] : nullptr, particle, &cix, &ciy)) {   // :156
                    const double zdiff = w.z - patch.mean;      // :160
                    const double pose_var = patch.stdev * patch.stdev;           // :166
                    const double zvar = patch.stdev * patch.stdev + measVar;      // :167
                    const Vec3 surface_point = {w.x, w.y, patch.mean};           // :168
                    const double ratio = A::likelihood_ratio(zdiff, zvar, cfg.contact_likelihood_correction);  // :170
                    if (!valid) {                               // :172-177
                        p.point = surface_point; p.prob = 1.0; p.cell_ix = cix; p.cell_iy = ciy;
                        p.zdiff = zdiff * ratio; p.zvar = zvar * ratio;
                        contact_ratio = ratio;
                        pose_var_avg = pose_var * ratio;
                    } else if (A::shared) {
                        p.zdiff = em::fma_(zdiff, ratio, p.zdiff); p.zvar = em::fma_(zvar, ratio, p.zvar);
                        contact_ratio += ratio;
                        pose_var_avg = em::fma_(pose_var, ratio, pose_var_avg);
                    } else {                                    // :178-184
                        p.zdiff += zdiff * ratio; p.zvar += zvar * ratio;
                        contact_ratio += ratio;
                        pose_var_avg += pose_var * ratio;
                    }
                    valid = true;                               // :185
                } else
                    group_valid = false;                        // :189-190
            }
            if (valid && (groupId == -1 || i + 1 == contactPoints.size()
                          || groupId != contactPoints[i + 1].group_id)) {          // :194-197
                if (group_valid && contact_ratio > 1e-9) {      // :199
                    if (A::shared) {                            // one reciprocal instead of two divisions
                        const double ir = 1.0 / contact_ratio;
                        p.zdiff *= ir; p.zvar *= ir;
                    } else {
                        p.zdiff /= contact_ratio;               // :201-203
                        p.zvar /= contact_ratio;
                    }
                    // SHARED: sum(stdev^2 r)/sum(r) == sum(zvar r)/sum(r) - measVar (measVar is constant over the
                    // particle), so the fourth running sum is not carried
                    if (A::shared) m_poseVar += p.zvar - measVar;
                    else m_poseVar += pose_var_avg / contact_ratio;
                    contact_points.push_back(p);                // :205
                    if (A::shared) {                            // running sums for evaluateWeight
                        const double iv = 1.0 / p.zvar;
                        const double e = p.zdiff * iv;
                        sh_d1 += e; sh_d2 += iv; sh_A = em::fma_(p.zdiff, e, sh_A);
                    }
                    // :207-208 p.prob *= matchTerrain(...) happens after the push_back: inert (quirk Q5)
                }
                group_valid = true;                             // :210-213 (only when `valid`: quirk Q4)
                valid = false;
                pose_var_avg = 0;
                contact_ratio = 0;
            }
        }
        if (contact_points.size() >= cfg.min_contacts) {        // :217
            evaluateWeight<A>(measVar);
            return true;
        }
        return false;
    }

    // --- ContactModel::evaluateWeight  src/ContactModel.cpp:262-317
    template <class A>
    void evaluateWeight(double /*measVar*/) {
        if (A::shared) {
            // same quantities from three running sums: sum((zd-delta)^2/zv) = A - d1*delta
            const double delta = sh_d1 / sh_d2;
            double pz = 1.0;
            if (cfg.use_shape_update) {
                double q = em::fma_(-sh_d1, delta, sh_A);
                if (!(q > 0.0)) q = 0.0;
                pz = em::exp_(-0.5 * q);
            }
            m_weight = pz; m_zDelta = -delta; m_zVar = 1.0 / sh_d2;
            return;
        }
        double d1 = 0, d2 = 0;                                  // :266
        for (size_t i = 0; i < contact_points.size(); i++) {    // :267-273
            ContactPoint& p(contact_points[i]);
            d1 += p.zdiff / p.zvar;
            d2 += 1.0 / p.zvar;
        }
        const double delta = d1 / d2;                           // :274
        double pz = 1.0;                                        // :278
        for (size_t i = 0; i < contact_points.size(); i++) {    // :279-291
            ContactPoint& p(contact_points[i]);
            const double odiff = (p.zdiff - delta) / std::sqrt(p.zvar);
            const double zk = std::exp(-(odiff * odiff) / (2.0));
            if (cfg.use_shape_update) pz *= zk;
            if (cfg.use_slip_update) pz *= p.prob;              // prob is always 1 (Q5)
        }
        m_weight = pz;                                          // :299-301
        m_zDelta = -delta;
        m_zVar = 1.0 / d2;
    }

    // --- per-particle map write (north_star's contact-height variant of EmbodiedSlamFilter::processMap,
    //     src/EmbodiedSlamFilter.cpp:209-227; definition in DESIGN.md "particle-local patches"): every fused contact of
    //     the particle's last evaluation writes {height = point.z + zdiff, stdev = sqrt(zvar)} into the particle's own
    //     P x P patch, Kalman-merged (float arithmetic) with what the cell already holds.
    void patch_write(size_t i, double px, double py) {
        const int P = cfg.patch_size;
        int32_t* org = &patch_origin[2 * i];
        if (org[0] == INT32_MIN) {                              // anchor the patch on the particle's own cell
            int ix, iy;
            double gx = px, gy = py;
            if (!grid.identity) { const double* g = grid.g2l; gx = em::fma_(g[0], px, em::fma_(g[1], py, g[3])); gy = em::fma_(g[4], px, em::fma_(g[5], py, g[7])); }
            if (!em::cell_index_exact(gx, grid.ox, grid.sx, grid.w, ix) || !em::cell_index_exact(gy, grid.oy, grid.sy, grid.h, iy)) return;
            org[0] = ix - P / 2; org[1] = iy - P / 2;
        }
        const size_t ncp = std::min<size_t>(contact_points.size(), ESLAM_MAX_CP);
        for (size_t k = 0; k < ncp; k++) {
            const ContactPoint& q = contact_points[k];
            const int lx = q.cell_ix - org[0], ly = q.cell_iy - org[1];
            if (lx < 0 || lx >= P || ly < 0 || ly >= P) continue;
            float* cell = &patches[(i * P * P + (size_t)ly * P + lx) * 2];
            const float zm = (float)(q.point.z + q.zdiff);      // contact height
            const float sd = (float)std::sqrt(q.zvar);
            if (std::isnan(cell[0]) || std::isnan(cell[1])) { cell[0] = zm; cell[1] = sd; }
            else {                                              // Kalman merge of two height estimates
                const float v0 = cell[1] * cell[1], v1 = sd * sd;
                const float kg = v0 / (v0 + v1);
                cell[0] = cell[0] + kg * (zm - cell[0]);
                cell[1] = std::sqrt((1.0f - kg) * v0);
            }
        }
    }

    // --- EmbodiedSlamFilter::processMap(scanMap, match, update)  src/EmbodiedSlamFilter.cpp:179-232, per-particle-map mode.
    //     The scan (MLSProjection's output in the yaw-compensated body frame, :137-176) arrives as a dense grid; per particle:
    //     C_s2p = Translation(x, y, 0) * AngleAxis(theta, Z) (:190-193,:209), offsetPatch(zPos, zSigma) (:212),
    //     match -> p.weight *= pow(weight, 0.1) (:214-221), merge (:222-227).  envire's MLSGrid::match / merge are not
    //     vendored (parity unpinned); THIS is their definition for the dense single-surface grid:
    //       merge : every scan cell with a surface, in row-major order: its centre goes through C_s2p into the particle's map
    //               cell; {mean + zPos, sqrt(stdev^2 + zSigma^2)} is Kalman-merged (float) into that cell of the particle's
    //               patch (anchored on the particle's own cell at its first write; cells outside the patch are dropped);
    //       match : `sampling` scan cells with a surface, evenly strided through the row-major list; each is looked up in
    //               the particle's map as it was before this scan's merge (patch first, then the shared grid, 3-sigma gate
    //               on the offset height); weight = mean over the hits of exp(-d^2 / (2 sigma^2)), 1 without any hit,
    //               returned as float like upstream.
    struct ScanCell { double x, y; float mean, stdev; };
    std::vector<ScanCell> scan_cells;
    template <class A>
    void patch_merge(size_t i, int ix, int iy, float zm, float sd, double px, double py) {
        const int P = cfg.patch_size;
        int32_t* org = &patch_origin[2 * i];
        if (org[0] == INT32_MIN) {                              // anchor the patch on the particle's own cell
            int cx, cy;
            double gx = px, gy = py;
            if (!grid.identity) { const double* g = grid.g2l; gx = em::fma_(g[0], px, em::fma_(g[1], py, g[3])); gy = em::fma_(g[4], px, em::fma_(g[5], py, g[7])); }
            if (!em::cell_index_exact(gx, grid.ox, grid.sx, grid.w, cx) || !em::cell_index_exact(gy, grid.oy, grid.sy, grid.h, cy)) return;
            org[0] = cx - P / 2; org[1] = cy - P / 2;
        }
        const int lx = ix - org[0], ly = iy - org[1];
        if (lx < 0 || lx >= P || ly < 0 || ly >= P) return;
        float* cell = &patches[(i * P * P + (size_t)ly * P + lx) * 2];
        if (std::isnan(cell[0]) || std::isnan(cell[1])) { cell[0] = zm; cell[1] = sd; }
        else {                                                  // Kalman merge of two height estimates
            const float v0 = cell[1] * cell[1], v1 = sd * sd;
            const float kg = v0 / (v0 + v1);
            cell[0] = cell[0] + kg * (zm - cell[0]);
            cell[1] = std::sqrt((1.0f - kg) * v0);
        }
    }
    template <class A>
    int process_map(const eslam_gpu_scan* scan, int match, int update) {
        if (cfg.patch_size <= 0 || patches.empty()) return ESLAM_ERR_INVALID;
        if (!grid.attached) return ESLAM_ERR_NO_ENVIRONMENT;
        scan_cells.clear();
        for (int j = 0; j < scan->height; j++)
            for (int k = 0; k < scan->width; k++) {
                const float m = scan->cells_mean_stdev[2 * ((size_t)j * scan->width + k)], sd = scan->cells_mean_stdev[2 * ((size_t)j * scan->width + k) + 1];
                if (std::isnan(m) || std::isnan(sd)) continue;
                ScanCell c; c.x = scan->offset_x + (k + 0.5) * scan->scale_x; c.y = scan->offset_y + (j + 0.5) * scan->scale_y; c.mean = m; c.stdev = sd;
                scan_cells.push_back(c);
            }
        const size_t M = scan_cells.size();
        const size_t sampling = 10; const double sigma = 0.2, visualWeighting = 0.1;   // src/EmbodiedSlamFilter.cpp:216-220
        const size_t n_samples = std::min(sampling, M);
        for (size_t i = 0; i < xi_k.size(); i++) {
            PoseParticle& p(xi_k[i]);
            double s, c; A::sincos(p.orientation, s, c);        // :190-193
            const double ox = p.px, oy = p.py;
            if (match && n_samples > 0) {                       // :214-221
                double wsum = 0; size_t hits = 0;
                for (size_t j = 0; j < n_samples; j++) {
                    const ScanCell& sc_ = scan_cells[j * M / n_samples];
                    Vec3 w; w.x = em::fma_(c, sc_.x, em::fma_(-s, sc_.y, ox)); w.y = em::fma_(s, sc_.x, em::fma_(c, sc_.y, oy));
                    w.z = (double)sc_.mean + p.zPos;
                    const double qvar = (double)sc_.stdev * (double)sc_.stdev + p.zSigma * p.zSigma;
                    SurfacePatch patch = {w.z, A::sqrt(qvar)};
                    int cix, ciy;
                    if (map_get<A>(w, patch, qvar, nullptr, (long)i, &cix, &ciy)) {
                        const double d = w.z - patch.mean;
                        wsum += A::exp(-(d * d) / (2.0 * sigma * sigma));
                        hits++;
                    }
                }
                const float weight = hits ? (float)(wsum / (double)hits) : 1.0f;
                p.weight *= A::shared ? em::exp_(visualWeighting * em::log_((double)weight)) : std::pow((double)weight, visualWeighting);   // :220
            }
            if (update) {                                       // :222-227
                for (size_t j = 0; j < M; j++) {
                    const ScanCell& sc_ = scan_cells[j];
                    const double wx = em::fma_(c, sc_.x, em::fma_(-s, sc_.y, ox)), wy = em::fma_(s, sc_.x, em::fma_(c, sc_.y, oy));
                    double gx = wx, gy = wy;
                    if (!grid.identity) { const double* g = grid.g2l; gx = em::fma_(g[0], wx, em::fma_(g[1], wy, g[3])); gy = em::fma_(g[4], wx, em::fma_(g[5], wy, g[7])); }
                    int ix, iy;
                    if (!em::cell_index_exact(gx, grid.ox, grid.sx, grid.w, ix) || !em::cell_index_exact(gy, grid.oy, grid.sy, grid.h, iy)) continue;
                    const float zm = (float)((double)sc_.mean + p.zPos);
                    const float sd = (float)A::sqrt((double)sc_.stdev * (double)sc_.stdev + p.zSigma * p.zSigma);
                    patch_merge<A>(i, ix, iy, zm, sd, ox, oy);
                }
            }
        }
        return ESLAM_OK;
    }

    // --- SurfaceHash::create (src/SurfaceHash.hpp:155-231): for every heading a of `steps` and every grid cell the least-squares
    //     plane under a 4-point footprint (base 0.5 m) -> (slope_x, slope_y) -> bucket; candidate pose = cell centre, heading
    //     a 2 pi / steps, z = mean height + 0.18.  Kept as in the reference: the footprint is rotated BEFORE it is used, i.e. by
    //     (a + 1) steps while the stored heading is a steps (:177-181); the plane is fitted over the UNROTATED footprint
    //     coordinates (:199).  Defined here (the reference reads uninitialised indices there): a footprint point outside the
    //     grid is skipped.  Order of the candidates: heading, then grid row, then column.
    int hash_create(int bins, int steps) {
        if (!grid.attached || !grid.identity || bins <= 0 || bins > 255 || steps <= 0 || steps > 255 || grid.w > 65535 || grid.h > 65535) return ESLAM_ERR_INVALID;
        hash_bins = bins; hash_steps = steps; hash_poses.clear();
        for (int a = 0; a < steps; a++) {
            const double rot = (a + 1) * 2.0 * M_PI / steps;
            double s, c; em::sincos_(rot, s, c);
            double rx[4], ry[4], fx[4], fy[4];
            for (int i = 0; i < 4; i++) { em::hash_footprint(i, fx[i], fy[i]); rx[i] = c * fx[i] - s * fy[i]; ry[i] = s * fx[i] + c * fy[i]; }
            for (int n = 0; n < grid.h; n++)
                for (int m = 0; m < grid.w; m++) {
                    const double x = grid.ox + (m + 0.5) * grid.sx, y = grid.oy + (n + 0.5) * grid.sy;   // fromGrid :189
                    double gx[4], gy[4], gz[4], mean_z = 0; int cnt = 0;
                    for (int i = 0; i < 4; i++) {                                                          // :193-206
                        int ix, iy;
                        if (!em::cell_index_exact(x + rx[i], grid.ox, grid.sx, grid.w, ix) || !em::cell_index_exact(y + ry[i], grid.oy, grid.sy, grid.h, iy)) continue;
                        const float mean = grid.cells[2 * ((size_t)iy * grid.w + ix)], sd = grid.cells[2 * ((size_t)iy * grid.w + ix) + 1];
                        if (std::isnan(mean) || std::isnan(sd)) continue;
                        gx[cnt] = fx[i]; gy[cnt] = fy[i]; gz[cnt] = (double)mean; mean_z += (double)mean; cnt++;
                    }
                    if (cnt < 3) continue;                                                                 // :209
                    mean_z /= cnt;
                    double slx, sly;
                    if (!em::surface_slopes(gx, gy, gz, cnt, slx, sly)) continue;                          // :211-212
                    HashPose hp; hp.m = (uint16_t)m; hp.n = (uint16_t)n; hp.a = (uint8_t)a;
                    hp.bucket = (uint16_t)(em::slope_bucket(slx, bins) * bins + em::slope_bucket(sly, bins));   // :218
                    hp.z = mean_z + 0.18;                                                                  // :216
                    hash_poses.push_back(hp);
                }
        }
        // the buckets: candidate indices grouped by bucket, creation order inside a bucket
        hash_offs.assign((size_t)bins * bins + 1, 0);
        for (size_t i = 0; i < hash_poses.size(); i++) hash_offs[hash_poses[i].bucket + 1]++;
        for (size_t b = 0; b < (size_t)bins * bins; b++) hash_offs[b + 1] += hash_offs[b];
        hash_by_bucket.resize(hash_poses.size());
        std::vector<uint32_t> fill(hash_offs.begin(), hash_offs.end() - 1);
        for (size_t i = 0; i < hash_poses.size(); i++) hash_by_bucket[fill[hash_poses[i].bucket]++] = (uint32_t)i;
        return ESLAM_OK;
    }
    void hash_pose(const HashPose& hp, double& x, double& y, double& theta) const {
        x = grid.ox + (hp.m + 0.5) * grid.sx; y = grid.oy + (hp.n + 0.5) * grid.sy;
        theta = hp.a * 2.0 * M_PI / hash_steps;                                                             // :175
    }
    // --- PoseEstimator::init(numParticles, hash) (src/PoseEstimator.cpp:75-86): every particle a uniformly drawn candidate
    //     (SurfaceHash::sample :129-133 uses libc rand(); here the Philox stream STREAM_HASH, counter = particle index)
    int init_from_hash(uint64_t n) {
        if (hash_poses.empty()) return ESLAM_ERR_EMPTY;                                                     // "could not sample from pose hash." :84
        PoseParticle p; std::memset(&p, 0, sizeof(p)); p.floating = true;
        xi_k.assign(n, p);
        for (uint64_t i = 0; i < n; i++) {
            const em::Philox4 r = noise.draw(i, 0xFFFFFFFFu, 0, em::STREAM_HASH);
            const size_t idx = std::min((size_t)(em::u01(r.v[0], r.v[1]) * (double)hash_poses.size()), hash_poses.size() - 1);
            double th; hash_pose(hash_poses[idx], xi_k[i].px, xi_k[i].py, th);
            xi_k[i].orientation = th; xi_k[i].zPos = hash_poses[idx].z;                                     // PoseParticle(position, orientation, zpos): zSigma 0, weight 0
        }
        if (cfg.patch_size > 0) {
            const size_t cells = (size_t)cfg.patch_size * cfg.patch_size * 2;
            patches.assign(n * cells, std::numeric_limits<float>::quiet_NaN()); patch_origin.assign(2 * n, INT32_MIN);
        }
        max_weight = 0; project_step = 0; resample_step = 0; hash_step = 0;
        return ESLAM_OK;
    }
    // --- PoseEstimator::sampleFromHash (src/PoseEstimator.cpp:130-182): the lowest-weight particles are replaced by candidates
    //     whose surface slopes match what the body's lowest contact points see
    template <class A>
    int sample_from_hash(double replace_percentage, double avg_factor, const eslam_gpu_contact* pts, int n_pts, uint64_t* replaced) {
        if (replaced) *replaced = 0;
        if (hash_poses.empty() || xi_k.empty()) return ESLAM_ERR_EMPTY;
        std::vector<double> low((size_t)n_pts * 3);
        const int nl = eslam_oracle_lowest_points(pts, n_pts, 0, low.data(), nullptr);                     // :139-140
        std::vector<double> lx(nl), ly(nl), lz(nl);
        for (int i = 0; i < nl; i++) { lx[i] = low[3 * i]; ly[i] = low[3 * i + 1]; lz[i] = low[3 * i + 2]; }
        double slx = 0, sly = 0;
        if (nl < 3 || !em::surface_slopes(lx.data(), ly.data(), lz.data(), nl, slx, sly)) { hash_step++; return ESLAM_OK; }   // :143-144
        const int b = em::slope_bucket(slx, hash_bins) * hash_bins + em::slope_bucket(sly, hash_bins);
        const size_t in_bucket = hash_offs[b + 1] - hash_offs[b];
        const double relevance = 1.0 - 1.0 * in_bucket / hash_poses.size();                                 // getRelevance :139-144
        const double relevance_factor = relevance * relevance * relevance;                                  // pow(., 3) :156
        size_t replace_count = (size_t)(xi_k.size() * replace_percentage * relevance_factor);               // :157
        if (relevance_factor < 0.8) replace_count = 0;                                                      // :158-159
        std::vector<double> w(xi_k.size());
        for (size_t i = 0; i < xi_k.size(); i++) w[i] = xi_k[i].weight;
        const double weight = (A::sum(w) / xi_k.size()) * avg_factor * relevance_factor;                   // :162 getWeightAvg()
        typedef std::pair<float, unsigned long> weight_index;                                               // :147-153
        std::vector<weight_index> widxs(xi_k.size());
        for (size_t i = 0; i < xi_k.size(); i++) widxs[i] = weight_index((float)xi_k[i].weight, i);
        std::stable_sort(widxs.begin(), widxs.end(), [](const weight_index& l, const weight_index& r) { return l.first < r.first; });
        uint64_t done = 0;
        for (size_t i = 0; i < replace_count && in_bucket > 0; i++) {                                       // :164-180 (sample(params) == NULL: nothing)
            const em::Philox4 r = noise.draw(i, hash_step, 1, em::STREAM_HASH);
            const size_t k = std::min((size_t)(em::u01(r.v[0], r.v[1]) * (double)in_bucket), in_bucket - 1);
            const HashPose& hp = hash_poses[hash_by_bucket[hash_offs[b] + k]];
            PoseParticle& pose = xi_k[widxs[i].second];
            double th; hash_pose(hp, pose.px, pose.py, th);
            pose.orientation = th; pose.zPos = hp.z; pose.zSigma = 0.5; pose.floating = true; pose.weight = weight;   // :170-175
            done++;
        }
        hash_step++;
        if (replaced) *replaced = done;
        return ESLAM_OK;
    }

    // --- weighted mean and covariance of the particles' (x, y): the single-component mixture a PoseDistribution carries
    //     (src/PoseParticle.hpp:88-114); sums in the policy's order
    template <class A>
    void pose_moments(eslam_gpu_gaussian2d* out) {
        const size_t n = xi_k.size();
        std::vector<double> w(n), wx(n), wy(n);
        for (size_t i = 0; i < n; i++) { w[i] = xi_k[i].weight; wx[i] = xi_k[i].px * w[i]; wy[i] = xi_k[i].py * w[i]; }
        const double W = A::sum(w), mx = A::sum(wx) / W, my = A::sum(wy) / W;
        std::vector<double> cxx(n), cxy(n), cyy(n);
        for (size_t i = 0; i < n; i++) {
            const double dx = xi_k[i].px - mx, dy = xi_k[i].py - my;
            cxx[i] = w[i] * dx * dx; cxy[i] = w[i] * dx * dy; cyy[i] = w[i] * dy * dy;
        }
        out->weight = W; out->mean[0] = mx; out->mean[1] = my;
        out->cov[0] = A::sum(cxx) / W; out->cov[1] = out->cov[2] = A::sum(cxy) / W; out->cov[3] = A::sum(cyy) / W;
    }

    // --- ContactModel::updateZPositionEstimate  src/ContactModel.cpp:319-340
    bool updateZPositionEstimate(double& z_pos, double& z_var) {
        const double pose_var = m_poseVar / contact_points.size();      // :321
        double delta_var = std::max(z_var - pose_var, 1e-9);            // :322
        const double z_delta = m_zDelta;                                // :327
        if (std::fabs(z_delta / std::sqrt(delta_var)) > 1.0)            // :328 (1 sigma, quirk Q6)
            return false;
        double gain = z_var / (z_var + m_zVar);                         // :332
        z_pos += gain * z_delta;
        double var_gain = delta_var / (delta_var + m_zVar);             // :335
        delta_var = (1.0 - var_gain) * delta_var;
        z_var = pose_var + delta_var;
        return true;
    }

    // --- PoseEstimator::samplePose2D  src/PoseEstimator.cpp:64-73 ; init :88-102
    template <class A>
    void init(uint64_t n, const double mu[3], const double sigma[3], double zpos, double zsigma) {
        xi_k.clear();
        const uint64_t base = shard_base(n);
        for (uint64_t i = 0; i < n; i++) {
            const uint64_t gi = base + i;
            em::Philox4 a = noise.draw(gi, 0, 0, em::STREAM_INIT), b = noise.draw(gi, 0, 1, em::STREAM_INIT);
            double nx, ny, nt, unused;
            A::normal_pair(em::u01_open(a.v[0], a.v[1]), em::u01_open(a.v[2], a.v[3]), nx, ny);
            A::normal_pair(em::u01_open(b.v[0], b.v[1]), em::u01_open(b.v[2], b.v[3]), nt, unused);
            PoseParticle p;
            p.px = nx * sigma[0] + mu[0];                       // :69-72
            p.py = ny * sigma[1] + mu[1];
            p.orientation = nt * sigma[2] + mu[2];
            p.zPos = zpos; p.zSigma = zsigma;
            p.floating = true; p.weight = 0;                    // ctor, src/PoseParticle.hpp:55-56 (quirk Q9)
            p.mprob = 0; p.n_cpoints = 0;
            xi_k.push_back(p);
        }
    }

    uint64_t shard_base(uint64_t n_local) const {
        if (shard.world <= 1) return 0;
        return (uint64_t)shard.rank * n_local;
    }

    // --- weightingFunction  src/PoseEstimator.cpp:114-128
    static double weightingFunction(double x, double alpha, double beta, double gamma) {
        if (x < alpha) return 1.0;
        if (x < beta) {
            double a = (1.0 - gamma) / (alpha - beta);
            double b = 1.0 - alpha * a;
            return a * x + b;
        }
        if (x >= beta) return gamma;
        return 0.0;
    }

    // --- PoseEstimator::project  src/PoseEstimator.cpp:184-242
    template <class A>
    void project(const eslam_gpu_step& st) {
        const double yaw = st.yaw;                              // :186
        const double z_delta = st.z_delta;                      // :188-189
        const double z_var = st.z_var;                          // :192
        double spread = weightingFunction(max_weight, 0.0, cfg.spread_threshold, 0.0);   // :194
        const uint64_t base = shard_base(xi_k.size());
        for (size_t i = 0; i < xi_k.size(); i++) {              // :196
            const uint64_t gi = base + i;
            // odometry.getPoseDeltaSample2D() :198 -- un-vendored; defined as mean + sigma o n
            em::Philox4 a = noise.draw(gi, project_step, 0, em::STREAM_PROJECT);
            em::Philox4 b = noise.draw(gi, project_step, 1, em::STREAM_PROJECT);
            double n0, n1, n2, unused;
            A::normal_pair(em::u01_open(a.v[0], a.v[1]), em::u01_open(a.v[2], a.v[3]), n0, n1);
            A::normal_pair(em::u01_open(b.v[0], b.v[1]), em::u01_open(b.v[2], b.v[3]), n2, unused);
            double dx = st.delta_mean[0] + st.delta_sigma[0] * n0;
            double dy = st.delta_mean[1] + st.delta_sigma[1] * n1;
            double dth = st.delta_mean[2] + st.delta_sigma[2] * n2;
            em::Philox4 u = noise.draw(gi, project_step, 2, em::STREAM_PROJECT);
            if (em::u01(u.v[0], u.v[1]) < cfg.slip_factor)      // :199-202
                dy *= em::u01(u.v[2], u.v[3]);
            PoseParticle& p(xi_k[i]);
            double s, c; A::sincos(p.orientation, s, c);
            p.px += c * dx - s * dy;                            // :205 Rotation2D(orientation) * delta.position
            p.py += s * dx + c * dy;
            p.orientation += dth;                               // :206
            if (cfg.max_yaw_deviation > 0.0) {                  // :208-216 (no angle wrap, quirk Q7)
                if (std::fabs(p.orientation - yaw) > cfg.max_yaw_deviation) p.weight *= 0.7;
            }
            p.zPos += z_delta;                                  // :218
            p.zSigma = std::sqrt(p.zSigma * p.zSigma + z_var);  // :219
            if (spread > 0 && !cfg.use_hash) {                  // :224-236
                const double trans_fac = cfg.spread_translation_factor * spread;
                const double rot_fac = cfg.spread_rotation_factor * spread;
                em::Philox4 e = noise.draw(gi, project_step, 3, em::STREAM_PROJECT);
                em::Philox4 f = noise.draw(gi, project_step, 4, em::STREAM_PROJECT);
                double sx, sy, sth;
                A::normal_pair(em::u01_open(e.v[0], e.v[1]), em::u01_open(e.v[2], e.v[3]), sx, sy);
                A::normal_pair(em::u01_open(f.v[0], f.v[1]), em::u01_open(f.v[2], f.v[3]), sth, unused);
                p.px += sx * trans_fac + 0.0;                   // samplePose2D(Pose2D(), ...) :66-72
                p.py += sy * trans_fac + 0.0;
                p.orientation += sth * rot_fac + 0.0;
            }
        }
        project_step++;
    }

    // --- PoseEstimator::updateWeights  src/PoseEstimator.cpp:257-352
    template <class A>
    int updateWeights(eslam_gpu_update_result* res) {
        if (!grid.attached && !fake.active) return ESLAM_ERR_NO_ENVIRONMENT;    // :259-260
        size_t total_points = 0;                                // :264-266
        size_t data_particles = 0;
        std::vector<double> data_weights;                       // terms of sum_data_weights
        data_weights.reserve(xi_k.size());
        double last_max_weight = max_weight;                    // :268-269
        max_weight = 0;
        const bool keep_cps = cfg.patch_size > 0 && !patches.empty();
        for (size_t i = 0; i < xi_k.size(); i++) {              // :276
            PoseParticle& pose(xi_k[i]);
            const double measVar = pose.zSigma * pose.zSigma + cfg.measurement_error * cfg.measurement_error;  // :290
            bool ok;
            const double eval_x = pose.px, eval_y = pose.py;
            try { ok = evaluatePose<A>(pose.px, pose.py, pose.zPos, pose.orientation, measVar, nullptr, (long)i); }     // :288-291
            catch (std::string&) { return ESLAM_ERR_ZERO_VARIANCE; }
            if (ok) {
                double zVar = pose.zSigma * pose.zSigma;        // :294
                updateZPositionEstimate(pose.zPos, zVar);       // :295 (return value ignored)
                pose.zSigma = A::sqrt(zVar);                    // :296
                const double weight = m_weight;                 // :299
                pose.weight *= weight;                          // :300 (first application, quirk Q1)
                pose.mprob = weight;
                pose.floating = false;
                max_weight = std::max(max_weight, weight);      // :305
                data_particles++;                               // :307
                const size_t found_points = contact_points.size();
                data_weights.push_back(A::root_n(weight, found_points));        // :309
                total_points += found_points;                   // :310
            } else {
                pose.floating = true;                           // :316-317
                pose.mprob = 1.0;
                if (A::shared) data_weights.push_back(0.0);     // keeps the pairwise tree index-aligned
            }
            pose.n_cpoints = contact_points.size();             // :322 cpoints.swap(contact_points)
            if (keep_cps) patch_write(i, eval_x, eval_y);       // per-particle map write, before any resampling
        }
        const double sum_data_weights = A::sum(data_weights);
        const double floating_weight = data_particles > 0 ? sum_data_weights / data_particles : 1.0;   // :329
        for (size_t i = 0; i < xi_k.size(); i++) {              // :332-345
            PoseParticle& it(xi_k[i]);
            double factor = it.mprob * A::pow_u64(cfg.discount_factor * floating_weight, (size_t)4 - it.n_cpoints);  // :336 (Q2)
            it.weight *= factor;                                // second application (Q1)
        }
        if (total_points == 0) max_weight = last_max_weight * cfg.discount_factor;   // :347-348
        if (res) {
            res->max_weight = max_weight; res->floating_weight = floating_weight;
            res->data_particles = data_particles; res->total_points = total_points;
        }
        return ESLAM_OK;
    }

    // --- ParticleFilter::getWeightsSum / normalizeWeights  src/ParticleFilter.hpp:34-70
    template <class A>
    double normalizeWeights(double* sum_out = nullptr) {
        std::vector<double> w(xi_k.size());
        for (size_t n = 0; n < xi_k.size(); n++) w[n] = xi_k[n].weight;
        double sumWeights = A::sum(w);                          // :48
        if (sum_out) *sum_out = sumWeights;
        std::vector<double> sq(xi_k.size());
        if (sumWeights <= 0.0) {                                // :51-59 (quirk Q8)
            for (size_t n = 0; n < xi_k.size(); n++) {
                double& wn(xi_k[n].weight);
                wn = 1.0 / xi_k.size();
                sq[n] = wn * wn;
            }
        } else {                                                // :60-67
            for (size_t n = 0; n < xi_k.size(); n++) {
                double& wn(xi_k[n].weight);
                wn /= sumWeights;
                sq[n] = wn * wn;
            }
        }
        double effective = A::sum(sq);
        return 1.0 / effective;                                 // :69
    }

    // --- ParticleFilter::resample_stratified  src/ParticleFilter.hpp:85-108
    // Both policies keep the reference's left-to-right running sum: the GPU reproduces it exactly.
    int resample_stratified(size_t samples) {
        if (xi_k.empty()) return ESLAM_ERR_EMPTY;               // assert :91
        const uint64_t base = shard_base(xi_k.size());
        std::vector<PoseParticle> xi_kp;
        ancestors.assign(samples, 0);
        cdf.assign(xi_k.size(), 0.0);
        {   // the running sum the loop below walks, recorded for the parity tap
            double s = xi_k[0].weight; cdf[0] = s;
            for (size_t i = 1; i < xi_k.size(); i++) { s += xi_k[i].weight; cdf[i] = s; }
        }
        em::Philox4 first = noise.draw(0, resample_step, 0, em::STREAM_RESAMPLE);
        size_t idx = 0;
        double sum_w = xi_k[idx].weight;                        // :95
        for (size_t k = 0; k < samples; ++k) {                  // :96
            double u;
            if (cfg.resample_mode == ESLAM_RESAMPLE_SYSTEMATIC) u = em::u01(first.v[0], first.v[1]);
            else { em::Philox4 r = noise.draw(base + k, resample_step, 0, em::STREAM_RESAMPLE); u = em::u01(r.v[0], r.v[1]); }
            double sum_r = (k + u) / samples;                   // :98
            while (sum_w < sum_r) {                             // :99-103
                if (idx + 1 >= xi_k.size()) break;              // the reference has no guard (reads past the end
                ++idx;                                          //  when the weights sum below sum_r); we stop at N-1
                sum_w += xi_k[idx].weight;
            }
            xi_kp.push_back(xi_k[idx]);                         // :104 copy keeps the weight (quirk Q3)
            ancestors[k] = (uint32_t)idx;
        }
        xi_k.swap(xi_kp);                                       // :107
        if (cfg.patch_size > 0 && !patches.empty()) {           // cloneMaps, src/PoseEstimator.cpp:31-45: maps travel with copies
            const size_t cells = (size_t)cfg.patch_size * cfg.patch_size * 2;
            std::vector<float> np(samples * cells); std::vector<int32_t> no(2 * samples);
            for (size_t k = 0; k < samples; k++) {
                std::memcpy(&np[k * cells], &patches[ancestors[k] * cells], cells * sizeof(float));
                no[2 * k] = patch_origin[2 * ancestors[k]]; no[2 * k + 1] = patch_origin[2 * ancestors[k] + 1];
            }
            patches.swap(np); patch_origin.swap(no);
        }
        resample_step++;
        return ESLAM_OK;
    }
    // --- ParticleFilter::resample_multinomial(samples)  src/ParticleFilter.hpp:120-147 (unused by the filter itself)
    int resample_multinomial(size_t samples) {
        if (xi_k.empty()) return ESLAM_ERR_EMPTY;               // assert :126
        std::vector<PoseParticle> xi_kp;
        ancestors.assign(samples, 0);
        cdf.assign(xi_k.size(), 0.0);
        { double s = xi_k[0].weight; cdf[0] = s; for (size_t i = 1; i < xi_k.size(); i++) { s += xi_k[i].weight; cdf[i] = s; } }
        for (size_t n = 0; n < samples; n++) {                  // :129
            em::Philox4 r = noise.draw(n, resample_step, 0, em::STREAM_RESAMPLE);
            const double r_n = em::u01(r.v[0], r.v[1]);         // :132
            double sum = 0; size_t pick = xi_k.size() - 1;      // the reference drops a draw above the total (:134-143 finds nothing);
            for (size_t i = 0; i < xi_k.size(); i++) {          //  we keep the set size and take the last particle
                sum = i == 0 ? xi_k[0].weight : sum + xi_k[i].weight;   // :136
                if (r_n <= sum) { pick = i; break; }            // :137
            }
            PoseParticle p(xi_k[pick]);                         // :139
            p.weight = 1.0 / xi_k.size();                       // :140
            xi_kp.push_back(p);
            ancestors[n] = (uint32_t)pick;
        }
        if (cfg.patch_size > 0 && !patches.empty()) {
            const size_t cells = (size_t)cfg.patch_size * cfg.patch_size * 2;
            std::vector<float> np(samples * cells); std::vector<int32_t> no(2 * samples);
            for (size_t k = 0; k < samples; k++) {
                std::memcpy(&np[k * cells], &patches[ancestors[k] * cells], cells * sizeof(float));
                no[2 * k] = patch_origin[2 * ancestors[k]]; no[2 * k + 1] = patch_origin[2 * ancestors[k] + 1];
            }
            patches.swap(np); patch_origin.swap(no);
        }
        xi_k.swap(xi_kp);                                       // :146
        resample_step++;
        return ESLAM_OK;
    }

    // --- sharded resampling, LOCAL_MIGRATE mode (north_star / SURVEY.md §8e; no reference code behind it: this IS the definition,
    //     include/eslam_gpu.h ESLAM_SHARD_LOCAL_MIGRATE).  The oracle holds all N particles; the G logical shards are the
    //     contiguous ranges [g n_s, (g+1) n_s).
    int logical_shards() const {
        if (cfg.shard_mode != ESLAM_SHARD_LOCAL_MIGRATE || cfg.patch_size > 0) return 1;   // patches: exact-global only
        const int G = shard.world > 1 ? shard.world : cfg.virtual_shards;
        return (G > 1 && xi_k.size() % ((size_t)G * 1024) == 0) ? G : 1;    // shards are whole 1024-particle tiles
    }
    // decided on the weights as they are BEFORE normalisation (what every rank holds after the exchange of the per-tile sums):
    // migrate = some shard's share of the total weight, times G, is below migrate_min_share, or its own effective sample
    // size W_g^2 / sum w^2 is below migrate_min_ess n_s.  A non-positive total (uniform reset, Q8) never migrates.
    template <class A>
    void shard_decision() {
        const int G = logical_shards();
        migrate = 1;
        if (G <= 1) return;
        const size_t ns = xi_k.size() / G;
        std::vector<double> w(xi_k.size());
        for (size_t i = 0; i < xi_k.size(); i++) w[i] = xi_k[i].weight;
        const double total = A::sum(w);
        migrate = 0;
        if (total <= 0.0) return;
        for (int g = 0; g < G; g++) {
            std::vector<double> ws(w.begin() + g * ns, w.begin() + (g + 1) * ns), qs(ns);
            for (size_t i = 0; i < ns; i++) qs[i] = ws[i] * ws[i];
            const double Wg = A::sum(ws), Qg = A::sum(qs);
            const double share = (Wg / total) * (double)G;
            const double essg = (Wg * Wg) / Qg;
            if (share < cfg.migrate_min_share || essg < cfg.migrate_min_ess * (double)ns) migrate = 1;
        }
    }
    // every shard walks its own strata over its own left-to-right running sum (restarted at 0), scaled by the shard's total
    int resample_local(int G) {
        if (xi_k.empty()) return ESLAM_ERR_EMPTY;
        const size_t N = xi_k.size(), ns = N / G;
        std::vector<PoseParticle> xi_kp;
        ancestors.assign(N, 0);
        cdf.assign(N, 0.0);
        for (int g = 0; g < G; g++) {
            const size_t o = (size_t)g * ns;
            double s = xi_k[o].weight; cdf[o] = s;
            for (size_t i = 1; i < ns; i++) { s += xi_k[o + i].weight; cdf[o + i] = s; }
            const double Wg = s;
            size_t idx = 0;
            double sum_w = xi_k[o].weight;
            for (size_t k = 0; k < ns; ++k) {
                em::Philox4 r = noise.draw(o + k, resample_step, 0, em::STREAM_RESAMPLE);
                const double u = em::u01(r.v[0], r.v[1]);
                const double sum_r = (((double)k + u) / (double)ns) * Wg;
                while (sum_w < sum_r) {
                    if (idx + 1 >= ns) break;
                    ++idx;
                    sum_w += xi_k[o + idx].weight;
                }
                xi_kp.push_back(xi_k[o + idx]);                 // copies keep their weight, as in the global walk (Q3)
                ancestors[o + k] = (uint32_t)(o + idx);
            }
        }
        xi_k.swap(xi_kp);
        resample_step++;
        return ESLAM_OK;
    }
    int resample_any() {
        const int G = logical_shards();
        if (G > 1 && !migrate && cfg.resample_mode == ESLAM_RESAMPLE_STRATIFIED) return resample_local(G);
        return resample_stratified(xi_k.size());
    }

    // --- PoseEstimator::update  src/PoseEstimator.cpp:244-255
    template <class A>
    int update(const eslam_gpu_contact* pts, int n_pts, eslam_gpu_update_result* res) {
        contactPoints.assign(pts, pts + n_pts);                 // setContactPoints :262 (rotation done by the caller)
        eslam_gpu_update_result local; std::memset(&local, 0, sizeof(local));
        int st = updateWeights<A>(&local);                      // :247
        if (st != ESLAM_OK) { if (res) { *res = local; res->status = st; } return st; }
        double sum = 0;
        shard_decision<A>();
        double eff = normalizeWeights<A>(&sum);                 // :248
        local.ess = eff; local.weights_sum = sum; local.resampled = 0;
        if (eff < cfg.min_effective) {                          // :249
            st = resample_any();                                // :251 (resample_stratified(N) unless LOCAL_MIGRATE says otherwise)
            local.resampled = 1;
        }
        local.status = st;
        if (res) *res = local;
        return st;
    }

    // --- PoseEstimator::getCentroid  src/PoseEstimator.cpp:354-383
    template <class A>
    void centroid(double out[4]) {
        normalizeWeights<A>();                                  // :356
        std::vector<double> vx(xi_k.size()), vy(xi_k.size()), vt(xi_k.size()), vz(xi_k.size()), vw(xi_k.size());
        for (size_t i = 0; i < xi_k.size(); i++) {              // :362-372
            const PoseParticle& p(xi_k[i]);
            vx[i] = p.px * p.weight; vy[i] = p.py * p.weight; vt[i] = p.orientation * p.weight;
            vz[i] = p.zPos * p.weight; vw[i] = p.weight;
        }
        const double sumWeights = A::sum(vw);
        out[0] = A::sum(vx) / sumWeights; out[1] = A::sum(vy) / sumWeights;     // :373-375
        out[2] = A::sum(vt) / sumWeights; out[3] = A::sum(vz) / sumWeights;
    }
};

// ------------------------------------------------------------------------------------------------
// C API
// ------------------------------------------------------------------------------------------------
#define DISPATCH(ctx, expr_lit, expr_sh) ((ctx)->policy == ESLAM_ORACLE_SHARED ? (expr_sh) : (expr_lit))

extern "C" {

int eslam_oracle_create(const eslam_gpu_config* cfg, int policy, const eslam_gpu_shard* shard, eslam_oracle_ctx** out) {
    if (!cfg || !out) return ESLAM_ERR_INVALID;
    *out = new eslam_oracle_ctx(*cfg, policy);
    if (shard) (*out)->shard = *shard;
    return ESLAM_OK;
}
void eslam_oracle_destroy(eslam_oracle_ctx* ctx) { delete ctx; }

int eslam_oracle_set_grid(eslam_oracle_ctx* ctx, const float* cells, int width, int height, double sx, double sy,
                          double ox, double oy, const double* g2l) {
    if (!ctx || !cells || width <= 0 || height <= 0) return ESLAM_ERR_INVALID;
    Grid& g = ctx->grid;
    g.cells.assign(cells, cells + (size_t)width * height * 2);
    g.w = width; g.h = height; g.sx = sx; g.sy = sy; g.ox = ox; g.oy = oy;
    g.identity = (g2l == nullptr);
    if (g2l) std::memcpy(g.g2l, g2l, sizeof(g.g2l));
    g.attached = true; ctx->fake.active = false;
    return ESLAM_OK;
}

int eslam_oracle_set_fake_map(eslam_oracle_ctx* ctx, const double z[4], const double stddev[4], const uint8_t* res) {
    if (!ctx) return ESLAM_ERR_INVALID;
    FakeMap& f = ctx->fake;
    f.active = true; f.points.clear();
    for (int i = 0; i < 4; i++) { f.z[i] = z[i]; f.stddev[i] = stddev[i]; f.res[i] = res ? res[i] != 0 : true; }
    f.has_res = res != nullptr;
    return ESLAM_OK;
}

int eslam_oracle_get_queried_points(eslam_oracle_ctx* ctx, double* xyz, int max_points) {
    int n = (int)std::min<size_t>(ctx->fake.points.size(), (size_t)max_points);
    for (int i = 0; i < n; i++) { xyz[3 * i] = ctx->fake.points[i].x; xyz[3 * i + 1] = ctx->fake.points[i].y; xyz[3 * i + 2] = ctx->fake.points[i].z; }
    return (int)ctx->fake.points.size();
}

int eslam_oracle_init_gaussian(eslam_oracle_ctx* ctx, uint64_t n, const double mu[3], const double sigma[3], double zpos, double zsigma) {
    if (!ctx) return ESLAM_ERR_INVALID;
    if (ctx->policy == ESLAM_ORACLE_SHARED) ctx->init<Shared>(n, mu, sigma, zpos, zsigma);
    else ctx->init<Literal>(n, mu, sigma, zpos, zsigma);
    ctx->max_weight = 0;
    if (ctx->cfg.patch_size > 0) {
        const size_t cells = (size_t)ctx->cfg.patch_size * ctx->cfg.patch_size * 2;
        ctx->patches.assign(n * cells, std::numeric_limits<float>::quiet_NaN());
        ctx->patch_origin.assign(2 * n, INT32_MIN);
    }
    return ESLAM_OK;
}

int eslam_oracle_set_particles(eslam_oracle_ctx* ctx, uint64_t n, const double* x, const double* y, const double* theta,
                               const double* zpos, const double* zsigma, const double* weight) {
    if (!ctx) return ESLAM_ERR_INVALID;
    if (ctx->xi_k.size() != n) {
        PoseParticle p; std::memset(&p, 0, sizeof(p)); p.floating = true;
        ctx->xi_k.assign(n, p);
        if (ctx->cfg.patch_size > 0) {                          // a fresh set: empty maps, like eslam_gpu_set_particles
            const size_t cells = (size_t)ctx->cfg.patch_size * ctx->cfg.patch_size * 2;
            ctx->patches.assign(n * cells, std::numeric_limits<float>::quiet_NaN());
            ctx->patch_origin.assign(2 * n, INT32_MIN);
        }
    }
    for (uint64_t i = 0; i < n; i++) {
        PoseParticle& p = ctx->xi_k[i];
        if (x) p.px = x[i];
        if (y) p.py = y[i];
        if (theta) p.orientation = theta[i];
        if (zpos) p.zPos = zpos[i];
        if (zsigma) p.zSigma = zsigma[i];
        if (weight) p.weight = weight[i];
    }
    return ESLAM_OK;
}

int eslam_oracle_get_particles(eslam_oracle_ctx* ctx, double* x, double* y, double* theta, double* zpos, double* zsigma,
                               double* weight, double* mprob, uint8_t* floating, uint8_t* n_found) {
    if (!ctx) return ESLAM_ERR_INVALID;
    for (size_t i = 0; i < ctx->xi_k.size(); i++) {
        const PoseParticle& p = ctx->xi_k[i];
        if (x) x[i] = p.px;
        if (y) y[i] = p.py;
        if (theta) theta[i] = p.orientation;
        if (zpos) zpos[i] = p.zPos;
        if (zsigma) zsigma[i] = p.zSigma;
        if (weight) weight[i] = p.weight;
        if (mprob) mprob[i] = p.mprob;
        if (floating) floating[i] = p.floating ? 1 : 0;
        if (n_found) n_found[i] = (uint8_t)std::min<size_t>(p.n_cpoints, 255);
    }
    return ESLAM_OK;
}

uint64_t eslam_oracle_particle_count(const eslam_oracle_ctx* ctx) { return ctx ? ctx->xi_k.size() : 0; }

int eslam_oracle_project(eslam_oracle_ctx* ctx, const eslam_gpu_step* step) {
    if (!ctx || !step) return ESLAM_ERR_INVALID;
    if (ctx->policy == ESLAM_ORACLE_SHARED) ctx->project<Shared>(*step); else ctx->project<Literal>(*step);
    return ESLAM_OK;
}

int eslam_oracle_update(eslam_oracle_ctx* ctx, const eslam_gpu_contact* pts, int n_pts, eslam_gpu_update_result* result) {
    if (!ctx || !pts || n_pts < 0 || n_pts > ESLAM_MAX_CONTACTS) return ESLAM_ERR_INVALID;
    return DISPATCH(ctx, ctx->update<Literal>(pts, n_pts, result), ctx->update<Shared>(pts, n_pts, result));
}

int eslam_oracle_normalize_weights(eslam_oracle_ctx* ctx, double* ess) {
    if (!ctx) return ESLAM_ERR_INVALID;
    if (ctx->xi_k.empty()) return ESLAM_ERR_EMPTY;
    double e = DISPATCH(ctx, ctx->normalizeWeights<Literal>(), ctx->normalizeWeights<Shared>());
    if (ess) *ess = e;
    return ESLAM_OK;
}

int eslam_oracle_resample(eslam_oracle_ctx* ctx) {
    if (!ctx) return ESLAM_ERR_INVALID;
    if (ctx->xi_k.empty()) return ESLAM_ERR_EMPTY;
    DISPATCH(ctx, ctx->shard_decision<Literal>(), ctx->shard_decision<Shared>());   // on the weights as they are
    return ctx->resample_any();
}

int eslam_oracle_resample_n(eslam_oracle_ctx* ctx, uint64_t samples, int multinomial) {
    if (!ctx || samples == 0) return ESLAM_ERR_INVALID;
    if (ctx->xi_k.empty()) return ESLAM_ERR_EMPTY;
    return multinomial ? ctx->resample_multinomial((size_t)samples) : ctx->resample_stratified((size_t)samples);
}

int eslam_oracle_get_ancestors(eslam_oracle_ctx* ctx, uint32_t* a) {
    if (!ctx || !a) return ESLAM_ERR_INVALID;
    std::copy(ctx->ancestors.begin(), ctx->ancestors.end(), a);
    return ESLAM_OK;
}
int eslam_oracle_get_cdf(eslam_oracle_ctx* ctx, double* c) {
    if (!ctx || !c) return ESLAM_ERR_INVALID;
    std::copy(ctx->cdf.begin(), ctx->cdf.end(), c);
    return ESLAM_OK;
}

int eslam_oracle_centroid(eslam_oracle_ctx* ctx, double out[4]) {
    if (!ctx || ctx->xi_k.empty()) return ESLAM_ERR_EMPTY;
    if (ctx->policy == ESLAM_ORACLE_SHARED) ctx->centroid<Shared>(out); else ctx->centroid<Literal>(out);
    return ESLAM_OK;
}

int eslam_oracle_best_index(eslam_oracle_ctx* ctx, uint64_t* index) {   // src/ParticleFilter.hpp:160-173
    if (!ctx || !index) return ESLAM_ERR_INVALID;
    size_t best = 0;
    double weight = -std::numeric_limits<double>::infinity();
    for (size_t i = 0; i < ctx->xi_k.size(); i++) {
        if (ctx->xi_k[i].weight > weight) { best = i; weight = ctx->xi_k[i].weight; }
    }
    *index = best;
    return ESLAM_OK;
}

int eslam_oracle_evaluate_poses(eslam_oracle_ctx* ctx, uint64_t n, const double* x, const double* y, const double* theta,
                                const double* zpos, const double* meas_var, const double* z_var_in,
                                const eslam_gpu_contact* pts, int n_pts,
                                uint8_t* ok, double* weight, double* z_delta, double* z_var, double* pose_var,
                                int32_t* n_contacts, double* zpos_out, double* z_var_out,
                                int32_t* cell_index, double* cp) {
    if (!ctx || !pts || n_pts < 0 || n_pts > ESLAM_MAX_CONTACTS) return ESLAM_ERR_INVALID;
    if (!ctx->grid.attached && !ctx->fake.active) return ESLAM_ERR_NO_ENVIRONMENT;
    ctx->contactPoints.assign(pts, pts + n_pts);
    for (uint64_t i = 0; i < n; i++) {
        bool r;
        int32_t* cells = cell_index ? cell_index + i * n_pts : nullptr;
        try {
            r = DISPATCH(ctx, ctx->evaluatePose<Literal>(x[i], y[i], zpos[i], theta[i], meas_var[i], cells),
                         ctx->evaluatePose<Shared>(x[i], y[i], zpos[i], theta[i], meas_var[i], cells));
        } catch (std::string&) { return ESLAM_ERR_ZERO_VARIANCE; }
        if (ok) ok[i] = r;
        if (weight) weight[i] = r ? ctx->m_weight : 0.0;
        if (z_delta) z_delta[i] = r ? ctx->m_zDelta : 0.0;
        if (z_var) z_var[i] = r ? ctx->m_zVar : 0.0;
        if (pose_var) pose_var[i] = ctx->m_poseVar;
        if (n_contacts) n_contacts[i] = (int32_t)ctx->contact_points.size();
        double zp = zpos[i], zv = z_var_in ? z_var_in[i] : 0.0;
        if (r && z_var_in) ctx->updateZPositionEstimate(zp, zv);
        if (zpos_out) zpos_out[i] = zp;
        if (z_var_out) z_var_out[i] = zv;
        if (cp) {
            for (int k = 0; k < ESLAM_MAX_CP; k++) {
                double* o = cp + (i * ESLAM_MAX_CP + k) * 5;
                if ((size_t)k < ctx->contact_points.size()) {
                    const ContactPoint& q = ctx->contact_points[k];
                    o[0] = q.point.x; o[1] = q.point.y; o[2] = q.point.z; o[3] = q.zdiff; o[4] = q.zvar;
                } else o[0] = o[1] = o[2] = o[3] = o[4] = 0.0;
            }
        }
    }
    return ESLAM_OK;
}

// ContactModel::lowestPointHeuristic  src/ContactModel.cpp:48-79
int eslam_oracle_lowest_points(const eslam_gpu_contact* pts, int n_pts, int update_probabilities,
                               double* lowest_xyz, float* contact_out) {
    std::vector<eslam_gpu_contact> contactPoints(pts, pts + n_pts);
    std::vector<std::pair<double, int> > group;
    int count = 0;
    for (size_t i = 0; i < contactPoints.size(); i++) {
        if (contactPoints[i].group_id >= 0) {                   // :56-62
            group.push_back(std::make_pair(contactPoints[i].position[2], (int)i));
            if (update_probabilities) contactPoints[i].contact = 0;
        }
        if (group.empty()) {                                    // :64-65
            std::memcpy(lowest_xyz + 3 * count++, contactPoints[i].position, 24);
        } else if (i + 1 == contactPoints.size() || contactPoints[i + 1].group_id != contactPoints[i].group_id) {   // :66-77
            std::sort(group.begin(), group.end());
            eslam_gpu_contact& contact = contactPoints[group[0].second];
            std::memcpy(lowest_xyz + 3 * count++, contact.position, 24);
            if (update_probabilities) contact.contact = 1;
            group.clear();
        }
    }
    if (contact_out) for (int i = 0; i < n_pts; i++) contact_out[i] = contactPoints[i].contact;
    return count;
}

int eslam_oracle_process_map(eslam_oracle_ctx* ctx, const eslam_gpu_scan* scan, int match, int update) {
    if (!ctx || !scan || !scan->cells_mean_stdev || scan->width <= 0 || scan->height <= 0) return ESLAM_ERR_INVALID;
    if (ctx->xi_k.empty()) return ESLAM_ERR_EMPTY;
    return DISPATCH(ctx, ctx->process_map<Literal>(scan, match, update), ctx->process_map<Shared>(scan, match, update));
}

int eslam_oracle_pose_moments(eslam_oracle_ctx* ctx, eslam_gpu_gaussian2d* out) {
    if (!ctx || !out) return ESLAM_ERR_INVALID;
    if (ctx->xi_k.empty()) return ESLAM_ERR_EMPTY;
    DISPATCH(ctx, ctx->pose_moments<Literal>(out), ctx->pose_moments<Shared>(out));
    return ESLAM_OK;
}

int eslam_oracle_hash_create(eslam_oracle_ctx* ctx, int slope_bins, int angular_steps) {
    if (!ctx) return ESLAM_ERR_INVALID;
    return ctx->hash_create(slope_bins, angular_steps);
}
int eslam_oracle_hash_info(eslam_oracle_ctx* ctx, uint64_t* n_poses, uint32_t* bucket_counts) {
    if (!ctx) return ESLAM_ERR_INVALID;
    if (n_poses) *n_poses = ctx->hash_poses.size();
    if (bucket_counts && !ctx->hash_offs.empty()) for (size_t b = 0; b + 1 < ctx->hash_offs.size(); b++) bucket_counts[b] = ctx->hash_offs[b + 1] - ctx->hash_offs[b];
    return ESLAM_OK;
}
int eslam_oracle_hash_get(eslam_oracle_ctx* ctx, uint64_t first, uint64_t count, double* x, double* y, double* theta, double* z, int32_t* bucket) {
    if (!ctx || first + count > ctx->hash_poses.size()) return ESLAM_ERR_INVALID;
    for (uint64_t i = 0; i < count; i++) {
        const eslam_oracle_ctx::HashPose& hp = ctx->hash_poses[first + i];
        double px, py, th; ctx->hash_pose(hp, px, py, th);
        if (x) x[i] = px;
        if (y) y[i] = py;
        if (theta) theta[i] = th;
        if (z) z[i] = hp.z;
        if (bucket) bucket[i] = hp.bucket;
    }
    return ESLAM_OK;
}
int eslam_oracle_init_from_hash(eslam_oracle_ctx* ctx, uint64_t n) {
    if (!ctx || n == 0) return ESLAM_ERR_INVALID;
    return ctx->init_from_hash(n);
}
int eslam_oracle_sample_from_hash(eslam_oracle_ctx* ctx, double replace_percentage, double avg_factor, const eslam_gpu_contact* pts, int n_pts,
                                  uint64_t* replaced) {
    if (!ctx || !pts || n_pts < 0 || n_pts > ESLAM_MAX_CONTACTS) return ESLAM_ERR_INVALID;
    return DISPATCH(ctx, ctx->sample_from_hash<Literal>(replace_percentage, avg_factor, pts, n_pts, replaced),
                    ctx->sample_from_hash<Shared>(replace_percentage, avg_factor, pts, n_pts, replaced));
}

int eslam_oracle_get_patches(eslam_oracle_ctx* ctx, float* cells, int32_t* origin_xy) {
    if (!ctx || ctx->cfg.patch_size <= 0) return ESLAM_ERR_INVALID;
    if (cells) std::copy(ctx->patches.begin(), ctx->patches.end(), cells);
    if (origin_xy) std::copy(ctx->patch_origin.begin(), ctx->patch_origin.end(), origin_xy);
    return ESLAM_OK;
}

double eslam_oracle_math(int fn, double a, double b) {
    switch (fn) {
        case 0: return em::exp_(a);
        case 1: return em::log_(a);
        case 2: { double s, c; em::sincos_(a, s, c); return s; }
        case 3: { double s, c; em::sincos_(a, s, c); return c; }
        case 4: { double s, c; em::sincos_2pi(a, s, c); return s; }
        case 5: { double s, c; em::sincos_2pi(a, s, c); return c; }
        case 6: return em::rsqrt_(a);
        case 7: return em::h_ratio(a, H_TABLE);
        case 8: return em::likelihood_ratio(a, b, 1.0 / 0.33, H_TABLE);
        case 9: return Literal::likelihood_ratio(a, b, 0.33);
        case 10: return em::root_n(a, (unsigned long long)b);
        case 11: return em::pow_u64(a, (unsigned long long)b);
        case 12: { double n0, n1; em::box_muller(a, b, n0, n1); return n0; }
        case 13: { double n0, n1; em::box_muller(a, b, n0, n1); return n1; }
        default: return std::numeric_limits<double>::quiet_NaN();
    }
}

}  // extern "C"
