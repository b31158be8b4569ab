// ref_driver.cpp -- TEST INFRASTRUCTURE ONLY (same rules as eslam_oracle.h: only tests/, smoke() and bench.py's CPU legs
// may load what this builds).
//
// oracle/_ref/libeslam_ref.so = the REFERENCE'S OWN SOURCES, compiled unmodified from where they lie
// (/root/reference/src/ContactModel.cpp, src/PoseEstimator.cpp and the headers they include: ParticleFilter.hpp,
// PoseParticle.hpp, Configuration.hpp, ContactModel.hpp, PoseEstimator.hpp, SurfaceHash.hpp) against the stand-in headers
// of oracle/ref_shim/ for the dependencies that are absent here (boost, Eigen, base/types, envire, odometry,
// terrain_estimator), plus this driver, which exposes eslam::PoseEstimator / eslam::ContactModel through the same C
// signatures as the oracle (eslam_ref_* instead of eslam_oracle_*).
//
// What runs is therefore the reference's code for everything SURVEY.md §8(a) lists -- project, updateWeights loops A and
// B, ContactModel::evaluatePose / contactLikelihoodRatio / evaluateWeight / updateZPositionEstimate, normalizeWeights,
// resample_stratified, getCentroid, getBestParticleIndex, with all its quirks -- and what the stand-ins DEFINE is the
// behaviour of the un-vendored dependencies: the dense single-patch MLS grid, the odometry sampler
// mean + sigma o n, Eigen's small fixed-size algebra, boost's pdf/cdf expressions on libm.
//
// Noise: north_star compares the filters on identical noise streams.  The stand-in variate generators take their numbers
// from a hook (boost/random/linear_congruential.hpp of the shim); the hook below serves every draw site of the reference
// from the Philox counters of SURVEY.md §7.3 H3 in the order the reference consumes them:
//   init      : samplePose2D -> rand_norm x3 per particle            = stream INIT, slots 0 (pair), 1 (first)
//   project   : getPoseDeltaSample2D -> 3 normals per particle        = stream PROJECT, slots 0 (pair), 1 (first)
//               rand_uni (slip test), rand_uni (slip factor)         = slot 2, words (0,1) and (2,3)
//               samplePose2D (spread) -> rand_norm x3                 = slots 3 (pair), 4 (first)
//   resample  : the local `rand` of resample_stratified, stratum k    = stream RESAMPLE, counter k, slot 0, words (0,1)
#include <eslam/PoseEstimator.hpp>
#include <boost/bind.hpp>

#include "../eslam_oracle.h"                       // POD structs + status codes (via include/eslam_gpu.h)
#include "../../slam-eslam_b200/csrc/eslam_math.h"     // Philox4x32-10, u01 (integer/bit definitions of the noise stream)

#include <cmath>
#include <cstring>
#include <mutex>
#include <sstream>

namespace em = eslam_math;

namespace {

struct NullBuf : std::streambuf { int overflow(int c) { return c; } };

// eslam::ContactModel with its protected results readable
struct ProbeModel : eslam::ContactModel {
    double poseVar() const { return m_poseVar; }
};

enum Phase { PH_NONE, PH_INIT, PH_PROJECT, PH_UPDATE };

struct Hook;
thread_local Hook* g_hook = 0;         // one filter per thread at a time (bench.py runs independent filters on threads)

struct Hook {
    uint32_t k0, k1;
    Phase phase;
    const void* odo_site;
    uint32_t project_step, resample_step;
    uint64_t n;                      // particles
    long particle;                   // current particle of the phase
    int odo_k, uni_k, norm_k;        // ordinals inside the current particle
    uint64_t init_draws, res_draws;
    double pair_second; bool have_second;

    em::Philox4 draw(uint64_t index, uint32_t step, uint32_t slot, uint32_t stream) const {
        return em::philox4x32_10((uint32_t)index, step, slot, stream ^ ((uint32_t)(index >> 32) << 8), k0, k1);
    }
    // the LITERAL normal pair (oracle Literal::normal_pair): Box-Muller on two open uniforms, libm
    static void normal_pair(const em::Philox4& p, double& n0, double& n1) {
        const double u1 = em::u01_open(p.v[0], p.v[1]), u2 = em::u01_open(p.v[2], p.v[3]);
        const double r = std::sqrt(-2.0 * std::log(u1));
        n0 = r * std::cos(2.0 * M_PI * u2);
        n1 = r * std::sin(2.0 * M_PI * u2);
    }
    double normal(uint64_t index, uint32_t step, uint32_t slot_base, uint32_t stream, int k) const {
        double a, b;
        normal_pair(draw(index, step, slot_base + (k >> 1), stream), a, b);
        return (k & 1) ? b : a;
    }
    double next(int kind, const void* site) {
        switch (phase) {
        case PH_INIT: {                                          // src/PoseEstimator.cpp:66 x, y, theta
            const uint64_t i = init_draws / 3; const int k = (int)(init_draws % 3); ++init_draws;
            return normal(i, 0, 0, em::STREAM_INIT, k);
        }
        case PH_PROJECT:
            if (site == odo_site) {                              // :198
                if (odo_k == 0) { ++particle; uni_k = 0; norm_k = 0; }
                const double v = normal((uint64_t)particle, project_step, 0, em::STREAM_PROJECT, odo_k);
                odo_k = (odo_k + 1) % 3;
                return v;
            }
            if (kind == eslam_ref_shim::DRAW_UNIFORM) {          // :199, :201
                const em::Philox4 u = draw((uint64_t)particle, project_step, 2, em::STREAM_PROJECT);
                const int k = uni_k++;
                return k == 0 ? em::u01(u.v[0], u.v[1]) : em::u01(u.v[2], u.v[3]);
            }
            return normal((uint64_t)particle, project_step, 3, em::STREAM_PROJECT, norm_k++);   // :228-230
        case PH_UPDATE: {                                        // src/ParticleFilter.hpp:98
            const em::Philox4 r = draw(res_draws++, resample_step, 0, em::STREAM_RESAMPLE);
            return em::u01(r.v[0], r.v[1]);
        }
        default: return 0.5;
        }
    }
};

double hook_fn(int kind, const void* site) { return g_hook->next(kind, site); }

}  // namespace

struct eslam_ref_ctx {
    eslam_gpu_config cfg;
    eslam::Configuration rcfg;
    odometry::FootContact odo;
    eslam::PoseEstimator* pe;
    envire::Environment env;
    envire::FrameNode* frame;
    envire::MLSMap::Ptr map;
    boost::intrusive_ptr<envire::MLSGrid> grid;
    bool env_dirty, debug;
    Hook hook;
    std::vector<uint32_t> ancestors;

    eslam_ref_ctx(const eslam_gpu_config& c, bool dbg) : cfg(c), pe(0), frame(0), env_dirty(false), debug(dbg) {
        rcfg.seed = c.seed;                                      // src/Configuration.hpp:85-111
        rcfg.minEffective = c.min_effective;
        rcfg.measurementError = c.measurement_error;
        rcfg.discountFactor = c.discount_factor;
        rcfg.spreadThreshold = c.spread_threshold;
        rcfg.spreadTranslationFactor = c.spread_translation_factor;
        rcfg.spreadRotationFactor = c.spread_rotation_factor;
        rcfg.slipFactor = c.slip_factor;
        rcfg.maxYawDeviation = c.max_yaw_deviation;
        rcfg.contactModel.minContacts = c.min_contacts;
        rcfg.contactModel.contactLikelihoodCorrection = c.contact_likelihood_correction;
        rcfg.contactModel.contactPointRadius = c.contact_point_radius;
        rcfg.contactModel.useSlipUpdate = c.use_slip_update != 0;
        rcfg.contactModel.useShapeUpdate = c.use_shape_update != 0;
        rcfg.logDebug = dbg;                                     // keeps cpoints (n_found) readable after the update
        pe = new eslam::PoseEstimator(odo, rcfg);
        std::memset(&hook, 0, sizeof(hook));
        hook.k0 = (uint32_t)c.seed; hook.k1 = (uint32_t)(c.seed >> 32);
        hook.odo_site = &odo;
    }
    ~eslam_ref_ctx() { delete pe; }

    void enter(Phase ph) {
        g_hook = &hook; eslam_ref_shim::noise_hook() = &hook_fn;
        hook.phase = ph; hook.particle = -1; hook.odo_k = hook.uni_k = hook.norm_k = 0; hook.init_draws = hook.res_draws = 0;
    }
    void leave() { hook.phase = PH_NONE; }
    void attach_env() {                                          // PoseEstimator::setEnvironment, src/PoseEstimator.cpp:47-62
        if (map.get() && env_dirty) { pe->setEnvironment(&env, map, true); env_dirty = false; }
    }
};

namespace {
odometry::BodyContactState make_state(const eslam_gpu_contact* pts, int n) {
    odometry::BodyContactState s;
    s.points.resize(n);
    for (int i = 0; i < n; ++i) {
        s.points[i].position = base::Vector3d(pts[i].position[0], pts[i].position[1], pts[i].position[2]);
        s.points[i].contact = pts[i].contact; s.points[i].slip = pts[i].slip; s.points[i].groupId = pts[i].group_id;
    }
    return s;
}
// process-wide, counted: std::cerr is shared by the filters that bench.py runs on several threads
struct CerrMute {
    static std::mutex& mu() { static std::mutex m; return m; }
    static int& depth() { static int d = 0; return d; }
    static std::streambuf*& saved() { static std::streambuf* b = 0; return b; }
    static NullBuf& sink() { static NullBuf n; return n; }
    CerrMute() { std::lock_guard<std::mutex> g(mu()); if (depth()++ == 0) saved() = std::cerr.rdbuf(&sink()); }
    ~CerrMute() { std::lock_guard<std::mutex> g(mu()); if (--depth() == 0) std::cerr.rdbuf(saved()); }
};
}

extern "C" {

int eslam_ref_create(const eslam_gpu_config* cfg, int debug, eslam_ref_ctx** out) {
    if (!cfg || !out) return ESLAM_ERR_INVALID;
    if (cfg->resample_mode != ESLAM_RESAMPLE_STRATIFIED || cfg->patch_size != 0 || cfg->use_hash) return ESLAM_ERR_INVALID;
    *out = new eslam_ref_ctx(*cfg, debug != 0);
    return ESLAM_OK;
}
void eslam_ref_destroy(eslam_ref_ctx* ctx) { delete ctx; }

int eslam_ref_set_grid(eslam_ref_ctx* ctx, const float* cells, int width, int height, double sx, double sy, double ox,
                       double oy, const double* g2l) {
    if (!ctx || !cells || width <= 0 || height <= 0) return ESLAM_ERR_INVALID;
    ctx->grid = boost::intrusive_ptr<envire::MLSGrid>(new envire::MLSGrid(width, height, sx, sy, ox, oy));
    ctx->grid->cells.assign(cells, cells + (size_t)width * height * 2);
    ctx->frame = new envire::FrameNode();                        // lives as long as the process (a few bytes per set_grid)
    ctx->frame->env = &ctx->env;
    if (g2l) {                                                   // C_global2local = relativeTransform(root, frame) = frame.toRoot^-1
        Eigen::Affine3d g;
        for (int i = 0; i < 3; ++i) { for (int j = 0; j < 3; ++j) g.linear()(i, j) = g2l[4 * i + j]; g.translation()[i] = g2l[4 * i + 3]; }
        ctx->frame->toRoot = g.inverse();
    }
    envire::MLSMap* m = new envire::MLSMap();
    m->grids.push_back(ctx->grid);
    ctx->map = envire::MLSMap::Ptr(m);
    ctx->env.setFrameNode(m, ctx->frame);
    ctx->env_dirty = true;
    return ESLAM_OK;
}

int eslam_ref_init_gaussian(eslam_ref_ctx* ctx, uint64_t n, const double mu[3], const double sigma[3], double zpos, double zsigma) {
    if (!ctx) return ESLAM_ERR_INVALID;
    ctx->pe->getParticles().clear();
    ctx->enter(PH_INIT);
    ctx->pe->init((int)n, base::Pose2D(base::Vector2d(mu[0], mu[1]), mu[2]), base::Pose2D(base::Vector2d(sigma[0], sigma[1]), sigma[2]),
                  zpos, zsigma);                                 // src/PoseEstimator.cpp:88-102
    ctx->leave();
    ctx->env_dirty = true;
    return ESLAM_OK;
}

int eslam_ref_set_particles(eslam_ref_ctx* ctx, uint64_t n, const double* x, const double* y, const double* theta,
                            const double* zpos, const double* zsigma, const double* weight) {
    if (!ctx) return ESLAM_ERR_INVALID;
    std::vector<eslam::PoseEstimator::Particle>& v = ctx->pe->getParticles();
    if (v.size() != n) {
        v.clear();
        for (uint64_t i = 0; i < n; ++i) v.push_back(eslam::PoseEstimator::Particle(base::Vector2d(0, 0), 0.0));
        ctx->env_dirty = true;
    }
    for (uint64_t i = 0; i < n; ++i) {
        if (x) v[i].position.x() = x[i];
        if (y) v[i].position.y() = y[i];
        if (theta) v[i].orientation = theta[i];
        if (zpos) v[i].zPos = zpos[i];
        if (zsigma) v[i].zSigma = zsigma[i];
        if (weight) v[i].weight = weight[i];
    }
    return ESLAM_OK;
}

int eslam_ref_get_particles(eslam_ref_ctx* ctx, double* x, double* y, double* theta, double* zpos, double* zsigma,
                            double* weight, double* mprob, uint8_t* floating, uint8_t* n_found) {
    if (!ctx) return ESLAM_ERR_INVALID;
    const std::vector<eslam::PoseEstimator::Particle>& v = ctx->pe->getParticles();
    for (size_t i = 0; i < v.size(); ++i) {
        if (x) x[i] = v[i].position.x();
        if (y) y[i] = v[i].position.y();
        if (theta) theta[i] = v[i].orientation;
        if (zpos) zpos[i] = v[i].zPos;
        if (zsigma) zsigma[i] = v[i].zSigma;
        if (weight) weight[i] = v[i].weight;
        if (mprob) mprob[i] = v[i].mprob;
        if (floating) floating[i] = v[i].floating ? 1 : 0;
        if (n_found) n_found[i] = ctx->debug ? (uint8_t)std::min<size_t>(v[i].cpoints.size(), 255) : 255;
    }
    return ESLAM_OK;
}

uint64_t eslam_ref_particle_count(const eslam_ref_ctx* ctx) { return ctx ? ctx->pe->getParticles().size() : 0; }

int eslam_ref_project(eslam_ref_ctx* ctx, const eslam_gpu_step* st) {
    if (!ctx || !st) return ESLAM_ERR_INVALID;
    // the facade's derivation in reverse: orientation = yaw only, pose delta = (0, 0, z_delta), position error zz = z_var / 2
    const base::Quaterniond q(Eigen::AngleAxisd(st->yaw, Eigen::Vector3d::UnitZ()));
    ctx->odo.delta = base::Pose(base::Vector3d(0, 0, st->z_delta), base::Quaterniond::Identity());
    ctx->odo.positionError = base::Matrix3d::Zero();
    ctx->odo.positionError(2, 2) = st->z_var / 2.0;
    for (int i = 0; i < 3; ++i) { ctx->odo.mean2d[i] = st->delta_mean[i]; ctx->odo.sigma2d[i] = st->delta_sigma[i]; }
    ctx->attach_env();
    ctx->enter(PH_PROJECT);
    ctx->pe->project(odometry::BodyContactState(), q);           // src/PoseEstimator.cpp:184-242
    ctx->leave();
    ctx->hook.project_step++;
    return ESLAM_OK;
}

int eslam_ref_update(eslam_ref_ctx* ctx, const eslam_gpu_contact* pts, int n_pts, eslam_gpu_update_result* res) {
    if (!ctx || (!pts && n_pts > 0)) return ESLAM_ERR_INVALID;
    std::vector<eslam::PoseEstimator::Particle>& v = ctx->pe->getParticles();
    ctx->attach_env();
    if (ctx->debug) {                                            // ancestor tags travel in the debug vector of every particle
        for (size_t i = 0; i < v.size(); ++i) {
            v[i].spoints.assign(1, eslam::SlipPoint());
            v[i].spoints[0].prob = (double)i;
            v[i].cpoints.clear();
        }
    }
    const void* before = v.empty() ? 0 : (const void*)&v[0];
    eslam_gpu_update_result local; std::memset(&local, 0, sizeof(local));
    const odometry::BodyContactState state = make_state(pts, n_pts);
    ctx->enter(PH_UPDATE);
    int status = ESLAM_OK;
    try {
        CerrMute mute;                            // "iteration: ..." progress line, src/PoseEstimator.cpp:351
        // orientation = identity: the contact positions arrive already rotated by removeYaw(orientation) (eslam_gpu_contact)
        ctx->pe->update(state, base::Quaterniond::Identity(), std::vector<terrain_estimator::TerrainClassification>());
    } catch (std::runtime_error& e) {
        status = std::strstr(e.what(), "No environment") ? ESLAM_ERR_NO_ENVIRONMENT : ESLAM_ERR_ZERO_VARIANCE;
    }
    ctx->leave();
    std::vector<eslam::PoseEstimator::Particle>& w = ctx->pe->getParticles();
    const bool resampled = ctx->hook.res_draws > 0 || (!w.empty() && (const void*)&w[0] != before);
    if (resampled) ctx->hook.resample_step++;
    local.resampled = resampled ? 1 : 0;
    local.status = status;
    local.ess = std::nan("");
    if (status == ESLAM_OK) {
        for (size_t i = 0; i < w.size(); ++i) if (!w[i].floating) local.data_particles++;
        if (!resampled) {                                        // the value normalizeWeights() returned: same terms, same order
            double effective = 0;
            for (size_t i = 0; i < w.size(); ++i) effective += w[i].weight * w[i].weight;
            local.ess = 1.0 / effective;
        }
        if (ctx->debug) {
            ctx->ancestors.resize(w.size());
            for (size_t i = 0; i < w.size(); ++i) {
                ctx->ancestors[i] = resampled ? (uint32_t)w[i].spoints[0].prob : (uint32_t)i;
                if (!w[i].floating) local.total_points += w[i].cpoints.size();   // :310 counts successful particles only
            }
            if (resampled) local.data_particles = 0, local.total_points = 0;   // counted over copies: not the update's numbers
        }
    }
    local.max_weight = local.floating_weight = local.weights_sum = std::nan("");   // private to the reference
    if (res) *res = local;
    return status;
}

int eslam_ref_normalize_weights(eslam_ref_ctx* ctx, double* ess) {
    if (!ctx) return ESLAM_ERR_INVALID;
    const double e = ctx->pe->normalizeWeights();                // src/ParticleFilter.hpp:46-70
    if (ess) *ess = e;
    return ESLAM_OK;
}

int eslam_ref_resample(eslam_ref_ctx* ctx) {
    if (!ctx) return ESLAM_ERR_INVALID;
    std::vector<eslam::PoseEstimator::Particle>& v = ctx->pe->getParticles();
    if (v.empty()) return ESLAM_ERR_EMPTY;                       // assert at src/ParticleFilter.hpp:91
    for (size_t i = 0; i < v.size(); ++i) { v[i].spoints.assign(1, eslam::SlipPoint()); v[i].spoints[0].prob = (double)i; }
    ctx->enter(PH_UPDATE);
    ctx->pe->resample();                                         // src/ParticleFilter.hpp:72-108
    ctx->leave();
    ctx->hook.resample_step++;
    std::vector<eslam::PoseEstimator::Particle>& w = ctx->pe->getParticles();
    ctx->ancestors.resize(w.size());
    for (size_t i = 0; i < w.size(); ++i) ctx->ancestors[i] = (uint32_t)w[i].spoints[0].prob;
    return ESLAM_OK;
}

int eslam_ref_get_ancestors(eslam_ref_ctx* ctx, uint32_t* a) {
    if (!ctx || !a || ctx->ancestors.empty()) return ESLAM_ERR_INVALID;
    std::memcpy(a, ctx->ancestors.data(), ctx->ancestors.size() * sizeof(uint32_t));
    return ESLAM_OK;
}

int eslam_ref_centroid(eslam_ref_ctx* ctx, double out[4]) {
    if (!ctx || !out) return ESLAM_ERR_INVALID;
    // getCentroid returns Pose(position, AngleAxis(mean theta, Z) * zCompensatedOrientation): with the yaw-only orientations
    // this driver passes, the mean heading is read back from the rotation (src/PoseEstimator.cpp:354-383)
    const base::Pose p = ctx->pe->getCentroid();
    out[0] = p.position.x(); out[1] = p.position.y(); out[3] = p.position.z();
    out[2] = base::getYaw(p.orientation);
    return ESLAM_OK;
}

int eslam_ref_best_index(eslam_ref_ctx* ctx, uint64_t* index) {
    if (!ctx || !index) return ESLAM_ERR_INVALID;
    *index = ctx->pe->getBestParticleIndex();                    // src/ParticleFilter.hpp:160-173
    return ESLAM_OK;
}

// FakeMLSAccess of test/testContactModel.cpp:8-38 (set with eslam_ref_set_fake_map) or the grid, behind the map callback
struct RefMapProbe {
    eslam_ref_ctx* ctx;
    const double* fz; const double* fs; const uint8_t* fres;
    std::vector<int32_t> cells; std::vector<uint8_t> hit; std::vector<base::Vector3d> points;
    bool get(const base::Vector3d& pos, envire::MLSGrid::SurfacePatch& patch) {
        points.push_back(pos);
        if (fz) {
            int index = 0;
            if (pos.y() > 0) index += 2;
            if (pos.x() > 0) index++;
            patch.mean = fz[index]; patch.stdev = fs[index]; patch.height = 0; patch.setHorizontal();
            cells.push_back(index);
            const bool r = fres ? fres[index] != 0 : true;
            hit.push_back(r); return r;
        }
        ctx->grid->last_cell = -1;
        const bool r = ctx->map->getPatch(ctx->env.relativeTransform(ctx->env.getRootNode(), ctx->frame) * pos, patch, 3.0);
        cells.push_back((int32_t)ctx->grid->last_cell);
        hit.push_back(r); return r;
    }
};

static double g_fake_z[4], g_fake_s[4]; static uint8_t g_fake_res[4]; static int g_fake = 0;   // 0 off, 1 on, 2 on with res
static std::vector<base::Vector3d> g_queried;

int eslam_ref_set_fake_map(eslam_ref_ctx*, const double z[4], const double stddev[4], const uint8_t* res) {
    for (int i = 0; i < 4; ++i) { g_fake_z[i] = z[i]; g_fake_s[i] = stddev[i]; g_fake_res[i] = res ? res[i] : 1; }
    g_fake = res ? 2 : 1; g_queried.clear();
    return ESLAM_OK;
}
int eslam_ref_get_queried_points(eslam_ref_ctx*, double* xyz, int max_points) {
    const int n = (int)std::min<size_t>(g_queried.size(), (size_t)max_points);
    for (int i = 0; i < n; ++i) { xyz[3 * i] = g_queried[i].x(); xyz[3 * i + 1] = g_queried[i].y(); xyz[3 * i + 2] = g_queried[i].z(); }
    return (int)g_queried.size();
}

int eslam_ref_evaluate_poses(eslam_ref_ctx* ctx, uint64_t n, const double* x, const double* y, const double* theta,
                             const double* zpos, const double* meas_var, const double* z_var_in,
                             const eslam_gpu_contact* pts, int n_pts,
                             uint8_t* ok, double* weight, double* z_delta, double* z_var, double* pose_var,
                             int32_t* n_contacts, double* zpos_out, double* z_var_out, int32_t* cell_index, double* cp) {
    if (!ctx || !pts || n_pts <= 0) return ESLAM_ERR_INVALID;
    if (!g_fake && !ctx->map.get()) return ESLAM_ERR_NO_ENVIRONMENT;
    ProbeModel model;
    model.setConfiguration(ctx->rcfg.contactModel);
    const odometry::BodyContactState state = make_state(pts, n_pts);
    for (uint64_t i = 0; i < n; ++i) {
        model.setContactPoints(state, base::Quaterniond::Identity());          // src/ContactModel.cpp:21-41
        RefMapProbe probe; probe.ctx = ctx;
        probe.fz = g_fake ? g_fake_z : 0; probe.fs = g_fake_s; probe.fres = g_fake == 2 ? g_fake_res : 0;
        const base::Affine3d t = Eigen::Translation3d(base::Vector3d(x[i], y[i], zpos[i])) * Eigen::AngleAxisd(theta[i], Eigen::Vector3d::UnitZ());
        bool good;
        try { good = model.evaluatePose(t, meas_var[i], boost::bind(&RefMapProbe::get, &probe, _1, _2)); }   // :117-224
        catch (std::runtime_error&) { return ESLAM_ERR_ZERO_VARIANCE; }
        if (g_fake) g_queried.insert(g_queried.end(), probe.points.begin(), probe.points.end());
        const std::vector<eslam::ContactPoint>& cps = model.getContactPoints();
        if (ok) ok[i] = good;
        if (n_contacts) n_contacts[i] = (int32_t)cps.size();
        if (pose_var) pose_var[i] = model.poseVar();
        if (good) {
            if (weight) weight[i] = model.getWeight();
            if (z_delta) z_delta[i] = model.getZDelta();
            if (z_var) z_var[i] = model.getZVar();
            if (z_var_in && zpos_out && z_var_out) {
                double zp = zpos[i], zv = z_var_in[i];
                model.updateZPositionEstimate(zp, zv);                          // :319-340
                zpos_out[i] = zp; z_var_out[i] = zv;
            }
        }
        if (cp) for (size_t k = 0; k < cps.size() && k < ESLAM_MAX_CP; ++k) {
            double* o = cp + (i * ESLAM_MAX_CP + k) * 5;
            o[0] = cps[k].point.x(); o[1] = cps[k].point.y(); o[2] = cps[k].point.z(); o[3] = cps[k].zdiff; o[4] = cps[k].zvar;
        }
        if (cell_index) {
            // which contact each recorded lookup belongs to: lookups happen in contact order for the contacts that pass
            // `group_valid && !(contact < 0.2)` (:154); group_valid falls on a miss (:189-190) and is restored at the end of a
            // group that had a hit (:194-213)
            int32_t* row = cell_index + i * (uint64_t)n_pts;
            size_t k = 0; bool valid = false, gvalid = true;
            for (int j = 0; j < n_pts; ++j) {
                row[j] = -1;
                if (gvalid && !(pts[j].contact < 0.2f)) {
                    if (k >= probe.cells.size()) return ESLAM_ERR_INVALID;
                    row[j] = probe.cells[k];
                    if (probe.hit[k]) valid = true; else gvalid = false;
                    ++k;
                }
                if (valid && (pts[j].group_id == -1 || j + 1 == n_pts || pts[j].group_id != pts[j + 1].group_id)) { gvalid = true; valid = false; }
            }
            if (k != probe.cells.size()) return ESLAM_ERR_INVALID;
        }
    }
    return ESLAM_OK;
}

// ContactModel::getLowestPointPerGroup / updateContactStateUsingLowestPointHeuristic (src/ContactModel.cpp:48-92)
int eslam_ref_lowest_points(const eslam_gpu_contact* pts, int n_pts, int update_probabilities, double* lowest_xyz, float* contact_out) {
    eslam::ContactModel model;
    model.setContactPoints(make_state(pts, n_pts), base::Quaterniond::Identity());
    if (update_probabilities) model.updateContactStateUsingLowestPointHeuristic();
    const std::vector<base::Vector3d>& low = model.getLowestPointPerGroup();
    for (size_t i = 0; i < low.size(); ++i) { lowest_xyz[3 * i] = low[i].x(); lowest_xyz[3 * i + 1] = low[i].y(); lowest_xyz[3 * i + 2] = low[i].z(); }
    if (contact_out) for (int i = 0; i < n_pts; ++i) contact_out[i] = model.getContactState().points[i].contact;
    return (int)low.size();
}

void eslam_ref_clear_fake_map(void) { g_fake = 0; g_queried.clear(); }

}  // extern "C"
