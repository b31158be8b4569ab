/* eslam_gpu.h -- C ABI of libeslam_gpu.so: the B200 implementation of eSLAM's particle-filter
 * update path.  Plain pointers and sizes only; every call returns an int status (0 = ok) and
 * never throws.  The host C++ facade (include/eslam/ *.hpp) keeps the reference's class API
 * (PoseEstimator / ContactModel / EmbodiedSlamFilter / ParticleFilter<PoseParticle>) on top of
 * these entry points; INTEGRATION.md shows the binding a maintainer of the reference would add.
 *
 * Each entry point names the reference interface it replaces (paths relative to the reference
 * checkout, liyangSKD/slam-eslam).
 */
#ifndef ESLAM_GPU_H
#define ESLAM_GPU_H
#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define ESLAM_MAX_CONTACTS 128   /* body contact points per update (Asguard: 4 wheels x 5 feet) */
#define ESLAM_MAX_CP 8           /* fused contact points reported per pose by the diagnostic call */

enum {
    ESLAM_OK = 0,
    ESLAM_ERR_INVALID = 1,       /* bad argument */
    ESLAM_ERR_CUDA = 2,          /* CUDA runtime failure, see eslam_gpu_last_error */
    ESLAM_ERR_NO_ENVIRONMENT = 3,/* "No environment attached." src/PoseEstimator.cpp:259-260 */
    ESLAM_ERR_ZERO_VARIANCE = 4, /* zero measurement variance, src/ContactModel.cpp:122-123 */
    ESLAM_ERR_EMPTY = 5,         /* empty particle set, assert at src/ParticleFilter.hpp:91 */
    ESLAM_ERR_COMM = 6           /* communicator / peer-memory exchange failed (a rank did not arrive, no peer access) */
};

enum { ESLAM_RESAMPLE_STRATIFIED = 0,   /* reference: src/ParticleFilter.hpp:85-108 */
       ESLAM_RESAMPLE_SYSTEMATIC = 1 }; /* north_star: one uniform for all strata */

/* How a SHARDED filter (particles split into equal contiguous shards, one per GPU) resamples (north_star):
 *   EXACT_GLOBAL  : one stratified walk over all particles; ancestors may lie on any shard and are identical to the
 *                   one-GPU / reference run.  The parity mode and the default.
 *   LOCAL_MIGRATE : every shard resamples its own slots from its own particles (stratum k of shard g draws
 *                   (k + u)/n_s * W_g on the shard's own left-to-right running sum, W_g its total); particles cross shards
 *                   only in a step where some shard has collapsed -- its share of the total weight fell below
 *                   migrate_min_share of the fair share 1/G, or its own effective sample size W_g^2 / sum w^2 fell below
 *                   migrate_min_ess * n_s -- and that step resamples exact-globally (the migration).  Statistically a
 *                   valid particle filter, but not index-identical to the reference: defined by the oracle. */
enum { ESLAM_SHARD_EXACT_GLOBAL = 0, ESLAM_SHARD_LOCAL_MIGRATE = 1 };

/* Hot-path fields of eslam::Configuration / ContactModelConfiguration
 * (src/Configuration.hpp:51-111).  eslam_gpu_default_config() fills the reference defaults. */
typedef struct eslam_gpu_config {
    uint64_t seed;                       /* :114 seed */
    uint64_t min_effective;              /* :120 minEffective (size_t) */
    double measurement_error;            /* :136 */
    double discount_factor;              /* :141 */
    double spread_threshold;             /* :146 */
    double spread_translation_factor;    /* :149 */
    double spread_rotation_factor;       /* :152 */
    double slip_factor;                  /* :158 */
    double max_yaw_deviation;            /* :163 */
    uint64_t min_contacts;               /* ContactModelConfiguration::minContacts :72 */
    double contact_likelihood_correction;/* :76 */
    double contact_point_radius;         /* :80 */
    int32_t use_slip_update;             /* :64 (inert in the reference, quirk Q5) */
    int32_t use_shape_update;            /* :67 */
    int32_t use_hash;                    /* SurfaceHash attached: disables spreading, src/PoseEstimator.cpp:224 */
    int32_t resample_mode;               /* ESLAM_RESAMPLE_* */
    int32_t patch_size;                  /* 0: shared map only; P>0: per-particle PxP map patches */
    int32_t shard_mode;                  /* ESLAM_SHARD_* (only matters with more than one shard) */
    double migrate_min_share;            /* LOCAL_MIGRATE: migrate when a shard's weight share * G falls below this (default 0.25) */
    double migrate_min_ess;              /* LOCAL_MIGRATE: ... or its own ESS / n_s falls below this (default 0.05) */
    int32_t virtual_shards;              /* > 1 on an unsharded context: treat the particles as that many logical shards (the
                                            shard logic on one GPU; results identical to the same number of real shards) */
    int32_t reserved;
} eslam_gpu_config;

/* Particle shard owned by this context (multi-GPU: one context per rank). */
typedef struct eslam_gpu_shard {
    int32_t rank, world;
    uint64_t global_count;               /* total particles over all ranks */
} eslam_gpu_shard;

/* Per-step inputs of PoseEstimator::project (src/PoseEstimator.cpp:184-194); the facade derives
 * them from the orientation quaternion and odometry::FootContact. */
typedef struct eslam_gpu_step {
    double yaw;            /* base::getYaw(orientation) :186 */
    double z_delta;        /* (orientation * poseDelta).translation().z() :188-189 */
    double z_var;          /* odometry.getPositionError()(2,2) * 2 :192 */
    double delta_mean[3];  /* odometry pose delta mean (x, y, theta) */
    double delta_sigma[3]; /* std-dev of odometry.getPoseDeltaSample2D() :198 */
} eslam_gpu_step;

/* odometry::BodyContactPoint after ContactModel::setContactPoints (src/ContactModel.cpp:21-41):
 * position already rotated by removeYaw(orientation). */
typedef struct eslam_gpu_contact {
    double position[3];
    float contact;         /* contact probability; NaN = unknown (passes the 0.2 gate) */
    float slip;
    int32_t group_id;      /* -1 = ungrouped */
    int32_t pad;
} eslam_gpu_contact;

/* What PoseEstimator::update leaves behind (src/PoseEstimator.cpp:244-255, 329-351). */
typedef struct eslam_gpu_update_result {
    double ess;              /* normalizeWeights() return value */
    double max_weight;       /* max contact weight, carried into the next project (Q10) */
    double floating_weight;  /* :329 */
    double weights_sum;      /* getWeightsSum() before normalising */
    uint64_t data_particles; /* :307 */
    uint64_t total_points;   /* :310 */
    int32_t resampled;       /* eff < minEffective */
    int32_t status;          /* ESLAM_OK or ESLAM_ERR_ZERO_VARIANCE */
} eslam_gpu_update_result;

typedef struct eslam_gpu_ctx eslam_gpu_ctx;

void eslam_gpu_default_config(eslam_gpu_config* cfg);                 /* src/Configuration.hpp:53-59,85-111 */
int eslam_gpu_create(const eslam_gpu_config* cfg, int device, const eslam_gpu_shard* shard_or_null,
                     eslam_gpu_ctx** out);                            /* PoseEstimator ctor, src/PoseEstimator.cpp:13-25 */
/* Sharded filter (one context per GPU / rank of ONE box; shards are equal, 1024 * 2^k particles each): rank 0 creates a
 * NCCL unique id, the launcher broadcasts the 128 bytes, every rank joins.  The library resolves NCCL at run time (dlopen of
 * the libnccl.so.2 already in the process) and uses it for the bootstrap only -- the all-gather of the CUDA IPC handles that
 * map every rank's buffers into every process, and the per-tile partials of getCentroid.  The per-step exchanges (loop-A
 * partials, weight sums, ESS, the exact running sum's per-rank summary, "my running sums are final") are O(1) per rank and run
 * INSIDE the step's kernels over NVLink peer memory: the last CTA of the producing kernel pushes the rank's summary into every
 * rank's exchange buffer and releases an epoch flag, the consuming kernel's CTAs poll their own flags.  Resampling reads the
 * ancestors (state, running sums, map patches) in place from the owning rank.  Peer access between the GPUs is required. */
int eslam_gpu_nccl_unique_id(uint8_t out[128]);
int eslam_gpu_comm_init(eslam_gpu_ctx* ctx, const uint8_t unique_id[128]);
void eslam_gpu_destroy(eslam_gpu_ctx* ctx);
const char* eslam_gpu_last_error(const eslam_gpu_ctx* ctx);

/* PoseEstimator::setEnvironment + GridAccess::setMap (src/PoseEstimator.cpp:47-62, src/PoseEstimator.hpp:68-77):
 * dense elevation grid, cells[iy*width+ix] = {mean, stdev}, NaN mean = empty cell;
 * cell index = floor((p - offset)/scale); g2l = row-major 3x4 C_global2local (NULL = identity). */
int eslam_gpu_set_grid(eslam_gpu_ctx* ctx, const float* cells_mean_stdev, int width, int height,
                       double scale_x, double scale_y, double offset_x, double offset_y, const double* g2l);

/* PoseEstimator::init(n, mu, sigma, zpos, zsigma) (src/PoseEstimator.cpp:88-102). n = local shard size. 
 * Sharded contexts: re-initialising after steps restarts the exchange epochs; every rank must have finished its steps before
 * any rank calls this again, and no rank may step before all have (a barrier on either side of the call). */
int eslam_gpu_init_gaussian(eslam_gpu_ctx* ctx, uint64_t n, const double mu[3], const double sigma[3],
                            double zpos, double zsigma);
/* ParticleFilter::getParticles() (src/ParticleFilter.hpp:150-158) as SoA; NULL pointers are skipped. */
int eslam_gpu_set_particles(eslam_gpu_ctx* ctx, uint64_t n, const double* x, const double* y, const double* theta,
                            const double* zpos, const double* zsigma, const double* weight);
int eslam_gpu_get_particles(eslam_gpu_ctx* ctx, double* x, double* y, double* theta, double* zpos, double* zsigma,
                            double* weight, double* mprob, uint8_t* floating, uint8_t* n_found);
uint64_t eslam_gpu_particle_count(const eslam_gpu_ctx* ctx);

/* PoseEstimator::project (src/PoseEstimator.cpp:184-242). */
int eslam_gpu_project(eslam_gpu_ctx* ctx, const eslam_gpu_step* step);
/* PoseEstimator::update = updateWeights + normalizeWeights + conditional resample
 * (src/PoseEstimator.cpp:244-352).  result may be NULL (no host sync). */
int eslam_gpu_update(eslam_gpu_ctx* ctx, const eslam_gpu_contact* pts, int n_pts, eslam_gpu_update_result* result);
/* project + update as one fused launch sequence (EmbodiedSlamFilter::update(contact),
 * src/EmbodiedSlamFilter.cpp:353-369 with the measurement threshold met). */
int eslam_gpu_step_fused(eslam_gpu_ctx* ctx, const eslam_gpu_step* step, const eslam_gpu_contact* pts, int n_pts,
                         eslam_gpu_update_result* result);
/* Enqueue `steps` identical fused steps without host synchronisation (benchmark loop; inputs stay
 * resident).  elapsed_ms (may be NULL) = CUDA-event time of the whole region on the context's stream;
 * top_kernel_ms = summed event time of the dominant kernel alone (k_weight_lean; k_project_weight when the lean path
 * does not apply), from events around it on every step; they cost the programmatic overlap at two places per step, so a
 * run that measures the step itself passes NULL: no event then sits between its first and its last launch. */
int eslam_gpu_run_steps(eslam_gpu_ctx* ctx, const eslam_gpu_step* step, const eslam_gpu_contact* pts, int n_pts,
                        int steps, float* elapsed_ms, float* top_kernel_ms, eslam_gpu_update_result* last_result);
/* Step periods of the last eslam_gpu_run_steps that asked for top_kernel_ms (steps - 1 values, CUDA events): copies min(count, cap) values, returns count. */
int eslam_gpu_step_times(eslam_gpu_ctx* ctx, float* out_ms, int cap);
/* Bytes one eslam_gpu_step_fused call moves between host and device: out = {host -> device (step + contact array, as kernel
 * parameters), device -> host (the scalar block behind eslam_gpu_update_result)} (bench "e2e"). */
int eslam_gpu_io_bytes(const eslam_gpu_ctx* ctx, int n_pts, uint64_t out[2]);

/* ParticleFilter::getWeightsSum (src/ParticleFilter.hpp:34-39). */
int eslam_gpu_weights_sum(eslam_gpu_ctx* ctx, double* sum);
/* ParticleFilter::normalizeWeights (src/ParticleFilter.hpp:46-70); returns ESS. */
int eslam_gpu_normalize_weights(eslam_gpu_ctx* ctx, double* ess);
/* ParticleFilter::resample -> resample_stratified(N) (src/ParticleFilter.hpp:72-108). */
int eslam_gpu_resample(eslam_gpu_ctx* ctx);
/* ParticleFilter::resample_stratified(samples) / resample_multinomial(samples) (src/ParticleFilter.hpp:85-147) with a sample
 * count of their own: the set has `samples` particles afterwards (at most the capacity the context was initialised with; not
 * on sharded contexts).  multinomial != 0: one uniform per sample, the first particle whose running sum reaches it, weights
 * reset to 1 / old size (:140); a draw above the total weight takes the last particle (the reference drops that sample). */
int eslam_gpu_resample_n(eslam_gpu_ctx* ctx, uint64_t samples, int multinomial);
/* parity taps: ancestor index of every particle after the last resample; the exact running-sum CDF. */
int eslam_gpu_get_ancestors(eslam_gpu_ctx* ctx, uint32_t* ancestors);
int eslam_gpu_get_cdf(eslam_gpu_ctx* ctx, double* cdf);

/* PoseEstimator::getCentroid (src/PoseEstimator.cpp:354-383): out = {x, y, theta, z} weighted means. */
int eslam_gpu_centroid(eslam_gpu_ctx* ctx, double out[4]);
/* ParticleFilter::getBestParticleIndex (src/ParticleFilter.hpp:160-173). */
int eslam_gpu_best_index(eslam_gpu_ctx* ctx, uint64_t* index);

/* ContactModel::evaluatePose + evaluateWeight + updateZPositionEstimate for explicit poses
 * (src/ContactModel.cpp:117-224,262-340) -- the seam the reference's testContactModel.cpp drives.
 * All output pointers may be NULL.  cell_index: n*n_pts flat cell indices (iy*width+ix), -1 = not looked up / outside.
 * cp: n*ESLAM_MAX_CP*5 doubles (point x, y, z, zdiff, zvar) of the fused contact points. */
int eslam_gpu_evaluate_poses(eslam_gpu_ctx* ctx, uint64_t n, const double* x, const double* y, const double* theta,
                             const double* zpos, const double* meas_var, const double* z_var_in,
                             const eslam_gpu_contact* pts, int n_pts,
                             uint8_t* ok, double* weight, double* z_delta, double* z_var, double* pose_var,
                             int32_t* n_contacts, double* zpos_out, double* z_var_out,
                             int32_t* cell_index, double* cp);

/* EmbodiedSlamFilter per-particle maps (useSharedMap == false, src/EmbodiedSlamFilter.cpp:131-134): with cfg.patch_size = P
 * every particle owns a P x P overlay of the shared grid.  update() writes the fused contact heights of each particle into
 * its patch (north_star's narrowing of processMap, src/EmbodiedSlamFilter.cpp:209-227), lookups consult the patch first,
 * resampling copies it with the particle (cloneMaps, src/PoseEstimator.cpp:31-45).  Read-back: [n][P][P]{mean, stdev}
 * cells and [n][2] grid coordinates of each patch's cell (0, 0) (INT32_MIN: not anchored yet). */
int eslam_gpu_get_patches(eslam_gpu_ctx* ctx, float* cells_mean_stdev, int32_t* origin_xy);

/* EmbodiedSlamFilter::processMap(scanMap, match, update) (src/EmbodiedSlamFilter.cpp:179-232) in per-particle-map mode
 * (cfg.patch_size > 0).  The scan is what envire's MLSProjection leaves in scanMap (:137-176, out of scope): a dense grid in the
 * yaw-compensated body frame, cell (i, j) centred on offset + (i + 0.5, j + 0.5) * scale, {mean, stdev}, NaN = no surface.
 * Per particle: C_s2p = Translation(x, y, 0) * AngleAxis(theta, Z) (:190-193,:209), offsetPatch(zPos, zSigma) (:212);
 *   match  != 0: weight = MLSGrid::match(scan, C_s2p, offset, sampling = 10, sigma = 0.2); p.weight *= pow(weight, 0.1) (:214-221)
 *   update != 0: MLSGrid::merge(scan, C_s2p, offset) into the particle's own patch (:222-227).
 * envire's match / merge are not vendored: include/eslam_gpu.h's definition is the one of oracle/eslam_oracle.cpp::process_map
 * (merge: every scan surface, row-major, Kalman-merged in float into the patch cell its centre falls in; match: 10 evenly
 * strided scan surfaces looked up in the particle's map -- patch first, then the shared grid, 3-sigma gate -- weight = mean of
 * exp(-d^2 / (2 sigma^2)) over the hits, 1 without a hit). */
typedef struct eslam_gpu_scan {
    const float* cells_mean_stdev; int32_t width, height;
    double scale_x, scale_y, offset_x, offset_y;
} eslam_gpu_scan;
int eslam_gpu_process_map(eslam_gpu_ctx* ctx, const eslam_gpu_scan* scan, int match, int update);

/* The mixture part of eslam::PoseDistribution (src/PoseParticle.hpp:88-114; the Rock task around the library logs particles +
 * a Gaussian mixture over the poses): weighted mean and covariance of the particles' (x, y) as ONE Gaussian, reduced on the
 * device over all shards (pairwise trees: the result does not depend on the number of GPUs).  cov is row major. */
typedef struct eslam_gpu_gaussian2d { double weight; double mean[2]; double cov[4]; } eslam_gpu_gaussian2d;
int eslam_gpu_pose_moments(eslam_gpu_ctx* ctx, eslam_gpu_gaussian2d* out);

/* SurfaceHash, the relocaliser (src/SurfaceHash.hpp:113-232): for every grid cell and each of `angular_steps` headings the
 * least-squares plane under a 4-point footprint gives (slope_x, slope_y); candidate poses are bucketed by slope
 * (slope_bins x slope_bins over [-1, 1]^2).  Built on the device from the attached grid (identity map frame).
 *   eslam_gpu_hash_create       SurfaceHash::create (:155-231)
 *   eslam_gpu_hash_info         size of the table, candidates per bucket (slope_bins^2 counts; either pointer may be NULL)
 *   eslam_gpu_hash_get          parity tap: candidates [first, first + count) in creation order (heading, row, column)
 *   eslam_gpu_init_from_hash    PoseEstimator::init(numParticles, hash) (src/PoseEstimator.cpp:75-86); ESLAM_ERR_EMPTY =
 *                               "could not sample from pose hash." (:84)
 *   eslam_gpu_sample_from_hash  PoseEstimator::sampleFromHash (src/PoseEstimator.cpp:130-182), which project() calls every
 *                               hash period (:239-241): the replace_percentage * relevance^3 lowest-weight particles become
 *                               candidates whose slopes match the plane through the lowest contact point of every group;
 *                               their weight is getWeightAvg() * avg_factor * relevance^3, zSigma 0.5.  pts as for update().
 * The reference draws with libc rand(); here the draws come from the Philox stream (noise is unpinned either way). */
int eslam_gpu_hash_create(eslam_gpu_ctx* ctx, int slope_bins, int angular_steps);
int eslam_gpu_hash_info(eslam_gpu_ctx* ctx, uint64_t* n_poses, uint32_t* bucket_counts);
int eslam_gpu_hash_get(eslam_gpu_ctx* ctx, uint64_t first, uint64_t count, double* x, double* y, double* theta, double* z, int32_t* bucket);
int eslam_gpu_init_from_hash(eslam_gpu_ctx* ctx, uint64_t n);
int eslam_gpu_sample_from_hash(eslam_gpu_ctx* ctx, double replace_percentage, double avg_factor, const eslam_gpu_contact* pts, int n_pts,
                               uint64_t* replaced);

int eslam_gpu_synchronize(eslam_gpu_ctx* ctx);
/* diagnostics of the exact running-sum machinery: out = {sequential-fallback runs, boundary elements of the last
 * resample, pending error flags, resamples done} */
int eslam_gpu_debug_counters(eslam_gpu_ctx* ctx, uint64_t out[4]);
/* which weight kernel did the work in the last update: out = {particles binned for the lean kernel (0: the update
 * ran on the generic kernel alone), particles evaluated by the generic kernel (map border, out-of-range values)} */
int eslam_gpu_path_counters(eslam_gpu_ctx* ctx, uint64_t out[2]);
/* number of kernels this library has launched on ctx since creation (bench "gpu_launches") */
uint64_t eslam_gpu_launch_count(const eslam_gpu_ctx* ctx);

#ifdef __cplusplus
}
#endif
#endif /* ESLAM_GPU_H */
