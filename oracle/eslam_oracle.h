/* eslam_oracle.h -- C API of the CPU oracle.  TEST INFRASTRUCTURE ONLY: nothing under
 * slam-eslam_b200/, include/ or the shipped library may include, link or call this.  Only
 * tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs use it.
 *
 * The oracle is a single-threaded CPU restatement of the reference's particle-filter update
 * path (liyangSKD/slam-eslam: src/ParticleFilter.hpp, src/PoseEstimator.cpp, src/ContactModel.cpp),
 * reading its noise from Philox counters instead of boost::minstd_rand (north_star: "identical
 * noise streams").  It has two arithmetic policies:
 *   ESLAM_ORACLE_LITERAL (0): the reference's expression order, libm exp/erfc/sin/cos/pow/log,
 *                             left-to-right sums -- the line-by-line restatement;
 *   ESLAM_ORACLE_SHARED  (1): the same algorithm with the arithmetic of slam-eslam_b200/csrc/eslam_math.h
 *                             (explicit fma, table-driven pdf/cdf, pairwise sums), which the CUDA kernels
 *                             reproduce bit-for-bit.  tests/ bound LITERAL vs SHARED by tolerance and
 *                             SHARED vs GPU bit-exactly.
 * PARITY PINNING: the LITERAL policy is pinned against the reference's OWN compiled code.  The reference as a whole
 * cannot be built here (Rock CMake, boost, Eigen, envire, odometry are absent), but its particle-filter
 * sources compile unmodified against stand-in dependency headers (oracle/ref_shim, oracle/ref/Makefile ->
 * oracle/_ref/libeslam_ref.so), and tests/test_reference_parity.py requires this oracle to reproduce that library BIT FOR
 * BIT -- every double of every particle, ESS, flags, ancestors over whole trajectories (live, and from the committed
 * outputs tests/golden/reference_trajectories.npz where /root/reference is absent), the ContactModel seam with its map-cell
 * indices, resampling on adversarial weights -- next to the known-answer tests of test/testContactModel.cpp
 * (tests/golden/contact_model_kat.json).  What stays defined by this repo rather than pinned is the behaviour of the
 * un-vendored dependencies (dense-grid map semantics, odometry sampler, Philox noise): oracle/ref_shim/README.md.
 */
#ifndef ESLAM_ORACLE_H
#define ESLAM_ORACLE_H
#include "../include/eslam_gpu.h"   /* POD structs only */

#ifdef __cplusplus
extern "C" {
#endif

enum { ESLAM_ORACLE_LITERAL = 0, ESLAM_ORACLE_SHARED = 1 };

typedef struct eslam_oracle_ctx eslam_oracle_ctx;

int eslam_oracle_create(const eslam_gpu_config* cfg, int policy, const eslam_gpu_shard* shard_or_null, eslam_oracle_ctx** out);
void eslam_oracle_destroy(eslam_oracle_ctx* ctx);
int eslam_oracle_set_grid(eslam_oracle_ctx* ctx, const float* cells_mean_stdev, int width, int height,
                          double scale_x, double scale_y, double offset_x, double offset_y, const double* g2l);
int eslam_oracle_init_gaussian(eslam_oracle_ctx* ctx, uint64_t n, const double mu[3], const double sigma[3],
                               double zpos, double zsigma);
int eslam_oracle_set_particles(eslam_oracle_ctx* ctx, uint64_t n, const double* x, const double* y, const double* theta,
                               const double* zpos, const double* zsigma, const double* weight);
int eslam_oracle_get_particles(eslam_oracle_ctx* ctx, double* x, double* y, double* theta, double* zpos, double* zsigma,
                               double* weight, double* mprob, uint8_t* floating, uint8_t* n_found);
uint64_t eslam_oracle_particle_count(const eslam_oracle_ctx* ctx);
int eslam_oracle_project(eslam_oracle_ctx* ctx, const eslam_gpu_step* step);
int eslam_oracle_update(eslam_oracle_ctx* ctx, const eslam_gpu_contact* pts, int n_pts, eslam_gpu_update_result* result);
int eslam_oracle_normalize_weights(eslam_oracle_ctx* ctx, double* ess);
int eslam_oracle_resample(eslam_oracle_ctx* ctx);
int eslam_oracle_resample_n(eslam_oracle_ctx* ctx, uint64_t samples, int multinomial);
int eslam_oracle_get_ancestors(eslam_oracle_ctx* ctx, uint32_t* ancestors);
int eslam_oracle_get_cdf(eslam_oracle_ctx* ctx, double* cdf);
int eslam_oracle_centroid(eslam_oracle_ctx* ctx, double out[4]);
int eslam_oracle_best_index(eslam_oracle_ctx* ctx, uint64_t* index);
int eslam_oracle_evaluate_poses(eslam_oracle_ctx* ctx, uint64_t n, const double* x, const double* y, const double* theta,
                                const double* zpos, const double* meas_var, const double* z_var_in,
                                const eslam_gpu_contact* pts, int n_pts,
                                uint8_t* ok, double* weight, double* z_delta, double* z_var, double* pose_var,
                                int32_t* n_contacts, double* zpos_out, double* z_var_out,
                                int32_t* cell_index, double* cp);
int eslam_oracle_get_patches(eslam_oracle_ctx* ctx, float* cells_mean_stdev, int32_t* origin_xy);
int eslam_oracle_process_map(eslam_oracle_ctx* ctx, const eslam_gpu_scan* scan, int match, int update);
int eslam_oracle_pose_moments(eslam_oracle_ctx* ctx, eslam_gpu_gaussian2d* out);
int eslam_oracle_hash_create(eslam_oracle_ctx* ctx, int slope_bins, int angular_steps);
int eslam_oracle_hash_info(eslam_oracle_ctx* ctx, uint64_t* n_poses, uint32_t* bucket_counts);
int eslam_oracle_hash_get(eslam_oracle_ctx* ctx, uint64_t first, uint64_t count, double* x, double* y, double* theta, double* z, int32_t* bucket);
int eslam_oracle_init_from_hash(eslam_oracle_ctx* ctx, uint64_t n);
int eslam_oracle_sample_from_hash(eslam_oracle_ctx* ctx, double replace_percentage, double avg_factor, const eslam_gpu_contact* pts, int n_pts,
                                  uint64_t* replaced);

/* FakeMLSAccess of test/testContactModel.cpp:8-38 as the map: 4 quadrants, index = 2*(y>0) + (x>0);
 * res may be NULL (all present).  Replaces the grid for evaluate_poses; queried points are recorded. */
int eslam_oracle_set_fake_map(eslam_oracle_ctx* ctx, const double z[4], const double stddev[4], const uint8_t* res);
int eslam_oracle_get_queried_points(eslam_oracle_ctx* ctx, double* xyz, int max_points);

/* ContactModel::getLowestPointPerGroup / updateContactStateUsingLowestPointHeuristic
 * (src/ContactModel.cpp:48-92); host-side helper pinned by the reference's KATs.
 * lowest_xyz: up to n_pts*3 doubles, returns the count; contact_out (may be NULL): rewritten probabilities. */
int eslam_oracle_lowest_points(const eslam_gpu_contact* pts, int n_pts, int update_probabilities,
                               double* lowest_xyz, float* contact_out);

/* math probes for tests (shared policy primitives vs libm) */
double eslam_oracle_math(int fn, double a, double b);

#ifdef __cplusplus
}
#endif
#endif
