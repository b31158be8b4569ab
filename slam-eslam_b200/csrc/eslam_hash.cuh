// eslam_hash.cuh -- SurfaceHash on the device (src/SurfaceHash.hpp:113-232, src/PoseEstimator.cpp:75-86,130-182): the slope-indexed
// table of candidate poses the filter is seeded from (init(n, hash)) and re-seeded with every `period` steps (sampleFromHash).
// Off the per-step hot path: built once per map, consulted every `period` steps.  Sorting and scanning use CUB (the library
// that ships with the toolkit); the candidate generation, seeding and replacement kernels are below.
//
// Definition shared with the oracle (oracle/eslam_oracle.cpp hash_create / init_from_hash / sample_from_hash): candidate order =
// heading, grid row, column; the footprint is rotated by (a + 1) steps while the stored heading is a steps (as the reference
// does, :177-181); a footprint point outside the grid is skipped; draws come from the Philox stream STREAM_HASH.
#pragma once
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "eslam_kernels.cuh"

struct HashPose { unsigned short m, n; unsigned char a, pad; unsigned short bucket; double z; };   // 16 bytes

struct HashArgs { MapParams map; int bins, steps; };

// the candidate of (heading a, cell m, n), if the footprint sees at least 3 surfaces and they span a plane
__device__ __forceinline__ bool hash_candidate(const HashArgs& A, const double* rx, const double* ry, const double* fx, const double* fy,
                                               int a, int m, int n, HashPose& hp) {
    const MapParams& M = A.map;
    const double x = M.ox + (m + 0.5) * M.sx, y = M.oy + (n + 0.5) * M.sy;                  // fromGrid :189
    double gx[4], gy[4], gz[4], mean_z = 0; int cnt = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {                                                            // :193-206
        int ix, iy;
        if (!em::cell_index_exact(x + rx[i], M.ox, M.sx, M.W, ix) || !em::cell_index_exact(y + ry[i], M.oy, M.sy, M.H, iy)) continue;
        const float2 cell = __ldg(&M.cells[(size_t)iy * M.W + ix]);
        if (!(cell.x == cell.x) || !(cell.y == cell.y)) continue;
        gx[cnt] = fx[i]; gy[cnt] = fy[i]; gz[cnt] = (double)cell.x; mean_z += (double)cell.x; cnt++;
    }
    if (cnt < 3) return false;                                                               // :209
    mean_z /= cnt;
    double slx, sly;
    if (!em::surface_slopes(gx, gy, gz, cnt, slx, sly)) return false;                        // :211-212
    hp.m = (unsigned short)m; hp.n = (unsigned short)n; hp.a = (unsigned char)a; hp.pad = 0;
    hp.bucket = (unsigned short)(em::slope_bucket(slx, A.bins) * A.bins + em::slope_bucket(sly, A.bins));   // :218
    hp.z = mean_z + 0.18;                                                                    // :216
    return true;
}

// one CTA per (heading, grid row): WRITE == false counts the row's candidates, WRITE == true writes them behind the row's offset
// in column order (block-wide ranks: ballots inside the warp, warp totals through shared memory, a carry from chunk to chunk)
template <bool WRITE>
__global__ void __launch_bounds__(256) k_hash_rows(const HashArgs A, unsigned* row_count, const unsigned* row_offs, HashPose* out, unsigned* bucket_count) {
    __shared__ unsigned wsum[8];
    const int row = blockIdx.x, a = row / A.map.H, n = row - a * A.map.H;
    const double rot = (a + 1) * 2.0 * M_PI / A.steps;                                       // :177-181 (rotated before use)
    double s, c; em::sincos_(rot, s, c);
    double rx[4], ry[4], fx[4], fy[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) { em::hash_footprint(i, fx[i], fy[i]); rx[i] = c * fx[i] - s * fy[i]; ry[i] = s * fx[i] + c * fy[i]; }
    const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned carry = 0;
    for (int m0 = 0; m0 < A.map.W; m0 += 256) {
        const int m = m0 + (int)threadIdx.x;
        HashPose hp; bool ok = false;
        if (m < A.map.W) ok = hash_candidate(A, rx, ry, fx, fy, a, m, n, hp);
        const unsigned bal = __ballot_sync(0xffffffffu, ok);
        if (lane == 0) wsum[warp] = __popc(bal);
        __syncthreads();
        unsigned before = 0, total = 0;
        for (int w = 0; w < 8; ++w) { if (w < (int)warp) before += wsum[w]; total += wsum[w]; }
        if (WRITE && ok) {
            out[(size_t)row_offs[row] + carry + before + __popc(bal & ((1u << lane) - 1u))] = hp;
            atomicAdd(&bucket_count[hp.bucket], 1u);
        }
        carry += total;
        __syncthreads();
    }
    if (!WRITE && threadIdx.x == 0) row_count[row] = carry;
}

__global__ void k_hash_keys(const HashPose* poses, unsigned n, unsigned short* keys, unsigned* idx) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = poses[i].bucket; idx[i] = i; }
}

__device__ __forceinline__ void hash_pose_xyt(const MapParams& M, int steps, const HashPose& hp, double& x, double& y, double& th) {
    x = M.ox + (hp.m + 0.5) * M.sx; y = M.oy + (hp.n + 0.5) * M.sy;
    th = hp.a * 2.0 * M_PI / steps;                                                          // :175
}

// PoseEstimator::init(numParticles, hash) (src/PoseEstimator.cpp:75-86): every particle a uniformly drawn candidate
__global__ void k_init_from_hash(StateBufs st, const DevScalars* sc, unsigned long long n, unsigned long long base, uint32_t k0, uint32_t k1,
                                 const HashPose* poses, unsigned n_poses, MapParams M, int steps) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cur = sc->cur;
    const unsigned long long gi = base + i;
    const em::Philox4 r = em::philox4x32_10((uint32_t)gi, 0xFFFFFFFFu, 0, stream_word(em::STREAM_HASH, gi), k0, k1);
    unsigned idx = (unsigned)(em::u01(r.v[0], r.v[1]) * (double)n_poses);
    if (idx >= n_poses) idx = n_poses - 1;
    double x, y, th; hash_pose_xyt(M, steps, poses[idx], x, y, th);
    st.f[cur][F_X][i] = x; st.f[cur][F_Y][i] = y; st.f[cur][F_TH][i] = th;
    st.f[cur][F_Z][i] = poses[idx].z; st.f[cur][F_ZS][i] = 0.0;                              // PoseParticle(position, orientation, zpos): zSigma 0
    st.f[cur][F_W][i] = 0.0; st.f[cur][F_MP][i] = 0.0;                                       // weight(0): src/PoseParticle.hpp:56
    st.b[cur][B_FLOAT][i] = 1; st.b[cur][B_NF][i] = 0;
}

__global__ void k_weight_keys(StateBufs st, const DevScalars* sc, unsigned n, float* keys, unsigned* idx) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) { keys[i] = (float)st.f[sc->cur][F_W][i]; idx[i] = i; }                       // weight_index(float, index) :147-151
}

// sampleFromHash :164-180: the replace_count lowest-weight particles become candidates of the matching bucket
__global__ void k_hash_replace(StateBufs st, const DevScalars* sc, const unsigned* sorted_idx, unsigned replace_count, uint32_t k0, uint32_t k1,
                               uint32_t hash_step, const HashPose* poses, const unsigned* by_bucket, unsigned bucket_first, unsigned in_bucket,
                               MapParams M, int steps, double weight) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= replace_count) return;
    const int cur = sc->cur;
    const em::Philox4 r = em::philox4x32_10(i, hash_step, 1, stream_word(em::STREAM_HASH, i), k0, k1);
    unsigned k = (unsigned)(em::u01(r.v[0], r.v[1]) * (double)in_bucket);
    if (k >= in_bucket) k = in_bucket - 1;
    const HashPose hp = poses[by_bucket[bucket_first + k]];
    const unsigned t = sorted_idx[i];
    double x, y, th; hash_pose_xyt(M, steps, hp, x, y, th);
    st.f[cur][F_X][t] = x; st.f[cur][F_Y][t] = y; st.f[cur][F_TH][t] = th;                  // :170-175
    st.f[cur][F_Z][t] = hp.z; st.f[cur][F_ZS][t] = 0.5; st.f[cur][F_W][t] = weight;
    st.b[cur][B_FLOAT][t] = 1;
}
