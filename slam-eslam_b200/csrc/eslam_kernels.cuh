// eslam_kernels.cuh -- sm_100a kernels of the eSLAM particle-filter update path.
//
// Reference functions replaced (paths relative to liyangSKD/slam-eslam):
//   k_init_gaussian      PoseEstimator::init                      src/PoseEstimator.cpp:88-102
//   k_project_weight     PoseEstimator::project                   src/PoseEstimator.cpp:184-242
//                        + updateWeights loop A                   src/PoseEstimator.cpp:276-327
//                        + ContactModel::evaluatePose/evaluateWeight/updateZPositionEstimate
//                                                                 src/ContactModel.cpp:117-224,262-340
//                        + GridAccess::get                        src/PoseEstimator.hpp:97-105
//   k_reduce_a           floating_weight, max_weight              src/PoseEstimator.cpp:329,347-348
//   k_loop_b             updateWeights loop B                     src/PoseEstimator.cpp:332-345
//   k_reduce_b,k_norm_scan,k_chain   ParticleFilter::normalizeWeights  src/ParticleFilter.hpp:46-70
//   k_norm_scan,k_chain,k_cdf        running sum of resample_stratified src/ParticleFilter.hpp:95,102
//   k_search             the stratified walk                      src/ParticleFilter.hpp:96-105
//   k_gather             xi_kp.push_back(xi_k[idx]); swap         src/ParticleFilter.hpp:104-107
//   k_centroid,k_best    getCentroid / getBestParticleIndex       src/PoseEstimator.cpp:354-383, src/ParticleFilter.hpp:160-173
//
// Arithmetic is that of eslam_math.h (compile with -fmad=false); all global sums are pairwise over the
// index-aligned binary tree, so results do not depend on grid size, SM count or number of GPUs.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>
#include "eslam_math.h"
#include "../../include/eslam_gpu.h"

namespace em = eslam_math;

#define TILE 1024                // particles per reduction tile (fixed: defines the summation tree)
#define WT_THREADS 1024          // threads of the generic project+weight kernel
#define MAX_BND_TOTAL 2048       // boundary elements per resample (pool + chain capacity); more -> sequential fallback
#define BIN_NONE (-100000)       // "no binade": piece made of zero / sub-quarter-ulp weights only
#define ESLAM_MAX_RANKS 8        // GPUs of one box
#define ESLAM_MAX_SHARDS 64      // logical shards (ranks, or virtual shards of one context)
#define ERR_ZERO_VARIANCE_BIT 1
#define ERR_CDF_FALLBACK_BIT 2

struct StateBufs {               // SoA particle state, double buffered (src/PoseParticle.hpp:52-86 hot fields)
    double* f[2][7];             // x, y, theta, zPos, zSigma, weight, mprob
    uint8_t* b[2][2];            // floating, n_found (= cpoints.size())
};
enum { F_X = 0, F_Y, F_TH, F_Z, F_ZS, F_W, F_MP };
enum { B_FLOAT = 0, B_NF };

struct DevScalars {              // device-resident scalars of one context
    double max_weight;           // PoseEstimator::max_weight, carried to the next project (Q10)
    double floating_weight;
    double sum_w;                // getWeightsSum()
    double ess;
    double pow_table[6];         // pow(discount*floating_weight, 4 - n), n = 0..4, [5] = wrapped exponent (Q2)
    double tile_cx, tile_cy;     // world position the shared-memory map tile is centred on
    double centroid[5];          // sum x w, y w, theta w, z w, w
    unsigned long long data_particles, total_points;
    unsigned long long best_index;
    unsigned int resample_step;
    int cur;                     // which state buffer is current
    int do_resample;
    int uniform;                 // sum of weights <= 0 -> 1/N (Q8)
    int error_flags;
    int n_boundaries;
    int pool_count;              // boundary-element pool allocator (reset by k_reduce_b)
    unsigned int fallbacks;      // times the sequential CDF fallback ran
    unsigned int defer_count;    // particles the binning flagged for contact_eval (map border, out-of-range values; live counter)
    unsigned int defer_last;     // its value at the end of the last update (diagnostics)
    int migrate;                 // LOCAL_MIGRATE sharding: 1 = this step resamples exact-globally (decided in k_norm_scan)
    int cdf_seq;                 // the running sum of this resample was computed sequentially by k_chain (k_cdf skips)
    unsigned int ticket[4];      // "last CTA" counters: 0 k_tile_part, 1 k_loop_b, 2 k_norm_scan, 3 k_cdf
    unsigned int x_epoch[8];     // epoch of every exchange (X_*): bumped by this rank's producer, read by its consumers
    unsigned int project_step;   // Philox counter word of project: projects done so far (device side: a captured step is replayable)
    int rank_start_valid;        // rank_start[] below holds the exact running sum in front of every rank (k_chain, exact-global walk)
    double rank_start[9];        // [q] = running sum before rank q's first particle, [world] = the total (ESLAM_MAX_RANKS + 1 entries)
};

// The UpdateResult of a step, written by k_chain into host-mapped pinned memory as soon as the resampling decision is taken
// (the host reads it without a copy or a stream synchronise while k_cdf / the resampling kernel still run); seq last.
struct ResultBox {
    double ess, max_weight, floating_weight, sum_w;
    unsigned long long data_particles, total_points;
    int do_resample, error_flags;
    unsigned long long seq;      // number of k_chain launches of this context that have published
};

struct TilePartA {               // per-tile partials of loop A
    double sum_pw, max_w;
    unsigned int cnt, pts;
};

struct ContactSet {              // odometry::BodyContactState after setContactPoints, as kernel parameters
    int n;
    int pad;
    double px[ESLAM_MAX_CONTACTS], py[ESLAM_MAX_CONTACTS], pzr[ESLAM_MAX_CONTACTS];   // pzr = z - contactPointRadius
    uint8_t flags[ESLAM_MAX_CONTACTS];   // bit0: passes !(contact < 0.2); bit1: last point of its group
};

struct MapParams {
    const float2* cells; int W, H;
    double sx, sy, ox, oy, inv_sx, inv_sy;
    double c0x, c0y;             // 1.5*2^34 - offset*inv_scale (fixed-point cell index, see contact_eval)
    double g2l[8];               // rows 0 and 1 of C_global2local
    int identity;
    int tile_w, tile_h;          // shared-memory tile (cells)
    int tile_x0, tile_y0;        // its origin (inside the grid, x0 even): a placement choice of the host, any value
                                 // gives the same results
};

struct ModelParams {             // eslam::Configuration hot fields
    uint32_t k0, k1;
    double slip_factor, max_yaw_dev, spread_thr, spread_tf, spread_rf;
    double merr2, inv_corr, discount;
    unsigned long long min_contacts;
    int use_shape, use_hash;
};

// The lean weight kernel's view of one update: only the contact points that pass !(contact < 0.2)
// (src/ContactModel.cpp:154), grouped (:194-197), and the integer thresholds of the 3-sigma gate on t = zdiff/(sigma corr).
struct FastContacts {
    int n;                       // == n_eval: with px/py/pzr/flags below the compact list reads like a ContactSet (contact_eval)
    int n_eval, n_groups;
    int margin;                  // cells: every contact of a particle whose own cell is >= margin inside the tile is in the tile
    uint32_t gate_lo, gate_span; // high word of |t|: below gate_lo -> inside the gate; within [gate_lo, gate_lo+gate_span] -> undecided
    double c0x, c0y;             // 1.5*2^28 - offset/scale: cell quotient with 24 fractional bits (minus the tile origin: tile-relative)
    int bsx, bsy, nbx, nby, nb;  // bins: a lattice of tile cores (tile minus the margin on every side), bsx x bsy cells each
    uint32_t inv_bsx, inv_bsy;   // ceil(2^32 / d), d >= 2: n / d == umulhi(n, inv) for n < 2^16 (cell numbers)
    uint32_t inv_nbx;            // ceil(2^31 / nbx) (nbx may be 1): b / nbx == umulhi(2 b, inv) for b < 2^15
    uint8_t gend[ESLAM_MAX_CONTACTS];   // one past the last evaluated contact of each group (groups without any are dropped)
    uint8_t flags[ESLAM_MAX_CONTACTS];  // per compact contact: bit0 = 1 (evaluated), bit1 = last of its group
    double px[ESLAM_MAX_CONTACTS + 1], py[ESLAM_MAX_CONTACTS + 1], pzr[ESLAM_MAX_CONTACTS + 1];   // [n_eval] = 0: read ahead by the loop
};

struct WeightCommon {
    StateBufs st; DevScalars* sc;
    MapParams map; ModelParams mp;
    const em::dbl2* htab;
    unsigned long long n, base;  // local particle count, global index of local particle 0
    int use_tma;                 // the map tile is staged by one TMA bulk tensor copy (tmap valid)
    int pad_tma;
    const CUtensorMap* tmap;     // device copy of the 2-D tensor map of the elevation grid (8-byte cells, box = the tile)
    double* hint;                // [gridDim.x][2]: sums of x and y over the particles a CTA handled (tile placement hint)
    double yaw, z_delta, z_var, dmean[3], dsigma[3];
};

struct FastParams { WeightCommon c; FastContacts fc; };

struct WeightParams {
    StateBufs st; DevScalars* sc;
    MapParams map; ModelParams mp;
    const em::dbl2* htab;
    unsigned long long n, base;  // local particle count, global index of local particle 0
    int bump_project;            // k_tile_part: this update contained a project on the generic kernel -> count it
    int patch_size;              // P (0: no per-particle patches)
    int use_tma;                 // the map tile is staged by one TMA bulk tensor copy (tmap valid)
    const CUtensorMap* tmap;     // device copy of the 2-D tensor map of the elevation grid (8-byte cells, box = the tile)
    float2* patch[2]; int* porigin[2];   // double buffered with the particle state: [N][P][P] cells, [N][2] origins
    double* hint;                // [gridDim.x][2]: sums of x and y over the particles a CTA handled (tile placement hint)
    double yaw, z_delta, z_var, dmean[3], dsigma[3];
    ContactSet cs;
};

struct BndEntry { double w; int idx, b; long long d0, d1; };   // a boundary element and the rounding function of the run after it
struct TileDesc {                // what a tile contributes to the exact running sum (see DESIGN.md "exact CDF")
    int n_bnd, overflow;
    int head_b, pad;
    long long head_d0, head_d1;  // rounding function of the elements before the tile's first boundary element
    int bnd_base, pad2;          // the tile's n_bnd boundary elements live at pool[bnd_base ..]
};

struct SegF { long long d0, d1; int b; int pad; };   // 2-state rounding function of a run of interior elements

// ------------------------------------------------------------------------------------------------ small helpers
// First statement of every kernel of a filter step (see launch_pdl in eslam_gpu.cu): let the next kernel of the stream
// be scheduled early, then wait until everything launched before this kernel has completed and is visible.
__device__ __forceinline__ void pdl_start() {
    asm volatile("griddepcontrol.launch_dependents;");
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// ------------------------------------------------------------------------------------------------ peer exchange
// The sharded filter's per-step exchanges run inside the step's own kernels, over NVLink peer memory (CUDA IPC mappings of
// every rank's exchange buffer; no collective library, no extra launch).  What crosses GPUs is O(1) per rank: because a shard
// holds 2^k tiles, the index-aligned pairwise tree over all G T tiles factors into G per-rank subtrees, so each rank reduces
// its own tiles and publishes the subtree root; the exact running sum likewise travels as one composed rounding function +
// boundary list per rank.  Producer: the LAST CTA of a kernel (ticket) reduces the rank's per-tile partials, PUSHES the
// summary into slot `me` of every rank's buffer (plain stores through the peer mapping, __threadfence_system) and releases
// an epoch flag there (st.release.sys).  Consumer: every CTA of the next kernel polls the flags in its OWN memory
// (ld.acquire.sys) and combines the G summaries in tree order.  Epochs live in DevScalars (device side), so a step has no
// host-supplied exchange state.  On one GPU the same code runs with world == 1 (no flags: the kernel boundary orders).
#define X_A 0        // loop-A rank partials            k_tile_part (last CTA) -> k_loop_b
#define X_B 1        // rank weight sums                k_loop_b (last CTA)    -> k_norm_scan
#define X_C 2        // rank sum of w^2 + chain summary k_chain                -> k_chain (all ranks)
#define X_D 3        // "my running sums are final"     k_resample_local CTA 0 -> the CTAs that read this rank
#define X_SEQ 4      // sequential fallback: S_out      k_chain of rank r      -> k_chain of rank r+1
#define X_BAR 5      // plain rendezvous                k_barrier
#define X_COUNT 6
#define ERR_XCHG_TIMEOUT_BIT 4
#define ERR_INTERNAL_BIT 8

// Payloads travel as "LL words" (the low-latency protocol of collective libraries): 8 bytes = {32 bits of payload, 32-bit
// epoch}.  An aligned 8-byte store is atomic, so a word whose epoch matches carries valid payload: no fence, no separate
// flag, one NVLink write of latency.  Epoch 0 never occurs (buffers start zeroed).  Only X_D / X_BAR, which order bulk data
// written by whole grids, are release/acquire flags.
struct RankPartA { double sum_pw, max_w; unsigned long long cnt, pts; };                     // 8 payload words
struct RankPartB { double sum_w, sum_sq; };                                                  // 4 words (sum_sq: raw weights, LOCAL_MIGRATE decision)
struct ChainEntry { double w; long long d0, d1; int b, pad; };                               // 8 words: a boundary element and the rounding function of the run after it
struct RankChainHead { double sum_sq; int n_bnd, bad; long long head_d0, head_d1; int head_b, head_pad; double pad[3]; };   // 16 words

#define XA_WORDS 8
#define XB_WORDS 4
#define XH_WORDS 16
#define XE_WORDS 8
#define XOFF_A 0                                               // ESLAM_MAX_RANKS slots of XA_WORDS LL words
#define XOFF_B (XOFF_A + ESLAM_MAX_RANKS * XA_WORDS * 8)
#define XOFF_SEQ (XOFF_B + ESLAM_MAX_RANKS * XB_WORDS * 8)
#define XOFF_FLAGS (XOFF_SEQ + ESLAM_MAX_RANKS * 2 * 8)
#define XOFF_C 2048
#define XC_SLOT ((XH_WORDS + MAX_BND_TOTAL * XE_WORDS) * 8)    // header + entries of one rank, as LL words
#define XOFF_TSIN (XOFF_C + ESLAM_MAX_RANKS * XC_SLOT)         // then world x tiles_per_rank doubles

struct Xchg {
    unsigned char* buf[ESLAM_MAX_RANKS];   // rank q's exchange buffer as mapped into this process (q == me: the local one)
    int world, me;
    unsigned tiles_per_rank;               // stride of the global tile-start table at XOFF_TSIN
    unsigned pad;
};

__device__ __forceinline__ unsigned ld_acquire_sys(const unsigned* p) {
    unsigned v; asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_release_sys(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned long long ld_sys_u64(const void* p) {
    unsigned long long v; asm volatile("ld.relaxed.sys.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void st_sys_u64(void* p, unsigned long long v) {
    asm volatile("st.relaxed.sys.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ unsigned* x_flags(unsigned char* buf, int x) { return reinterpret_cast<unsigned*>(buf + XOFF_FLAGS) + x * ESLAM_MAX_RANKS; }

// Producer, any number of threads of one CTA: `words` 32-bit payload words from src (shared / global / local memory) become
// LL words at byte offset `off` of EVERY rank's buffer.  Threads stride over (rank, word).
__device__ __forceinline__ void ll_publish(const Xchg& X, size_t off, const void* src, unsigned words, unsigned epoch) {
    const unsigned* s = reinterpret_cast<const unsigned*>(src);
    const unsigned total = words * (unsigned)X.world;
    for (unsigned i = threadIdx.x; i < total; i += blockDim.x) {
        const unsigned q = i / words, k = i - q * words;
        st_sys_u64(X.buf[q] + off + (size_t)k * 8, ((unsigned long long)epoch << 32) | s[k]);
    }
}
// Consumer, one thread: LL word at p of this GPU's own memory, once its epoch has arrived.  A rank that never arrives (a dead
// peer) would hang the GPU: after ~4 s the wait gives up and flags the step.
__device__ __forceinline__ unsigned ll_read(const unsigned char* p, unsigned epoch, int* error_flags) {
    unsigned long long v = ld_sys_u64(p);
    if ((unsigned)(v >> 32) != epoch) {
        const long long t0 = clock64();
        do {
            __nanosleep(20);
            v = ld_sys_u64(p);
            if (clock64() - t0 > 8000000000ll) { atomicOr(error_flags, ERR_XCHG_TIMEOUT_BIT); break; }
        } while ((unsigned)(v >> 32) != epoch);
    }
    return (unsigned)v;
}
// Consumer, all threads of a CTA: slot `words` payload words of every rank (slots `words` LL words apart at `off` of the own
// buffer) into dst[rank * words + k] (shared memory), then a barrier.
__device__ __forceinline__ void ll_gather(const Xchg& X, size_t off, unsigned words, unsigned epoch, unsigned* dst, int* error_flags) {
    const unsigned total = words * (unsigned)X.world;
    for (unsigned i = threadIdx.x; i < total; i += blockDim.x) dst[i] = ll_read(X.buf[X.me] + off + (size_t)i * 8, epoch, error_flags);
    __syncthreads();
}
// Release / acquire flags for the exchanges that order bulk data (X_D, X_BAR): all threads of ONE CTA call.
__device__ __forceinline__ void x_signal(const Xchg& X, int x, unsigned epoch) {
    if (X.world <= 1) return;
    __threadfence_system();
    __syncthreads();
    if ((int)threadIdx.x < X.world) st_release_sys(x_flags(X.buf[threadIdx.x], x) + X.me, epoch);
}
__device__ __forceinline__ void x_wait(const Xchg& X, int x, unsigned epoch, int* error_flags) {
    if (X.world > 1) {
        const int q = (int)threadIdx.x;
        if (q < X.world) {
            const unsigned* f = x_flags(X.buf[X.me], x) + q;
            const long long t0 = clock64();
            while ((int)(ld_acquire_sys(f) - epoch) < 0) {
                __nanosleep(20);
                if (clock64() - t0 > 8000000000ll) { atomicOr(error_flags, ERR_XCHG_TIMEOUT_BIT); break; }
            }
        }
    }
    __syncthreads();
}
// pairwise tree over m <= 64 values held by one thread: the top levels of the index-aligned tree over all ranks' tiles
__device__ __forceinline__ double tree_small(double* v, int m) {
    while (m > 1) { const int h = (m + 1) >> 1; for (int j = 0; j < h; ++j) v[j] = (2 * j + 1 < m) ? v[2 * j] + v[2 * j + 1] : v[2 * j]; m = h; }
    return v[0];
}
// "am I the last CTA of this grid to get here": what THREAD 0 of the other CTAs wrote before their call (their tile's partial)
// is visible to the last CTA.  Thread 0 takes the ticket with a release/acquire atomic at gpu scope -- no block-wide fence: the
// last CTA only reads what the thread-0s published.
__device__ __forceinline__ bool cta_is_last(unsigned* ticket) {
    __shared__ int s_last;
    if (threadIdx.x == 0) {
        unsigned t;
        asm volatile("atom.add.acq_rel.gpu.global.u32 %0, [%1], 1;" : "=r"(t) : "l"(ticket) : "memory");
        s_last = (t == gridDim.x - 1) ? 1 : 0;
        if (s_last) *ticket = 0;
    }
    __syncthreads();
    return s_last != 0;
}
// rendezvous of all ranks on their streams (eslam_gpu_run_steps: all ranks start the timed region together)
__global__ void k_barrier(DevScalars* sc, const __grid_constant__ Xchg X) {
    __shared__ unsigned s_e;
    if (threadIdx.x == 0) { s_e = sc->x_epoch[X_BAR] + 1; }
    __syncthreads();
    const unsigned e = s_e;
    x_signal(X, X_BAR, e);
    x_wait(X, X_BAR, e, &sc->error_flags);
    if (threadIdx.x == 0) sc->x_epoch[X_BAR] = e;
}

__device__ __forceinline__ uint32_t stream_word(uint32_t stream, unsigned long long idx) {
    return stream ^ ((uint32_t)(idx >> 32) << 8);
}

__device__ __forceinline__ double warp_tree_sum(double v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;   // xor butterfly = balanced binary tree over lane order; every lane holds the total
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}
__device__ __forceinline__ unsigned warp_sum_u(unsigned v) {
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Tree sum over the 1024 leaves of a tile held one per thread (blockDim.x == 1024). red: >= 32 doubles.
__device__ __forceinline__ double tile_tree_sum_1024(double v, double* red) {
    v = warp_tree_sum(v);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = red[threadIdx.x & 31];
    return warp_tree_sum(t);
}

// Pairwise tree over n values in global memory, continuing the per-tile trees; one CTA.  With shared scratch
// (sa: TREE_SMEM doubles, sb: TREE_SMEM/2) and n <= 2*TREE_SMEM all levels but the first run in shared memory;
// otherwise the levels ping-pong between the global arrays a and b.  Same tree either way.
#define TREE_SMEM 2048
#define TREE_SMEM_S 1024         // capacity used by the 256-thread kernels' last-CTA trees (8 + 4 KB of shared memory)
__device__ double block_tree_reduce(double* a, double* b, unsigned n, double* sa = nullptr, double* sb = nullptr, unsigned cap = TREE_SMEM) {
    if (sa && ((n + 1) >> 1) > cap && n <= 16u * cap) {
        // many tiles (sharded runs: all ranks' tiles): the first L levels of the same tree run in registers, 2^L consecutive
        // leaves per thread (a missing right operand passes the left one through, as below), the rest in shared memory
        unsigned L = 1; while ((((n + (1u << L) - 1) >> L) + 1) >> 1 > cap) ++L;
        const unsigned groups = (n + (1u << L) - 1) >> L;                 // <= 2 * TREE_SMEM values left
        for (unsigned g = threadIdx.x; g < ((groups + 1) >> 1); g += blockDim.x) {
            double pair[2]; unsigned have = 0;
            for (unsigned h = 0; h < 2; ++h) {
                const unsigned gi = 2 * g + h;
                if (gi >= groups) break;
                double v[16]; unsigned m = 0;
                const unsigned base = gi << L;
                for (unsigned j = 0; j < (1u << L) && base + j < n; ++j) v[m++] = a[base + j];
                while (m > 1) { const unsigned hh = (m + 1) >> 1; for (unsigned j = 0; j < hh; ++j) v[j] = (2 * j + 1 < m) ? v[2 * j] + v[2 * j + 1] : v[2 * j]; m = hh; }
                pair[have++] = v[0];
            }
            sa[g] = have == 2 ? pair[0] + pair[1] : pair[0];              // one more level: sa holds ceil(groups/2) values
        }
        __syncthreads();
        unsigned m = (groups + 1) >> 1;
        double *x = sa, *y = sb;
        while (m > 1) {
            const unsigned h = (m + 1) >> 1;
            for (unsigned i = threadIdx.x; i < h; i += blockDim.x) y[i] = (2 * i + 1 < m) ? x[2 * i] + x[2 * i + 1] : x[2 * i];
            __syncthreads();
            double* t = x; x = y; y = t; m = h;
        }
        const double r = x[0];
        __syncthreads();
        return r;
    }
    if (sa && n > 1 && ((n + 1) >> 1) <= cap) {
        unsigned h = (n + 1) >> 1;
        for (unsigned i = threadIdx.x; i < h; i += blockDim.x) sa[i] = (2 * i + 1 < n) ? a[2 * i] + a[2 * i + 1] : a[2 * i];
        __syncthreads();
        n = h;
        double *x = sa, *y = sb;
        while (n > 1) {
            h = (n + 1) >> 1;
            for (unsigned i = threadIdx.x; i < h; i += blockDim.x) y[i] = (2 * i + 1 < n) ? x[2 * i] + x[2 * i + 1] : x[2 * i];
            __syncthreads();
            double* t = x; x = y; y = t; n = h;
        }
        const double r = x[0];
        __syncthreads();
        return r;
    }
    while (n > 1) {
        unsigned h = (n + 1) >> 1;
        for (unsigned i = threadIdx.x; i < h; i += blockDim.x)
            b[i] = (2 * i + 1 < n) ? a[2 * i] + a[2 * i + 1] : a[2 * i];
        __syncthreads();
        double* t = a; a = b; b = t; n = h;
    }
    return n ? a[0] : 0.0;
}

#define SCAN_THREADS 256         // scan-type kernels: 256 threads x 4 consecutive elements = one 1024-tile
#define EPT 4

// Tree sum over the 1024 leaves of a tile held 4 consecutive per thread (blockDim.x == 256): same tree as
// tile_tree_sum_1024 (levels 0-1 in the thread, 2-6 by the warp butterfly, 7-9 over the 8 warp partials). red: >= 8.
__device__ __forceinline__ double tile_tree_sum4(const double v[EPT], double* red) {
    double t = (v[0] + v[1]) + (v[2] + v[3]);
    t = warp_tree_sum(t);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = t;
    __syncthreads();
    double r = red[threadIdx.x & 7];
    r += __shfl_xor_sync(0xffffffffu, r, 1);
    r += __shfl_xor_sync(0xffffffffu, r, 2);
    r += __shfl_xor_sync(0xffffffffu, r, 4);
    return r;
}

// Pairwise tree over n per-tile values in global memory (written by other CTAs: read through L2), continuing the per-tile
// trees; all threads of ONE CTA of SCAN_THREADS threads.  The index-aligned tree over 1024 consecutive leaves is what
// tile_tree_sum4 computes from registers (a missing right operand passes the left one through == adding +0.0, exact), so a
// rank's 2^k <= 1024 tiles take one call; more tiles go 1024 at a time and their roots through the same tree once more.
// Same tree, same bits as block_tree_reduce.  red: >= 8 doubles of shared memory, scratch: >= ceil(n/1024) doubles.
__device__ __forceinline__ double rank_tree_sum(const double* vals, unsigned stride_doubles, unsigned n, double* red, double* scratch) {
    auto chunk = [&](const double* base, unsigned stride, unsigned count) -> double {
        double v[EPT];
#pragma unroll
        for (int j = 0; j < EPT; ++j) { const unsigned i = threadIdx.x * EPT + j; v[j] = i < count ? __ldcg(base + (size_t)i * stride) : 0.0; }
        return tile_tree_sum4(v, red);
    };
    const double* src = vals; unsigned stride = stride_doubles, count = n;
    while (count > TILE) {                                 // 1024 values at a time; the roots go through the same tree again
        const unsigned chunks = (count + TILE - 1) / TILE;
        for (unsigned c = 0; c < chunks; ++c) {
            const double r = chunk(src + (size_t)c * TILE * stride, stride, min((unsigned)TILE, count - c * TILE));
            __syncthreads();                               // everyone has read chunk c (scratch[c] may be one of its inputs)
            if (threadIdx.x == 0) scratch[c] = r;
        }
        __threadfence_block();
        __syncthreads();
        src = scratch; stride = 1; count = chunks;
    }
    return chunk(src, stride, count);
}

// the same for ONE CTA of 1024 threads (one value per thread and chunk); red: >= 32 doubles
__device__ __forceinline__ double rank_tree_sum_1024(const double* vals, unsigned n, double* red, double* scratch) {
    const double* src = vals; unsigned count = n;
    while (count > TILE) {
        const unsigned chunks = (count + TILE - 1) / TILE;
        for (unsigned c = 0; c < chunks; ++c) {
            const unsigned i = c * TILE + threadIdx.x;
            const double r = tile_tree_sum_1024(i < count ? __ldcg(src + i) : 0.0, red);
            __syncthreads();
            if (threadIdx.x == 0) scratch[c] = r;
        }
        __threadfence_block();
        __syncthreads();
        src = scratch; count = chunks;
    }
    return tile_tree_sum_1024(threadIdx.x < count ? __ldcg(src + threadIdx.x) : 0.0, red);
}

// ------------------------------------------------------------------------------------------------ map lookup
struct MapTile { const float2* s; int x0, y0, w, h; };

#define IDX_FRAC_BITS 18         // fractional bits of the fixed-point cell quotient (see contact_eval)

// ------------------------------------------------------------------------------------------------ contact model
struct PoseEval { int n_found; double d1, d2, A, poseVar; };

// One map cell from the shared-memory tile or from global memory (L2), predicated: no branch, no generic pointer.
__device__ __forceinline__ float2 load_cell(bool in_tile, uint32_t tile_saddr, const float2* gptr) {
    float2 v;
    asm volatile(
        "{\n"
        " .reg .pred p;\n"
        " setp.ne.s32 p, %4, 0;\n"
        " @p ld.shared.v2.f32 {%0, %1}, [%2];\n"
        " @!p ld.global.nc.v2.f32 {%0, %1}, [%3];\n"
        "}\n"
        : "=f"(v.x), "=f"(v.y)
        : "r"(tile_saddr), "l"(gptr), "r"((int)in_tile));
    return v;
}

// ContactModel::evaluatePose (src/ContactModel.cpp:117-224) for one pose; SHARED arithmetic policy.
// Written straight-line: every lane computes every evaluated contact; lanes whose group is already invalid, or whose
// lookup misses, contribute a zero ratio.  The rare cases (quotient within 2^-18 of a cell border, variance or t
// outside the table) take one slow branch each.
// PATCH: the particle owns a P x P overlay of the shared grid (north_star's per-particle map patches); its cells are
// looked up first (MLSMap::getPatch walks the newest grid first) and the fused contacts are recorded for the write.
struct PatchRef { const float2* cells; int ox, oy, P; };                 // ox == INT_MIN: not anchored yet
struct CpRec { int ix[ESLAM_MAX_CP], iy[ESLAM_MAX_CP]; double h[ESLAM_MAX_CP], zv[ESLAM_MAX_CP]; };

template <bool TAP, bool PATCH, class CS>
__device__ __forceinline__ void contact_eval(const CS& cs, const MapParams& M, const MapTile& mt,
                                             const em::dbl2* htab, double inv_corr,
                                             double x, double y, double z, double s, double c, double mv,
                                             PoseEval& out, int32_t* tap_cells, double* tap_cp,
                                             const PatchRef& pr, CpRec* rec) {
    int valid = 0, gvalid = 1;                                           // :134-135
    double acc_d = 0, acc_v = 0, acc_r = 0;
    double pt_x = 0, pt_y = 0, pt_z = 0;
    int pt_ix = 0, pt_iy = 0;
    out.n_found = 0; out.d1 = 0; out.d2 = 0; out.A = 0; out.poseVar = 0;
    const int n = cs.n;
    const int FRAC_MASK = (1 << IDX_FRAC_BITS) - 1;
    const uint32_t tile_sa = (uint32_t)__cvta_generic_to_shared(mt.s);
#pragma unroll 1
    for (int i = 0; i < n; ++i) {
        const unsigned f = cs.flags[i];
        if (TAP) tap_cells[i] = -1;
        if (f & 1u) {                                                    // :154 !(contact < 0.2): warp-uniform
            const double px = cs.px[i], py = cs.py[i];
            const double wx = em::fma_(c, px, em::fma_(-s, py, x));      // :147
            const double wy = em::fma_(s, px, em::fma_(c, py, y));
            const double wz = z + cs.pzr[i];
            double gx = wx, gy = wy;
            if (!M.identity) {                                           // C_global2local, src/PoseEstimator.hpp:101
                gx = em::fma_(M.g2l[0], wx, em::fma_(M.g2l[1], wy, em::fma_(M.g2l[2], wz, M.g2l[3])));
                gy = em::fma_(M.g2l[4], wx, em::fma_(M.g2l[5], wy, em::fma_(M.g2l[6], wz, M.g2l[7])));
            }
            // cell index floor((g - offset)/scale): g*inv + (1.5*2^34 - offset*inv) puts the quotient with 18 fractional
            // bits into the low mantissa word (error < 2^-18); when the fraction is within one step of an integer
            // (or the quotient is out of the fixed-point range) the exact IEEE division of eslam_math.h decides.
            const uint64_t ux = em::d2u(em::fma_(gx, M.inv_sx, M.c0x)), uy = em::d2u(em::fma_(gy, M.inv_sy, M.c0y));
            const int kx = (int)(uint32_t)ux, ky = (int)(uint32_t)uy;
            int ix = kx >> IDX_FRAC_BITS, iy = ky >> IDX_FRAC_BITS;
            const bool fast = (((uint32_t)(ux >> 32) - 0x4217FFFFu) <= 1u) & (((kx + 1) & FRAC_MASK) > 2) &
                              (((uint32_t)(uy >> 32) - 0x4217FFFFu) <= 1u) & (((ky + 1) & FRAC_MASK) > 2);
            bool inb = ((unsigned)ix < (unsigned)M.W) & ((unsigned)iy < (unsigned)M.H);
            if (!fast) {                                                 // rare
                inb = em::cell_index_exact(gx, M.ox, M.sx, M.W, ix);
                if (inb) inb = em::cell_index_exact(gy, M.oy, M.sy, M.H, iy);
            }
            const int flat = inb ? iy * M.W + ix : 0;
            const unsigned lx = (unsigned)(ix - mt.x0), ly = (unsigned)(iy - mt.y0);
            const bool in_tile = inb & (lx < (unsigned)mt.w) & (ly < (unsigned)mt.h);
            float2 cell = load_cell(in_tile, tile_sa + (ly * mt.w + lx) * 8u, M.cells + flat);
            if (TAP) tap_cells[i] = (inb && gvalid) ? flat : -1;
            if (PATCH) {
                const unsigned plx = (unsigned)(ix - pr.ox), ply = (unsigned)(iy - pr.oy);
                if (inb && pr.ox != INT_MIN && plx < (unsigned)pr.P && ply < (unsigned)pr.P) {
                    const float2 pc = pr.cells[ply * pr.P + plx];
                    if (pc.x == pc.x && pc.y == pc.y) {
                        const double dp = wz - (double)pc.x;
                        if (dp * dp < 9.0 * ((double)pc.y * (double)pc.y + mv)) cell = pc;
                    }
                }
            }
            // candidate lanes: group still valid, inside the grid, cell not empty (NaN mean; a NaN stdev misses too).
            // Non-candidate lanes compute on zeros so that everything below stays finite for them.
            const bool cand = (gvalid != 0) & inb & (cell.x == cell.x) & (cell.y == cell.y);
            const float cm = cand ? cell.x : 0.0f, csd = cand ? cell.y : 0.0f;
            const double d = wz - (double)cm;                            // :160
            const double zvar = em::fma_((double)csd, (double)csd, mv);  // :166-167 (the product of two floats is exact)
            const bool use = cand & (d * d < 9.0 * zvar);                // 3-sigma patch gate of MLSMap::getPatch
            bool rare;
            double ratio = em::likelihood_ratio_fast(d, zvar, inv_corr, htab, rare);           // :170
            if (rare & use) ratio = em::likelihood_ratio(d, zvar, inv_corr, htab);
            const double r_eff = use ? ratio : 0.0;
            if ((TAP || PATCH) && use && !valid) { pt_x = wx; pt_y = wy; pt_z = (double)cell.x; pt_ix = ix; pt_iy = iy; }
            acc_d = em::fma_(d, r_eff, acc_d); acc_v = em::fma_(zvar, r_eff, acc_v);           // :172-184
            acc_r += r_eff;
            valid |= (int)use;                                           // :185
            gvalid &= (int)use;                                          // :189-190
        }
        if (f & 2u) {                                                    // :194-197 end of group: warp-uniform
            if (valid) {
                if (gvalid && acc_r > 1e-9) {                            // :199
                    const double ir = 1.0 / acc_r;
                    const double zd = acc_d * ir, zv = acc_v * ir;       // :201-202
                    out.poseVar += zv - mv;                              // :203 sum(stdev^2 r)/sum(r) = sum(zvar r)/sum(r) - measVar
                    const double iv = 1.0 / zv;
                    const double e = zd * iv;
                    out.d1 += e; out.d2 += iv; out.A = em::fma_(zd, e, out.A);
                    if (TAP && out.n_found < ESLAM_MAX_CP) {
                        double* o = tap_cp + out.n_found * 5;
                        o[0] = pt_x; o[1] = pt_y; o[2] = pt_z; o[3] = zd; o[4] = zv;
                    }
                    if (PATCH && out.n_found < ESLAM_MAX_CP) {
                        rec->ix[out.n_found] = pt_ix; rec->iy[out.n_found] = pt_iy;
                        rec->h[out.n_found] = pt_z + zd; rec->zv[out.n_found] = zv;
                    }
                    out.n_found++;
                }
                gvalid = 1; valid = 0;                                   // :210-213 (only when valid: Q4)
                acc_d = acc_v = acc_r = 0;
            }
        }
    }
}

// ContactModel::evaluateWeight (src/ContactModel.cpp:262-317), three-sum form
__device__ __forceinline__ void evaluate_weight(const PoseEval& e, int use_shape, double& weight, double& z_delta,
                                                double& z_var) {
    const double delta = e.d1 / e.d2;
    double pz = 1.0;
    if (use_shape) {
        double q = em::fma_(-e.d1, delta, e.A);
        if (!(q > 0.0)) q = 0.0;
        pz = em::exp_(-0.5 * q);
    }
    weight = pz; z_delta = -delta; z_var = 1.0 / e.d2;
}

// ContactModel::updateZPositionEstimate (src/ContactModel.cpp:319-340)
__device__ __forceinline__ bool update_z(const PoseEval& e, double m_zDelta, double m_zVar, double& z_pos, double& z_var) {
    const double pose_var = e.poseVar / (double)e.n_found;
    double delta_var = z_var - pose_var;
    if (delta_var < 1e-9) delta_var = 1e-9;                              // std::max(z_var - pose_var, 1e-9)
    if (fabs(m_zDelta / em::sqrt_(delta_var)) > 1.0) return false;      // 1 sigma (Q6)
    const double gain = z_var / (z_var + m_zVar);
    z_pos += gain * m_zDelta;
    const double var_gain = delta_var / (delta_var + m_zVar);
    delta_var = (1.0 - var_gain) * delta_var;
    z_var = pose_var + delta_var;
    return true;
}

// Per-particle map write (north_star's contact-height variant of EmbodiedSlamFilter::processMap,
// src/EmbodiedSlamFilter.cpp:209-227): each fused contact writes {height, sqrt(zvar)} into the particle's own patch,
// Kalman-merged in float with what the cell holds.  The patch is anchored on the particle's cell at its first write.
__device__ __forceinline__ void patch_write(const MapParams& M, float2* cells, int* origin, int P, double x, double y,
                                            int n_found, const CpRec& rec) {
    int ox = origin[0], oy = origin[1];
    if (ox == INT_MIN) {
        double gx = x, gy = y;
        if (!M.identity) {
            gx = em::fma_(M.g2l[0], x, em::fma_(M.g2l[1], y, M.g2l[3]));
            gy = em::fma_(M.g2l[4], x, em::fma_(M.g2l[5], y, M.g2l[7]));
        }
        int ix, iy;
        if (!em::cell_index_exact(gx, M.ox, M.sx, M.W, ix) || !em::cell_index_exact(gy, M.oy, M.sy, M.H, iy)) return;
        ox = ix - P / 2; oy = iy - P / 2;
        origin[0] = ox; origin[1] = oy;
    }
    const int ncp = n_found < ESLAM_MAX_CP ? n_found : ESLAM_MAX_CP;
    for (int k = 0; k < ncp; ++k) {
        const int lx = rec.ix[k] - ox, ly = rec.iy[k] - oy;
        if (lx < 0 || lx >= P || ly < 0 || ly >= P) continue;
        float2* cell = cells + ly * P + lx;
        const float zm = (float)rec.h[k];
        const float sd = (float)em::sqrt_(rec.zv[k]);
        float2 cv = *cell;
        if (!(cv.x == cv.x) || !(cv.y == cv.y)) { cv.x = zm; cv.y = sd; }
        else {
            const float v0 = cv.y * cv.y, v1 = sd * sd;
            const float kg = __fdiv_rn(v0, v0 + v1);
            cv.x = cv.x + kg * (zm - cv.x);
            cv.y = __fsqrt_rn((1.0f - kg) * v0);
        }
        *cell = cv;
    }
}

__device__ __forceinline__ double weighting_function(double x, double alpha, double beta, double gamma) {
    if (x < alpha) return 1.0;                                           // src/PoseEstimator.cpp:114-128
    if (x < beta) {
        const double a = (1.0 - gamma) / (alpha - beta);
        const double b = 1.0 - alpha * a;
        return a * x + b;
    }
    if (x >= beta) return gamma;
    return 0.0;
}

__device__ __forceinline__ void normal_pair(const em::Philox4& p, double& n0, double& n1) {
    em::box_muller(em::u01_open(p.v[0], p.v[1]), em::u01_open(p.v[2], p.v[3]), n0, n1);
}

// ------------------------------------------------------------------------------------------------ init
__global__ void k_init_gaussian(StateBufs st, DevScalars* sc, unsigned long long n, unsigned long long base,
                                uint32_t k0, uint32_t k1, double mx, double my, double mt, double sgx, double sgy,
                                double sgt, double zpos, double zsigma) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const int cur = sc->cur;
    const unsigned long long gi = base + i;
    const uint32_t sw = stream_word(em::STREAM_INIT, gi);
    em::Philox4 a = em::philox4x32_10((uint32_t)gi, 0, 0, sw, k0, k1);
    em::Philox4 b = em::philox4x32_10((uint32_t)gi, 0, 1, sw, k0, k1);
    double nx, ny, nt, unused;
    normal_pair(a, nx, ny);
    normal_pair(b, nt, unused);
    st.f[cur][F_X][i] = nx * sgx + mx;                                   // src/PoseEstimator.cpp:69-72
    st.f[cur][F_Y][i] = ny * sgy + my;
    st.f[cur][F_TH][i] = nt * sgt + mt;
    st.f[cur][F_Z][i] = zpos; st.f[cur][F_ZS][i] = zsigma;
    st.f[cur][F_W][i] = 0.0; st.f[cur][F_MP][i] = 0.0;                   // weight(0): src/PoseParticle.hpp:56 (Q9)
    st.b[cur][B_FLOAT][i] = 1; st.b[cur][B_NF][i] = 0;
}

// ------------------------------------------------------------------------------------------------ project + weight
// Shared-memory layout of the weight kernels: [h table | lean contact list | scratch | map tile (128-byte aligned)]
#define HTAB_BYTES ((ESLAM_H_TABLE_ENTRIES * 16 + 127) / 128 * 128)
#define WT_CONTACT_BYTES (ESLAM_MAX_CONTACTS * 32)         // {px, py, pzr, -} per evaluated contact
#define WT_SCRATCH_BYTES 1024                              // hint reduction (2 x 32 doubles), TMA mbarrier
#define WT_HEAD_BYTES ((HTAB_BYTES + WT_CONTACT_BYTES + WT_SCRATCH_BYTES + 127) / 128 * 128)

// TMA (cp.async.bulk.tensor) + mbarrier helpers: one thread arms the barrier with the byte count and issues the copy,
// everyone waits on the barrier's phase before reading the tile.
__device__ __forceinline__ void mbar_init(uint32_t mbar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(mbar), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t mbar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const void* tmap, uint32_t mbar, int c0, int c1) {
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
                 ::"r"(dst), "l"(tmap), "r"(mbar), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t mbar, uint32_t phase) {
    uint32_t ok;
    do {
        asm volatile("{\n .reg .pred p;\n mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n selp.u32 %0, 1, 0, p;\n}"
                     : "=r"(ok) : "r"(mbar), "r"(phase) : "memory");
    } while (!ok);
}

// Stage the map tile (one TMA bulk tensor copy: box tile_w x tile_h cells at (tile_x0, tile_y0) of the 2-D grid, or
// ordinary loads when the grid cannot be described to TMA) and the pdf/cdf table into shared memory.  All threads call.
__device__ __forceinline__ MapTile stage_tile_and_table(const MapParams& M, int use_tma, const CUtensorMap* tmap,
                                                        const em::dbl2* htab_g, unsigned char* smem_raw) {
    em::dbl2* htab = reinterpret_cast<em::dbl2*>(smem_raw);
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(smem_raw + HTAB_BYTES + WT_CONTACT_BYTES + 512);
    float2* tile = reinterpret_cast<float2*>(smem_raw + WT_HEAD_BYTES);
    const int tid = threadIdx.x;
    MapTile mt; mt.s = tile; mt.w = 0; mt.h = 0; mt.x0 = 0; mt.y0 = 0;
    const bool tma = use_tma && M.tile_w > 0;
    if (M.tile_w > 0) {
        mt.w = M.tile_w; mt.h = M.tile_h; mt.x0 = M.tile_x0; mt.y0 = M.tile_y0;
        if (tma) {
            if (tid == 0) mbar_init(mbar, 1);
            __syncthreads();
            if (tid == 0) {
                mbar_expect_tx(mbar, (uint32_t)(mt.w * mt.h) * (uint32_t)sizeof(float2));
                tma_load_2d((uint32_t)__cvta_generic_to_shared(tile), tmap, mbar, mt.x0, mt.y0);
            }
        } else {
            const int cells = mt.w * mt.h;
            for (int i = tid; i < cells; i += blockDim.x) {
                const int r = i / mt.w, cc = i - r * mt.w;
                tile[i] = __ldg(&M.cells[(size_t)(mt.y0 + r) * M.W + mt.x0 + cc]);
            }
        }
    }
    for (int i = tid; i < ESLAM_H_TABLE_ENTRIES; i += blockDim.x) htab[i] = htab_g[i];   // overlaps the TMA copy
    __syncthreads();
    if (tma) mbar_wait(mbar, 0);
    return mt;
}

// PoseEstimator::project for one particle (src/PoseEstimator.cpp:196-237)
template <class PP>
__device__ __forceinline__ void project_particle(const PP& P, uint32_t pstep, unsigned long long gi, double spread, double& x, double& y,
                                                 double& th, double& z, double& zs, double& w) {
    const uint32_t sw = stream_word(em::STREAM_PROJECT, gi);
    double n0, n1, n2, unused;
    normal_pair(em::philox4x32_10((uint32_t)gi, pstep, 0, sw, P.mp.k0, P.mp.k1), n0, n1);
    normal_pair(em::philox4x32_10((uint32_t)gi, pstep, 1, sw, P.mp.k0, P.mp.k1), n2, unused);
    const double dx = P.dmean[0] + P.dsigma[0] * n0;                     // :198
    double dy = P.dmean[1] + P.dsigma[1] * n1;
    const double dth = P.dmean[2] + P.dsigma[2] * n2;
    const em::Philox4 u = em::philox4x32_10((uint32_t)gi, pstep, 2, sw, P.mp.k0, P.mp.k1);
    if (em::u01(u.v[0], u.v[1]) < P.mp.slip_factor) dy *= em::u01(u.v[2], u.v[3]);   // :199-202
    double s, c; em::sincos_(th, s, c);
    x += c * dx - s * dy;                                                // :205
    y += s * dx + c * dy;
    th += dth;                                                           // :206
    if (P.mp.max_yaw_dev > 0.0) {                                        // :208-216
        if (fabs(th - P.yaw) > P.mp.max_yaw_dev) w *= 0.7;
    }
    z += P.z_delta;                                                      // :218
    zs = em::sqrt_(zs * zs + P.z_var);                                   // :219
    if (spread > 0 && !P.mp.use_hash) {                                  // :224-236
        const double trans_fac = P.mp.spread_tf * spread, rot_fac = P.mp.spread_rf * spread;
        double sx, sy, sth;
        normal_pair(em::philox4x32_10((uint32_t)gi, pstep, 3, sw, P.mp.k0, P.mp.k1), sx, sy);
        normal_pair(em::philox4x32_10((uint32_t)gi, pstep, 4, sw, P.mp.k0, P.mp.k1), sth, unused);
        x += sx * trans_fac + 0.0;
        y += sy * trans_fac + 0.0;
        th += sth * rot_fac + 0.0;
    }
}

// The per-particle tail of updateWeights loop A (src/PoseEstimator.cpp:293-322) once the contact model has run.
template <class PP>
__device__ __forceinline__ void finish_particle(const PP& P, int cur, unsigned long long i, const PoseEval& ev, double z,
                                                double zs2, double w) {
    double mprob = 1.0; uint8_t floating = 1;                            // :316-317
    if ((unsigned long long)ev.n_found >= P.mp.min_contacts) {           // ContactModel.cpp:217
        double weight, zdl, zvr;
        evaluate_weight(ev, P.mp.use_shape, weight, zdl, zvr);
        double zVar = zs2;                                               // :294
        update_z(ev, zdl, zvr, z, zVar);                                 // :295
        P.st.f[cur][F_Z][i] = z; P.st.f[cur][F_ZS][i] = em::sqrt_(zVar); // :296
        P.st.f[cur][F_W][i] = w * weight;                                // :300 (Q1: first application)
        mprob = weight; floating = 0;
    }
    P.st.f[cur][F_MP][i] = mprob;
    P.st.b[cur][B_FLOAT][i] = floating; P.st.b[cur][B_NF][i] = (uint8_t)min(ev.n_found, 255);   // :322 cpoints.size()
}

// sums of x, y over the particles this CTA handled -> hint[blockIdx.x] (only steers where the next tile is put)
__device__ __forceinline__ void write_hint(double hx, double hy, double* red, double* hint) {
    hx = warp_tree_sum(hx); hy = warp_tree_sum(hy);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) { red[warp] = hx; red[32 + warp] = hy; }
    __syncthreads();
    if (warp == 0) {
        const int nw = blockDim.x >> 5;
        const double a = warp_tree_sum(lane < nw ? red[lane] : 0.0), b = warp_tree_sum(lane < nw ? red[32 + lane] : 0.0);
        if (lane == 0) { hint[2 * blockIdx.x] = a; hint[2 * blockIdx.x + 1] = b; }
    }
}

// Warp-chunk schedule shared by the weight kernels: chunk c = 32 consecutive work items; warp w of CTA b takes chunks
// it*(G*W) + w*G + b, so a partly filled last round is spread over all SMs instead of filling a few of them.
__device__ __forceinline__ unsigned long long first_chunk() { return (unsigned long long)(threadIdx.x >> 5) * gridDim.x + blockIdx.x; }
__device__ __forceinline__ unsigned long long chunk_stride() { return (unsigned long long)(blockDim.x >> 5) * gridDim.x; }

// ---- the lean path ------------------------------------------------------------------------------------------------
// Particles are binned by the map tile that holds their whole footprint (k_project_bin), listed per bin
// (k_bin_scatter), and evaluated bin by bin from shared memory (k_weight_lean): every CTA stages the tile of the bin
// it is working on with one TMA bulk tensor copy, so the lookups never leave the SM wherever the cloud lies.
//
// ContactModel::evaluatePose for a particle whose footprint lies inside the staged tile uses the same arithmetic as
// contact_eval for everything that reaches the result (zdiff, zvar, 1/sqrt, table, running sums, group fusion); what
// differs is HOW two decisions are taken:
//   * cell index: one fused fixed-point quotient per axis (tile-relative, 24 fractional bits) instead of transform +
//     divide; if the fraction is within 4 steps of a cell border the decision is not trusted;
//   * 3-sigma gate d^2 < 9 zvar: decided on the high word of |t| = |zdiff| / (sigma corr) against 3/corr; within one
//     high-word step (2^-20 relative) of the threshold the decision is not trusted.
// An untrusted decision is taken again for that contact by lean_contact_exact (IEEE division, the gate as written):
// results are bit-identical to the generic path by construction.
#define MAX_BINS 4096
#define LEAN_REDO_CAP (WT_CONTACT_BYTES / 4)   // the lean kernel's list of deferred exact evaluations lives where the generic kernel keeps its contact list

__device__ __forceinline__ void bin_origin(const MapParams& M, const FastContacts& F, int b, int& x0, int& y0) {
    const int by = (int)__umulhi((unsigned)b << 1, F.inv_nbx), bx = b - by * F.nbx;
    x0 = max(0, min(bx * F.bsx - F.margin, M.W - M.tile_w)) & ~1;        // W - tile_w is even when the lean path is on
    y0 = max(0, min(by * F.bsy - F.margin, M.H - M.tile_h));
}

// project (optional) + which bin.  One particle per thread, one 1024-particle index tile per CTA: ranks inside a bin are
// handed out through a shared-memory histogram, one global atomic per (CTA, non-empty bin).
template <bool DO_PROJECT>
__global__ void __launch_bounds__(1024, 1) k_project_bin(const __grid_constant__ FastParams P, int* __restrict__ bin_count,
                                                         uint2* __restrict__ keys) {
    pdl_start();
    __shared__ int hist[MAX_BINS], gbase[MAX_BINS];
    const WeightCommon& C = P.c; const FastContacts& F = P.fc;
    const int tid = threadIdx.x, lane = tid & 31;
    const int cur = C.sc->cur;
    for (int b = tid; b < F.nb; b += 1024) hist[b] = 0;
    __syncthreads();
    const unsigned long long i = (unsigned long long)blockIdx.x * 1024 + tid;
    int bin = -1, rank = 0; bool flagged = false;
    if (i < C.n) {
        double x = C.st.f[cur][F_X][i], y = C.st.f[cur][F_Y][i], zs = C.st.f[cur][F_ZS][i];
        if (DO_PROJECT) {
            double th = C.st.f[cur][F_TH][i], z = C.st.f[cur][F_Z][i], w = C.st.f[cur][F_W][i];
            const double spread = weighting_function(C.sc->max_weight, 0.0, C.mp.spread_thr, 0.0);   // :194
            project_particle(C, C.sc->project_step, C.base + i, spread, x, y, th, z, zs, w);
            C.st.f[cur][F_X][i] = x; C.st.f[cur][F_Y][i] = y; C.st.f[cur][F_TH][i] = th;
            C.st.f[cur][F_Z][i] = z; C.st.f[cur][F_ZS][i] = zs; C.st.f[cur][F_W][i] = w;   // everything project changed
        }
        const double mv = zs * zs + C.mp.merr2;                          // :290 (zero or out of range: generic kernel)
        // the particle's own cell (grid-relative fixed point, 18 fractional bits; may be off by one cell: the margin
        // has room for that), its bin, and whether its footprint is inside that bin's tile
        const uint64_t ux = em::d2u(em::fma_(x, C.map.inv_sx, C.map.c0x)), uy = em::d2u(em::fma_(y, C.map.inv_sy, C.map.c0y));
        const int cx = (int)(uint32_t)ux >> IDX_FRAC_BITS, cy = (int)(uint32_t)uy >> IDX_FRAC_BITS;
        // in the grid and inside the fixed-point range: the particle has a bin (its tile, clamped into the grid, holds every
        // in-grid cell of its footprint).  Inside the core and measVar inside rsqrt_core's range: the lean evaluation applies;
        // otherwise the particle is flagged and evaluated by contact_eval from the same tile (map border) or through L2.
        const bool placed = (((uint32_t)(ux >> 32) - 0x4217FFFFu) <= 1u) & (((uint32_t)(uy >> 32) - 0x4217FFFFu) <= 1u) &
                            ((unsigned)cx < (unsigned)C.map.W) & ((unsigned)cy < (unsigned)C.map.H);
        bool ok = placed & (((uint32_t)(em::d2u(mv) >> 52) - 923u) < 199u);   // measVar inside rsqrt_core's range
        int b = 0;
        if (placed) {
            b = (int)__umulhi((unsigned)cy, F.inv_bsy) * F.nbx + (int)__umulhi((unsigned)cx, F.inv_bsx);
            int x0, y0; bin_origin(C.map, F, b, x0, y0);
            ok = ok & ((unsigned)(cx - x0 - F.margin) < (unsigned)(C.map.tile_w - 2 * F.margin)) &
                 ((unsigned)(cy - y0 - F.margin) < (unsigned)(C.map.tile_h - 2 * F.margin));
        }
        bin = b;
        flagged = !ok;
    }
    {   // rank inside the bin: one shared-memory atomic per (warp, bin) instead of one per particle (a warp's particles fall
        // into one to three bins, so per-particle atomics on the same address serialise 32-fold)
        const unsigned peers_m = __match_any_sync(0xffffffffu, bin);
        const int leader = __ffs(peers_m) - 1;
        int base_rank = 0;
        if (lane == leader && bin >= 0) base_rank = atomicAdd(&hist[bin], __popc(peers_m));
        base_rank = __shfl_sync(0xffffffffu, base_rank, leader);
        rank = base_rank + __popc(peers_m & ((1u << lane) - 1u));
    }
    const unsigned m = __ballot_sync(0xffffffffu, flagged);
    if (m && lane == (__ffs(m) - 1)) atomicAdd(&C.sc->defer_count, (unsigned)__popc(m));   // diagnostics only
    __syncthreads();
    for (int b = tid; b < F.nb; b += 1024) { const int c = hist[b]; if (c > 0) gbase[b] = atomicAdd(&bin_count[b], c); }
    __syncthreads();
    if (i < C.n) keys[i] = make_uint2((unsigned)bin | (flagged ? 0x80000000u : 0u), (unsigned)(gbase[bin] + rank));
}

// bin-sorted particle list, the offset of every bin in it, and the bin each CTA of the lean kernel starts on
// (one CTA per non-empty bin, the others in proportion to the bins' particles)
__global__ void __launch_bounds__(1024, 1) k_bin_scatter(const int* __restrict__ bin_count, int nb, const uint2* __restrict__ keys,
                                                         unsigned long long n, uint32_t* __restrict__ list, unsigned* __restrict__ bin_offs,
                                                         int* __restrict__ cta_bin, int n_cta, DevScalars* sc, int did_project) {
    pdl_start();
    if (did_project && blockIdx.x == 0 && threadIdx.x == 0) sc->project_step += 1;   // k_project_bin has finished: its project is counted
    __shared__ unsigned offs[MAX_BINS + 1];
    __shared__ unsigned wsum[32];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    // exclusive scan over the bins (4 consecutive bins per thread)
    unsigned c[4], lo = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) { const int b = tid * 4 + j; c[j] = b < nb ? (unsigned)bin_count[b] : 0u; lo += c[j]; }
    unsigned so = lo;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned a = __shfl_up_sync(0xffffffffu, so, o); if (lane >= o) so += a; }
    if (lane == 31) wsum[warp] = so;
    __syncthreads();
    if (warp == 0) {
        unsigned a = wsum[lane];
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned x = __shfl_up_sync(0xffffffffu, a, o); if (lane >= o) a += x; }
        wsum[lane] = a;
    }
    __syncthreads();
    unsigned eo = so - lo + (warp ? wsum[warp - 1] : 0u);
#pragma unroll
    for (int j = 0; j < 4; ++j) { offs[tid * 4 + j] = eo; eo += c[j]; }
    if (tid == 1023) offs[MAX_BINS] = eo;
    __syncthreads();
    const unsigned long long i = (unsigned long long)blockIdx.x * 1024 + tid;
    if (i < n) { const uint2 k = keys[i]; list[offs[k.x & 0x7FFFFFFFu] + k.y] = (uint32_t)i | (k.x & 0x80000000u); }   // bit 31: flagged
    if (blockIdx.x == 0) {
        for (int b = tid; b < nb; b += 1024) bin_offs[b] = offs[b];
        const unsigned M = offs[MAX_BINS];
        // Where the CTAs of the lean kernel start.  Every non-empty bin gets one CTA first (up to n_cta bins): a wide cloud has a
        // fringe of small bins, and a bin nobody starts on is only reached by stealing, i.e. AFTER the big bins are exhausted --
        // one more tile load and chunk round at the very end of the kernel.  Started at once, the small bins are done long before
        // the big ones, whose shared queues then absorb every CTA and end the kernel at chunk granularity.  The remaining CTAs
        // start where sorted position (i + 1/2) M / R lies: in proportion to the bins' particles.
        unsigned nz = 0;
#pragma unroll
        for (int j = 0; j < 4; ++j) nz += c[j] > 0;
        unsigned sz = nz;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned a = __shfl_up_sync(0xffffffffu, sz, o); if (lane >= o) sz += a; }
        __syncthreads();                                  // wsum is free again
        if (lane == 31) wsum[warp] = sz;
        __syncthreads();
        if (warp == 0) {
            unsigned a = wsum[lane];
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const unsigned x = __shfl_up_sync(0xffffffffu, a, o); if (lane >= o) a += x; }
            wsum[lane] = a;
        }
        __syncthreads();
        unsigned r = sz - nz + (warp ? wsum[warp - 1] : 0u);                   // non-empty bins in front of this thread's four
        const unsigned B = wsum[31];
#pragma unroll
        for (int j = 0; j < 4; ++j) if (c[j] > 0) { if (r < (unsigned)n_cta) cta_bin[r] = tid * 4 + j; ++r; }
        const unsigned first = B < (unsigned)n_cta ? B : (unsigned)n_cta, R = (unsigned)n_cta - first;
        if ((unsigned)tid < R) {
            const unsigned p = (unsigned)(((unsigned long long)(2 * tid + 1) * M) / (2ull * R));
            int lo_b = 0, hi_b = nb;                      // last bin whose offset is <= p: it holds p when it is not empty
            while (hi_b - lo_b > 1) { const int mid = (lo_b + hi_b) >> 1; if (offs[mid] <= p) lo_b = mid; else hi_b = mid; }
            cta_bin[first + tid] = lo_b;
        }
    }
}

// h(t) outside the table (|t| >= 10): kept out of line, it is never on a hot path
__device__ __noinline__ double h_slow_call(double t) { return em::h_slow(t); }

// shared-memory reads by 32-bit shared address (pure: the staged data do not change while they are read, so the
// compiler may schedule these freely)
__device__ __forceinline__ void lds_f64x2(uint32_t sa, double& a, double& b) { asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(a), "=d"(b) : "r"(sa)); }
__device__ __forceinline__ void lds_f32x4(uint32_t sa, float& a, float& b, float& c, float& d) {
    asm("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(a), "=f"(b), "=f"(c), "=f"(d) : "r"(sa));
}
__device__ __forceinline__ double lds_f64(uint32_t sa) { double a; asm("ld.shared.f64 %0, [%1];" : "=d"(a) : "r"(sa)); return a; }
__device__ __forceinline__ float2 lds_f32x2(uint32_t sa) { float2 v; asm("ld.shared.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(sa)); return v; }

struct LeanTile { uint32_t sa; int w, x0, y0; double c0x, c0y; };

// EXACT == false: the lean decisions; returns false when one of them was too close to its threshold to be trusted.
// EXACT == true: every decision as contact_eval takes it (cell index through the IEEE division, the gate as
// d^2 < 9 zvar); always returns true.  Everything that reaches the result is the same arithmetic in both.
// (top byte of ky) << 8 | top byte of kx, for ky < 2^31: prmt's sign-replicating selectors (PTX ISA, prmt default mode, selector
// bit 3) give the two zero bytes -- __byte_perm documents only three selector bits, hence the PTX
__device__ __forceinline__ uint32_t row_col_bytes(uint32_t kx, uint32_t ky) {
    uint32_t r;
    asm("prmt.b32 %0, %1, %2, 0xFF73;" : "=r"(r) : "r"(kx), "r"(ky));
    return r;
}

template <bool EXACT, int UNROLL, bool P256>
__device__ __forceinline__ bool lean_eval(const FastContacts& F, const MapParams& M, const LeanTile& T,
                                          uint32_t htab_sa, double inv_corr, double x, double y, double s, double c,
                                          double z, double mv, PoseEval& out) {
    const double axx = c * M.inv_sx, axy = -s * M.inv_sx, ayx = s * M.inv_sy, ayy = c * M.inv_sy;
    const double ucx = em::fma_(x, M.inv_sx, T.c0x), ucy = em::fma_(y, M.inv_sy, T.c0y);
    const uint32_t gate_lo = F.gate_lo;
    // Flags as integers (no loop-carried predicates): pass = -1 while every evaluated contact since the last reset was
    // inside the gate (gvalid of :135,:189-190: once a contact misses, nothing is used until the next reset); n_used
    // counts the contacts used since the last reset (valid of :134,:185 == n_used > 0).  border_min / gate_min keep the
    // minimum distance of any lean decision from its threshold; both are examined once, after the loop.
    int pass = -1, n_used = 0;
    uint32_t border_min = 0xFFFFFFFFu, gate_min = 0xFFFFFFFFu;
    double acc_d = 0, acc_v = 0, acc_r = 0;
    out.n_found = 0; out.d1 = 0; out.d2 = 0; out.A = 0; out.poseVar = 0;
    const int ng = F.n_groups;
    // Software pipeline of the lean loop: the cell of contact i+1 is looked up (constant-bank read of the contact, fused
    // index, shared-memory load) while contact i goes through its dependent chain (1/sqrt, table, sums).  The look-up of
    // contact n_eval (zeros: the particle's own cell) is issued and dropped.
    float2 cell_n = make_float2(0.f, 0.f); double wz_n = 0.0;
    auto look_up = [&](int j) {
        const double px = F.px[j], py = F.py[j];                         // warp-uniform: constant bank, not the shared-memory pipe
        wz_n = z + F.pzr[j];                                             // :147
        const uint64_t ux = em::d2u(em::fma_(axx, px, em::fma_(axy, py, ucx)));
        const uint64_t uy = em::d2u(em::fma_(ayx, px, em::fma_(ayy, py, ucy)));
        const uint32_t kx = (uint32_t)ux, ky = (uint32_t)uy;
        // (k + 4) << 8 drops the cell bits: below 8 << 8 when the fraction is within 4/2^24 of a cell border
        border_min = min(border_min, min(kx * 256u + 1024u, ky * 256u + 1024u));
        // tile rows are 256 cells apart when the grid is at least that wide: the cell's offset is then (row byte, column byte), one
        // byte permute (the tile has fewer than 128 rows, so the row byte's sign bit supplies the two zero bytes)
        if (P256) cell_n = lds_f32x2(T.sa + (row_col_bytes(kx, ky) << 3));
        else cell_n = lds_f32x2(T.sa + ((ky >> 24) * (uint32_t)T.w + (kx >> 24)) * 8u);
    };
    int i = 0;
    if (!EXACT) look_up(0);
#pragma unroll 1
    for (int g = 0; g < ng; ++g) {
        const int e = F.gend[g];
#pragma unroll UNROLL
        for (; i < e; ++i) {
            double d, zvar, ratio; int inside;
            if (!EXACT) {
                const float2 cell = cell_n; const double wz = wz_n;
                look_up(i + 1);
                d = wz - (double)cell.x;                                 // :160
                zvar = em::fma_((double)cell.y, (double)cell.y, mv);     // :166-167
                const double qc = em::rsqrt_core(zvar) * inv_corr;       // 1 / (sigma corr)  :109,:170
                const double t = d * qc;
                const uint32_t dlt = ((uint32_t)(em::d2u(t) >> 32) & 0x7FFFFFFFu) - gate_lo;   // < 0: inside the gate
                gate_min = min(gate_min, dlt);                           // NaN cells (empty): |t| high word >= 0x7FF00000, a miss
                inside = (int)dlt >> 31;
                uint32_t j; double zc; em::h_index(t, j, zc);
                double c0, c1; float c2, c3, c4, c5;
                lds_f64x2(htab_sa + 16u * j, c0, c1); lds_f32x4(htab_sa + 16u * (ESLAM_H_NINT + j), c2, c3, c4, c5);
                ratio = em::h_poly(zc, c0, c1, c2, c3, c4, c5) * qc;     // :111-112
            } else {
                const double px = F.px[i], py = F.py[i];
                const double wz = z + F.pzr[i];                          // :147
                const double wx = em::fma_(c, px, em::fma_(-s, py, x));
                const double wy = em::fma_(s, px, em::fma_(c, py, y));
                int ix = 0, iy = 0;
                bool inb = em::cell_index_exact(wx, M.ox, M.sx, M.W, ix);
                if (inb) inb = em::cell_index_exact(wy, M.oy, M.sy, M.H, iy);
                const unsigned lx = (unsigned)(ix - T.x0), ly = (unsigned)(iy - T.y0);
                inb = inb & (lx < (unsigned)M.tile_w) & (ly < (unsigned)M.tile_h);   // always true for a binned particle
                float2 cl = make_float2(0.f, 0.f);
                if (inb) cl = lds_f32x2(T.sa + (ly * (unsigned)M.tile_w + lx) * 8u);
                d = wz - (double)cl.x;
                zvar = em::fma_((double)cl.y, (double)cl.y, mv);
                const bool hit = inb & (cl.x == cl.x) & (cl.y == cl.y) & (d * d < 9.0 * zvar);
                ratio = 0.0;
                if (hit) {                                               // em::likelihood_ratio
                    const double q2 = (em::rsqrt_in_range(zvar) ? em::rsqrt_core(zvar) : 1.0 / em::sqrt_(zvar)) * inv_corr;
                    const double t2 = d * q2;
                    double hv;
                    if (em::h_in_range(t2)) {
                        uint32_t j; double zc; em::h_index(t2, j, zc);
                        double c0, c1; float c2, c3, c4, c5;
                        lds_f64x2(htab_sa + 16u * j, c0, c1); lds_f32x4(htab_sa + 16u * (ESLAM_H_NINT + j), c2, c3, c4, c5);
                        hv = em::h_poly(zc, c0, c1, c2, c3, c4, c5);
                    } else hv = h_slow_call(t2);
                    ratio = hv * q2;
                }
                inside = hit ? -1 : 0;
            }
            pass &= inside;                                              // use = gvalid && hit; gvalid = use  :154,:189-190
            n_used -= pass;                                              // :185
            // :172-184, unconditionally: once a contact of the group has missed (pass == 0) the group can no longer be
            // pushed (:199 needs gvalid) and the sums are cleared or never read again, so what they hold does not matter
            acc_d = em::fma_(d, ratio, acc_d); acc_v = em::fma_(zvar, ratio, acc_v); acc_r += ratio;
        }
        if (n_used > 0) {                                                // :194-197 (the group's last evaluated point)
            if (pass && acc_r > 1e-9) {                                  // :199
                const double ir = 1.0 / acc_r;
                const double zd = acc_d * ir, zv = acc_v * ir;           // :201-202
                out.poseVar += zv - mv;                                  // :203
                const double iv = 1.0 / zv;
                const double q = zd * iv;
                out.d1 += q; out.d2 += iv; out.A = em::fma_(zd, q, out.A);
                out.n_found++;
            }
            pass = -1; n_used = 0;                                       // :210-213 (only when valid: Q4)
            acc_d = acc_v = acc_r = 0;
        }
    }
    // a decision taken while the group was already invalid cannot matter but is counted as well: it is rare
    return EXACT || (border_min >= 2048u && gate_min > F.gate_span);
}

// a flagged particle (footprint over the map border, out-of-range values): ContactModel::evaluatePose as the generic kernel
// does it -- tile first, L2 otherwise, exact cell index, the gate as written -- on the compact contact list.  Out of line.
__device__ __noinline__ void lean_generic_eval(const FastParams* P, const LeanTile* T, const float2* tile, const em::dbl2* htab, int cur,
                                               unsigned long long i, PoseEval* out) {
    const WeightCommon& C = P->c;
    const double x = C.st.f[cur][F_X][i], y = C.st.f[cur][F_Y][i], th = C.st.f[cur][F_TH][i];
    const double z = C.st.f[cur][F_Z][i], zs = C.st.f[cur][F_ZS][i];
    double s, c; em::sincos_(th, s, c);
    const double mv = zs * zs + C.mp.merr2;
    if (mv == 0) atomicOr(&C.sc->error_flags, ERR_ZERO_VARIANCE_BIT);   // src/ContactModel.cpp:122-123
    MapTile mt; mt.s = tile; mt.x0 = T->x0; mt.y0 = T->y0; mt.w = C.map.tile_w; mt.h = C.map.tile_h;
    PatchRef pr; pr.cells = nullptr; pr.ox = INT_MIN; pr.oy = 0; pr.P = 0;
    contact_eval<false, false>(P->fc, C.map, mt, htab, C.mp.inv_corr, x, y, z, s, c, mv, *out, nullptr, nullptr, pr, nullptr);
}

// the whole particle again with exact decisions (~2e-4 of the particles); out of line so that it costs the lean loop no registers
__device__ __noinline__ void lean_redo_exact(const FastParams* P, const LeanTile* T, uint32_t htab_sa, int cur,
                                             unsigned long long i, PoseEval* out) {
    const WeightCommon& C = P->c;
    const double x = C.st.f[cur][F_X][i], y = C.st.f[cur][F_Y][i], th = C.st.f[cur][F_TH][i];
    const double z = C.st.f[cur][F_Z][i], zs = C.st.f[cur][F_ZS][i];
    double s, c; em::sincos_(th, s, c);
    lean_eval<true, 1, false>(P->fc, C.map, *T, htab_sa, C.mp.inv_corr, x, y, s, c, z, zs * zs + C.mp.merr2, *out);
}

// The same, by all 32 lanes of the warp for ONE particle (i is warp-uniform): one lane redoing 64 contacts alone keeps its warp
// for two more chunk times, and when that happens in the last round of a bin the whole CTA (and the kernel) waits for it.
// Lane l takes the decisions and the likelihood of contacts l, l + 32, ... exactly as lean_eval<EXACT> does per contact; the
// running sums and the group fusion (:172-213) then run over the contacts in order, every lane the same, the per-contact
// values fetched by shuffles.  Every lane returns the particle's PoseEval.
__device__ __noinline__ void lean_redo_exact_warp(const FastParams* P, const LeanTile* Tp, uint32_t htab_sa, int cur,
                                                  unsigned long long i, PoseEval* out) {
    const WeightCommon& C = P->c; const FastContacts& F = P->fc; const MapParams& M = C.map; const LeanTile& T = *Tp;
    const int lane = threadIdx.x & 31;
    const double x = C.st.f[cur][F_X][i], y = C.st.f[cur][F_Y][i], th = C.st.f[cur][F_TH][i];
    const double z = C.st.f[cur][F_Z][i], zs = C.st.f[cur][F_ZS][i];
    double s, c; em::sincos_(th, s, c);
    const double mv = zs * zs + C.mp.merr2, inv_corr = C.mp.inv_corr;
    int pass = -1, n_used = 0, g = 0;
    double acc_d = 0, acc_v = 0, acc_r = 0;
    PoseEval ev; ev.n_found = 0; ev.d1 = 0; ev.d2 = 0; ev.A = 0; ev.poseVar = 0;
    const int n = F.n_eval;
    for (int j0 = 0; j0 < n; j0 += 32) {
        const int j = j0 + lane;
        double d = 0, zvar = 0, ratio = 0; int inside = 0;
        if (j < n) {
            const double px = F.px[j], py = F.py[j];
            const double wz = z + F.pzr[j];                              // :147
            const double wx = em::fma_(c, px, em::fma_(-s, py, x));
            const double wy = em::fma_(s, px, em::fma_(c, py, y));
            int ix = 0, iy = 0;
            bool inb = em::cell_index_exact(wx, M.ox, M.sx, M.W, ix);
            if (inb) inb = em::cell_index_exact(wy, M.oy, M.sy, M.H, iy);
            const unsigned lx = (unsigned)(ix - T.x0), ly = (unsigned)(iy - T.y0);
            inb = inb & (lx < (unsigned)M.tile_w) & (ly < (unsigned)M.tile_h);   // always true for a binned particle
            float2 cl = make_float2(0.f, 0.f);
            if (inb) cl = lds_f32x2(T.sa + (ly * (unsigned)M.tile_w + lx) * 8u);
            d = wz - (double)cl.x;
            zvar = em::fma_((double)cl.y, (double)cl.y, mv);
            const bool hit = inb & (cl.x == cl.x) & (cl.y == cl.y) & (d * d < 9.0 * zvar);
            if (hit) {                                                   // em::likelihood_ratio
                const double q2 = (em::rsqrt_in_range(zvar) ? em::rsqrt_core(zvar) : 1.0 / em::sqrt_(zvar)) * inv_corr;
                const double t2 = d * q2;
                double hv;
                if (em::h_in_range(t2)) {
                    uint32_t hj; double zc; em::h_index(t2, hj, zc);
                    double c0, c1; float c2, c3, c4, c5;
                    lds_f64x2(htab_sa + 16u * hj, c0, c1); lds_f32x4(htab_sa + 16u * (ESLAM_H_NINT + hj), c2, c3, c4, c5);
                    hv = em::h_poly(zc, c0, c1, c2, c3, c4, c5);
                } else hv = h_slow_call(t2);
                ratio = hv * q2;
            }
            inside = hit ? -1 : 0;
        }
        const int m = min(32, n - j0);
        for (int l = 0; l < m; ++l) {
            const double dl = __shfl_sync(0xffffffffu, d, l), vl = __shfl_sync(0xffffffffu, zvar, l), rl = __shfl_sync(0xffffffffu, ratio, l);
            const int il = __shfl_sync(0xffffffffu, inside, l);
            pass &= il; n_used -= pass;                                  // :154,:185,:189-190
            acc_d = em::fma_(dl, rl, acc_d); acc_v = em::fma_(vl, rl, acc_v); acc_r += rl;
            if (j0 + l + 1 == (int)F.gend[g]) {                          // the group's last evaluated point
                if (n_used > 0) {                                        // :194-197
                    if (pass && acc_r > 1e-9) {                          // :199
                        const double ir = 1.0 / acc_r;
                        const double zd = acc_d * ir, zv = acc_v * ir;   // :201-202
                        ev.poseVar += zv - mv;                           // :203
                        const double iv = 1.0 / zv;
                        const double q = zd * iv;
                        ev.d1 += q; ev.d2 += iv; ev.A = em::fma_(zd, q, ev.A);
                        ev.n_found++;
                    }
                    pass = -1; n_used = 0;                               // :210-213 (only when valid: Q4)
                    acc_d = acc_v = acc_r = 0;
                }
                ++g;
            }
        }
    }
    *out = ev;
}

// Work distribution: a chunk is 32 consecutive entries of a bin's list.  Every CTA works on one bin at a time (its tile
// is what the CTA has in shared memory); its warps pull chunks from the bin's global counter, so all CTAs on a bin finish
// it together.  A CTA whose bin is exhausted joins the bin with the most chunks left (restaging the tile) or exits.
template <int THREADS, bool P256>
__global__ void __launch_bounds__(THREADS, 1) k_weight_lean(const __grid_constant__ FastParams P, const uint32_t* __restrict__ list,
                                                            const int* __restrict__ bin_count, const unsigned* __restrict__ bin_offs,
                                                            unsigned* __restrict__ bin_next, const int* __restrict__ cta_bin) {
    pdl_start();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const WeightCommon& C = P.c;
    em::dbl2* htab = reinterpret_cast<em::dbl2*>(smem_raw);
    const uint32_t htab_sa = (uint32_t)__cvta_generic_to_shared(smem_raw);
    const uint32_t mbar = htab_sa + HTAB_BYTES + WT_CONTACT_BYTES + 512;
    int* s_pick = reinterpret_cast<int*>(smem_raw + HTAB_BYTES + WT_CONTACT_BYTES);          // [0] best remaining, [1] its bin, [2] the chunk claimed for the CTA (-1: none), [3] entries of s_redo
    uint32_t* s_redo = reinterpret_cast<uint32_t*>(smem_raw + HTAB_BYTES);                  // particles of this bin visit waiting for their exact evaluation
    float2* tile = reinterpret_cast<float2*>(smem_raw + WT_HEAD_BYTES);
    const int tid = threadIdx.x, lane = tid & 31;
    const int cur = C.sc->cur;
    const int nb = P.fc.nb;
    if (tid == 0) mbar_init(mbar, 1);
    __syncthreads();
    for (int i = tid; i < ESLAM_H_TABLE_ENTRIES; i += THREADS) htab[i] = C.htab[i];   // visible after the barrier below
    LeanTile T; T.sa = (uint32_t)__cvta_generic_to_shared(tile); T.w = C.map.tile_w; T.x0 = 0; T.y0 = 0; T.c0x = 0; T.c0y = 0;
    int bin = cta_bin[blockIdx.x];
    uint32_t phase = 0;
    while (true) {
        const unsigned cnt = (unsigned)bin_count[bin], nch = (cnt + 31u) >> 5, base = bin_offs[bin];
        // one thread claims a first chunk for the CTA: the tile is only staged when there really is work left
        if (tid == 0) { const unsigned c0 = atomicAdd(&bin_next[bin], 1u); s_pick[2] = c0 < nch ? (int)c0 : -1; s_pick[3] = 0; }
        __syncthreads();
        const int first = s_pick[2];
        if (first >= 0) {
            bin_origin(C.map, P.fc, bin, T.x0, T.y0);
            T.c0x = P.fc.c0x - (double)T.x0; T.c0y = P.fc.c0y - (double)T.y0;   // exact
            if (C.use_tma) {
                if (tid == 0) {
                    mbar_expect_tx(mbar, (uint32_t)(C.map.tile_w * C.map.tile_h) * (uint32_t)sizeof(float2));
                    tma_load_2d(T.sa, C.tmap, mbar, T.x0, T.y0);
                }
                mbar_wait(mbar, phase); phase ^= 1u;
            } else {
                const int cells = C.map.tile_w * C.map.tile_h;
                for (int i = tid; i < cells; i += THREADS) {
                    const int rr = i / C.map.tile_w, cc = i - rr * C.map.tile_w;
                    tile[i] = __ldg(&C.map.cells[(size_t)(T.y0 + rr) * C.map.W + T.x0 + cc]);
                }
                __syncthreads();
            }
            bool have_first = tid < 32;                                  // warp 0 starts with the claimed chunk
            while (true) {
                unsigned ch = (unsigned)first;
                if (!have_first) {
                    if (lane == 0) ch = atomicAdd(&bin_next[bin], 1u);
                    ch = __shfl_sync(0xffffffffu, ch, 0);
                }
                have_first = false;
                if (ch >= nch) break;
                const unsigned k = (ch << 5) + lane;
                const bool have = k < cnt;
                const uint32_t entry = have ? list[base + k] : 0u;
                const unsigned long long i = entry & 0x7FFFFFFFu;
                double z = 0, zs2 = 0;
                PoseEval ev;
                bool redo = false;
                if (have) {
                    z = C.st.f[cur][F_Z][i]; const double zs = C.st.f[cur][F_ZS][i];
                    zs2 = zs * zs;
                    if (entry >> 31) lean_generic_eval(&P, &T, tile, htab, cur, i, &ev);
                    else {
                        const double x = C.st.f[cur][F_X][i], y = C.st.f[cur][F_Y][i], th = C.st.f[cur][F_TH][i];
                        double s, c; em::sincos_(th, s, c);              // :279-282
                        redo = !lean_eval<false, 2, P256>(P.fc, C.map, T, htab_sa, C.mp.inv_corr, x, y, s, c, z, zs2 + C.mp.merr2 /* :290 */, ev);
                    }
                }
                // A decision too close to its threshold (~1e-4 of the particles; copies of one ancestor that fall into the same cell
                // share it, so they come in runs): the particle is evaluated again with exact decisions, by a whole warp.  Not here --
                // a chunk with eight of them would keep this warp for five chunk times, and the CTA with it -- but from the CTA's
                // list once its warps have run out of chunks: one particle per warp, all warps at once.
                bool redo_here = false;
                if (redo) {
                    const int slot = atomicAdd(&s_pick[3], 1);
                    if (slot < LEAN_REDO_CAP) s_redo[slot] = (uint32_t)i; else redo_here = true;
                }
                for (unsigned m = __ballot_sync(0xffffffffu, redo_here); m; m &= m - 1) {   // the list is full: as before
                    const int src = __ffs(m) - 1;
                    PoseEval e2;
                    lean_redo_exact_warp(&P, &T, htab_sa, cur, __shfl_sync(0xffffffffu, i, src), &e2);
                    if (lane == src) ev = e2;
                }
                if (have && (!redo || redo_here)) finish_particle(C, cur, i, ev, z, zs2, C.st.f[cur][F_W][i]);
            }
            __syncthreads();                                             // the CTA's list of deferred particles is complete
            const int n_redo = min(s_pick[3], LEAN_REDO_CAP);
            for (int r = tid >> 5; r < n_redo; r += THREADS / 32) {
                const unsigned long long i = s_redo[r];
                PoseEval ev;
                lean_redo_exact_warp(&P, &T, htab_sa, cur, i, &ev);
                if (lane == 0) {
                    const double z = C.st.f[cur][F_Z][i], zs = C.st.f[cur][F_ZS][i];
                    finish_particle(C, cur, i, ev, z, zs * zs, C.st.f[cur][F_W][i]);
                }
            }
        }
        // everyone is done with this tile; pick the bin with the most chunks left
        if (tid == 0) { s_pick[0] = 0; s_pick[1] = -1; }
        __syncthreads();
        int best = 0, best_bin = -1;
        for (int b = tid; b < nb; b += THREADS) {
            const int left = (int)(((unsigned)bin_count[b] + 31u) >> 5) - (int)min(bin_next[b], 0x7FFFFFFFu);
            if (left > best) { best = left; best_bin = b; }
        }
        if (best > 0) atomicMax(&s_pick[0], best);
        __syncthreads();
        if (best > 0 && best == s_pick[0]) s_pick[1] = best_bin;         // any of the ties
        __syncthreads();
        bin = s_pick[1];
        if (bin < 0) break;
        __syncthreads();
    }
}

// ---- the generic path ---------------------------------------------------------------------------------------------
// Any particle, any map (tile + L2 fallback, exact cell index, per-particle patches, C_global2local).
template <bool DO_PROJECT, bool DO_WEIGHT, bool PATCH>
__global__ void __launch_bounds__(WT_THREADS, 1) k_project_weight(const __grid_constant__ WeightParams P) {
    pdl_start();
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const em::dbl2* htab = reinterpret_cast<const em::dbl2*>(smem_raw);
    double* red = reinterpret_cast<double*>(smem_raw + HTAB_BYTES + WT_CONTACT_BYTES);
    const int tid = threadIdx.x, lane = tid & 31;
    const DevScalars* sc = P.sc;
    const int cur = sc->cur;
    const unsigned long long count = P.n;
    const unsigned long long n_chunks = (count + 31) >> 5;
    MapTile mt; mt.s = nullptr; mt.w = 0; mt.h = 0; mt.x0 = 0; mt.y0 = 0;
    if (DO_WEIGHT) mt = stage_tile_and_table(P.map, P.use_tma, P.tmap, P.htab, smem_raw);
    double spread = 0.0;
    if (DO_PROJECT) spread = weighting_function(sc->max_weight, 0.0, P.mp.spread_thr, 0.0);   // :194
    const uint32_t pstep = sc->project_step;

    double* X = P.st.f[cur][F_X]; double* Y = P.st.f[cur][F_Y]; double* TH = P.st.f[cur][F_TH];
    double* Z = P.st.f[cur][F_Z]; double* ZS = P.st.f[cur][F_ZS]; double* WG = P.st.f[cur][F_W];
    double hx = 0, hy = 0;
    for (unsigned long long ch = first_chunk(); ch < n_chunks; ch += chunk_stride()) {
        const unsigned long long k = (ch << 5) + lane;
        if (k >= count) continue;
        const unsigned long long i = k;
        double x = X[i], y = Y[i], th = TH[i], z = Z[i], zs = ZS[i], w = WG[i];
        if (DO_PROJECT) {                                                        // :196-237
            project_particle(P, pstep, P.base + i, spread, x, y, th, z, zs, w);
            X[i] = x; Y[i] = y; TH[i] = th; Z[i] = z; ZS[i] = zs; WG[i] = w;     // everything project changed
        }
        hx += x; hy += y;
        if (DO_WEIGHT) {                                                         // :276-327
            double s, c; em::sincos_(th, s, c);
            const double zs2 = zs * zs;
            const double mv = zs2 + P.mp.merr2;                                  // :290
            if (mv == 0) atomicOr(&P.sc->error_flags, ERR_ZERO_VARIANCE_BIT);   // src/ContactModel.cpp:122-123
            PoseEval ev;
            PatchRef pr; pr.cells = nullptr; pr.ox = INT_MIN; pr.oy = 0; pr.P = 0;
            CpRec rec;
            if (PATCH) {
                const size_t pp = (size_t)P.patch_size * P.patch_size;
                pr.cells = P.patch[cur] + i * pp; pr.P = P.patch_size;
                pr.ox = P.porigin[cur][2 * i]; pr.oy = P.porigin[cur][2 * i + 1];
            }
            contact_eval<false, PATCH>(P.cs, P.map, mt, htab, P.mp.inv_corr, x, y, z, s, c, mv, ev, nullptr, nullptr,
                                       pr, &rec);
            if (PATCH)
                patch_write(P.map, P.patch[cur] + i * (size_t)P.patch_size * P.patch_size, P.porigin[cur] + 2 * i,
                            P.patch_size, x, y, ev.n_found, rec);
            finish_particle(P, cur, i, ev, z, zs2, w);
        }
    }
    if (DO_WEIGHT) write_hint(hx, hy, red, P.hint);
}

// ------------------------------------------------------------------------------------------------ loop A partials
// data_particles, total_points, max_weight and the terms pow(w, 1/found) of sum_data_weights (src/PoseEstimator.cpp:
// 305-310) per 1024-particle tile, from what the weight kernels left per particle (mprob, floating, cpoints.size()).
// The last CTA to finish reduces the rank's tiles (pairwise tree, continuing the per-tile trees) and publishes the rank's
// partial to every rank (exchange X_A); it also clears the lean path's bin counters for the next update and refreshes the
// generic path's tile placement hint.
struct TilePartArgs {
    TilePartA* part; unsigned n_tiles;
    double *scratch_a, *scratch_b;
    const double* hint; unsigned n_hint; unsigned long long n_local;
    int* bin_count; unsigned* bin_next;
    int bump_project;              // this update projected on the generic kernel: count the project (device-side Philox step)
};
__global__ void __launch_bounds__(SCAN_THREADS) k_tile_part(StateBufs st, DevScalars* sc, unsigned long long n, const TilePartArgs A,
                                                            const __grid_constant__ Xchg X) {
    pdl_start();
    __shared__ double red[8];
    __shared__ double redm[8];
    __shared__ unsigned redu[16];
    __shared__ unsigned long long redull[16];
    __shared__ RankPartA s_out;
    __shared__ unsigned s_epoch;
    const int cur = sc->cur;
    if (A.bump_project && blockIdx.x == 0 && threadIdx.x == 0) sc->project_step += 1;
    const unsigned long long i0 = (unsigned long long)blockIdx.x * TILE + threadIdx.x * EPT;
    double mp[EPT]; unsigned nf[EPT], fl[EPT];
    if (i0 + EPT <= n) {
        const double2 a = *reinterpret_cast<const double2*>(st.f[cur][F_MP] + i0), b = *reinterpret_cast<const double2*>(st.f[cur][F_MP] + i0 + 2);
        const uchar4 f = *reinterpret_cast<const uchar4*>(st.b[cur][B_FLOAT] + i0), c = *reinterpret_cast<const uchar4*>(st.b[cur][B_NF] + i0);
        mp[0] = a.x; mp[1] = a.y; mp[2] = b.x; mp[3] = b.y;
        fl[0] = f.x; fl[1] = f.y; fl[2] = f.z; fl[3] = f.w; nf[0] = c.x; nf[1] = c.y; nf[2] = c.z; nf[3] = c.w;
    } else {
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            const bool in = i0 + j < n;
            mp[j] = in ? st.f[cur][F_MP][i0 + j] : 0.0; fl[j] = in ? st.b[cur][B_FLOAT][i0 + j] : 1u; nf[j] = in ? st.b[cur][B_NF][i0 + j] : 0u;
        }
    }
    double pw[EPT], mx = 0; unsigned cnt = 0, pts = 0;
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
        pw[j] = 0.0;
        if (!fl[j]) {
            pw[j] = em::root_n(mp[j], (unsigned long long)nf[j]);        // :309
            mx = fmax(mx, mp[j]); cnt += 1; pts += nf[j];                // :305-310
        }
    }
    const double s = tile_tree_sum4(pw, red);
    mx = warp_max(mx); cnt = warp_sum_u(cnt); pts = warp_sum_u(pts);
    if ((threadIdx.x & 31) == 0) { redm[threadIdx.x >> 5] = mx; redu[threadIdx.x >> 5] = cnt; redu[8 + (threadIdx.x >> 5)] = pts; }
    __syncthreads();
    if (threadIdx.x == 0) {
        TilePartA o; o.sum_pw = s; o.max_w = 0; o.cnt = 0; o.pts = 0;
        for (int k = 0; k < SCAN_THREADS / 32; ++k) { o.max_w = fmax(o.max_w, redm[k]); o.cnt += redu[k]; o.pts += redu[8 + k]; }
        A.part[blockIdx.x] = o;
    }
    if (!cta_is_last(&sc->ticket[0])) return;
    // ---- last CTA: the rank's partial (floating_weight :329 and max_weight :305,:347-348 need all ranks': k_loop_b combines them)
    for (int b = threadIdx.x; b < MAX_BINS; b += blockDim.x) { A.bin_count[b] = 0; A.bin_next[b] = 0; }   // for the next update's lean path
    double rmx = 0; unsigned long long rcnt = 0, rpts = 0;
    for (unsigned i = threadIdx.x; i < A.n_tiles; i += blockDim.x) {
        const double* pp = reinterpret_cast<const double*>(A.part + i);   // 24 bytes, written by another CTA: read through L2
        const uint2 cp = __ldcg(reinterpret_cast<const uint2*>(pp + 2));
        rmx = fmax(rmx, __ldcg(pp + 1)); rcnt += cp.x; rpts += cp.y;
    }
    const double sum_pw = rank_tree_sum(reinterpret_cast<const double*>(A.part), sizeof(TilePartA) / 8, A.n_tiles, red, A.scratch_a);
    rmx = warp_max(rmx);
    for (int o = 1; o < 32; o <<= 1) { rcnt += __shfl_xor_sync(0xffffffffu, rcnt, o); rpts += __shfl_xor_sync(0xffffffffu, rpts, o); }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { redm[threadIdx.x >> 5] = rmx; redull[threadIdx.x >> 5] = rcnt; redull[8 + (threadIdx.x >> 5)] = rpts; }
    __syncthreads();
    if (threadIdx.x == 0) {
        RankPartA o; o.sum_pw = sum_pw; o.max_w = 0; o.cnt = 0; o.pts = 0;
        for (int k = 0; k < SCAN_THREADS / 32; ++k) { o.max_w = fmax(o.max_w, redm[k]); o.cnt += redull[k]; o.pts += redull[8 + k]; }
        s_out = o;
        s_epoch = sc->x_epoch[X_A] + 1;
        if (A.n_local > 0 && A.n_hint > 0) {
            double tsx = 0, tsy = 0;
            for (unsigned i = 0; i < A.n_hint; ++i) { tsx += __ldcg(A.hint + 2 * i); tsy += __ldcg(A.hint + 2 * i + 1); }
            sc->tile_cx = tsx / (double)A.n_local; sc->tile_cy = tsy / (double)A.n_local;
        }
        sc->defer_last = sc->defer_count; sc->defer_count = 0;
    }
    __syncthreads();
    ll_publish(X, XOFF_A + (size_t)X.me * XA_WORDS * 8, &s_out, XA_WORDS, s_epoch);
    if (threadIdx.x == 0) sc->x_epoch[X_A] = s_epoch;
}

// LOCAL_MIGRATE sharding (include/eslam_gpu.h): the logical shards are runs of tiles_per_shard consecutive tiles of the
// global tile index space; this context holds shards [first, first + count).  mode == 0: everything below is inert.
struct ShardPlan {
    int mode;                        // ESLAM_SHARD_LOCAL_MIGRATE when there is more than one logical shard, else 0
    unsigned n_shards, tiles_per_shard, first, count;
    unsigned long long n_shard;      // particles per shard
    double min_share, min_ess;
    double* shard_tot;               // [n_shards] exact left-to-right total of every local shard (written by k_chain)
};

// ------------------------------------------------------------------------------------------------ loop B + raw sums
// weight *= mprob * pow(discount*floating_weight, 4 - cpoints.size())  (:332-345); per-tile sums of the result.
// Prologue (APPLY): every CTA combines the ranks' loop-A partials (exchange X_A) in tree order into floating_weight (:329)
// and the power table; CTA 0 also records them, and max_weight (:305, :347-348), in DevScalars.
// Last CTA: the rank's weight sum (pairwise tree over its tiles), the raw exclusive prefix of its tile sums (only classifies
// elements later: approximate is enough) and, for LOCAL_MIGRATE, the per-shard sums; publishes the rank's sums (X_B).
struct LoopBArgs {
    double2* tile_sum; unsigned n_tiles;
    double discount;
    double *scratch_a, *scratch_b;
    double* prefix_raw;            // [n_tiles] raw (un-normalised) sum of the tiles before each local tile
    double2* shard_sums;           // [sp.count] {sum w, sum w^2} of every logical shard held here (virtual shards)
};
template <bool APPLY>
__global__ void __launch_bounds__(SCAN_THREADS) k_loop_b(StateBufs st, DevScalars* sc, unsigned long long n, const LoopBArgs A,
                                                         const ShardPlan sp, const __grid_constant__ Xchg X) {
    pdl_start();
    __shared__ double red[8];
    __shared__ double s_pow[6];
    __shared__ __align__(8) unsigned s_in[ESLAM_MAX_RANKS * XA_WORDS];
    __shared__ double sh[64];
    __shared__ RankPartB s_out;
    __shared__ unsigned s_epoch;
    const int cur = sc->cur;
    // this tile's weights (and mprob, cpoints.size()) are on their way while the CTA waits for the ranks' loop-A partials
    const unsigned long long i0 = (unsigned long long)blockIdx.x * TILE + threadIdx.x * EPT;
    double w[EPT], mp[EPT]; unsigned nfv[EPT];
    double* W = st.f[cur][F_W];
    const bool full = i0 + EPT <= n;
    if (full) {
        const double2 a = *reinterpret_cast<const double2*>(W + i0), b = *reinterpret_cast<const double2*>(W + i0 + 2);
        w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y;
        if (APPLY) {
            const double2 ma = *reinterpret_cast<const double2*>(st.f[cur][F_MP] + i0), mb = *reinterpret_cast<const double2*>(st.f[cur][F_MP] + i0 + 2);
            const uchar4 nf = *reinterpret_cast<const uchar4*>(st.b[cur][B_NF] + i0);
            mp[0] = ma.x; mp[1] = ma.y; mp[2] = mb.x; mp[3] = mb.y;
            nfv[0] = nf.x; nfv[1] = nf.y; nfv[2] = nf.z; nfv[3] = nf.w;
        }
    } else {
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            const bool in = i0 + j < n;
            w[j] = in ? W[i0 + j] : 0.0;
            if (APPLY) { mp[j] = in ? st.f[cur][F_MP][i0 + j] : 0.0; nfv[j] = in ? st.b[cur][B_NF][i0 + j] : 0u; }
        }
    }
    if (APPLY) {
        ll_gather(X, XOFF_A, XA_WORDS, sc->x_epoch[X_A], s_in, &sc->error_flags);
        if (threadIdx.x == 0) {
            double v[ESLAM_MAX_RANKS]; double smx = 0; unsigned long long scnt = 0, spts = 0;
            for (int q = 0; q < X.world; ++q) {
                const RankPartA r = *reinterpret_cast<const RankPartA*>(s_in + q * XA_WORDS);
                v[q] = r.sum_pw; smx = fmax(smx, r.max_w); scnt += r.cnt; spts += r.pts;
            }
            const double sum_pw = tree_small(v, X.world);
            const double fw = scnt > 0 ? sum_pw / (double)scnt : 1.0;                     // :329
            const double xb = A.discount * fw;
            for (int k = 0; k <= 4; ++k) s_pow[k] = em::pow_u64(xb, (unsigned long long)(4 - k));   // :336
            s_pow[5] = em::pow_u64(xb, ~0ull);                                            // wrapped size_t exponent (Q2)
            if (blockIdx.x == 0) {
                const double last_max = sc->max_weight;                                   // :268
                sc->floating_weight = fw;
                sc->data_particles = scnt; sc->total_points = spts;
                sc->max_weight = (spts == 0) ? last_max * A.discount : smx;               // :347-348
                for (int k = 0; k < 6; ++k) sc->pow_table[k] = s_pow[k];
            }
        }
        __syncthreads();
#pragma unroll
        for (int j = 0; j < EPT; ++j) { const double factor = mp[j] * s_pow[nfv[j] > 4 ? 5 : nfv[j]]; w[j] *= factor; }   // Q1: second application
        if (full) {
            *reinterpret_cast<double2*>(W + i0) = make_double2(w[0], w[1]);
            *reinterpret_cast<double2*>(W + i0 + 2) = make_double2(w[2], w[3]);
        } else {
#pragma unroll
            for (int j = 0; j < EPT; ++j) if (i0 + j < n) W[i0 + j] = w[j];
        }
    }
    const double s = tile_tree_sum4(w, red);
    double q[EPT];
#pragma unroll
    for (int j = 0; j < EPT; ++j) q[j] = w[j] * w[j];
    const double sq = tile_tree_sum4(q, red);                        // per-shard ESS of the LOCAL_MIGRATE decision
    if (threadIdx.x == 0) A.tile_sum[blockIdx.x] = make_double2(s, sq);
    if (!cta_is_last(&sc->ticket[1])) return;
    // ---- last CTA: getWeightsSum (src/ParticleFilter.hpp:34-39) of this rank's particles
    const unsigned T = A.n_tiles;
    const double rank_w = rank_tree_sum(reinterpret_cast<const double*>(A.tile_sum), 2, T, red, A.scratch_a);
    double rank_sq = 0.0;
    if (sp.mode) {
        // LOCAL_MIGRATE: per logical shard held here the pairwise trees over its tiles' sums of w and of w^2
        for (unsigned g = 0; g < sp.count; ++g) {
            const double* src = reinterpret_cast<const double*>(A.tile_sum + (size_t)g * sp.tiles_per_shard);
            const double w0 = rank_tree_sum(src, 2, sp.tiles_per_shard, red, A.scratch_a);
            __syncthreads();
            const double w1 = rank_tree_sum(src + 1, 2, sp.tiles_per_shard, red, A.scratch_a);
            __syncthreads();
            if (threadIdx.x == 0) A.shard_sums[g] = make_double2(w0, w1);
            rank_sq = w1;                                             // real shards: the rank is one shard
        }
    }
    // raw exclusive scan of the tile sums: rounds of blockDim tiles, one per thread, a carry from round to round
    {
        const unsigned lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
        double carry = 0.0;
        __syncthreads();
        for (unsigned r0 = 0; r0 < T; r0 += blockDim.x) {
            const unsigned t = r0 + threadIdx.x;
            const double m = t < T ? __ldcg(&A.tile_sum[t]).x : 0.0;
            double inc = m;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const double v = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= (unsigned)o) inc += v; }
            if (lane == 31) sh[warp] = inc;
            __syncthreads();
            if (warp == 0) {
                double ws = lane < nw ? sh[lane] : 0.0;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const double v = __shfl_up_sync(0xffffffffu, ws, o); if (lane >= (unsigned)o) ws += v; }
                sh[32 + lane] = ws;
            }
            __syncthreads();
            if (t < T) A.prefix_raw[t] = carry + ((warp ? sh[32 + warp - 1] : 0.0) + (inc - m));
            carry += sh[32 + nw - 1];                                            // every thread keeps the same carry
            __syncthreads();
        }
    }
    if (threadIdx.x == 0) {
        s_out.sum_w = rank_w; s_out.sum_sq = rank_sq;
        s_epoch = sc->x_epoch[X_B] + 1;
        sc->pool_count = 0;                                                      // boundary-element pool of k_norm_scan
    }
    __syncthreads();
    ll_publish(X, XOFF_B + (size_t)X.me * XB_WORDS * 8, &s_out, XB_WORDS, s_epoch);
    if (threadIdx.x == 0) sc->x_epoch[X_B] = s_epoch;
}

// What every CTA that consumes exchange X_B derives from the ranks' sums: getWeightsSum() over all ranks (tree order), the Q8
// flag, the LOCAL_MIGRATE decision, and the raw weight in front of this rank.  One thread computes, the CTA shares it.
struct SumInfo { double total, rank_prefix; int uniform, migrate; };
__device__ __noinline__ SumInfo combine_sums(const Xchg& X, const unsigned* gathered, const ShardPlan& sp, const double2* shard_sums, int normalise) {
    SumInfo o;
    double v[ESLAM_MAX_RANKS], sqv[ESLAM_MAX_RANKS];
    o.rank_prefix = 0.0;
    for (int q = 0; q < X.world; ++q) {
        const RankPartB r = *reinterpret_cast<const RankPartB*>(gathered + q * XB_WORDS);
        v[q] = r.sum_w; sqv[q] = r.sum_sq;
        if (q < X.me) o.rank_prefix += v[q];
    }
    double wv[ESLAM_MAX_RANKS];
    for (int q = 0; q < X.world; ++q) wv[q] = v[q];
    o.total = tree_small(v, X.world);
    o.uniform = (normalise && o.total <= 0.0) ? 1 : 0;                            // :51 (Q8)
    // LOCAL_MIGRATE: does any shard's share of the weight or its own effective sample size collapse?  Real shards: one per
    // rank, from the exchange; virtual shards: from this context's own per-shard sums.  Every rank decides the same.
    o.migrate = 1;
    if (sp.mode) {
        o.migrate = 0;
        if (o.total > 0.0) {
            for (unsigned g = 0; g < sp.n_shards; ++g) {
                double w0, w1;
                if (X.world > 1) { w0 = wv[g]; w1 = sqv[g]; } else { const double2 t = __ldcg(&shard_sums[g]); w0 = t.x; w1 = t.y; }
                const double share = (w0 / o.total) * (double)sp.n_shards;
                const double essg = (w0 * w0) / w1;
                if (share < sp.min_share || essg < sp.min_ess * (double)sp.n_shard) o.migrate = 1;
            }
        }
    }
    return o;
}

// getWeightsSum() alone (eslam_gpu_weights_sum): combine the ranks' sums, then a rendezvous so that no rank overwrites the
// exchange slots while another still reads them.
__global__ void k_sum_finish(DevScalars* sc, const ShardPlan sp, const double2* shard_sums, const __grid_constant__ Xchg X) {
    __shared__ unsigned s_e;
    __shared__ __align__(8) unsigned s_in[ESLAM_MAX_RANKS * XB_WORDS];
    ll_gather(X, XOFF_B, XB_WORDS, sc->x_epoch[X_B], s_in, &sc->error_flags);
    if (threadIdx.x == 0) {
        const SumInfo si = combine_sums(X, s_in, sp, shard_sums, 0);
        sc->sum_w = si.total;
        s_e = sc->x_epoch[X_BAR] + 1;
    }
    __syncthreads();
    const unsigned e = s_e;
    x_signal(X, X_BAR, e);
    x_wait(X, X_BAR, e, &sc->error_flags);
    if (threadIdx.x == 0) sc->x_epoch[X_BAR] = e;
}

// ------------------------------------------------------------------------------------------------ exact running sum
// Element classes for the exact emulation of  sum_w += xi_k[idx].weight  (src/ParticleFilter.hpp:95,102):
//   neutral  : adding it never changes the running sum (zero, or below a quarter ulp of the sum)
//   interior : the sum stays inside one binade 2^b across this add -> the add is an integer add of
//              rne(w / ulp_b), with ties depending only on the parity of the sum: a 2-state function (d0, d1)
//   boundary : anything else (first element, binade crossings, ambiguous cases): added in real fp64 by the chain
struct ElemClass { int kind; int b; long long d0, d1; };   // kind: 0 neutral, 1 interior, 2 boundary

__device__ __forceinline__ ElemClass classify_elem(double w, double a_prev, double a_cur, double eps) {
    ElemClass c; c.kind = 0; c.b = BIN_NONE; c.d0 = 0; c.d1 = 0;
    if (w == 0.0) return c;
    c.kind = 2;
    if (!(w > 0.0) || !(a_cur < 1.7e308)) return c;                    // negative / NaN / inf: let the chain add it
    const double lo = em::fma_(-a_prev, eps, a_prev), hi = em::fma_(a_cur, eps, a_cur);
    const uint64_t ulo = em::d2u(lo), uhi = em::d2u(hi), uw = em::d2u(w);
    if (!(lo > 0.0) || (ulo >> 52) == 0) return c;
    const int blo = (int)(ulo >> 52) - 1023, bhi = (int)(uhi >> 52) - 1023;
    const int ew = ((uw >> 52) == 0) ? -1023 : (int)(uw >> 52) - 1023;
    if (ew + 1 <= blo - 54) { c.kind = 0; return c; }                   // w < 2^(blo-54): under a quarter ulp
    if (blo != bhi || (uw >> 52) == 0) return c;
    const int k = blo - ew;
    if (k < 0) return c;
    const uint64_t M = (uw & 0x000FFFFFFFFFFFFFull) | 0x0010000000000000ull;
    c.kind = 1; c.b = blo;
    if (k == 0) { c.d0 = c.d1 = (long long)M; return c; }
    const uint64_t fl = M >> k, frac = M & ((1ull << k) - 1), half = 1ull << (k - 1);
    const long long up = frac > half ? 1 : 0;
    const bool tie = frac == half;
    c.d0 = (long long)fl + up + ((tie && (fl & 1)) ? 1 : 0);
    c.d1 = (long long)fl + up + ((tie && !(fl & 1)) ? 1 : 0);
    return c;
}

__device__ __forceinline__ SegF segf_identity() { SegF f; f.d0 = 0; f.d1 = 0; f.b = BIN_NONE; f.pad = 0; return f; }
// a then b.  pad carries a "binade mismatch" flag (must never happen; triggers the sequential fallback)
__device__ __forceinline__ SegF segf_compose(const SegF& a, const SegF& b) {
    SegF c;
    c.b = (b.b != BIN_NONE) ? b.b : a.b;
    c.pad = a.pad | b.pad | ((a.b != BIN_NONE && b.b != BIN_NONE && a.b != b.b) ? 1 : 0);
    c.d0 = a.d0 + ((a.d0 & 1) ? b.d1 : b.d0);
    c.d1 = a.d1 + (((1 + a.d1) & 1) ? b.d1 : b.d0);
    return c;
}
__device__ __forceinline__ double segf_apply(double S, const SegF& f, int& bad) {
    if (f.pad) bad = 1;
    if (f.b == BIN_NONE) return S;
    const uint64_t u = em::d2u(S);
    if ((int)(u >> 52) - 1023 != f.b) { bad = 1; return S; }
    return em::u2d(u + (uint64_t)((u & 1) ? f.d1 : f.d0));
}

struct SegE { SegF f; int head; };   // scan element: function since the last boundary at or before (head = its index, -1 none)
__device__ __forceinline__ SegE sege_combine(const SegE& L, const SegE& R) {
    if (R.head >= 0) return R;
    SegE o; o.head = L.head; o.f = segf_compose(L.f, R.f); return o;
}
__device__ __forceinline__ SegE sege_shfl_up(const SegE& v, int o) {
    SegE r;
    r.f.d0 = __shfl_up_sync(0xffffffffu, v.f.d0, o); r.f.d1 = __shfl_up_sync(0xffffffffu, v.f.d1, o);
    r.f.b = __shfl_up_sync(0xffffffffu, v.f.b, o); r.f.pad = __shfl_up_sync(0xffffffffu, v.f.pad, o);
    r.head = __shfl_up_sync(0xffffffffu, v.head, o);
    return r;
}

struct TileScan {          // per-thread results of the tile-local scans (EPT consecutive elements per thread)
    ElemClass cls[EPT];
    SegE inc[EPT];         // inclusive segmented function
    int bnd_before[EPT];   // boundary elements strictly before this one in the tile
    int bnd_total;
};

// Shared scratch for tile_scan: one entry per warp (8 warps)
struct TileScanSmem { double d[8]; SegE e[8]; int c[8]; };

__device__ __forceinline__ SegE sege_identity() { SegE e; e.f = segf_identity(); e.head = -1; return e; }
__device__ __forceinline__ SegE sege_shfl(const SegE& v, int src) {
    SegE r;
    r.f.d0 = __shfl_sync(0xffffffffu, v.f.d0, src); r.f.d1 = __shfl_sync(0xffffffffu, v.f.d1, src);
    r.f.b = __shfl_sync(0xffffffffu, v.f.b, src); r.f.pad = __shfl_sync(0xffffffffu, v.f.pad, src);
    r.head = __shfl_sync(0xffffffffu, v.head, src);
    return r;
}

// Tile-local approximate inclusive prefix + classification + segmented function scan.
// blockDim.x == SCAN_THREADS; thread t holds elements 4t .. 4t+3 of the tile (inactive ones as w = 0).
// The order of the floating-point adds is fixed, so k_norm_scan and k_cdf see identical bits.
__device__ __forceinline__ void tile_scan(const double w[EPT], double tile_prefix, double eps, TileScanSmem& sm, TileScan& r) {
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    // approximate prefix: local inclusive sums, warp scan of the thread totals, scan of the 8 warp totals
    double l[EPT];
    l[0] = w[0]; l[1] = l[0] + w[1]; l[2] = l[1] + w[2]; l[3] = l[2] + w[3];
    double a = l[3];
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const double v = __shfl_up_sync(0xffffffffu, a, o); if (lane >= o) a += v; }
    __syncthreads();
    if (lane == 31) sm.d[warp] = a;
    __syncthreads();
    double wp = sm.d[lane & 7];
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) { const double v = __shfl_up_sync(0xffffffffu, wp, o); if ((lane & 7) >= o) wp += v; }
    const double warp_excl = __shfl_sync(0xffffffffu, wp, max(warp - 1, 0));
    double lane_excl = __shfl_up_sync(0xffffffffu, a, 1);
    if (lane == 0) lane_excl = 0.0;
    const double base = tile_prefix + ((warp > 0 ? warp_excl : 0.0) + lane_excl);

    SegE loc[EPT]; int cnt = 0;
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
        const double a_prev = (j == 0) ? base : base + l[j - 1], a_cur = base + l[j];
        r.cls[j] = classify_elem(w[j], a_prev, a_cur, eps);
        SegE e = sege_identity();
        if (r.cls[j].kind == 1) { e.f.d0 = r.cls[j].d0; e.f.d1 = r.cls[j].d1; e.f.b = r.cls[j].b; }
        if (r.cls[j].kind == 2) e.head = (int)threadIdx.x * EPT + j;
        loc[j] = (j == 0) ? e : sege_combine(loc[j - 1], e);
        r.bnd_before[j] = cnt;
        cnt += (r.cls[j].kind == 2) ? 1 : 0;
    }
    // scan of the per-thread aggregates
    SegE agg = loc[EPT - 1]; int c = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const SegE v = sege_shfl_up(agg, o); const int cv = __shfl_up_sync(0xffffffffu, c, o);
        if (lane >= o) { agg = sege_combine(v, agg); c += cv; }
    }
    __syncthreads();
    if (lane == 31) { sm.e[warp] = agg; sm.c[warp] = c; }
    __syncthreads();
    SegE we = sm.e[lane & 7]; int wc = sm.c[lane & 7];
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
        const SegE v = sege_shfl_up(we, o); const int cv = __shfl_up_sync(0xffffffffu, wc, o);
        if ((lane & 7) >= o) { we = sege_combine(v, we); wc += cv; }
    }
    r.bnd_total = __shfl_sync(0xffffffffu, wc, 7);
    const SegE wpre = sege_shfl(we, max(warp - 1, 0)); const int wcpre = __shfl_sync(0xffffffffu, wc, max(warp - 1, 0));
    SegE excl = sege_shfl_up(agg, 1); int cexcl = __shfl_up_sync(0xffffffffu, c, 1);
    if (lane == 0) { excl = sege_identity(); cexcl = 0; }
    if (warp > 0) { excl = sege_combine(wpre, excl); cexcl += wcpre; }
#pragma unroll
    for (int j = 0; j < EPT; ++j) { r.inc[j] = sege_combine(excl, loc[j]); r.bnd_before[j] += cexcl; }
}

__device__ __forceinline__ void load4(const double* p, unsigned long long i0, unsigned long long n, double w[EPT]) {
    if (i0 + EPT <= n) {
        const double2 a = *reinterpret_cast<const double2*>(p + i0), b = *reinterpret_cast<const double2*>(p + i0 + 2);
        w[0] = a.x; w[1] = a.y; w[2] = b.x; w[3] = b.y;
    } else {
#pragma unroll
        for (int j = 0; j < EPT; ++j) w[j] = (i0 + j < n) ? p[i0 + j] : 0.0;
    }
}
__device__ __forceinline__ void store4(double* p, unsigned long long i0, unsigned long long n, const double w[EPT]) {
    if (i0 + EPT <= n) {
        *reinterpret_cast<double2*>(p + i0) = make_double2(w[0], w[1]);
        *reinterpret_cast<double2*>(p + i0 + 2) = make_double2(w[2], w[3]);
    } else {
#pragma unroll
        for (int j = 0; j < EPT; ++j) if (i0 + j < n) p[i0 + j] = w[j];
    }
}

// normalise (src/ParticleFilter.hpp:51-67), per-tile sum of w^2, and the tile descriptors of the exact running sum.
// Prologue: every CTA combines the ranks' weight sums (exchange X_B) into getWeightsSum() and derives its tile's approximate
// prefix (rank prefix + raw tile prefix, normalised); CTA 0 records sum / Q8 flag / LOCAL_MIGRATE decision in DevScalars.
struct NormScanArgs {
    unsigned long long n_global;
    const double* prefix_raw; const double2* shard_sums;
    double* tile_prefix;           // [all tiles] approximate prefix this kernel used for its tile (k_cdf reads the same bits)
    double* tile_sq; TileDesc* desc; BndEntry* pool;
    double *scratch_a, *scratch_b;
    unsigned n_tiles, tile_ofs;
    int normalise; double eps;
    long long* el_d0; unsigned short* el_code;   // per element: what k_cdf needs to materialise its running sum (see k_cdf)
};
__global__ void __launch_bounds__(SCAN_THREADS, 4) k_norm_scan(StateBufs st, DevScalars* sc, unsigned long long n, const NormScanArgs A,
                                                            const ShardPlan sp, const __grid_constant__ Xchg X) {
    pdl_start();
    __shared__ TileScanSmem sm;
    __shared__ double red[8];
    __shared__ int slot_of[TILE];
    __shared__ int first_kind[SCAN_THREADS / 32 + 1];
    __shared__ int s_base;
    __shared__ double s_total, s_prefix; __shared__ int s_uniform;
    __shared__ __align__(8) unsigned s_in[ESLAM_MAX_RANKS * XB_WORDS];
    const int cur = sc->cur;
    const int normalise = A.normalise;
    const unsigned long long i0 = (unsigned long long)blockIdx.x * TILE + threadIdx.x * EPT;
    double w[EPT];
    load4(st.f[cur][F_W], i0, n, w);                                 // on its way while the CTA waits for the ranks' sums
    ll_gather(X, XOFF_B, XB_WORDS, sc->x_epoch[X_B], s_in, &sc->error_flags);
    if (threadIdx.x == 0) {
        const SumInfo si = combine_sums(X, s_in, sp, A.shard_sums, normalise);
        const bool local = sp.mode && !si.migrate;
        // approximate running sum in front of this tile (only classifies elements).  Local resampling: it restarts at every shard.
        const unsigned bt = blockIdx.x;
        const unsigned seg0 = local ? (bt / sp.tiles_per_shard) * sp.tiles_per_shard : 0u;   // first local tile of this tile's walk
        double prefix;
        if (si.uniform) {
            const unsigned long long before = (unsigned long long)(bt - seg0) * TILE + (local ? 0ull : (unsigned long long)A.tile_ofs * TILE);
            prefix = (double)before / (double)A.n_global;
        } else {
            double raw = A.prefix_raw[bt] - (seg0 ? A.prefix_raw[seg0] : 0.0);
            if (!local) raw += si.rank_prefix;
            prefix = normalise ? raw / si.total : raw;
        }
        s_total = si.total; s_uniform = si.uniform; s_prefix = prefix;
        A.tile_prefix[A.tile_ofs + bt] = prefix;
        if (blockIdx.x == 0) { sc->sum_w = si.total; sc->uniform = si.uniform; sc->migrate = si.migrate; }
    }
    __syncthreads();
    if (normalise) {
        const int uniform = s_uniform; const double sum_w = s_total;
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            if (i0 + j < n) {
                if (uniform) w[j] = 1.0 / (double)A.n_global;                         // :55
                else w[j] /= sum_w;                                                  // :64
            }
        }
        store4(st.f[cur][F_W], i0, n, w);
    }
    double q[EPT];
#pragma unroll
    for (int j = 0; j < EPT; ++j) q[j] = w[j] * w[j];                                 // effective += w*w
    const double sq = tile_tree_sum4(q, red);
    if (threadIdx.x == 0) A.tile_sq[blockIdx.x] = sq;

    TileScan ts;
    tile_scan(w, s_prefix, A.eps, sm, ts);
    TileDesc* D = &A.desc[blockIdx.x];
    BndEntry* pool = A.pool;
#pragma unroll
    for (int j = 0; j < EPT; ++j) slot_of[threadIdx.x * EPT + j] = ts.bnd_before[j];
    if ((threadIdx.x & 31) == 0) first_kind[threadIdx.x >> 5] = ts.cls[0].kind;
    if (threadIdx.x == 0) {
        int base = 0;
        if (ts.bnd_total > 0) base = atomicAdd(&sc->pool_count, ts.bnd_total);
        s_base = base;
        D->n_bnd = ts.bnd_total; D->overflow = (base + ts.bnd_total > MAX_BND_TOTAL) ? 1 : 0; D->bnd_base = base; D->pad2 = 0;
        D->head_b = BIN_NONE; D->head_d0 = 0; D->head_d1 = 0; D->pad = 0;
        first_kind[SCAN_THREADS / 32] = 2;                                           // past the tile end: closes the last piece
    }
    __syncthreads();
    if (s_base + ts.bnd_total <= MAX_BND_TOTAL) {
        BndEntry* B = pool + s_base;
        // kind of the element after this thread's last one
        int next_first = __shfl_down_sync(0xffffffffu, ts.cls[0].kind, 1);
        if ((threadIdx.x & 31) == 31) next_first = first_kind[(threadIdx.x >> 5) + 1];
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            if (ts.cls[j].kind == 2) { B[ts.bnd_before[j]].w = w[j]; B[ts.bnd_before[j]].idx = (int)threadIdx.x * EPT + j; }
            // piece ends: the next element is a boundary, or this is the last element of the tile
            const int nk = (j + 1 < EPT) ? ts.cls[(j + 1) & (EPT - 1)].kind : next_first;
            if (nk == 2) {
                const SegE& e = ts.inc[j];
                if (e.head < 0) { D->head_b = e.f.b; D->head_d0 = e.f.d0; D->head_d1 = e.f.d1; D->pad = e.f.pad; }
                else {
                    const int sl = slot_of[e.head];
                    B[sl].b = e.f.b; B[sl].d0 = e.f.d0; B[sl].d1 = e.f.d1;
                    if (e.f.pad) D->overflow = 1;
                }
            }
        }
    }
    // Per element, what k_cdf needs: the running sum after element i is  bits(base) + d0 (+ delta if base's mantissa is odd),
    // base = the exact sum at the tile start (ref 0) or at the tile's boundary element `ref - 1` -- the composed rounding
    // function since that point, whose two branches differ by -1, 0 or +1 (two walks that start one ulp apart stay 0, 1 or 2
    // apart).  10 bytes per element instead of a second tile scan.
    {
        long long d0v[EPT]; unsigned short cv[EPT];
#pragma unroll
        for (int j = 0; j < EPT; ++j) {
            long long d0 = 0; int delta = 0; unsigned ref;
            if (ts.cls[j].kind == 2) ref = (unsigned)ts.bnd_before[j] + 1u;
            else {
                const SegE& e = ts.inc[j];
                if (e.f.b != BIN_NONE) { d0 = e.f.d0; delta = (int)(e.f.d1 - e.f.d0); }
                ref = e.head < 0 ? 0u : (unsigned)slot_of[e.head] + 1u;
            }
            d0v[j] = d0; cv[j] = (unsigned short)(ref | ((unsigned)(delta + 1) << 12));
        }
        if (i0 + EPT <= n) {
            *reinterpret_cast<longlong2*>(A.el_d0 + i0) = make_longlong2(d0v[0], d0v[1]);
            *reinterpret_cast<longlong2*>(A.el_d0 + i0 + 2) = make_longlong2(d0v[2], d0v[3]);
            *reinterpret_cast<ushort4*>(A.el_code + i0) = make_ushort4(cv[0], cv[1], cv[2], cv[3]);
        } else {
#pragma unroll
            for (int j = 0; j < EPT; ++j) if (i0 + j < n) { A.el_d0[i0 + j] = d0v[j]; A.el_code[i0 + j] = cv[j]; }
        }
    }
}

// one CTA per rank: ESS, resample decision, and the exact value of the running sum at every tile start / boundary element of
// this rank.  The rank scans ITS tiles' descriptors into one summary -- the rounding function of everything before its first
// boundary element, then per boundary element its weight and the function of the run after it -- publishes it to every rank
// (exchange X_C), and walks the summaries of the ranks in front of it and its own with real fp64 adds.
struct ChainSmem { double bw[MAX_BND_TOTAL]; SegF seg[MAX_BND_TOTAL]; ChainEntry stage[MAX_BND_TOTAL]; };

struct ChainArgs {
    const double* tile_sq; const TileDesc* desc; const BndEntry* pool;   // this rank's tiles
    unsigned n_tiles, tile_ofs;      // local tile count; index of local tile 0 in the global tile-start table
    int* tile_bofs; SegF* tile_pre;  // [local tiles]
    double* sb_out;                  // [MAX_BND_TOTAL] exact running sum at every local boundary element
    unsigned long long min_effective; int decide, force;
    unsigned long long n; double* cdf;   // sequential fallback: local particle count, running-sum array
    ResultBox* box; unsigned long long* box_seq;   // host-mapped result mailbox and its device-side sequence counter
};

// Everything the caller's UpdateResult holds is final: publish it to the host (fields, system fence, then the sequence number).
__device__ __forceinline__ void publish_result(const ChainArgs& A, DevScalars* sc) {
    if (!A.box) return;
    const unsigned long long q = *A.box_seq + 1; *A.box_seq = q;
    ResultBox* b = A.box;
    b->ess = sc->ess; b->max_weight = sc->max_weight; b->floating_weight = sc->floating_weight; b->sum_w = sc->sum_w;
    b->data_particles = sc->data_particles; b->total_points = sc->total_points;
    b->do_resample = sc->do_resample; b->error_flags = *(volatile int*)&sc->error_flags;
    __threadfence_system();
    *(volatile unsigned long long*)&b->seq = q;
}

// the running sum through one rank's summary (entries in shared memory); sb != nullptr: record the sum at every boundary
__device__ __forceinline__ double chain_walk(double S, const SegF& head, const ChainEntry* e, int n_bnd, int& bad) {
    S = segf_apply(S, head, bad);
    for (int k = 0; k < n_bnd; ++k) {
        S = S + e[k].w;
        SegF f; f.d0 = e[k].d0; f.d1 = e[k].d1; f.b = e[k].b; f.pad = e[k].pad;
        S = segf_apply(S, f, bad);
    }
    return S;
}

__global__ void __launch_bounds__(1024, 1) k_chain(StateBufs st, DevScalars* sc, const ChainArgs A, const ShardPlan sp,
                                                   const __grid_constant__ Xchg X) {
    pdl_start();
    extern __shared__ __align__(16) unsigned char chain_raw[];
    ChainSmem& cs = *reinterpret_cast<ChainSmem*>(chain_raw);
    __shared__ int sh_i[64];             // [0,32): warp totals, [32,64): their inclusive scan
    __shared__ SegE sh_e[64];
    __shared__ int s_bad, s_go, s_late;
    __shared__ SegF s_head;
    __shared__ __align__(8) RankChainHead s_hdr[ESLAM_MAX_RANKS];
    __shared__ int s_ofs[ESLAM_MAX_RANKS + 1];
    __shared__ double s_sin, cs_own_start;
    __shared__ unsigned s_epoch;
    __shared__ double red[32];
    const unsigned tid = threadIdx.x;
    const unsigned n_tiles = A.n_tiles;
    double* tile_sin = reinterpret_cast<double*>(X.buf[X.me] + XOFF_TSIN);
    // this rank's part of  effective = sum w^2  (src/ParticleFilter.hpp:58-66): pairwise tree over its tiles
    const double rank_sq = rank_tree_sum_1024(A.tile_sq, n_tiles, red, A.sb_out);
    if (tid == 0) { s_bad = 0; s_late = 0; sc->cdf_seq = 0; s_epoch = sc->x_epoch[X_C] + 1; s_head = segf_identity(); }
    __syncthreads();
    // is a resampling possible at all in this call (every rank answers the same)?  Otherwise only the ESS is exchanged.
    const bool want_chain = A.force || (A.decide && A.min_effective > 0);

    // Rounds of 1024 tiles, one tile per thread (coalesced, independent loads): per round an exclusive scan of the boundary
    // counts and a segmented exclusive scan of the tile functions (shuffles inside the warp, warp 0 across the warp totals),
    // continued from round to round by a carry.  Each thread then places its tile's boundary elements and segment
    // functions into the chain arrays straight away.
    // The chain runs over one segment of tiles -- all of this rank's (one GPU, exact-global sharding) -- or, when the shards
    // resample locally (LOCAL_MIGRATE and no shard has collapsed), once per shard held here, each restarting from a zero sum.
    const unsigned lane = tid & 31, warp = tid >> 5;
    __shared__ int s_carry_c; __shared__ SegE s_carry_e;
    const bool local = sp.mode && !sc->migrate;
    const unsigned n_seg = !want_chain ? 0u : (local ? sp.count : 1u);
    int bnd_base = 0;
    for (unsigned sg = 0; sg < n_seg; ++sg) {
        const unsigned t_begin = local ? sg * sp.tiles_per_shard : 0u;
        const unsigned t_end = local ? t_begin + sp.tiles_per_shard : n_tiles;
        __syncthreads();
        if (tid == 0) { s_carry_c = bnd_base; s_carry_e.f = segf_identity(); s_carry_e.head = -1; s_head = segf_identity(); }
        __syncthreads();
        for (unsigned r0 = t_begin; r0 < t_end; r0 += 1024) {
            const unsigned t = r0 + tid;
            const bool have = t < t_end;
            TileDesc D; D.n_bnd = 0; D.overflow = 0; D.head_b = BIN_NONE; D.pad = 0; D.head_d0 = 0; D.head_d1 = 0; D.bnd_base = 0; D.pad2 = 0;
            if (have) D = A.desc[t];
            const BndEntry* B = A.pool + D.bnd_base;
            if (D.overflow) atomicOr(&s_bad, 1);
            const int cnt = D.overflow ? 0 : D.n_bnd;
            SegE e;                                                          // what the tile does to the running function
            if (cnt > 0) { const BndEntry& Lb = B[cnt - 1]; e.head = (int)t; e.f.b = Lb.b; e.f.d0 = Lb.d0; e.f.d1 = Lb.d1; e.f.pad = 0; }
            else { e.head = -1; e.f.b = D.head_b; e.f.d0 = D.head_d0; e.f.d1 = D.head_d1; e.f.pad = D.pad; }
            int inc_c = cnt; SegE inc_e = e;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) {
                const SegE v = sege_shfl_up(inc_e, o); const int cv = __shfl_up_sync(0xffffffffu, inc_c, o);
                if (lane >= (unsigned)o) { inc_e = sege_combine(v, inc_e); inc_c += cv; }
            }
            if (lane == 31) { sh_i[warp] = inc_c; sh_e[warp] = inc_e; }
            __syncthreads();
            if (warp == 0) {
                int wc = sh_i[lane]; SegE we = sh_e[lane];
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) {
                    const SegE v = sege_shfl_up(we, o); const int cv = __shfl_up_sync(0xffffffffu, wc, o);
                    if (lane >= (unsigned)o) { we = sege_combine(v, we); wc += cv; }
                }
                sh_i[32 + lane] = wc; sh_e[32 + lane] = we;                  // inclusive over warps
            }
            __syncthreads();
            // exclusive value of this thread: carry o (all earlier warps) o (earlier lanes of this warp)
            SegE run_e = sege_shfl_up(inc_e, 1); int run_c = __shfl_up_sync(0xffffffffu, inc_c, 1);
            if (lane == 0) { run_e.f = segf_identity(); run_e.head = -1; run_c = 0; }
            if (warp > 0) { run_e = sege_combine(sh_e[32 + warp - 1], run_e); run_c += sh_i[32 + warp - 1]; }
            run_e = sege_combine(s_carry_e, run_e); run_c += s_carry_c;
            const SegE round_e = sh_e[32 + 31]; const int round_c = sh_i[32 + 31];
            if (have) {
                A.tile_bofs[t] = run_c; A.tile_pre[t] = run_e.f;
                // boundary k = run_c + j: weight, and the function of the segment that follows it
                for (int j = 0; j < cnt; ++j) {
                    if (run_c + j < MAX_BND_TOTAL) {
                        cs.bw[run_c + j] = B[j].w;
                        if (j + 1 < cnt) { SegF f; f.b = B[j].b; f.d0 = B[j].d0; f.d1 = B[j].d1; f.pad = 0; cs.seg[run_c + j] = f; }
                    }
                }
                if (cnt > 0) {
                    // the run in front of this tile's first boundary element ends with the tile's head piece: it closes the
                    // segment after boundary run_c-1, or -- for the first boundary of the walk -- it is the walk's head function
                    SegF hf; hf.b = D.head_b; hf.d0 = D.head_d0; hf.d1 = D.head_d1; hf.pad = D.pad;
                    const SegF cf = segf_compose(run_e.f, hf);
                    if (run_c > bnd_base) { if (run_c - 1 < MAX_BND_TOTAL) cs.seg[run_c - 1] = cf; }
                    else s_head = cf;
                }
            }
            __syncthreads();                                                 // everyone has read the carry and the round totals
            if (tid == 0) { s_carry_e = sege_combine(s_carry_e, round_e); s_carry_c += round_c; }
            __syncthreads();
        }
        const int total = s_carry_c;
        if (tid == 0) {
            if (total > MAX_BND_TOTAL) s_bad = 1;
            else if (total > bnd_base) cs.seg[total - 1] = s_carry_e.f;     // the run after the last boundary element, to the end of the walk
            else s_head = s_carry_e.f;                                       // no boundary element at all
            sc->n_boundaries = total;
        }
        __syncthreads();
        if (local) {
            // this shard's own walk from a zero sum: nothing to exchange
            if (!s_bad && tid == 0) {
                int b2 = 0;
                double S = segf_apply(0.0, s_head, b2);
                for (int k = bnd_base; k < total; ++k) { S = S + cs.bw[k]; cs.bw[k] = S; S = segf_apply(S, cs.seg[k], b2); }
                sp.shard_tot[sp.first + sg] = S;                             // the shard's total: what its strata are scaled by
                if (b2) s_bad = 1;
            }
            __syncthreads();
            if (!s_bad) {
                int b3 = 0;
                for (unsigned t = t_begin + tid; t < t_end; t += blockDim.x) {
                    const int ofs = A.tile_bofs[t];
                    tile_sin[A.tile_ofs + t] = segf_apply(ofs > bnd_base ? cs.bw[ofs - 1] : 0.0, A.tile_pre[t], b3);
                }
                if (b3) atomicOr(&s_bad, 1);
            }
            __syncthreads();
        }
        bnd_base = total;
        if (s_bad) break;
    }
    __syncthreads();
    const int n_bnd = bnd_base;

    // ---- exchange: every rank's sum of w^2 (-> ESS, the decision) and, for the exact-global walk, its chain summary
    const bool send_entries = want_chain && !local;
    if (send_entries) {
        for (unsigned k = tid; k < (unsigned)n_bnd && k < MAX_BND_TOTAL; k += blockDim.x) {
            ChainEntry ce; ce.w = cs.bw[k]; ce.d0 = cs.seg[k].d0; ce.d1 = cs.seg[k].d1; ce.b = cs.seg[k].b; ce.pad = cs.seg[k].pad;
            cs.stage[k] = ce;
            if (ce.pad) atomicOr(&s_bad, 1);
        }
        __syncthreads();
    }
    if (tid == 0) {
        RankChainHead h; memset(&h, 0, sizeof(h));
        if (s_head.pad) s_bad = 1;
        h.sum_sq = rank_sq;
        h.n_bnd = (send_entries && !s_bad) ? n_bnd : 0; h.bad = send_entries ? s_bad : 0;
        h.head_d0 = s_head.d0; h.head_d1 = s_head.d1; h.head_b = s_head.b; h.head_pad = s_head.pad;
        s_hdr[X.me] = h;
    }
    __syncthreads();
    if (X.world > 1) {
        const size_t slot = XOFF_C + (size_t)X.me * XC_SLOT;
        ll_publish(X, slot, &s_hdr[X.me], XH_WORDS, s_epoch);
        if (X.world > 1 && s_hdr[X.me].n_bnd > 0) ll_publish(X, slot + XH_WORDS * 8, cs.stage, (unsigned)s_hdr[X.me].n_bnd * XE_WORDS, s_epoch);
        __syncthreads();                                   // the own header in shared memory is complete before it is overwritten below
        // headers of all ranks (the own one travels the same way)
        unsigned* hw = reinterpret_cast<unsigned*>(s_hdr);
        for (unsigned i = tid; i < (unsigned)X.world * XH_WORDS; i += blockDim.x) {
            const unsigned q = i / XH_WORDS, k = i - q * XH_WORDS;
            hw[q * (sizeof(RankChainHead) / 4) + k] = ll_read(X.buf[X.me] + XOFF_C + (size_t)q * XC_SLOT + (size_t)k * 8, s_epoch, &sc->error_flags);
        }
        __syncthreads();
    }
    if (tid == 0) {
        sc->x_epoch[X_C] = s_epoch;
        // effective sample size over all ranks (src/ParticleFilter.hpp:69) and the decision (src/PoseEstimator.cpp:249)
        if (A.decide) {
            double v[ESLAM_MAX_RANKS];
            for (int q = 0; q < X.world; ++q) v[q] = s_hdr[q].sum_sq;
            const double ess = 1.0 / tree_small(v, X.world);
            sc->ess = ess;
            sc->do_resample = (A.force || ess < (double)A.min_effective) ? 1 : 0;
        } else sc->do_resample = A.force ? 1 : 0;
        s_go = sc->do_resample;
        int any = 0, ofs = 0;
        for (int q = 0; q < X.world; ++q) { any |= s_hdr[q].bad; s_ofs[q] = ofs; if (q != X.me) ofs += s_hdr[q].n_bnd; }
        s_ofs[X.world] = ofs;                                            // boundary elements of all the other ranks
        if (!local) s_bad = any;                                         // every rank sees the same flags: same decision
        s_sin = 0.0;
        sc->rank_start_valid = 0;
    }
    __syncthreads();
    if (!s_go) { if (tid == 0) publish_result(A, sc); return; }

    if (!local && !s_bad) {
        // The walk: through the summaries of ALL ranks in order (the other ranks' entries polled into shared memory, the own
        // ones already there), one real fp64 add per boundary element, one integer add per run.  It yields the running sum in
        // front of every rank (rank_start: lets the resampling kernel tell which ranks hold its ancestors) and, inside the own
        // rank, at every boundary element.
        const int M = s_ofs[X.world];
        const bool at_once = M <= MAX_BND_TOTAL;                         // cs.stage is free again (the own entries live in cs.bw / cs.seg)
        if (at_once) {
            for (int q = 0; q < X.world; ++q) {
                if (q == X.me) continue;
                const unsigned char* src = X.buf[X.me] + XOFF_C + (size_t)q * XC_SLOT + XH_WORDS * 8;
                unsigned* d = reinterpret_cast<unsigned*>(cs.stage + s_ofs[q]);
                const unsigned words = (unsigned)s_hdr[q].n_bnd * XE_WORDS;
                for (unsigned i = tid; i < words; i += blockDim.x) d[i] = ll_read(src + (size_t)i * 8, s_epoch, &sc->error_flags);
            }
            __syncthreads();
        }
        for (int q = 0; q < X.world; ++q) {
            if (!at_once && q != X.me) {                                 // long lists: one rank at a time
                const unsigned char* src = X.buf[X.me] + XOFF_C + (size_t)q * XC_SLOT + XH_WORDS * 8;
                unsigned* d = reinterpret_cast<unsigned*>(cs.stage);
                const unsigned words = (unsigned)s_hdr[q].n_bnd * XE_WORDS;
                __syncthreads();
                for (unsigned i = tid; i < words; i += blockDim.x) d[i] = ll_read(src + (size_t)i * 8, s_epoch, &sc->error_flags);
                __syncthreads();
            }
            if (tid == 0) {
                int b2 = 0;
                double S = s_sin;
                sc->rank_start[q] = S;
                if (q == X.me) {
                    cs_own_start = S;
                    S = segf_apply(S, s_head, b2);
                    for (int k = 0; k < n_bnd; ++k) { S = S + cs.bw[k]; cs.bw[k] = S; S = segf_apply(S, cs.seg[k], b2); }
                } else {
                    SegF hf; hf.d0 = s_hdr[q].head_d0; hf.d1 = s_hdr[q].head_d1; hf.b = s_hdr[q].head_b; hf.pad = s_hdr[q].head_pad;
                    S = chain_walk(S, hf, cs.stage + (at_once ? s_ofs[q] : 0), s_hdr[q].n_bnd, b2);
                }
                s_sin = S;
                if (b2) s_late = 1;
                if (q + 1 == X.world) { sc->rank_start[X.world] = S; sc->rank_start_valid = 1; }
            }
        }
        __syncthreads();
        if (tid == 0) s_sin = cs_own_start;                              // the running sum in front of this rank
        __syncthreads();
        int b3 = 0;
        for (unsigned t = tid; t < n_tiles; t += blockDim.x) {
            const int ofs = A.tile_bofs[t];
            const double v = segf_apply(ofs > 0 ? cs.bw[ofs - 1] : s_sin, A.tile_pre[t], b3);
            for (int q = 0; q < X.world; ++q) reinterpret_cast<double*>(X.buf[q] + XOFF_TSIN)[A.tile_ofs + t] = v;   // every rank's tile-start table
        }
        if (b3) atomicOr(&s_late, 1);
        __syncthreads();
        // a consistency check that failed after the exchange cannot be turned into a fallback every rank agrees on; it is a
        // "cannot happen" (the element classes guarantee the binades): one GPU falls back, a sharded run reports it
        if (s_late) { if (X.world > 1) { if (tid == 0) atomicOr(&sc->error_flags, ERR_INTERNAL_BIT); } else if (tid == 0) s_bad = 1; }
        __syncthreads();
    }
    // the decision and every consistency check of the walk are in: the host may have the step's result while the running sum,
    // the resampling copy (and a sequential fallback below) are still to run
    if (tid == 0) publish_result(A, sc);
    if (s_bad) {
        // ---- fallback: the literal left-to-right running sum by one thread (boundary pool overflow, ambiguous binades).  Sharded
        // exact-global walk: rank r continues from the sum rank r-1 ended with (exchange X_SEQ), one rank after the other.
        if (tid == 0) {
            const unsigned e = sc->x_epoch[X_SEQ] + 1;
            s_epoch = e;
            double s_in = 0.0;
            if (!local && X.me > 0) {                                    // the sum the rank in front ended with
                const unsigned char* p = X.buf[X.me] + XOFF_SEQ + (size_t)(X.me - 1) * 16;
                const unsigned lo = ll_read(p, e, &sc->error_flags), hi = ll_read(p + 8, e, &sc->error_flags);
                s_in = __longlong_as_double((long long)(((unsigned long long)hi << 32) | lo));
            }
            const double* w = st.f[sc->cur][F_W];
            const unsigned long long seg_n = local ? sp.n_shard : A.n;
            const bool continues = !local && X.me > 0;
            double s = 0.0;
            for (unsigned long long i = 0; i < A.n; ++i) {
                const bool first = (i % seg_n) == 0;
                if (first) {
                    if (local && i > 0) sp.shard_tot[sp.first + (unsigned)(i / seg_n) - 1] = s;
                    s = continues ? s_in : 0.0;
                }
                if ((i & (TILE - 1)) == 0) tile_sin[A.tile_ofs + i / TILE] = s;
                s = (first && !continues) ? w[i] : s + w[i];
                A.cdf[i] = s;
            }
            if (local) sp.shard_tot[sp.first + (unsigned)((A.n - 1) / seg_n)] = s;
            s_sin = s;
            sc->cdf_seq = 1; sc->fallbacks += 1;
        }
        __syncthreads();
        if (!local) {
            ll_publish(X, XOFF_SEQ + (size_t)X.me * 16, &s_sin, 2, s_epoch);
            if (X.world > 1) {
                for (unsigned t = tid; t < n_tiles; t += blockDim.x) {
                    const double v = tile_sin[A.tile_ofs + t];
                    for (int q = 0; q < X.world; ++q) if (q != X.me) reinterpret_cast<double*>(X.buf[q] + XOFF_TSIN)[A.tile_ofs + t] = v;
                }
            }
        }
        __syncthreads();
        if (tid == 0) sc->x_epoch[X_SEQ] = s_epoch;
        __threadfence_system();
        return;
    }
    for (unsigned k = tid; k < (unsigned)n_bnd; k += blockDim.x) A.sb_out[k] = cs.bw[k];
    if (X.world > 1) __threadfence_system();                             // the tile starts pushed to the peers land before this kernel ends
}

// exact running sum of every element: cdf[i] == the reference's sum_w after adding xi_k[i].weight.  A streaming pass: the sum
// at the element's base (tile start, or one of the tile's boundary elements: k_chain) plus the integer offset k_norm_scan left.
struct CdfArgs {
    const int* tile_bofs; const double* sb; double* cdf;
    const long long* el_d0; const unsigned short* el_code;
    unsigned tile_ofs;
};
__global__ void __launch_bounds__(SCAN_THREADS) k_cdf(DevScalars* sc, unsigned long long n, const CdfArgs A, const __grid_constant__ Xchg X) {
    pdl_start();
    if (!sc->do_resample || sc->cdf_seq) return;                     // the same decision on every rank; sequential fallback: k_chain wrote it
    const unsigned long long i0 = (unsigned long long)blockIdx.x * TILE + threadIdx.x * EPT;
    if (i0 >= n) return;
    const int ofs = A.tile_bofs[blockIdx.x];
    const double s_in = reinterpret_cast<const double*>(X.buf[X.me] + XOFF_TSIN)[A.tile_ofs + blockIdx.x];
    long long d0[EPT]; unsigned short code[EPT];
    const bool full = i0 + EPT <= n;
    if (full) {
        const longlong2 a = *reinterpret_cast<const longlong2*>(A.el_d0 + i0), b = *reinterpret_cast<const longlong2*>(A.el_d0 + i0 + 2);
        const ushort4 c = *reinterpret_cast<const ushort4*>(A.el_code + i0);
        d0[0] = a.x; d0[1] = a.y; d0[2] = b.x; d0[3] = b.y; code[0] = c.x; code[1] = c.y; code[2] = c.z; code[3] = c.w;
    } else {
#pragma unroll
        for (int j = 0; j < EPT; ++j) { const bool in = i0 + j < n; d0[j] = in ? A.el_d0[i0 + j] : 0; code[j] = in ? A.el_code[i0 + j] : (unsigned short)(1u << 12); }
    }
    double S[EPT];
#pragma unroll
    for (int j = 0; j < EPT; ++j) {
        const unsigned ref = code[j] & 0xFFFu; const long long delta = (long long)((code[j] >> 12) & 3u) - 1;
        const double base = ref ? A.sb[ofs + (int)ref - 1] : s_in;
        const uint64_t u = em::d2u(base);
        S[j] = em::u2d(u + (uint64_t)(d0[j] + ((u & 1) ? delta : 0)));
    }
    store4(A.cdf, i0, n, S);
}

// ------------------------------------------------------------------------------------------------ resample
// The stratified walk (src/ParticleFilter.hpp:96-105) as a lower-bound search, fused with the copy (:104):
// ancestor[k] = first idx with !(cdf[idx] < sum_r); out[k] = in[ancestor[k]].  tile_start[t] holds the exact running sum
// before the t-th 1024-tile (== cdf[1024 t - 1]), so the search is two-level: everything before the last tile whose start
// is below sum_r is below sum_r too, and the next tile's start is not.
// Resampling with the search in shared memory, on one GPU and sharded (exact-global mode) alike.  A CTA owns RS_THREADS
// consecutive strata; their draws lie in [k0/N, (k0+RS_THREADS)/N), so their ancestors lie in the tiles from the last one
// starting below k0/N to the last one starting below (k0+1024)/N: with balanced weights one or two tiles.  When that is at
// most RS_TILES tiles the CTA copies their exact running sums into shared memory -- coalesced, from the rank that owns each
// tile, over NVLink when that is a peer -- and every thread searches there; otherwise the two-level search runs on global
// (peer) memory.  The ancestor's state is copied straight from the owner's current buffers (:104).  Fused with
// xi_k.swap(xi_kp) (:107): the last CTA to finish flips the current buffer.  tile_start holds the exact running sum before
// every tile of every rank (k_chain); peers maps every rank's running-sum and state buffers (own pointers on one GPU).
#define RS_TILES 3
struct PeerTable {
    const double* cdf[ESLAM_MAX_RANKS];
    const double* f[ESLAM_MAX_RANKS][2][7];
    const uint8_t* b[ESLAM_MAX_RANKS][2][2];
    const float2* patch[ESLAM_MAX_RANKS][2];      // per-particle map patches and their origins (cfg.patch_size > 0)
    const int* porigin[ESLAM_MAX_RANKS][2];
};
// cloneMaps (src/PoseEstimator.cpp:31-45): every copy of a particle owns a copy of its map patch.  The destination buffers
// of this rank; the sources come from the PeerTable (the owner of the ancestor, over NVLink when that is another GPU).
struct PatchGather { int P; int pad; float2* patch[2]; int* porigin[2]; };

template <int RS_THREADS>
__global__ void __launch_bounds__(RS_THREADS, 2048 / RS_THREADS) k_resample_local(StateBufs st, DevScalars* sc, const double* __restrict__ tile_start,
                                                            unsigned n_tiles, unsigned tiles_per_rank, unsigned long long n,
                                                            unsigned long long n_global, unsigned long long base, uint32_t k0key,
                                                            uint32_t k1key, int systematic, uint32_t* ancestors, unsigned* ticket,
                                                            const __grid_constant__ PeerTable peers, const ShardPlan sp, unsigned tile_ofs,
                                                            unsigned my_rank, const __grid_constant__ Xchg X) {
    pdl_start();
    __shared__ double s_cdf[RS_TILES * TILE];
    if (!sc->do_resample) return;                                        // the same decision on every rank
    int world = X.world;
    const bool local = sp.mode && !sc->migrate;
    const unsigned tid = threadIdx.x;
    // "my running sums are final" (this kernel runs after this rank's k_cdf): CTA 0 tells every rank, so that the signal
    // travels while the other CTAs work.  Epoch = number of this resampling (equal on every rank).
    const unsigned xd_epoch = sc->resample_step + 1u;
    if (!local && blockIdx.x == 0) x_signal(X, X_D, xd_epoch);
    const unsigned long long kb = (unsigned long long)blockIdx.x * RS_THREADS, k = kb + tid;   // local stratum
    // Local resampling (LOCAL_MIGRATE sharding, no shard collapsed): the CTA's strata belong to ONE logical shard held here; they
    // walk that shard's own running sum (restarted at zero, tile starts at tile_ofs + s T_s of the table), scaled by the shard's
    // total, and their ancestors are that shard's own particles.  Everything below runs on the rebound quantities; in the exact
    // mode they are the arguments themselves and scale == 1.0 (a multiplication that changes no bit).
    const unsigned long long n_ctx = n;                                  // particles of this context
    unsigned long long po = 0, anc_base = 0, sbase = base + kb;          // sbase: index of the CTA's first stratum in its walk
    unsigned q_fixed = 0; double scale = 1.0;
    if (local) {
        const unsigned s = (unsigned)(kb / sp.n_shard);
        po = (unsigned long long)s * sp.n_shard; anc_base = base + po; sbase = kb - po;
        tile_start += tile_ofs + s * sp.tiles_per_shard;
        n_tiles = sp.tiles_per_shard; tiles_per_rank = sp.tiles_per_shard; n = sp.n_shard; n_global = sp.n_shard;
        q_fixed = my_rank; world = 1; scale = sp.shard_tot[sp.first + s];
    }
    // the tile starts are non-decreasing, so "the last tile starting below thr" is a count: one coalesced read of the table and
    // two block-wide counts instead of two dependent binary searches.  thr_lo / thr_hi bound every sum_r of this CTA.
    const double thr_lo = ((double)sbase / (double)n_global) * scale, thr_hi = ((double)(sbase + (unsigned long long)RS_THREADS) / (double)n_global) * scale;
    // Sharded exact-global walk: which ranks can hold this CTA's ancestors?  From the running sum in front of every rank
    // (k_chain): ranks q_a .. q_b.  Only their part of the tile-start table is searched, and the CTA waits only for those of
    // them that are other GPUs (their k_chain has pushed their tile starts, their k_cdf has finished) -- with balanced
    // weights that is nobody for all but the CTAs at the ends of the shard.
    unsigned t_first = 0, t_count = n_tiles;                             // slice of the table to count over
    if (world > 1) {
        unsigned q_a = 0, q_b = (unsigned)world - 1;
        if (sc->rank_start_valid) {
            unsigned ca = 0, cb = 0;
            for (int q = 0; q < world; ++q) { const double rs = sc->rank_start[q]; ca += rs < thr_lo; cb += rs < thr_hi; }
            q_a = ca ? ca - 1 : 0; q_b = cb ? cb - 1 : 0;
        }
        t_first = q_a * tiles_per_rank; t_count = (q_b - q_a + 1) * tiles_per_rank;
        if (tid < (unsigned)world && tid >= q_a && tid <= q_b && tid != my_rank) {
            const unsigned* f = x_flags(X.buf[X.me], X_D) + tid;
            const long long t0 = clock64();
            while ((int)(ld_acquire_sys(f) - xd_epoch) < 0) {
                __nanosleep(20);
                if (clock64() - t0 > 8000000000ll) { atomicOr(&sc->error_flags, ERR_XCHG_TIMEOUT_BIT); break; }
            }
        }
        __syncthreads();
    }
    unsigned c_lo = 0, c_hi = 0;
    for (unsigned t0 = 0; t0 < t_count; t0 += RS_THREADS) {
        const double ts = (t0 + tid < t_count) ? __ldcg(&tile_start[t_first + t0 + tid]) : 1.7e308;
        c_lo += (unsigned)__syncthreads_count(ts < thr_lo);
        c_hi += (unsigned)__syncthreads_count(ts < thr_hi);
    }
    c_lo += t_first; c_hi += t_first;                                    // the tiles in front of the slice all start below thr_lo
    const unsigned t_lo = c_lo == 0 ? 0u : c_lo - 1, t_hi = c_hi == 0 ? 0u : c_hi - 1;
    const bool in_smem = (t_hi - t_lo) < RS_TILES;
    const unsigned long long g0 = (unsigned long long)t_lo * TILE;                        // window [g0, g1) of the walk's elements
    const unsigned long long g1 = min(((unsigned long long)t_hi + 1) * TILE, n_global);
    if (in_smem) {
        // tile by tile: the owner of a tile is a 32-bit quotient, uniform over the CTA (no per-element 64-bit division)
        for (unsigned t = t_lo; t <= t_hi; ++t) {
            const unsigned q = world > 1 ? t / tiles_per_rank : 0u;
            const unsigned long long l0 = (unsigned long long)(t - q * tiles_per_rank) * TILE;   // owner-local start of the tile
            const unsigned cnt = (unsigned)min((unsigned long long)TILE, n - l0);                // short only at the very end (one GPU)
            const double* __restrict__ src = peers.cdf[world > 1 ? q : q_fixed] + po + l0;
            for (unsigned e = tid; e < cnt; e += RS_THREADS) s_cdf[(t - t_lo) * TILE + e] = src[e];
        }
        __syncthreads();
    }
    if (k < n_ctx) {
        const unsigned long long gk = systematic ? 0ull : base + k;
        const em::Philox4 r = em::philox4x32_10((uint32_t)gk, sc->resample_step, 0, stream_word(em::STREAM_RESAMPLE, gk), k0key, k1key);
        const double u = em::u01(r.v[0], r.v[1]);
        const double sum_r = (((double)(sbase + tid) + u) / (double)n_global) * scale;    // :98
        unsigned long long a;                                                             // index of the ancestor in the walk
        if (in_smem) {
            unsigned lo = 0, hi = (unsigned)(g1 - g0);                                    // first element with !(cdf < sum_r)
            while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if (s_cdf[mid] < sum_r) lo = mid + 1; else hi = mid; }
            a = g0 + lo;
        } else {
            unsigned tlo = t_lo, thi = min(t_hi + 1, n_tiles);               // the first tile that does not start below sum_r lies in (t_lo, t_hi + 1]
            while (tlo < thi) { const unsigned mid = (tlo + thi) >> 1; if (__ldcg(&tile_start[mid]) < sum_r) tlo = mid + 1; else thi = mid; }
            const unsigned t = tlo == 0 ? 0 : tlo - 1;
            const unsigned q = t / tiles_per_rank;
            const double* cdf_q = peers.cdf[world > 1 ? q : q_fixed] + po;
            unsigned long long lo = (unsigned long long)(t - q * tiles_per_rank) * TILE, hi = min(lo + TILE, n);   // owner-local
            while (lo < hi) { const unsigned long long mid = (lo + hi) >> 1; if (cdf_q[mid] < sum_r) lo = mid + 1; else hi = mid; }
            a = (unsigned long long)q * n + lo;           // lo == n: the first element of the next rank (its tile start is >= sum_r)
        }
        if (a >= n_global) a = n_global - 1;              // the reference walks off the end here; we stop at the last particle
        ancestors[k] = (uint32_t)(a + anc_base);
        const unsigned q = world > 1 ? (unsigned)(a >> 10) / tiles_per_rank : 0u;         // owner: shards are whole tiles (TILE == 1024)
        const unsigned long long la = a - (unsigned long long)q * n + po;
        const unsigned qs = world > 1 ? q : q_fixed;
        const int cur = sc->cur, nxt = cur ^ 1;
#pragma unroll
        for (int f = 0; f < 7; ++f) st.f[nxt][f][k] = peers.f[qs][cur][f][la];            // :104 copy keeps the weight (Q3)
#pragma unroll
        for (int f = 0; f < 2; ++f) st.b[nxt][f][k] = peers.b[qs][cur][f][la];
    }
    // xi_k.swap(xi_kp) :107 by the last CTA (every CTA has read sc->cur by then)
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(ticket, 1u) == gridDim.x - 1) { *ticket = 0; sc->cur ^= 1; sc->resample_step += 1; }
    }
}

// Resampling as a stream (no patches, one walk over all particles; one GPU or the exact-global walk of a sharded filter): a
// persistent CTA owns a run of consecutive strata [K0, K1) of this rank.  Strata and their ancestors advance together (the
// walk is monotone, src/ParticleFilter.hpp:96-105), so the CTA pulls the SOURCE tiles that hold its ancestors -- exact running
// sum + the 58 bytes of state of 1024 particles, 66 KB -- one after the other through a three-stage ring of bulk copies
// (cp.async.bulk + mbarrier: tiles i+1 and i+2 are on their way while tile i is searched), from the rank that owns each tile: its own
// memory or a peer's over NVLink.  Every source tile is read once per CTA that needs it.  For tile t the candidate strata are
// floor(N S_t) - 2 .. floor(N S_t+1) + 2 (S_t = exact sum before the tile); a stratum belongs to the tile iff
//   (t == 0 or S_t < sum_r) and (t last or !(cdf[last of t] < sum_r))      <=>  lower_bound(cdf, sum_r) lies in tile t
// so every stratum is served by exactly one tile; tiles without candidates are never loaded.  Outputs leave coalesced.
// Sharded: the running sum in front of every rank (k_chain) tells which ranks can hold the CTA's ancestors; it searches only
// their part of the tile-start table and waits only for those of them that are other GPUs ("my running sums are final", X_D,
// signalled by CTA 0 of this same kernel on every rank) -- with balanced weights only the CTAs at the ends of a shard wait.
#define RS_STAGE_BYTES (8 * TILE * 8 + 2 * TILE)          // cdf + 7 doubles + 2 bytes per particle of one source tile
#define RS_STAGES 3                                       // ring depth: two source tiles in flight while one is searched
#define RS_STREAM_SMEM (RS_STAGES * RS_STAGE_BYTES + 64)
#define RS_TS_CACHE 128                                   // tile starts of a CTA's source tiles kept in shared memory

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t mbar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(__cvta_generic_to_global(src)), "r"(bytes), "r"(mbar) : "memory");
}

struct StreamArgs {
    const double* tile_start;          // exact running sum before every tile of every rank (global tile index)
    unsigned n_tiles, tiles_per_rank;  // tiles of all ranks, of one rank
    unsigned long long n, n_global, base;   // strata (= particles) of this rank, of all ranks, global index of local stratum 0
    uint32_t k0key, k1key; int systematic;
    uint32_t* ancestors; unsigned* ticket;
    unsigned my_rank;
};

__global__ void __launch_bounds__(1024, 1) k_resample_stream(StateBufs st, DevScalars* sc, const StreamArgs A,
                                                             const __grid_constant__ PeerTable peers, const __grid_constant__ Xchg X) {
    pdl_start();
    extern __shared__ __align__(128) unsigned char rs_raw[];
    if (!sc->do_resample) return;                                              // the same decision on every rank
    const unsigned tid = threadIdx.x;
    const int cur = sc->cur, nxt = cur ^ 1;
    const unsigned long long n = A.n, base = A.base;
    const unsigned T = A.tiles_per_rank;
    const int world = X.world;
    const unsigned xd_epoch = sc->resample_step + 1u;                          // number of this resampling: equal on every rank
    if (blockIdx.x == 0) x_signal(X, X_D, xd_epoch);                           // this rank's k_cdf has finished (stream order)
    const uint32_t smem0 = (uint32_t)__cvta_generic_to_shared(rs_raw);
    const uint32_t mbar0 = smem0 + RS_STAGES * RS_STAGE_BYTES;                // the stages' mbarriers behind them
    const unsigned long long per = (((n + gridDim.x - 1) / gridDim.x) + 31ull) & ~31ull;
    const unsigned long long K0 = min(n, (unsigned long long)blockIdx.x * per), K1 = min(n, K0 + per);
    if (K0 < K1) {
        const double dn = (double)A.n_global;
        const double thr_lo = (double)(base + K0) / dn, thr_hi = (double)(base + K1) / dn;
        // ranks that can hold this CTA's ancestors, and the wait for those that are other GPUs
        unsigned t_first = 0, t_count = A.n_tiles;
        if (world > 1) {
            unsigned q_a = 0, q_b = (unsigned)world - 1;
            if (sc->rank_start_valid) {
                unsigned ca = 0, cb = 0;
                for (int q = 0; q < world; ++q) { const double rs = sc->rank_start[q]; ca += rs < thr_lo; cb += rs < thr_hi; }
                q_a = ca ? ca - 1 : 0; q_b = cb ? cb - 1 : 0;
            }
            t_first = q_a * T; t_count = (q_b - q_a + 1) * T;
            if (tid < (unsigned)world && tid >= q_a && tid <= q_b && tid != A.my_rank) {
                const unsigned* f = x_flags(X.buf[X.me], X_D) + tid;
                const long long t0 = clock64();
                while ((int)(ld_acquire_sys(f) - xd_epoch) < 0) {
                    __nanosleep(20);
                    if (clock64() - t0 > 8000000000ll) { atomicOr(&sc->error_flags, ERR_XCHG_TIMEOUT_BIT); break; }
                }
            }
            __syncthreads();
        }
        // source tiles of this CTA: from the last tile starting below K0/N to the last one starting below K1/N (two counts)
        unsigned c_lo = 0, c_hi = 0;
        for (unsigned t0 = 0; t0 < t_count; t0 += 1024) {
            const double ts = (t0 + tid < t_count) ? __ldcg(&A.tile_start[t_first + t0 + tid]) : 1.7e308;
            c_lo += (unsigned)__syncthreads_count(ts < thr_lo);
            c_hi += (unsigned)__syncthreads_count(ts < thr_hi);
        }
        c_lo += t_first; c_hi += t_first;                                      // the tiles in front of the slice all start below thr_lo
        const unsigned tA = c_lo == 0 ? 0u : c_lo - 1, tB = c_hi == 0 ? 0u : c_hi - 1;
        if (tid == 0) { for (int g = 0; g < RS_STAGES; ++g) mbar_init(mbar0 + 8u * g, 1); }
        // exact running sum before tile t.  The first tile of a rank: from the walk's rank starts (this rank's own memory) --
        // the table entry is pushed by that rank, which this CTA has not necessarily waited for (t may be one past its slice)
        const bool rs_valid = world > 1 && sc->rank_start_valid;
        auto tstart_mem = [&](unsigned t) -> double {
            if (rs_valid) { const unsigned q = t / T; if (q * T == t) return sc->rank_start[q]; }
            return __ldcg(&A.tile_start[t]);
        };
        // the starts of the CTA's own source tiles tA .. tB + 1 are asked for again and again (candidate ranges, two tiles ahead of
        // the ring): once into shared memory -- from global memory every look-ahead costs two dependent L2 round trips per tile
        __shared__ double s_ts[RS_TS_CACHE];
        for (unsigned i = tid; i < RS_TS_CACHE && tA + i < A.n_tiles && tA + i <= tB + 1; i += 1024) s_ts[i] = tstart_mem(tA + i);
        __syncthreads();
        auto tstart = [&](unsigned t) -> double { return (t - tA) < RS_TS_CACHE ? s_ts[t - tA] : tstart_mem(t); };
        // candidate strata of tile t inside [K0, K1) (local stratum numbers; uniform over the CTA)
        const double dbase = (double)base;
        auto range = [&](unsigned t, unsigned long long& klo, unsigned long long& khi) {
            klo = K0; khi = K1;
            if (t > 0) {
                const double f = floor(tstart(t) * dn) - 2.0 - dbase;
                if (f > (double)K0) klo = f >= (double)K1 ? K1 : (unsigned long long)f;
            }
            if (t + 1 < A.n_tiles) {
                const double c = floor(tstart(t + 1) * dn) + 3.0 - dbase;
                if (c < (double)K1) khi = c <= (double)K0 ? K0 : (unsigned long long)c;
            }
            if (klo > khi) klo = khi;
        };
        // A source tile is ten bulk copies (running sum, seven doubles, two bytes per particle) onto one mbarrier.  Issuing them
        // takes ~0.1 us each -- from one thread that was 1 us per tile with everyone waiting for that thread's warp -- so thread 0
        // arms the barrier with the byte count and, after a CTA barrier, lane 0 of warps 0..9 issues one copy each.
        auto tile_geometry = [&](unsigned t, unsigned& q, unsigned long long& e0, uint32_t& c8, uint32_t& c1) {
            q = world > 1 ? t / T : 0u;
            e0 = (unsigned long long)(t - q * T) * TILE;                       // owner-local first element
            const unsigned cnt = (unsigned)min((unsigned long long)TILE, n - e0);
            c8 = ((cnt + 1u) & ~1u) * 8u; c1 = (cnt + 15u) & ~15u;             // multiples of 16 bytes (buffers are padded)
        };
        auto arm = [&](unsigned t, int stage) {                               // one thread
            unsigned q; unsigned long long e0; uint32_t c8, c1; tile_geometry(t, q, e0, c8, c1);
            mbar_expect_tx(mbar0 + 8u * (uint32_t)stage, 8u * c8 + 2u * c1);
        };
        auto issue_part = [&](unsigned t, int stage, unsigned part) {         // part 0: running sum, 1..7: doubles, 8..9: bytes
            unsigned q; unsigned long long e0; uint32_t c8, c1; tile_geometry(t, q, e0, c8, c1);
            const uint32_t S = smem0 + (uint32_t)stage * RS_STAGE_BYTES, mb = mbar0 + 8u * (uint32_t)stage;
            if (part == 0) bulk_g2s(S, peers.cdf[q] + e0, c8, mb);
            else if (part < 8) bulk_g2s(S + part * TILE * 8u, peers.f[q][cur][part - 1] + e0, c8, mb);
            else bulk_g2s(S + 8u * TILE * 8u + (part - 8u) * TILE, peers.b[q][cur][part - 8] + e0, c1, mb);
        };
        const unsigned lane = tid & 31u, warp = tid >> 5;
        // the candidate ranges of the CTA's source tiles, once (thread j: tile tA + j): every thread walks them below
        __shared__ unsigned long long s_klo[RS_TS_CACHE], s_khi[RS_TS_CACHE];
        for (unsigned j = tid; j < RS_TS_CACHE && tA + j <= tB; j += 1024) { unsigned long long a, b; range(tA + j, a, b); s_klo[j] = a; s_khi[j] = b; }
        __syncthreads();
        // the next source tile after `from` that has candidate strata (tB + 1: none)
        auto next_tile = [&](unsigned from, unsigned long long& lo, unsigned long long& hi) -> unsigned {
            unsigned t = from; lo = hi = 0;
            while (t <= tB) {
                if (t - tA < RS_TS_CACHE) { lo = s_klo[t - tA]; hi = s_khi[t - tA]; } else range(t, lo, hi);
                if (hi > lo) break;
                ++t;
            }
            return t;
        };
        // Three tiles ahead: t is searched, tn and tm are on their way or about to be issued.  A stage's mbarrier is armed by
        // thread 0 BEFORE the CTA barrier that precedes the issue of its copies, so the top of an iteration only issues.
        unsigned long long klo = 0, khi = 0, nlo = 0, nhi = 0, mlo = 0, mhi = 0, qlo = 0, qhi = 0;
        unsigned t = next_tile(tA, klo, khi);
        unsigned tn = t <= tB ? next_tile(t + 1, nlo, nhi) : tB + 1;
        unsigned tm = tn <= tB ? next_tile(tn + 1, mlo, mhi) : tB + 1;
        if (tid == 0) { if (t <= tB) arm(t, 0); if (tn <= tB) arm(tn, 1); if (tm <= tB) arm(tm, 2); }
        __syncthreads();
        if (lane == 0 && warp < 10) { if (t <= tB) issue_part(t, 0, warp); if (tn <= tB) issue_part(tn, 1, warp); }
        int it = 0;
        while (t <= tB) {
            const int stage = it % RS_STAGES;
            if (tm <= tB && lane == 0 && warp < 10) issue_part(tm, (it + 2) % RS_STAGES, warp);   // that stage was read one iteration ago
            mbar_wait(mbar0 + 8u * (uint32_t)stage, (uint32_t)(it / RS_STAGES) & 1u);
            {
                const unsigned char* S = rs_raw + (size_t)stage * RS_STAGE_BYTES;
                const double* s_cdf = reinterpret_cast<const double*>(S);
                const unsigned q = world > 1 ? t / T : 0u;
                const unsigned long long e0 = (unsigned long long)(t - q * T) * TILE;
                const unsigned cnt = (unsigned)min((unsigned long long)TILE, n - e0);
                const unsigned long long g0 = (unsigned long long)q * n + e0;  // global index of the tile's first element
                const double start = tstart(t), last = s_cdf[cnt - 1];
                const bool first_tile = t == 0, last_tile = t + 1 == A.n_tiles;
                for (unsigned long long k = klo + tid; k < khi; k += 1024) {
                    const unsigned long long gk = A.systematic ? 0ull : base + k;
                    const em::Philox4 r = em::philox4x32_10((uint32_t)gk, sc->resample_step, 0, stream_word(em::STREAM_RESAMPLE, gk), A.k0key, A.k1key);
                    const double u = em::u01(r.v[0], r.v[1]);
                    const double sum_r = ((double)(base + k) + u) / dn;                          // :98
                    if ((first_tile || start < sum_r) && (last_tile || !(last < sum_r))) {       // the ancestor lies in this tile
                        unsigned lo = 0, hi = cnt;                                               // first element with !(cdf < sum_r)
                        while (lo < hi) { const unsigned mid = (lo + hi) >> 1; if (s_cdf[mid] < sum_r) lo = mid + 1; else hi = mid; }
                        if (lo >= cnt) lo = cnt - 1;      // the reference walks off the end here; we stop at the last particle
                        A.ancestors[k] = (uint32_t)(g0 + lo);
#pragma unroll
                        for (int f = 0; f < 7; ++f) st.f[nxt][f][k] = reinterpret_cast<const double*>(S + (size_t)(1 + f) * TILE * 8)[lo];   // :104 (Q3)
#pragma unroll
                        for (int f = 0; f < 2; ++f) st.b[nxt][f][k] = (S + (size_t)8 * TILE * 8 + (size_t)f * TILE)[lo];
                    }
                }
            }
            // the tile after tm goes into THIS stage at the top of the next iteration: found and armed before the barrier
            const unsigned tq = tm <= tB ? next_tile(tm + 1, qlo, qhi) : tB + 1;
            if (tid == 0 && tq <= tB) {
                asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the stage was read through the generic proxy
                arm(tq, stage);
            }
            __syncthreads();                                                   // everyone is done with this stage
            t = tn; klo = nlo; khi = nhi; tn = tm; nlo = mlo; nhi = mhi; tm = tq; mlo = qlo; mhi = qhi; ++it;
        }
    }
    // xi_k.swap(xi_kp) :107 by the last CTA (every CTA has read sc->cur by then)
    __syncthreads();
    if (tid == 0) {
        __threadfence();
        if (atomicAdd(A.ticket, 1u) == gridDim.x - 1) { *A.ticket = 0; sc->cur ^= 1; sc->resample_step += 1; }
    }
}

// cloneMaps (src/PoseEstimator.cpp:31-45, :252-253): every copy of a particle owns a copy of its map.  Runs after the
// resampling kernel (which has swapped the state buffers: sc->cur is the NEW buffer): one warp per particle copies the
// ancestor's patch and origin from the old buffer of the rank that owns the ancestor (a peer's memory over NVLink when that is
// another GPU; its old buffer stays untouched until its next resampling, which cannot start before this rank's next k_chain).
__global__ void __launch_bounds__(256) k_gather_patches(const DevScalars* sc, unsigned long long n, unsigned long long n_per_rank,
                                                        const uint32_t* __restrict__ anc, const PatchGather pg,
                                                        const __grid_constant__ PeerTable peers) {
    pdl_start();
    if (!sc->do_resample) return;
    const unsigned long long k = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (k >= n) return;
    const unsigned lane = threadIdx.x & 31;
    const int now = sc->cur, old = now ^ 1;
    const unsigned long long a = anc[k];
    const unsigned q = (unsigned)(a / n_per_rank);
    const unsigned long long la = a - (unsigned long long)q * n_per_rank;
    const unsigned pp = (unsigned)(pg.P * pg.P);
    const float2* src = peers.patch[q][old] + la * pp;
    float2* dst = pg.patch[now] + k * pp;
    if ((pp & 1u) == 0) {
        const float4* s4 = reinterpret_cast<const float4*>(src); float4* d4 = reinterpret_cast<float4*>(dst);
        const unsigned m = pp >> 1;
        for (unsigned c0 = 0; c0 < m; c0 += 128) {                        // four 16-byte loads in flight per lane
            float4 v[4];
#pragma unroll
            for (int j = 0; j < 4; ++j) { const unsigned c = c0 + j * 32 + lane; if (c < m) v[j] = s4[c]; }
#pragma unroll
            for (int j = 0; j < 4; ++j) { const unsigned c = c0 + j * 32 + lane; if (c < m) d4[c] = v[j]; }
        }
    } else for (unsigned c = lane; c < pp; c += 32) dst[c] = src[c];
    if (lane < 2) pg.porigin[now][2 * k + lane] = peers.porigin[q][old][2 * la + lane];
}

// ------------------------------------------------------------------------------------------------ processMap
// EmbodiedSlamFilter::processMap(scanMap, match, update) (src/EmbodiedSlamFilter.cpp:179-232) in per-particle-map mode: one
// warp per particle.  match: lane j evaluates the j-th sampled scan surface against the particle's map (patch first, then the
// shared grid, 3-sigma gate), lane 0 adds the hits in sample order; p.weight *= pow(weight, 0.1) (:214-221).  merge: the scan
// surfaces go through C_s2p 32 at a time in row-major order; lanes whose surfaces fall into the SAME patch cell apply their
// Kalman merges one after the other in lane order, so every cell sees its merges in the order of the sequential loop
// (:222-227).  Definition of match / merge: include/eslam_gpu.h eslam_gpu_process_map.
struct ScanCell { double x, y; float mean, stdev; };   // centre of a scan cell with a surface (scan frame), its {mean, stdev}
struct ProcessMapArgs {
    const ScanCell* cells; unsigned n_cells, n_samples;
    int match, update, P;
    double two_sigma2, visual;
    float2* patch[2]; int* porigin[2];
    const float2* grid; MapParams map;
};

__device__ __forceinline__ void patch_merge_cell(float2* cell, float zm, float sd) {
    float2 cv = *cell;
    if (!(cv.x == cv.x) || !(cv.y == cv.y)) { cv.x = zm; cv.y = sd; }
    else {                                                               // Kalman merge of two height estimates (float)
        const float v0 = cv.y * cv.y, v1 = sd * sd;
        const float kg = __fdiv_rn(v0, v0 + v1);
        cv.x = cv.x + kg * (zm - cv.x);
        cv.y = __fsqrt_rn((1.0f - kg) * v0);
    }
    *cell = cv;
}

__global__ void __launch_bounds__(256) k_process_map(StateBufs st, DevScalars* sc, unsigned long long n, const ProcessMapArgs A) {
    const unsigned long long i = ((unsigned long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    if (i >= n) return;                                                  // whole warps leave together
    const unsigned lane = threadIdx.x & 31;
    const int cur = sc->cur, P = A.P;
    const MapParams& M = A.map;
    const double x = st.f[cur][F_X][i], y = st.f[cur][F_Y][i], th = st.f[cur][F_TH][i];
    const double zpos = st.f[cur][F_Z][i], zs = st.f[cur][F_ZS][i];
    double s, c; em::sincos_(th, s, c);                                  // :190-193
    float2* patch = A.patch[cur] + i * (size_t)P * P;
    int* origin = A.porigin[cur] + 2 * i;
    int ox = origin[0], oy = origin[1];
    if (A.match && A.n_samples > 0) {                                    // :214-221
        double wj = 0.0; int hit = 0;
        if (lane < A.n_samples) {
            const ScanCell q = A.cells[(unsigned long long)lane * A.n_cells / A.n_samples];
            const double wx = em::fma_(c, q.x, em::fma_(-s, q.y, x)), wy = em::fma_(s, q.x, em::fma_(c, q.y, y));
            const double wz = (double)q.mean + zpos;
            const double qvar = (double)q.stdev * (double)q.stdev + zs * zs;
            double gx = wx, gy = wy;
            if (!M.identity) {
                gx = em::fma_(M.g2l[0], wx, em::fma_(M.g2l[1], wy, em::fma_(M.g2l[2], wz, M.g2l[3])));
                gy = em::fma_(M.g2l[4], wx, em::fma_(M.g2l[5], wy, em::fma_(M.g2l[6], wz, M.g2l[7])));
            }
            int ix, iy;
            if (em::cell_index_exact(gx, M.ox, M.sx, M.W, ix) && em::cell_index_exact(gy, M.oy, M.sy, M.H, iy)) {
                float2 cell = A.grid[(size_t)iy * M.W + ix];
                bool found = false;
                const unsigned lx = (unsigned)(ix - ox), ly = (unsigned)(iy - oy);
                if (ox != INT_MIN && lx < (unsigned)P && ly < (unsigned)P) {     // the particle's own map first
                    const float2 pc = patch[ly * P + lx];
                    if (pc.x == pc.x && pc.y == pc.y) {
                        const double dp = wz - (double)pc.x;
                        if (dp * dp < 9.0 * ((double)pc.y * (double)pc.y + qvar)) { cell = pc; found = true; }
                    }
                }
                if (!found && cell.x == cell.x && cell.y == cell.y) {
                    const double d0 = wz - (double)cell.x;
                    const double zvar = (double)cell.y * (double)cell.y + qvar;
                    found = d0 * d0 < 9.0 * zvar;
                }
                if (found) { const double d = wz - (double)cell.x; wj = em::exp_(-(d * d) / A.two_sigma2); hit = 1; }
            }
        }
        double wsum = 0.0; unsigned hits = 0;
        for (unsigned j = 0; j < A.n_samples; ++j) {                     // in sample order (every lane keeps the same sums)
            const double v = __shfl_sync(0xffffffffu, wj, j); const int h = __shfl_sync(0xffffffffu, hit, j);
            if (h) { wsum += v; hits++; }
        }
        if (lane == 0) {
            const float weight = hits ? (float)(wsum / (double)hits) : 1.0f;
            st.f[cur][F_W][i] *= em::exp_(A.visual * em::log_((double)weight));   // pow(weight, visualWeighting) :220
        }
    }
    if (!A.update) return;
    for (unsigned b0 = 0; b0 < A.n_cells; b0 += 32) {                    // :222-227
        const unsigned k = b0 + lane;
        bool valid = false; int ix = 0, iy = 0; float zm = 0.f, sd = 0.f;
        if (k < A.n_cells) {
            const ScanCell q = A.cells[k];
            const double wx = em::fma_(c, q.x, em::fma_(-s, q.y, x)), wy = em::fma_(s, q.x, em::fma_(c, q.y, y));
            double gx = wx, gy = wy;
            if (!M.identity) {
                gx = em::fma_(M.g2l[0], wx, em::fma_(M.g2l[1], wy, M.g2l[3]));
                gy = em::fma_(M.g2l[4], wx, em::fma_(M.g2l[5], wy, M.g2l[7]));
            }
            valid = em::cell_index_exact(gx, M.ox, M.sx, M.W, ix) && em::cell_index_exact(gy, M.oy, M.sy, M.H, iy);
            zm = (float)((double)q.mean + zpos);
            sd = (float)em::sqrt_((double)q.stdev * (double)q.stdev + zs * zs);
        }
        if (ox == INT_MIN && __any_sync(0xffffffffu, valid)) {           // anchor the patch on the particle's own cell
            int ax = 0, ay = 0; bool ok = false;
            if (lane == 0) {
                double gx = x, gy = y;
                if (!M.identity) { gx = em::fma_(M.g2l[0], x, em::fma_(M.g2l[1], y, M.g2l[3])); gy = em::fma_(M.g2l[4], x, em::fma_(M.g2l[5], y, M.g2l[7])); }
                ok = em::cell_index_exact(gx, M.ox, M.sx, M.W, ax) && em::cell_index_exact(gy, M.oy, M.sy, M.H, ay);
                if (ok) { origin[0] = ax - P / 2; origin[1] = ay - P / 2; }
            }
            ok = __shfl_sync(0xffffffffu, (int)ok, 0) != 0;
            if (ok) { ox = __shfl_sync(0xffffffffu, ax, 0) - P / 2; oy = __shfl_sync(0xffffffffu, ay, 0) - P / 2; }
        }
        const unsigned lx = (unsigned)(ix - ox), ly = (unsigned)(iy - oy);
        const bool act = valid && ox != INT_MIN && lx < (unsigned)P && ly < (unsigned)P;
        const unsigned target = act ? ly * (unsigned)P + lx : 0x80000000u | lane;
        const unsigned grp = __match_any_sync(0xffffffffu, target);
        const unsigned rank = __popc(grp & ((1u << lane) - 1u));
        unsigned max_rank = act ? rank : 0u;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) max_rank = max(max_rank, __shfl_xor_sync(0xffffffffu, max_rank, o));
        for (unsigned r = 0; r <= max_rank; ++r) {
            if (act && rank == r) patch_merge_cell(patch + target, zm, sd);
            __syncwarp();
        }
    }
}

// second moments of (x, y) about the weighted mean: per-tile sums of w dx^2, w dx dy, w dy^2 (eslam_gpu_pose_moments)
__global__ void __launch_bounds__(1024, 1) k_moment_tiles(StateBufs st, const DevScalars* sc, unsigned long long n, double mx, double my,
                                                          double* part /* n_tiles x 5 */) {
    __shared__ double red[32];
    const int cur = sc->cur;
    const unsigned long long i = (unsigned long long)blockIdx.x * TILE + threadIdx.x;
    double cxx = 0, cxy = 0, cyy = 0;
    if (i < n) {
        const double w = st.f[cur][F_W][i], dx = st.f[cur][F_X][i] - mx, dy = st.f[cur][F_Y][i] - my;
        cxx = w * dx * dx; cxy = w * dx * dy; cyy = w * dy * dy;
    }
    const double a = tile_tree_sum_1024(cxx, red), b = tile_tree_sum_1024(cxy, red), c = tile_tree_sum_1024(cyy, red);
    if (threadIdx.x == 0) { double* o = part + (size_t)blockIdx.x * 5; o[0] = a; o[1] = b; o[2] = c; o[3] = 0; o[4] = 0; }
}

// ParticleFilter::resample_stratified(samples) / resample_multinomial(samples) (src/ParticleFilter.hpp:85-147) with a sample count
// of their own (one GPU, off the step path): one thread per sample, two-level lower-bound search of the exact running sum (tile
// starts, then the tile), copy of the ancestor (state and map patch).  Multinomial: one uniform per sample, weight 1 / old size.
struct ResampleNArgs {
    const double* cdf; const double* tile_start; unsigned n_tiles;
    unsigned long long n, samples; uint32_t k0key, k1key; int systematic, multinomial;
    uint32_t* ancestors; int P; float2* patch[2]; int* porigin[2];
};
__global__ void __launch_bounds__(256) k_resample_n(StateBufs st, const DevScalars* sc, const ResampleNArgs A) {
    const unsigned long long k = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= A.samples) return;
    const int cur = sc->cur, nxt = cur ^ 1;
    const unsigned long long gk = (A.systematic && !A.multinomial) ? 0ull : k;
    const em::Philox4 r = em::philox4x32_10((uint32_t)gk, sc->resample_step, 0, stream_word(em::STREAM_RESAMPLE, gk), A.k0key, A.k1key);
    const double u = em::u01(r.v[0], r.v[1]);
    const double sum_r = A.multinomial ? u : ((double)k + u) / (double)A.samples;    // :132 / :98
    unsigned tlo = 0, thi = A.n_tiles;                                               // first tile with !(start < sum_r)
    while (tlo < thi) { const unsigned mid = (tlo + thi) >> 1; if (__ldg(&A.tile_start[mid]) < sum_r) tlo = mid + 1; else thi = mid; }
    const unsigned t = tlo == 0 ? 0 : tlo - 1;
    unsigned long long lo = (unsigned long long)t * TILE, hi = min(lo + TILE, A.n);
    while (lo < hi) { const unsigned long long mid = (lo + hi) >> 1; if (__ldg(&A.cdf[mid]) < sum_r) lo = mid + 1; else hi = mid; }   // :99-103 / :137
    const unsigned long long a = lo >= A.n ? A.n - 1 : lo;
    A.ancestors[k] = (uint32_t)a;
#pragma unroll
    for (int f = 0; f < 7; ++f) st.f[nxt][f][k] = st.f[cur][f][a];                   // :104 / :139
    if (A.multinomial) st.f[nxt][F_W][k] = 1.0 / (double)A.n;                        // :140
#pragma unroll
    for (int f = 0; f < 2; ++f) st.b[nxt][f][k] = st.b[cur][f][a];
    if (A.P > 0) {
        const size_t pp = (size_t)A.P * A.P;
        const float2* src = A.patch[cur] + a * pp; float2* dst = A.patch[nxt] + k * pp;
        for (size_t c = 0; c < pp; ++c) dst[c] = src[c];
        A.porigin[nxt][2 * k] = A.porigin[cur][2 * a]; A.porigin[nxt][2 * k + 1] = A.porigin[cur][2 * a + 1];
    }
}
__global__ void k_after_resample_n(DevScalars* sc) { sc->cur ^= 1; sc->resample_step += 1; }   // xi_k.swap(xi_kp) :107 / :146

__global__ void k_bump_project(DevScalars* sc) { sc->project_step += 1; }   // eslam_gpu_project alone: count the project

__global__ void k_fill_i32(int* p, size_t count, int v) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < count) p[i] = v;
}

// ------------------------------------------------------------------------------------------------ outputs
__global__ void __launch_bounds__(1024, 1) k_centroid_tiles(StateBufs st, const DevScalars* sc, unsigned long long n,
                                                            unsigned long long base, double* part /* n_tiles x 5 */,
                                                            unsigned long long* best_idx, double* best_w) {
    __shared__ double red[32];
    __shared__ unsigned long long bi[32];
    const int cur = sc->cur;
    const unsigned long long i = (unsigned long long)blockIdx.x * TILE + threadIdx.x;
    double x = 0, y = 0, th = 0, z = 0, w = 0;
    if (i < n) { w = st.f[cur][F_W][i]; x = st.f[cur][F_X][i] * w; y = st.f[cur][F_Y][i] * w; th = st.f[cur][F_TH][i] * w; z = st.f[cur][F_Z][i] * w; }
    const double sx = tile_tree_sum_1024(x, red), sy = tile_tree_sum_1024(y, red);
    const double st_ = tile_tree_sum_1024(th, red), sz = tile_tree_sum_1024(z, red), sw = tile_tree_sum_1024(w, red);
    if (threadIdx.x == 0) {
        double* o = part + (size_t)blockIdx.x * 5;
        o[0] = sx; o[1] = sy; o[2] = st_; o[3] = sz; o[4] = sw;
    }
    // first index of the strict maximum (src/ParticleFilter.hpp:160-173)
    double bw = (i < n) ? st.f[cur][F_W][i] : -INFINITY; unsigned long long bidx = (i < n) ? base + i : ~0ull;
    if (!(bw == bw)) bw = -INFINITY;   // NaN never wins the > comparison
    for (int o = 1; o < 32; o <<= 1) {
        const double ow = __shfl_xor_sync(0xffffffffu, bw, o); const unsigned long long oi = __shfl_xor_sync(0xffffffffu, bidx, o);
        if (ow > bw || (ow == bw && oi < bidx)) { bw = ow; bidx = oi; }
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) { red[threadIdx.x >> 5] = bw; bi[threadIdx.x >> 5] = bidx; }
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < 32; ++k) if (red[k] > bw || (red[k] == bw && bi[k] < bidx)) { bw = red[k]; bidx = bi[k]; }
        best_idx[blockIdx.x] = bidx; best_w[blockIdx.x] = bw;
    }
}

__global__ void __launch_bounds__(1024, 1) k_centroid_final(const double* part, unsigned n_tiles, double* scratch_a,
                                                            double* scratch_b, const unsigned long long* best_idx,
                                                            const double* best_w, DevScalars* sc) {
    __shared__ double tree_a[TREE_SMEM], tree_b[TREE_SMEM / 2];
    for (int q = 0; q < 5; ++q) {
        for (unsigned t = threadIdx.x; t < n_tiles; t += blockDim.x) scratch_a[t] = part[(size_t)t * 5 + q];
        __syncthreads();
        const double v = block_tree_reduce(scratch_a, scratch_b, n_tiles, tree_a, tree_b);
        if (threadIdx.x == 0) sc->centroid[q] = v;
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        double bw = -INFINITY; unsigned long long bi = 0;
        for (unsigned t = 0; t < n_tiles; ++t) if (best_w[t] > bw) { bw = best_w[t]; bi = best_idx[t]; }
        sc->best_index = (bi == ~0ull) ? 0 : bi;
    }
}

// ------------------------------------------------------------------------------------------------ diagnostic seam
// ContactModel::evaluatePose / evaluateWeight / updateZPositionEstimate for explicit poses
// (the seam test/testContactModel.cpp drives); same device functions as the fused kernel.
struct EvalParams {
    unsigned long long n;
    const double *x, *y, *th, *z, *mv, *zvin;
    MapParams map; const em::dbl2* htab;
    double inv_corr; unsigned long long min_contacts; int use_shape;
    uint8_t* ok; double *weight, *zdelta, *zvar, *posevar; int32_t* ncont; double *zout, *zvout;
    int32_t* cells; double* cp; int* err;
    ContactSet cs;
};

__global__ void k_evaluate_poses(const __grid_constant__ EvalParams P) {
    const unsigned long long i = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n) return;
    MapTile mt; mt.s = nullptr; mt.w = 0; mt.h = 0; mt.x0 = 0; mt.y0 = 0;
    double s, c; em::sincos_(P.th[i], s, c);
    const double mv = P.mv[i];
    if (mv == 0) { atomicOr(P.err, ERR_ZERO_VARIANCE_BIT); return; }
    PoseEval ev;
    PatchRef pr; pr.cells = nullptr; pr.ox = INT_MIN; pr.oy = 0; pr.P = 0;
    contact_eval<true, false>(P.cs, P.map, mt, P.htab, P.inv_corr, P.x[i], P.y[i], P.z[i], s, c, mv, ev,
                              P.cells + i * P.cs.n, P.cp + i * ESLAM_MAX_CP * 5, pr, nullptr);
    const bool ok = (unsigned long long)ev.n_found >= P.min_contacts;
    double weight = 0, zdl = 0, zvr = 0, z = P.z[i], zv = P.zvin ? P.zvin[i] : 0.0;
    if (ok) {
        evaluate_weight(ev, P.use_shape, weight, zdl, zvr);
        if (P.zvin) update_z(ev, zdl, zvr, z, zv);
    }
    P.ok[i] = ok; P.weight[i] = weight; P.zdelta[i] = zdl; P.zvar[i] = zvr; P.posevar[i] = ev.poseVar;
    P.ncont[i] = ev.n_found; P.zout[i] = z; P.zvout[i] = zv;
}
