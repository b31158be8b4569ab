// eslam_gpu.cu -- C ABI of libeslam_gpu.so (include/eslam_gpu.h): context, memory, launch sequencing and the
// peer-memory bootstrap of the sharded (multi-GPU) filter (the step's exchanges run inside its kernels).  All arithmetic lives in eslam_kernels.cuh; there is no CPU
// implementation of any step in this library.
#include "eslam_kernels.cuh"
#include "eslam_hash.cuh"
#include "eslam_nccl.h"
#include <nvtx3/nvToolsExt.h>   // header-only; a range costs a few ns when no tool is attached

#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <chrono>
#include <algorithm>
#include <atomic>
#include <string>
#include <utility>
#include <vector>

static const em::dbl2 H_TABLE_HOST[ESLAM_H_TABLE_ENTRIES] = {
#include "eslam_h_table.inc"
};

static NcclApi g_nccl;

// NVTX ranges around the entry points of the step path (what nsys / ncu group the kernels under)
struct NvtxRange { explicit NvtxRange(const char* name) { nvtxRangePushA(name); } ~NvtxRange() { nvtxRangePop(); } };

// Development switches, read ONCE when a context is created (tools/README.md); never consulted on the step path.
struct EnvSwitches {
    bool no_fast, no_tma, no_local_resample, stream_resample, trace, timing, use_pdl, no_p2p, result_sync;
    int lean_threads, tile_override, graph;
    void read() {
        no_fast = getenv("ESLAM_GPU_NO_FAST") != nullptr;
        no_tma = getenv("ESLAM_GPU_NO_TMA") != nullptr;
        no_local_resample = getenv("ESLAM_GPU_NO_LOCAL_RESAMPLE") != nullptr;
        stream_resample = getenv("ESLAM_GPU_STREAM_RESAMPLE") == nullptr || atoi(getenv("ESLAM_GPU_STREAM_RESAMPLE")) != 0;   // default on
        trace = getenv("ESLAM_GPU_TRACE") != nullptr;
        timing = getenv("ESLAM_GPU_TIMING") != nullptr;
        use_pdl = getenv("ESLAM_GPU_NO_PDL") == nullptr;
        no_p2p = getenv("ESLAM_GPU_NO_P2P") != nullptr;
        result_sync = getenv("ESLAM_GPU_RESULT_SYNC") != nullptr && atoi(getenv("ESLAM_GPU_RESULT_SYNC")) != 0;
        lean_threads = getenv("ESLAM_GPU_LEAN_THREADS") ? atoi(getenv("ESLAM_GPU_LEAN_THREADS")) : 1024;
        graph = getenv("ESLAM_GPU_GRAPH") ? atoi(getenv("ESLAM_GPU_GRAPH")) : -1;   // -1: by size, 0: never, 1: always
        tile_override = getenv("ESLAM_GPU_TILE") ? atoi(getenv("ESLAM_GPU_TILE")) : 0;
    }
};

struct eslam_gpu_ctx {
    eslam_gpu_config cfg;
    eslam_gpu_shard shard;
    EnvSwitches env;
    int device;
    cudaStream_t stream;
    std::string error;
    uint64_t launches;
    eslam_ncclComm_t comm; bool has_comm;

    // particles (local shard)
    uint64_t n, cap;
    StateBufs st;
    DevScalars* sc; DevScalars* sc_host;          // sc_host: pinned staging
    DevScalars* sc_async;                         // pinned mirror refreshed after every update on a side stream (never waited for)
    ResultBox* box_host; ResultBox* box_dev;      // host-mapped pinned mailbox k_chain publishes the UpdateResult into (two views)
    unsigned long long* box_seq;                  // device-side sequence counter of the mailbox
    uint64_t chain_launched;                      // k_chain launches so far: the sequence number the newest result will carry
    cudaStream_t side; cudaEvent_t side_ev;
    uint32_t* ancestors; double* cdf;
    long long* el_d0; unsigned short* el_code;    // per element: offset of its running sum from its base (k_norm_scan -> k_cdf)
    double* hint; unsigned hint_n;                // per-CTA sums of x, y (tile placement hint), entries written by the last update
    unsigned* ticket;                             // "last CTA" counter of the resampling kernels
    bool last_lean;                               // the last update went through the lean path
    FastParams* fast;                             // the lean kernels' parameter block of the current update (per context: contexts are independent)
    int* bin_count; uint2* keys; uint32_t* bin_list; unsigned* bin_offs; unsigned* bin_next; int* cta_bin;   // lean path: bins, per-bin lists, chunk counters
    bool grid_sane;                               // every stdev of the grid is NaN or |stdev| < 2^49 (lean path precondition)
    float2* patch[2]; int* porigin[2];            // per-particle map patches (cfg.patch_size > 0), double buffered
    // per-tile arrays of this rank's tiles
    unsigned t_local, t_global, t_cap;
    TilePartA* l_part_a; double2* l_tile_sum; double* l_tile_sq; TileDesc* l_desc; BndEntry* l_pool;
    double* prefix_raw; double2* shard_sums;
    double* shard_tot;                            // LOCAL_MIGRATE: exact total of every logical shard held here (k_chain -> k_resample_local)
    double *tile_prefix, *scratch_a, *scratch_b, *sb; int* tile_bofs; SegF* tile_pre;
    double *l_cent, *g_cent, *l_best_w, *g_best_w; unsigned long long *l_best_idx, *g_best_idx;   // centroid / best particle (g_*: all ranks' tiles)
    // exchange buffer of this rank (exchange slots, epoch flags, the global tile-start table) and every rank's mapping of it;
    // peer-mapped running-sum and state buffers of the other ranks (CUDA IPC over NVLink): resampling reads ancestors in place
    unsigned char* xchg; size_t x_bytes; Xchg X;
    PeerTable peers; bool has_peers; std::vector<void*> ipc_opened;
    // map
    float2* cells; MapParams map; bool has_map;
    void* scan_cells; size_t scan_cap;                // processMap: the scan's surfaces on the device
    // SurfaceHash: candidate poses (creation order), their indices grouped by slope bucket, per-bucket counts / offsets (host)
    HashPose* hash_poses; unsigned* hash_by_bucket; unsigned hash_n; int hash_bins, hash_steps; uint32_t hash_step;
    std::vector<unsigned> hash_counts, hash_offs;
    CUtensorMap tmap; CUtensorMap* d_tmap; bool has_tmap;   // TMA descriptor of the grid (host copy, device copy)
    em::dbl2* htab;
    int max_smem_optin, sm_count;
    std::vector<cudaEvent_t> events;
    bool timing; std::vector<std::pair<int, cudaEvent_t> > marks; size_t marks_used;
    std::vector<float> step_ms;                   // step periods of the last eslam_gpu_run_steps
    // The fused step (project + update) as ONE CUDA graph with programmatic edges, for launch-bound sizes: the step's decisions
    // (resample?, which buffer, Philox step, exchange epochs) all live on the device, so a captured step is replayable; what
    // changes from call to call -- odometry step and contact points -- lives in the parameters of three kernel nodes, which
    // are updated in place (cudaGraphExecKernelNodeSetParams).
    struct StepGraph {
        cudaGraph_t graph; cudaGraphExec_t exec; cudaGraphNode_t n_project, n_scatter, n_lean;
        bool valid, failed; uint64_t n; int nb, lean_threads; unsigned kernels;
        eslam_gpu_step step; ContactSet cs; cudaEvent_t ev0, ev1;
    } sg;
    bool capturing;
    bool use_pdl;
};

static void graph_drop(eslam_gpu_ctx* ctx);

template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(eslam_gpu_ctx* ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args) {
    cudaLaunchConfig_t cfg; memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = ctx->stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = ctx->use_pdl ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

static void timing_mark(eslam_gpu_ctx* ctx, int line) {
    if (ctx->marks_used == ctx->marks.size()) { cudaEvent_t e; cudaEventCreate(&e); ctx->marks.push_back(std::make_pair(line, e)); }
    ctx->marks[ctx->marks_used].first = line;
    cudaEventRecord(ctx->marks[ctx->marks_used].second, ctx->stream);
    ctx->marks_used++;
}
static void timing_report(eslam_gpu_ctx* ctx) {
    if (!ctx->timing || ctx->marks_used < 2) return;
    cudaEventSynchronize(ctx->marks[ctx->marks_used - 1].second);
    std::vector<std::pair<int, std::pair<double, int> > > acc;
    for (size_t i = 1; i < ctx->marks_used; ++i) {
        float ms = 0; cudaEventElapsedTime(&ms, ctx->marks[i - 1].second, ctx->marks[i].second);
        size_t j = 0; for (; j < acc.size(); ++j) if (acc[j].first == ctx->marks[i].first) break;
        if (j == acc.size()) acc.push_back(std::make_pair(ctx->marks[i].first, std::make_pair(0.0, 0)));
        acc[j].second.first += ms; acc[j].second.second++;
    }
    double tot = 0;
    for (size_t j = 0; j < acc.size(); ++j) {
        fprintf(stderr, "[eslam timing] line %4d: %8.2f us x %d\n", acc[j].first, 1e3 * acc[j].second.first / acc[j].second.second, acc[j].second.second);
        tot += 1e3 * acc[j].second.first / acc[j].second.second;
    }
    fprintf(stderr, "[eslam timing] sum of means %.2f us\n", tot);
    ctx->marks_used = 0;
}

#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e__ = (call);                                                                  \
        if (e__ != cudaSuccess) {                                                                  \
            char buf__[256];                                                                       \
            snprintf(buf__, sizeof buf__, "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
            ctx->error = buf__;                                                                    \
            return ESLAM_ERR_CUDA;                                                                 \
        }                                                                                          \
    } while (0)

#define NK(call)                                                                                   \
    do {                                                                                           \
        int r__ = (call);                                                                          \
        if (r__ != 0) {                                                                            \
            char buf__[256];                                                                       \
            snprintf(buf__, sizeof buf__, "%s failed: %s", #call, g_nccl.GetErrorString ? g_nccl.GetErrorString(r__) : "?"); \
            ctx->error = buf__;                                                                    \
            return ESLAM_ERR_COMM;                                                                 \
        }                                                                                          \
    } while (0)

// ESLAM_GPU_TIMING=1: an event after every launch; eslam_gpu_run_steps prints the mean time between consecutive events
// (kernel + the gap before it) per launch site to stderr.  Development aid; off by default.
#define LAUNCHED() do { ctx->launches++; CK(cudaGetLastError()); if (ctx->timing) timing_mark(ctx, __LINE__); } while (0)

struct eslam_gpu_ctx;
static void timing_mark(eslam_gpu_ctx* ctx, int line);
// Launch with programmatic dependent launch allowed: the kernel's CTAs may be scheduled while the previous kernel of the
// stream is still draining; every kernel of the step starts with pdl_start() (griddepcontrol.launch_dependents + .wait), so
// nothing is read before the previous kernel has completed and flushed.  Hides the launch latency of the step's ~13 kernels.
template <typename... KArgs, typename... Args>
static cudaError_t launch_pdl(eslam_gpu_ctx* ctx, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, Args&&... args);
template <class T> static cudaError_t dalloc(T** p, size_t count) { return cudaMalloc((void**)p, count * sizeof(T)); }
template <class T> static void dfree(T*& p) { if (p) cudaFree(p); p = nullptr; }

static inline bool sharded(const eslam_gpu_ctx* ctx) { return ctx->shard.world > 1; }

static void free_particles(eslam_gpu_ctx* ctx) {
    for (int b = 0; b < 2; ++b) {
        for (int f = 0; f < 7; ++f) dfree(ctx->st.f[b][f]);
        for (int f = 0; f < 2; ++f) dfree(ctx->st.b[b][f]);
    }
    dfree(ctx->ancestors); dfree(ctx->cdf); dfree(ctx->keys); dfree(ctx->bin_list); dfree(ctx->el_d0); dfree(ctx->el_code);
    for (int b = 0; b < 2; ++b) { dfree(ctx->patch[b]); dfree(ctx->porigin[b]); }
    if (ctx->g_cent != ctx->l_cent) { dfree(ctx->g_cent); dfree(ctx->g_best_w); dfree(ctx->g_best_idx); }
    ctx->g_cent = ctx->g_best_w = nullptr; ctx->g_best_idx = nullptr;
    dfree(ctx->l_part_a); dfree(ctx->l_tile_sum); dfree(ctx->l_tile_sq); dfree(ctx->l_desc); dfree(ctx->l_pool);
    dfree(ctx->prefix_raw);
    dfree(ctx->l_cent); dfree(ctx->l_best_w); dfree(ctx->l_best_idx);
    dfree(ctx->tile_prefix); dfree(ctx->scratch_a); dfree(ctx->scratch_b); dfree(ctx->sb);
    dfree(ctx->tile_bofs); dfree(ctx->tile_pre);
    dfree(ctx->xchg); ctx->x_bytes = 0;
    ctx->cap = 0; ctx->t_cap = 0;
}

static int allgather(eslam_gpu_ctx* ctx, const void* send, void* recv, size_t bytes_per_rank);

// Exchange CUDA IPC handles of the running-sum, state, patch and exchange buffers and map every peer's buffers into this
// process.  NCCL only carries this bootstrap (and the per-tile centroid partials of getCentroid); the step's own exchanges are
// stores and flag polls on these mappings (eslam_kernels.cuh "peer exchange").
static int setup_peers(eslam_gpu_ctx* ctx) {
    const int G = ctx->shard.world, me = ctx->shard.rank;
    if (G > ESLAM_MAX_RANKS) { ctx->error = "more ranks than ESLAM_MAX_RANKS"; return ESLAM_ERR_INVALID; }
    if (!ctx->has_comm) { ctx->error = "sharded context without a communicator (call eslam_gpu_comm_init before init)"; return ESLAM_ERR_COMM; }
    if (ctx->env.no_p2p) { ctx->error = "sharded contexts need peer access between the GPUs (ESLAM_GPU_NO_P2P is set)"; return ESLAM_ERR_COMM; }
    const bool patches = ctx->cfg.patch_size > 0;
    const int NH = 20 + (patches ? 4 : 0);                              // cdf + 2 x 7 doubles + 2 x 2 bytes + exchange buffer (+ patches, origins)
    std::vector<void*> mine;
    mine.push_back(ctx->cdf);
    for (int b = 0; b < 2; ++b) for (int f = 0; f < 7; ++f) mine.push_back(ctx->st.f[b][f]);
    for (int b = 0; b < 2; ++b) for (int f = 0; f < 2; ++f) mine.push_back(ctx->st.b[b][f]);
    mine.push_back(ctx->xchg);
    if (patches) { for (int b = 0; b < 2; ++b) mine.push_back(ctx->patch[b]); for (int b = 0; b < 2; ++b) mine.push_back(ctx->porigin[b]); }
    std::vector<cudaIpcMemHandle_t> h_mine(NH), h_all((size_t)NH * G);
    for (int i = 0; i < NH; ++i) CK(cudaIpcGetMemHandle(&h_mine[i], mine[i]));
    cudaIpcMemHandle_t *d_send = nullptr, *d_recv = nullptr;
    CK(dalloc(&d_send, (size_t)NH)); CK(dalloc(&d_recv, (size_t)NH * G));
    CK(cudaMemcpyAsync(d_send, h_mine.data(), NH * sizeof(cudaIpcMemHandle_t), cudaMemcpyHostToDevice, ctx->stream));
    int st = allgather(ctx, d_send, d_recv, NH * sizeof(cudaIpcMemHandle_t));
    if (st == ESLAM_OK && cudaMemcpyAsync(h_all.data(), d_recv, (size_t)NH * G * sizeof(cudaIpcMemHandle_t), cudaMemcpyDeviceToHost, ctx->stream) != cudaSuccess) st = ESLAM_ERR_CUDA;
    if (st == ESLAM_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = ESLAM_ERR_CUDA;
    cudaFree(d_send); cudaFree(d_recv);
    if (st != ESLAM_OK) { if (ctx->error.empty()) ctx->error = "exchange of the IPC handles failed"; return st; }
    for (int q = 0; q < G; ++q) {
        std::vector<void*> p(NH);
        for (int i = 0; i < NH; ++i) {
            if (q == me) p[i] = mine[i];
            else { CK(cudaIpcOpenMemHandle(&p[i], h_all[(size_t)q * NH + i], cudaIpcMemLazyEnablePeerAccess)); ctx->ipc_opened.push_back(p[i]); }
        }
        int j = 0;
        ctx->peers.cdf[q] = (const double*)p[j++];
        for (int b = 0; b < 2; ++b) for (int f = 0; f < 7; ++f) ctx->peers.f[q][b][f] = (const double*)p[j++];
        for (int b = 0; b < 2; ++b) for (int f = 0; f < 2; ++f) ctx->peers.b[q][b][f] = (const uint8_t*)p[j++];
        ctx->X.buf[q] = (unsigned char*)p[j++];
        if (patches) { for (int b = 0; b < 2; ++b) ctx->peers.patch[q][b] = (const float2*)p[j++]; for (int b = 0; b < 2; ++b) ctx->peers.porigin[q][b] = (const int*)p[j++]; }
    }
    // nobody may publish into a buffer that its owner has not cleared yet: the handle exchange above is after the memset on
    // every rank's stream; one more collective orders "everyone has mapped everything" before the first step
    int* d_one = nullptr; int* d_all = nullptr;
    CK(dalloc(&d_one, 1)); CK(dalloc(&d_all, (size_t)G));
    st = allgather(ctx, d_one, d_all, sizeof(int));
    if (st == ESLAM_OK && cudaStreamSynchronize(ctx->stream) != cudaSuccess) st = ESLAM_ERR_CUDA;
    cudaFree(d_one); cudaFree(d_all);
    if (st != ESLAM_OK) return st;
    ctx->has_peers = true;
    return ESLAM_OK;
}

static int ensure_capacity(eslam_gpu_ctx* ctx, uint64_t n) {
    if (n == 0) return ESLAM_ERR_EMPTY;
    if (n > 0xFFFFFFFFull) { ctx->error = "more than 2^32 particles per shard"; return ESLAM_ERR_INVALID; }
    const int G = sharded(ctx) ? ctx->shard.world : 1;
    if (G > 1) {
        // equal shards of 1024 * 2^k particles: the pairwise summation tree over all tiles then factors into per-rank subtrees,
        // which is what lets every rank reduce its own tiles and exchange O(1) values (results identical to one GPU)
        const uint64_t t = n / TILE;
        if (n % TILE != 0 || (t & (t - 1)) != 0 || ctx->shard.global_count != n * (uint64_t)G) {
            ctx->error = "sharded contexts need equal shards of 1024 * 2^k particles"; return ESLAM_ERR_INVALID;
        }
    }
    if (n != ctx->n) graph_drop(ctx);                                   // the captured step holds the old sizes / buffers
    if (n > ctx->cap) {
        for (void* p : ctx->ipc_opened) cudaIpcCloseMemHandle(p);   // peer mappings of the old buffers are stale
        ctx->ipc_opened.clear(); ctx->has_peers = false;
        free_particles(ctx);
        for (int b = 0; b < 2; ++b) {
            for (int f = 0; f < 7; ++f) CK(dalloc(&ctx->st.f[b][f], n + 32));    // + 32: bulk copies of the last tile round their
            for (int f = 0; f < 2; ++f) CK(dalloc(&ctx->st.b[b][f], n + 32));    //   size up to 16 bytes (k_resample_stream)
        }
        CK(dalloc(&ctx->ancestors, n)); CK(dalloc(&ctx->cdf, n + 32)); CK(dalloc(&ctx->el_d0, n + 4)); CK(dalloc(&ctx->el_code, n + 4));
        CK(cudaMemsetAsync(ctx->ancestors, 0, n * sizeof(uint32_t), ctx->stream));
        CK(dalloc(&ctx->keys, n)); CK(dalloc(&ctx->bin_list, n));
        if (ctx->cfg.patch_size > 0) {
            const size_t pp = (size_t)ctx->cfg.patch_size * ctx->cfg.patch_size;
            for (int b = 0; b < 2; ++b) { CK(dalloc(&ctx->patch[b], n * pp)); CK(dalloc(&ctx->porigin[b], 2 * n)); }
        }
        const size_t t = (size_t)((n + TILE - 1) / TILE), tg = t * G;
        CK(dalloc(&ctx->l_part_a, t)); CK(dalloc(&ctx->l_tile_sum, t)); CK(dalloc(&ctx->l_tile_sq, t));
        CK(dalloc(&ctx->l_desc, t)); CK(dalloc(&ctx->l_pool, (size_t)MAX_BND_TOTAL)); CK(dalloc(&ctx->prefix_raw, t));
        CK(dalloc(&ctx->l_cent, 5 * t)); CK(dalloc(&ctx->l_best_w, t)); CK(dalloc(&ctx->l_best_idx, t));
        if (G > 1) { CK(dalloc(&ctx->g_cent, 5 * tg)); CK(dalloc(&ctx->g_best_w, tg)); CK(dalloc(&ctx->g_best_idx, tg)); }
        else { ctx->g_cent = ctx->l_cent; ctx->g_best_w = ctx->l_best_w; ctx->g_best_idx = ctx->l_best_idx; }
        CK(dalloc(&ctx->tile_prefix, tg)); CK(dalloc(&ctx->scratch_a, tg)); CK(dalloc(&ctx->scratch_b, tg));
        CK(dalloc(&ctx->sb, (size_t)MAX_BND_TOTAL));
        CK(dalloc(&ctx->tile_bofs, t)); CK(dalloc(&ctx->tile_pre, t));
        // exchange buffer: slots + flags (fixed offsets), then the tile-start table of all ranks
        ctx->x_bytes = (size_t)XOFF_TSIN + tg * sizeof(double);
        CK(dalloc(&ctx->xchg, ctx->x_bytes));
        CK(cudaMemsetAsync(ctx->xchg, 0, ctx->x_bytes, ctx->stream));
        // exchange epochs restart with the buffer (flags are zero again on every rank: they all reallocate together)
        CK(cudaMemsetAsync(ctx->sc->x_epoch, 0, sizeof(ctx->sc->x_epoch), ctx->stream));
        ctx->cap = n; ctx->t_cap = (unsigned)t;
    }
    ctx->n = n; ctx->t_local = (unsigned)((n + TILE - 1) / TILE); ctx->t_global = ctx->t_local * G;
    ctx->X.world = G; ctx->X.me = sharded(ctx) ? ctx->shard.rank : 0; ctx->X.tiles_per_rank = ctx->t_cap;
    if (G == 1) {                                                       // one GPU: the "peer" tables map this context's own buffers
        ctx->X.buf[0] = ctx->xchg;
        ctx->peers.cdf[0] = ctx->cdf;
        for (int b = 0; b < 2; ++b) {
            for (int f = 0; f < 7; ++f) ctx->peers.f[0][b][f] = ctx->st.f[b][f];
            for (int f = 0; f < 2; ++f) ctx->peers.b[0][b][f] = ctx->st.b[b][f];
            ctx->peers.patch[0][b] = ctx->patch[b]; ctx->peers.porigin[0][b] = ctx->porigin[b];
        }
        return ESLAM_OK;
    }
    if (ctx->t_local != ctx->t_cap) { ctx->error = "a sharded context cannot shrink its shard (the tile-start table is laid out by capacity)"; return ESLAM_ERR_INVALID; }
    if (!ctx->has_peers) { const int st = setup_peers(ctx); if (st != ESLAM_OK) return st; }
    return ESLAM_OK;
}

static inline uint64_t global_count(const eslam_gpu_ctx* ctx) { return sharded(ctx) ? ctx->shard.global_count : ctx->n; }
static inline uint64_t base_index(const eslam_gpu_ctx* ctx) { return sharded(ctx) ? (uint64_t)ctx->shard.rank * ctx->n : 0; }
static inline unsigned tile_ofs(const eslam_gpu_ctx* ctx) { return sharded(ctx) ? (unsigned)ctx->shard.rank * ctx->t_local : 0; }
static inline double* tile_sin_table(const eslam_gpu_ctx* ctx) { return reinterpret_cast<double*>(ctx->xchg + XOFF_TSIN); }
// LOCAL_MIGRATE sharding: the logical shards are the ranks of a sharded filter, or cfg.virtual_shards equal parts of an
// unsharded context (the same logic on one GPU; identical results).  Inert (mode 0) with one shard or in EXACT_GLOBAL mode.
static ShardPlan shard_plan(const eslam_gpu_ctx* ctx) {
    ShardPlan sp; memset(&sp, 0, sizeof(sp));
    if (ctx->cfg.shard_mode != ESLAM_SHARD_LOCAL_MIGRATE || ctx->cfg.resample_mode != ESLAM_RESAMPLE_STRATIFIED || ctx->cfg.patch_size > 0) return sp;
    unsigned G = 1, first = 0, count = 1;
    if (sharded(ctx)) { G = (unsigned)ctx->shard.world; first = (unsigned)ctx->shard.rank; }
    else if (ctx->cfg.virtual_shards > 1 && ctx->cfg.virtual_shards <= ESLAM_MAX_SHARDS && ctx->n % ((uint64_t)ctx->cfg.virtual_shards * TILE) == 0) {
        G = (unsigned)ctx->cfg.virtual_shards; count = G;
    }
    if (G <= 1) return sp;
    sp.mode = 1; sp.n_shards = G; sp.first = first; sp.count = count;
    sp.tiles_per_shard = ctx->t_global / G; sp.n_shard = (unsigned long long)sp.tiles_per_shard * TILE;
    sp.min_share = ctx->cfg.migrate_min_share; sp.min_ess = ctx->cfg.migrate_min_ess;
    sp.shard_tot = ctx->shard_tot;
    return sp;
}
static inline double scan_eps(const eslam_gpu_ctx* ctx) {
    // |sequential fp64 running sum - approximate parallel prefix| <= N 2^-53 (relative); 8x margin, floor 2^-40
    const double e = ldexp((double)global_count(ctx), -50);
    return e > ldexp(1.0, -40) ? e : ldexp(1.0, -40);
}

// all-gather of equal-sized per-rank blocks (device pointers, stream ordered); identity on one GPU.  Off the step path:
// bootstrap of the peer mappings and the per-tile partials of getCentroid / getBestParticleIndex.
static int allgather(eslam_gpu_ctx* ctx, const void* send, void* recv, size_t bytes_per_rank) {
    if (!sharded(ctx)) return ESLAM_OK;
    if (!ctx->has_comm) { ctx->error = "sharded context without a communicator (call eslam_gpu_comm_init)"; return ESLAM_ERR_COMM; }
    NK(g_nccl.AllGather(send, recv, bytes_per_rank, ESLAM_NCCL_UINT8, ctx->comm, ctx->stream));
    return ESLAM_OK;
}
#define AG(send, recv, bytes) do { int s__ = allgather(ctx, send, recv, bytes); if (s__ != ESLAM_OK) return s__; } while (0)

static ModelParams model_params(const eslam_gpu_ctx* ctx) {
    const eslam_gpu_config& c = ctx->cfg;
    ModelParams m;
    m.k0 = (uint32_t)c.seed; m.k1 = (uint32_t)(c.seed >> 32);
    m.slip_factor = c.slip_factor; m.max_yaw_dev = c.max_yaw_deviation; m.spread_thr = c.spread_threshold;
    m.spread_tf = c.spread_translation_factor; m.spread_rf = c.spread_rotation_factor;
    m.merr2 = c.measurement_error * c.measurement_error;
    m.inv_corr = 1.0 / c.contact_likelihood_correction;
    m.discount = c.discount_factor;
    m.min_contacts = c.min_contacts; m.use_shape = c.use_shape_update; m.use_hash = c.use_hash;
    return m;
}

// ContactModel::setContactPoints bookkeeping (src/ContactModel.cpp:21-41 + the group logic of :136,153-154,194-197)
static int build_contact_set(eslam_gpu_ctx* ctx, const eslam_gpu_contact* pts, int n, ContactSet* cs) {
    if (!pts || n < 0 || n > ESLAM_MAX_CONTACTS) { ctx->error = "contact point count out of range"; return ESLAM_ERR_INVALID; }
    memset(cs, 0, sizeof(*cs));
    cs->n = n;
    const double contact_threshold = 0.2;                                 // src/ContactModel.cpp:136
    for (int i = 0; i < n; ++i) {
        cs->px[i] = pts[i].position[0]; cs->py[i] = pts[i].position[1];
        cs->pzr[i] = pts[i].position[2] - ctx->cfg.contact_point_radius;  // :147
        const float contactProbability = pts[i].contact;                  // :153
        unsigned f = 0;
        if (!(contactProbability < contact_threshold)) f |= 1u;           // :154 (NaN passes)
        const int gid = pts[i].group_id;
        if (gid == -1 || i + 1 == n || gid != pts[i + 1].group_id) f |= 2u;   // :195-197
        cs->flags[i] = (uint8_t)f;
    }
    return ESLAM_OK;
}

static size_t weight_smem_bytes(const eslam_gpu_ctx* ctx) {
    return WT_HEAD_BYTES + (size_t)ctx->map.tile_w * ctx->map.tile_h * sizeof(float2);
}

// empty maps for every particle: NaN cells (0xFF bytes), un-anchored origins
static int reset_patches(eslam_gpu_ctx* ctx) {
    if (ctx->cfg.patch_size <= 0) return ESLAM_OK;
    const size_t pp = (size_t)ctx->cfg.patch_size * ctx->cfg.patch_size;
    for (int b = 0; b < 2; ++b) {
        CK(cudaMemsetAsync(ctx->patch[b], 0xFF, ctx->n * pp * sizeof(float2), ctx->stream));
        k_fill_i32<<<(unsigned)((2 * ctx->n + 255) / 256), 256, 0, ctx->stream>>>(ctx->porigin[b], 2 * ctx->n, INT_MIN);
        ctx->launches++;
        CK(cudaGetLastError());
    }
    return ESLAM_OK;
}

static int read_scalars(eslam_gpu_ctx* ctx) {
    CK(cudaMemcpyAsync(ctx->sc_host, ctx->sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ESLAM_OK;
}

extern "C" {

void eslam_gpu_default_config(eslam_gpu_config* c) {
    if (!c) return;
    memset(c, 0, sizeof(*c));
    c->seed = 42u; c->min_effective = 50;                                 // src/Configuration.hpp:86-88
    c->measurement_error = 0.1; c->discount_factor = 0.9; c->spread_threshold = 0.9;
    c->spread_translation_factor = 0.1; c->spread_rotation_factor = 0.05; c->slip_factor = 0.05;
    c->max_yaw_deviation = 15 * M_PI / 180.0;
    c->min_contacts = 3; c->contact_likelihood_correction = 0.33; c->contact_point_radius = 0.01;   // :55-59
    c->use_slip_update = 0; c->use_shape_update = 1; c->use_hash = 0;
    c->resample_mode = ESLAM_RESAMPLE_STRATIFIED; c->patch_size = 0;
    c->shard_mode = ESLAM_SHARD_EXACT_GLOBAL; c->migrate_min_share = 0.25; c->migrate_min_ess = 0.05; c->virtual_shards = 0;
}

static int create_impl(eslam_gpu_ctx* ctx, int device) {
    CK(cudaSetDevice(device));
    CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CK(cudaDeviceGetAttribute(&ctx->max_smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, device));
    CK(cudaDeviceGetAttribute(&ctx->sm_count, cudaDevAttrMultiProcessorCount, device));
    CK(dalloc(&ctx->sc, 1));
    CK(cudaMemsetAsync(ctx->sc, 0, sizeof(DevScalars), ctx->stream));
    CK(cudaMallocHost((void**)&ctx->sc_host, sizeof(DevScalars)));
    CK(cudaMallocHost((void**)&ctx->sc_async, sizeof(DevScalars)));
    memset(ctx->sc_async, 0, sizeof(DevScalars));
    CK(cudaHostAlloc((void**)&ctx->box_host, sizeof(ResultBox), cudaHostAllocMapped));
    memset(ctx->box_host, 0, sizeof(ResultBox));
    CK(cudaHostGetDevicePointer((void**)&ctx->box_dev, ctx->box_host, 0));
    CK(cudaMalloc((void**)&ctx->box_seq, sizeof(unsigned long long)));
    CK(cudaMemsetAsync(ctx->box_seq, 0, sizeof(unsigned long long), ctx->stream));
    CK(cudaStreamCreateWithFlags(&ctx->side, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->side_ev, cudaEventDisableTiming));
    CK(dalloc(&ctx->ticket, 1)); CK(cudaMemsetAsync(ctx->ticket, 0, sizeof(unsigned), ctx->stream));
    CK(dalloc(&ctx->shard_tot, (size_t)ESLAM_MAX_SHARDS)); CK(cudaMemsetAsync(ctx->shard_tot, 0, ESLAM_MAX_SHARDS * sizeof(double), ctx->stream));
    CK(dalloc(&ctx->shard_sums, (size_t)ESLAM_MAX_SHARDS)); CK(cudaMemsetAsync(ctx->shard_sums, 0, ESLAM_MAX_SHARDS * sizeof(double2), ctx->stream));
    CK(dalloc(&ctx->bin_count, (size_t)MAX_BINS)); CK(dalloc(&ctx->bin_offs, (size_t)MAX_BINS)); CK(dalloc(&ctx->bin_next, (size_t)MAX_BINS));
    CK(dalloc(&ctx->cta_bin, (size_t)ctx->sm_count));
    CK(cudaMemsetAsync(ctx->bin_count, 0, MAX_BINS * sizeof(int), ctx->stream));
    CK(cudaMemsetAsync(ctx->bin_next, 0, MAX_BINS * sizeof(unsigned), ctx->stream));
    CK(dalloc(&ctx->hint, (size_t)2 * ctx->sm_count));
    CK(cudaMemsetAsync(ctx->hint, 0, 2 * ctx->sm_count * sizeof(double), ctx->stream));
    CK(dalloc(&ctx->htab, (size_t)ESLAM_H_TABLE_ENTRIES));
    CK(cudaMemcpyAsync(ctx->htab, H_TABLE_HOST, sizeof(H_TABLE_HOST), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaFuncSetAttribute(k_chain, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ChainSmem)));
    CK(cudaFuncSetAttribute(k_resample_stream, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)RS_STREAM_SMEM));
    CK(cudaStreamSynchronize(ctx->stream));
    return ESLAM_OK;
}

int eslam_gpu_create(const eslam_gpu_config* cfg, int device, const eslam_gpu_shard* shard, eslam_gpu_ctx** out) {
    if (!cfg || !out) return ESLAM_ERR_INVALID;
    *out = nullptr;
    int count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return ESLAM_ERR_CUDA;
    eslam_gpu_ctx* ctx = new eslam_gpu_ctx();                          // value-initialised: every pointer null, every count zero
    ctx->cfg = *cfg;
    ctx->shard.rank = 0; ctx->shard.world = 1; ctx->shard.global_count = 0;
    if (shard) ctx->shard = *shard;
    ctx->device = device;
    ctx->env.read();
    ctx->timing = ctx->env.timing; ctx->use_pdl = ctx->env.use_pdl;
    ctx->fast = new FastParams();
    const int st = create_impl(ctx, device);
    if (st != ESLAM_OK) { eslam_gpu_destroy(ctx); return st; }         // no half-built context leaves this function
    *out = ctx;
    return ESLAM_OK;
}

void eslam_gpu_destroy(eslam_gpu_ctx* ctx) {
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    if (ctx->stream) cudaStreamSynchronize(ctx->stream);
    if (ctx->has_comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm);
    for (void* p : ctx->ipc_opened) cudaIpcCloseMemHandle(p);
    free_particles(ctx);
    if (ctx->side) { cudaStreamSynchronize(ctx->side); cudaStreamDestroy(ctx->side); }
    if (ctx->side_ev) cudaEventDestroy(ctx->side_ev);
    if (ctx->sc_async) cudaFreeHost(ctx->sc_async);
    if (ctx->box_host) cudaFreeHost(ctx->box_host);
    if (ctx->box_seq) cudaFree(ctx->box_seq);
    if (ctx->sc_host) cudaFreeHost(ctx->sc_host);
    cudaFree(ctx->sc); cudaFree(ctx->d_tmap); cudaFree(ctx->htab); cudaFree(ctx->cells); cudaFree(ctx->hint); cudaFree(ctx->ticket);
    cudaFree(ctx->bin_count); cudaFree(ctx->bin_offs); cudaFree(ctx->bin_next); cudaFree(ctx->cta_bin);
    cudaFree(ctx->shard_tot); cudaFree(ctx->shard_sums); cudaFree(ctx->scan_cells); cudaFree(ctx->hash_poses); cudaFree(ctx->hash_by_bucket);
    graph_drop(ctx);
    if (ctx->sg.ev0) { cudaEventDestroy(ctx->sg.ev0); cudaEventDestroy(ctx->sg.ev1); }
    for (cudaEvent_t e : ctx->events) cudaEventDestroy(e);
    for (auto& m : ctx->marks) cudaEventDestroy(m.second);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    delete ctx->fast;
    delete ctx;
}

const char* eslam_gpu_last_error(const eslam_gpu_ctx* ctx) { return ctx ? ctx->error.c_str() : "null context"; }
uint64_t eslam_gpu_launch_count(const eslam_gpu_ctx* ctx) { return ctx ? ctx->launches : 0; }
uint64_t eslam_gpu_particle_count(const eslam_gpu_ctx* ctx) { return ctx ? ctx->n : 0; }

int eslam_gpu_nccl_unique_id(uint8_t out[128]) {
    if (!out || !g_nccl.load()) return ESLAM_ERR_COMM;
    eslam_ncclUniqueId id;
    if (g_nccl.GetUniqueId(&id) != 0) return ESLAM_ERR_COMM;
    memcpy(out, id.internal, 128);
    return ESLAM_OK;
}

int eslam_gpu_comm_init(eslam_gpu_ctx* ctx, const uint8_t id_bytes[128]) {
    if (!ctx || !id_bytes) return ESLAM_ERR_INVALID;
    if (!sharded(ctx)) return ESLAM_OK;
    if (!g_nccl.load()) { ctx->error = "libnccl.so.2 not found"; return ESLAM_ERR_COMM; }
    CK(cudaSetDevice(ctx->device));
    eslam_ncclUniqueId id; memcpy(id.internal, id_bytes, 128);
    NK(g_nccl.CommInitRank(&ctx->comm, ctx->shard.world, id, ctx->shard.rank));
    ctx->has_comm = true;
    return ESLAM_OK;
}

int eslam_gpu_synchronize(eslam_gpu_ctx* ctx) {
    if (!ctx) return ESLAM_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    return ESLAM_OK;
}

int eslam_gpu_set_grid(eslam_gpu_ctx* ctx, const float* cells, int width, int height, double sx, double sy, double ox,
                       double oy, const double* g2l) {
    if (!ctx || !cells || width <= 0 || height <= 0 || !(sx > 0) || !(sy > 0)) return ESLAM_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    CK(cudaStreamSynchronize(ctx->stream));
    cudaFree(ctx->cells); ctx->cells = nullptr;
    graph_drop(ctx);                                                      // the captured step holds the old grid's pointers
    const size_t count = (size_t)width * height;
    CK(dalloc(&ctx->cells, count));
    CK(cudaMemcpyAsync(ctx->cells, cells, count * sizeof(float2), cudaMemcpyHostToDevice, ctx->stream));
    // lean-path precondition: stdev^2 + measVar stays inside rsqrt_core's validated range (see k_weight_fast)
    ctx->grid_sane = true;
    for (size_t i = 0; i < count; ++i) {
        const float sd = cells[2 * i + 1];
        if (sd == sd && !(fabsf(sd) < 5.6e14f)) { ctx->grid_sane = false; break; }   // 2^49 = 5.63e14
    }
    MapParams& m = ctx->map;
    m.cells = ctx->cells; m.W = width; m.H = height; m.sx = sx; m.sy = sy; m.ox = ox; m.oy = oy;
    m.inv_sx = 1.0 / sx; m.inv_sy = 1.0 / sy;
    m.c0x = 25769803776.0 - ox * m.inv_sx; m.c0y = 25769803776.0 - oy * m.inv_sy;   // 1.5 * 2^34 - offset/scale
    m.identity = g2l == nullptr;
    for (int i = 0; i < 8; ++i) m.g2l[i] = g2l ? g2l[i] : (i == 0 || i == 5 ? 1.0 : 0.0);
    // shared-memory map tile: as many cells as fit next to the h-table
    const size_t avail = (size_t)ctx->max_smem_optin - WT_HEAD_BYTES;
    int side = (int)floor(sqrt((double)(avail / sizeof(float2))));
    if (ctx->env.tile_override > 0) side = ctx->env.tile_override;
    m.tile_w = width < side ? width : side;
    // a grid at least 256 cells wide: tiles of 256 columns (the row pitch the lean kernel turns into a byte permute, and the TMA
    // box limit), as many rows as fit (about 100: below the 128 the permute needs)
    if (width >= 256 && ctx->env.tile_override <= 0) m.tile_w = 256;
    m.tile_h = (int)((avail / sizeof(float2)) / (size_t)m.tile_w);
    if (m.tile_w == 256 && m.tile_h > 127) m.tile_h = 127;
    if (ctx->env.tile_override > 0) m.tile_h = ctx->env.tile_override;
    if (m.tile_h > height) m.tile_h = height;
    if (m.tile_w > 256) m.tile_w = 256;                           // TMA box limit per dimension
    if (m.tile_h > 256) m.tile_h = 256;
    // TMA descriptor: 2-D tensor of 8-byte cells, box = the tile.  Needs 16-byte multiples for the row pitch and the
    // box's inner extent; otherwise the kernel stages the tile with ordinary loads.
    ctx->has_tmap = false;
    if ((width % 2) == 0 && (m.tile_w % 2) == 0) {
        typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                     const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                     CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
        void* fn = nullptr; cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess && fn &&
            qres == cudaDriverEntryPointSuccess) {
            const cuuint64_t gdim[2] = {(cuuint64_t)width, (cuuint64_t)height};
            const cuuint64_t gstride[1] = {(cuuint64_t)width * sizeof(float2)};
            const cuuint32_t box[2] = {(cuuint32_t)m.tile_w, (cuuint32_t)m.tile_h};
            const cuuint32_t estr[2] = {1, 1};
            const CUresult r = ((EncodeFn)fn)(&ctx->tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, (void*)ctx->cells, gdim, gstride, box, estr,
                                              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
                                              CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
            ctx->has_tmap = (r == CUDA_SUCCESS);
            if (ctx->env.trace) {
                const unsigned long long* q = reinterpret_cast<const unsigned long long*>(&ctx->tmap);
                fprintf(stderr, "[eslam] tensor map encode=%d cells=%p W=%d H=%d box=%dx%d:", (int)r, (void*)ctx->cells, width, height, m.tile_w, m.tile_h);
                for (int i = 0; i < 16; ++i) fprintf(stderr, " %016llx", q[i]);
                fprintf(stderr, "\n");
            }
            if (ctx->has_tmap) {
                if (!ctx->d_tmap) CK(cudaMalloc((void**)&ctx->d_tmap, sizeof(CUtensorMap)));
                CK(cudaMemcpyAsync(ctx->d_tmap, &ctx->tmap, sizeof(CUtensorMap), cudaMemcpyHostToDevice, ctx->stream));
            }
        }
    }
    const int bytes = (int)weight_smem_bytes(ctx);
    CK(cudaFuncSetAttribute(k_project_weight<true, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CK(cudaFuncSetAttribute(k_project_weight<false, true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CK(cudaFuncSetAttribute(k_project_weight<true, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CK(cudaFuncSetAttribute(k_project_weight<false, true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CK(cudaFuncSetAttribute(k_weight_lean<1024, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CK(cudaFuncSetAttribute(k_weight_lean<1024, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CK(cudaFuncSetAttribute(k_weight_lean<512, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->has_map = true;
    return ESLAM_OK;
}

int eslam_gpu_init_gaussian(eslam_gpu_ctx* ctx, uint64_t n, const double mu[3], const double sigma[3], double zpos,
                            double zsigma) {
    if (!ctx || !mu || !sigma) return ESLAM_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    int st = ensure_capacity(ctx, n);
    if (st != ESLAM_OK) return st;
    DevScalars init; memset(&init, 0, sizeof(init));
    init.tile_cx = mu[0]; init.tile_cy = mu[1];
    CK(cudaStreamSynchronize(ctx->side));
    *ctx->sc_async = init;
    *ctx->sc_host = init;
    CK(cudaMemcpyAsync(ctx->sc, ctx->sc_host, sizeof(DevScalars), cudaMemcpyHostToDevice, ctx->stream));
    // Re-initialising a sharded context that has stepped: the exchange epochs restart with the scalars, so the words of the old
    // epochs must go.  The caller's part of the contract (include/eslam_gpu.h): every rank has finished its steps before any rank
    // re-initialises, and no rank steps before all have -- a barrier on either side.
    if (sharded(ctx) && ctx->xchg && ctx->chain_launched > 0) CK(cudaMemsetAsync(ctx->xchg, 0, ctx->x_bytes, ctx->stream));
    st = reset_patches(ctx);
    if (st != ESLAM_OK) return st;
    const ModelParams mp = model_params(ctx);
    k_init_gaussian<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->st, ctx->sc, n, base_index(ctx), mp.k0, mp.k1,
                                                                         mu[0], mu[1], mu[2], sigma[0], sigma[1], sigma[2],
                                                                         zpos, zsigma);
    LAUNCHED();
    CK(cudaStreamSynchronize(ctx->stream));
    return ESLAM_OK;
}

int eslam_gpu_set_particles(eslam_gpu_ctx* ctx, uint64_t n, const double* x, const double* y, const double* theta,
                            const double* zpos, const double* zsigma, const double* weight) {
    if (!ctx) return ESLAM_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    const bool fresh = n != ctx->n;
    int st = ensure_capacity(ctx, n);
    if (st != ESLAM_OK) return st;
    st = read_scalars(ctx);
    if (st != ESLAM_OK) return st;
    const int cur = ctx->sc_host->cur;
    if (fresh) {
        for (int f = 0; f < 7; ++f) CK(cudaMemsetAsync(ctx->st.f[cur][f], 0, n * sizeof(double), ctx->stream));
        CK(cudaMemsetAsync(ctx->st.b[cur][B_FLOAT], 1, n, ctx->stream));
        CK(cudaMemsetAsync(ctx->st.b[cur][B_NF], 0, n, ctx->stream));
        st = reset_patches(ctx);
        if (st != ESLAM_OK) return st;
    }
    const double* src[6] = {x, y, theta, zpos, zsigma, weight};
    for (int f = 0; f < 6; ++f)
        if (src[f]) CK(cudaMemcpyAsync(ctx->st.f[cur][f], src[f], n * sizeof(double), cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ESLAM_OK;
}

int eslam_gpu_get_particles(eslam_gpu_ctx* ctx, double* x, double* y, double* theta, double* zpos, double* zsigma,
                            double* weight, double* mprob, uint8_t* floating, uint8_t* n_found) {
    if (!ctx) return ESLAM_ERR_INVALID;
    if (ctx->n == 0) return ESLAM_ERR_EMPTY;
    CK(cudaSetDevice(ctx->device));
    int st = read_scalars(ctx);
    if (st != ESLAM_OK) return st;
    const int cur = ctx->sc_host->cur;
    double* dst[7] = {x, y, theta, zpos, zsigma, weight, mprob};
    for (int f = 0; f < 7; ++f)
        if (dst[f]) CK(cudaMemcpyAsync(dst[f], ctx->st.f[cur][f], ctx->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (floating) CK(cudaMemcpyAsync(floating, ctx->st.b[cur][B_FLOAT], ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
    if (n_found) CK(cudaMemcpyAsync(n_found, ctx->st.b[cur][B_NF], ctx->n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ESLAM_OK;
}

}  // extern "C"

// ---------------------------------------------------------------------------------------------- launch sequences
// where to put the shared-memory map tile: around the cloud centre the device reported last (a hint: it may be a few
// steps old when steps are enqueued without synchronisation; results do not depend on it).  Inside the grid, x0 even
// (TMA needs a 16-byte aligned box start for 8-byte cells).
static void place_tile(const eslam_gpu_ctx* ctx, MapParams* m) {
    volatile DevScalars* as = ctx->sc_async;                          // as of the last update that has reached the host
    double cx = (as->tile_cx - ctx->map.ox) * ctx->map.inv_sx, cy = (as->tile_cy - ctx->map.oy) * ctx->map.inv_sy;
    if (!(cx > -1e9 && cx < 1e9)) cx = 0;
    if (!(cy > -1e9 && cy < 1e9)) cy = 0;
    const int x0 = (int)cx - m->tile_w / 2, y0 = (int)cy - m->tile_h / 2;
    m->tile_x0 = std::max(0, std::min(x0, m->W - m->tile_w)) & ~1;
    m->tile_y0 = std::max(0, std::min(y0, m->H - m->tile_h));
}

static void fill_weight_params(eslam_gpu_ctx* ctx, WeightParams* P, const eslam_gpu_step* step, const ContactSet* cs) {
    memset(P, 0, sizeof(*P));
    P->st = ctx->st; P->sc = ctx->sc; P->map = ctx->map; P->mp = model_params(ctx);
    P->htab = ctx->htab; P->n = ctx->n; P->base = base_index(ctx);
    P->patch_size = ctx->cfg.patch_size > 0 ? ctx->cfg.patch_size : 0;
    P->use_tma = ctx->has_tmap && !ctx->env.no_tma;
    P->tmap = ctx->d_tmap;
    P->hint = ctx->hint;
    for (int b = 0; b < 2; ++b) { P->patch[b] = ctx->patch[b]; P->porigin[b] = ctx->porigin[b]; }
    if (step) {
        P->yaw = step->yaw; P->z_delta = step->z_delta; P->z_var = step->z_var;
        for (int i = 0; i < 3; ++i) { P->dmean[i] = step->delta_mean[i]; P->dsigma[i] = step->delta_sigma[i]; }
    }
    if (cs) P->cs = *cs;
    place_tile(ctx, &P->map);
}

static inline uint32_t hi_word(double v) { uint64_t u; memcpy(&u, &v, 8); return (uint32_t)(u >> 32); }

// The lean kernel's inputs (k_weight_fast).  Returns false when the lean path does not apply to this context / update:
// per-particle patches, a non-identity C_global2local, no shared-memory tile, a grid with out-of-range deviations.
static bool fill_fast_params(eslam_gpu_ctx* ctx, FastParams* F, const WeightParams& P, const ContactSet& cs) {
    if (ctx->env.no_fast) return false;
    if (ctx->cfg.patch_size > 0 || !ctx->map.identity || !ctx->grid_sane || P.map.tile_w <= 0 || P.map.tile_h <= 0) return false;
    if (ctx->n >= 0x80000000ull) return false;                            // bit 31 of a list entry flags the particle
    memset(F, 0, sizeof(*F));
    WeightCommon& c = F->c;
    c.st = P.st; c.sc = P.sc; c.map = P.map; c.mp = P.mp; c.htab = P.htab; c.n = P.n; c.base = P.base;
    c.use_tma = P.use_tma; c.tmap = P.tmap; c.hint = P.hint;
    c.yaw = P.yaw; c.z_delta = P.z_delta; c.z_var = P.z_var;
    for (int i = 0; i < 3; ++i) { c.dmean[i] = P.dmean[i]; c.dsigma[i] = P.dsigma[i]; }
    FastContacts& fc = F->fc;
    double r2 = 0; int last_end = 0;
    for (int i = 0; i < cs.n; ++i) {
        if (cs.flags[i] & 1u) {
            fc.px[fc.n_eval] = cs.px[i]; fc.py[fc.n_eval] = cs.py[i]; fc.pzr[fc.n_eval] = cs.pzr[i]; fc.n_eval++;
            const double q = cs.px[i] * cs.px[i] + cs.py[i] * cs.py[i];
            if (!(q < 1e12)) return false;                                // NaN / absurd contact positions: generic path
            if (q > r2) r2 = q;
        }
        if ((cs.flags[i] & 2u) && fc.n_eval > last_end) { fc.gend[fc.n_groups++] = (uint8_t)fc.n_eval; last_end = fc.n_eval; }
    }
    fc.n = fc.n_eval;
    for (int i = 0; i < fc.n_eval; ++i) fc.flags[i] = 1u;
    for (int g = 0; g < fc.n_groups; ++g) fc.flags[fc.gend[g] - 1] |= 2u;   // the compact list in the ContactSet convention
    // bins: tile cores on a lattice; a particle whose own cell is in a core has its whole footprint (contact offsets
    // <= R, +1 cell for the approximate own cell, +1 for rounding) inside that bin's tile.  Even margin and core width
    // keep the tile origins even (TMA: 16-byte aligned box start).
    const double inv_max = std::max(P.map.inv_sx, P.map.inv_sy);
    const double mc = ceil(sqrt(r2) * inv_max) + 2.0;
    if (!(mc < 100.0)) return false;
    fc.margin = ((int)mc + 1) & ~1;
    if ((P.map.W & 1) || (P.map.tile_w & 1)) return false;
    fc.bsx = (P.map.tile_w - 2 * fc.margin) & ~1; fc.bsy = P.map.tile_h - 2 * fc.margin;
    if (fc.bsx < 8 || fc.bsy < 8) return false;
    fc.nbx = (P.map.W + fc.bsx - 1) / fc.bsx; fc.nby = (P.map.H + fc.bsy - 1) / fc.bsy;
    fc.nb = fc.nbx * fc.nby;
    auto magic = [](uint64_t two_k, int d) { return (uint32_t)((two_k + (uint64_t)d - 1) / (uint64_t)d); };
    fc.inv_bsx = magic(1ull << 32, fc.bsx); fc.inv_bsy = magic(1ull << 32, fc.bsy); fc.inv_nbx = magic(1ull << 31, fc.nbx);
    if (fc.nb > MAX_BINS || P.map.W > 16000 || P.map.H > 16000) return false;
    // 3-sigma gate d^2 < 9 zvar  <=>  |t| < 3/corr with t = d / (sigma corr); decided on the high word of |t|
    const double T = 3.0 * P.mp.inv_corr;
    if (!(T > 0.0) || !(T < 1e300)) return false;
    const uint32_t hiT = hi_word(T), hi10 = hi_word(ESLAM_H_TMAX);
    fc.gate_lo = std::min(hiT, hi10) - 1u;                                // also keeps t inside the table
    fc.gate_span = (hiT + 1u) - fc.gate_lo;
    fc.c0x = 402653184.0 - P.map.ox * P.map.inv_sx;                       // 1.5 * 2^28 - offset/scale
    fc.c0y = 402653184.0 - P.map.oy * P.map.inv_sy;
    return true;
}

static unsigned weight_grid(const eslam_gpu_ctx* ctx) {
    return ctx->t_local < (unsigned)ctx->sm_count ? ctx->t_local : (unsigned)ctx->sm_count;
}

// normalise + (optional) resample tail shared by update / normalize_weights / resample.
// mode: 0 = do_resample := force, 1 = decide from ESS (update), 2 = ESS only
#ifndef RS_CTA
#define RS_CTA 1024               // strata (threads) per CTA of k_resample_local (512 measured the same)
#endif
static int launch_tail(eslam_gpu_ctx* ctx, bool apply_loop_b, int normalise, int mode, int force, bool cdf_only = false) {
    const unsigned T = ctx->t_local, TG = ctx->t_global;
    const uint64_t n = ctx->n, ng = global_count(ctx);
    const ShardPlan sp = shard_plan(ctx);
    LoopBArgs lb; lb.tile_sum = ctx->l_tile_sum; lb.n_tiles = T; lb.discount = ctx->cfg.discount_factor;
    lb.scratch_a = ctx->scratch_a; lb.scratch_b = ctx->scratch_b; lb.prefix_raw = ctx->prefix_raw; lb.shard_sums = ctx->shard_sums;
    if (apply_loop_b) CK(launch_pdl(ctx, k_loop_b<true>, T, SCAN_THREADS, 0, ctx->st, ctx->sc, n, lb, sp, ctx->X));
    else CK(launch_pdl(ctx, k_loop_b<false>, T, SCAN_THREADS, 0, ctx->st, ctx->sc, n, lb, sp, ctx->X));
    LAUNCHED();
    NormScanArgs ns; ns.n_global = ng; ns.prefix_raw = ctx->prefix_raw; ns.shard_sums = ctx->shard_sums; ns.tile_prefix = ctx->tile_prefix;
    ns.tile_sq = ctx->l_tile_sq; ns.desc = ctx->l_desc; ns.pool = ctx->l_pool; ns.scratch_a = ctx->scratch_a; ns.scratch_b = ctx->scratch_b;
    ns.n_tiles = T; ns.tile_ofs = tile_ofs(ctx); ns.normalise = normalise; ns.eps = scan_eps(ctx); ns.el_d0 = ctx->el_d0; ns.el_code = ctx->el_code;
    CK(launch_pdl(ctx, k_norm_scan, T, SCAN_THREADS, 0, ctx->st, ctx->sc, n, ns, sp, ctx->X));
    LAUNCHED();
    ChainArgs ca; ca.tile_sq = ctx->l_tile_sq; ca.desc = ctx->l_desc; ca.pool = ctx->l_pool; ca.n_tiles = T; ca.tile_ofs = tile_ofs(ctx);
    ca.tile_bofs = ctx->tile_bofs; ca.tile_pre = ctx->tile_pre; ca.sb_out = ctx->sb;
    ca.min_effective = mode == 2 ? 0ull : ctx->cfg.min_effective; ca.decide = mode == 0 ? 0 : 1; ca.force = force;
    ca.n = n; ca.cdf = ctx->cdf; ca.box = ctx->box_dev; ca.box_seq = ctx->box_seq;
    if (!ctx->capturing) ctx->chain_launched++;
    CK(launch_pdl(ctx, k_chain, 1, 1024, sizeof(ChainSmem), ctx->st, ctx->sc, ca, sp, ctx->X));
    LAUNCHED();
    if (mode == 2) return ESLAM_OK;
    CdfArgs cd; cd.tile_bofs = ctx->tile_bofs; cd.sb = ctx->sb; cd.cdf = ctx->cdf; cd.el_d0 = ctx->el_d0; cd.el_code = ctx->el_code; cd.tile_ofs = tile_ofs(ctx);
    CK(launch_pdl(ctx, k_cdf, T, SCAN_THREADS, 0, ctx->sc, n, cd, ctx->X));
    LAUNCHED();
    if (cdf_only) return ESLAM_OK;                                        // the caller walks the running sum itself
    const ModelParams mp = model_params(ctx);
    const int systematic = ctx->cfg.resample_mode == ESLAM_RESAMPLE_SYSTEMATIC;
    double* tile_sin = tile_sin_table(ctx);
    PatchGather pg; memset(&pg, 0, sizeof(pg));
    pg.P = ctx->cfg.patch_size > 0 ? ctx->cfg.patch_size : 0;
    for (int b = 0; b < 2; ++b) { pg.patch[b] = ctx->patch[b]; pg.porigin[b] = ctx->porigin[b]; }
    if (ctx->env.stream_resample && !sp.mode) {
        // persistent CTAs, source tiles streamed through a bulk-copy ring from the rank that owns them: own memory or a peer's
        // over NVLink (see k_resample_stream); sharded: tile starts from the global table the ranks' k_chain pushed
        StreamArgs sa; sa.tile_start = tile_sin; sa.n_tiles = TG; sa.tiles_per_rank = T; sa.n = n; sa.n_global = ng; sa.base = base_index(ctx);
        sa.k0key = mp.k0; sa.k1key = mp.k1; sa.systematic = systematic; sa.ancestors = ctx->ancestors; sa.ticket = ctx->ticket;
        sa.my_rank = sharded(ctx) ? (unsigned)ctx->shard.rank : 0u;
        CK(launch_pdl(ctx, k_resample_stream, (unsigned)ctx->sm_count, 1024, (size_t)RS_STREAM_SMEM, ctx->st, ctx->sc, sa, ctx->peers, ctx->X));
        LAUNCHED();
    } else if (sharded(ctx)) {
        // LOCAL_MIGRATE: every rank searches the global tile-start table (pushed by the ranks' k_chain), then the owner's running
        // sums, and copies its ancestors' state straight from the owner's buffers over NVLink
        CK(launch_pdl(ctx, k_resample_local<RS_CTA>, (unsigned)((n + RS_CTA - 1) / RS_CTA), RS_CTA, 0, ctx->st, ctx->sc, tile_sin, TG, T, n, ng, base_index(ctx),
                      mp.k0, mp.k1, systematic, ctx->ancestors, ctx->ticket, ctx->peers, sp, tile_ofs(ctx), (unsigned)ctx->shard.rank, ctx->X));
        LAUNCHED();
    } else {
        // search in shared memory, copy and buffer swap in one kernel (virtual shards, ESLAM_GPU_STREAM_RESAMPLE=0)
        CK(launch_pdl(ctx, k_resample_local<RS_CTA>, (unsigned)((n + RS_CTA - 1) / RS_CTA), RS_CTA, 0, ctx->st, ctx->sc, tile_sin, T, T, n, n, 0ull, mp.k0, mp.k1,
                      systematic, ctx->ancestors, ctx->ticket, ctx->peers, sp, 0u, 0u, ctx->X));
        LAUNCHED();
    }
    if (pg.P > 0) {
        // cloneMaps: the maps follow their particles (the state buffers are swapped already)
        CK(launch_pdl(ctx, k_gather_patches, (unsigned)((n * 32 + 255) / 256), 256, 0, ctx->sc, n, n, ctx->ancestors, pg, ctx->peers));
        LAUNCHED();
    }
    return ESLAM_OK;
}

static int launch_update(eslam_gpu_ctx* ctx, const eslam_gpu_step* step_or_null, const ContactSet* cs,
                         cudaEvent_t ev0, cudaEvent_t ev1) {
    WeightParams P;
    fill_weight_params(ctx, &P, step_or_null, cs);
    const size_t smem = weight_smem_bytes(ctx);
    const unsigned grid = weight_grid(ctx);
    const bool patches = ctx->cfg.patch_size > 0;
    FastParams& F = *ctx->fast;                                           // large (lists of 128 contacts): per context, not on the stack
    const bool lean = fill_fast_params(ctx, &F, P, *cs);
    if (!lean && ev0) CK(cudaEventRecord(ev0, ctx->stream));
    if (lean) {
        // bin the particles by map tile, evaluate bin by bin from shared memory (particles whose footprint leaves the map or
        // whose values are out of range are flagged and go through contact_eval inside the same kernel)
        const unsigned T = ctx->t_local;
        if (step_or_null) CK(launch_pdl(ctx, k_project_bin<true>, T, 1024, 0, F, ctx->bin_count, ctx->keys));
        else CK(launch_pdl(ctx, k_project_bin<false>, T, 1024, 0, F, ctx->bin_count, ctx->keys));
        LAUNCHED();
        CK(launch_pdl(ctx, k_bin_scatter, T, 1024, 0, ctx->bin_count, F.fc.nb, ctx->keys, ctx->n, ctx->bin_list, ctx->bin_offs, ctx->cta_bin, ctx->sm_count, ctx->sc,
                      step_or_null ? 1 : 0));
        LAUNCHED();
        if (ev0) CK(cudaEventRecordWithFlags(ev0, ctx->stream, ctx->capturing ? cudaEventRecordExternal : cudaEventRecordDefault));   // the events bracket the dominant kernel alone
        const bool p256 = F.c.map.tile_w == 256 && F.c.map.tile_h < 128;   // rows 256 cells apart: cell offsets by byte permute
        if (ctx->env.lean_threads == 512)
            CK(launch_pdl(ctx, (k_weight_lean<512, false>), (unsigned)ctx->sm_count, 512, smem, F, ctx->bin_list, ctx->bin_count, ctx->bin_offs, ctx->bin_next, ctx->cta_bin));
        else if (p256)
            CK(launch_pdl(ctx, (k_weight_lean<1024, true>), (unsigned)ctx->sm_count, 1024, smem, F, ctx->bin_list, ctx->bin_count, ctx->bin_offs, ctx->bin_next, ctx->cta_bin));
        else
            CK(launch_pdl(ctx, (k_weight_lean<1024, false>), (unsigned)ctx->sm_count, 1024, smem, F, ctx->bin_list, ctx->bin_count, ctx->bin_offs, ctx->bin_next, ctx->cta_bin));
        LAUNCHED();
        if (ev1) CK(cudaEventRecordWithFlags(ev1, ctx->stream, ctx->capturing ? cudaEventRecordExternal : cudaEventRecordDefault));
        ctx->hint_n = 0;
    } else {
        if (step_or_null) {
            if (patches) CK(launch_pdl(ctx, (k_project_weight<true, true, true>), grid, WT_THREADS, smem, P));
            else CK(launch_pdl(ctx, (k_project_weight<true, true, false>), grid, WT_THREADS, smem, P));
        } else if (patches) CK(launch_pdl(ctx, (k_project_weight<false, true, true>), grid, WT_THREADS, smem, P));
        else CK(launch_pdl(ctx, (k_project_weight<false, true, false>), grid, WT_THREADS, smem, P));
        LAUNCHED();
        ctx->hint_n = grid;
        if (ev1) CK(cudaEventRecord(ev1, ctx->stream));
    }
    ctx->last_lean = lean;
    TilePartArgs ta; ta.part = ctx->l_part_a; ta.n_tiles = ctx->t_local; ta.scratch_a = ctx->scratch_a; ta.scratch_b = ctx->scratch_b;
    ta.hint = ctx->hint; ta.n_hint = ctx->hint_n; ta.n_local = ctx->n; ta.bin_count = ctx->bin_count; ta.bin_next = ctx->bin_next;
    ta.bump_project = (!lean && step_or_null) ? 1 : 0;
    CK(launch_pdl(ctx, k_tile_part, ctx->t_local, SCAN_THREADS, 0, ctx->st, ctx->sc, ctx->n, ta, ctx->X));
    LAUNCHED();
    // cloud centre (tile placement of the generic path): copied to the host on a side stream, read by later launches
    // without waiting -- a hint
    if (!lean) {
        CK(cudaEventRecord(ctx->side_ev, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->side, ctx->side_ev, 0));
        CK(cudaMemcpyAsync(ctx->sc_async, ctx->sc, sizeof(DevScalars), cudaMemcpyDeviceToHost, ctx->side));
    }
    return launch_tail(ctx, true, 1, 1, 0);
}

// The step's UpdateResult.  k_chain publishes it into host-mapped memory the moment the resampling decision is taken; the host
// waits for that sequence number only, so the call returns while k_cdf and the resampling kernel are still running -- they overlap
// the caller's next step, and every later call is ordered after them on the context's stream.  Errors raised by those trailing
// kernels (a peer that never arrived) surface in the next result.  ESLAM_GPU_RESULT_SYNC=1: wait for the whole step instead.
static int fetch_result(eslam_gpu_ctx* ctx, eslam_gpu_update_result* r) {
    const uint64_t want = ctx->chain_launched;
    volatile ResultBox* b = ctx->box_host;
    for (unsigned spin = 1; b->seq < want; ++spin) {
        if ((spin & 2047u) == 0) {
            const cudaError_t e = cudaStreamQuery(ctx->stream);
            if (e == cudaSuccess) {
                if (b->seq >= want) break;
                ctx->error = "the step finished without publishing its result"; return ESLAM_ERR_CUDA;
            }
            if (e != cudaErrorNotReady) CK(e);
        }
#if defined(__x86_64__)
        __builtin_ia32_pause();
#endif
    }
    std::atomic_thread_fence(std::memory_order_acquire);
    if (ctx->env.result_sync) CK(cudaStreamSynchronize(ctx->stream));
    ResultBox s;
    s.ess = b->ess; s.max_weight = b->max_weight; s.floating_weight = b->floating_weight; s.sum_w = b->sum_w;
    s.data_particles = b->data_particles; s.total_points = b->total_points; s.do_resample = b->do_resample; s.error_flags = b->error_flags;
    r->ess = s.ess; r->max_weight = s.max_weight; r->floating_weight = s.floating_weight; r->weights_sum = s.sum_w;
    r->data_particles = s.data_particles; r->total_points = s.total_points; r->resampled = s.do_resample;
    r->status = ESLAM_OK;
    if (s.error_flags & (ERR_XCHG_TIMEOUT_BIT | ERR_INTERNAL_BIT)) {
        r->status = ESLAM_ERR_COMM;
        ctx->error = (s.error_flags & ERR_XCHG_TIMEOUT_BIT) ? "a rank did not arrive at a peer-memory exchange within 4 s"
                                                            : "internal consistency check of the exact running sum failed";
        CK(cudaMemsetAsync(&ctx->sc->error_flags, 0, sizeof(int), ctx->stream));
    }
    if (s.error_flags & ERR_ZERO_VARIANCE_BIT) {
        r->status = ESLAM_ERR_ZERO_VARIANCE;
        ctx->error = "using a zero measurement variance leads to singularities";
        CK(cudaMemsetAsync(&ctx->sc->error_flags, 0, sizeof(int), ctx->stream));
    }
    return r->status;
}

static int check_ready(eslam_gpu_ctx* ctx, bool need_map) {
    if (!ctx) return ESLAM_ERR_INVALID;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return ESLAM_ERR_CUDA;
    if (ctx->n == 0) { ctx->error = "empty particle set"; return ESLAM_ERR_EMPTY; }
    if (need_map && !ctx->has_map) { ctx->error = "No environment attached."; return ESLAM_ERR_NO_ENVIRONMENT; }
    return ESLAM_OK;
}

static int launch_centroid(eslam_gpu_ctx* ctx) {
    k_centroid_tiles<<<ctx->t_local, 1024, 0, ctx->stream>>>(ctx->st, ctx->sc, ctx->n, base_index(ctx), ctx->l_cent,
                                                            ctx->l_best_idx, ctx->l_best_w);
    LAUNCHED();
    AG(ctx->l_cent, ctx->g_cent, ctx->t_local * 5 * sizeof(double));
    AG(ctx->l_best_idx, ctx->g_best_idx, ctx->t_local * sizeof(unsigned long long));
    AG(ctx->l_best_w, ctx->g_best_w, ctx->t_local * sizeof(double));
    k_centroid_final<<<1, 1024, 0, ctx->stream>>>(ctx->g_cent, ctx->t_global, ctx->scratch_a, ctx->scratch_b, ctx->g_best_idx,
                                                  ctx->g_best_w, ctx->sc);
    LAUNCHED();
    return read_scalars(ctx);
}

extern "C" {

int eslam_gpu_project(eslam_gpu_ctx* ctx, const eslam_gpu_step* step) {
    NvtxRange nvtx_("eslam project");
    int st = check_ready(ctx, false);
    if (st != ESLAM_OK) return st;
    if (!step) return ESLAM_ERR_INVALID;
    WeightParams P;
    fill_weight_params(ctx, &P, step, nullptr);
    k_project_weight<true, false, false><<<weight_grid(ctx), WT_THREADS, 0, ctx->stream>>>(P);
    LAUNCHED();
    k_bump_project<<<1, 1, 0, ctx->stream>>>(ctx->sc);
    LAUNCHED();
    return ESLAM_OK;
}

int eslam_gpu_update(eslam_gpu_ctx* ctx, const eslam_gpu_contact* pts, int n_pts, eslam_gpu_update_result* result) {
    NvtxRange nvtx_("eslam update");
    int st = check_ready(ctx, true);
    if (st != ESLAM_OK) return st;
    ContactSet cs;
    st = build_contact_set(ctx, pts, n_pts, &cs);
    if (st != ESLAM_OK) return st;
    st = launch_update(ctx, nullptr, &cs, nullptr, nullptr);
    if (st != ESLAM_OK) return st;
    return result ? fetch_result(ctx, result) : ESLAM_OK;
}

}  // extern "C"

static void graph_drop(eslam_gpu_ctx* ctx) {
    if (ctx->sg.exec) cudaGraphExecDestroy(ctx->sg.exec);
    if (ctx->sg.graph) cudaGraphDestroy(ctx->sg.graph);
    ctx->sg.exec = nullptr; ctx->sg.graph = nullptr; ctx->sg.valid = false;
}

// launch-bound sizes only: with a million particles the step's ~9 launches hide behind 0.4 ms of kernels
static bool graph_wanted(const eslam_gpu_ctx* ctx) {
    if (ctx->timing || ctx->sg.failed || ctx->env.graph == 0) return false;
    if (ctx->env.graph == 1) return true;
    return ctx->n <= 262144;
}

// One fused step through the graph: capture on first use (or when the step's structure changed), afterwards update the three
// parameter blocks that carry the call's inputs and launch.  Returns ESLAM_OK with *done = false when the graph does not apply
// (the generic kernel, a failed capture): the caller launches the kernels directly.
static int graph_step(eslam_gpu_ctx* ctx, const eslam_gpu_step* step, const ContactSet* cs, bool* done, bool launch = true) {
    *done = false;
    eslam_gpu_ctx::StepGraph& G = ctx->sg;
    WeightParams P;
    fill_weight_params(ctx, &P, step, cs);
    FastParams& F = *ctx->fast;
    if (!fill_fast_params(ctx, &F, P, *cs)) return ESLAM_OK;              // generic path: tile placement is a host-side hint per launch
    if (G.valid && (G.n != ctx->n || G.nb != F.fc.nb || G.lean_threads != ctx->env.lean_threads)) graph_drop(ctx);
    if (!G.valid) {
        if (!G.ev0) { CK(cudaEventCreate(&G.ev0)); CK(cudaEventCreate(&G.ev1)); }
        CK(cudaStreamSynchronize(ctx->stream));
        if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) != cudaSuccess) { cudaGetLastError(); G.failed = true; return ESLAM_OK; }
        ctx->capturing = true;
        const uint64_t before = ctx->launches;
        const int st = launch_update(ctx, step, cs, G.ev0, G.ev1);
        ctx->capturing = false;
        cudaGraph_t g = nullptr;
        const cudaError_t e = cudaStreamEndCapture(ctx->stream, &g);
        G.kernels = (unsigned)(ctx->launches - before); ctx->launches = before;
        if (st != ESLAM_OK || e != cudaSuccess || !g) { cudaGetLastError(); if (g) cudaGraphDestroy(g); G.failed = true; return st; }
        G.graph = g;
        if (cudaGraphInstantiate(&G.exec, G.graph, 0) != cudaSuccess) { cudaGetLastError(); graph_drop(ctx); G.failed = true; return ESLAM_OK; }
        // the nodes whose parameters carry the call's inputs
        size_t nn = 0; CK(cudaGraphGetNodes(G.graph, nullptr, &nn));
        std::vector<cudaGraphNode_t> nodes(nn); CK(cudaGraphGetNodes(G.graph, nodes.data(), &nn));
        G.n_project = G.n_scatter = G.n_lean = nullptr;
        for (size_t i = 0; i < nn; ++i) {
            cudaGraphNodeType ty; CK(cudaGraphNodeGetType(nodes[i], &ty));
            if (ty != cudaGraphNodeTypeKernel) continue;
            cudaKernelNodeParams kp; CK(cudaGraphKernelNodeGetParams(nodes[i], &kp));
            if (kp.func == (void*)k_project_bin<true>) G.n_project = nodes[i];
            else if (kp.func == (void*)k_bin_scatter) G.n_scatter = nodes[i];
            else if (kp.func == (void*)k_weight_lean<1024, true> || kp.func == (void*)k_weight_lean<1024, false> || kp.func == (void*)k_weight_lean<512, false>) G.n_lean = nodes[i];
        }
        if (!G.n_project || !G.n_scatter || !G.n_lean) { graph_drop(ctx); G.failed = true; return ESLAM_OK; }
        G.valid = true; G.n = ctx->n; G.nb = F.fc.nb; G.lean_threads = ctx->env.lean_threads; G.step = *step; G.cs = *cs;
    } else if (memcmp(&G.step, step, sizeof(*step)) != 0 || memcmp(&G.cs, cs, sizeof(*cs)) != 0) {
        // new odometry step / contact points: the three kernels that take them
        cudaKernelNodeParams kp;
        int nb = F.fc.nb, n_cta = ctx->sm_count, did = 1; unsigned long long n = ctx->n;
        void* a_project[] = {&F, &ctx->bin_count, &ctx->keys};
        CK(cudaGraphKernelNodeGetParams(G.n_project, &kp)); kp.kernelParams = a_project; kp.extra = nullptr;
        CK(cudaGraphExecKernelNodeSetParams(G.exec, G.n_project, &kp));
        void* a_scatter[] = {&ctx->bin_count, &nb, &ctx->keys, &n, &ctx->bin_list, &ctx->bin_offs, &ctx->cta_bin, &n_cta, &ctx->sc, &did};
        CK(cudaGraphKernelNodeGetParams(G.n_scatter, &kp)); kp.kernelParams = a_scatter; kp.extra = nullptr;
        CK(cudaGraphExecKernelNodeSetParams(G.exec, G.n_scatter, &kp));
        void* a_lean[] = {&F, &ctx->bin_list, &ctx->bin_count, &ctx->bin_offs, &ctx->bin_next, &ctx->cta_bin};
        CK(cudaGraphKernelNodeGetParams(G.n_lean, &kp)); kp.kernelParams = a_lean; kp.extra = nullptr;
        CK(cudaGraphExecKernelNodeSetParams(G.exec, G.n_lean, &kp));
        G.step = *step; G.cs = *cs;
    }
    *done = true;
    if (!launch) return ESLAM_OK;                                          // prepared (captured, parameters current): nothing has run
    CK(cudaGraphLaunch(G.exec, ctx->stream));
    ctx->launches += G.kernels; ctx->chain_launched++;
    ctx->last_lean = true; ctx->hint_n = 0;
    return ESLAM_OK;
}

extern "C" {

int eslam_gpu_step_fused(eslam_gpu_ctx* ctx, const eslam_gpu_step* step, const eslam_gpu_contact* pts, int n_pts,
                         eslam_gpu_update_result* result) {
    NvtxRange nvtx_("eslam step (project + update)");
    int st = check_ready(ctx, true);
    if (st != ESLAM_OK) return st;
    if (!step) return ESLAM_ERR_INVALID;
    ContactSet cs;
    st = build_contact_set(ctx, pts, n_pts, &cs);
    if (st != ESLAM_OK) return st;
    bool done = false;
    if (graph_wanted(ctx)) { st = graph_step(ctx, step, &cs, &done); if (st != ESLAM_OK) return st; }
    if (!done) st = launch_update(ctx, step, &cs, nullptr, nullptr);
    if (st != ESLAM_OK) return st;
    return result ? fetch_result(ctx, result) : ESLAM_OK;
}

int eslam_gpu_run_steps(eslam_gpu_ctx* ctx, const eslam_gpu_step* step, const eslam_gpu_contact* pts, int n_pts, int steps,
                        float* elapsed_ms, float* top_kernel_ms, eslam_gpu_update_result* last_result) {
    NvtxRange nvtx_("eslam run_steps");
    int st = check_ready(ctx, true);
    if (st != ESLAM_OK) return st;
    if (!step || steps <= 0) return ESLAM_ERR_INVALID;
    ContactSet cs;
    st = build_contact_set(ctx, pts, n_pts, &cs);
    if (st != ESLAM_OK) return st;
    const size_t need = 2 + 2 * (size_t)steps;
    while (ctx->events.size() < need) { cudaEvent_t e; CK(cudaEventCreate(&e)); ctx->events.push_back(e); }
    if (sharded(ctx)) {
        // all ranks enter the timed region together: a rendezvous on the streams (flags over peer memory), so that no rank's
        // events include waiting for a rank that was enqueued later
        k_barrier<<<1, 32, 0, ctx->stream>>>(ctx->sc, ctx->X);
        LAUNCHED();
    }
    bool graphed = false;
    if (graph_wanted(ctx)) { st = graph_step(ctx, step, &cs, &graphed, false); if (st != ESLAM_OK) return st; }   // capture outside the timed region
    CK(cudaEventRecord(ctx->events[0], ctx->stream));
    ctx->marks_used = 0;
    if (graphed) {
        // launch-bound sizes: every step is one graph launch.  The graph holds ONE pair of events around the dominant kernel
        // (re-recorded by every replay), so its time is that of the last step, scaled to the run; step periods: the mean.
        for (int s = 0; s < steps; ++s) { bool done = false; st = graph_step(ctx, step, &cs, &done); if (st != ESLAM_OK) return st; }
        CK(cudaEventRecord(ctx->events[1], ctx->stream));
        CK(cudaEventSynchronize(ctx->events[1]));
        float ms = 0; CK(cudaEventElapsedTime(&ms, ctx->events[0], ctx->events[1]));
        if (elapsed_ms) *elapsed_ms = ms;
        if (top_kernel_ms) { float t = 0; CK(cudaEventElapsedTime(&t, ctx->sg.ev0, ctx->sg.ev1)); *top_kernel_ms = t * (float)steps; }
        ctx->step_ms.assign((size_t)(steps > 1 ? steps - 1 : 0), ms / (float)steps);
        return last_result ? fetch_result(ctx, last_result) : ESLAM_OK;
    }
    // The events around the dominant kernel sit between launches and so take the programmatic overlap away at those two
    // places (~4 us per step): they are recorded only when the caller asks for the kernel's time (top_kernel_ms != NULL), then
    // on every step.  A run that measures the step itself passes NULL and carries no event between its first and its last launch.
    const bool sampling = top_kernel_ms != nullptr;
    const int every = 1;
    for (int s = 0; s < steps; ++s) {
        st = launch_update(ctx, step, &cs, sampling ? ctx->events[2 + 2 * s] : nullptr, sampling ? ctx->events[3 + 2 * s] : nullptr);
        if (st != ESLAM_OK) return st;
    }
    CK(cudaEventRecord(ctx->events[1], ctx->stream));
    CK(cudaEventSynchronize(ctx->events[1]));
    timing_report(ctx);
    if (elapsed_ms) CK(cudaEventElapsedTime(elapsed_ms, ctx->events[0], ctx->events[1]));
    if (top_kernel_ms) {
        float tot = 0; int cnt = 0;
        for (int s = 0; s < steps; s += every) { float t = 0; CK(cudaEventElapsedTime(&t, ctx->events[2 + 2 * s], ctx->events[3 + 2 * s])); tot += t; ++cnt; }
        *top_kernel_ms = tot * (float)steps / (float)cnt;
    }
    ctx->step_ms.clear();
    for (int s = 0; sampling && s + every < steps; s += every) {
        float t = 0; CK(cudaEventElapsedTime(&t, ctx->events[2 + 2 * s], ctx->events[2 + 2 * (s + every)]));
        ctx->step_ms.push_back(t / (float)every);
    }
    return last_result ? fetch_result(ctx, last_result) : ESLAM_OK;
}

int eslam_gpu_step_times(eslam_gpu_ctx* ctx, float* out_ms, int cap) {
    if (!ctx || (!out_ms && cap > 0)) return -1;
    const int n = (int)std::min<size_t>(ctx->step_ms.size(), (size_t)(cap > 0 ? cap : 0));
    for (int i = 0; i < n; ++i) out_ms[i] = ctx->step_ms[i];
    return (int)ctx->step_ms.size();
}

int eslam_gpu_io_bytes(const eslam_gpu_ctx* ctx, int n_pts, uint64_t out[2]) {
    if (!ctx || !out || n_pts < 0) return ESLAM_ERR_INVALID;
    out[0] = sizeof(eslam_gpu_step) + (uint64_t)n_pts * sizeof(eslam_gpu_contact);   // host inputs of one eslam_gpu_step_fused call
    out[1] = sizeof(ResultBox);                                                       // what k_chain writes into host memory for the UpdateResult
    return ESLAM_OK;
}

int eslam_gpu_weights_sum(eslam_gpu_ctx* ctx, double* sum) {
    int st = check_ready(ctx, false);
    if (st != ESLAM_OK) return st;
    if (!sum) return ESLAM_ERR_INVALID;
    ShardPlan sp0; memset(&sp0, 0, sizeof(sp0));
    LoopBArgs lb; lb.tile_sum = ctx->l_tile_sum; lb.n_tiles = ctx->t_local; lb.discount = ctx->cfg.discount_factor;
    lb.scratch_a = ctx->scratch_a; lb.scratch_b = ctx->scratch_b; lb.prefix_raw = ctx->prefix_raw; lb.shard_sums = ctx->shard_sums;
    CK(launch_pdl(ctx, k_loop_b<false>, ctx->t_local, SCAN_THREADS, 0, ctx->st, ctx->sc, ctx->n, lb, sp0, ctx->X));
    LAUNCHED();
    k_sum_finish<<<1, 32, 0, ctx->stream>>>(ctx->sc, sp0, ctx->shard_sums, ctx->X);
    LAUNCHED();
    st = read_scalars(ctx);
    if (st != ESLAM_OK) return st;
    *sum = ctx->sc_host->sum_w;
    return ESLAM_OK;
}

int eslam_gpu_normalize_weights(eslam_gpu_ctx* ctx, double* ess) {
    int st = check_ready(ctx, false);
    if (st != ESLAM_OK) return st;
    st = launch_tail(ctx, false, 1, 2, 0);
    if (st != ESLAM_OK) return st;
    if (ess) { st = read_scalars(ctx); if (st != ESLAM_OK) return st; *ess = ctx->sc_host->ess; }
    return ESLAM_OK;
}

int eslam_gpu_resample(eslam_gpu_ctx* ctx) {
    NvtxRange nvtx_("eslam resample");
    int st = check_ready(ctx, false);
    if (st != ESLAM_OK) return st;
    return launch_tail(ctx, false, 0, 0, 1);
}

int eslam_gpu_resample_n(eslam_gpu_ctx* ctx, uint64_t samples, int multinomial) {
    int st = check_ready(ctx, false);
    if (st != ESLAM_OK) return st;
    if (samples == 0) return ESLAM_ERR_INVALID;
    if (sharded(ctx)) { ctx->error = "resample_stratified(samples) / resample_multinomial are not available on sharded contexts"; return ESLAM_ERR_INVALID; }
    if (samples > ctx->cap) { ctx->error = "more samples than the capacity the context was initialised with"; return ESLAM_ERR_INVALID; }
    if (samples == ctx->n && !multinomial) return launch_tail(ctx, false, 0, 0, 1);   // resample(): the step's own kernels
    eslam_gpu_config saved = ctx->cfg;
    ctx->cfg.shard_mode = ESLAM_SHARD_EXACT_GLOBAL;                       // one walk over all particles
    st = launch_tail(ctx, false, 0, 0, 1, true);                          // exact running sum of the weights as they are
    ctx->cfg = saved;
    if (st != ESLAM_OK) return st;
    const ModelParams mp = model_params(ctx);
    ResampleNArgs A; memset(&A, 0, sizeof(A));
    A.cdf = ctx->cdf; A.tile_start = tile_sin_table(ctx); A.n_tiles = ctx->t_local; A.n = ctx->n; A.samples = samples;
    A.k0key = mp.k0; A.k1key = mp.k1; A.systematic = ctx->cfg.resample_mode == ESLAM_RESAMPLE_SYSTEMATIC; A.multinomial = multinomial;
    A.ancestors = ctx->ancestors; A.P = ctx->cfg.patch_size > 0 ? ctx->cfg.patch_size : 0;
    for (int b = 0; b < 2; ++b) { A.patch[b] = ctx->patch[b]; A.porigin[b] = ctx->porigin[b]; }
    k_resample_n<<<(unsigned)((samples + 255) / 256), 256, 0, ctx->stream>>>(ctx->st, ctx->sc, A);
    LAUNCHED();
    k_after_resample_n<<<1, 1, 0, ctx->stream>>>(ctx->sc);
    LAUNCHED();
    CK(cudaStreamSynchronize(ctx->stream));
    graph_drop(ctx);
    ctx->n = samples; ctx->t_local = (unsigned)((samples + TILE - 1) / TILE); ctx->t_global = ctx->t_local;
    return ESLAM_OK;
}

int eslam_gpu_get_ancestors(eslam_gpu_ctx* ctx, uint32_t* a) {
    int st = check_ready(ctx, false);
    if (st != ESLAM_OK) return st;
    if (!a) return ESLAM_ERR_INVALID;
    CK(cudaMemcpyAsync(a, ctx->ancestors, ctx->n * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ESLAM_OK;
}

int eslam_gpu_get_cdf(eslam_gpu_ctx* ctx, double* c) {
    int st = check_ready(ctx, false);
    if (st != ESLAM_OK) return st;
    if (!c) return ESLAM_ERR_INVALID;
    CK(cudaMemcpyAsync(c, ctx->cdf, ctx->n * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ESLAM_OK;
}

int eslam_gpu_centroid(eslam_gpu_ctx* ctx, double out[4]) {
    int st = check_ready(ctx, false);
    if (st != ESLAM_OK) return st;
    if (!out) return ESLAM_ERR_INVALID;
    st = launch_tail(ctx, false, 1, 2, 0);                        // normalizeWeights(), src/PoseEstimator.cpp:356
    if (st != ESLAM_OK) return st;
    st = launch_centroid(ctx);
    if (st != ESLAM_OK) return st;
    const double* c = ctx->sc_host->centroid;
    for (int i = 0; i < 4; ++i) out[i] = c[i] / c[4];             // :373-375
    return ESLAM_OK;
}

int eslam_gpu_best_index(eslam_gpu_ctx* ctx, uint64_t* index) {
    int st = check_ready(ctx, false);
    if (st != ESLAM_OK) return st;
    if (!index) return ESLAM_ERR_INVALID;
    st = launch_centroid(ctx);
    if (st != ESLAM_OK) return st;
    *index = ctx->sc_host->best_index;
    return ESLAM_OK;
}

int eslam_gpu_evaluate_poses(eslam_gpu_ctx* ctx, uint64_t n, const double* x, const double* y, const double* theta,
                             const double* zpos, const double* meas_var, const double* z_var_in,
                             const eslam_gpu_contact* pts, int n_pts,
                             uint8_t* ok, double* weight, double* z_delta, double* z_var, double* pose_var,
                             int32_t* n_contacts, double* zpos_out, double* z_var_out, int32_t* cell_index, double* cp) {
    if (!ctx || n == 0 || !x || !y || !theta || !zpos || !meas_var) return ESLAM_ERR_INVALID;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return ESLAM_ERR_CUDA;
    if (!ctx->has_map) { ctx->error = "No environment attached."; return ESLAM_ERR_NO_ENVIRONMENT; }
    EvalParams P; memset(&P, 0, sizeof(P));
    int st = build_contact_set(ctx, pts, n_pts, &P.cs);
    if (st != ESLAM_OK) return st;
    const size_t nd = n * sizeof(double);
    struct Temps {                                                       // freed on every path out of this function
        double* din = nullptr; double* dout = nullptr; int32_t* dint = nullptr; uint8_t* dok = nullptr; int* derr = nullptr;
        ~Temps() { cudaFree(din); cudaFree(dout); cudaFree(dint); cudaFree(dok); cudaFree(derr); }
    } tmp;
    double*& din = tmp.din; double*& dout = tmp.dout; int32_t*& dint = tmp.dint; uint8_t*& dok = tmp.dok; int*& derr = tmp.derr;
    const size_t cpn = n * ESLAM_MAX_CP * 5, celln = n * (size_t)(n_pts > 0 ? n_pts : 1);
    CK(dalloc(&din, 6 * n)); CK(dalloc(&dout, 6 * n + cpn)); CK(dalloc(&dint, n + celln)); CK(dalloc(&dok, n)); CK(dalloc(&derr, 1));
    const double* src[6] = {x, y, theta, zpos, meas_var, z_var_in};
    for (int i = 0; i < 6; ++i)
        if (src[i]) CK(cudaMemcpyAsync(din + i * n, src[i], nd, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaMemsetAsync(derr, 0, sizeof(int), ctx->stream));
    CK(cudaMemsetAsync(dout, 0, (6 * n + cpn) * sizeof(double), ctx->stream));
    CK(cudaMemsetAsync(dint, 0xFF, (n + celln) * sizeof(int32_t), ctx->stream));
    CK(cudaMemsetAsync(dok, 0, n, ctx->stream));
    P.n = n; P.x = din; P.y = din + n; P.th = din + 2 * n; P.z = din + 3 * n; P.mv = din + 4 * n;
    P.zvin = z_var_in ? din + 5 * n : nullptr;
    P.map = ctx->map; P.map.tile_w = 0; P.map.tile_h = 0; P.htab = ctx->htab;
    const ModelParams mp = model_params(ctx);
    P.inv_corr = mp.inv_corr; P.min_contacts = mp.min_contacts; P.use_shape = mp.use_shape;
    P.ok = dok; P.weight = dout; P.zdelta = dout + n; P.zvar = dout + 2 * n; P.posevar = dout + 3 * n;
    P.zout = dout + 4 * n; P.zvout = dout + 5 * n; P.cp = dout + 6 * n; P.ncont = dint; P.cells = dint + n; P.err = derr;
    k_evaluate_poses<<<(unsigned)((n + 127) / 128), 128, 0, ctx->stream>>>(P);
    LAUNCHED();
    int herr = 0;
    CK(cudaMemcpyAsync(&herr, derr, sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    double* dst[6] = {weight, z_delta, z_var, pose_var, zpos_out, z_var_out};
    for (int i = 0; i < 6; ++i)
        if (dst[i]) CK(cudaMemcpyAsync(dst[i], dout + i * n, nd, cudaMemcpyDeviceToHost, ctx->stream));
    if (cp) CK(cudaMemcpyAsync(cp, dout + 6 * n, cpn * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
    if (ok) CK(cudaMemcpyAsync(ok, dok, n, cudaMemcpyDeviceToHost, ctx->stream));
    if (n_contacts) CK(cudaMemcpyAsync(n_contacts, dint, n * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    if (cell_index && n_pts > 0) CK(cudaMemcpyAsync(cell_index, dint + n, celln * sizeof(int32_t), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    if (herr & ERR_ZERO_VARIANCE_BIT) {
        ctx->error = "using a zero measurement variance leads to singularities";
        return ESLAM_ERR_ZERO_VARIANCE;
    }
    return ESLAM_OK;
}

int eslam_gpu_debug_counters(eslam_gpu_ctx* ctx, uint64_t out[4]) {
    if (!ctx || !out) return ESLAM_ERR_INVALID;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return ESLAM_ERR_CUDA;
    int st = read_scalars(ctx);
    if (st != ESLAM_OK) return st;
    out[0] = ctx->sc_host->fallbacks; out[1] = (uint64_t)ctx->sc_host->n_boundaries;
    out[2] = (uint64_t)ctx->sc_host->error_flags; out[3] = ctx->sc_host->resample_step;
    return ESLAM_OK;
}

int eslam_gpu_path_counters(eslam_gpu_ctx* ctx, uint64_t out[2]) {
    if (!ctx || !out) return ESLAM_ERR_INVALID;
    if (cudaSetDevice(ctx->device) != cudaSuccess) return ESLAM_ERR_CUDA;
    int st = read_scalars(ctx);
    if (st != ESLAM_OK) return st;
    out[0] = ctx->last_lean ? ctx->n : 0; out[1] = ctx->last_lean ? ctx->sc_host->defer_last : ctx->n;
    return ESLAM_OK;
}

int eslam_gpu_process_map(eslam_gpu_ctx* ctx, const eslam_gpu_scan* scan, int match, int update) {
    NvtxRange nvtx_("eslam processMap");
    int st = check_ready(ctx, true);
    if (st != ESLAM_OK) return st;
    if (!scan || !scan->cells_mean_stdev || scan->width <= 0 || scan->height <= 0) return ESLAM_ERR_INVALID;
    if (ctx->cfg.patch_size <= 0) {
        ctx->error = "processMap needs per-particle maps (patch_size > 0: EmbodiedSlamFilter::init with useSharedMap == false)";
        return ESLAM_ERR_INVALID;
    }
    // the scan cells that hold a surface, in row-major order, with their centres in the scan frame
    std::vector<ScanCell> cells;
    cells.reserve((size_t)scan->width * scan->height);
    for (int j = 0; j < scan->height; ++j)
        for (int k = 0; k < scan->width; ++k) {
            const float m = scan->cells_mean_stdev[2 * ((size_t)j * scan->width + k)], sd = scan->cells_mean_stdev[2 * ((size_t)j * scan->width + k) + 1];
            if (m != m || sd != sd) continue;
            ScanCell c; c.x = scan->offset_x + (k + 0.5) * scan->scale_x; c.y = scan->offset_y + (j + 0.5) * scan->scale_y; c.mean = m; c.stdev = sd;
            cells.push_back(c);
        }
    if (cells.empty()) return ESLAM_OK;
    if (cells.size() > ctx->scan_cap) {
        cudaFree(ctx->scan_cells); ctx->scan_cells = nullptr; ctx->scan_cap = 0;
        CK(cudaMalloc((void**)&ctx->scan_cells, cells.size() * sizeof(ScanCell)));
        ctx->scan_cap = cells.size();
    }
    CK(cudaMemcpyAsync(ctx->scan_cells, cells.data(), cells.size() * sizeof(ScanCell), cudaMemcpyHostToDevice, ctx->stream));
    ProcessMapArgs A; memset(&A, 0, sizeof(A));
    const double sigma = 0.2;                                            // src/EmbodiedSlamFilter.cpp:216-217
    A.cells = reinterpret_cast<const ScanCell*>(ctx->scan_cells); A.n_cells = (unsigned)cells.size();
    A.n_samples = (unsigned)std::min<size_t>(10, cells.size());          // sampling = 10  :216
    A.match = match; A.update = update; A.P = ctx->cfg.patch_size;
    A.two_sigma2 = 2.0 * sigma * sigma; A.visual = 0.1;                  // visualWeighting :219
    for (int b = 0; b < 2; ++b) { A.patch[b] = ctx->patch[b]; A.porigin[b] = ctx->porigin[b]; }
    A.grid = ctx->cells; A.map = ctx->map;
    k_process_map<<<(unsigned)((ctx->n * 32 + 255) / 256), 256, 0, ctx->stream>>>(ctx->st, ctx->sc, ctx->n, A);
    LAUNCHED();
    CK(cudaStreamSynchronize(ctx->stream));                              // `cells` (host) is borrowed until the copy has run
    return ESLAM_OK;
}

int eslam_gpu_pose_moments(eslam_gpu_ctx* ctx, eslam_gpu_gaussian2d* out) {
    int st = check_ready(ctx, false);
    if (st != ESLAM_OK) return st;
    if (!out) return ESLAM_ERR_INVALID;
    st = launch_centroid(ctx);                                           // sums of w x, w y, ..., w over all shards
    if (st != ESLAM_OK) return st;
    const double W = ctx->sc_host->centroid[4];
    const double mx = ctx->sc_host->centroid[0] / W, my = ctx->sc_host->centroid[1] / W;
    k_moment_tiles<<<ctx->t_local, 1024, 0, ctx->stream>>>(ctx->st, ctx->sc, ctx->n, mx, my, ctx->l_cent);
    LAUNCHED();
    AG(ctx->l_cent, ctx->g_cent, ctx->t_local * 5 * sizeof(double));
    k_centroid_final<<<1, 1024, 0, ctx->stream>>>(ctx->g_cent, ctx->t_global, ctx->scratch_a, ctx->scratch_b, ctx->g_best_idx,
                                                  ctx->g_best_w, ctx->sc);
    LAUNCHED();
    st = read_scalars(ctx);
    if (st != ESLAM_OK) return st;
    const double* c = ctx->sc_host->centroid;
    out->weight = W; out->mean[0] = mx; out->mean[1] = my;
    out->cov[0] = c[0] / W; out->cov[1] = out->cov[2] = c[1] / W; out->cov[3] = c[2] / W;
    return ESLAM_OK;
}

// ---- SurfaceHash (src/SurfaceHash.hpp:113-232; src/PoseEstimator.cpp:75-86,130-182)
int eslam_gpu_hash_create(eslam_gpu_ctx* ctx, int slope_bins, int angular_steps) {
    NvtxRange nvtx_("eslam SurfaceHash::create");
    if (!ctx) return ESLAM_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (!ctx->has_map) { ctx->error = "No environment attached."; return ESLAM_ERR_NO_ENVIRONMENT; }
    if (!ctx->map.identity || slope_bins <= 0 || slope_bins > 255 || angular_steps <= 0 || angular_steps > 255 || ctx->map.W > 65535 || ctx->map.H > 65535) {
        ctx->error = "SurfaceHash needs an identity map frame, at most 255 slope bins / headings and a grid of at most 65535 cells a side";
        return ESLAM_ERR_INVALID;
    }
    cudaFree(ctx->hash_poses); cudaFree(ctx->hash_by_bucket); ctx->hash_poses = nullptr; ctx->hash_by_bucket = nullptr; ctx->hash_n = 0;
    ctx->hash_bins = slope_bins; ctx->hash_steps = angular_steps; ctx->hash_step = 0;
    HashArgs A; A.map = ctx->map; A.bins = slope_bins; A.steps = angular_steps;
    const unsigned rows = (unsigned)angular_steps * (unsigned)ctx->map.H, nb = (unsigned)slope_bins * slope_bins;
    struct Temps { unsigned *cnt = nullptr, *offs = nullptr, *bcnt = nullptr, *idx = nullptr; unsigned short *keys = nullptr, *keys2 = nullptr; void* tmp = nullptr;
                   ~Temps() { cudaFree(cnt); cudaFree(offs); cudaFree(bcnt); cudaFree(idx); cudaFree(keys); cudaFree(keys2); cudaFree(tmp); } } T;
    CK(dalloc(&T.cnt, (size_t)rows + 1)); CK(dalloc(&T.offs, (size_t)rows + 1)); CK(dalloc(&T.bcnt, (size_t)nb));
    CK(cudaMemsetAsync(T.cnt, 0, (rows + 1) * sizeof(unsigned), ctx->stream));
    CK(cudaMemsetAsync(T.bcnt, 0, nb * sizeof(unsigned), ctx->stream));
    k_hash_rows<false><<<rows, 256, 0, ctx->stream>>>(A, T.cnt, nullptr, nullptr, nullptr);
    LAUNCHED();
    size_t tb = 0;
    cub::DeviceScan::ExclusiveSum(nullptr, tb, T.cnt, T.offs, (int)rows + 1, ctx->stream);
    CK(cudaMalloc(&T.tmp, tb));
    cub::DeviceScan::ExclusiveSum(T.tmp, tb, T.cnt, T.offs, (int)rows + 1, ctx->stream);
    unsigned total = 0;
    CK(cudaMemcpyAsync(&total, T.offs + rows, sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    ctx->hash_counts.assign(nb, 0); ctx->hash_offs.assign(nb + 1, 0);
    if (total == 0) return ESLAM_OK;
    CK(dalloc(&ctx->hash_poses, (size_t)total)); CK(dalloc(&ctx->hash_by_bucket, (size_t)total));
    k_hash_rows<true><<<rows, 256, 0, ctx->stream>>>(A, nullptr, T.offs, ctx->hash_poses, T.bcnt);
    LAUNCHED();
    // candidate indices grouped by bucket, creation order inside a bucket: a stable sort of (bucket, index)
    CK(dalloc(&T.keys, (size_t)total)); CK(dalloc(&T.keys2, (size_t)total)); CK(dalloc(&T.idx, (size_t)total));
    k_hash_keys<<<(total + 255) / 256, 256, 0, ctx->stream>>>(ctx->hash_poses, total, T.keys, T.idx);
    LAUNCHED();
    cudaFree(T.tmp); T.tmp = nullptr; tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, T.keys, T.keys2, T.idx, ctx->hash_by_bucket, (int)total, 0, 16, ctx->stream);
    CK(cudaMalloc(&T.tmp, tb));
    cub::DeviceRadixSort::SortPairs(T.tmp, tb, T.keys, T.keys2, T.idx, ctx->hash_by_bucket, (int)total, 0, 16, ctx->stream);
    CK(cudaMemcpyAsync(ctx->hash_counts.data(), T.bcnt, nb * sizeof(unsigned), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaGetLastError());
    for (unsigned b = 0; b < nb; ++b) ctx->hash_offs[b + 1] = ctx->hash_offs[b] + ctx->hash_counts[b];
    ctx->hash_n = total;
    return ESLAM_OK;
}

int eslam_gpu_hash_info(eslam_gpu_ctx* ctx, uint64_t* n_poses, uint32_t* bucket_counts) {
    if (!ctx) return ESLAM_ERR_INVALID;
    if (n_poses) *n_poses = ctx->hash_n;
    if (bucket_counts) for (size_t b = 0; b < ctx->hash_counts.size(); ++b) bucket_counts[b] = ctx->hash_counts[b];
    return ESLAM_OK;
}

int eslam_gpu_hash_get(eslam_gpu_ctx* ctx, uint64_t first, uint64_t count, double* x, double* y, double* theta, double* z, int32_t* bucket) {
    if (!ctx || first + count > ctx->hash_n) return ESLAM_ERR_INVALID;
    if (count == 0) return ESLAM_OK;
    CK(cudaSetDevice(ctx->device));
    std::vector<HashPose> h(count);
    CK(cudaMemcpyAsync(h.data(), ctx->hash_poses + first, count * sizeof(HashPose), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    const MapParams& M = ctx->map;
    for (uint64_t i = 0; i < count; ++i) {
        if (x) x[i] = M.ox + (h[i].m + 0.5) * M.sx;
        if (y) y[i] = M.oy + (h[i].n + 0.5) * M.sy;
        if (theta) theta[i] = h[i].a * 2.0 * M_PI / ctx->hash_steps;
        if (z) z[i] = h[i].z;
        if (bucket) bucket[i] = h[i].bucket;
    }
    return ESLAM_OK;
}

int eslam_gpu_init_from_hash(eslam_gpu_ctx* ctx, uint64_t n) {
    if (!ctx || n == 0) return ESLAM_ERR_INVALID;
    CK(cudaSetDevice(ctx->device));
    if (ctx->hash_n == 0) { ctx->error = "could not sample from pose hash."; return ESLAM_ERR_EMPTY; }   // src/PoseEstimator.cpp:84
    int st = ensure_capacity(ctx, n);
    if (st != ESLAM_OK) return st;
    DevScalars init; memset(&init, 0, sizeof(init));
    CK(cudaStreamSynchronize(ctx->side));
    *ctx->sc_async = init; *ctx->sc_host = init;
    CK(cudaMemcpyAsync(ctx->sc, ctx->sc_host, sizeof(DevScalars), cudaMemcpyHostToDevice, ctx->stream));
    if (sharded(ctx)) CK(cudaMemsetAsync(ctx->xchg, 0, ctx->x_bytes, ctx->stream));   // the exchange epochs restart with the scalars
    st = reset_patches(ctx);
    if (st != ESLAM_OK) return st;
    ctx->hash_step = 0;
    const ModelParams mp = model_params(ctx);
    k_init_from_hash<<<(unsigned)((n + 255) / 256), 256, 0, ctx->stream>>>(ctx->st, ctx->sc, n, base_index(ctx), mp.k0, mp.k1, ctx->hash_poses, ctx->hash_n,
                                                                          ctx->map, ctx->hash_steps);
    LAUNCHED();
    CK(cudaStreamSynchronize(ctx->stream));
    return ESLAM_OK;
}

}  // extern "C"

// ContactModel::lowestPointHeuristic (src/ContactModel.cpp:48-79) on the rotated contact points: ungrouped points stand for
// themselves, a run of one group id is represented by its lowest member (the first one on ties)
static int lowest_points_per_group(const eslam_gpu_contact* pts, int n, double* lx, double* ly, double* lz) {
    int count = 0;
    for (int first = 0; first < n;) {
        int low = first, last = first;
        if (pts[first].group_id >= 0) {
            while (last + 1 < n && pts[last + 1].group_id == pts[first].group_id) ++last;
            for (int k = first + 1; k <= last; ++k) if (pts[k].position[2] < pts[low].position[2]) low = k;
        }
        lx[count] = pts[low].position[0]; ly[count] = pts[low].position[1]; lz[count] = pts[low].position[2]; ++count;
        first = last + 1;
    }
    return count;
}

extern "C" {

int eslam_gpu_sample_from_hash(eslam_gpu_ctx* ctx, double replace_percentage, double avg_factor, const eslam_gpu_contact* pts, int n_pts,
                               uint64_t* replaced) {
    if (replaced) *replaced = 0;
    int st = check_ready(ctx, false);
    if (st != ESLAM_OK) return st;
    if (!pts || n_pts < 0 || n_pts > ESLAM_MAX_CONTACTS) return ESLAM_ERR_INVALID;
    if (sharded(ctx)) { ctx->error = "sampleFromHash is not available on sharded contexts"; return ESLAM_ERR_INVALID; }
    if (ctx->hash_n == 0) { ctx->error = "no pose hash (eslam_gpu_hash_create)"; return ESLAM_ERR_EMPTY; }
    double lx[ESLAM_MAX_CONTACTS], ly[ESLAM_MAX_CONTACTS], lz[ESLAM_MAX_CONTACTS];
    const int nl = lowest_points_per_group(pts, n_pts, lx, ly, lz);                   // src/PoseEstimator.cpp:139-140
    double slx = 0, sly = 0;
    const uint32_t hstep = ctx->hash_step++;
    if (nl < 3 || !em::surface_slopes(lx, ly, lz, nl, slx, sly)) return ESLAM_OK;     // :143-144
    const int b = em::slope_bucket(slx, ctx->hash_bins) * ctx->hash_bins + em::slope_bucket(sly, ctx->hash_bins);
    const unsigned in_bucket = ctx->hash_counts[b];
    const double relevance = 1.0 - 1.0 * in_bucket / ctx->hash_n;                     // SurfaceHash::getRelevance :139-144
    const double relevance_factor = relevance * relevance * relevance;                // pow(., 3) :156
    size_t replace_count = (size_t)(ctx->n * replace_percentage * relevance_factor);  // :157
    if (relevance_factor < 0.8) replace_count = 0;                                    // :158-159
    if (replace_count == 0 || in_bucket == 0) return ESLAM_OK;                        // nothing to replace / sample(params) == NULL
    if (replace_count > ctx->n) replace_count = (size_t)ctx->n;
    double sum = 0;
    st = eslam_gpu_weights_sum(ctx, &sum);                                            // getWeightAvg() :162
    if (st != ESLAM_OK) return st;
    const double weight = (sum / (double)ctx->n) * avg_factor * relevance_factor;
    // the particles by ascending weight (as float, ties by index: std::sort of pair<float, index> :147-153)
    const unsigned n = (unsigned)ctx->n;
    struct Temps { float *k = nullptr, *k2 = nullptr; unsigned *i = nullptr, *i2 = nullptr; void* tmp = nullptr;
                   ~Temps() { cudaFree(k); cudaFree(k2); cudaFree(i); cudaFree(i2); cudaFree(tmp); } } T;
    CK(dalloc(&T.k, (size_t)n)); CK(dalloc(&T.k2, (size_t)n)); CK(dalloc(&T.i, (size_t)n)); CK(dalloc(&T.i2, (size_t)n));
    k_weight_keys<<<(n + 255) / 256, 256, 0, ctx->stream>>>(ctx->st, ctx->sc, n, T.k, T.i);
    LAUNCHED();
    size_t tb = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, tb, T.k, T.k2, T.i, T.i2, (int)n, 0, 32, ctx->stream);
    CK(cudaMalloc(&T.tmp, tb));
    cub::DeviceRadixSort::SortPairs(T.tmp, tb, T.k, T.k2, T.i, T.i2, (int)n, 0, 32, ctx->stream);
    const ModelParams mp = model_params(ctx);
    k_hash_replace<<<(unsigned)((replace_count + 255) / 256), 256, 0, ctx->stream>>>(ctx->st, ctx->sc, T.i2, (unsigned)replace_count, mp.k0, mp.k1, hstep,
                                                                                     ctx->hash_poses, ctx->hash_by_bucket, ctx->hash_offs[b], in_bucket,
                                                                                     ctx->map, ctx->hash_steps, weight);
    LAUNCHED();
    CK(cudaStreamSynchronize(ctx->stream));
    if (replaced) *replaced = replace_count;
    return ESLAM_OK;
}

int eslam_gpu_get_patches(eslam_gpu_ctx* ctx, float* cells, int32_t* origin_xy) {
    if (!ctx) return ESLAM_ERR_INVALID;
    if (ctx->cfg.patch_size <= 0 || ctx->n == 0) {
        ctx->error = "per-particle map patches are not enabled in this context (patch_size == 0)";
        return ESLAM_ERR_INVALID;
    }
    CK(cudaSetDevice(ctx->device));
    int st = read_scalars(ctx);
    if (st != ESLAM_OK) return st;
    const int cur = ctx->sc_host->cur;
    const size_t pp = (size_t)ctx->cfg.patch_size * ctx->cfg.patch_size;
    if (cells) CK(cudaMemcpyAsync(cells, ctx->patch[cur], ctx->n * pp * sizeof(float2), cudaMemcpyDeviceToHost, ctx->stream));
    if (origin_xy) CK(cudaMemcpyAsync(origin_xy, ctx->porigin[cur], 2 * ctx->n * sizeof(int), cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return ESLAM_OK;
}

}  // extern "C"
