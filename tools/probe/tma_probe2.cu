// TMA probe 2: mimic k_project_weight's layout: mbarrier at byte 32000, tile at 32128, 1024 threads, many CTAs.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>
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
struct Big { int x0, y0, bw, bh, W, H, mbar_ofs, tile_ofs; const CUtensorMap* tm; double pad[400]; };
__global__ void __launch_bounds__(1024, 1) k(const __grid_constant__ Big P, double* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(smem + P.mbar_ofs);
    double* tile = reinterpret_cast<double*>(smem + P.tile_ofs);
    double* htab = reinterpret_cast<double*>(smem);
    const int x0 = max(0, min(P.x0, P.W - P.bw)), y0 = max(0, min(P.y0, P.H - P.bh));
    if (threadIdx.x == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(mbar, (uint32_t)(P.bw * P.bh) * 8u); tma_load_2d((uint32_t)__cvta_generic_to_shared(tile), P.tm, mbar, x0, y0); }
    for (int i = threadIdx.x; i < 3840; i += blockDim.x) htab[i] = P.pad[i % 400];
    __syncthreads();
    mbar_wait(mbar, 0);
    if (blockIdx.x == 0) for (int i = threadIdx.x; i < P.bw * P.bh; i += blockDim.x) out[i] = tile[i];
}
int main(int argc, char** argv) {
    int W = argc > 1 ? atoi(argv[1]) : 2048, bw = argc > 2 ? atoi(argv[2]) : 158, bh = argc > 3 ? atoi(argv[3]) : 158;
    int threads = argc > 4 ? atoi(argv[4]) : 1024, grid = argc > 5 ? atoi(argv[5]) : 148;
    int mbar_ofs = argc > 6 ? atoi(argv[6]) : 32000, tile_ofs = argc > 7 ? atoi(argv[7]) : 32128;
    const int H = W;
    std::vector<double> h((size_t)W * H);
    for (size_t i = 0; i < h.size(); i++) h[i] = (double)i;
    double* d; cudaMalloc(&d, h.size() * 8); cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    double* out; cudaMalloc(&out, (size_t)bw * bh * 8);
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    CUtensorMap tm;
    cuuint64_t gdim[2] = {(cuuint64_t)W, (cuuint64_t)H}; cuuint64_t gstr[1] = {(cuuint64_t)W * 8};
    cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}; cuuint32_t es[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(&tm, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    { const unsigned long long* qq = reinterpret_cast<const unsigned long long*>(&tm); printf("tmap d=%p:", (void*)d); for (int i = 0; i < 16; ++i) printf(" %016llx", qq[i]); printf("\n"); }
    CUtensorMap* dtm; cudaMalloc(&dtm, sizeof(tm)); cudaMemcpy(dtm, &tm, sizeof(tm), cudaMemcpyHostToDevice);
    Big P; memset(&P, 0, sizeof(P)); P.x0 = 100; P.y0 = 20; P.bw = bw; P.bh = bh; P.W = W; P.H = H; P.mbar_ofs = mbar_ofs; P.tile_ofs = tile_ofs; P.tm = dtm;
    const size_t smem = tile_ofs + (size_t)bw * bh * 8;
    cudaError_t e = cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    printf("encode %d W %d box %dx%d threads %d grid %d mbar@%d tile@%d smem %zu attr:%s sizeof(Big)=%zu\n", (int)r, W, bw, bh, threads, grid, mbar_ofs, tile_ofs, smem, cudaGetErrorString(e), sizeof(Big));
    k<<<grid, threads, smem>>>(P, out);
    e = cudaDeviceSynchronize();
    printf("  run: %s\n", cudaGetErrorString(e));
    if (e != cudaSuccess) return 1;
    std::vector<double> o((size_t)bw * bh);
    cudaMemcpy(o.data(), out, o.size() * 8, cudaMemcpyDeviceToHost);
    const int x0 = std::max(0, std::min(P.x0, W - bw)), y0 = std::max(0, std::min(P.y0, H - bh));
    size_t bad = 0;
    for (int yy = 0; yy < bh; yy++) for (int xx = 0; xx < bw; xx++) if (o[(size_t)yy * bw + xx] != (double)((size_t)(y0 + yy) * W + x0 + xx)) bad++;
    printf("  mismatches: %zu\n", bad);
    return 0;
}
