// TMA probe: load a (bw x bh) box of 8-byte cells from a W x H grid into shared memory; variants of descriptor placement.
#include <cuda.h>
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
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
struct Params { int x0, y0, bw, bh; int pad[12]; alignas(64) CUtensorMap tmap; };
__global__ void k_param(const __grid_constant__ Params P, double* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(smem);
    double* tile = reinterpret_cast<double*>(smem + 128);
    if (threadIdx.x == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(mbar, P.bw * P.bh * 8); tma_load_2d((uint32_t)__cvta_generic_to_shared(tile), &P.tmap, mbar, P.x0, P.y0); }
    mbar_wait(mbar, 0);
    for (int i = threadIdx.x; i < P.bw * P.bh; i += blockDim.x) out[i] = tile[i];
}
__global__ void k_global(const CUtensorMap* tm, int x0, int y0, int bw, int bh, double* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const uint32_t mbar = (uint32_t)__cvta_generic_to_shared(smem);
    double* tile = reinterpret_cast<double*>(smem + 128);
    if (threadIdx.x == 0) mbar_init(mbar, 1);
    __syncthreads();
    if (threadIdx.x == 0) { mbar_expect_tx(mbar, bw * bh * 8); tma_load_2d((uint32_t)__cvta_generic_to_shared(tile), tm, mbar, x0, y0); }
    mbar_wait(mbar, 0);
    for (int i = threadIdx.x; i < bw * bh; i += blockDim.x) out[i] = tile[i];
}
int main(int argc, char** argv) {
    const int W = 2048, H = 2048; int bw = argc > 1 ? atoi(argv[1]) : 158, bh = argc > 2 ? atoi(argv[2]) : 158;
    std::vector<double> h((size_t)W * H);
    for (size_t i = 0; i < h.size(); i++) h[i] = (double)i;
    double* d; cudaMalloc(&d, h.size() * 8); cudaMemcpy(d, h.data(), h.size() * 8, cudaMemcpyHostToDevice);
    double* out; cudaMalloc(&out, (size_t)bw * bh * 8);
    typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                 const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
    void* fn = nullptr; cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    Params P; memset(&P, 0, sizeof(P)); P.x0 = 100; P.y0 = 200; P.bw = bw; P.bh = bh;
    cuuint64_t gdim[2] = {(cuuint64_t)W, (cuuint64_t)H}; cuuint64_t gstr[1] = {(cuuint64_t)W * 8};
    cuuint32_t box[2] = {(cuuint32_t)bw, (cuuint32_t)bh}; cuuint32_t es[2] = {1, 1};
    CUresult r = ((EncodeFn)fn)(&P.tmap, CU_TENSOR_MAP_DATA_TYPE_UINT64, 2, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("encode -> %d, box %dx%d, sizeof(Params)=%zu\n", (int)r, bw, bh, sizeof(Params));
    const size_t smem = 128 + (size_t)bw * bh * 8;
    cudaFuncSetAttribute(k_param, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    cudaFuncSetAttribute(k_global, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    std::vector<double> o((size_t)bw * bh);
    for (int variant = 0; variant < 2; variant++) {
        cudaMemset(out, 0, o.size() * 8);
        if (variant == 0) k_param<<<1, 256, smem>>>(P, out);
        else { CUtensorMap* dtm; cudaMalloc(&dtm, sizeof(CUtensorMap)); cudaMemcpy(dtm, &P.tmap, sizeof(CUtensorMap), cudaMemcpyHostToDevice); k_global<<<1, 256, smem>>>(dtm, P.x0, P.y0, bw, bh, out); }
        cudaError_t e = cudaDeviceSynchronize();
        printf("variant %d (%s): %s\n", variant, variant ? "descriptor in global memory" : "descriptor in __grid_constant__ param", cudaGetErrorString(e));
        if (e != cudaSuccess) return 1;
        cudaMemcpy(o.data(), out, o.size() * 8, cudaMemcpyDeviceToHost);
        size_t bad = 0;
        for (int yy = 0; yy < bh; yy++) for (int xx = 0; xx < bw; xx++) if (o[(size_t)yy * bw + xx] != (double)((size_t)(P.y0 + yy) * W + P.x0 + xx)) bad++;
        printf("  mismatches: %zu\n", bad);
    }
    return 0;
}
