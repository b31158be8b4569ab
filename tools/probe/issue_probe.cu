// Does an fp64 instruction take one issue slot or two?  1024 threads (8 warps per scheduler), per iteration: ND independent
// DFMA chains + NF independent FFMA chains, every chain long enough apart to hide latency.  cycles/iteration/warp-scheduler.
#include <cstdio>
#include <cuda_runtime.h>
template <int ND, int NF> __global__ void k(double* out, long long* cyc, double b, float g) {
    double a[8]; float f[16];
    for (int i = 0; i < 8; ++i) a[i] = threadIdx.x + i; for (int i = 0; i < 16; ++i) f[i] = threadIdx.x + i;
    __syncthreads();
    long long t0 = clock64();
    for (int it = 0; it < 4096; ++it) {
#pragma unroll
        for (int i = 0; i < ND; ++i) a[i] = __fma_rn(a[i], b, 1e-9);
#pragma unroll
        for (int i = 0; i < NF; ++i) f[i] = __fmaf_rn(f[i], g, 1e-9f);
    }
    long long t1 = clock64();
    double s = 0; for (int i = 0; i < 8; ++i) s += a[i]; for (int i = 0; i < 16; ++i) s += f[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int ND, int NF> void run(double* out, long long* cyc) {
    k<ND, NF><<<1, 1024>>>(out, cyc, 1.0000001, 1.0001f);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    // 8 warps per scheduler: cycles per (warp-iteration) on one scheduler = c / 4096 / 8
    printf("DFMA %d + FFMA %2d per iteration: %.2f issue cycles per warp-iteration (instructions: %d)\n", ND, NF, (double)c / 4096 / 8, ND + NF);
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 1024 * 8); cudaMalloc(&cyc, 8);
    run<8, 0>(out, cyc); run<0, 16>(out, cyc); run<4, 4>(out, cyc); run<4, 8>(out, cyc); run<4, 12>(out, cyc); run<4, 16>(out, cyc);
    run<8, 8>(out, cyc); run<8, 16>(out, cyc); run<2, 16>(out, cyc);
    return 0;
}
