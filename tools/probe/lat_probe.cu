// Dependent-issue latency of the instructions the lean weight kernel chains (one warp, clock64 around N dependent ops).
#include <cstdio>
#include <cuda_runtime.h>
#define N 2048
__constant__ double cdata[256];
template <int OP> __global__ void k(double* out, long long* cyc, double a0, float f0, int idx) {
    double a = a0, b = 1.0000001, c = 1e-9; float f = f0, g = 1.0001f; unsigned u = (unsigned)idx;
    __shared__ double sm[256];
    for (int i = threadIdx.x; i < 256; i += blockDim.x) sm[i] = (double)((i * 7 + 1) & 255);
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; ++i) {
        if (OP == 0) a = __fma_rn(a, b, c);
        if (OP == 1) a = __dmul_rn(a, b);
        if (OP == 2) a = __dadd_rn(a, c);
        if (OP == 3) f = __fmaf_rn(f, g, 1e-9f);
        if (OP == 4) { f = (float)a; a = (double)f + 0.0; }                 // F2F.F32.F64 + F2F.F64.F32 (+DADD)
        if (OP == 5) { u = (unsigned)sm[u & 255]; }                          // LDS + F2I
        if (OP == 6) { u = (unsigned)cdata[u & 255]; }                       // LDC indexed + F2I
        if (OP == 7) { a = (double)(float)a; }                               // two conversions
        if (OP == 8) { u = u * 3u + 1u; }                                    // IMAD
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { cyc[OP] = t1 - t0; out[OP] = a + f + u; }
}
int main() {
    double h[256]; for (int i = 0; i < 256; ++i) h[i] = (double)((i * 7 + 1) & 255);
    cudaMemcpyToSymbol(cdata, h, sizeof h);
    double* out; long long* cyc; cudaMalloc(&out, 128); cudaMalloc(&cyc, 128);
    k<0><<<1, 32>>>(out, cyc, 1.0, 1.f, 3); k<1><<<1, 32>>>(out, cyc, 1.0, 1.f, 3); k<2><<<1, 32>>>(out, cyc, 1.0, 1.f, 3);
    k<3><<<1, 32>>>(out, cyc, 1.0, 1.f, 3); k<4><<<1, 32>>>(out, cyc, 1.0, 1.f, 3); k<5><<<1, 32>>>(out, cyc, 1.0, 1.f, 3);
    k<6><<<1, 32>>>(out, cyc, 1.0, 1.f, 3); k<7><<<1, 32>>>(out, cyc, 1.0, 1.f, 3); k<8><<<1, 32>>>(out, cyc, 1.0, 1.f, 3);
    long long c[16]; cudaMemcpy(c, cyc, 9 * 8, cudaMemcpyDeviceToHost);
    const char* nm[] = {"DFMA", "DMUL", "DADD", "FFMA", "F2F.F32.F64+F2F.F64.F32+DADD", "LDS+F2I(f64)", "LDC(indexed)+F2I(f64)", "F2F+F2F", "IMAD"};
    for (int i = 0; i < 9; ++i) printf("%-32s %.2f cycles per iteration\n", nm[i], (double)c[i] / N);
    // throughput: 8 warps per scheduler (1024 threads), independent chains
    return 0;
}
