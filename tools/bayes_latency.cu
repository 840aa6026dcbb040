// Dependent-chain latency of the occupancy update (two IEEE double divisions) on this GPU.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ double clampProb(double v) { const double lo = 1e-3, hi = 1.0 - 1e-3; return v < lo ? lo : (hi < v ? hi : v); }
__device__ __forceinline__ double bayesUpdate(double v, double p, double oddsP) {
    if (v == 0.0) return clampProb(p);
    const double cv = clampProb(v);
    const double oldOdds = __ddiv_rn(cv, __dsub_rn(1.0, cv));
    const double o = __dmul_rn(oldOdds, oddsP);
    const double nv = clampProb(__ddiv_rn(o, __dadd_rn(1.0, o)));
    return clampProb(nv);
}
__global__ void chain(double* out, long long* cyc, int n, double oh, double om) {
    double v = 0.5;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) v = bayesUpdate(v, (i & 1) ? 0.6 : 0.45, (i & 1) ? oh : om);
    long long t1 = clock64();
    out[threadIdx.x] = v; if (threadIdx.x == 0) *cyc = t1 - t0;
}
__global__ void chain_dfma(double* out, long long* cyc, int n, double a, double b) {
    double v = 0.5;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) v = fma(v, a, b);
    long long t1 = clock64();
    out[threadIdx.x] = v; if (threadIdx.x == 0) *cyc = t1 - t0;
}
int main() {
    double* out; long long* cyc; cudaMalloc(&out, 8 * 1024); cudaMallocManaged(&cyc, 8);
    for (int threads : {1, 32, 256}) {
        chain<<<1, threads>>>(out, cyc, 10000, 0.6 / 0.4, 0.45 / 0.55); cudaDeviceSynchronize();
        printf("bayesUpdate chain, %d thread(s): %.1f cycles per update\n", threads, *cyc / 10000.0);
        chain_dfma<<<1, threads>>>(out, cyc, 10000, 0.999, 1e-4); cudaDeviceSynchronize();
        printf("dependent DFMA chain, %d thread(s): %.1f cycles per DFMA\n", threads, *cyc / 10000.0);
    }
    return 0;
}
