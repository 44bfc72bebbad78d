// DFMA dependent-issue latency and per-warp throughput on sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dfma_lat dfma_lat.cu && ./dfma_lat
// One CTA per SM, W warps per CTA, NCH independent chains per thread, unroll 1 so every loop trip
// issues NCH independent DFMAs that each depend on the previous trip.
#include <cstdio>
#include <cuda_runtime.h>

template <int NCH>
__global__ void k(double* out, long long* cyc, int iters, double m, double c) {
    double x[NCH];
#pragma unroll
    for (int j = 0; j < NCH; ++j) x[j] = threadIdx.x + j;
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < NCH; ++j) x[j] = fma(x[j], m, c);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < NCH; ++j) s += x[j];
    if (s == 123.456) out[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int NCH>
void run(int warps) {
    double* out; long long* cyc;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    const int iters = 1 << 16;
    k<NCH><<<148, warps * 32>>>(out, cyc, iters, 0.999999, 1e-9);
    cudaDeviceSynchronize();
    k<NCH><<<148, warps * 32>>>(out, cyc, iters, 0.999999, 1e-9);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_trip = (double)h / iters;
    printf("warps/SM %2d (per SMSP %4.1f) chains %d: %.2f cycles/trip, %.2f cycles per DFMA per warp, SMSP DFMA/cycle %.3f\n",
           warps, warps / 4.0, NCH, per_trip, per_trip / NCH, (warps / 4.0) * NCH / per_trip);
    cudaFree(out); cudaFree(cyc);
}

int main() {
    for (int w : {4, 8, 16, 20}) {
        run<1>(w); run<2>(w); run<4>(w); run<8>(w);
    }
    return 0;
}
