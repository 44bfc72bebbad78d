// Issue rates of the integer instructions the u16 prefilter is made of.  sm_100a.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o int_rate int_rate.cu && ./int_rate
// Every trip issues 16 independent instructions of one kind (8 chains per thread, 8 warps per SMSP hide the
// latencies), or a mix of two kinds, and reports cycles per warp instruction and SMSP.
//   OP 0 IDP.2A (dp2a)   1 IMAD (32-bit, accumulate)   2 IMAD.WIDE.U32 + 64-bit accumulate   3 IADD3   4 PRMT   5 LOP3
#include <cstdio>
#include <cuda_runtime.h>

template <int OP>
__device__ __forceinline__ void step(unsigned& q, unsigned long long& w, unsigned salt) {
    if (OP == 0) q = __dp2a_lo(q, salt, q);
    if (OP == 1) asm volatile("mad.lo.u32 %0, %0, %1, %0;" : "+r"(q) : "r"(salt));
    if (OP == 2) asm volatile("mad.wide.u32 %0, %1, %1, %0;" : "+l"(w) : "r"(q));
    if (OP == 3) asm volatile("add.u32 %0, %0, %1;" : "+r"(q) : "r"(salt));
    if (OP == 4) q = __byte_perm(q, salt, 0x3120);
    if (OP == 5) q = (q & salt) ^ 0x55aa55aau;
}

template <int OPA, int NA, int OPB, int NB>
__global__ void k(unsigned* out, long long* cyc, int iters, unsigned salt) {
    unsigned q[8];
    unsigned long long w[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { q[j] = threadIdx.x * 7 + j + salt; w[j] = j; }
    long long t0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int j = 0; j < (NA > NB ? NA : NB); ++j) {
            if (j < NA) step<OPA>(q[j & 7], w[j & 7], salt);
            if (j < NB) step<OPB>(q[(j + 4) & 7], w[(j + 4) & 7], salt);
        }
    }
    long long t1 = clock64();
    unsigned u = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) u ^= q[j] ^ (unsigned)w[j] ^ (unsigned)(w[j] >> 32);
    if (u == 0x12345u) out[0] = u;
    if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OPA, int NA, int OPB, int NB>
void run(const char* name) {
    unsigned* out; long long* cyc;
    cudaMalloc(&out, 8); cudaMalloc(&cyc, 8);
    const int iters = 1 << 14, warps = 32;
    k<OPA, NA, OPB, NB><<<148, warps * 32>>>(out, cyc, iters, 0x0101u);
    cudaDeviceSynchronize();
    k<OPA, NA, OPB, NB><<<148, warps * 32>>>(out, cyc, iters, 0x0101u);
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double per_warp = (double)h / iters / (warps / 4.0);
    printf("%-34s %6.2f cycles per trip per warp = %.2f per instruction\n", name, per_warp, per_warp / (NA + NB));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    run<0, 16, 0, 0>("16 IDP.2A");
    run<1, 16, 1, 0>("16 IMAD");
    run<2, 16, 2, 0>("16 IMAD.WIDE.U32 (64-bit acc)");
    run<3, 16, 3, 0>("16 IADD3");
    run<4, 16, 4, 0>("16 PRMT");
    run<5, 16, 5, 0>("16 LOP3 pairs");
    run<0, 12, 4, 4>("12 IDP.2A + 4 PRMT");
    run<0, 12, 3, 12>("12 IDP.2A + 12 IADD3");
    run<1, 12, 3, 12>("12 IMAD + 12 IADD3");
    run<0, 8, 1, 8>("8 IDP.2A + 8 IMAD");
    run<2, 8, 3, 16>("8 IMAD.WIDE + 16 IADD3");
    return 0;
}
