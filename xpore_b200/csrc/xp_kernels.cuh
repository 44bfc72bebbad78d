// xp_kernels.cuh — sm_100a kernels of the diffmod hot path.
//
//   K1  xp_prefilter_kernel   two-sample t-test per position (statstest.py:4-18)  — HBM streaming
//   W*  xp_worklist_*         selected positions, longest first (replaces the gene queue,
//                             scripts/diffmod.py:152-175, helper.py:96-116)
//   K3  xp_fit_kernel         VB-EM of the 2-component mixture (gmm.py:11-127)  — xp_fit.cuh: one
//                             warp (or a 4-warp team) per position, reads resident in shared memory
//                             for all sweeps, persistent CTAs pulling tickets from a global worklist
//   K4  xp_stats_kernel       cluster assignment, z-tests, overlap, row filter (io.py:255-367)
//
// FP64 throughout (the reference is float64 and its stopping rule rounds an ELBO difference at
// 1e-8 absolute, gmm.py:110).  No tensor cores: nothing here is a dense contraction.
#pragma once
#include <cuda_runtime.h>

#include "xp_core.cuh"

namespace xpk {

using namespace xpc;

constexpr int kMaxGroups = 64;
constexpr int kMaxConditions = 32;
constexpr int kBuckets = 1024;   // worklist buckets: [0, 512) by read count, [512, 1024) single-warp class by expected sweeps x read count
constexpr unsigned kFull = 0xffffffffu;

// ------------------------------------------------------------------------------------------
// small device helpers
// ------------------------------------------------------------------------------------------
// shuffles of a double through the 32-bit intrinsics (the toolkit's double overloads split and rejoin the value
// with volatile asm, which costs two extra moves per shuffle in SASS)
__device__ __forceinline__ double shfl_xor_f64(double v, int m) {
    const int lo = __shfl_xor_sync(kFull, __double2loint(v), m);
    const int hi = __shfl_xor_sync(kFull, __double2hiint(v), m);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double shfl_f64(double v, int src) {
    const int lo = __shfl_sync(kFull, __double2loint(v), src);
    const int hi = __shfl_sync(kFull, __double2hiint(v), src);
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) v += shfl_xor_f64(v, m);
    return v;
}
__device__ __forceinline__ uint64_t warp_min_u64(uint64_t v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        const uint64_t o = __shfl_xor_sync(kFull, v, m);
        v = o < v ? o : v;
    }
    return v;
}
__device__ __forceinline__ uint64_t warp_max_u64(uint64_t v) {
#pragma unroll
    for (int m = 16; m > 0; m >>= 1) {
        const uint64_t o = __shfl_xor_sync(kFull, v, m);
        v = o > v ? o : v;
    }
    return v;
}

// order-preserving map double -> uint64
__device__ __forceinline__ uint64_t dkey(double x) {
    const uint64_t u = d2u(x);
    return (u >> 63) ? ~u : (u | 0x8000000000000000ull);
}
__device__ __forceinline__ double keyd(uint64_t k) {
    return u2d((k >> 63) ? (k & 0x7FFFFFFFFFFFFFFFull) : ~k);
}

// ---- mbarrier + 1-D TMA bulk copy (global -> shared) ----------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "XP_WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra XP_DONE_%=;\n"
        "bra XP_WAIT_%=;\n"
        "XP_DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}

// ------------------------------------------------------------------------------------------
// K1: prefilter t-test (statstest.py:4-21; the gate is scripts/diffmod.py:52-58).
// ------------------------------------------------------------------------------------------
// Decode of the scaled-integer read formats: y = q / scale, correctly rounded, without a division
// instruction sequence.  With rs = RN(1/scale):  a = q*rs;  r = fma(-a, scale, q) (exact residual);
// y = fma(r, rs, a)  (Markstein's correction step; equal to IEEE q/scale - verified exhaustively over
// the u16 range and on random i32 values in tests/test_gpu_parity.py).
enum { kYF64 = 0, kYI32 = 1, kYU16 = 2 };
__device__ __forceinline__ int y_elem_bytes(int fmt) { return fmt == kYF64 ? 8 : (fmt == kYI32 ? 4 : 2); }
__device__ __forceinline__ double y_decode(double q, double scale, double rscale) {
    const double a = q * rscale;
    const double r = fma(-a, scale, q);
    return fma(r, rscale, a);
}
__device__ __forceinline__ double y_load_global(const void* y, int fmt, double scale, double rscale, int64_t i) {
    if (fmt == kYF64) return __ldg(reinterpret_cast<const double*>(y) + i);
    if (fmt == kYI32) return y_decode((double)__ldg(reinterpret_cast<const int*>(y) + i), scale, rscale);
    return y_decode((double)__ldg(reinterpret_cast<const unsigned short*>(y) + i), scale, rscale);
}

struct PrefilterArgs {
    const void* y;
    int y_format;
    double y_scale;
    const int64_t* read_off;
    const int32_t* grp_len;
    const double* kmer_mean;
    const int32_t* grp_cond;
    int P, G;
    int enabled;       // 1 = t-test, 0 = prefilter configured with an unknown method (NaN => fit all)
    double threshold;
    uint32_t* status;  // [P] written: SELECTED | TTEST_VALID
    double* pval;      // [P]
};

// One warp owns a tile of 32 consecutive positions.  Four lanes share one position (8 positions in
// flight per warp, four rounds per tile): they stream its reads in 8-read blocks (one aligned
// 16-byte load per block for the u16 format, two blocks ahead), walk the group boundaries to know
// which condition a read belongs to, and keep the centred moments (n, sum y', sum y'^2) of the two
// conditions present.  A two-step shuffle reduction over the four lanes and one shuffle hand the
// moments of position j to lane j; after the four rounds all 32 lanes evaluate their Student-t
// p-value in parallel.  Fixed summation order: results are deterministic.
#ifndef XP_PF_MINB
#define XP_PF_MINB 6
#endif
constexpr int kPfLanes = 4;                 // lanes per position
constexpr int kPfPerRound = 32 / kPfLanes;  // positions per round

struct PfMoments {
    double s1a, s2a, s1b, s2b;
    int na, nb;
};
// The u16 read format carries integers q = 100 y, so its moments are kept as exact integer sums
// (sum q <= 2^31 and sum q^2 <= 2^47 for up to 32767 reads): three integer instructions per read instead of
// a decode and four FP64 operations, and a t statistic that carries only the final roundings.
struct PfIntMoments {
    unsigned s1a, s1b;
    unsigned long long s2a, s2b;
    int na, nb;
};
constexpr int kPfIntMaxReads = 32767;

// group slots blocked by condition: every condition's slots are adjacent
__device__ __forceinline__ bool pf_cond_sorted(const int* s_cond, int G) {
    bool sorted = true;
    unsigned seen = 1u << s_cond[0];
    for (int g = 1; g < G; ++g) {
        if (s_cond[g] != s_cond[g - 1]) {
            sorted = sorted && !((seen >> s_cond[g]) & 1u);
            seen |= 1u << s_cond[g];
        }
    }
    return sorted;
}

// U16X: the batch is u16-encoded and no position has more than kPfIntMaxReads reads (the host checks
// max_reads), so every position takes the exact-integer path and the FP accumulation is compiled out.
template <int WARPS, bool U16X>
__global__ void __launch_bounds__(WARPS * 32, XP_PF_MINB) xp_prefilter_kernel(PrefilterArgs a) {
    __shared__ int s_cond[kMaxGroups];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = a.G;
    for (int g = threadIdx.x; g < G; g += blockDim.x) s_cond[g] = a.grp_cond[g];
    __syncthreads();
    // u16 batches whose group slots are blocked by condition (every condition's slots are adjacent - all batch
    // builders emit the runs condition by condition) are xp_prefilter_u16_kernel's, launched in front of this one
    if (U16X && pf_cond_sorted(s_cond, G)) return;
    const int tile = blockIdx.x * WARPS + warp;
    const int p0 = tile * 32;
    if (p0 >= a.P) return;
    const int sub = lane / kPfLanes, sl = lane % kPfLanes;
    const double rscale = 1.0 / a.y_scale;

    double k_na = 0, k_s1a = 0, k_s2a = 0, k_nb = 0, k_s1b = 0, k_s2b = 0;
    bool k_valid = false, k_exact = false;
    for (int round = 0; round < kPfLanes; ++round) {
        const int p = p0 + round * kPfPerRound + sub;
        PfMoments m = {0.0, 0.0, 0.0, 0.0, 0, 0};
        PfIntMoments mi = {0u, 0u, 0ull, 0ull, 0, 0};
        bool valid = false, exact = false;
        int n_pos = 0;
        if (p < a.P) {
            const int64_t off = a.read_off[p];
            const int n = (int)(a.read_off[p + 1] - off);
            n_pos = n;
            const double m0 = a.kmer_mean[p];
            const int32_t* glen = a.grp_len + (size_t)p * G;
            // conditions present at this position; the lower-numbered one is side 0
            unsigned cmask = 0;
            for (int g = 0; g < G; ++g)
                if (glen[g] > 0) cmask |= 1u << s_cond[g];
            valid = a.enabled && (__popc(cmask) == 2);
            if (valid) {
                const int ca = __ffs(cmask) - 1;
                int g = 0, gend = glen[0];
                auto add = [&](double v, int i) {  // read i of the position
                    while (i >= gend) {
                        ++g;
                        gend += glen[g];
                    }
                    const double c = v - m0;
                    if (s_cond[g] == ca) {
                        m.s1a += c;
                        m.s2a = fma(c, c, m.s2a);
                        ++m.na;
                    } else {
                        m.s1b += c;
                        m.s2b = fma(c, c, m.s2b);
                        ++m.nb;
                    }
                };
                // Lane sl takes the 8-read blocks sl, sl+4, ... of the position (position-local, so the
                // moments depend neither on the read encoding nor on where the position sits in the
                // buffer).
                bool side_a = (s_cond[0] == ca);  // side of the current group; looked up only at a boundary
                // Integer sums are exact, so their order is free: a lane sums the reads of its current run inside one
                // group (one add and one 64-bit multiply-add per read) and moves the run's sums to the group's side
                // when it crosses into another group.  Counts come from the group lengths.
                unsigned run1 = 0u;
                unsigned long long run2 = 0ull;
                int n_a = 0;
                if (U16X)
                    for (int gg = 0; gg < G; ++gg)
                        if (s_cond[gg] == ca) n_a += glen[gg];
                auto flush = [&]() {
                    if (side_a) {
                        mi.s1a += run1;
                        mi.s2a += run2;
                    } else {
                        mi.s1b += run1;
                        mi.s2b += run2;
                    }
                    run1 = 0u;
                    run2 = 0ull;
                };
                auto advance = [&](int i) {  // make the current group the one holding read i
                    if (i >= gend) {
                        flush();
                        do {
                            ++g;
                            gend += glen[g];
                        } while (i >= gend);
                        side_a = (s_cond[g] == ca);
                    }
                };
                auto addq = [&](unsigned q) {
                    run1 += q;
                    run2 += (unsigned long long)q * q;
                };
                exact = U16X;
                if (U16X || a.y_format == kYU16) {
                    // One 16-byte load per lane and step.  The buffer's aligned 16-byte chunk c holds the
                    // reads [8c - head, 8c - head + 8) of the position, so block b = the tail of chunk b
                    // + the first `head` reads of chunk b+1, which the next lane of the quad has
                    // (lane 0's prefetched chunk for lane 3).
                    const unsigned quad = 0xFu << (lane & ~3);
                    const int64_t e0 = off & ~(int64_t)7;
                    const int head = (int)(off - e0), h2 = head >> 1, hs = (head & 1) * 16;
                    const int nchunk = (head + n + 7) >> 3, nblock = (n + 7) >> 3;
                    const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned short*>(a.y) + e0);
                    const uint4 zero4 = make_uint4(0, 0, 0, 0);
                    uint4 cur = (sl < nchunk) ? __ldg(src + sl) : zero4;
                    uint4 nxt = (sl + kPfLanes < nchunk) ? __ldg(src + sl + kPfLanes) : zero4;
                    for (int b0 = 0; b0 < nblock; b0 += kPfLanes) {
                        const int bq = b0 + sl;
                        const uint4 nn = (bq + 2 * kPfLanes < nchunk) ? __ldg(src + bq + 2 * kPfLanes) : zero4;
                        const uint4 give = (sl == 0) ? nxt : cur;
                        const int from = (lane & ~3) | ((sl + 1) & 3);
                        uint4 got;
                        got.x = __shfl_sync(quad, give.x, from);
                        got.y = __shfl_sync(quad, give.y, from);
                        got.z = __shfl_sync(quad, give.z, from);
                        got.w = __shfl_sync(quad, give.w, from);
                        // words [h2, h2 + 5) of {cur, got}, then a funnel shift by the odd half-word
                        unsigned w0 = cur.x, w1 = cur.y, w2 = cur.z, w3 = cur.w, w4 = got.x, w5 = got.y, w6 = got.z,
                                 w7 = got.w;
                        if (h2 & 1) { w0 = w1; w1 = w2; w2 = w3; w3 = w4; w4 = w5; w5 = w6; w6 = w7; }
                        if (h2 & 2) { w0 = w2; w1 = w3; w2 = w4; w3 = w5; w4 = w6; }
                        const unsigned x[4] = {__funnelshift_r(w0, w1, hs), __funnelshift_r(w1, w2, hs),
                                               __funnelshift_r(w2, w3, hs), __funnelshift_r(w3, w4, hs)};
                        const int ibase = bq << 3;
                        if (U16X) {
                            if (ibase < n) advance(ibase);
                            if (ibase + 8 <= n && ibase + 8 <= gend) {  // a full block inside one group: no per-read checks
#pragma unroll
                                for (int k = 0; k < 8; ++k) addq((x[k >> 1] >> ((k & 1) * 16)) & 0xffffu);
                            } else {
#pragma unroll
                                for (int k = 0; k < 8; ++k) {
                                    if (ibase + k < n) {
                                        advance(ibase + k);
                                        addq((x[k >> 1] >> ((k & 1) * 16)) & 0xffffu);
                                    }
                                }
                            }
                        } else {
#pragma unroll
                            for (int k = 0; k < 8; ++k) {
                                const unsigned q = (x[k >> 1] >> ((k & 1) * 16)) & 0xffffu;
                                if (ibase + k < n) add(y_decode((double)q, a.y_scale, rscale), ibase + k);
                            }
                        }
                        cur = nxt;
                        nxt = nn;
                    }
                    if (U16X) {
                        flush();
                        if (sl == 0) mi.na = n_a;  // the quad reduction below sums the lanes: lane 0 carries the count
                    }
                } else {
                    for (int c = sl * 8; c < n; c += kPfLanes * 8) {  // 8 independent loads in flight
                        double v[8];
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            v[k] = (c + k < n) ? y_load_global(a.y, a.y_format, a.y_scale, rscale, off + c + k) : 0.0;
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            if (c + k < n) add(v[k], c + k);
                    }
                }
            }
        }
        // the four lanes of a position: fixed-order butterfly
#pragma unroll
        for (int d = 1; d < kPfLanes; d <<= 1) {
            if (!U16X) {
                m.s1a += __shfl_xor_sync(kFull, m.s1a, d);
                m.s2a += __shfl_xor_sync(kFull, m.s2a, d);
                m.s1b += __shfl_xor_sync(kFull, m.s1b, d);
                m.s2b += __shfl_xor_sync(kFull, m.s2b, d);
                m.na += __shfl_xor_sync(kFull, m.na, d);
                m.nb += __shfl_xor_sync(kFull, m.nb, d);
                continue;
            }
            mi.s1a += __shfl_xor_sync(kFull, mi.s1a, d);
            mi.s1b += __shfl_xor_sync(kFull, mi.s1b, d);
            mi.s2a += __shfl_xor_sync(kFull, mi.s2a, d);
            mi.s2b += __shfl_xor_sync(kFull, mi.s2b, d);
            mi.na += __shfl_xor_sync(kFull, mi.na, d);
            mi.nb += __shfl_xor_sync(kFull, mi.nb, d);
        }
        if (U16X) mi.nb = n_pos - mi.na;
        if (U16X) {  // exact sums -> the same six doubles (n, sum, sum of squares about the side's own origin)
            m.na = mi.na;
            m.nb = mi.nb;
            m.s1a = (double)mi.s1a;
            m.s1b = (double)mi.s1b;
            // n S2 - S1^2 is exact in 64 bits (< 2^62); stored through s2 as "S2 - S1^2/n + S1^2/n" would
            // round, so the t statistic below takes the exact numerators instead: s2 carries N = n S2 - S1^2.
            m.s2a = (double)((unsigned long long)mi.na * mi.s2a - (unsigned long long)mi.s1a * mi.s1a);
            m.s2b = (double)((unsigned long long)mi.nb * mi.s2b - (unsigned long long)mi.s1b * mi.s1b);
        }
        const bool t_exact = __shfl_sync(kFull, (int)exact, ((lane - round * kPfPerRound) & (kPfPerRound - 1)) * kPfLanes) != 0;
        // position round * 8 + j goes to lane round * 8 + j
        const int src = ((lane - round * kPfPerRound) & (kPfPerRound - 1)) * kPfLanes;
        const double t1a = __shfl_sync(kFull, m.s1a, src), t2a = __shfl_sync(kFull, m.s2a, src);
        const double t1b = __shfl_sync(kFull, m.s1b, src), t2b = __shfl_sync(kFull, m.s2b, src);
        const int tna = __shfl_sync(kFull, m.na, src), tnb = __shfl_sync(kFull, m.nb, src);
        const bool tv = __shfl_sync(kFull, (int)valid, src) != 0;
        if ((lane / kPfPerRound) == round) {
            k_na = tna; k_s1a = t1a; k_s2a = t2a;
            k_nb = tnb; k_s1b = t1b; k_s2b = t2b;
            k_valid = tv;
            k_exact = t_exact;
        }
    }
    const int p = p0 + lane;
    if (p < a.P) {
        double pv = NAN;
        uint32_t st = 0;
        if (k_valid) {
            double t, df;
            if (k_exact) ttest_from_exact_sums(k_na, k_s1a, k_s2a, k_nb, k_s1b, k_s2b, t, df);
            else ttest_from_moments(k_na, k_s1a, k_s2a, k_nb, k_s1b, k_s2b, t, df);
            pv = student_t_two_sided(t, df);
            st |= STATUS_TTEST_VALID;
        }
        if (pv != pv || pv < a.threshold) st |= STATUS_SELECTED;  // diffmod.py:54
        // a p-value this close to the threshold could fall on the other side of it under another (equally accurate)
        // evaluation of the incomplete beta function: flagged for the host (bit-exact selection)
        if (fabs(pv - a.threshold) <= 1e-9 * a.threshold) st |= STATUS_NEAR_THRESHOLD;
        a.pval[p] = pv;
        a.status[p] = st;
    }
}

// ------------------------------------------------------------------------------------------
// K1 for u16 batches whose group slots are blocked by condition (what every batch builder emits): exact integer
// sums over ALIGNED 16-byte chunks, no masks and no shuffles in the streaming loop.
// ------------------------------------------------------------------------------------------
// Integer sums do not care about order or grouping, so a quad of lanes simply strides over the aligned 16-byte
// chunks [c0, cE] that cover its position - lane sl takes c0 + sl, c0 + sl + 4, ... - and adds all eight codes of
// every chunk: per pair of codes two extracts, one three-input add and two 32 x 32 + 64-bit multiply-adds.  The
// reads of the condition whose block of slots comes first are a prefix of the position (`split` reads), so a lane
// keeps, besides its running total, a copy of the total as it stood after its last chunk in front of the chunk
// that holds the first read of the other condition (cS): the quad's sum of those copies is the prefix side, the
// other side follows by subtraction.  What the whole chunks added too much - the neighbours' codes in the first
// and the last chunk, and the prefix part of chunk cS that went to the wrong side - is computed once per position
// by three lanes of the quad (one masked chunk each) and taken out again: the arithmetic is exact, so the result
// equals the sums of the position's own reads bit for bit, and the p-values those of xp_prefilter_kernel<W, true>.
// The per-position bookkeeping (offsets, split, conditions present) is done once per tile by the lane that also
// evaluates the position's p-value, and handed to the quads by shuffles.
// The kernel returns at once when the slots are not blocked by condition; xp_prefilter_kernel<W, true> is
// launched behind it and returns at once when they are (both read the same grp_cond): the host-side entry points
// take device pointers and never synchronise, so the choice is made on the device.
__device__ __forceinline__ unsigned long long mad_wide_u32(unsigned a, unsigned b, unsigned long long c) {
    unsigned long long d;
    asm("mad.wide.u32 %0, %1, %2, %3;" : "=l"(d) : "r"(a), "r"(b), "l"(c));
    return d;
}
// Sums of a chunk's eight codes q = 256 h + l: s1 += sum q, ql += sum q l, qh += sum q h (sum q^2 = ql + 256 qh).
// One byte permute per pair of codes gathers {l0, l1, h0, h1}; three two-way 16 x 8-bit dot products (IDP.2A) do
// the rest.  32-bit accumulators: a term is < 2^24, so 256 codes (32 chunks) fit before they are folded into 64 bits.
__device__ __forceinline__ void pf_add_chunk(const uint4& w, unsigned& s1, unsigned& ql, unsigned& qh) {
    const unsigned x[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const unsigned b = __byte_perm(x[j], 0u, 0x3120);
        s1 = __dp2a_lo(x[j], 0x0101u, s1);
        ql = __dp2a_lo(x[j], b, ql);
        qh = __dp2a_hi(x[j], b, qh);
    }
}
constexpr int kPfFoldTrips = 32;
// codes k in [lo, hi) of a chunk
__device__ __forceinline__ void pf_add_chunk_range(const uint4& w, int lo, int hi, unsigned& s1, unsigned long long& s2) {
    const unsigned x[4] = {w.x, w.y, w.z, w.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
        const unsigned q = (k >= lo && k < hi) ? ((x[k >> 1] >> ((k & 1) * 16)) & 0xffffu) : 0u;
        s1 += q;
        s2 = mad_wide_u32(q, q, s2);
    }
}

#ifndef XP_PF16_MINB
#define XP_PF16_MINB 8
#endif
template <int WARPS>
__global__ void __launch_bounds__(WARPS * 32, XP_PF16_MINB) xp_prefilter_u16_kernel(PrefilterArgs a) {
    __shared__ int s_cond[kMaxGroups];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int G = a.G;
    for (int g = threadIdx.x; g < G; g += blockDim.x) s_cond[g] = a.grp_cond[g];
    __syncthreads();
    if (!pf_cond_sorted(s_cond, G)) return;  // xp_prefilter_kernel<W, true> takes the batch
    const int p0 = (blockIdx.x * WARPS + warp) * 32;
    if (p0 >= a.P) return;

    // ---- once per tile: lane j looks at position p0 + j ---------------------------------------------------
    const int p = p0 + lane;
    int64_t off = 0;
    int n = 0, split = 0;
    bool valid = false, prefix_a = true;
    if (p < a.P) {
        off = a.read_off[p];
        n = (int)(a.read_off[p + 1] - off);
        const int32_t* glen = a.grp_len + (size_t)p * G;
        unsigned cmask = 0;
        int cfirst = -1;
        for (int g = 0; g < G; ++g) {
            const int len = glen[g], c = s_cond[g];
            if (len > 0) {
                cmask |= 1u << c;
                if (cfirst < 0) cfirst = c;
            }
            if (c == cfirst) split += len;  // reads of the condition whose slots come first: a prefix of the position
        }
        valid = a.enabled && (__popc(cmask) == 2);
        prefix_a = (cfirst == __ffs(cmask) - 1);  // the lower-numbered condition is side a (as in xp_prefilter_kernel)
    }
    // offsets relative to the aligned chunk that holds the tile's first read (a tile has < 2^20 reads)
    const int64_t off0 = __shfl_sync(kFull, off, 0);
    const int64_t e0 = off0 & ~(int64_t)7;
    const int rel = (int)(off - e0);
    const int n_eff = valid ? n : 0;
    const uint4* src = reinterpret_cast<const uint4*>(reinterpret_cast<const unsigned short*>(a.y) + e0);

    const int sub = lane >> 2, sl = lane & 3;
    unsigned k_t1 = 0u, k_a1 = 0u;
    unsigned long long k_t2 = 0ull, k_a2 = 0ull;
#pragma unroll 1
    for (int round = 0; round < 4; ++round) {
        const int from = round * 8 + sub;
        const int r0 = __shfl_sync(kFull, rel, from), np = __shfl_sync(kFull, n_eff, from);
        const int sp = __shfl_sync(kFull, split, from);
        const int c0 = r0 >> 3, cS = (r0 + sp) >> 3;
        const int cE = np > 0 ? (r0 + np - 1) >> 3 : c0 - 1;
        unsigned t1 = 0u, a1 = 0u;
        unsigned long long t2 = 0ull, a2 = 0ull;
        const uint4 zero4 = make_uint4(0u, 0u, 0u, 0u);
        int c = c0 + sl;
        uint4 cur = (c <= cE) ? __ldg(src + c) : zero4;
        uint4 nxt = (c + 4 <= cE) ? __ldg(src + c + 4) : zero4;
        while (c <= cE) {  // blocks of kPfFoldTrips chunks (one block unless the position has > 1024 reads)
            unsigned s1 = 0u, ql = 0u, qh = 0u, p1 = 0u, pl = 0u, ph = 0u;
            const bool in_prefix = c < cS;
            const int cend = min(cE, c + 4 * (kPfFoldTrips - 1));
#pragma unroll 3
            for (; c <= cend; c += 4) {
                const uint4 nn = (c + 8 <= cE) ? __ldg(src + c + 8) : zero4;
                pf_add_chunk(cur, s1, ql, qh);
                if (c < cS) {  // still in front of the other condition's first read: the prefix follows the total
                    p1 = s1;
                    pl = ql;
                    ph = qh;
                }
                cur = nxt;
                nxt = nn;
            }
            if (in_prefix) {
                a1 = t1 + p1;
                a2 = t2 + pl + ((unsigned long long)ph << 8);
            }
            t1 += s1;
            t2 += ql + ((unsigned long long)qh << 8);
        }
        // ---- what the whole chunks added too much: lane 0 the codes in front of the position in chunk c0, lane 1
        // those behind it in chunk cE, lane 2 the prefix part of chunk cS (counted for the other side above) ----
        {
            const int first = r0 & 7, last = (r0 + np - 1) & 7;  // of the position's reads inside c0 / cE
            const int cc = sl == 0 ? c0 : (sl == 1 ? cE : cS);
            int lo = 0, hi = 0;
            if (np > 0) {
                if (sl == 0) hi = first;
                else if (sl == 1) { lo = last + 1; hi = 8; }
                else if (sl == 2) { lo = cS == c0 ? first : 0; hi = (r0 + sp) & 7; }
            }
            unsigned x1 = 0u;
            unsigned long long x2 = 0ull;
            if (hi > lo) {
                const uint4 w = __ldg(src + cc);
                pf_add_chunk_range(w, lo, hi, x1, x2);
            }
            if (sl < 2) {
                t1 -= x1;
                t2 -= x2;
                if (sl == 0 && c0 < cS) {
                    a1 -= x1;
                    a2 -= x2;
                }
            } else {
                a1 += x1;
                a2 += x2;
            }
        }
#pragma unroll
        for (int d = 1; d < 4; d <<= 1) {
            t1 += __shfl_xor_sync(kFull, t1, d);
            a1 += __shfl_xor_sync(kFull, a1, d);
            t2 += __shfl_xor_sync(kFull, t2, d);
            a2 += __shfl_xor_sync(kFull, a2, d);
        }
        // position round * 8 + j goes to lane round * 8 + j
        const int own = (lane & 7) * 4;
        const unsigned g_t1 = __shfl_sync(kFull, t1, own), g_a1 = __shfl_sync(kFull, a1, own);
        const unsigned long long g_t2 = __shfl_sync(kFull, t2, own), g_a2 = __shfl_sync(kFull, a2, own);
        if ((lane >> 3) == round) {
            k_t1 = g_t1; k_a1 = g_a1; k_t2 = g_t2; k_a2 = g_a2;
        }
    }
    if (p < a.P) {
        double pv = NAN;
        uint32_t st = 0;
        if (valid) {
            const unsigned s1a = prefix_a ? k_a1 : k_t1 - k_a1, s1b = prefix_a ? k_t1 - k_a1 : k_a1;
            const unsigned long long s2a = prefix_a ? k_a2 : k_t2 - k_a2, s2b = prefix_a ? k_t2 - k_a2 : k_a2;
            const int na = prefix_a ? split : n - split, nb = n - na;
            // n S2 - S1^2 is exact in 64 bits (< 2^62): the t statistic takes the exact numerators
            const double nssa = (double)((unsigned long long)na * s2a - (unsigned long long)s1a * s1a);
            const double nssb = (double)((unsigned long long)nb * s2b - (unsigned long long)s1b * s1b);
            double t, df;
            ttest_from_exact_sums((double)na, (double)s1a, nssa, (double)nb, (double)s1b, nssb, t, df);
            pv = student_t_two_sided(t, df);
            st |= STATUS_TTEST_VALID;
        }
        if (pv != pv || pv < a.threshold) st |= STATUS_SELECTED;  // diffmod.py:54
        if (fabs(pv - a.threshold) <= 1e-9 * a.threshold) st |= STATUS_NEAR_THRESHOLD;  // see xp_prefilter_kernel
        a.pval[p] = pv;
        a.status[p] = st;
    }
}

// ------------------------------------------------------------------------------------------
// Worklist: selected positions ordered by descending read count (semi-log buckets), so the
// persistent fit kernel starts the longest positions first.
// ------------------------------------------------------------------------------------------
constexpr int kCostBuckets = 512;  // semi-log read-count classes (16 per octave)
__device__ __forceinline__ int cost_class(int n) {
    int cls;
    if (n < 16) cls = n < 0 ? 0 : n;
    else {
        const int e = 31 - __clz(n);
        cls = 16 + (e - 4) * 16 + ((n >> (e - 4)) & 15);
    }
    return min(cls, kCostBuckets - 1);
}
__device__ __forceinline__ int cost_bucket(int n) { return kCostBuckets - 1 - cost_class(n); }  // larger n -> smaller bucket id
// Bucket of a position in the worklist.  Positions of the team / streaming classes (n >= split > 0) are ordered by read
// count alone, deepest first: buckets [0, 512), the class boundaries are bucket boundaries.  The single-warp class follows
// in [512, 1024), ordered by how long a position is expected to sweep, then by read count: VB-EM needs the more sweeps
// the less the two conditions differ (c2: 66 sweeps on average, up to 270, where the prefilter's p >= 1e-2; 34, up to 92,
// where p < 1e-10), so the prefilter's p-value sorts the long runners to the front and the persistent kernel's tail is made
// of short positions (longest-processing-time-first scheduling; the order changes no result).
__device__ __forceinline__ int worklist_bucket(int n, int split, double pval) {
    if (split > 0 && n >= split) return cost_bucket(n);
    const int pc = !(pval < 1e-2) ? 0 : !(pval < 1e-5) ? 1 : !(pval < 1e-10) ? 2 : 3;  // NaN (no test): class 0
    return kCostBuckets + pc * 128 + (127 - min(cost_class(n), 127));
}

struct WorklistArgs {
    const uint32_t* status;  // may be null => every position selected
    const int64_t* read_off;
    int P;
    uint32_t* hist;      // [kBuckets] zeroed
    uint32_t* cursor;    // [kBuckets]
    uint32_t* n_work;    // [1]
    int32_t* worklist;   // [P]
    uint32_t* n_large;   // [1] entries of the worklist with at least `split` reads (they come first)
    int split;           // a bucket boundary (see cost_bucket), or 0: no large class
    uint32_t* n_deep;    // [1] entries with at least `split_deep` reads (the first ones of the worklist)
    int split_deep;      // a bucket boundary above `split`, or 0: no streaming class
    const double* pval;  // prefilter p-values (null: none), see worklist_bucket
};

__global__ void xp_worklist_count(WorklistArgs a) {
    __shared__ uint32_t h[kBuckets];
    for (int i = threadIdx.x; i < kBuckets; i += blockDim.x) h[i] = 0;
    __syncthreads();
    for (int p = blockIdx.x * blockDim.x + threadIdx.x; p < a.P; p += gridDim.x * blockDim.x) {
        if (a.status && !(a.status[p] & STATUS_SELECTED)) continue;
        atomicAdd(&h[worklist_bucket((int)(a.read_off[p + 1] - a.read_off[p]), a.split, a.pval ? a.pval[p] : NAN)], 1u);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kBuckets; i += blockDim.x)
        if (h[i]) atomicAdd(&a.hist[i], h[i]);
}

__global__ void xp_worklist_scan(WorklistArgs a) {  // <<<1, kBuckets>>>
    __shared__ uint32_t s[kBuckets];
    const int t = threadIdx.x;
    s[t] = a.hist[t];
    __syncthreads();
    for (int d = 1; d < kBuckets; d <<= 1) {
        const uint32_t v = (t >= d) ? s[t - d] : 0;
        __syncthreads();
        s[t] += v;
        __syncthreads();
    }
    a.cursor[t] = s[t] - a.hist[t];  // exclusive start
    if (t == kBuckets - 1) *a.n_work = s[t];
    // buckets 0..cost_bucket(x) hold the positions with n >= x when x is the first count of its bucket
    const uint32_t nl = a.split > 0 ? s[cost_bucket(a.split)] : 0u;
    if (t == 0) {
        *a.n_large = nl;
        *a.n_deep = a.split_deep > 0 ? s[cost_bucket(a.split_deep)] : 0u;
    }
}

// Warp-aggregated: the lanes of a warp that fall into the same bucket take their slots with one atomic (a design
// with similar coverage everywhere - pooled c3, c5 - puts a whole batch into a handful of buckets, and a million
// single atomics on three addresses ran 0.2 - 0.26 ms at 0.5 % issue utilisation).
__global__ void xp_worklist_scatter(WorklistArgs a) {
    const int lane = threadIdx.x & 31;
    const int stride = gridDim.x * blockDim.x;
    for (int p0 = blockIdx.x * blockDim.x + (threadIdx.x & ~31); p0 < a.P; p0 += stride) {  // warp-uniform trip count
        const int p = p0 + lane;
        const bool sel = p < a.P && !(a.status && !(a.status[p] & STATUS_SELECTED));
        const int b = sel ? worklist_bucket((int)(a.read_off[p + 1] - a.read_off[p]), a.split, a.pval ? a.pval[p] : NAN)
                          : -1 - lane;  // unselected: a group of one
        const unsigned peers = __match_any_sync(kFull, b);
        const int leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (sel && lane == leader) base = atomicAdd(&a.cursor[b], (uint32_t)__popc(peers));
        base = __shfl_sync(kFull, base, leader);
        if (sel) a.worklist[base + __popc(peers & ((1u << lane) - 1u))] = p;
    }
}

}  // namespace xpk
#include "xp_fit.cuh"  // K3: variational EM
#include "xp_fit_tm.cuh"  // K3 for u16 batches: reads resident in tensor memory, two positions per warp
namespace xpk {

// ------------------------------------------------------------------------------------------
// K4: statistics, one thread per fitted position.
// ------------------------------------------------------------------------------------------
struct StatsArgs {
    const int32_t* worklist;  // null => identity over P
    const uint32_t* n_work;   // null => P
    int P, G, C;
    const int32_t* grp_cond;
    const int32_t* grp_len;
    const double* kmer_mean;
    const double* kmer_std;
    double conc0;
    const double* fit;
    double* stats;
    uint32_t* status;
};

__global__ void xp_stats_kernel(StatsArgs a) {
    __shared__ int s_cond[kMaxGroups];
    for (int g = threadIdx.x; g < a.G; g += blockDim.x) s_cond[g] = a.grp_cond[g];
    __syncthreads();
    const unsigned n = a.n_work ? *a.n_work : (unsigned)a.P;
    const unsigned w = blockIdx.x * blockDim.x + threadIdx.x;
    if (w >= n) return;
    const int p = a.worklist ? a.worklist[w] : (int)w;
    double w_mod[kMaxGroups], cov[kMaxGroups];
    const uint32_t fl = compute_stats(a.G, a.C, s_cond, a.grp_len + (size_t)p * a.G, a.conc0, a.kmer_mean[p],
                                      a.kmer_std[p], a.fit + (size_t)p * fit_width(a.G),
                                      a.stats + (size_t)p * stats_width(a.G, a.C), w_mod, cov);
    a.status[p] = (a.status[p] & ~(STATUS_KEEP_ROW | STATUS_MOD_IS_1 | STATUS_HIGHER)) | fl;
}

// ------------------------------------------------------------------------------------------
// Compact fixed-width records of the fitted positions (worklist order), one warp per record:
//   [global position index, status, n_iter, pval, fit[FW], stats[SW]]  as f64
// This is what the final NCCL gather ships to the writer rank.
// ------------------------------------------------------------------------------------------
struct PackArgs {
    const int32_t* worklist;
    const uint32_t* n_work;
    const uint32_t* status;
    const int32_t* n_iter;
    const double* pval;
    const double* fit;
    const double* stats;
    int FW, SW;
    long long index_offset, capacity;
    double* out;
};

__global__ void xp_pack_records_kernel(PackArgs a) {
    const unsigned n = *a.n_work;
    const int lane = threadIdx.x & 31;
    const int W = 4 + a.FW + a.SW;
    for (unsigned w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); w < n && (long long)w < a.capacity;
         w += gridDim.x * (blockDim.x >> 5)) {
        const int p = a.worklist[w];
        double* o = a.out + (size_t)w * W;
        if (lane == 0) {
            o[0] = (double)(a.index_offset + p);
            o[1] = (double)a.status[p];
            o[2] = (double)a.n_iter[p];
            o[3] = a.pval[p];
        }
        const double* f = a.fit + (size_t)p * a.FW;
        const double* t = a.stats + (size_t)p * a.SW;
        for (int c = lane; c < a.FW; c += 32) o[4 + c] = f[c];
        for (int c = lane; c < a.SW; c += 32) o[4 + a.FW + c] = t[c];
    }
}

// Rows of diffmod.table only: records (same layout) of the positions [0, P) whose status carries every bit of
// `need` (SELECTED | KEEP_ROW: fitted and past the row filter of io.py:363-367), compacted in position
// order.  Two small launches: every warp owns a contiguous span of positions and counts its rows
// (xp_count_rows_kernel); then it adds up the counts of the warps before it and writes its rows in order
// (xp_pack_rows_kernel).  Deterministic, no atomics; the last warp leaves the total in n_rows.
struct PackRowsArgs {
    const uint32_t* status;
    const int32_t* n_iter;
    const double* pval;
    const double* fit;
    const double* stats;
    int P, FW, SW;
    int span;              // positions per warp, a multiple of 32
    int n_warps;           // warps of the launch (gridDim.x * blockDim.x / 32)
    uint32_t need;
    long long index_offset, capacity;
    unsigned int* counts;  // [n_warps]
    unsigned int* n_rows;
    double* out;
};

__global__ void xp_count_rows_kernel(PackRowsArgs a) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= a.n_warps) return;
    const int lo = w * a.span, hi = min(a.P, lo + a.span);
    unsigned c = 0;
    for (int p = lo + lane; p < hi; p += 32) c += ((a.status[p] & a.need) == a.need) ? 1u : 0u;
    c = __reduce_add_sync(kFull, c);
    if (lane == 0) a.counts[w] = c;
}

__global__ void xp_pack_rows_kernel(PackRowsArgs a) {
    const int lane = threadIdx.x & 31;
    const int w = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (w >= a.n_warps) return;
    const int W = 4 + a.FW + a.SW;
    unsigned before = 0;
    for (int j = lane; j < w; j += 32) before += a.counts[j];
    long long row = __reduce_add_sync(kFull, before);
    if (w == a.n_warps - 1 && lane == 0) *a.n_rows = (unsigned)row + a.counts[w];
    const int lo = w * a.span, hi = min(a.P, lo + a.span);
    for (int p0 = lo; p0 < hi; p0 += 32) {
        const int p = p0 + lane;
        const bool keep = p < hi && (a.status[p] & a.need) == a.need;
        for (unsigned rest = __ballot_sync(kFull, keep); rest; rest &= rest - 1, ++row) {
            if (row >= a.capacity) return;
            const int q = p0 + __ffs(rest) - 1;
            double* o = a.out + (size_t)row * W;
            if (lane == 0) {
                o[0] = (double)(a.index_offset + q);
                o[1] = (double)a.status[q];
                o[2] = (double)a.n_iter[q];
                o[3] = a.pval[q];
            }
            const double* f = a.fit + (size_t)q * a.FW;
            const double* t = a.stats + (size_t)q * a.SW;
            for (int c = lane; c < a.FW; c += 32) o[4 + c] = f[c];
            for (int c = lane; c < a.SW; c += 32) o[4 + a.FW + c] = t[c];
        }
    }
}

// Counts of an asynchronous call in the chained form of the host pipeline: {rows needed, positions fitted}; all ones
// in the first when the chained launch gave up waiting for a chunk (the caller cannot be told any other way).
__global__ void xp_chain_counts_kernel(const unsigned int* n_rows, const unsigned int* n_fitted, int n_chunks,
                                       const unsigned int* abort_flag, unsigned int* out_counts) {
    if (threadIdx.x == 0 && blockIdx.x == 0) {
        unsigned int fitted = 0;
        for (int c = 0; c < n_chunks; ++c) fitted += n_fitted[c];
        out_counts[0] = *abort_flag ? 0xffffffffu : n_rows[0];
        out_counts[1] = fitted;
    }
}

// Rows of several chunks (xpb200_fit_host_rows_async: chunk c's rows sit at row chunk_start[c] of `rows`, n_rows[c] of
// them) concatenated in chunk order into `out`; one warp per row.  out_counts = {rows needed, positions fitted}.
constexpr int kMaxRowChunks = 64;
struct ConcatRowsArgs {
    const double* rows;
    const unsigned int* n_rows;    // [n_chunks]
    const unsigned int* n_fitted;  // [n_chunks]
    int n_chunks, RW;
    long long capacity;
    int chunk_start[kMaxRowChunks];
    double* out;
    unsigned int* out_counts;  // [2]
};

__global__ void xp_concat_rows_kernel(ConcatRowsArgs a) {
    __shared__ unsigned int s_first[kMaxRowChunks + 1];  // first output row of every chunk
    if (threadIdx.x == 0) {
        unsigned int acc = 0, fitted = 0;
        for (int c = 0; c < a.n_chunks; ++c) {
            s_first[c] = acc;
            acc += a.n_rows[c];
            fitted += a.n_fitted[c];
        }
        s_first[a.n_chunks] = acc;
        if (blockIdx.x == 0) {
            a.out_counts[0] = acc;
            a.out_counts[1] = fitted;
        }
    }
    __syncthreads();
    const int lane = threadIdx.x & 31, wpb = blockDim.x >> 5;
    const long long total = min((long long)s_first[a.n_chunks], a.capacity);
    for (long long row = (long long)blockIdx.x * wpb + (threadIdx.x >> 5); row < total; row += (long long)gridDim.x * wpb) {
        int c = 0;
        while (row >= (long long)s_first[c + 1]) ++c;
        const double* src = a.rows + ((size_t)a.chunk_start[c] + (size_t)(row - s_first[c])) * a.RW;
        double* dst = a.out + (size_t)row * a.RW;
        for (int k = lane; k < a.RW; k += 32) dst[k] = src[k];
    }
}

// ------------------------------------------------------------------------------------------
// z node for --save_models (io.py:99-139): responsibilities of the last E-step, one warp per position.
// ------------------------------------------------------------------------------------------
struct PosteriorArgs {
    const void* y;
    int y_format;
    double y_scale;
    const int64_t* read_off;
    const int32_t* grp_len;
    const double* kmer_mean;
    const uint32_t* status;
    const int32_t* n_iter;
    const double* coef;  // [P, G+2]
    int P, G;
    double* prob0;    // [R]
    double* ln_prob;  // [R, 2]
};

__global__ void xp_posterior_kernel(PosteriorArgs a) {
    const int lane = threadIdx.x & 31;
    const int wpb = blockDim.x >> 5;
    const double rscale = 1.0 / a.y_scale;
    for (int p = blockIdx.x * wpb + (threadIdx.x >> 5); p < a.P; p += gridDim.x * wpb) {
        if (!(a.status[p] & STATUS_SELECTED) || a.n_iter[p] < 1) continue;
        const int64_t off = a.read_off[p];
        const int n = (int)(a.read_off[p + 1] - off);
        const double m0 = a.kmer_mean[p];
        const double* cf = a.coef + (size_t)p * (a.G + 2);
        const double B = cf[a.G], C = cf[a.G + 1];
        const int32_t* glen = a.grp_len + (size_t)p * a.G;
        for (int i = lane; i < n; i += 32) {
            int g = 0, gend = glen[0];
            while (i >= gend) gend += glen[++g];
            const double yc = y_load_global(a.y, a.y_format, a.y_scale, rscale, off + i) - m0;
            const double d = fma(yc, fma(C, yc, B), cf[g]);
            double p0, l0, l1;
            posterior_read(HostTab(), d, p0, l0, l1);
            a.prob0[off + i] = p0;
            a.ln_prob[2 * (off + i)] = l0;
            a.ln_prob[2 * (off + i) + 1] = l1;
        }
    }
}

// ------------------------------------------------------------------------------------------
// FP64 peak: 8 independent DFMA chains per thread, no memory traffic.
// ------------------------------------------------------------------------------------------
__global__ void xp_dfma_peak_kernel(double* out, int iters, double seed) {
    double x0 = seed + threadIdx.x, x1 = x0 + 1, x2 = x0 + 2, x3 = x0 + 3, x4 = x0 + 4, x5 = x0 + 5, x6 = x0 + 6,
           x7 = x0 + 7;
    const double m = 0.999999, c = 1e-9;
    for (int i = 0; i < iters; ++i) {
        x0 = fma(x0, m, c); x1 = fma(x1, m, c); x2 = fma(x2, m, c); x3 = fma(x3, m, c);
        x4 = fma(x4, m, c); x5 = fma(x5, m, c); x6 = fma(x6, m, c); x7 = fma(x7, m, c);
    }
    const double s = ((x0 + x1) + (x2 + x3)) + ((x4 + x5) + (x6 + x7));
    if (s == 123.456) out[0] = s;  // never true; keeps the chains alive
}

// Special functions as the kernels evaluate them (accuracy tests).
__global__ void xp_eval_special_kernel(int which, const double* x, const double* aux, double* out, long long n) {
    const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
    const bool valid = i < n;
    const double v = valid ? x[i] : 0.0;
    double r = NAN, ps, lg;
    if (which == 5 || which == 6) {
        double inv, rmin, h;
        bool pos;
        estep_read(HostTab(), v, inv, rmin, h, pos);
        r = (which == 5) ? (pos ? inv : rmin) : -h;  // r_n1 ; r0 ln r0 + r1 ln r1
    } else if (valid) {
        switch (which) {
            case 0: psi_lgamma(v, &ps, &lg); r = ps; break;
            case 1: psi_lgamma(v, &ps, &lg); r = lg; break;
            case 2: r = xlog(v); break;
            case 3: r = xexp(v); break;
            case 4: r = student_t_two_sided(v, aux[i]); break;
        }
    }
    if (valid) out[i] = r;
}

}  // namespace xpk
