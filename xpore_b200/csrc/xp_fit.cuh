// xp_fit.cuh — K3: variational EM of the two-component mixture (gmm.py:11-127) on sm_100a.
//
// Work unit = one position.  A *team* of TW warps owns one position at a time (TW = 1: one warp per
// position, the common case; TW = 4: high-coverage positions whose reads would not leave room for
// enough resident warps otherwise).  A CTA holds several teams; every team pulls tickets from a
// global worklist (longest positions first), so ragged convergence (2...300+ sweeps) never idles an SM.
//
// Per position:
//   1. one TMA bulk copy (cp.async.bulk + mbarrier) lands the position's reads in the team's slab;
//      they are decoded / centred on the k-mer mean in place and stay resident for all sweeps;
//   2. initial locations (np.percentile, gmm.py:39-44) by an exact radix select: on the integer codes of a u16
//      batch before the decode (two 8-bit digits), on order-preserving 64-bit keys otherwise;
//   3. sweeps.  The team leader (warp 0 of the team) runs the O(G) scalar phase — psi / lgamma of the
//      2G+2 posterior parameters spread over lanes, ELBO in closed form from sufficient statistics, stopping
//      rule (gmm.py:110-123), M-step — keeping component k in the lanes of parity k; all warps of the team
//      run the E-step over their share of the resident reads (two reads per lane per trip, four in 4-warp teams).
//
// E-step bookkeeping: only component 0 is accumulated per read (N_g0, S1_0, S2_0 and the entropy);
// component 1 follows from the per-position totals because r_n1 = 1 - r_n0 (gmm.py:242):
//   N_g1 = n_g - N_g0,  S1_1 = sum y' - S1_0,  S2_1 = sum y'^2 - S2_0.
// The 256-entry 2^(j/256) table and the 129-entry {1/c, ln c} table live in shared memory (4 KB per CTA).
//
// CTA shared memory:  [ exp2 table 256 f64 | {1/c, ln c} table 129 double2 (+pad) | team 0 | team 1 | ... ]
// team region, in this order (doubles unless noted).  The fixed-size arrays and the scratch columns come first:
// their addresses are the team base plus a constant; the offsets of the rest are kernel arguments (fit_fill_layout).
//   sc[8]              psi(alpha_k), lgamma(alpha_k), ln beta_k, ln lambda_k published by their lanes
//   bc[8]              leader -> team broadcast: B, C, continue flag, ticket, reduction slots
//   pc[8]              per-position values the leader needs once per sweep (m0, a0, b0, the two ELBO constants,
//                      sum y', sum y'^2, previous ELBO): parked here so they hold no registers during the E-step
//   mbar (u64) + pad
//   scratch[(G+3) x (32 TW + 1)]  column c holds one double per thread of the team: its partial N_g0 of
//                      group c < G, then S1_0, S2_0 and the entropy of the current sweep (odd column pitch: the
//                      column sums read conflict free, and a thread's entry is scratch + tid + c * pitch);
//                      at least 2 KB: the two 256-bin select histograms live here before the first sweep
//   tot[G+4]           column totals
//   Asm[G]             per-group constant A_g of d_n, as used by the E-step of the current sweep
//   dps[G]             psi(c_g1) - psi(c_g0) of the scalar phase in progress (becomes Asm only if another sweep follows,
//                      so Asm / bc still describe the last E-step when the position is done: --save_models)
//   goff[G+1] (int)    group offsets inside the position
//   slab[cap+2]        16-byte aligned: centred reads y' (TMA destination; +2 = alignment slack)
//   gid[cap+2] (u8)    group slot of every read
#pragma once
#include <cuda_runtime.h>

#ifndef XP_FIT_NR
#define XP_FIT_NR 2
#endif
#ifndef XP_FIT_NR_TEAM  // reads per lane and trip of the 4-warp teams (16 warps per SM: more chains per warp)
#define XP_FIT_NR_TEAM 4
#endif
#ifndef XP_FIT_FLOOR
#define XP_FIT_FLOOR 0
#endif
// per-position values the leader needs once per sweep are parked in shared memory (pc[]): XP_PK(i, name) re-reads slot i
#define XP_PK(i, var) parked(pc + (i))

namespace xpk {

struct FitArgs {
    const void* y;
    int y_format;
    double y_scale;
    const int64_t* read_off;
    const int32_t* grp_len;
    const double* kmer_mean;
    const double* kmer_std;
    const int32_t* worklist;   // null => identity
    const uint32_t* w_begin;   // device: first worklist entry of this launch (null => 0)
    const uint32_t* w_end;     // device: one past the last worklist entry of this launch
    unsigned int* ticket;      // zeroed before launch
    int G, max_iters, cap;
    double stopping, conc0, lg_prior_w;  // lg_prior_w = lgamma(2 c0) - 2 lgamma(c0)
    double* fit;
    int32_t* n_iter;
    uint32_t* status;
    double* coef;  // optional [P, G+2]: A_g, B, C of the last E-step (for xp_posterior_kernel)
    // byte offsets inside a team's shared-memory region and its size, filled by fit_fill_layout: the kernel adds
    // them to the team base as constant-bank operands instead of re-deriving them from G and cap
    int team_bytes, off_tot, off_asm, off_dps, off_goff, off_slab, off_gid;
    // xp_fit_tm.cuh (reads resident in tensor memory): bytes of one position's header, offset of the warp's scratch
    // columns inside the warp's region, tensor-memory columns per position
    int hdr_bytes, off_scratch, tm_cols;
};

// The chunked host call's chain of worklists (xp_fit_tm_kernel only; n = 0: a plain launch over FitArgs' worklist).
// ONE persistent launch serves the whole batch: a warp that finds chunk c's worklist exhausted goes on with chunk c + 1,
// waiting (nanosleep polling) until that chunk's reads have arrived and its prefilter and worklist launches - which run
// beside this kernel on the SMs its grid leaves free - have set ready[c + 1].  No chunk ends with a tail of idle warps,
// and the early, small chunks no longer leave most of the machine empty.  FitArgs' position arrays are the batch's
// (chunk c's worklist entries are relative to position p0 of the batch).
constexpr int kMaxChainChunks = 64;
struct FitChainChunk {
    const int32_t* worklist;
    unsigned int* ticket;
    const uint32_t* n_work;
    const uint32_t* w_begin;  // entries in front of the single-warp class (the chunk's team / streaming classes, which get launches of their own); null: none
    int p0, pad;
};
struct FitChain {
    int n;
    unsigned int spin_limit;  // polls (about a microsecond each) before a warp gives up and raises `abort`
    const uint32_t* ready;    // [n] device flags
    unsigned int* abort;      // device flag: set by the host on an error path or by a warp whose wait timed out
    unsigned int* wait_polls; // experiment builds: [n] polls spent waiting for each chunk, summed over the warps
    unsigned int* done;       // [n] positions of each chunk fitted so far (their records are visible device-wide): the host
                              // polls these and runs a chunk's statistics / row compaction / copy-out while later chunks are fitted
    FitChainChunk c[kMaxChainChunks];
};

constexpr int kTabExpDoubles = 256;
constexpr int kTabLogDoubles = 130 * 2;
constexpr size_t kFitTableBytes = (size_t)(kTabExpDoubles + kTabLogDoubles) * 8;

__host__ __device__ inline int fit_rows(int G) { return G + 3; }
// scratch: column c (N_g0 of group c < G, then S1_0, S2_0, entropy) holds one double per thread of the team;
// columns are 32 TW + 1 doubles apart (odd: the column sums read conflict free)
// (streaming flavour: one double per WARP of the team, columns TW + 1 doubles apart)
__host__ __device__ inline size_t fit_scratch_doubles(int G, int TW, bool stream = false) {
    const size_t s = (size_t)fit_rows(G) * (stream ? TW + 1 : 32 * TW + 1);
    return s > 256 ? s : 256;  // >= 2 KB: the two select histograms live here before the first sweep
}
// fixed-size arrays first (sc, bc, pc, mbar at constant offsets from the team base), then the scratch columns
// (constant offset too: the E-step's stores into them need no G- or cap-dependent address arithmetic), then the
// G-sized arrays, then - 16-byte aligned - the slab
constexpr int kFitFixedDoubles = 8 + 8 + 8 + 2;
__host__ __device__ inline size_t fit_team_header_bytes(int G, int TW, bool stream = false) {
    const size_t bytes = ((size_t)kFitFixedDoubles + fit_scratch_doubles(G, TW, stream) + (G + 4) + 2 * G) * 8 + (size_t)((G + 2) & ~1) * 4;
    return (bytes + 15) & ~(size_t)15;
}
// a streaming team keeps no reads in shared memory: its region is the header alone
__host__ __device__ inline size_t fit_team_smem_bytes(int cap, int G, int TW, bool stream = false) {
    if (stream) return fit_team_header_bytes(G, TW, true);
    const size_t bytes = fit_team_header_bytes(G, TW) + (size_t)(cap + 2) * 8 + (size_t)(cap + 2);
    return (bytes + 15) & ~(size_t)15;
}

inline void fit_fill_layout(FitArgs& a, int TW, bool stream = false) {
    const int G = a.G;
    a.team_bytes = (int)fit_team_smem_bytes(a.cap, G, TW, stream);
    a.off_tot = (int)((kFitFixedDoubles + fit_scratch_doubles(G, TW, stream)) * 8);
    a.off_asm = a.off_tot + (G + 4) * 8;
    a.off_dps = a.off_asm + G * 8;
    a.off_goff = a.off_dps + G * 8;
    a.off_slab = (int)fit_team_header_bytes(G, TW, stream);
    a.off_gid = a.off_slab + (a.cap + 2) * 8;
}

// shared-memory tables, addressed through the 32-bit shared window.  ES / LS: bytes between consecutive entries
// (8 / 16: one copy of each table; 128 / 128: the lane-replicated, bank-conflict-free copies of xp_fit_tm.cuh, where
// e2 and lg already carry the lane's offset inside an entry's row)
__device__ __forceinline__ void lds_tab2(uint32_t a, double& x, double& y);
template <unsigned ES, unsigned LS>
struct SmemTabT {
    static constexpr unsigned kES = ES, kLS = LS;
    uint32_t e2;  // 2^(j/256), entry j at e2 + ES * j
    uint32_t lg;  // {1/c, ln c}, biased so that entry i of u = 1 + t in [1, 2] is found at lg + LS * (hi32(u) >> 13)
    __device__ __forceinline__ void logc(int i, double& ic, double& lc) const {  // entry i (xlog_tab)
        lds_tab2(lg + ((0x3ff00000u >> 13) + (unsigned)i) * LS, ic, lc);
    }
};
using SmemTab = SmemTabT<8u, 16u>;
__device__ __forceinline__ double lds_tab(uint32_t a) {  // tables are immutable after the prologue
    double v;
    asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void lds_tab2(uint32_t a, double& x, double& y) {
    asm("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(x), "=d"(y) : "r"(a));
}
// a * b + c as one IMAD (written in C, the compiler turns a * 2^s into shift + mask + add)
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c) {
    uint32_t r;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(r) : "r"(a), "r"(b), "r"(c));
    return r;
}
__device__ __forceinline__ double lds_f64(uint32_t a) {  // mutable shared data: ordered like a volatile access
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a) : "memory");
    return v;
}
// a value parked in shared memory: re-read where it is used instead of living in a register pair
__device__ __forceinline__ double parked(const double* p) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(smem_u32(p)) : "memory");
    return v;
}
__device__ __forceinline__ void sts_f64(uint32_t a, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}

template <int TW>
__device__ __forceinline__ void team_sync(int team) {
    if (TW == 1) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(team + 1), "n"(TW * 32) : "memory");
}

// Order statistics sorted[r] and sorted[min(r+1, n-1)] of the n doubles in shared memory by an MSD
// radix select over the order-preserving 64-bit keys: 8-bit digits, 256-bin histogram (shared-memory
// integer atomics by the whole team), starting below the common prefix of the smallest and largest
// key and stopping as soon as one candidate is left.  Exact for any input.  Every warp of the team
// runs the (cheap) scan redundantly, so no broadcast is needed.  `yv(i)` yields read i: a shared-memory load for a
// resident position, a decoding global load for a streamed one.
template <int TW, class YF>
__device__ void team_select2(YF yv, int n, int r, uint64_t kmin, uint64_t kmax, int* hist, int lane,
                             int tid, int team, double& a, double& b) {
    const uint64_t diff = kmin ^ kmax;
    int hi = diff ? 64 - __clzll((long long)diff) : 0;  // bits [hi,64) are common to every key
    uint64_t pre = (hi >= 64) ? 0ull : (kmin >> hi);    // value of the fixed bits
    int rank = r, cnt = n;
    while (hi > 0 && cnt > 1) {
        const int lo = hi > 8 ? hi - 8 : 0;
        const unsigned mask = (1u << (hi - lo)) - 1u;
        for (int j = tid; j < 256; j += TW * 32) hist[j] = 0;
        team_sync<TW>(team);
        for (int i = tid; i < n; i += TW * 32) {
            const uint64_t k = dkey(yv(i));
            if (hi >= 64 || (k >> hi) == pre) atomicAdd(&hist[(unsigned)(k >> lo) & mask], 1);
        }
        team_sync<TW>(team);
        int h[8], s = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            h[j] = hist[lane * 8 + j];
            s += h[j];
        }
        int incl = s;
#pragma unroll
        for (int m = 1; m < 32; m <<= 1) {
            const int v = __shfl_up_sync(kFull, incl, m);
            if (lane >= m) incl += v;
        }
        const unsigned bal = __ballot_sync(kFull, incl > rank);
        const int L = __ffs(bal) - 1;  // lane whose bins hold the wanted rank
        int below = incl - s, d = 0, c = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (c == 0) {
                if (below + h[j] > rank) {
                    d = j;
                    c = h[j];
                } else below += h[j];
            }
        }
        d = __shfl_sync(kFull, lane * 8 + d, L);
        below = __shfl_sync(kFull, below, L);
        cnt = __shfl_sync(kFull, c, L);
        rank -= below;
        pre = (hi >= 64) ? (uint64_t)d : ((pre << (hi - lo)) | (uint64_t)d);
        hi = lo;
        team_sync<TW>(team);  // everybody has read the histogram before it is cleared again
    }
    uint64_t ka;
    if (hi == 0) ka = pre;
    else {  // one candidate left: fetch it
        uint64_t found = 0;
        for (int i = lane; i < n; i += 32) {
            const uint64_t k = dkey(yv(i));
            if ((k >> hi) == pre) found = k;
        }
        ka = warp_max_u64(found);
    }
    a = keyd(ka);
    b = a;
    if (r + 1 < n && !(hi == 0 && rank + 1 < cnt)) {
        uint64_t mn = ~0ull;
        for (int i = lane; i < n; i += 32) {
            const uint64_t k = dkey(yv(i));
            if (k > ka && k < mn) mn = k;
        }
        b = keyd(warp_min_u64(mn));
    }
}

// ---- order statistics of the u16 read format, taken on the integer codes before they are decoded -----------
// (the decode is strictly increasing, so rank r of the codes is rank r of the reads).  16-bit keys need at most
// two 8-bit digits, and the histogram of the high byte serves both percentiles of a position.
//
// Bin of a 256-bin histogram holding rank `rank`, that bin's count and the rank inside it (every lane gets them).
__device__ __forceinline__ void hist_find(const int* hist, int rank, int lane, int& bin, int& within, int& cnt) {
    int h[8], s = 0;
    {
        const int4 v0 = reinterpret_cast<const int4*>(hist)[lane * 2], v1 = reinterpret_cast<const int4*>(hist)[lane * 2 + 1];
        h[0] = v0.x, h[1] = v0.y, h[2] = v0.z, h[3] = v0.w, h[4] = v1.x, h[5] = v1.y, h[6] = v1.z, h[7] = v1.w;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) s += h[j];
    int incl = s;
#pragma unroll
    for (int m = 1; m < 32; m <<= 1) {
        const int v = __shfl_up_sync(kFull, incl, m);
        if (lane >= m) incl += v;
    }
    const int L = __ffs(__ballot_sync(kFull, incl > rank)) - 1;  // lane whose bins hold the wanted rank
    int below = incl - s, d = 0, c = 0;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        if (c == 0) {
            if (below + h[j] > rank) {
                d = j;
                c = h[j];
            } else below += h[j];
        }
    }
    bin = __shfl_sync(kFull, lane * 8 + d, L);
    within = rank - __shfl_sync(kFull, below, L);
    cnt = __shfl_sync(kFull, c, L);
}
// codes sorted[r] and sorted[min(r + 1, n - 1)]; hist1 = histogram of the high bytes (already built), hist2 = work space
template <int TW>
__device__ void team_select2_u16(const unsigned short* q, int n, int r, const int* hist1, int* hist2, int lane, int tid,
                                 int team, unsigned& qa, unsigned& qb) {
    constexpr int NT = TW * 32;
    int bh, wh, ch;
    hist_find(hist1, r, lane, bh, wh, ch);
    for (int j = tid; j < 256; j += NT) hist2[j] = 0;
    team_sync<TW>(team);
    for (int i = tid; i < n; i += NT) {
        const unsigned v = q[i];
        if ((int)(v >> 8) == bh) atomicAdd(&hist2[v & 255u], 1);
    }
    team_sync<TW>(team);
    int bl, wl, cl;
    hist_find(hist2, wh, lane, bl, wl, cl);
    qa = ((unsigned)bh << 8) | (unsigned)bl;
    qb = qa;
    if (r + 1 < n && wl + 1 >= cl) {  // the next order statistic is the smallest code above qa
        unsigned mn = 0xffffffffu;
        for (int i = lane; i < n; i += 32) {
            const unsigned v = q[i];
            if (v > qa && v < mn) mn = v;
        }
        qb = __reduce_min_sync(kFull, mn);
    }
    team_sync<TW>(team);  // everybody is done with hist2 before it is cleared again
}

// The two start locations' inputs from centred reads (gmm.py:39-44): returns loc1' = P75 or P25 of y', chosen by the
// reference's own comparison on the uncentred median (xp_core.cuh:prior_below_median).
template <int TW, class YF>
__device__ double team_second_location(YF yv, int n, double m0, uint64_t kmin, uint64_t kmax, int* hist,
                                       int lane, int tid, int team) {
    int lo, hi;
    double gamma;
    percentile_index(n, 0.5, lo, hi, gamma);
    double a, b;
    team_select2<TW>(yv, n, lo, kmin, kmax, hist, lane, tid, team, a, b);
    const bool up = prior_below_median(a + m0, b + m0, gamma, m0);
    percentile_index(n, up ? 0.75 : 0.25, lo, hi, gamma);
    team_select2<TW>(yv, n, lo, kmin, kmax, hist, lane, tid, team, a, b);
    return np_lerp(a, b, gamma);
}

// The constants of the E-step kept in registers across a sweep.
// The series coefficients of the E-step.  They are left in the constant bank on purpose: ptxas feeds
// them to DFMA through uniform registers, which costs a few LDCU per trip but no vector registers - and
// this kernel is bound by the number of independent FP64 chains in flight (DFMA latency is ~23 cycles on
// B200, tools/microbench/dfma_lat.cu), not by issue slots, so registers go to reads in flight instead.
struct EConst {
    double a256, lhi, llo, c3, l3;
    __device__ __forceinline__ void load() {
        a256 = XPT(kE256)[0];
        lhi = XPT(kE256)[1];
        llo = XPT(kE256)[2];
        c3 = XPT(kE3)[0];
        l3 = XPT(kE3)[1];
    }
};

// E-step of one read given d (device flavour of xpc::estep_read: same operations, same order).
template <class Tab>
__device__ __forceinline__ void estep_dev(const Tab& tab, const EConst& K, double d, double& r0, double& h) {
    const int dhi = __double2hiint(d);
    int ahi = dhi & 0x7fffffff;
    ahi = min(ahi, 0x407f4000);  // |d| clipped at 500 (gmm.py:240); low word kept: [500, 500 + 2^-12)
    const double ax = __hiloint2double(ahi, __double2loint(d));
    const double magic = 6755399441055744.0;
    const double t0 = fma(-ax, K.a256, magic);
    const double kf = t0 - magic;
    const unsigned k = (unsigned)__double2loint(t0);
    double r = fma(-kf, K.lhi, -ax);
    r = fma(-kf, K.llo, r);
    double p = fma(kExpC4, r, K.c3);
    p = fma(p, r, 0.5);
    p = fma(p, r, 1.0);
    p = fma(p, r, 1.0);
    double t = lds_tab(mad_u32(k & 255u, Tab::kES, tab.e2)) * p;
    t = __hiloint2double((int)mad_u32(k & ~255u, 4096u, (unsigned)__double2hiint(t)), __double2loint(t));  // * 2^(k >> 8)
    const double u = 1.0 + t;
    double ic, lc;
    lds_tab2(mad_u32((unsigned)__double2hiint(u) >> 13, Tab::kLS, tab.lg), ic, lc);  // top seven mantissa bits; u == 2 -> entry 128
    const double rr = fma(u, ic, -1.0);
    double q = fma(kLogC5, rr, -0.25);
    q = fma(q, rr, K.l3);
    q = fma(q, rr, -0.5);
    q = fma(q, rr, 1.0);
    const double l1p = fma(q, rr, lc);
    const double inv = xrcp3(u);
    const double rmin = t * inv;
    r0 = (dhi >= 0) ? rmin : inv;
    h = fma(rmin, ax, l1p);  // |d| > 500 only when rmin < 1e-217: the clip does not show
}

// Column sums of the team's rows of partial sums for R <= 32 / PARTS columns: PARTS lanes share a column
// (each sums a contiguous block of rows with two accumulators), then butterfly shuffles combine the
// parts, so EVERY lane ends up with the total of column lane % (32 / PARTS) (0 for columns >= R).
template <int TW, int PARTS>
__device__ __forceinline__ double team_colsum(const double* scratch, int R, int lane) {
    constexpr int RPP = 32 / PARTS, LEN = 32 * TW / PARTS, CS = 32 * TW + 1;
    const int row = lane % RPP, part = lane / RPP;
    double t0 = 0.0, t1 = 0.0;
    if (row < R) {
        const double* q = scratch + row * CS + part * LEN;
#pragma unroll(LEN <= 16 ? LEN / 2 : 4)
        for (int l = 0; l < LEN; l += 2) {
            t0 += q[l];
            t1 += q[l + 1];
        }
    }
    double t = t0 + t1;
    if (PARTS >= 2) t += shfl_xor_f64(t, RPP);
    if (PARTS >= 4) t += shfl_xor_f64(t, 2 * RPP);
    return t;
}
// sum over the columns c < G of the per-lane column totals of team_colsum (every lane gets it)
template <int PARTS>
__device__ __forceinline__ double colsum_first(double t, int G, int lane) {
    constexpr int RPP = 32 / PARTS;
    double v = (lane % RPP < G) ? t : 0.0;
#pragma unroll
    for (int m = RPP / 2; m > 0; m >>= 1) v += shfl_xor_f64(v, m);
    return v;
}

// STREAM: the position's reads stay in global memory / L2 and are decoded in every sweep - the class for positions
// deeper than one CTA's shared memory (the reference has no cap on n, gmm.py:93-127; pooled mode keeps the reads of
// conditions that failed the count test, io.py:46-58,70-73).  One team (= the CTA) per position; per read the same
// arithmetic as the resident flavour; per-thread partial sums go through a warp butterfly and one shared-memory entry
// per warp and column, which the leader adds in warp order (deterministic).
template <int TW, int MAXT, bool STREAM = false>
__global__ void __launch_bounds__(MAXT, 1) xp_fit_kernel(FitArgs a) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    const int warp = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;  // broadcast: warp-uniform for the compiler
    const int team = warp / TW, tw = warp % TW, tid = tw * 32 + lane;  // tid: thread index inside the team
    const bool leader = (tw == 0);
    const int G = a.G;
    const int R = fit_rows(G);
    constexpr int CS = STREAM ? TW + 1 : 32 * TW + 1;  // doubles between scratch columns

    double* tab_e2 = reinterpret_cast<double*>(smem_raw);
    double2* tab_lg = reinterpret_cast<double2*>(tab_e2 + kTabExpDoubles);
    for (int i = threadIdx.x; i < kTabExpDoubles; i += blockDim.x) tab_e2[i] = XPT(kExp2)[i];
    for (int i = threadIdx.x; i < 129; i += blockDim.x) tab_lg[i] = make_double2(XPT(kInvC)[i], XPT(kLogC)[i]);
    SmemTab tab;
    tab.e2 = smem_u32(tab_e2);
    tab.lg = smem_u32(tab_lg) - (0x3ff00000u >> 13) * 16u;

    unsigned char* wbase = smem_raw + kFitTableBytes + team * a.team_bytes;
    double* sc = reinterpret_cast<double*>(wbase);
    double* bc = sc + 8;
    double* pc = bc + 8;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(pc + 8);
    double* scratch = pc + 10;
    double* tot = reinterpret_cast<double*>(wbase + a.off_tot);
    double* Asm = reinterpret_cast<double*>(wbase + a.off_asm);
    double* dps = reinterpret_cast<double*>(wbase + a.off_dps);
    int* goff = reinterpret_cast<int*>(wbase + a.off_goff);
    double* slab = reinterpret_cast<double*>(wbase + a.off_slab);
    unsigned char* gid = wbase + a.off_gid;

    if (tid == 0) {
        mbar_init(mbar, 1);
        fence_mbar_init();
    }
    __syncthreads();
    uint32_t parity = 0;
    const unsigned w_begin = a.w_begin ? *a.w_begin : 0u;
    const unsigned w_end = *a.w_end;
    const double conc0 = a.conc0;
    const int kpar = lane & 1;  // component owned by this lane in the scalar phase

    for (;;) {
        // ---- next ticket -----------------------------------------------------------------------
        unsigned w = 0;
        if (TW == 1) {
            if (lane == 0) w = atomicAdd(a.ticket, 1u);
            w = __shfl_sync(kFull, w, 0);
        } else {
            if (tid == 0) reinterpret_cast<unsigned*>(bc + 3)[0] = atomicAdd(a.ticket, 1u);
            team_sync<TW>(team);
            w = reinterpret_cast<const unsigned*>(bc + 3)[0];
        }
        w += w_begin;
        if (w >= w_end) break;
        const int p = a.worklist ? a.worklist[w] : (int)w;
        const int64_t off = a.read_off[p];
        const int n = (int)(a.read_off[p + 1] - off);
        if ((!STREAM && n > a.cap) || n <= 0) {  // only when the caller understated batch.max_reads
            if (tid == 0) {
                a.status[p] = (a.status[p] & (STATUS_TTEST_VALID | STATUS_NEAR_THRESHOLD)) | STATUS_SELECTED | STATUS_NOT_FITTED;
                a.n_iter[p] = -1;
            }
            continue;
        }
        // ---- stage the reads: one TMA bulk copy of the 16-byte aligned superset -------------------
        const int fmt = a.y_format;
        const int eb = y_elem_bytes(fmt);
        const int per16 = 16 / eb;                         // elements per 16 bytes
        const int sh = (int)(off & (per16 - 1));           // elements in front of this position
        const uint32_t bytes = (uint32_t)(((n + sh) * eb + 15) & ~15);
        if (!STREAM) {
            fence_proxy_async();  // order this team's generic-proxy slab accesses before the async-proxy write
            team_sync<TW>(team);
            if (tid == 0) {
                mbar_expect_tx(mbar, bytes);
                tma_bulk_g2s(slab, reinterpret_cast<const char*>(a.y) + (off - sh) * eb, bytes, mbar);
            }
        }
        const double m0 = a.kmer_mean[p];
        const double sd = a.kmer_std[p];
        const double tau = 1.0 / (sd * sd);          // diffmod.py:24
        const double a0 = tau;                       // diffmod.py:40
        const double b0 = 0.5 * (1.0 / tau);         // diffmod.py:41
        for (int g = tid; g < G; g += TW * 32) goff[g + 1] = a.grp_len[(size_t)p * G + g];
        team_sync<TW>(team);
        if (tid == 0) {
            int acc = 0;
            goff[0] = 0;
            for (int g = 0; g < G; ++g) {
                acc += goff[g + 1];
                goff[g + 1] = acc;
            }
        }
        team_sync<TW>(team);
        if (!STREAM) {
            int g = 0, gend = goff[1];
            for (int i = tid; i < n; i += TW * 32) {
                while (i >= gend) {
                    ++g;
                    gend = goff[g + 1];
                }
                gid[i] = (unsigned char)g;
            }
        }
        // per-position constants (leader): lgamma(sum_k c_gk) of the present groups, lgamma(a0)
        double Kconst = 0.0, prior_gamma = 0.0;
        if (leader) {
            double kpart = 0.0, lga0 = 0.0;
            for (int s = lane; s <= G; s += 32) {
                double ps, lg;
                if (s < G) {
                    const int ng = goff[s + 1] - goff[s];
                    psi_lgamma(2.0 * conc0 + (double)ng, &ps, &lg);
                    if (ng > 0) kpart += a.lg_prior_w - lg;  // lnC(prior_g) - lgamma(sum_k c_gk)
                } else {
                    psi_lgamma(a0, &ps, &lg);
                    lga0 = lg;
                }
            }
            Kconst = warp_sum(kpart);
            lga0 = shfl_f64(lga0, G & 31);
            prior_gamma = a0 * xlog(b0) - lga0;  // a0 ln b0 - lgamma(a0)   (gmm.py:334)
        }

        if (!STREAM) {
            mbar_wait(mbar, parity);
            parity ^= 1u;
        }
        // decode in place and centre on the k-mer mean.  Scaled integers expand (last step first, so no
        // unread raw bytes are overwritten); doubles move down by the `sh` leading elements of the aligned
        // superset (first step first), so the reads start 16-byte aligned in either case.
        double* yv = slab;
        const double rscale = 1.0 / a.y_scale;
        constexpr int TT = TW * 32;
        double loc1 = 0.0;
        // read i of a streamed position, decoded and centred exactly as the resident flavour does it in place
        auto yg = [&](int i) -> double { return y_load_global(a.y, fmt, a.y_scale, rscale, off + i) - m0; };
        if (!STREAM && fmt == kYU16) {
            // initial locations (gmm.py:39-44) from the integer codes, before the decode overwrites them
            const unsigned short* q = reinterpret_cast<const unsigned short*>(slab) + sh;
            int* hist1 = reinterpret_cast<int*>(scratch);  // scratch is idle until the first sweep
            int* hist2 = hist1 + 256;
            for (int j = tid; j < 256; j += TT) hist1[j] = 0;
            team_sync<TW>(team);
            for (int i = tid; i < n; i += TT) atomicAdd(&hist1[q[i] >> 8], 1);
            team_sync<TW>(team);
            int lo, hi;
            double gamma;
            unsigned qa, qb;
            percentile_index(n, 0.5, lo, hi, gamma);
            team_select2_u16<TW>(q, n, lo, hist1, hist2, lane, tid, team, qa, qb);
            const bool up = prior_below_median(y_decode((double)qa, a.y_scale, rscale), y_decode((double)qb, a.y_scale, rscale), gamma, m0);
            percentile_index(n, up ? 0.75 : 0.25, lo, hi, gamma);
            team_select2_u16<TW>(q, n, lo, hist1, hist2, lane, tid, team, qa, qb);
            loc1 = np_lerp(y_decode((double)qa, a.y_scale, rscale) - m0, y_decode((double)qb, a.y_scale, rscale) - m0, gamma);
        }
        if (!STREAM) {
            const int last = ((n - 1) / TT) * TT;
            const int step = (fmt == kYF64) ? TT : -TT;
            for (int i0 = (fmt == kYF64) ? 0 : last; i0 >= 0 && i0 <= last; i0 += step) {
                const int i = i0 + tid;
                double v = 0.0;
                if (i < n) {
                    if (fmt == kYF64) v = slab[sh + i];
                    else if (fmt == kYI32) v = y_decode((double)reinterpret_cast<const int*>(slab)[sh + i], a.y_scale, rscale);
                    else v = y_decode((double)reinterpret_cast<const unsigned short*>(slab)[sh + i], a.y_scale, rscale);
                    v -= m0;
                }
                team_sync<TW>(team);
                if (i < n) yv[i] = v;
            }
        }
        // totals (and, for the generic select, min/max keys) in one fixed order, whatever the encoding was
        uint64_t kmin = ~0ull, kmax = 0ull;
        double T1 = 0.0, T2 = 0.0;
        if (STREAM) {
            for (int i = tid; i < n; i += TT) {
                const double v = yg(i);
                const uint64_t k = dkey(v);
                kmin = k < kmin ? k : kmin;
                kmax = k > kmax ? k : kmax;
                T1 += v;
                T2 = fma(v, v, T2);
            }
            kmin = warp_min_u64(kmin);
            kmax = warp_max_u64(kmax);
        } else if (fmt == kYU16) {
            for (int i = tid; i < n; i += TT) {
                const double v = yv[i];
                T1 += v;
                T2 = fma(v, v, T2);
            }
        } else {
            for (int i = tid; i < n; i += TT) {
                const double v = yv[i];
                const uint64_t k = dkey(v);
                kmin = k < kmin ? k : kmin;
                kmax = k > kmax ? k : kmax;
                T1 += v;
                T2 = fma(v, v, T2);
            }
            kmin = warp_min_u64(kmin);
            kmax = warp_max_u64(kmax);
        }
        T1 = warp_sum(T1);
        T2 = warp_sum(T2);
        if (TW > 1) {  // combine the warps of the team (fixed order: deterministic)
            uint64_t* sk = reinterpret_cast<uint64_t*>(scratch);
            team_sync<TW>(team);  // the u16 select is done with its histograms
            if (lane == 0) {
                sk[tw] = kmin;
                sk[TW + tw] = kmax;
                scratch[2 * TW + tw] = T1;
                scratch[3 * TW + tw] = T2;
            }
            team_sync<TW>(team);
            kmin = sk[0];
            kmax = sk[TW];
            T1 = scratch[2 * TW];
            T2 = scratch[3 * TW];
#pragma unroll
            for (int j = 1; j < TW; ++j) {
                kmin = sk[j] < kmin ? sk[j] : kmin;
                kmax = sk[TW + j] > kmax ? sk[TW + j] : kmax;
                T1 += scratch[2 * TW + j];
                T2 += scratch[3 * TW + j];
            }
        }
        team_sync<TW>(team);
        // ---- initial locations (gmm.py:39-44) -------------------------------------------------------
        if (STREAM) {
            int* hist = reinterpret_cast<int*>(scratch);
            loc1 = team_second_location<TW>(yg, n, m0, kmin, kmax, hist, lane, tid, team);
        } else if (fmt != kYU16) {
            int* hist = reinterpret_cast<int*>(scratch);  // scratch is idle until the first sweep
            const auto ys = [yv](int i) -> double { return yv[i]; };
            loc1 = team_second_location<TW>(ys, n, m0, kmin, kmax, hist, lane, tid, team);
        }
        team_sync<TW>(team);
        // this thread's row of partial sums.  The groups a thread meets are the same in every sweep, so the
        // entries it never writes are cleared once per position.  (Streaming: entry of this thread's WARP; every
        // warp writes every column in every sweep.)
        double* scol = scratch + (STREAM ? tw : tid);  // this thread's entry of column 0
        if (!STREAM)
            for (int g = 0; g < G; ++g) scol[g * CS] = 0.0;
        const int g_first = STREAM ? 0 : gid[0];

        // ---- leader state: lanes of parity k hold component k ---------------------------------------
        Comp c = mstep(0.0, 0.0, 0.0, a0, b0);
        if (kpar) c.mp = loc1;
        double Hent = 0.0;
        double elbo = -INFINITY, elbo_old = -INFINITY;
        if (leader && lane == 0) {
            pc[0] = m0;
            pc[1] = a0;
            pc[2] = b0;
            pc[3] = Kconst;
            pc[4] = prior_gamma;
            pc[5] = T1;
            pc[6] = T2;
            pc[7] = -INFINITY;  // ELBO of the previous sweep
        }
        int it = 0;
        uint32_t st = STATUS_SELECTED;
        if (leader)
            for (int s = lane; s < G; s += 32) tot[s] = 0.0;  // N_g0 = 0: c_gk = c0 at the first scalar phase
        // first sweep: component 1 has N = 0 too (q(w) starts at the prior, gmm.py:66) -> handled below
        bool first = true;
        __syncwarp();
#if XP_FIT_FLOOR
        // Register floor.  ptxas picks its register budget from occupancy levels (.., 72, 80, 96, 128) and, once
        // the minimal pressure fits a level, spends only that level's slack on interleaving the E-step chains:
        // a kernel that *could* run in 96 registers gets its NR chains serialised to stay there, although this
        // kernel's occupancy is set by shared memory, not registers.  A few values held across the sweep loop
        // push the minimum over that level, and the next level's slack goes to instruction-level parallelism.
        double floor_[XP_FIT_FLOOR];
#pragma unroll
        for (int q = 0; q < XP_FIT_FLOOR; ++q) floor_[q] = *reinterpret_cast<const volatile double*>(bc + (q & 7));
#endif

        for (;;) {
            bool go = true;
            double B = 0.0, C = 0.0;
            if (leader) {
                // ---- scalar phase: special functions spread over lanes, one uniform code path ----
                // slots: [0,2G) c_gk -> psi, lgamma;  2G+k alpha_k -> psi, lgamma;
                //        2G+2+k -> ln beta_k;  2G+4+k -> ln lambda_k       (slot parity = component)
                double e_loc = 0.0;
                for (int s0 = 0; s0 < 2 * G + 6; s0 += 32) {
                    const int s = s0 + lane;
                    const int role = (s < 2 * G) ? 0 : (s < 2 * G + 2) ? 1 : (s < 2 * G + 6) ? 2 : 3;
                    double x = 16.0;
                    bool counted = false;  // lgamma of this slot enters the ELBO
                    if (role == 0) {
                        const int g = s >> 1;
                        const int ng = goff[g + 1] - goff[g];
                        const double v = tot[g];
                        counted = ng > 0;
                        x = conc0 + ((kpar && !first) ? ((double)ng - v) : v);
                    } else if (role == 1) {
                        x = c.alpha;
                        counted = true;
                    } else if (role == 2) x = (s < 2 * G + 4) ? c.beta : c.lam;
                    double z, num, den;
                    shift10(role == 2 ? 16.0 : x, z, num, den);
                    if (role == 2) z = x;
                    const double lz = xlog_tab(tab, z);
                    const double lden = xlog_tab(tab, den);  // den = 1 -> exactly 0
                    const double q = xrcp3(z * den);
                    double ps, lg;
                    psi_lgamma_shifted(z, lz, q * den, num * (q * z), &ps, &lg);
                    lg -= lden;
                    const double pso = shfl_xor_f64(ps, 1);
                    if (role == 0) {
                        if (kpar) dps[s >> 1] = ps - pso;  // psi(c_g1) - psi(c_g0)  (gmm.py:283-284)
                        if (counted) e_loc += lg;          // + sum_k lgamma(c_gk)
                    } else if (role == 1) {
                        sc[kpar] = ps;
                        sc[2 + kpar] = lg;                 // lgamma(alpha_k)
                    } else if (role == 2) {
                        sc[4 + (s - 2 * G - 2)] = lz;
                    }
                }
                __syncwarp();
                double tb, lt, ek;
                if (first)
                    ek = elbo_component(c, 0.0, 0.0, 0.0, XP_PK(1, a0), XP_PK(2, b0), XP_PK(4, prior_gamma), sc[kpar], sc[2 + kpar],
                                        sc[4 + kpar], sc[6 + kpar], tb, lt);
                else
                    ek = elbo_component_closed(c, sc[kpar], sc[2 + kpar], sc[4 + kpar], sc[6 + kpar], tb, lt);
                if (lane < 2) e_loc += ek;
                elbo = warp_sum(e_loc) + XP_PK(3, Kconst) + Hent;
                if (first && lane == 0) pc[3] = pc[3] + elbo_closed_const(pc[4], n);  // constants of the closed form from now on
                if (it >= 1) {
                    const int rule = stop_rule(elbo, XP_PK(7, elbo_old), a.stopping);
                    if (rule == 2) {
                        st |= STATUS_ELBO_DECREASED;
                        go = false;
                    } else {
                        elbo_old = elbo;
                        if (lane == 0) pc[7] = elbo;
                        if (rule == 1) {
                            st |= STATUS_CONVERGED;
                            go = false;
                        }
                    }
                }
                if (go && it >= a.max_iters - 1) {
                    if (it >= 1) st |= STATUS_HIT_MAX;
                    go = false;
                }
                if (go) {
                    ++it;
                    // coefficients of d_n = A_g + y'(B + C y'): every term is (component 1) - (component 0)
                    const double u = tb * c.mp;
                    const double wq = lt - c.il - u * c.mp;
                    const double uo = shfl_xor_f64(u, 1);
                    const double tbo = shfl_xor_f64(tb, 1);
                    const double wo = shfl_xor_f64(wq, 1);
                    // (component 1) - (component 0) in every lane: the even lanes flip the sign of own - other
                    const int flip = kpar ? 0 : (int)0x80000000;
                    const double du = u - uo, dt = tb - tbo, dw = wq - wo;
                    B = __hiloint2double(__double2hiint(du) ^ flip, __double2loint(du));
                    C = -0.5 * __hiloint2double(__double2hiint(dt) ^ flip, __double2loint(dt));
                    const double base = 0.5 * __hiloint2double(__double2hiint(dw) ^ flip, __double2loint(dw));
                    if (lane < G) Asm[lane] = dps[lane] + base;
                    if (G > 32 && lane + 32 < G) Asm[lane + 32] = dps[lane + 32] + base;  // G <= kMaxGroups = 64
                    if (lane == 0) {
                        bc[0] = B;
                        bc[1] = C;
                    }
                }
                if (TW > 1 && lane == 0) bc[2] = go ? 1.0 : 0.0;
            }
            team_sync<TW>(team);
            if (TW > 1) {
                go = bc[2] != 0.0;
                B = bc[0];
                C = bc[1];
            }
            if (!go) break;
            first = false;
            // ---- E-step + component-0 moments: one pass over the resident reads ------------------
            // A thread takes the reads lane, lane + 32, ... of every full trip of 32 NR reads (one load each), then
            // single reads of the tail, so the groups it meets never decrease; its N_g0 partial goes to the
            // scratch column of the group it leaves.  The loop body is branch free.
            double accN = 0.0, S1 = 0.0, S2 = 0.0, Hh = 0.0;
            if (STREAM) {
                // group by group (the reads of a group are contiguous): thread t takes the reads t, t + TT, ... of the
                // group, two in flight; the group's N_g0 goes through a warp butterfly to the warp's entry of column g
                EConst K;
                K.load();
                for (int g = 0; g < G; ++g) {
                    const int ge = goff[g + 1];
                    const double Ag = Asm[g];
                    double aN = 0.0;
                    int i = goff[g] + tid;
                    for (; i + TT < ge; i += 2 * TT) {
                        const double y0 = yg(i), y1 = yg(i + TT);
                        const double d0 = fma(y0, fma(C, y0, B), Ag), d1 = fma(y1, fma(C, y1, B), Ag);
                        double r0, h0, r1, h1;
                        estep_dev(tab, K, d0, r0, h0);
                        estep_dev(tab, K, d1, r1, h1);
                        aN += r0;
                        aN += r1;
                        const double w0 = r0 * y0, w1 = r1 * y1;
                        S1 += w0;
                        S1 += w1;
                        S2 = fma(w0, y0, S2);
                        S2 = fma(w1, y1, S2);
                        Hh += h0;
                        Hh += h1;
                    }
                    if (i < ge) {
                        const double y0 = yg(i);
                        const double d0 = fma(y0, fma(C, y0, B), Ag);
                        double r0, h0;
                        estep_dev(tab, K, d0, r0, h0);
                        aN += r0;
                        const double w0 = r0 * y0;
                        S1 += w0;
                        S2 = fma(w0, y0, S2);
                        Hh += h0;
                    }
                    __syncwarp();
                    aN = warp_sum(aN);
                    if (lane == 0) scol[g * CS] = aN;
                }
                S1 = warp_sum(S1);
                S2 = warp_sum(S2);
                Hh = warp_sum(Hh);
            } else {
                EConst K;
                K.load();
                unsigned pg = g_first;  // group of this thread's last read; accN goes to column pg on leaving it
                constexpr int NR = (TW > 1) ? XP_FIT_NR_TEAM : XP_FIT_NR;  // reads per thread and trip
                const int nfull = n / (32 * NR);
#define XP_READ(yy, gg)                              \
    {                                                \
        const double d_ = fma((yy), fma(C, (yy), B), Asm[gg]); \
        double r_, h_;                               \
        estep_dev(tab, K, d_, r_, h_);               \
        if ((gg) != pg) {                            \
            scol[pg * CS] = accN;                         \
            accN = 0.0;                              \
        }                                            \
        pg = (gg);                                   \
        accN += r_;                                  \
        const double w_ = r_ * (yy);                 \
        S1 += w_;                                    \
        S2 = fma(w_, (yy), S2);                      \
        Hh += h_;                                    \
    }
                const double* yp = yv + tw * (32 * NR) + lane;
                const unsigned char* gp = gid + tw * (32 * NR) + lane;
                const double* const yend = yv + nfull * (32 * NR) + lane;  // tw < TW <= trips apart: reached exactly
                for (; yp < yend; yp += TW * (32 * NR), gp += TW * (32 * NR)) {
                    double y4[NR], d4[NR], r4[NR], h4[NR];
                    unsigned gg[NR];
                    // one load per read (not one wide load per thread): ptxas interleaves the NR dependency
                    // chains only when they start from separate loads
#pragma unroll
                    for (int i = 0; i < NR; ++i) {
                        gg[i] = gp[32 * i];
                        y4[i] = yp[32 * i];
                    }
#pragma unroll
                    for (int i = 0; i < NR; ++i) d4[i] = fma(y4[i], fma(C, y4[i], B), Asm[gg[i]]);
#pragma unroll
                    for (int i = 0; i < NR; ++i) estep_dev(tab, K, d4[i], r4[i], h4[i]);
#pragma unroll
                    for (int i = 0; i < NR; ++i) {
                        if (gg[i] != pg) {
                            scol[pg * CS] = accN;
                            accN = 0.0;
                        }
                        pg = gg[i];
                        accN += r4[i];
                    }
#pragma unroll
                    for (int i = 0; i < NR; ++i) {
                        const double w_ = r4[i] * y4[i];
                        S1 += w_;
                        S2 = fma(w_, y4[i], S2);
                        Hh += h4[i];
                    }
                }
                if ((nfull % TW) == tw) {  // the reads past the last full trip (lanes may diverge here)
                    for (int i = nfull * (32 * NR) + lane; i < n; i += 32) {
                        const double ya = yv[i];
                        const unsigned ga = gid[i];
                        XP_READ(ya, ga)
                    }
                    __syncwarp();
                }
#undef XP_READ
                scol[pg * CS] = accN;
            }
            if (!STREAM || lane == 0) {
                double* sg = scol + G * CS;
                sg[0] = S1;
                sg[CS] = S2;
                sg[2 * CS] = Hh;
            }
            team_sync<TW>(team);
            if (leader) {
                // column sums of the rows of partial sums (several lanes per column when there are few columns)
                double n0;
                if (!STREAM && R <= 8) {
                    const double t = team_colsum<TW, 4>(scratch, R, lane);
                    if (lane < R) tot[lane] = t;
                    n0 = colsum_first<4>(t, G, lane);
                } else if (!STREAM && R <= 16) {
                    const double t = team_colsum<TW, 2>(scratch, R, lane);
                    if (lane < R) tot[lane] = t;
                    n0 = colsum_first<2>(t, G, lane);
                } else {
                    constexpr int NC = STREAM ? TW : 32 * TW;  // entries per column
                    for (int r0 = 0; r0 < R; r0 += 32) {
                        const int row = r0 + lane;
                        double t0 = 0.0, t1 = 0.0;
                        if (row < R) {
                            const double* q = scratch + row * CS;
                            for (int l = 0; l < NC; l += 2) {
                                t0 += q[l];
                                t1 += q[l + 1];
                            }
                            tot[row] = t0 + t1;
                        }
                    }
                    __syncwarp();
                    n0 = 0.0;
                    for (int g = lane; g < G; g += 32) n0 += tot[g];
                    n0 = warp_sum(n0);
                }
                __syncwarp();
                const double s10 = tot[G], s20 = tot[G + 1];
                Hent = tot[G + 2];
                const double Nk = kpar ? ((double)n - n0) : n0;
                const double S1k = kpar ? (XP_PK(5, T1) - s10) : s10;
                const double S2k = kpar ? (XP_PK(6, T2) - s20) : s20;
                // ---- M-step (gmm.py:292-294, 364-370) ---------------------------------------------
                c = mstep(Nk, S1k, S2k, XP_PK(1, a0), XP_PK(2, b0));
            }
        }
#if XP_FIT_FLOOR
#pragma unroll
        for (int q = 0; q < XP_FIT_FLOOR; ++q)
            if (floor_[q] == 1.2345e-300) *reinterpret_cast<volatile double*>(bc + (q & 7)) = floor_[q];
#endif
        // ---- write the fit record ------------------------------------------------------------------
        if (leader) {
            double* out = a.fit + (size_t)p * fit_width(G);
            if (lane < 2) {
                out[FIT_MU + kpar] = XP_PK(0, m0) + c.mp;
                out[FIT_LAM + kpar] = c.lam;
                out[FIT_ALPHA + kpar] = c.alpha;
                out[FIT_BETA + kpar] = c.beta;
            }
            if (a.coef != nullptr && it >= 1) {  // --save_models: coefficients of the last E-step
                double* cf = a.coef + (size_t)p * (G + 2);
                for (int g = lane; g < G; g += 32) cf[g] = Asm[g];
                if (lane < 2) cf[G + lane] = bc[lane];
            }
            if (lane == 0) {
                out[FIT_ELBO] = elbo;
                a.n_iter[p] = it;
                a.status[p] = (a.status[p] & (STATUS_TTEST_VALID | STATUS_NEAR_THRESHOLD)) | st;
            }
            for (int s = lane; s < 2 * G; s += 32) {
                const int g = s >> 1;
                const double v = tot[g];
                out[FIT_N + s] = (it == 0) ? 0.0 : (kpar ? (double)(goff[g + 1] - goff[g]) - v : v);
            }
        }
        team_sync<TW>(team);
    }
}

}  // namespace xpk
