// xp_fit_tm.cuh — K3 for u16 batches: one warp per position with the position's reads resident in TENSOR MEMORY.
//
// Nothing on this path is a matrix product, but the SM's 256 KB of tensor memory (TMEM) sit idle next to a kernel
// whose resident warps are limited by shared memory (a position's reads, 8 bytes each, stay on chip for ~50 sweeps)
// and whose LSU pipe is the second busiest unit (78 % of peak, a quarter of it bank conflicts of the exp / log table
// gathers, profiles/r1_fit_v14_c2_full.ncu_summary.txt).  TMEM is 128 lanes x 512 columns x 32 bit; warp w of a CTA
// reads and writes the 32 lanes of quadrant w % 4 with tcgen05.ld / tcgen05.st (SASS LDTM / STTM), thread i its own
// lane.  That is exactly the access pattern of the E-step: thread `lane` takes the reads lane, lane + 32, ... of its
// position.  So read i of a position lives in TMEM lane i % 32 of the warp's quadrant, columns 2 (i / 32) and
// 2 (i / 32) + 1 of the position's column range (a double = two 32-bit columns), and a trip of the E-step fetches
// its two reads with one tcgen05.ld.32x32b.x4 - no shared-memory wavefront, no bank.
//
// What the freed shared memory buys:
//   * the raw u16 codes are staged by TMA into a 2-byte-per-read buffer that only lives through the set-up (radix
//     select on the codes, decode, centre, tcgen05.st), 1.2 KB instead of a 5.3 KB slab on c2;
//   * the exp2 and {1/c, ln c} tables are replicated per lane slot (16 x 8-byte and 8 x 16-byte copies of every entry,
//     one 128-byte row per entry, 48 KB): the gathers of a half / quarter warp then hit distinct banks whatever the
//     indices are - the address arithmetic is the same single IMAD because the lane's offset sits in the base;
//   * NPW = 2 positions per warp (2 x 36 TMEM columns x 7 warps per quadrant = 504 of 512 columns): the warp runs the
//     E-steps of its two positions one after the other with all 32 lanes and then ONE scalar phase for both, 16 lanes
//     each (2G + 4 slots per position, slot parity = component: [0, 2G) c_gk -> psi, lgamma; 2G + k alpha_k -> psi,
//     lgamma; 2G + 2 + k -> ln beta_k and ln lambda_k, the slot's two table logarithms).  The scalar phase is 41 % of
//     the warp instructions of a c2 position-sweep in xp_fit.cuh, and the kernel is bound by dispatch slots.
//     (Sharing the scalar phase between two WARPS instead - one position each, a named barrier per sweep - removed
//     the same instructions but ran 1 % slower: the waiting warp is lost to latency hiding, and resident warps are
//     capped at 28 by the register file.  Measured)
//
// The arithmetic of a position is the one of xp_fit.cuh operation by operation, with the same summation order of its
// reads, columns and ELBO terms (lanes >= 2G + 2 add exact zeros there), so records are bit-identical to the
// shared-memory flavour - tests/test_gpu_parity.py compares u16 batches (this kernel) with their f64 twins (xp_fit.cuh).
//
// CTA shared memory: [ exp2 rows 256 x 128 B | {1/c, ln c} rows 129 x 128 B | tmem base (16 B) | warp 0 | warp 1 | ... ]
// warp region:       [ mbar (16 B) | header of position 0 | (header of position 1) | staging (u16 codes, TMA
//                      destination) | scratch columns (G + 3) x 33 doubles, >= 2 KB: select histograms before the first sweep ]
// position header:   sc[8] bc[8] pc[8] | tot[G + 4] | Asm[G] | dps[G] | goff[G + 1] (int) | gid[cap] (u8)
//   bc: 0 B, 1 C, 2 go flag, 3 state (int: 0 out of work, 1 swept, 2 new), 4 {p, n} (ints), 5 second start location,
//       6 chunk of the position (int; chained launch)
//   pc: m0, a0, b0, ELBO constant, prior_gamma, sum y', sum y'^2, previous ELBO;  tot[G + 3] = N_0
#pragma once
#include "xp_fit.cuh"

namespace xpk {

constexpr int kTmExpRowBytes = 128, kTmLogRowBytes = 128;
constexpr size_t kTmTableBytes = 256 * kTmExpRowBytes + 129 * kTmLogRowBytes + 16;  // + the TMEM base address
constexpr int kTmColumns = 512;
using SmemTabR = SmemTabT<128u, 128u>;

__host__ __device__ inline size_t fit_tm_hdr_bytes(int G, int cap) {
    const size_t b = (size_t)(24 + (G + 4) + 2 * G) * 8 + (size_t)((G + 2) & ~1) * 4 + (size_t)((cap + 15) & ~15);
    return (b + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t fit_tm_stage_bytes(int cap) { return ((size_t)(cap + 8) * 2 + 15) & ~(size_t)15; }
__host__ __device__ inline size_t fit_tm_scratch_bytes(int G) {
    const size_t s = (size_t)fit_rows(G) * 33 * 8;
    return ((s > 2048 ? s : 2048) + 15) & ~(size_t)15;
}
__host__ __device__ inline size_t fit_tm_warp_bytes(int cap, int G, int npw) {
    return 16 + (size_t)npw * fit_tm_hdr_bytes(G, cap) + fit_tm_stage_bytes(cap) + fit_tm_scratch_bytes(G);
}
__host__ __device__ inline int fit_tm_cols(int cap) { return 2 * ((cap + 31) / 32); }  // TMEM columns of one position

inline void fit_tm_fill_layout(FitArgs& a, int npw) {
    const int G = a.G;
    a.hdr_bytes = (int)fit_tm_hdr_bytes(G, a.cap);
    a.team_bytes = (int)fit_tm_warp_bytes(a.cap, G, npw);
    a.off_tot = 24 * 8;  // relative to a position's header
    a.off_asm = a.off_tot + (G + 4) * 8;
    a.off_dps = a.off_asm + G * 8;
    a.off_goff = a.off_dps + G * 8;
    a.off_gid = a.off_goff + ((G + 2) & ~1) * 4;
    a.off_slab = 16 + npw * a.hdr_bytes;  // staging, relative to the warp's region
    a.off_scratch = a.off_slab + (int)fit_tm_stage_bytes(a.cap);
    a.tm_cols = fit_tm_cols(a.cap);
}

// ---- tensor memory ------------------------------------------------------------------------------------------
__device__ __forceinline__ void tm_alloc_all(uint32_t* smem_dst) {  // one warp; the whole TMEM of the SM (1 CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_dst)), "n"(kTmColumns) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tm_free_all(uint32_t base) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "n"(kTmColumns) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_u32(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// Load of an input array in the chained launch: the launch lives through the arrival of later chunks, whose offsets, group
// lengths, priors and prefilter status land in the same arrays - and in the same 32-byte sectors as the last positions of
// the chunk before, which this SM may have read (and its L1 kept) before they arrived.  L1 is coherent only at launch
// boundaries, so these loads go to L2 (ld.global.cg).  A plain launch reads arrays that were complete before it started.
template <bool CG, class T>
__device__ __forceinline__ T ld_in(const T* p) {
    return CG ? __ldcg(p) : *p;
}
__device__ __forceinline__ void tm_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tm_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
// one double of this thread's lane (two columns); the data are in the register when the statement is done
__device__ __forceinline__ double tm_ld1(uint32_t taddr) {
    int lo, hi;
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%2];\n"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(lo), "=r"(hi)
        : "r"(taddr)
        : "memory");
    return __hiloint2double(hi, lo);
}
__device__ __forceinline__ void tm_ld2(uint32_t taddr, double& y0, double& y1) {
    int l0, h0, l1, h1;
    asm volatile(  // two x2 loads: each lands in a register pair that is a double as it stands (one x4 load wants four
                   // consecutive registers, and the compiler then copies both doubles out of them)
        "tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0, %1}, [%4];\n"
        "tcgen05.ld.sync.aligned.32x32b.x2.b32 {%2, %3}, [%4 + 2];\n"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(l0), "=r"(h0), "=r"(l1), "=r"(h1)
        : "r"(taddr)
        : "memory");
    y0 = __hiloint2double(h0, l0);
    y1 = __hiloint2double(h1, l1);
}
__device__ __forceinline__ void tm_st1(uint32_t taddr, double v) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1, %2};" ::"r"(taddr), "r"(__double2loint(v)), "r"(__double2hiint(v)) : "memory");
}
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// GT: the number of group slots as a compile-time constant (0 = read it from the arguments): the common designs get
// their G-sized loops unrolled and the variants of the column sums chosen at compile time.
// CHAIN: the launch of the chunked host call that follows the chain of chunk worklists (FitChain, xp_fit.cuh).
template <int NPW, int MAXT, int GT = 0, bool CHAIN = false>
__global__ void __launch_bounds__(MAXT, 1) xp_fit_tm_kernel(FitArgs a, FitChain ch) {
    extern __shared__ __align__(16) unsigned char smem_raw[];
    constexpr int LP = 32 / NPW;  // lanes per position in the scalar phase
    constexpr int CS = 33;        // doubles between scratch columns
    const int warp = __shfl_sync(kFull, (int)(threadIdx.x >> 5), 0), lane = threadIdx.x & 31;
    const int G = GT ? GT : a.G;
    const int R = fit_rows(G);

    // ---- tables: one 128-byte row per entry, the entry repeated in every lane slot of the row ----------------
    double* tab_e2 = reinterpret_cast<double*>(smem_raw);
    double2* tab_lg = reinterpret_cast<double2*>(smem_raw + 256 * kTmExpRowBytes);
    uint32_t* tm_base_smem = reinterpret_cast<uint32_t*>(smem_raw + 256 * kTmExpRowBytes + 129 * kTmLogRowBytes);
    for (int i = threadIdx.x; i < 256 * 16; i += blockDim.x) tab_e2[i] = XPT(kExp2)[i >> 4];
    for (int i = threadIdx.x; i < 129 * 8; i += blockDim.x) tab_lg[i] = make_double2(XPT(kInvC)[i >> 3], XPT(kLogC)[i >> 3]);
    SmemTabR tab;
    tab.e2 = smem_u32(tab_e2) + (lane & 15) * 8;
    tab.lg = smem_u32(tab_lg) + (lane & 7) * 16 - (0x3ff00000u >> 13) * (unsigned)kTmLogRowBytes;

    unsigned char* wbase = smem_raw + kTmTableBytes + (size_t)warp * a.team_bytes;
    uint64_t* mbar = reinterpret_cast<uint64_t*>(wbase);
    unsigned char* hdr0 = wbase + 16;
    unsigned char* stage = wbase + a.off_slab;
    double* scratch = reinterpret_cast<double*>(wbase + a.off_scratch);
    // the position this lane serves in the scalar phase
    const int sub = lane / LP, sl = lane % LP;
    unsigned char* lbase = hdr0 + sub * a.hdr_bytes;
    double* Lsc = reinterpret_cast<double*>(lbase);
    double* Lbc = Lsc + 8;
    double* Lpc = Lbc + 8;
    double* Ltot = reinterpret_cast<double*>(lbase + a.off_tot);
    double* LAsm = reinterpret_cast<double*>(lbase + a.off_asm);
    double* Ldps = reinterpret_cast<double*>(lbase + a.off_dps);
    const int* Lgoff = reinterpret_cast<const int*>(lbase + a.off_goff);

    if (warp == 0) tm_alloc_all(tm_base_smem);
    if (lane == 0) {
        mbar_init(mbar, 1);
        fence_mbar_init();
    }
    tm_fence_before();
    __syncthreads();
    tm_fence_after();
    // warp-uniform for the compiler (LDTM / STTM take their address from a uniform register): REDUX results are
    const uint32_t tm_base = __reduce_max_sync(kFull, *tm_base_smem);
    {   // the biased table base through a volatile shared-memory round trip: ptxas otherwise splits it into a register, an
        // add per lookup and an immediate offset (the bias times the 128-byte row pitch exceeds the immediate field)
        uint32_t* slot = reinterpret_cast<uint32_t*>(scratch) + lane;
        *reinterpret_cast<volatile uint32_t*>(slot) = tab.lg;
        __syncwarp();
        tab.lg = *reinterpret_cast<volatile uint32_t*>(slot);
        __syncwarp();
    }
    // this warp's columns: quadrant warp % 4 (the lanes a warp may touch), column block warp / 4
    const uint32_t tm_warp = __reduce_max_sync(kFull, tm_base + ((uint32_t)(warp & 3) << 21) + (uint32_t)(warp >> 2) * (uint32_t)(NPW * a.tm_cols));

    uint32_t parity = 0;
    const unsigned w_begin = (!CHAIN && a.w_begin) ? *a.w_begin : 0u;
    const unsigned w_end = CHAIN ? 0xffffffffu : *a.w_end;  // CHAIN: unused
    const double conc0 = a.conc0;
    const double rscale = 1.0 / a.y_scale;
    const int kpar = lane & 1;  // component owned by this lane in the scalar phase (LP is even)

    unsigned alive = (1u << NPW) - 1u, need = alive;  // per position slot of this warp (warp-uniform)
    int* wcur = reinterpret_cast<int*>(mbar) + 2;  // CHAIN: the chunk this warp takes its positions from (second half of the barrier's 16 bytes)
    if (CHAIN && lane == 0) *wcur = 0;
    int it = 0;  // sweeps of position `sub` so far: all the scalar phase carries from sweep to sweep in registers

    for (;;) {
        // =========================== new positions =====================================================
        for (int h = 0; h < NPW; ++h) {
            if (!((need & alive) >> h & 1u)) continue;
            unsigned char* hb = hdr0 + h * a.hdr_bytes;
            double* bc = reinterpret_cast<double*>(hb) + 8;
            double* pc = bc + 8;
            double* tot = reinterpret_cast<double*>(hb + a.off_tot);
            int* goff = reinterpret_cast<int*>(hb + a.off_goff);
            unsigned char* gid = hb + a.off_gid;
            int p = 0, n = 0, cc = 0;
            int64_t off = 0;
            bool got = false;
            for (;;) {
                unsigned w = 0;
                if (CHAIN) {
                    // lane 0 walks the chain of chunk worklists from the warp's current chunk (kept in shared memory)
                    if (lane == 0) {
                        cc = *wcur;
                        while (cc < ch.n) {
                            unsigned spins = 0, ns = 250;
                            while (ld_acquire_u32(ch.ready + cc) == 0u) {  // the chunk's reads / prefilter / worklist are still on their way
                                if (*reinterpret_cast<volatile unsigned int*>(ch.abort) != 0u || ++spins > ch.spin_limit) {
                                    atomicExch(ch.abort, 1u);
                                    cc = ch.n;
                                    break;
                                }
                                __nanosleep(ns);
                                ns = min(ns * 2u, 4000u);
                            }
                            if (cc >= ch.n) break;
#ifdef XPB200_EXPERIMENTS
                            if (spins) atomicAdd(ch.wait_polls + cc, spins);
#endif
                            w = atomicAdd(ch.c[cc].ticket, 1u);
                            if (ch.c[cc].w_begin) w += *reinterpret_cast<const volatile uint32_t*>(ch.c[cc].w_begin);
                            if (w < *reinterpret_cast<const volatile uint32_t*>(ch.c[cc].n_work)) break;
                            ++cc;
                        }
                        *wcur = cc;
                    }
                    cc = __shfl_sync(kFull, cc, 0);
                    if (cc >= ch.n) break;
                    w = __shfl_sync(kFull, w, 0);
                    p = __ldcg(ch.c[cc].worklist + w) + ch.c[cc].p0;
                } else {
                    if (lane == 0) w = atomicAdd(a.ticket, 1u);
                    w = __shfl_sync(kFull, w, 0) + w_begin;
                    if (w >= w_end) break;
                    p = a.worklist ? a.worklist[w] : (int)w;
                }
                off = ld_in<CHAIN>(reinterpret_cast<const long long*>(a.read_off) + p);
                n = (int)(ld_in<CHAIN>(reinterpret_cast<const long long*>(a.read_off) + p + 1) - off);
                if (n > a.cap || n <= 0) {  // only when the caller understated batch.max_reads
                    if (lane == 0) {
                        a.status[p] = (ld_in<CHAIN>(a.status + p) & (STATUS_TTEST_VALID | STATUS_NEAR_THRESHOLD)) | STATUS_SELECTED | STATUS_NOT_FITTED;
                        a.n_iter[p] = -1;
                        if (CHAIN) {
                            __threadfence();
                            atomicAdd(ch.done + cc, 1u);
                        }
                    }
                    continue;
                }
                got = true;
                break;
            }
            if (!got) {
                alive &= ~(1u << h);
                if (lane == 0) reinterpret_cast<int*>(bc + 3)[0] = 0;
                continue;
            }
            // ---- stage the u16 codes: one TMA bulk copy of the 16-byte aligned superset ---------------------
            const int sh = (int)(off & 7);  // codes in front of this position
            const uint32_t bytes = (uint32_t)(((n + sh) * 2 + 15) & ~15);
            fence_proxy_async();  // order this warp's generic-proxy accesses of the staging buffer before the async-proxy write
            __syncwarp();
            if (lane == 0) {
                mbar_expect_tx(mbar, bytes);
                tma_bulk_g2s(stage, reinterpret_cast<const char*>(a.y) + (off - sh) * 2, bytes, mbar);
            }
            const double m0 = ld_in<CHAIN>(a.kmer_mean + p);
            const double sd = ld_in<CHAIN>(a.kmer_std + p);
            const double tau = 1.0 / (sd * sd);   // diffmod.py:24
            const double a0 = tau;                // diffmod.py:40
            const double b0 = 0.5 * (1.0 / tau);  // diffmod.py:41
            for (int g = lane; g < G; g += 32) goff[g + 1] = ld_in<CHAIN>(a.grp_len + (size_t)p * G + g);
            __syncwarp();
            if (lane == 0) {
                int acc = 0;
                goff[0] = 0;
                for (int g = 0; g < G; ++g) {
                    acc += goff[g + 1];
                    goff[g + 1] = acc;
                }
            }
            __syncwarp();
            {
                int g = 0, gend = goff[1];
                for (int i = lane; i < n; i += 32) {
                    while (i >= gend) {
                        ++g;
                        gend = goff[g + 1];
                    }
                    gid[i] = (unsigned char)g;
                }
            }
            // per-position constants: lgamma(sum_k c_gk) of the present groups, lgamma(a0)
            double Kconst, prior_gamma;
            {
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
            mbar_wait(mbar, parity);
            parity ^= 1u;
            // ---- initial locations (gmm.py:39-44) from the integer codes (the decode is strictly increasing) ----
            const unsigned short* q = reinterpret_cast<const unsigned short*>(stage) + sh;
            double loc1;
            {
                int* hist1 = reinterpret_cast<int*>(scratch);  // scratch is idle between a position's sweeps
                int* hist2 = hist1 + 256;
                for (int j = lane; j < 256; j += 32) hist1[j] = 0;
                __syncwarp();
                for (int i = lane; i < n; i += 32) atomicAdd(&hist1[q[i] >> 8], 1);
                __syncwarp();
                int lo, hi;
                double gamma;
                unsigned qa, qb;
                percentile_index(n, 0.5, lo, hi, gamma);
                team_select2_u16<1>(q, n, lo, hist1, hist2, lane, lane, 0, qa, qb);
                const bool up = prior_below_median(y_decode((double)qa, a.y_scale, rscale), y_decode((double)qb, a.y_scale, rscale), gamma, m0);
                percentile_index(n, up ? 0.75 : 0.25, lo, hi, gamma);
                team_select2_u16<1>(q, n, lo, hist1, hist2, lane, lane, 0, qa, qb);
                loc1 = np_lerp(y_decode((double)qa, a.y_scale, rscale) - m0, y_decode((double)qb, a.y_scale, rscale) - m0, gamma);
            }
            // ---- decode, centre on the k-mer mean, park in tensor memory; totals in the order of xp_fit.cuh ----
            const uint32_t tpos = tm_warp + (uint32_t)(h * a.tm_cols);
            double T1 = 0.0, T2 = 0.0;
            {
                const int rows = (n + 31) >> 5;
                for (int j = 0; j < rows; ++j) {
                    const int i = j * 32 + lane;
                    double v = 0.0;
                    if (i < n) {
                        v = y_decode((double)q[i], a.y_scale, rscale) - m0;
                        T1 += v;
                        T2 = fma(v, v, T2);
                    }
                    tm_st1(tpos + 2u * (uint32_t)j, v);
                }
                tm_wait_st();
            }
            T1 = warp_sum(T1);
            T2 = warp_sum(T2);
            __syncwarp();
            if (NPW == 1)  // the groups a thread meets are the same in every sweep: its other entries are cleared once
                for (int g = 0; g < G; ++g) scratch[lane + g * CS] = 0.0;
            for (int s = lane; s < G; s += 32) tot[s] = 0.0;  // N_g0 = 0: c_gk = c0 at the first scalar phase
            if (lane == 0) {
                pc[0] = m0;
                pc[1] = a0;
                pc[2] = b0;
                pc[3] = Kconst;
                pc[4] = prior_gamma;
                pc[5] = T1;
                pc[6] = T2;
                pc[7] = -INFINITY;  // ELBO of the previous sweep
                reinterpret_cast<int*>(bc + 3)[0] = 2;
                reinterpret_cast<int*>(bc + 4)[0] = p;
                reinterpret_cast<int*>(bc + 4)[1] = n;
                if (CHAIN) reinterpret_cast<int*>(bc + 6)[0] = cc;
                bc[5] = loc1;
            }
        }
        if (alive == 0u) break;
        __syncwarp();
        // =========================== scalar phase of the warp's positions ================================
        {
            const int state = reinterpret_cast<const int*>(Lbc + 3)[0];
            const bool live = state != 0, fresh = state == 2;
            double Nk = 0.0, S1k = 0.0, S2k = 0.0, Hent = 0.0, elbo;
            uint32_t st = STATUS_SELECTED;
            const int n_s = reinterpret_cast<const int*>(Lbc + 4)[1];
            if (fresh) it = 0;
            const bool first = it == 0;  // the first scalar phase of a position: N = 0, locations = the percentile start
            if (!fresh) {
                const double n0 = Ltot[G + 3], s10 = Ltot[G], s20 = Ltot[G + 1];
                Hent = Ltot[G + 2];
                Nk = kpar ? ((double)n_s - n0) : n0;
                S1k = kpar ? (Lpc[5] - s10) : s10;
                S2k = kpar ? (Lpc[6] - s20) : s20;
            }
            // ---- M-step (gmm.py:292-294, 364-370); a new position starts from the prior and the percentile ----
            Comp c = mstep(Nk, S1k, S2k, Lpc[1], Lpc[2]);
            if (fresh && kpar) c.mp = Lbc[5];
            // ---- special functions spread over the lanes of the position, one uniform code path ----
            double e_loc = 0.0;
            for (int s0 = 0; s0 < 2 * G + 4; s0 += LP) {
                const int s = s0 + sl;
                const int role = (s < 2 * G) ? 0 : (s < 2 * G + 2) ? 1 : (s < 2 * G + 4) ? 2 : 3;
                double x = 16.0, x2 = 1.0;
                bool counted = false;  // lgamma of this slot enters the ELBO
                if (role == 0) {
                    const int g = s >> 1;
                    const int ng = Lgoff[g + 1] - Lgoff[g];
                    const double v = Ltot[g];
                    counted = ng > 0;
                    x = conc0 + ((kpar && !first) ? ((double)ng - v) : v);
                } else if (role == 1) {
                    x = c.alpha;
                    counted = true;
                } else if (role == 2) {
                    x = c.beta;
                    x2 = c.lam;
                }
                double z, num, den;
                shift10(role == 2 ? 16.0 : x, z, num, den);
                if (role == 2) z = x;
                const double lz = xlog_tab(tab, z);
                const double lden = xlog_tab(tab, role == 2 ? x2 : den);  // den = 1 -> exactly 0
                const double q = xrcp3(z * den);
                double ps, lg;
                psi_lgamma_shifted(z, lz, q * den, num * (q * z), &ps, &lg);
                lg -= lden;
                const double pso = shfl_xor_f64(ps, 1);
                if (role == 0) {
                    if (kpar) Ldps[s >> 1] = ps - pso;  // psi(c_g1) - psi(c_g0)  (gmm.py:283-284)
                    if (counted) e_loc += lg;           // + sum_k lgamma(c_gk)
                } else if (role == 1) {
                    Lsc[kpar] = ps;
                    Lsc[2 + kpar] = lg;  // lgamma(alpha_k)
                } else if (role == 2) {
                    Lsc[4 + kpar] = lz;    // ln beta_k
                    Lsc[6 + kpar] = lden;  // ln lambda_k
                }
            }
            __syncwarp();
            double tb, lt, ek;
            ek = elbo_component_closed(c, Lsc[kpar], Lsc[2 + kpar], Lsc[4 + kpar], Lsc[6 + kpar], tb, lt);
            if (__any_sync(kFull, first)) {
                double tb9, lt9;
                const double ek9 = elbo_component(c, 0.0, 0.0, 0.0, Lpc[1], Lpc[2], Lpc[4], Lsc[kpar], Lsc[2 + kpar],
                                                  Lsc[4 + kpar], Lsc[6 + kpar], tb9, lt9);
                if (first) ek = ek9;
            }
            if (sl < 2) e_loc += ek;
            const double elbo_prev = Lpc[7];
            {
                double v = e_loc;  // sum over the position's lanes (lanes >= 2G + 2 hold exact zeros: the order of xp_fit.cuh)
#pragma unroll
                for (int m = LP / 2; m > 0; m >>= 1) v += shfl_xor_f64(v, m);
                elbo = v + Lpc[3] + Hent;
            }
            __syncwarp();  // every lane has read the parked values before lane 0 of the group replaces them
            if (first && sl == 0) Lpc[3] = Lpc[3] + elbo_closed_const(Lpc[4], n_s);  // constants of the closed form from now on
            bool go = true;
            if (it >= 1) {
                const int rule = stop_rule(elbo, elbo_prev, a.stopping);
                if (rule == 2) {
                    st |= STATUS_ELBO_DECREASED;
                    go = false;
                } else {
                    if (sl == 0) Lpc[7] = elbo;
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
            go = go && live;
            {
                // coefficients of d_n = A_g + y'(B + C y'): every term is (component 1) - (component 0).  Evaluated by
                // every lane (the shuffles must stay convergent: the positions of a warp stop at different sweeps)
                const double u = tb * c.mp;
                const double wq = lt - c.il - u * c.mp;
                const double uo = shfl_xor_f64(u, 1);
                const double tbo = shfl_xor_f64(tb, 1);
                const double wo = shfl_xor_f64(wq, 1);
                // (component 1) - (component 0) in every lane: the even lanes flip the sign of own - other
                const int flip = kpar ? 0 : (int)0x80000000;
                const double du = u - uo, dt = tb - tbo, dw = wq - wo;
                const double B = __hiloint2double(__double2hiint(du) ^ flip, __double2loint(du));
                const double C = -0.5 * __hiloint2double(__double2hiint(dt) ^ flip, __double2loint(dt));
                const double base = 0.5 * __hiloint2double(__double2hiint(dw) ^ flip, __double2loint(dw));
                if (go) {
                    ++it;
                    for (int g = sl; g < G; g += LP) LAsm[g] = Ldps[g] + base;
                    if (sl == 0) {
                        Lbc[0] = B;
                        Lbc[1] = C;
                    }
                }
            }
            if (live && sl == 0) Lbc[2] = go ? 1.0 : 0.0;
            // ---- the record of a finished position ---------------------------------------------------
            if (live && !go) {
                const int p_s = reinterpret_cast<const int*>(Lbc + 4)[0];
                double* out = a.fit + (size_t)p_s * fit_width(G);
                if (sl < 2) {
                    out[FIT_MU + kpar] = Lpc[0] + c.mp;
                    out[FIT_LAM + kpar] = c.lam;
                    out[FIT_ALPHA + kpar] = c.alpha;
                    out[FIT_BETA + kpar] = c.beta;
                }
                if (a.coef != nullptr && it >= 1) {  // --save_models: coefficients of the last E-step
                    double* cf = a.coef + (size_t)p_s * (G + 2);
                    for (int g = sl; g < G; g += LP) cf[g] = LAsm[g];
                    if (sl < 2) cf[G + sl] = Lbc[sl];
                }
                if (sl == 0) {
                    out[FIT_ELBO] = elbo;
                    a.n_iter[p_s] = it;
                    a.status[p_s] = (ld_in<CHAIN>(a.status + p_s) & (STATUS_TTEST_VALID | STATUS_NEAR_THRESHOLD)) | st;
                }
                for (int s = sl; s < 2 * G; s += LP) {
                    const int g = s >> 1;
                    const double v = Ltot[g];
                    out[FIT_N + s] = (it == 0) ? 0.0 : (kpar ? (double)(Lgoff[g + 1] - Lgoff[g]) - v : v);
                }
                if (CHAIN) __threadfence();  // the record before the chunk's count of finished positions
            }
            if (CHAIN) {
                __syncwarp();
                if (live && !go && sl == 0) atomicAdd(ch.done + reinterpret_cast<const int*>(Lbc + 6)[0], 1u);
            }
        }
        __syncwarp();
        // =========================== E-step + component-0 moments of every position that goes on ===========
        for (int h = 0; h < NPW; ++h) {
            if (!((alive >> h) & 1u)) continue;
            unsigned char* hb = hdr0 + h * a.hdr_bytes;
            double* bc = reinterpret_cast<double*>(hb) + 8;
            if (bc[2] == 0.0) {
                need |= 1u << h;
                continue;
            }
            need &= ~(1u << h);
            double* tot = reinterpret_cast<double*>(hb + a.off_tot);
            const double* Asm = reinterpret_cast<const double*>(hb + a.off_asm);
            const unsigned char* gid = hb + a.off_gid;
            const int n = __reduce_max_sync(kFull, reinterpret_cast<const int*>(bc + 4)[1]);  // warp-uniform trip counts and TMEM addresses
            const uint32_t tpos = tm_warp + (uint32_t)(h * a.tm_cols);
            const double B = bc[0], C = bc[1];
            double accN = 0.0, S1 = 0.0, S2 = 0.0, Hh = 0.0;
            EConst K;
            K.load();
            double* scol = scratch + lane;
            if (NPW > 1)  // the scratch columns are shared by the warp's positions: clear this thread's N_g0 entries
                for (int g = 0; g < G; ++g) scol[g * CS] = 0.0;
            unsigned pg = gid[0];  // group of this thread's last read; accN goes to column pg on leaving it
            const int nfull = n >> 6;
            const unsigned char* gp = gid + lane;
            for (int t = 0; t < nfull; ++t, gp += 64) {
                double y4[2], d4[2], r4[2], h4[2];
                unsigned gg[2];
                gg[0] = gp[0];
                gg[1] = gp[32];
                tm_ld2(tpos + 4u * (uint32_t)t, y4[0], y4[1]);
#pragma unroll
                for (int i = 0; i < 2; ++i) d4[i] = fma(y4[i], fma(C, y4[i], B), Asm[gg[i]]);
#pragma unroll
                for (int i = 0; i < 2; ++i) estep_dev(tab, K, d4[i], r4[i], h4[i]);
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    if (gg[i] != pg) {
                        scol[pg * CS] = accN;
                        accN = 0.0;
                    }
                    pg = gg[i];
                    accN += r4[i];
                }
#pragma unroll
                for (int i = 0; i < 2; ++i) {
                    const double w_ = r4[i] * y4[i];
                    S1 += w_;
                    S2 = fma(w_, y4[i], S2);
                    Hh += h4[i];
                }
            }
            for (int i0 = nfull << 6; i0 < n; i0 += 32) {  // the rows past the last full trip (at most two)
                const double ya = tm_ld1(tpos + (uint32_t)(i0 >> 4));  // row i0 / 32 -> column 2 (i0 / 32)
                const int i = i0 + lane;
                if (i < n) {
                    const unsigned ga = gid[i];
                    const double d_ = fma(ya, fma(C, ya, B), Asm[ga]);
                    double r_, h_;
                    estep_dev(tab, K, d_, r_, h_);
                    if (ga != pg) {
                        scol[pg * CS] = accN;
                        accN = 0.0;
                    }
                    pg = ga;
                    accN += r_;
                    const double w_ = r_ * ya;
                    S1 += w_;
                    S2 = fma(w_, ya, S2);
                    Hh += h_;
                }
                __syncwarp();
            }
            scol[pg * CS] = accN;
            double* sg = scol + G * CS;
            sg[0] = S1;
            sg[CS] = S2;
            sg[2 * CS] = Hh;
            __syncwarp();
            // column sums of the rows of partial sums (several lanes per column when there are few columns)
            double n0;
            if (R <= 8) {
                const double t = team_colsum<1, 4>(scratch, R, lane);
                if (lane < R) tot[lane] = t;
                n0 = colsum_first<4>(t, G, lane);
            } else if (R <= 16) {
                const double t = team_colsum<1, 2>(scratch, R, lane);
                if (lane < R) tot[lane] = t;
                n0 = colsum_first<2>(t, G, lane);
            } else {
                for (int r0 = 0; r0 < R; r0 += 32) {
                    const int row = r0 + lane;
                    double t0 = 0.0, t1 = 0.0;
                    if (row < R) {
                        const double* qq = scratch + row * CS;
                        for (int l = 0; l < 32; l += 2) {
                            t0 += qq[l];
                            t1 += qq[l + 1];
                        }
                        tot[row] = t0 + t1;
                    }
                }
                __syncwarp();
                n0 = 0.0;
                for (int g = lane; g < G; g += 32) n0 += tot[g];
                n0 = warp_sum(n0);
            }
            if (lane == 0) {
                tot[G + 3] = n0;
                reinterpret_cast<int*>(bc + 3)[0] = 1;
            }
            __syncwarp();  // the scratch columns are free for the warp's next position
        }
    }
    tm_fence_before();
    __syncthreads();
    if (warp == 0) tm_free_all(tm_base);
}

}  // namespace xpk
