// xp_api.cu — C ABI (include/xpore_b200.h) over the kernels in xp_kernels.cuh.
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <cstring>
#include <chrono>
#include <map>
#include <mutex>
#include <thread>
#include <string>
#include <vector>

#include "../../include/xpore_b200.h"
#include "xp_kernels.cuh"

namespace {

thread_local std::string g_err;
thread_local uint64_t g_launches = 0;

// optional per-stage CUDA-event timing of xpb200_fit_device (bench.py's roofline numbers)
struct StageTimer {
    bool on = false, valid = false;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr};
};
thread_local StageTimer g_timer;
void stage_mark(int i, cudaStream_t st) {
    if (g_timer.on && g_timer.ev[i]) cudaEventRecord(g_timer.ev[i], st);
}

int fail(int code, const std::string& msg) {
    g_err = msg;
    return code;
}
#define CK(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(XPB200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_));         \
    } while (0)

constexpr size_t kMaxSmemPerCta = 227 * 1024;
// Positions with at least this many reads are fitted by 4-warp teams, the rest by single warps, so
// a few deep positions do not dictate the slab size (and with it the resident warps) of all.
// Must be a boundary of xpk::cost_bucket (a power of two is).
constexpr int kSplitReads = 1024;

struct Workspace {  // carved out of the caller's buffer
    uint32_t* hist;
    uint32_t* cursor;
    uint32_t* n_work;
    unsigned int* ticket;
    uint32_t* n_large;
    unsigned int* ticket_large;
    uint32_t* n_deep;
    unsigned int* ticket_deep;
    uint32_t* pack_counts;  // [kPackWarps] scratch of xpb200_pack_rows
    int32_t* worklist;
};
constexpr int kPackWarps = 2048;  // warps of the row-compaction launches of one batch / chunk
size_t ws_header_bytes() { return (size_t)(2 * xpk::kBuckets + 64) * 4; }  // the part zeroed before every fit
size_t ws_fixed_bytes() { return ws_header_bytes() + (size_t)kPackWarps * 4; }
Workspace carve(void* ws) {
    Workspace w;
    uint32_t* u = reinterpret_cast<uint32_t*>(ws);
    w.hist = u;
    w.cursor = u + xpk::kBuckets;
    w.n_work = u + 2 * xpk::kBuckets;
    w.ticket = u + 2 * xpk::kBuckets + 1;
    w.n_large = u + 2 * xpk::kBuckets + 2;
    w.ticket_large = u + 2 * xpk::kBuckets + 3;
    w.n_deep = u + 2 * xpk::kBuckets + 4;
    w.ticket_deep = u + 2 * xpk::kBuckets + 5;
    w.pack_counts = u + 2 * xpk::kBuckets + 64;
    w.worklist = reinterpret_cast<int32_t*>(u + 2 * xpk::kBuckets + 64 + kPackWarps);
    return w;
}

int check_batch(const xpb200_batch* b, const xpb200_method* m) {
    if (!b || !m) return fail(XPB200_EINVAL, "null batch/method");
    if (b->n_positions < 0) return fail(XPB200_EINVAL, "n_positions < 0");
    if (b->n_groups < 1 || b->n_groups > xpk::kMaxGroups) return fail(XPB200_EINVAL, "n_groups must be 1..64");
    if (b->n_conditions < 1 || b->n_conditions > xpk::kMaxConditions)
        return fail(XPB200_EINVAL, "n_conditions must be 1..32");
    if (!(m->dirichlet_conc > 0.0)) return fail(XPB200_EINVAL, "dirichlet_conc must be > 0");
    if (b->y_format < 0 || b->y_format > 2) return fail(XPB200_EINVAL, "unknown y_format");
    if (b->y_format != XPB200_Y_F64 && !(b->y_scale > 0.0)) return fail(XPB200_EINVAL, "y_scale must be > 0");
    return 0;
}

// Launch geometry of one fit-kernel flavour: `tw` warps per team (position), `teams` teams per CTA.
struct Geometry {
    int tw, teams, threads, maxt, cap, ctas_per_sm, sms, regs;
    int np;  // > 0: the tensor-memory kernel of xp_fit_tm.cuh with np positions per warp; 0 = the kernels of xp_fit.cuh
    size_t smem;
    bool stream;
};

constexpr int kStreamWarps = 16;  // warps of a streaming team (one CTA, one position)

int exp_int(const char* name, int dflt);

// Shared-memory opt-in, occupancy and register count of a kernel at a launch shape, remembered per device: the chunked
// host calls ask for the same few shapes once per chunk, and the three runtime calls cost the enqueueing thread tens of
// microseconds each time - at the start of a call that is time the SMs spend waiting for their first launches.  (The
// opt-in is set to the full 227 KB once per kernel, so a later, larger shape of the same kernel needs no second call.)
struct PrepKey {
    int dev;
    const void* fn;
    int threads;
    size_t smem;
    bool operator<(const PrepKey& o) const {
        if (dev != o.dev) return dev < o.dev;
        if (fn != o.fn) return fn < o.fn;
        if (threads != o.threads) return threads < o.threads;
        return smem < o.smem;
    }
};
std::mutex g_prep_mu;
std::map<PrepKey, std::pair<int, int>> g_prep;      // -> {CTAs per SM, registers per thread}
std::map<std::pair<int, const void*>, bool> g_optin;  // kernels whose dynamic shared memory limit is raised
template <class K>
int prepare_kernel_fn(K* k, Geometry* g) {
    int dev = 0;
    CK(cudaGetDevice(&dev));
    const PrepKey key{dev, reinterpret_cast<const void*>(k), g->threads, g->smem};
    std::lock_guard<std::mutex> lk(g_prep_mu);
    auto it = g_prep.find(key);
    if (it != g_prep.end()) {
        g->ctas_per_sm = it->second.first;
        g->regs = it->second.second;
        return 0;
    }
    bool& opted = g_optin[std::make_pair(dev, key.fn)];
    if (!opted) {
        CK(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmemPerCta));
        opted = true;
    }
    CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&g->ctas_per_sm, k, g->threads, g->smem));
    cudaFuncAttributes fa;
    CK(cudaFuncGetAttributes(&fa, k));
    g->regs = fa.numRegs;
    g_prep[key] = std::make_pair(g->ctas_per_sm, g->regs);
    return 0;
}
template <int TW, int MAXT, bool STREAM = false>
int prepare_kernel(Geometry* g) {
    return prepare_kernel_fn(xpk::xp_fit_kernel<TW, MAXT, STREAM>, g);
}
// the tensor-memory flavour of the single-warp class (u16 batches), by positions per warp and thread limit
using FitKernelFn = void (*)(xpk::FitArgs, xpk::FitChain);
const xpk::FitChain& no_chain() {
    static const xpk::FitChain z{};
    return z;
}
template <int NPW, int GT>
FitKernelFn tm_kernel_g(int maxt) {
    switch (maxt) {
        case 640: return xpk::xp_fit_tm_kernel<NPW, 640, GT>;
        case 768: return xpk::xp_fit_tm_kernel<NPW, 768, GT>;
        case 896: return xpk::xp_fit_tm_kernel<NPW, 896, GT>;
        default: return xpk::xp_fit_tm_kernel<NPW, 1024, GT>;
    }
}
// two positions per warp always run 28 warps (their warp regions are small); the designs of the named workloads
// (G = 2 pooled two-condition, 4 = 2 x 2 runs, 6 = 2 x 3 runs) have the group count compiled in
FitKernelFn tm_kernel(int npw, int maxt, int G) {
    if (npw == 2) {
        if (maxt == 896 && exp_int("XPB200_TM_GT", 1)) {
            if (G == 2) return xpk::xp_fit_tm_kernel<2, 896, 2>;
            if (G == 4) return xpk::xp_fit_tm_kernel<2, 896, 4>;
            if (G == 6) return xpk::xp_fit_tm_kernel<2, 896, 6>;
        }
        return tm_kernel_g<2, 0>(maxt);
    }
    if (exp_int("XPB200_TM_GT", 1)) {  // 6 x 3 runs (c5: 19 warps), 2 x 2 runs of a deep batch (c4: 20 warps)
        if (G == 18 && maxt == 640) return xpk::xp_fit_tm_kernel<1, 640, 18>;
        if (G == 4 && maxt == 640) return xpk::xp_fit_tm_kernel<1, 640, 4>;
    }
    return tm_kernel_g<1, 0>(maxt);
}

// The same kernels following the chain of chunk worklists of the host pipeline (xpk::FitChain); null where no such
// instantiation exists (the pipeline then launches chunk by chunk).
FitKernelFn tm_chain_kernel(int npw, int maxt, int G) {
    if (npw == 2) {
        if (maxt != 896) return nullptr;
        if (G == 2) return xpk::xp_fit_tm_kernel<2, 896, 2, true>;
        if (G == 4) return xpk::xp_fit_tm_kernel<2, 896, 4, true>;
        if (G == 6) return xpk::xp_fit_tm_kernel<2, 896, 6, true>;
        return xpk::xp_fit_tm_kernel<2, 896, 0, true>;
    }
    if (G == 18 && maxt == 640) return xpk::xp_fit_tm_kernel<1, 640, 18, true>;
    switch (maxt) {
        case 640: return xpk::xp_fit_tm_kernel<1, 640, 0, true>;
        case 768: return xpk::xp_fit_tm_kernel<1, 768, 0, true>;
        case 896: return xpk::xp_fit_tm_kernel<1, 896, 0, true>;
        default: return xpk::xp_fit_tm_kernel<1, 1024, 0, true>;
    }
}

constexpr int kFitWarpsDefault = 28;

// Tuning knobs read from the environment exist only in experiment builds (-DXPB200_EXPERIMENTS); the shipped
// library always runs its defaults.
int exp_int(const char* name, int dflt) {
#ifdef XPB200_EXPERIMENTS
    const char* v = getenv(name);
    return (v && *v) ? atoi(v) : dflt;
#else
    (void)name;
    return dflt;
#endif
}

// Teams per CTA: as many as 227 KB of shared memory (after the 4 KB of tables) allows, up to 28
// single-warp teams.  The kernel is instantiated per thread limit
// (640 / 768 / 896 / 1024 threads => 96 / 80 / 72 / 64 registers per thread) so that ptxas knows its budget.
// tw = kStreamWarps: the streaming flavour (reads stay in global memory; no cap on the reads of a position).
int fit_geometry(int G, int cap_reads, int tw, Geometry* g) {
    int dev = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&g->sms, cudaDevAttrMultiProcessorCount, dev));
    g->tw = tw;
    g->np = 0;
    g->stream = (tw == kStreamWarps);
    g->cap = std::max(32, (cap_reads + 1 + 7) & ~7);
    const size_t per_team = xpk::fit_team_smem_bytes(g->cap, G, tw, g->stream);
    if (xpk::kFitTableBytes + per_team > kMaxSmemPerCta)
        return fail(XPB200_EINVAL, "internal: a resident-class slab for " + std::to_string(cap_reads) +
                                       " reads exceeds one CTA's shared memory (the streaming class takes those)");
    int teams = (int)((kMaxSmemPerCta - xpk::kFitTableBytes) / per_team);
    if (g->stream) {
        teams = 1;
        g->maxt = kStreamWarps * 32;
    } else if (tw == 1) {
        teams = std::min(teams, std::max(1, exp_int("XPB200_FIT_WARPS", kFitWarpsDefault)));
        teams = std::min(teams, 32);
        g->maxt = teams <= 20 ? 640 : teams <= 24 ? 768 : teams <= 28 ? 896 : 1024;
    } else {
        teams = std::min(teams, 4);
        g->maxt = 512;
    }
    g->teams = teams;
    g->threads = teams * tw * 32;
    g->smem = xpk::kFitTableBytes + per_team * teams;
    int rc;
    if (g->stream) rc = prepare_kernel<kStreamWarps, kStreamWarps * 32, true>(g);
    else if (tw == 1) {
        switch (g->maxt) {
            case 640: rc = prepare_kernel<1, 640>(g); break;
            case 768: rc = prepare_kernel<1, 768>(g); break;
            case 896: rc = prepare_kernel<1, 896>(g); break;
            default: rc = prepare_kernel<1, 1024>(g); break;
        }
    } else rc = prepare_kernel<4, 512>(g);
    if (rc) return rc;
    if (g->ctas_per_sm < 1) return fail(XPB200_ECUDA, "fit kernel does not fit on an SM");
    return 0;
}

// Positions per warp of the tensor-memory kernel: two when their 2G + 4 scalar-phase slots fit half a warp each - unless
// the batch is a deep one (positions for the 4-warp teams anyway): two positions per warp hold 576 reads each, and
// handing the teams everything from there up costs more than the shared scalar phase saves (c4: 8.3 against 7.6 ms).
int tm_positions_per_warp(int G, int max_reads) {
    return exp_int("XPB200_TM_NPW", (2 * G + 4 <= 16 && max_reads < kSplitReads) ? 2 : 1);
}
// First read count of the 4-warp teams' class when the single-warp class runs on the tensor-memory kernel (a bucket
// boundary of the worklist).  Two positions per warp: 576 (2 x 36 columns x 7 column blocks per quadrant = 504 of 512
// columns keep 28 warps resident).  One position per warp: 1536 - a warp is the more efficient unit for a position
// well beyond the former 1024 (no warps idling through a leader's scalar phase), although 96 columns per position leave
// only 5 column blocks per quadrant = 20 warps (c4: 7.62 ms at 1024 / 28 warps, 7.47 at 1152 / 28, 7.37 at 1536 / 20,
// 7.46 at 2048 / 16).
int tm_class_split(int npw) { return exp_int("XPB200_TM_SPLIT", npw == 2 ? 576 : 1536); }
// Launch geometry of xp_fit_tm_kernel: one warp per `np` positions, reads in tensor memory.
int fit_tm_geometry(int G, int cap_reads, int npw, Geometry* g) {
    int dev = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&g->sms, cudaDevAttrMultiProcessorCount, dev));
    g->tw = 1;
    g->np = npw;
    g->stream = false;
    g->cap = std::max(32, (cap_reads + 31) & ~31);
    const size_t per_warp = xpk::fit_tm_warp_bytes(g->cap, G, npw);
    int warps = (int)((kMaxSmemPerCta - xpk::kTmTableBytes) / per_warp);
    warps = std::min(warps, std::max(1, exp_int("XPB200_FIT_WARPS", kFitWarpsDefault)));
    warps = std::min(warps, 32);
    // tensor-memory columns: ceil(warps / 4) column blocks of npw positions each per quadrant
    const int blocks = xpk::kTmColumns / (npw * xpk::fit_tm_cols(g->cap));
    warps = std::min(warps, 4 * blocks);
    if (warps < 1) return fail(XPB200_EINVAL, "internal: tensor-memory class with a position too deep for it");
    g->maxt = warps <= 20 ? 640 : warps <= 24 ? 768 : warps <= 28 ? 896 : 1024;
    g->teams = warps * npw;  // positions in flight per CTA
    g->threads = warps * 32;
    g->smem = xpk::kTmTableBytes + per_warp * warps;
    if (int rc = prepare_kernel_fn(tm_kernel(npw, g->maxt, G), g)) return rc;
    if (g->ctas_per_sm < 1) return fail(XPB200_ECUDA, "fit kernel does not fit on an SM");
    return 0;
}

// The class launches of one fit (4-warp teams, full-slab single-warp teams, small-slab single-warp teams) work on
// disjoint positions.  On one stream each would wait for the stragglers of the one before; on side streams forked
// from and joined back into the caller's stream their CTAs take over SMs as the earlier launch's CTAs leave.
struct SideStreams {
    cudaStream_t s[2] = {nullptr, nullptr};
    cudaEvent_t fork = nullptr, join[2] = {nullptr, nullptr};
};
std::mutex g_side_mu;
std::map<std::pair<int, cudaStream_t>, SideStreams*> g_side;
int side_streams(cudaStream_t st, SideStreams** out) {
    int dev = 0;
    CK(cudaGetDevice(&dev));
    std::lock_guard<std::mutex> lk(g_side_mu);
    auto key = std::make_pair(dev, st);
    auto it = g_side.find(key);
    if (it == g_side.end()) {
        auto* ss = new SideStreams();
        for (int i = 0; i < 2; ++i) {
            CK(cudaStreamCreateWithFlags(&ss->s[i], cudaStreamNonBlocking));
            CK(cudaEventCreateWithFlags(&ss->join[i], cudaEventDisableTiming));
        }
        CK(cudaEventCreateWithFlags(&ss->fork, cudaEventDisableTiming));
        it = g_side.emplace(key, ss).first;
    }
    *out = it->second;
    return 0;
}

// Positions with at least this many reads go to the streaming class: the smallest bucket boundary S of the worklist
// (xp_kernels.cuh:cost_bucket, 16 buckets per octave) such that a 4-warp team's slab cannot hold S reads any more,
// i.e. the largest S whose predecessor S - 1 still fits one CTA.
int deep_class_split(int G) {
    int best = 0;
    for (int e = 10; e <= 17; ++e)
        for (int k = 0; k < 16; ++k) {
            const int S = (1 << e) + k * (1 << (e - 4));
            const int cap = std::max(32, (S - 1 + 1 + 7) & ~7);
            if (xpk::kFitTableBytes + xpk::fit_team_smem_bytes(cap, G, 4) <= kMaxSmemPerCta) best = S;
        }
    return best;  // > kSplitReads for every G <= kMaxGroups
}

// Chained form of the host pipeline: the class launches of a chunk run beside the chained launch, which holds all but a
// few SMs until the end of the batch - a grid larger than those few SMs would keep its stream busy until then, and with
// it the select phases of later chunks the chained launch is waiting for.
thread_local int g_fit_grid_cap = 0;
int launch_fit(const Geometry& g, const xpk::FitArgs& fa_in, int max_teams_needed, cudaStream_t st) {
    xpk::FitArgs fa = fa_in;
    if (g.np > 0) xpk::fit_tm_fill_layout(fa, g.np);
    else xpk::fit_fill_layout(fa, g.tw, g.stream);
    const int need = (max_teams_needed + g.teams - 1) / g.teams;
    int grid = std::max(1, std::min(g.sms * g.ctas_per_sm, need));
    if (g_fit_grid_cap > 0) grid = std::min(grid, g_fit_grid_cap);
    if (g.np > 0) tm_kernel(g.np, g.maxt, fa.G)<<<grid, g.threads, g.smem, st>>>(fa, no_chain());
    else if (g.stream) xpk::xp_fit_kernel<kStreamWarps, kStreamWarps * 32, true><<<grid, g.threads, g.smem, st>>>(fa);
    else if (g.tw == 1) {
        switch (g.maxt) {
            case 640: xpk::xp_fit_kernel<1, 640><<<grid, g.threads, g.smem, st>>>(fa); break;
            case 768: xpk::xp_fit_kernel<1, 768><<<grid, g.threads, g.smem, st>>>(fa); break;
            case 896: xpk::xp_fit_kernel<1, 896><<<grid, g.threads, g.smem, st>>>(fa); break;
            default: xpk::xp_fit_kernel<1, 1024><<<grid, g.threads, g.smem, st>>>(fa); break;
        }
    } else xpk::xp_fit_kernel<4, 512><<<grid, g.threads, g.smem, st>>>(fa);
    ++g_launches;
    CK(cudaGetLastError());
    return 0;
}

int launch_prefilter(const xpb200_batch* b, const xpb200_method* m, uint32_t* status, double* pval,
                     cudaStream_t st) {
    if (b->n_positions == 0) return 0;
    xpk::PrefilterArgs a;
    a.y = b->y; a.y_format = b->y_format; a.y_scale = b->y_scale;
    a.read_off = b->read_off; a.grp_len = b->grp_len; a.kmer_mean = b->kmer_mean;
    a.grp_cond = b->grp_cond; a.P = b->n_positions; a.G = b->n_groups;
    a.enabled = (m->prefilter == 1);
    a.threshold = m->prefilter_threshold;
    a.status = status; a.pval = pval;
    constexpr int W = 4;
    const int tiles = (b->n_positions + 31) / 32;
    const int grid = (tiles + W - 1) / W;
    if (b->y_format == XPB200_Y_U16 && b->max_reads <= xpk::kPfIntMaxReads) {
        // exact integer sums.  Group slots blocked by condition (every batch builder's order): the first kernel; slots
        // interleaved by condition: the second.  Each returns at once in the other's case - grp_cond is device memory
        // and this entry point does not synchronise, so the choice is made on the device.
        xpk::xp_prefilter_u16_kernel<W><<<grid, W * 32, 0, st>>>(a);
        xpk::xp_prefilter_kernel<W, true><<<grid, W * 32, 0, st>>>(a);
        ++g_launches;
    } else
        xpk::xp_prefilter_kernel<W, false><<<grid, W * 32, 0, st>>>(a);
    ++g_launches;
    CK(cudaGetLastError());
    return 0;
}

int launch_stats(const xpb200_batch* b, const xpb200_method* m, const xpb200_result* r, const int32_t* worklist,
                 const uint32_t* n_work, cudaStream_t st) {
    if (b->n_positions == 0) return 0;
    xpk::StatsArgs a;
    a.worklist = worklist; a.n_work = n_work; a.P = b->n_positions; a.G = b->n_groups; a.C = b->n_conditions;
    a.grp_cond = b->grp_cond; a.grp_len = b->grp_len; a.kmer_mean = b->kmer_mean; a.kmer_std = b->kmer_std;
    a.conc0 = m->dirichlet_conc; a.fit = r->fit; a.stats = r->stats; a.status = r->status;
    const int threads = 128;
    xpk::xp_stats_kernel<<<(b->n_positions + threads - 1) / threads, threads, 0, st>>>(a);
    ++g_launches;
    CK(cudaGetLastError());
    return 0;
}

}  // namespace

extern "C" {

const char* xpb200_last_error(void) { return g_err.c_str(); }
const char* xpb200_version(void) { return "xpore_b200 0.1 (sm_100a)"; }
uint64_t xpb200_launch_count(void) { return g_launches; }
int32_t xpb200_fit_width(int32_t G) { return xpc::fit_width(G); }
int32_t xpb200_stats_width(int32_t G, int32_t C) { return xpc::stats_width(G, C); }
size_t xpb200_workspace_bytes(int32_t P) { return ws_fixed_bytes() + (size_t)std::max(P, 1) * 4; }

void* xpb200_host_alloc(size_t bytes) {
    void* p = nullptr;
    if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocPortable) != cudaSuccess) {
        g_err = std::string("cudaHostAlloc: ") + cudaGetErrorString(cudaGetLastError());
        return nullptr;
    }
    return p;
}
void xpb200_host_free(void* p) {
    if (p) cudaFreeHost(p);
}
int32_t xpb200_set_device(int32_t device) {
    CK(cudaSetDevice(device));
    return 0;
}

int32_t xpb200_deep_split(int32_t G) {
    if (G < 1 || G > xpk::kMaxGroups) return fail(XPB200_EINVAL, "n_groups must be 1..64");
    return deep_class_split(G);
}

int32_t xpb200_fit_launch_info(int32_t device, int32_t G, int32_t max_reads, int32_t y_format, xpb200_launch_info* out) {
    if (!out) return fail(XPB200_EINVAL, "null out");
    if (G < 1 || G > xpk::kMaxGroups) return fail(XPB200_EINVAL, "n_groups must be 1..64");
    CK(cudaSetDevice(device));
    Geometry g;  // the class below the team split
    if (y_format == XPB200_Y_U16) {
        const int npw = tm_positions_per_warp(G, max_reads);
        if (int rc = fit_tm_geometry(G, std::min(max_reads, tm_class_split(npw) - 1), npw, &g)) return rc;
        out->warps_per_cta = g.threads / 32;
        out->positions_per_warp = npw;
        out->tmem_columns = xpk::fit_tm_cols(g.cap);
    } else {
        if (int rc = fit_geometry(G, std::min(max_reads, kSplitReads - 1), 1, &g)) return rc;
        out->warps_per_cta = g.teams;
        out->positions_per_warp = 0;
        out->tmem_columns = 0;
    }
    out->sm_count = g.sms; out->ctas_per_sm = g.ctas_per_sm;
    out->smem_per_cta = (int32_t)g.smem; out->slab_reads = g.cap; out->regs_per_thread = g.regs;
    return 0;
}

int32_t xpb200_prefilter_device(const xpb200_batch* b, const xpb200_method* m, uint32_t* status, double* pval,
                                void* stream) {
    if (int rc = check_batch(b, m)) return rc;
    return launch_prefilter(b, m, status, pval, (cudaStream_t)stream);
}

int32_t xpb200_stats_device(const xpb200_batch* b, const xpb200_method* m, const xpb200_result* r,
                            const int32_t* worklist, const uint32_t* n_work, void* stream) {
    if (int rc = check_batch(b, m)) return rc;
    return launch_stats(b, m, r, worklist, n_work, (cudaStream_t)stream);
}

namespace {
enum { kPhaseSelect = 1, kPhaseFit = 2, kPhaseStats = 4, kPhaseAll = 7, kPhaseFitLarge = 8 };  // FitLarge: the team / streaming class launches only
int fit_device_phases(const xpb200_batch* b, const xpb200_method* m, const xpb200_result* r, void* workspace,
                      size_t workspace_bytes, uint32_t* n_fitted_dev, void* stream, int phases);
}  // namespace

int32_t xpb200_fit_device(const xpb200_batch* b, const xpb200_method* m, const xpb200_result* r, void* workspace,
                          size_t workspace_bytes, uint32_t* n_fitted_dev, void* stream) {
    return fit_device_phases(b, m, r, workspace, workspace_bytes, n_fitted_dev, stream, kPhaseAll);
}

namespace {
// The phases of one fit: select (prefilter + worklist), fit (the class launches), stats.  The chunked host call, when
// ONE chained launch fits all chunks (fit_host_impl), enqueues a chunk's select phase when its reads have arrived and its
// stats phase when the chained launch is done.
int fit_device_phases(const xpb200_batch* b, const xpb200_method* m, const xpb200_result* r, void* workspace,
                      size_t workspace_bytes, uint32_t* n_fitted_dev, void* stream, int phases) {
    if (int rc = check_batch(b, m)) return rc;
    if (!r || !r->status || !r->n_iter || !r->pval || !r->fit || !r->stats)
        return fail(XPB200_EINVAL, "null result array");
    const int P = b->n_positions;
    cudaStream_t st = (cudaStream_t)stream;
    if (P == 0) {
        if (n_fitted_dev) CK(cudaMemsetAsync(n_fitted_dev, 0, 4, st));
        return 0;
    }
    if (!workspace || workspace_bytes < xpb200_workspace_bytes(P)) return fail(XPB200_ENOMEM, "workspace too small");
    Workspace ws = carve(workspace);
    const bool do_select = (phases & kPhaseSelect) != 0, do_fit = (phases & (kPhaseFit | kPhaseFitLarge)) != 0;
    const bool do_fit_small = (phases & kPhaseFit) != 0;
    if (do_select) CK(cudaMemsetAsync(workspace, 0, ws_header_bytes(), st));
    stage_mark(0, st);

    // K1 prefilter (or: everything selected, p-value NaN)
    const uint32_t* sel = nullptr;
    if (m->prefilter) {
        if (do_select)
            if (int rc = launch_prefilter(b, m, r->status, r->pval, st)) return rc;
        sel = r->status;
    } else if (do_select) {
        CK(cudaMemsetAsync(r->status, 0, (size_t)P * 4, st));
        CK(cudaMemsetAsync(r->pval, 0xFF, (size_t)P * 8, st));  // all-ones = NaN
    }
    stage_mark(1, st);
    // worklist, longest positions first
    xpk::WorklistArgs wa;
    wa.status = sel; wa.read_off = b->read_off; wa.P = P; wa.hist = ws.hist; wa.cursor = ws.cursor;
    wa.n_work = ws.n_work; wa.worklist = ws.worklist; wa.n_large = ws.n_large;
    wa.pval = (m->prefilter && exp_int("XPB200_LPT", 1)) ? r->pval : nullptr;
    // Three classes by depth, each with its own launch: n >= deep_split reads -> streaming CTAs (reads stay in global
    // memory: no cap), n >= kSplitReads -> 4-warp teams, the rest -> one warp per position.
    // u16 batches: the single-warp class runs on the tensor-memory kernel, which holds fewer reads per position when
    // two positions share a warp
    const int npw = tm_positions_per_warp(b->n_groups, b->max_reads);
    const bool use_tm = b->y_format == XPB200_Y_U16 && exp_int("XPB200_TM", 1) != 0;
    const int split1 = use_tm ? tm_class_split(npw) : kSplitReads;
    const int deep_split = deep_class_split(b->n_groups);
    const bool has_deep = b->max_reads >= deep_split;
    const bool two_class = b->max_reads >= split1;
    wa.split = two_class ? split1 : 0;
    wa.n_deep = ws.n_deep;
    wa.split_deep = has_deep ? deep_split : 0;
    const int tw1_reads = two_class ? split1 - 1 : b->max_reads;
    Geometry g;
    if (use_tm) {
        if (int rc = fit_tm_geometry(b->n_groups, tw1_reads, npw, &g)) return rc;
    } else if (int rc = fit_geometry(b->n_groups, tw1_reads, 1, &g)) return rc;
    if (do_select) {
        const int wthreads = 256;
        const int wblocks = std::min((P + wthreads - 1) / wthreads, 1184);
        xpk::xp_worklist_count<<<wblocks, wthreads, 0, st>>>(wa);
        xpk::xp_worklist_scan<<<1, xpk::kBuckets, 0, st>>>(wa);
        xpk::xp_worklist_scatter<<<wblocks, wthreads, 0, st>>>(wa);
        g_launches += 3;
        CK(cudaGetLastError());
    }

    stage_mark(2, st);
    if (do_fit) {
    // K3 fit: persistent CTAs; the deepest positions first
    xpk::FitArgs fa;
    fa.y = b->y; fa.y_format = b->y_format; fa.y_scale = b->y_scale;
    fa.read_off = b->read_off; fa.grp_len = b->grp_len; fa.kmer_mean = b->kmer_mean;
    fa.kmer_std = b->kmer_std; fa.worklist = ws.worklist;
    fa.G = b->n_groups; fa.max_iters = m->max_iters; fa.stopping = m->stopping_criteria;
    fa.conc0 = m->dirichlet_conc;
    fa.lg_prior_w = xpm::xlgamma(2.0 * m->dirichlet_conc) - 2.0 * xpm::xlgamma(m->dirichlet_conc);
    fa.fit = r->fit; fa.n_iter = r->n_iter; fa.status = r->status; fa.coef = r->coef;
    // class launches, longest positions first: the first on the caller's stream, the others on side streams
    SideStreams* ss = nullptr;
    const int n_class = 1 + (two_class ? 1 : 0) + (has_deep ? 1 : 0);
    if (n_class > 1) {
        if (int rc = side_streams(st, &ss)) return rc;
        CK(cudaEventRecord(ss->fork, st));
    }
    int k = 0;  // launches so far
    auto class_stream = [&](int i) -> cudaStream_t { return i == 0 ? st : ss->s[i - 1]; };
    auto begin_class = [&](int i) -> int {
        if (i > 0) CK(cudaStreamWaitEvent(ss->s[i - 1], ss->fork, 0));
        return 0;
    };
    auto end_class = [&](int i) -> int {
        if (i > 0) {
            CK(cudaEventRecord(ss->join[i - 1], ss->s[i - 1]));
            CK(cudaStreamWaitEvent(st, ss->join[i - 1], 0));
        }
        return 0;
    };
    if (has_deep) {
        Geometry gd;
        if (int rc = fit_geometry(b->n_groups, 0, kStreamWarps, &gd)) return rc;
        fa.cap = 0; fa.w_begin = nullptr; fa.w_end = ws.n_deep; fa.ticket = ws.ticket_deep;
        if (int rc = begin_class(k)) return rc;
        if (int rc = launch_fit(gd, fa, P, class_stream(k))) return rc;
        if (int rc = end_class(k)) return rc;
        ++k;
    }
    if (two_class) {
        Geometry gl;
        if (int rc = fit_geometry(b->n_groups, std::min(b->max_reads, deep_split - 1), 4, &gl)) return rc;
        fa.cap = gl.cap; fa.w_begin = has_deep ? ws.n_deep : nullptr; fa.w_end = ws.n_large; fa.ticket = ws.ticket_large;
        if (int rc = begin_class(k)) return rc;
        if (int rc = launch_fit(gl, fa, P, class_stream(k))) return rc;
        if (int rc = end_class(k)) return rc;
        ++k;
    }
    if (do_fit_small) {
        fa.cap = g.cap; fa.w_begin = two_class ? ws.n_large : nullptr; fa.w_end = ws.n_work; fa.ticket = ws.ticket;
        if (int rc = begin_class(k)) return rc;
        if (int rc = launch_fit(g, fa, P, class_stream(k))) return rc;
        if (int rc = end_class(k)) return rc;
        ++k;
    }
    }  // do_fit

    stage_mark(3, st);
    // K4 statistics
    if (phases & kPhaseStats)
        if (int rc = launch_stats(b, m, r, ws.worklist, ws.n_work, st)) return rc;
    stage_mark(4, st);
    g_timer.valid = g_timer.on && phases == kPhaseAll;
    if (n_fitted_dev && do_select) CK(cudaMemcpyAsync(n_fitted_dev, ws.n_work, 4, cudaMemcpyDeviceToDevice, st));
    return 0;
}
}  // namespace

// ---- host-buffer entry points -------------------------------------------------------------
// The batch is cut into chunks of positions (~kChunkBytes of reads each).  Chunk c's reads go up
// on the copy-in stream while chunk c-1 is being fitted on a compute stream and chunk c-2's
// records come back on the copy-out stream, so PCIe (full duplex) and the SMs work concurrently.
namespace {
constexpr int kMaxChunks = xpk::kMaxRowChunks;
constexpr size_t kChunkBytes = 112u << 20;
constexpr int kChunkGrowthPct = 150;

constexpr int kComputeStreams = 3;

struct DevCache {
    int device = -1;
    cudaStream_t s_in = nullptr, s_c[kComputeStreams] = {nullptr, nullptr, nullptr}, s_out = nullptr;
    cudaEvent_t e_in[kMaxChunks], e_done[kMaxChunks];
    cudaStream_t s_fit = nullptr, s_help = nullptr;  // the chained launch and its late helper (see fit_host_impl)
    cudaStream_t s_post[kComputeStreams] = {nullptr, nullptr, nullptr}, s_poll = nullptr;  // chained form: statistics / rows of finished chunks, progress polls
    cudaEvent_t e_cls[kMaxChunks];                   // a chunk's select phase and class launches are enqueued / done
    uint32_t* h_poll = nullptr;                      // page-locked: the flag words, then n_fitted and the row count of every chunk
    bool chain_off = false;  // a chained launch timed out waiting for a chunk (kernels serialised by a profiler, SMs held by
                             // somebody else): this process stays with chunk-by-chunk launches on this device
    cudaEvent_t e_sel[kComputeStreams] = {nullptr, nullptr, nullptr}, e_help = nullptr;
    cudaEvent_t e_zero = nullptr, e_fit = nullptr;   // its flags are cleared / it is done
    cudaEvent_t e_call = nullptr;  // end of the previous call's device work (the arena is reused by the next one)
    cudaEvent_t e_user = nullptr;  // the caller's stream at entry of an asynchronous call
    cudaEvent_t t_0 = nullptr, t_in[kMaxChunks], t_start[kMaxChunks], t_done[kMaxChunks];  // XPB200_TRACE=1 only
    cudaEvent_t t_fit = nullptr, t_post = nullptr;
    void* buf = nullptr;
    size_t cap = 0;
    long calls = 0;
};
std::recursive_mutex g_cache_mu;  // recursive: a chained call whose max_reads was understated starts over chunk by chunk
std::vector<DevCache*> g_cache;

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

// Output mode of fit_host_impl
struct HostOut {
    const xpb200_result* dense = nullptr;  // dense result arrays of every position (host or device memory)
    double* rows = nullptr;                // host: records of the table's rows, compacted
    int64_t rows_capacity = 0;
    int64_t* n_rows = nullptr;
    uint32_t* n_fitted = nullptr;
    double* rows_dev = nullptr;            // device: the same records, no host synchronisation
    uint32_t* counts_dev = nullptr;        // device [2]: rows needed, positions fitted
    int64_t index_offset = 0;
    cudaStream_t user_stream = nullptr;
};

int launch_pack_rows(const xpb200_result& dr, int pc, int FW, int SW, long long index_offset, long long capacity,
                     unsigned* counts, unsigned* n_rows_dev, double* out, cudaStream_t st) {
    xpk::PackRowsArgs pa;
    pa.status = dr.status; pa.n_iter = dr.n_iter; pa.pval = dr.pval; pa.fit = dr.fit; pa.stats = dr.stats;
    pa.P = pc; pa.FW = FW; pa.SW = SW; pa.need = xpc::STATUS_SELECTED | xpc::STATUS_KEEP_ROW;
    pa.index_offset = index_offset; pa.capacity = capacity; pa.n_rows = n_rows_dev; pa.out = out;
    const int threads = 256, wpb = threads / 32;
    pa.n_warps = std::max(1, std::min(kPackWarps, (pc + 255) / 256));
    pa.span = ((pc + pa.n_warps - 1) / pa.n_warps + 31) & ~31;
    pa.counts = counts;
    const int blocks = (pa.n_warps + wpb - 1) / wpb;
    xpk::xp_count_rows_kernel<<<blocks, threads, 0, st>>>(pa);
    xpk::xp_pack_rows_kernel<<<blocks, threads, 0, st>>>(pa);
    g_launches += 2;
    CK(cudaGetLastError());
    return 0;
}

int fit_host_impl(int32_t device, const xpb200_batch* b, const xpb200_method* m, const HostOut& ho, bool allow_chain = true);
}  // namespace

int32_t xpb200_fit_host(int32_t device, const xpb200_batch* b, const xpb200_method* m, const xpb200_result* r,
                        uint32_t* n_fitted) {
    if (int rc = check_batch(b, m)) return rc;
    if (!r || !r->status || !r->n_iter || !r->pval || !r->fit || !r->stats)
        return fail(XPB200_EINVAL, "null result array");
    HostOut ho;
    ho.dense = r;
    ho.n_fitted = n_fitted;
    return fit_host_impl(device, b, m, ho);
}

int32_t xpb200_fit_host_rows(int32_t device, const xpb200_batch* b, const xpb200_method* m, double* rows,
                             int64_t capacity_rows, int64_t* n_rows, uint32_t* n_fitted) {
    if (int rc = check_batch(b, m)) return rc;
    if (!rows || capacity_rows < 0 || !n_rows) return fail(XPB200_EINVAL, "null rows / n_rows");
    HostOut ho;
    ho.rows = rows;
    ho.rows_capacity = capacity_rows;
    ho.n_rows = n_rows;
    ho.n_fitted = n_fitted;
    return fit_host_impl(device, b, m, ho);
}

int32_t xpb200_fit_host_rows_async(int32_t device, const xpb200_batch* b, const xpb200_method* m, double* rows_dev,
                                   int64_t capacity_rows, int64_t index_offset, uint32_t* counts_dev, void* stream) {
    if (int rc = check_batch(b, m)) return rc;
    if (!rows_dev || capacity_rows < 0 || !counts_dev) return fail(XPB200_EINVAL, "null rows_dev / counts_dev");
    HostOut ho;
    ho.rows_dev = rows_dev;
    ho.rows_capacity = capacity_rows;
    ho.counts_dev = counts_dev;
    ho.index_offset = index_offset;
    ho.user_stream = (cudaStream_t)stream;
    return fit_host_impl(device, b, m, ho);
}

namespace {
// The chained form of the pipeline (rows out, synchronous call, u16 batch whose deepest position is below the team
// split, so that the whole batch is the tensor-memory kernel's single-warp class): ONE persistent fit launch on all but
// kChainFreeSms SMs walks the chunks' worklists in order (xpk::FitChain) while the chunks' reads arrive and their prefilter
// and worklist launches run on the SMs left free; statistics and the row compaction follow when it is done.  Chunk by
// chunk launches (the other form) end each chunk with a tail of half-empty CTAs and leave most SMs idle through the
// first, small chunks: c2 took 18.3 ms end to end against 14.1 ms resident, the copy alone 14.2 ms.
constexpr int kChainFreeSms = 6;
constexpr size_t kChainChunkBytes = 32u << 20;
constexpr unsigned kChainSpinLimit = 500000u;  // polls of a warp waiting for a chunk (4 us each: >= 2 s) before it gives up

// Chunk boundaries (positions) of a host call, cut where the read offsets cross the byte targets.
//   chunk by chunk: sizes grow by kChunkGrowthPct from an eighth of a full chunk (112 MB) - the copy is only ~1.2x faster
//     than the fit, so a chunk much larger than its predecessor would leave the SMs waiting for it to arrive; a remainder
//     of less than a quarter chunk joins the last one;
//   chained form: tails and small launches cost nothing there, so the chunks are small (32 MB; more for batches beyond
//     ~1.5 GB, the chain table holds kMaxChunks) - they only set the latency between the arrival of a read and its fit -
//     and double from a sixteenth of a chunk; the last ones halve again down to an eighth: what is left to do when the last
//     read has arrived is the last chunk plus the long positions of the late chunks, and a small chunk holds fewer of both.
std::vector<int> plan_chunks(const int64_t* roff, int P, size_t eb, bool chain) {
    const size_t total_bytes = (size_t)roff[P] * eb;
    const size_t chunk_cap = chain ? std::max(kChainChunkBytes, total_bytes / (size_t)(kMaxChunks - 16)) : kChunkBytes;
    const size_t chunk_bytes = (size_t)std::max(1, exp_int(chain ? "XPB200_CHAIN_CHUNK_MB" : "XPB200_CHUNK_MB", (int)(chunk_cap >> 20))) << 20;
    std::vector<int> cut(1, 0);
    std::vector<size_t> ends;  // byte targets of the chunk ends
    size_t pos = 0;
    size_t sz = std::max<size_t>(chunk_bytes / (chain ? (size_t)std::max(1, exp_int("XPB200_CHAIN_FIRST_DIV", 16)) : 8), 1u << 20);
    const double growth = chain ? 2.0 : std::max(1.0, exp_int("XPB200_CHUNK_GROWTH_PCT", kChunkGrowthPct) / 100.0);
    const int last_div = chain ? std::max(1, exp_int("XPB200_CHAIN_LAST_DIV", 8)) : 1;
    size_t tail_bytes = 0;
    for (int dv = 2; dv <= last_div; dv *= 2) tail_bytes += chunk_bytes / dv;
    if (tail_bytes * 3 > total_bytes) tail_bytes = 0;
    const size_t body_bytes = total_bytes - tail_bytes;
    while (pos + sz + sz / 4 < body_bytes && (int)ends.size() < kMaxChunks - 8) {
        pos += sz;
        ends.push_back(pos);
        sz = std::min(chunk_bytes, (size_t)((double)sz * growth));
    }
    if (tail_bytes) {
        pos = body_bytes;
        for (int dv = 2; dv <= last_div; dv *= 2) {
            ends.push_back(pos);
            pos += chunk_bytes / dv;
        }
    }
    for (size_t e : ends) {
        const int64_t target = (int64_t)(e / eb);
        int p = (int)(std::lower_bound(roff, roff + P + 1, target) - roff);
        p = std::min(std::max(p, cut.back()), P);
        if (p > cut.back() && p < P) cut.push_back(p);
    }
    cut.push_back(P);
    return cut;
}

// One call of a host entry point: what its two forms - chained and chunk by chunk - share (the chunks, the device arena).
struct HostCall {
    // chained form: ready[c], the abort flag, polls spent waiting (experiment builds), positions done, size of the team / streaming classes
    static constexpr int kFlagWords = 4 * kMaxChunks + 1;
    const int32_t device;
    const xpb200_batch* const b;
    const xpb200_method* const m;
    const HostOut& ho;
    DevCache* dc = nullptr;
    const xpb200_result* const r;  // dense results wanted (host or device memory), else null
    double* const rows;            // rows wanted in host memory, else null
    const bool want_rows, async;
    const int P, G, C, FW, SW;
    int RW = 0;
    const size_t eb;
    const int64_t* const roff;
    const size_t total_bytes;
    // the chained form, when it applies
    Geometry gch;
    FitKernelFn chain_fn = nullptr;
    int chain_npw = 0;
    bool chain = false, chain_two_class = false;
    int chain_free_sms = kChainFreeSms;
    // chunks, streams, tracing
    std::vector<int> cut;
    int nchunks = 0, n_streams = kComputeStreams;
    bool trace = false;
    std::chrono::steady_clock::time_point host_t0;
    // device arena: offsets and pointers
    size_t o_y = 0, o_off = 0, o_glen = 0, o_km = 0, o_ks = 0, o_gc = 0, o_status = 0, o_niter = 0, o_pval = 0, o_fit = 0, o_stats = 0,
           o_flags = 0, o_nf = 0, o_nr = 0, o_cnt = 0, o_rows = 0;
    std::vector<size_t> o_ws;
    char *d = nullptr, *d_y = nullptr;
    int64_t* d_off = nullptr;
    int32_t *d_glen = nullptr, *d_niter = nullptr;
    double *d_km = nullptr, *d_ks = nullptr, *d_pval = nullptr, *d_fit = nullptr, *d_stats = nullptr, *d_rows = nullptr;
    uint32_t *d_status = nullptr, *d_ready = nullptr, *d_abort = nullptr, *d_done = nullptr, *d_nl = nullptr, *d_nf = nullptr, *d_nr = nullptr,
             *d_cnt = nullptr;

    HostCall(int32_t device_, const xpb200_batch* b_, const xpb200_method* m_, const HostOut& ho_)
        : device(device_), b(b_), m(m_), ho(ho_), r(ho_.dense), rows(ho_.rows),
          want_rows(ho_.rows != nullptr || ho_.rows_dev != nullptr), async(ho_.rows_dev != nullptr),
          P(b_->n_positions), G(b_->n_groups), C(b_->n_conditions), FW(xpc::fit_width(b_->n_groups)),
          SW(xpc::stats_width(b_->n_groups, b_->n_conditions)),
          eb(b_->y_format == XPB200_Y_F64 ? 8 : (b_->y_format == XPB200_Y_I32 ? 4 : 2)), roff(b_->read_off),
          total_bytes((size_t)b_->n_reads * (b_->y_format == XPB200_Y_F64 ? 8 : (b_->y_format == XPB200_Y_I32 ? 4 : 2))) {}

    int open_cache();                    // the device's streams, events and arena (created at the first call)
    int decide_chain(bool allow_chain);  // chained form?
    int lay_out();                       // arena offsets, (re)allocation, pointers
    int begin();                         // ordering against the previous call / the caller's stream, the first copies
    int chained();
    int chunked();
};

int HostCall::open_cache() {
    for (auto* c : g_cache)
        if (c->device == device) dc = c;
    if (!dc) {
        dc = new DevCache();
        dc->device = device;
        CK(cudaStreamCreateWithFlags(&dc->s_in, cudaStreamNonBlocking));
        for (int i = 0; i < kComputeStreams; ++i) CK(cudaStreamCreateWithFlags(&dc->s_c[i], cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&dc->s_out, cudaStreamNonBlocking));
        for (int i = 0; i < kMaxChunks; ++i) {
            CK(cudaEventCreateWithFlags(&dc->e_in[i], cudaEventDisableTiming));
            CK(cudaEventCreateWithFlags(&dc->e_done[i], cudaEventDisableTiming));
        }
        CK(cudaEventCreateWithFlags(&dc->e_call, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&dc->e_user, cudaEventDisableTiming));
        CK(cudaStreamCreateWithFlags(&dc->s_fit, cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&dc->s_help, cudaStreamNonBlocking));
        for (int i = 0; i < kComputeStreams; ++i) CK(cudaStreamCreateWithFlags(&dc->s_post[i], cudaStreamNonBlocking));
        CK(cudaStreamCreateWithFlags(&dc->s_poll, cudaStreamNonBlocking));
        for (int i = 0; i < kMaxChunks; ++i) CK(cudaEventCreateWithFlags(&dc->e_cls[i], cudaEventDisableTiming));
        CK(cudaHostAlloc((void**)&dc->h_poll, (size_t)(6 * kMaxChunks + 8) * 4, cudaHostAllocPortable));
        for (int i = 0; i < kComputeStreams; ++i) CK(cudaEventCreateWithFlags(&dc->e_sel[i], cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&dc->e_help, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&dc->e_zero, cudaEventDisableTiming));
        CK(cudaEventCreateWithFlags(&dc->e_fit, cudaEventDisableTiming));
        g_cache.push_back(dc);
    }
    return 0;
}

// Chained form?  Decided on the caller's max_reads (checked chunk by chunk: an understated one aborts the chained launch
// and the call starts over chunk by chunk).
int HostCall::decide_chain(bool allow_chain) {
    // chained form?  decided on the caller's max_reads (checked chunk by chunk below: an understated one aborts the
    // chained launch and the call starts over chunk by chunk)
    if (allow_chain && !dc->chain_off && want_rows && b->y_format == XPB200_Y_U16 && b->max_reads > 0 &&
        exp_int("XPB200_TM", 1) != 0 && exp_int("XPB200_CHAIN", 1) != 0 && total_bytes >= (8u << 20)) {
        chain_npw = tm_positions_per_warp(G, b->max_reads);
        // the chained launch is the single-warp class's; the chunks' team / streaming classes keep their own launches,
        // on the few SMs left free - so only where they are a sliver of the batch (a sample of the positions decides:
        // a wrong guess costs time, not results)
        const int split1 = tm_class_split(chain_npw);
        int64_t seen = 0, large = 0;
        for (int p = 0; p < P; p += 61) {
            const int64_t n = roff[p + 1] - roff[p];
            seen += n;
            if (n >= split1) large += n;
        }
        if (large * 200 <= seen) {
            if (int rc = fit_tm_geometry(G, std::min(b->max_reads, split1 - 1), chain_npw, &gch)) return rc;
            chain_fn = tm_chain_kernel(chain_npw, gch.maxt, G);
            if (chain_fn)
                if (int rc = prepare_kernel_fn(chain_fn, &gch)) return rc;
            if (gch.ctas_per_sm < 1 || gch.sms <= 2 * kChainFreeSms) chain_fn = nullptr;
        }
    }
    chain = chain_fn != nullptr;
    chain_two_class = chain && b->max_reads >= tm_class_split(chain_npw);  // chunks have a team class of their own
    chain_free_sms = std::max(1, exp_int("XPB200_CHAIN_FREE_SMS", kChainFreeSms));
    return 0;
}

int HostCall::lay_out() {
    size_t o = 0;
    o_y = o; o += align_up((size_t)b->n_reads * eb + 64);
    o_off = o; o += align_up((size_t)(P + 1) * 8);
    o_glen = o; o += align_up((size_t)P * G * 4);
    o_km = o; o += align_up((size_t)P * 8);
    o_ks = o; o += align_up((size_t)P * 8);
    o_gc = o; o += align_up((size_t)G * 4);
    o_status = o; o += align_up((size_t)P * 4);
    o_niter = o; o += align_up((size_t)P * 4);
    o_pval = o; o += align_up((size_t)P * 8);
    o_fit = o; o += align_up((size_t)P * FW * 8);
    o_stats = o; o += align_up((size_t)P * SW * 8);
    o_flags = o; o += align_up((size_t)kFlagWords * 4);
    o_nf = o; o += align_up((size_t)kMaxChunks * 4);
    o_nr = o; o += align_up((size_t)kMaxChunks * 4);
    o_cnt = o; o += want_rows ? align_up((size_t)kMaxChunks * kPackWarps * 4) : 0;
    RW = 4 + FW + SW;  // record width
    o_rows = o; o += want_rows ? align_up((size_t)P * RW * 8) : 0;  // chunk c's rows start at row cut[c]
    o_ws.assign(nchunks, 0);
    for (int c = 0; c < nchunks; ++c) {
        o_ws[c] = o;
        o += align_up(xpb200_workspace_bytes(cut[c + 1] - cut[c]));
    }
    if (o > dc->cap) {
        CK(cudaDeviceSynchronize());  // an asynchronous call may still be using the arena
        if (dc->buf) CK(cudaFree(dc->buf));
        dc->buf = nullptr;
        dc->cap = 0;
        const size_t want = o + o / 8;
        if (cudaMalloc(&dc->buf, want) != cudaSuccess) {
            cudaGetLastError();
            return fail(XPB200_ENOMEM, "cudaMalloc of " + std::to_string(want) + " bytes failed");
        }
        dc->cap = want;
    }
    d = reinterpret_cast<char*>(dc->buf);
    d_y = d + o_y;
    d_off = (int64_t*)(d + o_off);
    d_glen = (int32_t*)(d + o_glen);
    d_km = (double*)(d + o_km);
    d_ks = (double*)(d + o_ks);
    d_status = (uint32_t*)(d + o_status);
    d_niter = (int32_t*)(d + o_niter);
    d_pval = (double*)(d + o_pval);
    d_fit = (double*)(d + o_fit);
    d_stats = (double*)(d + o_stats);
    d_ready = (uint32_t*)(d + o_flags);
    d_abort = d_ready + kMaxChunks;
    d_done = d_abort + 1 + kMaxChunks;
    d_nl = d_done + kMaxChunks;
    d_nf = (uint32_t*)(d + o_nf);
    d_nr = (uint32_t*)(d + o_nr);
    d_cnt = (uint32_t*)(d + o_cnt);
    d_rows = (double*)(d + o_rows);

    return 0;
}

int HostCall::begin() {
    // the arena is shared with the previous call, which an asynchronous caller may not have waited for; and an
    // asynchronous call starts after the work already enqueued on the caller's stream
    if (dc->calls > 1) {
        CK(cudaStreamWaitEvent(dc->s_in, dc->e_call, 0));
        for (int i = 0; i < kComputeStreams; ++i) CK(cudaStreamWaitEvent(dc->s_c[i], dc->e_call, 0));
        CK(cudaStreamWaitEvent(dc->s_fit, dc->e_call, 0));
        CK(cudaStreamWaitEvent(dc->s_help, dc->e_call, 0));
        for (int i = 0; i < kComputeStreams; ++i) CK(cudaStreamWaitEvent(dc->s_post[i], dc->e_call, 0));
    }
    if (async) {
        CK(cudaEventRecord(dc->e_user, ho.user_stream));
        CK(cudaStreamWaitEvent(dc->s_in, dc->e_user, 0));
    }
    if (trace) CK(cudaEventRecord(dc->t_0, dc->s_in));
    CK(cudaMemcpyAsync(d + o_gc, b->grp_cond, (size_t)G * 4, cudaMemcpyDefault, dc->s_in));
    CK(cudaMemsetAsync(d_nf, 0, (size_t)kMaxChunks * 4, dc->s_in));
    CK(cudaMemsetAsync(d_nr, 0, (size_t)kMaxChunks * 4, dc->s_in));
    return 0;
}

int HostCall::chained() {
    bool chain_launched = false, chain_understated = false, chain_helper = false;
    xpk::FitArgs chain_fa;
    xpk::FitChain chain_ch;
    // An error return after the chained launch must not leave it polling for chunks that will never come
    auto chain_bail = [&](int rc) -> int {
        if (chain_launched) {
            cudaMemsetAsync(d_abort, 0xFF, 4, dc->s_poll);  // a stream nothing is ever queued on behind the launch
            cudaStreamSynchronize(dc->s_poll);
            cudaStreamSynchronize(dc->s_fit);
            cudaStreamSynchronize(dc->s_help);
            for (int i = 0; i < kComputeStreams; ++i) cudaStreamSynchronize(dc->s_c[i]);
            for (int i = 0; i < kComputeStreams; ++i) cudaStreamSynchronize(dc->s_post[i]);
            cudaStreamSynchronize(dc->s_in);
        }
        return rc;
    };
#define CKB(call)                                                                                        \
    do {                                                                                                 \
        cudaError_t e_ = (call);                                                                         \
        if (e_ != cudaSuccess) return chain_bail(fail(XPB200_ECUDA, std::string(#call) + ": " + cudaGetErrorString(e_))); \
    } while (0)
    if (chain) {
        CK(cudaMemsetAsync(d_ready, 0, (size_t)kFlagWords * 4, dc->s_in));
        CK(cudaEventRecord(dc->e_zero, dc->s_in));
        CK(cudaStreamWaitEvent(dc->s_fit, dc->e_zero, 0));
    }
    double host_enq[kMaxChunks] = {0.0};  // XPB200_TRACE: when the host had enqueued the chunk (ms since entry)
    int meta_next = 0;
    const size_t chain_meta_bytes = (size_t)std::max(1, exp_int("XPB200_CHAIN_META_MB", 128)) << 20;  // reads per group of small-array copies
    for (int c = 0; c < nchunks; ++c) {
        const int p0 = cut[c], p1 = cut[c + 1], pc = p1 - p0;
        const int64_t r0 = roff[p0], r1 = roff[p1];
        if (c == meta_next) {
            // the small arrays (offsets, group lengths, priors: 6 % of the bytes) go up for several chunks at once, in front
            // of the first one's reads: four more copies per chunk cost the copy engine ~25 us of gaps each time
            int ce = c;
            const size_t group_bytes = c == 0 ? chain_meta_bytes / 8 : chain_meta_bytes;  // the first group small: it is in front of the first reads
            while (ce < nchunks && (ce == c || (size_t)(roff[cut[ce + 1]] - roff[cut[c]]) * eb <= group_bytes)) ++ce;
            const int q0 = cut[c], qc = cut[ce] - cut[c];
            CKB(cudaMemcpyAsync(d_off + q0, roff + q0, (size_t)(qc + 1) * 8, cudaMemcpyDefault, dc->s_in));
            CKB(cudaMemcpyAsync(d_glen + (size_t)q0 * G, b->grp_len + (size_t)q0 * G, (size_t)qc * G * 4, cudaMemcpyDefault, dc->s_in));
            CKB(cudaMemcpyAsync(d_km + q0, b->kmer_mean + q0, (size_t)qc * 8, cudaMemcpyDefault, dc->s_in));
            CKB(cudaMemcpyAsync(d_ks + q0, b->kmer_std + q0, (size_t)qc * 8, cudaMemcpyDefault, dc->s_in));
            meta_next = ce;
        }
        CKB(cudaMemcpyAsync(d_y + (size_t)r0 * eb, (const char*)b->y + (size_t)r0 * eb, (size_t)(r1 - r0) * eb,
                            cudaMemcpyDefault, dc->s_in));
        CKB(cudaEventRecord(dc->e_in[c], dc->s_in));
        if (trace) CKB(cudaEventRecord(dc->t_in[c], dc->s_in));
        // the chunk's select phase (prefilter, worklist) on the SMs the chained launch leaves free, then its ready flag
        cudaStream_t sc = dc->s_c[c % n_streams];
        CKB(cudaStreamWaitEvent(sc, dc->e_in[c], 0));
        if (trace) CKB(cudaEventRecord(dc->t_start[c], sc));
        int64_t deepest = 0;
        for (int p = p0; p < p1; ++p) deepest = std::max(deepest, roff[p + 1] - roff[p]);
        if (deepest > (int64_t)b->max_reads) {  // the caller understated max_reads: start over, chunk by chunk
            chain_understated = true;
            break;
        }
        xpb200_batch db = *b;
        db.n_positions = pc;
        db.n_reads = r1 - r0;
        db.y = d_y; db.read_off = d_off + p0; db.grp_len = d_glen + (size_t)p0 * G;
        db.kmer_mean = d_km + p0; db.kmer_std = d_ks + p0; db.grp_cond = (const int32_t*)(d + o_gc);
        xpb200_result dr;
        dr.status = d_status + p0; dr.n_iter = d_niter + p0; dr.pval = d_pval + p0;
        dr.fit = d_fit + (size_t)p0 * FW; dr.stats = d_stats + (size_t)p0 * SW;
        dr.coef = nullptr;
        CKB(cudaMemsetAsync(dr.n_iter, 0, (size_t)pc * 4, sc));
        g_fit_grid_cap = chain_free_sms;
        const int rc_sel = fit_device_phases(&db, m, &dr, d + o_ws[c], xpb200_workspace_bytes(pc), d_nf + c, sc, kPhaseSelect | kPhaseFitLarge);
        g_fit_grid_cap = 0;
        if (rc_sel) return chain_bail(rc_sel);
        if (chain_two_class) CKB(cudaMemcpyAsync(d_nl + c, carve(d + o_ws[c]).n_large, 4, cudaMemcpyDeviceToDevice, sc));
        CKB(cudaMemsetAsync(d_ready + c, 0x01, 4, sc));
        CKB(cudaEventRecord(dc->e_cls[c], sc));
        if (trace) host_enq[c] = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
        if (trace) CKB(cudaEventRecord(dc->t_done[c], sc));
        if (c == 0) {  // the chained launch, as soon as the first chunk is on its way
            xpk::FitArgs& fa = chain_fa;
            xpk::FitChain& ch = chain_ch;
            memset(&fa, 0, sizeof(fa));
            fa.y = d_y; fa.y_format = b->y_format; fa.y_scale = b->y_scale;
            fa.read_off = d_off; fa.grp_len = d_glen; fa.kmer_mean = d_km; fa.kmer_std = d_ks;
            fa.G = G; fa.max_iters = m->max_iters; fa.stopping = m->stopping_criteria; fa.conc0 = m->dirichlet_conc;
            fa.lg_prior_w = xpm::xlgamma(2.0 * m->dirichlet_conc) - 2.0 * xpm::xlgamma(m->dirichlet_conc);
            fa.fit = d_fit; fa.n_iter = d_niter; fa.status = d_status; fa.coef = nullptr;
            fa.cap = gch.cap;
            xpk::fit_tm_fill_layout(fa, gch.np);
            memset(&ch, 0, sizeof(ch));
            ch.n = nchunks;
            ch.spin_limit = kChainSpinLimit;
            ch.ready = d_ready;
            ch.abort = d_abort;
            ch.wait_polls = d_abort + 1;
            ch.done = d_done;
            for (int k = 0; k < nchunks; ++k) {
                const Workspace wk = carve(d + o_ws[k]);
                ch.c[k].worklist = wk.worklist; ch.c[k].ticket = wk.ticket; ch.c[k].n_work = wk.n_work; ch.c[k].p0 = cut[k];
                ch.c[k].w_begin = chain_two_class ? wk.n_large : nullptr;
            }
            const int need = (P + gch.teams - 1) / gch.teams;
            const int grid = std::max(1, std::min(gch.sms * gch.ctas_per_sm - chain_free_sms, need));
            chain_helper = need > grid;
            chain_fn<<<grid, gch.threads, gch.smem, dc->s_fit>>>(fa, ch);
            ++g_launches;
            CKB(cudaGetLastError());
            chain_launched = true;
        }
    }
    if (chain_understated) {
        chain_bail(0);
        return fit_host_impl(device, b, m, ho, false);
    }
    if (chain_helper && exp_int("XPB200_CHAIN_HELPER", 1)) {
        // Once every chunk's select phase is through, the SMs kept free have nothing left to do: a second launch of the
        // same kernel on them walks the same chain (a batch whose fit outlasts its copy - c3, c5 - would otherwise run
        // its whole second half on 142 of 148 SMs)
        for (int i = 0; i < n_streams; ++i) {
            CKB(cudaEventRecord(dc->e_sel[i], dc->s_c[i]));
            CKB(cudaStreamWaitEvent(dc->s_help, dc->e_sel[i], 0));
        }
        chain_fn<<<chain_free_sms, gch.threads, gch.smem, dc->s_help>>>(chain_fa, chain_ch);
        ++g_launches;
        CKB(cudaGetLastError());
        CKB(cudaEventRecord(dc->e_help, dc->s_help));
        CKB(cudaStreamWaitEvent(dc->s_fit, dc->e_help, 0));
    }
    if (chain) {
        // statistics of every chunk once the chained launch is done, then the rows of the whole batch in position order
        CKB(cudaEventRecord(dc->e_fit, dc->s_fit));
        if (trace) {
            if (!dc->t_fit) {
                CKB(cudaEventCreate(&dc->t_fit));
                CKB(cudaEventCreate(&dc->t_post));
            }
            CKB(cudaEventRecord(dc->t_fit, dc->s_fit));
        }
        if (!async && exp_int("XPB200_CHAIN_POST", 1)) {
            // A chunk's statistics, row compaction and copy-out as soon as all its positions are fitted, while the chained
            // launch works on later chunks: the host polls the per-chunk counts of finished positions the kernel keeps.
            uint32_t* h_flags = dc->h_poll;                  // [kFlagWords]
            uint32_t* h_nf = dc->h_poll + kFlagWords + 3;    // [kMaxChunks]
            uint32_t* h_cnt = h_nf + kMaxChunks;             // [kMaxChunks] rows of each chunk
            const uint32_t *h_ready = h_flags, *h_abort = h_flags + kMaxChunks, *h_done = h_abort + 1 + kMaxChunks, *h_nl = h_done + kMaxChunks;
            int next_post = 0, next_copy = 0;
            int64_t total_rows = 0;
            bool overflow = false, failed = false;
            while (next_copy < nchunks) {
                bool progressed = false;
                if (next_post < nchunks) {
                    const bool fit_over = cudaEventQuery(dc->e_fit) == cudaSuccess;  // read BEFORE the counts
                    CKB(cudaMemcpyAsync(h_flags, d_ready, (size_t)kFlagWords * 4, cudaMemcpyDefault, dc->s_poll));
                    CKB(cudaMemcpyAsync(h_nf, d_nf, (size_t)kMaxChunks * 4, cudaMemcpyDefault, dc->s_poll));
                    CKB(cudaStreamSynchronize(dc->s_poll));
                    if (*h_abort != 0u) {
                        failed = true;
                        break;
                    }
                    while (next_post < nchunks && h_ready[next_post] != 0u &&
                           h_done[next_post] == h_nf[next_post] - h_nl[next_post]) {
                        const int c = next_post, p0 = cut[c], pc = cut[c + 1] - cut[c];
                        cudaStream_t sp = dc->s_post[c % kComputeStreams];
                        CKB(cudaStreamWaitEvent(sp, dc->e_cls[c], 0));
                        xpb200_batch db = *b;
                        db.n_positions = pc;
                        db.y = d_y; db.read_off = d_off + p0; db.grp_len = d_glen + (size_t)p0 * G;
                        db.kmer_mean = d_km + p0; db.kmer_std = d_ks + p0; db.grp_cond = (const int32_t*)(d + o_gc);
                        xpb200_result dr;
                        dr.status = d_status + p0; dr.n_iter = d_niter + p0; dr.pval = d_pval + p0;
                        dr.fit = d_fit + (size_t)p0 * FW; dr.stats = d_stats + (size_t)p0 * SW;
                        dr.coef = nullptr;
                        const Workspace wk = carve(d + o_ws[c]);
                        if (int rc = launch_stats(&db, m, &dr, wk.worklist, wk.n_work, sp)) return chain_bail(rc);
                        if (int rc = launch_pack_rows(dr, pc, FW, SW, ho.index_offset + p0, pc, d_cnt + (size_t)c * kPackWarps, d_nr + c,
                                                      d_rows + (size_t)p0 * RW, sp))
                            return chain_bail(rc);
                        CKB(cudaMemcpyAsync(h_cnt + c, d_nr + c, 4, cudaMemcpyDefault, sp));
                        CKB(cudaEventRecord(dc->e_done[c], sp));
                        ++next_post;
                        progressed = true;
                    }
                    if (!progressed && fit_over && next_post < nchunks) {  // the launch is over, yet a chunk is incomplete
                        failed = true;
                        break;
                    }
                }
                while (next_copy < next_post && cudaEventQuery(dc->e_done[next_copy]) == cudaSuccess) {
                    const int c = next_copy;
                    const int64_t cnt = h_cnt[c];
                    if (total_rows + cnt > ho.rows_capacity) overflow = true;
                    else if (cnt)
                        CKB(cudaMemcpyAsync(rows + (size_t)total_rows * RW, d_rows + (size_t)cut[c] * RW, (size_t)cnt * RW * 8,
                                            cudaMemcpyDefault, dc->s_out));
                    total_rows += cnt;
                    ++next_copy;
                    progressed = true;
                }
                if (!progressed) std::this_thread::sleep_for(std::chrono::microseconds(20));
            }
            if (failed) {  // a warp gave up waiting for a chunk (the free SMs were not free?): chunk by chunk, then
                dc->chain_off = true;
                chain_bail(0);
                CK(cudaEventRecord(dc->e_call, dc->s_out));
                return fit_host_impl(device, b, m, ho, false);
            }
            if (trace) CKB(cudaEventRecord(dc->t_post, dc->s_out));
            CK(cudaStreamSynchronize(dc->s_out));
            for (int i = 0; i < kComputeStreams; ++i) CK(cudaStreamSynchronize(dc->s_c[i]));
            for (int i = 0; i < kComputeStreams; ++i) CK(cudaStreamSynchronize(dc->s_post[i]));
            CK(cudaStreamSynchronize(dc->s_in));
            CK(cudaStreamSynchronize(dc->s_fit));
            CK(cudaEventRecord(dc->e_call, dc->s_out));
            *ho.n_rows = total_rows;
            uint32_t total = 0;
            for (int c = 0; c < nchunks; ++c) total += h_nf[c];
            if (ho.n_fitted) *ho.n_fitted = total;
            if (overflow) return fail(XPB200_ENOMEM, "rows buffer too small: " + std::to_string(total_rows) + " rows needed");
            if (trace) {
                const double host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
                float tf = 0, tp = 0;
                cudaEventElapsedTime(&tf, dc->t_0, dc->t_fit);
                cudaEventElapsedTime(&tp, dc->t_0, dc->t_post);
                fprintf(stderr, "[xpb200 trace] chained: %d chunks, host %.3f ms, chained launch done %.3f, last rows on their way %.3f ms\n", nchunks, host_ms, tf, tp);
            }
            return 0;
        }
        for (int c = 0; c < nchunks; ++c) {
            const int p0 = cut[c], pc = cut[c + 1] - cut[c];
            cudaStream_t sc = dc->s_c[c % n_streams];
            CKB(cudaStreamWaitEvent(sc, dc->e_fit, 0));
            xpb200_batch db = *b;
            db.n_positions = pc;
            db.y = d_y; db.read_off = d_off + p0; db.grp_len = d_glen + (size_t)p0 * G;
            db.kmer_mean = d_km + p0; db.kmer_std = d_ks + p0; db.grp_cond = (const int32_t*)(d + o_gc);
            xpb200_result dr;
            dr.status = d_status + p0; dr.n_iter = d_niter + p0; dr.pval = d_pval + p0;
            dr.fit = d_fit + (size_t)p0 * FW; dr.stats = d_stats + (size_t)p0 * SW;
            dr.coef = nullptr;
            const Workspace wk = carve(d + o_ws[c]);
            if (int rc = launch_stats(&db, m, &dr, wk.worklist, wk.n_work, sc)) return chain_bail(rc);
            CKB(cudaEventRecord(dc->e_done[c], sc));
            CKB(cudaStreamWaitEvent(dc->s_out, dc->e_done[c], 0));
        }
        xpb200_result dr;
        dr.status = d_status; dr.n_iter = d_niter; dr.pval = d_pval; dr.fit = d_fit; dr.stats = d_stats; dr.coef = nullptr;
        if (async) {
            // the rows straight into the caller's device buffer, the counts beside them; nothing waits on the host
            if (int rc = launch_pack_rows(dr, P, FW, SW, ho.index_offset, ho.rows_capacity, d_cnt, d_nr, ho.rows_dev, dc->s_out))
                return chain_bail(rc);
            xpk::xp_chain_counts_kernel<<<1, 32, 0, dc->s_out>>>(d_nr, d_nf, nchunks, d_abort, ho.counts_dev);
            ++g_launches;
            CKB(cudaGetLastError());
            CKB(cudaEventRecord(dc->e_call, dc->s_out));
            CKB(cudaStreamWaitEvent(ho.user_stream, dc->e_call, 0));
            return 0;
        }
        if (int rc = launch_pack_rows(dr, P, FW, SW, ho.index_offset, P, d_cnt, d_nr, d_rows, dc->s_out)) return chain_bail(rc);
        if (trace) CKB(cudaEventRecord(dc->t_post, dc->s_out));
        uint32_t tail[2] = {0u, 0u};  // rows, abort flag
        CKB(cudaMemcpyAsync(&tail[0], d_nr, 4, cudaMemcpyDefault, dc->s_out));
        CKB(cudaMemcpyAsync(&tail[1], d_abort, 4, cudaMemcpyDefault, dc->s_out));
        uint32_t nf[kMaxChunks];
        CKB(cudaMemcpyAsync(nf, d_nf, sizeof(nf), cudaMemcpyDefault, dc->s_out));
        CKB(cudaStreamSynchronize(dc->s_out));
        for (int i = 0; i < kComputeStreams; ++i) CK(cudaStreamSynchronize(dc->s_c[i]));
        CK(cudaStreamSynchronize(dc->s_in));
        CK(cudaStreamSynchronize(dc->s_fit));
        if (tail[1] != 0u) {  // a warp gave up waiting for a chunk (the free SMs were not free?): chunk by chunk, then
            dc->chain_off = true;
            CK(cudaEventRecord(dc->e_call, dc->s_out));
            return fit_host_impl(device, b, m, ho, false);
        }
        const int64_t total_rows = tail[0];
        *ho.n_rows = total_rows;
        uint32_t total = 0;
        for (int c = 0; c < nchunks; ++c) total += nf[c];
        if (ho.n_fitted) *ho.n_fitted = total;
        if (total_rows > ho.rows_capacity) {
            CK(cudaEventRecord(dc->e_call, dc->s_out));
            return fail(XPB200_ENOMEM, "rows buffer too small: " + std::to_string(total_rows) + " rows needed");
        }
        if (total_rows) CK(cudaMemcpyAsync(rows, d_rows, (size_t)total_rows * RW * 8, cudaMemcpyDefault, dc->s_out));
        CK(cudaStreamSynchronize(dc->s_out));
        CK(cudaEventRecord(dc->e_call, dc->s_out));
        if (trace) {
            const double host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
            float tf = 0, tp = 0;
            cudaEventElapsedTime(&tf, dc->t_0, dc->t_fit);
            cudaEventElapsedTime(&tp, dc->t_0, dc->t_post);
            fprintf(stderr, "[xpb200 trace] chained: %d chunks, host %.3f ms, chained launch done %.3f, rows packed %.3f ms\n", nchunks, host_ms, tf, tp);
            uint32_t polls[kMaxChunks];
            cudaMemcpy(polls, d_abort + 1, sizeof(polls), cudaMemcpyDefault);
            for (int c = 0; c < nchunks; ++c) {
                float a = 0, s2 = 0, e = 0;
                cudaEventElapsedTime(&a, dc->t_0, dc->t_in[c]);
                cudaEventElapsedTime(&s2, dc->t_0, dc->t_start[c]);
                cudaEventElapsedTime(&e, dc->t_0, dc->t_done[c]);
                fprintf(stderr, "[xpb200 trace] chunk %2d: %8d positions %7.1f MB  copied %7.3f  select %7.3f .. %7.3f ms  fitted %u  polls waiting %u  host enqueued %.3f\n", c,
                        cut[c + 1] - cut[c], (double)(roff[cut[c + 1]] - roff[cut[c]]) * eb / 1048576.0, a, s2, e, nf[c], polls[c], host_enq[c]);
            }
        }
        return 0;
    }
    return 0;
#undef CKB
}

int HostCall::chunked() {
    for (int c = 0; c < nchunks; ++c) {
        const int p0 = cut[c], p1 = cut[c + 1], pc = p1 - p0;
        if (pc == 0) continue;
        const int64_t r0 = roff[p0], r1 = roff[p1];
        // copy-in
        CK(cudaMemcpyAsync(d_y + (size_t)r0 * eb, (const char*)b->y + (size_t)r0 * eb, (size_t)(r1 - r0) * eb,
                           cudaMemcpyDefault, dc->s_in));
        CK(cudaMemcpyAsync(d_off + p0, roff + p0, (size_t)(pc + 1) * 8, cudaMemcpyDefault, dc->s_in));
        CK(cudaMemcpyAsync(d_glen + (size_t)p0 * G, b->grp_len + (size_t)p0 * G, (size_t)pc * G * 4, cudaMemcpyDefault, dc->s_in));
        CK(cudaMemcpyAsync(d_km + p0, b->kmer_mean + p0, (size_t)pc * 8, cudaMemcpyDefault, dc->s_in));
        CK(cudaMemcpyAsync(d_ks + p0, b->kmer_std + p0, (size_t)pc * 8, cudaMemcpyDefault, dc->s_in));
        CK(cudaEventRecord(dc->e_in[c], dc->s_in));
        if (trace) CK(cudaEventRecord(dc->t_in[c], dc->s_in));
        // compute
        cudaStream_t sc = dc->s_c[c % n_streams];
        CK(cudaStreamWaitEvent(sc, dc->e_in[c], 0));
        if (trace) CK(cudaEventRecord(dc->t_start[c], sc));
        xpb200_batch db = *b;
        db.n_positions = pc;
        db.n_reads = r1 - r0;
        // the chunk's own deepest position decides its launch classes (the host has the offsets: the caller's
        // max_reads is not relied upon)
        int64_t deepest = 0;
        for (int p = p0; p < p1; ++p) deepest = std::max(deepest, roff[p + 1] - roff[p]);
        if (deepest > 0x7fffffff) return fail(XPB200_EINVAL, "a position has more than 2^31 - 1 reads");
        db.max_reads = (int32_t)deepest;
        db.y = d_y; db.read_off = d_off + p0; db.grp_len = d_glen + (size_t)p0 * G;
        db.kmer_mean = d_km + p0; db.kmer_std = d_ks + p0; db.grp_cond = (const int32_t*)(d + o_gc);
        xpb200_result dr;
        dr.status = d_status + p0; dr.n_iter = d_niter + p0; dr.pval = d_pval + p0;
        dr.fit = d_fit + (size_t)p0 * FW; dr.stats = d_stats + (size_t)p0 * SW;
        dr.coef = nullptr;
        CK(cudaMemsetAsync(dr.n_iter, 0, (size_t)pc * 4, sc));
        CK(cudaMemsetAsync(dr.fit, 0xFF, (size_t)pc * FW * 8, sc));
        CK(cudaMemsetAsync(dr.stats, 0xFF, (size_t)pc * SW * 8, sc));
        if (int rc_fit = xpb200_fit_device(&db, m, &dr, d + o_ws[c], xpb200_workspace_bytes(pc), d_nf + c, sc)) return rc_fit;
        if (want_rows)  // compact the table's rows of this chunk
            if (int rc = launch_pack_rows(dr, pc, FW, SW, ho.index_offset + p0, pc, d_cnt + (size_t)c * kPackWarps, d_nr + c,
                                          d_rows + (size_t)p0 * RW, sc))
                return rc;
        CK(cudaEventRecord(dc->e_done[c], sc));
        if (trace) CK(cudaEventRecord(dc->t_done[c], sc));
        if (want_rows) continue;  // the rows are fetched below, once their counts are known
        // copy-out
        CK(cudaStreamWaitEvent(dc->s_out, dc->e_done[c], 0));
        CK(cudaMemcpyAsync(r->status + p0, dr.status, (size_t)pc * 4, cudaMemcpyDefault, dc->s_out));
        CK(cudaMemcpyAsync(r->n_iter + p0, dr.n_iter, (size_t)pc * 4, cudaMemcpyDefault, dc->s_out));
        CK(cudaMemcpyAsync(r->pval + p0, dr.pval, (size_t)pc * 8, cudaMemcpyDefault, dc->s_out));
        CK(cudaMemcpyAsync(r->fit + (size_t)p0 * FW, dr.fit, (size_t)pc * FW * 8, cudaMemcpyDefault, dc->s_out));
        CK(cudaMemcpyAsync(r->stats + (size_t)p0 * SW, dr.stats, (size_t)pc * SW * 8, cudaMemcpyDefault, dc->s_out));
    }
    if (async) {
        // rows of all chunks, concatenated on the device into the caller's buffer; nothing waits on the host
        for (int c = 0; c < nchunks; ++c)
            if (cut[c + 1] > cut[c]) CK(cudaStreamWaitEvent(dc->s_out, dc->e_done[c], 0));
        xpk::ConcatRowsArgs ca;
        ca.rows = d_rows; ca.n_rows = d_nr; ca.n_fitted = d_nf; ca.n_chunks = nchunks; ca.RW = RW;
        ca.capacity = ho.rows_capacity; ca.out = ho.rows_dev; ca.out_counts = ho.counts_dev;
        for (int c = 0; c < nchunks; ++c) ca.chunk_start[c] = cut[c];
        xpk::xp_concat_rows_kernel<<<148 * 4, 256, 0, dc->s_out>>>(ca);
        ++g_launches;
        CK(cudaGetLastError());
        CK(cudaEventRecord(dc->e_call, dc->s_out));
        CK(cudaStreamWaitEvent(ho.user_stream, dc->e_call, 0));
        return 0;
    }
    if (rows) {
        // every chunk is enqueued; fetch each chunk's rows as soon as it is done (count first, then the rows)
        int64_t total_rows = 0;
        bool overflow = false;
        for (int c = 0; c < nchunks; ++c) {
            const int p0 = cut[c], pc = cut[c + 1] - cut[c];
            if (pc == 0) continue;
            uint32_t cnt = 0;
            CK(cudaStreamWaitEvent(dc->s_out, dc->e_done[c], 0));
            CK(cudaMemcpyAsync(&cnt, d_nr + c, 4, cudaMemcpyDefault, dc->s_out));
            CK(cudaStreamSynchronize(dc->s_out));
            if (total_rows + cnt > ho.rows_capacity) {
                overflow = true;
                total_rows += cnt;
                continue;
            }
            if (cnt)
                CK(cudaMemcpyAsync(rows + (size_t)total_rows * RW, d_rows + (size_t)p0 * RW, (size_t)cnt * RW * 8,
                                   cudaMemcpyDefault, dc->s_out));
            total_rows += cnt;
        }
        CK(cudaStreamSynchronize(dc->s_out));
        *ho.n_rows = total_rows;
        if (overflow) {
            for (int i = 0; i < kComputeStreams; ++i) CK(cudaStreamSynchronize(dc->s_c[i]));
            CK(cudaStreamSynchronize(dc->s_in));
            return fail(XPB200_ENOMEM, "rows buffer too small: " + std::to_string(total_rows) + " rows needed");
        }
    }
    uint32_t nf[kMaxChunks];
    CK(cudaMemcpyAsync(nf, d_nf, sizeof(nf), cudaMemcpyDefault, dc->s_out));
    CK(cudaStreamSynchronize(dc->s_out));
    for (int i = 0; i < kComputeStreams; ++i) CK(cudaStreamSynchronize(dc->s_c[i]));
    CK(cudaStreamSynchronize(dc->s_in));
    CK(cudaEventRecord(dc->e_call, dc->s_out));
    uint32_t total = 0;
    for (int c = 0; c < nchunks; ++c) total += nf[c];
    if (ho.n_fitted) *ho.n_fitted = total;
    if (trace) {
        const double host_ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - host_t0).count();
        fprintf(stderr, "[xpb200 trace] %d chunks, host %.3f ms\n", nchunks, host_ms);
        for (int c = 0; c < nchunks; ++c) {
            if (cut[c + 1] == cut[c]) continue;
            float a = 0, s2 = 0, e = 0;
            cudaEventElapsedTime(&a, dc->t_0, dc->t_in[c]);
            cudaEventElapsedTime(&s2, dc->t_0, dc->t_start[c]);
            cudaEventElapsedTime(&e, dc->t_0, dc->t_done[c]);
            fprintf(stderr, "[xpb200 trace] chunk %2d: %8d positions %7.1f MB  copied %7.3f  start %7.3f  done %7.3f ms\n", c,
                    cut[c + 1] - cut[c], (double)(roff[cut[c + 1]] - roff[cut[c]]) * eb / 1048576.0, a, s2, e);
        }
    }
    return 0;
}

int fit_host_impl(int32_t device, const xpb200_batch* b, const xpb200_method* m, const HostOut& ho, bool allow_chain) {
    CK(cudaSetDevice(device));
    std::lock_guard<std::recursive_mutex> lk(g_cache_mu);
    HostCall hc(device, b, m, ho);
    if (int rc = hc.open_cache()) return rc;
    if (ho.n_fitted) *ho.n_fitted = 0;
    if (ho.n_rows) *ho.n_rows = 0;
    if (hc.P == 0) {
        if (hc.async) CK(cudaMemsetAsync(ho.counts_dev, 0, 8, ho.user_stream));
        return 0;
    }
    if (int rc = hc.decide_chain(allow_chain)) return rc;
    hc.cut = plan_chunks(hc.roff, hc.P, hc.eb, hc.chain);
    hc.nchunks = (int)hc.cut.size() - 1;
    DevCache* dc = hc.dc;
    ++dc->calls;
    hc.trace = exp_int("XPB200_TRACE", 0) != 0 && !hc.async;  // experiment builds: timed events per chunk, printed to stderr
    if (hc.trace && !dc->t_0) {
        CK(cudaEventCreate(&dc->t_0));
        for (int i = 0; i < kMaxChunks; ++i) {
            CK(cudaEventCreate(&dc->t_in[i]));
            CK(cudaEventCreate(&dc->t_start[i]));
            CK(cudaEventCreate(&dc->t_done[i]));
        }
    }
    hc.host_t0 = std::chrono::steady_clock::now();
    hc.n_streams = std::min(kComputeStreams, std::max(1, exp_int("XPB200_STREAMS", 3)));
    if (int rc = hc.lay_out()) return rc;
    if (int rc = hc.begin()) return rc;
    return hc.chain ? hc.chained() : hc.chunked();
}
}  // namespace

int32_t xpb200_host_chunk_plan(const int64_t* read_off, int32_t n_positions, int32_t elem_bytes, int32_t chained,
                               int32_t* cuts, int32_t capacity) {
    if (!read_off || !cuts || n_positions < 0 || (elem_bytes != 2 && elem_bytes != 4 && elem_bytes != 8))
        return fail(XPB200_EINVAL, "bad arguments");
    const std::vector<int> cut = plan_chunks(read_off, n_positions, (size_t)elem_bytes, chained != 0);
    if ((int)cut.size() > capacity) return fail(XPB200_EINVAL, "cuts holds fewer than " + std::to_string(cut.size()) + " entries");
    for (size_t i = 0; i < cut.size(); ++i) cuts[i] = cut[i];
    return (int32_t)cut.size() - 1;
}

void xpb200_release(void) {
    std::lock_guard<std::recursive_mutex> lk(g_cache_mu);
    for (auto* c : g_cache) {
        cudaSetDevice(c->device);
        cudaDeviceSynchronize();  // an asynchronous call may still be using the arena
        if (c->buf) cudaFree(c->buf);
        for (cudaStream_t st : {c->s_in, c->s_c[0], c->s_c[1], c->s_c[2], c->s_out, c->s_fit, c->s_help})
            if (st) cudaStreamDestroy(st);
        for (int i = 0; i < kComputeStreams; ++i) {
            if (c->e_sel[i]) cudaEventDestroy(c->e_sel[i]);
            if (c->s_post[i]) cudaStreamDestroy(c->s_post[i]);
        }
        if (c->s_poll) cudaStreamDestroy(c->s_poll);
        for (int i = 0; i < kMaxChunks; ++i) cudaEventDestroy(c->e_cls[i]);
        if (c->h_poll) cudaFreeHost(c->h_poll);
        if (c->e_help) cudaEventDestroy(c->e_help);
        if (c->e_zero) cudaEventDestroy(c->e_zero);
        if (c->e_fit) cudaEventDestroy(c->e_fit);
        for (int i = 0; i < kMaxChunks; ++i) {
            cudaEventDestroy(c->e_in[i]);
            cudaEventDestroy(c->e_done[i]);
        }
        if (c->e_call) cudaEventDestroy(c->e_call);
        if (c->e_user) cudaEventDestroy(c->e_user);
        delete c;
    }
    g_cache.clear();
    std::lock_guard<std::mutex> lk2(g_side_mu);
    for (auto& kv : g_side) {
        cudaSetDevice(kv.first.first);
        for (int i = 0; i < 2; ++i) {
            if (kv.second->s[i]) cudaStreamDestroy(kv.second->s[i]);
            if (kv.second->join[i]) cudaEventDestroy(kv.second->join[i]);
        }
        if (kv.second->fork) cudaEventDestroy(kv.second->fork);
        delete kv.second;
    }
    g_side.clear();
}

int32_t xpb200_measure_fp64_peak(int32_t device, int32_t repeats, double* tflops) {
    if (!tflops || repeats < 1) return fail(XPB200_EINVAL, "bad arguments");
    CK(cudaSetDevice(device));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, device));
    double* out = nullptr;
    CK(cudaMalloc(&out, 8));
    const int iters = 1 << 16, threads = 256, blocks = sms * 8;
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    xpk::xp_dfma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0);  // warm-up
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int i = 0; i < repeats; ++i) {
        CK(cudaEventRecord(e0));
        xpk::xp_dfma_peak_kernel<<<blocks, threads>>>(out, iters, 1.0);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms = 0;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        best = std::min(best, ms);
    }
    g_launches += repeats + 1;
    CK(cudaGetLastError());
    const double flops = 2.0 * 8.0 * (double)iters * threads * (double)blocks;
    *tflops = flops / (best * 1e-3) / 1e12;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(out);
    return 0;
}

int32_t xpb200_posterior_device(const xpb200_batch* b, const xpb200_result* r, double* prob0, double* ln_prob,
                                void* stream) {
    if (!b || !r || !r->status || !r->n_iter || !r->coef || !prob0 || !ln_prob)
        return fail(XPB200_EINVAL, "null argument (the fit must have been given result->coef)");
    if (b->n_positions <= 0) return 0;
    xpk::PosteriorArgs a;
    a.y = b->y; a.y_format = b->y_format; a.y_scale = b->y_scale; a.read_off = b->read_off; a.grp_len = b->grp_len;
    a.kmer_mean = b->kmer_mean; a.status = r->status; a.n_iter = r->n_iter; a.coef = r->coef;
    a.P = b->n_positions; a.G = b->n_groups; a.prob0 = prob0; a.ln_prob = ln_prob;
    const int threads = 256, wpb = threads / 32;
    const int blocks = std::max(1, std::min((b->n_positions + wpb - 1) / wpb, 148 * 8));
    xpk::xp_posterior_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(a);
    ++g_launches;
    CK(cudaGetLastError());
    return 0;
}

int32_t xpb200_pack_records(const xpb200_result* r, int32_t P, int32_t G, int32_t C, const void* workspace,
                            int64_t index_offset, double* out, int64_t capacity_rows, void* stream) {
    if (!r || !workspace || !out) return fail(XPB200_EINVAL, "null argument");
    if (P <= 0) return 0;
    Workspace ws = carve(const_cast<void*>(workspace));
    xpk::PackArgs a;
    a.worklist = ws.worklist; a.n_work = ws.n_work; a.status = r->status; a.n_iter = r->n_iter; a.pval = r->pval;
    a.fit = r->fit; a.stats = r->stats; a.FW = xpc::fit_width(G); a.SW = xpc::stats_width(G, C);
    a.index_offset = index_offset; a.capacity = capacity_rows; a.out = out;
    const int threads = 256, wpb = threads / 32;
    const int blocks = std::max(1, std::min((P + wpb - 1) / wpb, 148 * 8));
    xpk::xp_pack_records_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(a);
    ++g_launches;
    CK(cudaGetLastError());
    return 0;
}

int32_t xpb200_pack_rows(const xpb200_result* r, int32_t P, int32_t G, int32_t C, void* workspace,
                         int64_t index_offset, double* out, int64_t capacity_rows, uint32_t* n_rows_dev, void* stream) {
    if (!r || !workspace || !out || !n_rows_dev) return fail(XPB200_EINVAL, "null argument");
    if (P <= 0) {
        CK(cudaMemsetAsync(n_rows_dev, 0, 4, (cudaStream_t)stream));
        return 0;
    }
    Workspace ws = carve(workspace);
    return launch_pack_rows(*r, P, xpc::fit_width(G), xpc::stats_width(G, C), index_offset, capacity_rows,
                            ws.pack_counts, n_rows_dev, out, (cudaStream_t)stream);
}

// Test utility: overwrite the shared memory of every SM with a 64-bit pattern (a later kernel that
// reads shared memory it has not written itself then sees that pattern instead of its own leftovers).
__global__ void xp_fill_smem_kernel(unsigned long long pattern, unsigned long long* sink) {
    extern __shared__ unsigned long long fill_buf[];
    const int nwords = (int)(kMaxSmemPerCta / 8);
    if (pattern != 0xdeadull)  // 0xdead: touch nothing (occupies the SM's shared memory only)
        for (int i = threadIdx.x; i < nwords; i += blockDim.x) fill_buf[i] = pattern;
    __syncthreads();
    if (fill_buf[(threadIdx.x * 97) % nwords] != pattern) sink[0] = 1;  // keeps the stores alive
}
int32_t xpb200_debug_fill_smem(uint64_t pattern, void* stream) {
    int dev = 0, sms = 0;
    CK(cudaGetDevice(&dev));
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    static unsigned long long* sink = nullptr;
    if (!sink) CK(cudaMalloc(&sink, 8));
    CK(cudaFuncSetAttribute(xp_fill_smem_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxSmemPerCta));
    xp_fill_smem_kernel<<<sms, 256, kMaxSmemPerCta, (cudaStream_t)stream>>>(pattern, sink);
    ++g_launches;
    CK(cudaGetLastError());
    return 0;
}

int32_t xpb200_set_stage_timing(int32_t on) {
    if (on && !g_timer.ev[0])
        for (auto& e : g_timer.ev) CK(cudaEventCreate(&e));
    g_timer.on = on != 0;
    g_timer.valid = false;
    return 0;
}

int32_t xpb200_get_stage_ms(float* ms4) {
    if (!ms4) return fail(XPB200_EINVAL, "null out");
    if (!g_timer.valid) return fail(XPB200_EINVAL, "no timed xpb200_fit_device call on this thread");
    CK(cudaEventSynchronize(g_timer.ev[4]));
    for (int i = 0; i < 4; ++i) CK(cudaEventElapsedTime(&ms4[i], g_timer.ev[i], g_timer.ev[i + 1]));
    return 0;
}

int32_t xpb200_eval_special(int32_t which, const double* x, const double* aux, double* out, int64_t n, void* stream) {
    if (n <= 0) return 0;
    const int threads = 128;
    xpk::xp_eval_special_kernel<<<(unsigned)((n + threads - 1) / threads), threads, 0, (cudaStream_t)stream>>>(
        which, x, aux, out, (long long)n);
    ++g_launches;
    CK(cudaGetLastError());
    return 0;
}

}  // extern "C"
