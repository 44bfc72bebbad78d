// xp_ingest.cpp — native ingest of dataprep output: data.json byte spans -> CSR batch (host, C++).
//
// Replaces, for many genes at once and on several threads, what the reference does per gene in Python:
//   xpore/scripts/diffmod.py:156-169   seek(start); read(end-start); ujson.loads  (one run at a time)
//   xpore/diffmod/io.py:21-90          load_data: (position,kmer) pairs = intersection over the runs that
//                                      have the gene; reads concatenated run by run in config order;
//                                      non-pooled: drop a run with n < min or n > max (:51-52); a
//                                      condition counts iff (pooled) sum n in [min,max] (:70-73) or
//                                      (non-pooled) any retained run (:75-77); skip the position if fewer
//                                      than two conditions count (:79-80) or it has no reads (:66-67)
// The line format is the one dataprep writes (dataprep.py:485-495,642-651):
//   {"<idx>":{"<position>":{"<kmer>":[floats]}}}
// Numbers are converted with std::from_chars (correctly rounded, i.e. the double Python's float() gives).
// Positions are emitted sorted by (int(position), kmer) inside a gene, genes in the order given (the
// reference iterates a Python set: hash-seed dependent order, SURVEY.md P-order).
#include <charconv>
#include <cstdio>
#include <cstring>
#include <fcntl.h>
#include <unistd.h>
#include <algorithm>
#include <atomic>
#include <map>
#include <string>
#include <string_view>
#include <thread>
#include <unordered_map>
#include <vector>

#include "../../include/xpore_b200.h"

namespace {

thread_local std::string g_ingest_err;

struct ReadsRef {  // reads of one (position, kmer) in one run: a slice of the run's value pool
    size_t begin;
    uint32_t n;
};

struct RunGene {  // one gene of one run, parsed
    std::vector<double> pool;
    // the same reads as scaled integers q = 100 y while every literal so far is a non-negative decimal with at most two
    // fraction digits below 655.36 (what dataprep writes, dataprep.py:638): q / 100.0 is then the very double in `pool`
    // (both are the correctly rounded value of the same rational), and the batch can travel as 2 bytes per read
    std::vector<uint16_t> qpool;
    bool all_q = true;
    // (position, kmer) -> slice; string_views point into the raw JSON text held by the caller
    std::map<std::pair<std::string_view, std::string_view>, ReadsRef> pairs;
};

struct Cursor {
    const char* p;
    const char* end;
    void ws() {
        while (p < end && (*p == ' ' || *p == '\n' || *p == '\r' || *p == '\t')) ++p;
    }
    bool eat(char c) {
        ws();
        if (p < end && *p == c) {
            ++p;
            return true;
        }
        return false;
    }
    // a JSON string without escape processing (dataprep writes ids, integers and ACGT k-mers)
    bool str(std::string_view& out) {
        ws();
        if (p >= end || *p != '"') return false;
        const char* s = ++p;
        while (p < end && *p != '"') {
            if (*p == '\\') ++p;  // keep escapes verbatim
            ++p;
        }
        if (p >= end) return false;
        out = std::string_view(s, (size_t)(p - s));
        ++p;
        return true;
    }
};

// A JSON number.  Fast path (Clinger): [-]digits[.digits] with at most 15 significant digits is
// m / 10^k with m < 2^53 and 10^k exact in a double, so one IEEE division gives the correctly rounded
// value - the same double strtod / Python's float() return.  Anything else (exponents, long
// mantissas) goes through std::from_chars, which is correctly rounded as well.
inline bool parse_number(const char*& p, const char* end, double& out, uint32_t& qv, bool& q_ok) {
    static const double kPow10[] = {1e0, 1e1, 1e2, 1e3, 1e4, 1e5, 1e6, 1e7, 1e8, 1e9, 1e10, 1e11, 1e12, 1e13, 1e14, 1e15};
    const char* q = p;
    bool neg = false;
    if (q < end && *q == '-') {
        neg = true;
        ++q;
    }
    unsigned long long m = 0;
    int digits = 0, frac = 0;
    const char* d0 = q;
    while (q < end && *q >= '0' && *q <= '9') {
        m = m * 10 + (unsigned)(*q - '0');
        ++q;
        ++digits;
    }
    if (q == d0) return false;
    if (q < end && *q == '.') {
        ++q;
        const char* f0 = q;
        while (q < end && *q >= '0' && *q <= '9') {
            m = m * 10 + (unsigned)(*q - '0');
            ++q;
            ++digits;
            ++frac;
        }
        if (q == f0) return false;
    }
    if (digits <= 15 && (q >= end || (*q != 'e' && *q != 'E'))) {
        const double v = (double)m / kPow10[frac];
        out = neg ? -v : v;
        p = q;
        const unsigned long long scaled = frac == 0 ? m * 100 : frac == 1 ? m * 10 : m;
        q_ok = !neg && frac <= 2 && scaled < 65536;
        qv = (uint32_t)scaled;
        return true;
    }
    q_ok = false;
    auto r = std::from_chars(p, end, out);
    if (r.ec != std::errc()) return false;
    p = r.ptr;
    return true;
}

// {"idx":{"pos":{"kmer":[...]}}}
bool parse_gene(const char* text, size_t len, RunGene& g, std::string& err) {
    Cursor c{text, text + len};
    g.pool.reserve(len / 6 + 16);
    g.qpool.reserve(len / 6 + 16);
    std::string_view idx;
    if (!c.eat('{') || !c.str(idx) || !c.eat(':') || !c.eat('{')) {
        err = "data.json: expected {\"<idx>\":{";
        return false;
    }
    if (!c.eat('}')) {
        for (;;) {
            std::string_view pos;
            if (!c.str(pos) || !c.eat(':') || !c.eat('{')) {
                err = "data.json: expected \"<position>\":{";
                return false;
            }
            if (!c.eat('}')) {
                for (;;) {
                    std::string_view kmer;
                    if (!c.str(kmer) || !c.eat(':') || !c.eat('[')) {
                        err = "data.json: expected \"<kmer>\":[";
                        return false;
                    }
                    ReadsRef ref{g.pool.size(), 0};
                    if (!c.eat(']')) {
                        for (;;) {
                            c.ws();
                            double v;
                            uint32_t qv = 0;
                            bool q_ok = false;
                            if (!parse_number(c.p, c.end, v, qv, q_ok)) {
                                err = "data.json: bad number";
                                return false;
                            }
                            g.pool.push_back(v);
                            if (g.all_q) {
                                if (q_ok) g.qpool.push_back((uint16_t)qv);
                                else {
                                    g.all_q = false;
                                    g.qpool.clear();
                                }
                            }
                            if (c.eat(',')) continue;
                            if (c.eat(']')) break;
                            err = "data.json: expected , or ] in a read list";
                            return false;
                        }
                    }
                    ref.n = (uint32_t)(g.pool.size() - ref.begin);
                    g.pairs[{pos, kmer}] = ref;  // a repeated key keeps the last value, like a Python dict
                    if (c.eat(',')) continue;
                    if (c.eat('}')) break;
                    err = "data.json: expected , or } after a k-mer";
                    return false;
                }
            }
            if (c.eat(',')) continue;
            if (c.eat('}')) break;
            err = "data.json: expected , or } after a position";
            return false;
        }
    }
    if (!c.eat('}')) {
        err = "data.json: expected } at the end of the gene";
        return false;
    }
    return true;
}

bool all_digits(std::string_view s, long long& v) {
    if (s.empty()) return false;
    size_t i = (s[0] == '-' || s[0] == '+') ? 1 : 0;
    if (i == s.size() || s.size() - i > 18) return false;
    v = 0;
    for (size_t k = i; k < s.size(); ++k) {
        if (s[k] < '0' || s[k] > '9') return false;
        v = v * 10 + (s[k] - '0');
    }
    if (s[0] == '-') v = -v;
    return true;
}

struct GeneOut {  // selected positions of one gene
    std::vector<double> y;
    std::vector<uint16_t> yq;       // the same reads as u16 codes when all_q
    bool all_q = true;
    int32_t max_reads = 0;
    std::vector<int32_t> grp_len;   // [npos * G]
    std::vector<int32_t> kmer_row;  // [npos]
    std::string strings;            // position \0 kmer \0 ...
    int32_t npos = 0;
    std::string err;
};

struct Ingest {
    int n_runs = 0, n_groups = 0, n_conditions = 0, pooling = 0;
    int min_count = 0, max_count = 0;
    std::vector<int> run_slot, run_cond;
    std::vector<int> fds;  // pread() is thread-safe: no lock around the reads
    std::unordered_map<std::string, int> kmer_row;
    std::vector<GeneOut> genes;
    bool all_q = false;     // every read of the last xpb200_ingest_genes call has a u16 code
    int32_t max_reads = 0;  // deepest position of that call
};

void process_gene(Ingest& in, const int64_t* starts, const int64_t* ends, GeneOut& out) {
    const int R = in.n_runs, G = in.n_groups, C = in.n_conditions;
    std::vector<std::string> text((size_t)R);
    std::vector<RunGene> rg((size_t)R);
    std::vector<int> have;
    for (int r = 0; r < R; ++r) {
        if (starts[r] < 0) continue;  // the run does not have this gene (diffmod.py:160-163)
        const size_t len = (size_t)(ends[r] - starts[r]);
        text[r].resize(len);
        size_t done = 0;
        while (done < len) {
            const ssize_t k = pread(in.fds[r], &text[r][done], len - done, (off_t)(starts[r] + (int64_t)done));
            if (k <= 0) {
                out.err = "short read from data.json of run " + std::to_string(r);
                return;
            }
            done += (size_t)k;
        }
        if (!parse_gene(text[r].data(), len, rg[r], out.err)) return;
        have.push_back(r);
    }
    if (have.empty()) return;
    for (int r : have) out.all_q = out.all_q && rg[r].all_q;
    // (position, kmer) pairs present in every run that has the gene (io.py:31-39), in table order
    struct Key {
        int cls;
        long long ipos;
        std::string_view pos, kmer;
    };
    std::vector<Key> keys;
    for (const auto& kv : rg[have[0]].pairs) {
        bool everywhere = true;
        for (size_t j = 1; j < have.size() && everywhere; ++j)
            everywhere = rg[have[j]].pairs.count(kv.first) != 0;
        if (!everywhere) continue;
        Key k{1, 0, kv.first.first, kv.first.second};
        if (all_digits(k.pos, k.ipos)) k.cls = 0;
        keys.push_back(k);
    }
    std::sort(keys.begin(), keys.end(), [](const Key& a, const Key& b) {
        if (a.cls != b.cls) return a.cls < b.cls;
        if (a.cls == 0) {
            if (a.ipos != b.ipos) return a.ipos < b.ipos;
            return a.kmer < b.kmer;
        }
        std::string sa(a.pos), sb(b.pos);  // non-integer positions: by position + kmer text
        sa += a.kmer;
        sb += b.kmer;
        return sa < sb;
    });
    std::vector<int32_t> glen((size_t)G);
    std::vector<long long> cond_sum((size_t)C);
    std::vector<char> cond_has((size_t)C);
    std::vector<std::pair<int, ReadsRef>> order;  // (run, slice) in emission order
    for (const Key& k : keys) {
        std::fill(glen.begin(), glen.end(), 0);
        std::fill(cond_sum.begin(), cond_sum.end(), 0);
        std::fill(cond_has.begin(), cond_has.end(), 0);
        long long total = 0;
        order.clear();
        for (int r : have) {
            const ReadsRef ref = rg[r].pairs.find({k.pos, k.kmer})->second;
            const int n = (int)ref.n;
            if (!in.pooling && (n < in.min_count || n > in.max_count)) continue;  // io.py:51-52
            const int c = in.run_cond[r];
            cond_sum[c] += n;
            cond_has[c] = 1;
            total += n;
            const int slot = in.pooling ? c : in.run_slot[r];
            glen[slot] += n;
            order.push_back({r, ref});
        }
        if (total == 0) continue;  // io.py:66-67
        int n_incl = 0;
        for (int c = 0; c < C; ++c)
            n_incl += in.pooling ? (cond_sum[c] >= in.min_count && cond_sum[c] <= in.max_count) : (cond_has[c] != 0);
        if (n_incl < 2) continue;  // io.py:79-80
        auto it = in.kmer_row.find(std::string(k.kmer));
        if (it == in.kmer_row.end()) {
            out.err = "k-mer not in the model: " + std::string(k.kmer);  // model_kmer.loc[kmer] -> KeyError
            return;
        }
        // reads: per group slot contiguous, slots in order; inside a slot run by run in config order
        auto emit = [&](const std::pair<int, ReadsRef>& o) {
            const RunGene& src = rg[o.first];
            if (out.all_q)
                out.yq.insert(out.yq.end(), src.qpool.begin() + (long)o.second.begin,
                              src.qpool.begin() + (long)(o.second.begin + o.second.n));
            else
                out.y.insert(out.y.end(), src.pool.begin() + (long)o.second.begin,
                             src.pool.begin() + (long)(o.second.begin + o.second.n));
        };
        if (in.pooling) {
            for (int c = 0; c < C; ++c)
                for (const auto& o : order)
                    if (in.run_cond[o.first] == c && o.second.n) emit(o);
        } else {
            for (int s = 0; s < G; ++s)
                for (const auto& o : order)
                    if (in.run_slot[o.first] == s && o.second.n) emit(o);
        }
        out.max_reads = std::max<long long>(out.max_reads, total);
        out.grp_len.insert(out.grp_len.end(), glen.begin(), glen.end());
        out.kmer_row.push_back(it->second);
        out.strings.append(k.pos.data(), k.pos.size());
        out.strings.push_back('\0');
        out.strings.append(k.kmer.data(), k.kmer.size());
        out.strings.push_back('\0');
        ++out.npos;
    }
}

}  // namespace

extern "C" {

const char* xpb200_ingest_last_error(void) { return g_ingest_err.c_str(); }

void* xpb200_ingest_open(int32_t n_runs, const char* const* json_paths, const int32_t* run_slot,
                         const int32_t* run_cond, int32_t n_groups, int32_t n_conditions, int32_t pooling,
                         int32_t min_count, int32_t max_count, int32_t n_kmers, const char* const* kmers) {
    if (n_runs < 1 || !json_paths || !run_slot || !run_cond || n_groups < 1 || n_conditions < 1) {
        g_ingest_err = "xpb200_ingest_open: bad arguments";
        return nullptr;
    }
    Ingest* in = new Ingest();
    in->n_runs = n_runs;
    in->n_groups = n_groups;
    in->n_conditions = n_conditions;
    in->pooling = pooling;
    in->min_count = min_count;
    in->max_count = max_count;
    in->run_slot.assign(run_slot, run_slot + n_runs);
    in->run_cond.assign(run_cond, run_cond + n_runs);
    for (int r = 0; r < n_runs; ++r) {
        const int fd = open(json_paths[r], O_RDONLY);
        if (fd < 0) {
            g_ingest_err = std::string("cannot open ") + json_paths[r];
            for (int o : in->fds) close(o);
            delete in;
            return nullptr;
        }
        in->fds.push_back(fd);
    }
    for (int i = 0; i < n_kmers; ++i) in->kmer_row.emplace(kmers[i], i);
    return in;
}

void xpb200_ingest_close(void* handle) {
    Ingest* in = static_cast<Ingest*>(handle);
    if (!in) return;
    for (int fd : in->fds) close(fd);
    delete in;
}

int32_t xpb200_ingest_genes(void* handle, int32_t n_genes, const int64_t* starts, const int64_t* ends,
                            int32_t n_threads, int64_t* n_positions, int64_t* n_reads, int64_t* string_bytes) {
    Ingest* in = static_cast<Ingest*>(handle);
    if (!in || n_genes < 0 || (n_genes > 0 && (!starts || !ends))) {
        g_ingest_err = "xpb200_ingest_genes: bad arguments";
        return XPB200_EINVAL;
    }
    in->genes.assign((size_t)n_genes, GeneOut());
    std::atomic<int> next(0);
    const int R = in->n_runs;
    auto work = [&]() {
        for (;;) {
            const int g = next.fetch_add(1);
            if (g >= n_genes) break;
            process_gene(*in, starts + (size_t)g * R, ends + (size_t)g * R, in->genes[(size_t)g]);
        }
    };
    const int nt = std::max(1, std::min(n_threads, n_genes));
    std::vector<std::thread> pool;
    for (int t = 1; t < nt; ++t) pool.emplace_back(work);
    work();
    for (auto& t : pool) t.join();
    int64_t P = 0, Rd = 0, S = 0;
    in->all_q = true;
    in->max_reads = 0;
    for (const GeneOut& g : in->genes) {
        if (!g.err.empty()) {
            g_ingest_err = g.err;
            return XPB200_EINVAL;
        }
        P += g.npos;
        Rd += (int64_t)(g.all_q ? g.yq.size() : g.y.size());
        S += (int64_t)g.strings.size();
        in->all_q = in->all_q && g.all_q;
        in->max_reads = std::max(in->max_reads, g.max_reads);
    }
    if (n_positions) *n_positions = P;
    if (n_reads) *n_reads = Rd;
    if (string_bytes) *string_bytes = S;
    return XPB200_OK;
}

int32_t xpb200_ingest_copy(void* handle, double* y, int64_t* read_off, int32_t* grp_len, int32_t* kmer_row,
                           int32_t* gene_of_position, char* strings) {
    Ingest* in = static_cast<Ingest*>(handle);
    if (!in || !read_off) {
        g_ingest_err = "xpb200_ingest_copy: bad arguments";
        return XPB200_EINVAL;
    }
    const int G = in->n_groups;
    int64_t p = 0, r = 0, s = 0;
    read_off[0] = 0;
    for (size_t gi = 0; gi < in->genes.size(); ++gi) {
        const GeneOut& g = in->genes[gi];
        const size_t nr = g.all_q ? g.yq.size() : g.y.size();
        if (y) {
            if (g.all_q)
                for (size_t i = 0; i < nr; ++i) y[r + (int64_t)i] = (double)g.yq[i] / 100.0;
            else if (nr) memcpy(y + r, g.y.data(), nr * sizeof(double));
        }
        if (grp_len && g.npos) memcpy(grp_len + p * G, g.grp_len.data(), g.grp_len.size() * sizeof(int32_t));
        if (kmer_row && g.npos) memcpy(kmer_row + p, g.kmer_row.data(), g.kmer_row.size() * sizeof(int32_t));
        if (strings && !g.strings.empty()) memcpy(strings + s, g.strings.data(), g.strings.size());
        int64_t rr = r;
        for (int32_t k = 0; k < g.npos; ++k) {
            int64_t n = 0;
            for (int j = 0; j < G; ++j) n += g.grp_len[(size_t)k * G + j];
            rr += n;
            read_off[p + k + 1] = rr;
            if (gene_of_position) gene_of_position[p + k] = (int32_t)gi;
        }
        p += g.npos;
        r += (int64_t)nr;
        s += (int64_t)g.strings.size();
    }
    return XPB200_OK;
}

int32_t xpb200_ingest_encoding(void* handle, int32_t* max_reads) {
    Ingest* in = static_cast<Ingest*>(handle);
    if (!in) return XPB200_EINVAL;
    if (max_reads) *max_reads = in->max_reads;
    return in->all_q ? XPB200_Y_U16 : XPB200_Y_F64;
}

int32_t xpb200_ingest_copy_u16(void* handle, uint16_t* yq) {
    Ingest* in = static_cast<Ingest*>(handle);
    if (!in || !yq || !in->all_q) {
        g_ingest_err = "xpb200_ingest_copy_u16: no u16 encoding of the last xpb200_ingest_genes call";
        return XPB200_EINVAL;
    }
    int64_t r = 0;
    for (const GeneOut& g : in->genes) {
        if (!g.yq.empty()) memcpy(yq + r, g.yq.data(), g.yq.size() * sizeof(uint16_t));
        r += (int64_t)g.yq.size();
    }
    return XPB200_OK;
}

}  // extern "C"
