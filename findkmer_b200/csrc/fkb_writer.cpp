// fkb_writer.cpp -- host output writer: the two consumers of the counting path, on the dense table.
//
//   fkb_write_base_stats  <-> statistics()       findKmer/src/findKmer.cpp:491-565
//   fkb_write_histogram   <-> histo_recursive()  findKmer/src/findKmer.cpp:699-942 (live part :727-891)
//   fkb_max_nodes         <-> estimate_RAM_usage() :1256-1260
//
// Byte-identical output needs the same arithmetic TYPES in the same ORDER as the reference (a mix of
// double and x87 80-bit long double) and the same libc formatting (%LE, %Lf, %d).  What is new:
//  * rows come from iterating the dense table in index order (== the trie's pre-order A<C<G<T walk, :719-724)
//    and skipping zeros, instead of chasing pointers;
//  * everything that depends only on a k-mer's base COMPOSITION (h, H, the estimated proportion p, n*p,
//    the standard deviation, the normal-approximation test and the formatted h/H text) is computed once
//    per composition -- there are at most C(k+3,3) of them (364 at k = 11) against 4^k rows;
//  * rows are formatted by several host threads into per-slice buffers and written in order.
// Compile WITHOUT -ffast-math / -march flags that enable FMA contraction.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "findkmer_b200.h"

extern "C" uint64_t fkb_max_nodes(int k)
{
    uint64_t n = 1;  // the head node
    for (int d = 1; d <= k; ++d) n += (uint64_t)1 << (2 * d);
    return n;
}

extern "C" int fkb_write_base_stats(FILE *stats_out, FILE *console, int k, const fkb_counts *counts, long double base_probability[4])
{
    if (!counts || !base_probability) return FKB_ERR_BAD_ARG;
    if (console) fprintf(console, "Statistics of occurrences and probability of A, C, G and T respectively: \n");
    for (int i = 0; i < 4; ++i) {
        unsigned int count = (unsigned int)counts->base_count[i];  // the reference keeps 32 bits (:94)
        if (console) fprintf(console, "%u", count);
        unsigned long long base_counter = counts->base_total;
        base_probability[i] = (double)count / base_counter;  // double division, widened on store (:520-521)
        if (base_probability[i] == 0.0) {
            if (console) fprintf(console, "Division overflow detected in statistics.\n");
            return FKB_ERR_ZERO_BASE_PROBABILITY;
        }
        if (console) fprintf(console, ", %Lf\n", base_probability[i]);
        if (stats_out) fprintf(stats_out, "%Lf\n", base_probability[i]);
    }
    if (console) fprintf(console, "Found %llu valid bases total INSIDE sequences >= k.\n", (unsigned long long)counts->base_total);
    const uint64_t max_nodes = fkb_max_nodes(k);
    if (console) fprintf(console, "%0.0f%% tree density.\n", ((double)counts->node_count / (double)max_nodes) * 100);
    if (counts->node_count == max_nodes) {
        if (stats_out) fprintf(stats_out, "All possible %dmers combinations were found.\n", k);
        if (console) fprintf(console, "All possible kmer combinations were found.\n");
    } else if (counts->node_count > max_nodes) {
        fprintf(stderr, "Error! too many nodes were created!\nThere may be a corruption of data!\n");
        if (stats_out) fprintf(stats_out, "too many nodes were created when looking for %dmers.\n", k);
    } else {
        if (console) fprintf(console, "FYI we did not find all possible combinations.\n");
        if (stats_out) fprintf(stats_out, "did not find all possible %dmers combinations.\n", k);
    }
    return FKB_OK;
}

namespace {

// per-composition constants of histo_recursive()'s leaf block
struct Composition {
    long double mean;          // n * p                                  (:838)
    long double sd;            // sqrt(n * p * q)                        (:837)
    bool normal_ok;            // n*p >= 5 && n*q >= 5                   (:189-198, :875)
    char text[64];             // ", <h>, <H>" formatted with %LE        (:866-869)
    int text_len;
    bool ready;
};

struct CompositionTable {
    int k;
    unsigned long long n;
    const long double *base_probability;
    std::vector<Composition> comps;  // indexed by (cA * (k+1) + cC) * (k+1) + cG

    CompositionTable(int k_, unsigned long long n_, const long double *bp)
        : k(k_), n(n_), base_probability(bp), comps((size_t)(k_ + 1) * (k_ + 1) * (k_ + 1))
    {
        for (auto &c : comps) c.ready = false;
    }

    void compute(Composition &c, const unsigned cnt[4]) const
    {
        long double prob[4];
        for (int i = 0; i < 4; ++i) prob[i] = (double)cnt[i] / (double)k;  // (:771-774)
        long double h = 0;
        for (int i = 0; i < 4; ++i)
            if (prob[i] != 0) h += (double)prob[i] * log2(1 / (double)prob[i]);  // (:793-801)
        long double H = h * k;  // (:807)
        double estimated = 1;
        for (int i = 0; i < 4; ++i) estimated *= pow((double)base_probability[i], (double)cnt[i]);  // (:819-823)
        long double p = estimated;
        long double q = 1 - p;
        c.sd = sqrtl(n * p * q);
        c.mean = n * p;
        c.normal_ok = (n * p >= 5) && (n * q >= 5);
        c.text_len = snprintf(c.text, sizeof c.text, ", %LE, %LE", h, H);
        c.ready = true;
    }

    // fill every composition up front (cheap: <= 969 entries at k = 16)
    void fill()
    {
        for (unsigned a = 0; a <= (unsigned)k; ++a)
            for (unsigned c = 0; a + c <= (unsigned)k; ++c)
                for (unsigned g = 0; a + c + g <= (unsigned)k; ++g) {
                    unsigned cnt[4] = {a, c, g, (unsigned)k - a - c - g};
                    compute(comps[((size_t)a * (k + 1) + c) * (k + 1) + g], cnt);
                }
    }
};

inline unsigned popc32(uint32_t x) { return (unsigned)__builtin_popcount(x); }

// ", <count>, <z>" of a row depends only on (composition, count): z = (x - n*p) / sd (:839) with n*p and sd fixed per
// composition.  A genome's counts cluster (Poisson around n/4^k on random sequence; a few hundred distinct values per
// composition on real ones), so the two printf("%LE")-class conversions per row -- the bulk of the writer's time -- are
// memoised in a direct-mapped cache per worker thread.  The text is produced by the same snprintf calls as before, so the
// bytes cannot differ.
struct TailEntry {
    uint64_t key;   // (composition index << 32 | count) + 1; 0 = empty
    uint8_t len;
    char text[39];  // ", %d" + ", %LE"  (11 + 14 characters at most... kept generous)
};
struct TailCache {
    static constexpr uint32_t kBits = 14;
    std::vector<TailEntry> e;
    TailCache() : e((size_t)1 << kBits) {
        for (auto &x : e) x.key = 0;
    }
    TailEntry &slot(uint32_t comp, uint32_t freq) { return e[((comp * 0x9E3779B1u) ^ (freq * 0x85EBCA77u)) >> (32 - kBits)]; }
};

// format the rows of table[lo, hi) into out
// tsv: the same row as `kmer \t h \t frequency \t H [\t z] \n` -- tab-separated with the count in column 2, the layout
// mergeFile4GNUPLOT.pl joins on (it splits on tabs, keys on column 0 and carries columns 1 and 2, :13-32)
void format_slice(const uint32_t *table, uint64_t lo, uint64_t hi, int k, const CompositionTable &ct, int z_enable, long double z_thr,
                  std::string &out, uint64_t &rows, TailCache &cache, bool tsv)
{
    static const char letters[4] = {'A', 'C', 'G', 'T'};
    const uint32_t lo_bits = 0x55555555u;
    char row[160];
    for (uint64_t code = lo; code < hi; ++code) {
        const uint32_t freq = table[code];
        if (!freq) continue;
        // composition from the 2-bit digits: digit==3 -> both bits, 2 -> hi only, 1 -> lo only, 0 -> none
        const uint32_t c32 = (uint32_t)code;
        const uint32_t lo1 = c32 & lo_bits, hi1 = (c32 >> 1) & lo_bits;
        const unsigned cT = popc32(lo1 & hi1), cG = popc32(hi1 & ~lo1), cC = popc32(lo1 & ~hi1);
        const unsigned cA = (unsigned)k - cT - cG - cC;
        const uint32_t comp = (uint32_t)(((size_t)cA * (k + 1) + cC) * (k + 1) + cG);
        const Composition &cp = ct.comps[comp];
        unsigned long long x = freq;
        TailEntry &te = cache.slot(comp, freq);
        const uint64_t key = (((uint64_t)comp << 32) | freq) + 1;
        const bool hit = te.key == key;
        if (z_enable != 0 || !hit) {
            long double z = (x - cp.mean) / cp.sd;  // (:839)
            if (!(z_enable == 0 || (z_enable > 0 && fabsl(z) >= z_thr))) continue;  // (:852-854)
            if (!hit) {
                int m = snprintf(te.text, sizeof te.text, ", %d", (int)freq);                        // (:872)  %d of an unsigned
                if (cp.normal_ok) m += snprintf(te.text + m, sizeof te.text - (size_t)m, ", %LE", z);  // (:875-885)
                te.len = (uint8_t)m;
                te.key = key;
            }
        }
        int n = 0;
        row[n++] = '\n';
        for (int i = k - 1; i >= 0; --i) row[n++] = letters[(c32 >> (2 * i)) & 3u];
        memcpy(row + n, cp.text, (size_t)cp.text_len);
        n += cp.text_len;
        memcpy(row + n, te.text, te.len);
        n += te.len;
        if (tsv) {
            // row = '\n' kmer ", " h ", " H ", " freq [", " z]: same text, fields re-ordered and tab-separated
            const char *f[5];
            int fl[5], nf = 0;
            const char *p = row + 1, *end = row + n;
            while (nf < 5) {
                const char *q = p;
                while (q + 1 < end && !(q[0] == ',' && q[1] == ' ')) ++q;
                const bool last = q + 1 >= end;
                f[nf] = p;
                fl[nf++] = (int)((last ? end : q) - p);
                if (last) break;
                p = q + 2;
            }
            static const int order[5] = {0, 1, 3, 2, 4};
            for (int i = 0; i < nf; ++i) {
                if (i) out.push_back('\t');
                out.append(f[order[i]], (size_t)fl[order[i]]);
            }
            out.push_back('\n');
        } else {
            out.append(row, (size_t)n);
        }
        ++rows;
    }
}

}  // namespace

static int write_histogram_impl(FILE *csv_out, int k, const uint32_t *table, const fkb_counts *counts, const long double base_probability[4],
                                int z_threshold_enable, long double z_threshold, int n_threads, uint64_t *rows_written, bool tsv)
{
    if (!csv_out || !table || !counts || !base_probability || k < 1 || k > FKB_MAX_K) return FKB_ERR_BAD_ARG;
    if (rows_written) *rows_written = 0;
    CompositionTable ct(k, counts->n_kmers, base_probability);
    ct.fill();
    const uint64_t entries = (uint64_t)1 << (2 * k);
    if (n_threads <= 0) {
        unsigned hc = std::thread::hardware_concurrency();
        n_threads = hc ? (int)(hc > 32 ? 32 : hc) : 4;
    }
    // Slices small enough to bound memory (a k = 11 table is ~250 MB of text), written strictly in order.  Workers format
    // slices into a small ring of buffers (reused, so only the first lap page-faults); the calling thread writes slice s as
    // soon as it is done and hands its buffer to slice s + ring.
    const uint64_t slice = entries < 65536 ? entries : 65536;
    const uint64_t n_slices = entries / slice;
    if ((uint64_t)n_threads > n_slices) n_threads = (int)n_slices;
    const uint64_t ring = (uint64_t)n_threads * 2;
    std::vector<std::string> bufs(ring);
    std::vector<uint64_t> rows(n_slices, 0);
    std::vector<char> done(n_slices, 0);
    std::vector<TailCache> caches((size_t)n_threads);  // one per worker
    std::mutex mu;
    std::condition_variable cv;
    uint64_t next = 0, written = 0;
    bool abort_all = false;
    auto worker = [&](int t) {
        for (;;) {
            uint64_t s;
            {
                std::unique_lock<std::mutex> lk(mu);
                s = next++;
                if (s >= n_slices) return;
                cv.wait(lk, [&] { return abort_all || s < written + ring; });  // the slice's buffer has been written out
                if (abort_all) return;
            }
            std::string &b = bufs[s % ring];
            b.clear();
            format_slice(table, s * slice, (s + 1) * slice, k, ct, z_threshold_enable, z_threshold, b, rows[s], caches[(size_t)t], tsv);
            {
                std::lock_guard<std::mutex> lk(mu);
                done[s] = 1;
            }
            cv.notify_all();
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t) pool.emplace_back(worker, t);
    uint64_t total_rows = 0;
    int status = FKB_OK;
    for (uint64_t s = 0; s < n_slices; ++s) {
        {
            std::unique_lock<std::mutex> lk(mu);
            cv.wait(lk, [&] { return done[s] != 0; });
        }
        const std::string &b = bufs[s % ring];
        if (!b.empty() && fwrite(b.data(), 1, b.size(), csv_out) != b.size()) status = FKB_ERR_IO;
        total_rows += rows[s];
        {
            std::lock_guard<std::mutex> lk(mu);
            written = s + 1;
            if (status != FKB_OK) abort_all = true;
        }
        cv.notify_all();
        if (status != FKB_OK) break;
    }
    for (auto &th : pool) th.join();
    if (status != FKB_OK) return status;
    if (rows_written) *rows_written = total_rows;
    return FKB_OK;
}

extern "C" int fkb_write_histogram(FILE *csv_out, int k, const uint32_t *table, const fkb_counts *counts,
                                   const long double base_probability[4], int z_threshold_enable, long double z_threshold, int n_threads,
                                   uint64_t *rows_written)
{
    return write_histogram_impl(csv_out, k, table, counts, base_probability, z_threshold_enable, z_threshold, n_threads, rows_written, false);
}

extern "C" int fkb_write_histogram_tsv(FILE *tsv_out, int k, const uint32_t *table, const fkb_counts *counts,
                                       const long double base_probability[4], int z_threshold_enable, long double z_threshold, int n_threads,
                                       uint64_t *rows_written)
{
    return write_histogram_impl(tsv_out, k, table, counts, base_probability, z_threshold_enable, z_threshold, n_threads, rows_written, true);
}
