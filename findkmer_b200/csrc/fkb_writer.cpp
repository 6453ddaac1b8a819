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

// format the rows of table[lo, hi) into out
void format_slice(const uint32_t *table, uint64_t lo, uint64_t hi, int k, const CompositionTable &ct, int z_enable, long double z_thr,
                  std::string &out, uint64_t &rows)
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
        const Composition &cp = ct.comps[((size_t)cA * (k + 1) + cC) * (k + 1) + cG];
        unsigned long long x = freq;
        long double z = (x - cp.mean) / cp.sd;  // (:839)
        if (!(z_enable == 0 || (z_enable > 0 && fabsl(z) >= z_thr))) continue;  // (:852-854)
        int n = 0;
        row[n++] = '\n';
        for (int i = k - 1; i >= 0; --i) row[n++] = letters[(c32 >> (2 * i)) & 3u];
        memcpy(row + n, cp.text, (size_t)cp.text_len);
        n += cp.text_len;
        n += snprintf(row + n, sizeof row - (size_t)n, ", %d", (int)freq);                    // (:872)  %d of an unsigned
        if (cp.normal_ok) n += snprintf(row + n, sizeof row - (size_t)n, ", %LE", z);        // (:875-885)
        out.append(row, (size_t)n);
        ++rows;
    }
}

}  // namespace

extern "C" int fkb_write_histogram(FILE *csv_out, int k, const uint32_t *table, const fkb_counts *counts,
                                   const long double base_probability[4], int z_threshold_enable, long double z_threshold, int n_threads,
                                   uint64_t *rows_written)
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
    // slices small enough to bound memory (a k = 11 table is ~170 MB of text), written strictly in order
    const uint64_t slice = entries < 65536 ? entries : 65536;
    const uint64_t n_slices = entries / slice;
    if ((uint64_t)n_threads > n_slices) n_threads = (int)n_slices;
    const uint64_t wave = (uint64_t)n_threads * 4;
    std::vector<std::string> bufs(wave);
    std::vector<uint64_t> rows(wave);
    uint64_t total_rows = 0;
    for (uint64_t s0 = 0; s0 < n_slices; s0 += wave) {
        const uint64_t s1 = s0 + wave < n_slices ? s0 + wave : n_slices;
        std::atomic<uint64_t> next{s0};
        auto worker = [&] {
            for (;;) {
                uint64_t s = next.fetch_add(1);
                if (s >= s1) return;
                std::string &b = bufs[s - s0];
                b.clear();
                rows[s - s0] = 0;
                format_slice(table, s * slice, (s + 1) * slice, k, ct, z_threshold_enable, z_threshold, b, rows[s - s0]);
            }
        };
        std::vector<std::thread> pool;
        for (int t = 1; t < n_threads; ++t) pool.emplace_back(worker);
        worker();
        for (auto &th : pool) th.join();
        for (uint64_t s = s0; s < s1; ++s) {
            const std::string &b = bufs[s - s0];
            if (!b.empty() && fwrite(b.data(), 1, b.size(), csv_out) != b.size()) return FKB_ERR_IO;
            total_rows += rows[s - s0];
        }
    }
    if (rows_written) *rows_written = total_rows;
    return FKB_OK;
}
