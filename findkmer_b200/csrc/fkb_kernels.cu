// fkb_kernels.cu -- hand-written sm_100a kernels of the k-mer counting path.
//
// Replaces (result-identically, not line by line) the reference's per-byte scan
//   findKmer()                     findKmer/src/findKmer.cpp:962-1069
//   base2int()                     :567-589
//   shift_left_and_insert()        :947-958      -> a 2k-bit rolling register
//   tree_create()/node_branch_enter_and_create()/node_create()  :612-690 -> dense uint32[4^k] table
//
// Input is the STRIPPED stream (include/findkmer_b200.h "stream contract"): no '\n', one '>' per
// header.  Device rule: byte in {A,C,G,T} => code 0..3, anything else => window reset.
//
// Work decomposition: the stream is cut into 64-byte thread chunks; a thread reads its chunk plus the
// 16 bytes in front of it (k <= 16, so 16 bytes of left context decide every window that ends in the
// chunk AND whether a run's length is exactly k -- the reference's `seqSize == k` branch, :1044).
// A window is counted by the thread (and the shard) that owns its LAST byte.
#include "fkb_kernels.cuh"

namespace fkb {

namespace {

constexpr int kThreads = 256;
constexpr int kChunk = 64;  // bytes owned per thread per iteration

// sum_{e<d} 4^e = (4^d - 4) / 3; (4^d - 1) / 3 is the bit pattern 0101..01 (d ones): no division on the event path
__host__ __device__ inline uint64_t flags_offset(int d) { return (0x5555555555555555ull & ((1ull << (2 * d)) - 1ull)) - 1ull; }  // sum_{e<d} 4^e
// look before storing: every run start writes the same few flag bytes, and plain stores from 148 SMs onto one L2 line serialise
// (fkb_stream.cuh set_flag has the measurement); flags only go 0 -> 1, so a stale 0 costs one redundant store
__device__ __forceinline__ void set_flag(uint8_t *flag)
{
    uint32_t v;
    asm volatile("ld.global.ca.u8 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v == 0) *flag = 1;
}

// 128-bit read-only load.  The stream is read once; neighbouring lanes share 32-byte sectors across the
// five loads of an iteration, so L1 allocation is kept (no .no_allocate).
__device__ __forceinline__ uint4 ldg128(const uint8_t *p)
{
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}

// bytes [off, off+16) of the stream with everything outside [0, limit) read as 0 (a reset byte)
__device__ __forceinline__ uint4 load16_guarded(const uint8_t *s, int64_t off, uint64_t limit)
{
    if (off >= 0 && (uint64_t)off + 16 <= limit) return ldg128(s + off);
    uint32_t w[4] = {0, 0, 0, 0};
    for (int i = 0; i < 16; ++i) {
        int64_t p = off + i;
        if (p >= 0 && (uint64_t)p < limit) w[i >> 2] |= (uint32_t)s[p] << (8 * (i & 3));
    }
    return make_uint4(w[0], w[1], w[2], w[3]);
}

__device__ __forceinline__ void red_add_u32(uint32_t *addr, uint32_t v)
{
    asm volatile("red.global.add.u32 [%0], %1;" ::"l"(addr), "r"(v) : "memory");
}

__device__ __forceinline__ unsigned long long warp_sum(unsigned long long v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Per-thread scan state: the rolling register and the current run length (the reference's seqSize,
// findKmer.cpp:977, saturating is unnecessary: a thread sees at most 16 + 64 bytes).
struct Scan {
    uint32_t kmer;
    int run;
};

// ASCII -> 2-bit code for A,C,G,T (A0 C1 G2 T3, base2int :569-576) and validity.
// code = bit1^bit2 : bit1 of (c>>1); letters: A 0x41, C 0x43, G 0x47, T 0x54.
__device__ __forceinline__ uint32_t base_code(uint32_t c) { return ((c >> 1) ^ (c >> 2)) & 3u; }
__device__ __forceinline__ bool base_valid(uint32_t c, uint32_t code) { return c == ((0x54474341u >> (code * 8)) & 0xffu); }

// context-only step (left halo): no counting
__device__ __forceinline__ void step_context(Scan &s, uint32_t c, uint32_t mask)
{
    uint32_t code = base_code(c);
    s.kmer = ((s.kmer << 2) | code) & mask;
    s.run = base_valid(c, code) ? s.run + 1 : 0;
}

struct Tally {
    uint32_t windows;   // positions with run >= k
    uint32_t unknown;   // bytes outside {A,C,G,T,N,'>'}
    uint32_t valid;     // A,C,G,T bytes
};

// Rare events, off the fast path (once per run, and k-1 times per run):
//  run == k : the reference adds all k bases of the first window to baseStatistics (:1050-1056); the LAST
//             base is recovered from the table at finalize, the first k-1 go to head_base here.
//  run <  k : the reference inserts the run so far as a short path (:1059-1062) => a depth-`run` trie node;
//             recorded as a prefix flag, and as a visit of the depth-1 node of the run's first base.
// ev[0..3] head_base, ev[4..7] short_first, ev[8] runs_ge_k: CTA-private shared-memory counters (run boundaries are
// frequent in soft-masked / N-rich genomes; global atomics on nine fixed addresses would serialise them).
__device__ __noinline__ void rare_event(uint32_t kmer, int run, int k, uint8_t *flags, uint32_t *ev)
{
    if (run == k) {
        const uint32_t head = kmer >> 2;  // the first k-1 bases; composition by popcounts of the 2-bit digits
        const uint32_t lo = head & 0x55555555u, hi = (head >> 1) & 0x55555555u;
        const uint32_t cT = __popc(lo & hi), cG = __popc(hi & ~lo), cC = __popc(lo & ~hi), cA = (uint32_t)(k - 1) - cT - cG - cC;
        if (cA) atomicAdd(&ev[0], cA);
        if (cC) atomicAdd(&ev[1], cC);
        if (cG) atomicAdd(&ev[2], cG);
        if (cT) atomicAdd(&ev[3], cT);
        atomicAdd(&ev[8], 1u);
    } else {  // 1 <= run < k
        uint32_t prefix = kmer & ((1u << (2 * run)) - 1u);
        set_flag(flags + flags_offset(run) + prefix);
        atomicAdd(&ev[4 + (prefix >> (2 * (run - 1)))], 1u);
    }
}

// PRIV: `table` is the CTA's private copy of the table in shared memory (k <= 7), otherwise the global table
template <bool EDGE, bool PRIV>
__device__ __forceinline__ void step_owned(Scan &s, uint32_t c, uint32_t mask, int k, uint32_t *table, uint8_t *flags,
                                           uint32_t *P, Tally &t, uint64_t pos, uint64_t begin, uint64_t end)
{
    uint32_t code = base_code(c);
    bool valid = base_valid(c, code);
    s.kmer = ((s.kmer << 2) | code) & mask;
    s.run = valid ? s.run + 1 : 0;
    bool own = EDGE ? (pos >= begin && pos < end) : true;
    if (own) {
        t.valid += valid;
        t.unknown += (!valid && c != 'N' && c != '>');
        if (s.run >= k) {
            if constexpr (PRIV) atomicAdd(table + s.kmer, 1u);
            else red_add_u32(table + s.kmer, 1u);
            t.windows++;
        }
        if ((unsigned)(s.run - 1) < (unsigned)k) rare_event(s.kmer, s.run, k, flags, P);
    }
}

template <bool EDGE, bool PRIV>
__device__ __forceinline__ void scan_word(Scan &s, uint32_t w, uint32_t mask, int k, uint32_t *table, uint8_t *flags,
                                          uint32_t *P, Tally &t, uint64_t pos, uint64_t begin, uint64_t end)
{
#pragma unroll
    for (int j = 0; j < 4; ++j) step_owned<EDGE, PRIV>(s, (w >> (8 * j)) & 0xffu, mask, k, table, flags, P, t, pos + j, begin, end);
}

// ------------------------------------------------------------------------------------------------
// VARIANT_DIRECT: fused encode + count, one red.global.add.u32 per window.
// The 4^k table (16 MiB at k = 11) is L2-resident; the atomics never reach HBM.
// PRIV (k <= 7, ranges of some length): the 4^k counters (<= 64 KiB) are privatised per CTA in shared memory -- with 4096 or
// 16384 bins the global reds of a whole GPU pile up on the same L2 lines (40 Gbases/s at k = 6) -- and added to the global
// table once at the end.
// ------------------------------------------------------------------------------------------------
// A second range [begin2, end2) may ride along (blockIdx.y == 1): the two edge slivers around a bucketed interior.
template <bool PRIV>
__global__ void __launch_bounds__(kThreads) count_direct_kernel(const uint8_t *__restrict__ s, uint64_t begin, uint64_t end, uint64_t begin2,
                                                                uint64_t end2, int k, uint32_t *__restrict__ gtable, uint8_t *__restrict__ flags,
                                                                fkb_partials *__restrict__ P)
{
    if (blockIdx.y == 1) { begin = begin2; end = end2; }
    if (end <= begin) return;
    extern __shared__ __align__(16) uint32_t priv_table[];
    __shared__ uint32_t ev[16];
    if (threadIdx.x < 16) ev[threadIdx.x] = 0;
    uint32_t *table = gtable;
    if constexpr (PRIV) {
        table = priv_table;
        for (uint32_t i = threadIdx.x; i < (1u << (2 * k)); i += blockDim.x) priv_table[i] = 0;
    }
    __syncthreads();
    const uint32_t mask = (k == 16) ? 0xffffffffu : ((1u << (2 * k)) - 1u);
    const uint64_t base = begin & ~15ull;                      // chunks are 16-byte aligned in the stream
    const uint64_t n_chunks = (end - base + kChunk - 1) / kChunk;
    Tally t = {0, 0, 0};

    for (uint64_t c = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; c < n_chunks; c += (uint64_t)gridDim.x * blockDim.x) {  // (any block size up to kThreads)
        const uint64_t p0 = base + c * kChunk;
        const bool edge = (p0 < begin) || (p0 + kChunk > end) || (p0 < 16);
        Scan sc = {0u, 0};
        if (!edge) {
            uint4 h = ldg128(s + p0 - 16);
            uint4 v0 = ldg128(s + p0), v1 = ldg128(s + p0 + 16), v2 = ldg128(s + p0 + 32), v3 = ldg128(s + p0 + 48);
            const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int j = 0; j < 4; ++j) step_context(sc, (hw[i] >> (8 * j)) & 0xffu, mask);
            const uint32_t w[16] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w, v2.x, v2.y, v2.z, v2.w, v3.x, v3.y, v3.z, v3.w};
#pragma unroll
            for (int i = 0; i < 16; ++i) scan_word<false, PRIV>(sc, w[i], mask, k, table, flags, ev, t, p0 + 4 * i, begin, end);
        } else {
            uint4 h = load16_guarded(s, (int64_t)p0 - 16, end);
            const uint32_t hw[4] = {h.x, h.y, h.z, h.w};
            for (int i = 0; i < 4; ++i)
                for (int j = 0; j < 4; ++j) step_context(sc, (hw[i] >> (8 * j)) & 0xffu, mask);
            for (int g = 0; g < 4; ++g) {
                uint4 v = load16_guarded(s, (int64_t)(p0 + 16 * g), end);
                const uint32_t w[4] = {v.x, v.y, v.z, v.w};
                for (int i = 0; i < 4; ++i) scan_word<true, PRIV>(sc, w[i], mask, k, table, flags, ev, t, p0 + 16 * g + 4 * i, begin, end);
            }
        }
    }
    __syncthreads();
    if constexpr (PRIV) {
        for (uint32_t i = threadIdx.x; i < (1u << (2 * k)); i += blockDim.x) {
            const uint32_t v = priv_table[i];
            if (v) red_add_u32(gtable + i, v);
        }
    }
    if (threadIdx.x < 9 && ev[threadIdx.x]) {
        unsigned long long *dst = threadIdx.x < 4 ? &P->head_base[threadIdx.x] : (threadIdx.x < 8 ? &P->short_first[threadIdx.x - 4] : &P->runs_ge_k);
        atomicAdd(dst, (unsigned long long)ev[threadIdx.x]);
    }
    unsigned long long win = warp_sum(t.windows), unk = warp_sum(t.unknown), val = warp_sum(t.valid);
    if ((threadIdx.x & 31) == 0) {
        if (win) atomicAdd(&P->n_windows, win);
        if (unk) atomicAdd(&P->unknown_chars, unk);
        if (val) atomicAdd(&P->valid_bases, val);
    }
}

// ------------------------------------------------------------------------------------------------
// finalize: (table, prefix flags, partials) -> fkb_counts
//   scratch[0]      sum of table               (must equal partials.n_windows, else a counter wrapped)
//   scratch[1..4]   sum of table over k-mers whose LAST  base is b   (baseStatistics, :1041)
//   scratch[5..8]   sum of table over k-mers whose FIRST base is b   (depth-1 trie node visits, :640)
//   scratch[9]      number of trie nodes below the head
// ------------------------------------------------------------------------------------------------
// The trie node at depth d with prefix x exists iff a short run set flags[d][x] or one of its four children exists
// (depth k: table[x] != 0).  One launch takes an INPUT level d_in (LEAF: the table, 4 leaves per thread; otherwise the flag
// bytes of depth d_in, 4 siblings per thread) and resolves three levels with warp ballots -- thread = node at depth
// d_in-1, 32 lanes = 8 nodes at depth d_in-2 = 2 nodes at depth d_in-3 -- counting the nodes of depths d_in, d_in-1 and
// d_in-2 and writing the presence of depth d_in-3, which is the next launch's input: ceil(k/3) launches instead of k.
// The launch that reaches depth 1 is a single block and also assembles fkb_counts.
__device__ void finalize_counts(const unsigned long long *scratch_g, const fkb_partials *__restrict__ P, uint64_t stream_bytes,
                                fkb_counts *__restrict__ out)
{
    unsigned long long scratch[10];
    for (int i = 0; i < 10; ++i) scratch[i] = __ldcg(scratch_g + i);  // written by atomics (some by this very block): read at L2
    fkb_counts c;
    c.n_kmers = scratch[0];
    c.base_total = 0;
    unsigned long long worst_visits = 0;
    for (int b = 0; b < 4; ++b) {
        c.base_count[b] = scratch[1 + b] + P->head_base[b];
        c.base_total += c.base_count[b];
        unsigned long long visits = scratch[5 + b] + P->short_first[b];
        worst_visits = visits > worst_visits ? visits : worst_visits;
    }
    c.node_count = scratch[9] ? scratch[9] + 1 : 0;  // + the lazily created head node (:666-670)
    c.unknown_chars = P->unknown_chars;
    c.stream_bytes = stream_bytes;
    c.runs_ge_k = P->runs_ge_k;
    // rollover (:640-648): a trie node visited 2^32 times.  Depth-1 nodes are visited at least as often as any
    // of their descendants, so they wrap first; a wrapped table shows up as sum(table) != independent window count.
    c.rollover = (worst_visits >= (1ull << 32)) || (scratch[0] != P->n_windows);
    c.valid_bases = P->valid_bases;
    *out = c;
}

template <bool LEAF>
__global__ void __launch_bounds__(kThreads) finalize_levels_kernel(const uint32_t *__restrict__ table, int k, int d_in, uint8_t *__restrict__ flags,
                                                                   unsigned long long *__restrict__ scratch, int is_last,
                                                                   const fkb_partials *__restrict__ P, uint64_t stream_bytes,
                                                                   fkb_counts *__restrict__ out)
{
    const uint64_t n1 = 1ull << (2 * (d_in - 1));        // nodes at depth d_in-1 == groups of 4 siblings at depth d_in
    const uint64_t n_round = (n1 + 31) & ~31ull;         // whole warps: every lane takes part in the ballots
    const uint32_t lane = threadIdx.x & 31;
    unsigned long long tot = 0, last[4] = {0, 0, 0, 0}, first[4] = {0, 0, 0, 0}, nodes = 0;
    for (uint64_t g = (uint64_t)blockIdx.x * kThreads + threadIdx.x; g < n_round; g += (uint64_t)gridDim.x * kThreads) {
        const bool in = g < n1;
        uint32_t children = 0;  // present children of node g
        if (in) {
            if constexpr (LEAF) {
                const uint4 v = reinterpret_cast<const uint4 *>(table)[g];
                const unsigned long long sum = (unsigned long long)v.x + v.y + v.z + v.w;
                tot += sum;
                last[0] += v.x; last[1] += v.y; last[2] += v.z; last[3] += v.w;
                if (k == 1) {
                    first[0] += v.x; first[1] += v.y; first[2] += v.z; first[3] += v.w;
                } else {
                    const uint32_t fb = (uint32_t)(g >> (2 * (k - 2)));
                    first[0] += (fb == 0) ? sum : 0; first[1] += (fb == 1) ? sum : 0;
                    first[2] += (fb == 2) ? sum : 0; first[3] += (fb == 3) ? sum : 0;
                }
                children = (v.x != 0) + (v.y != 0) + (v.z != 0) + (v.w != 0);
            } else {
                const uint32_t v = reinterpret_cast<const uint32_t *>(flags + flags_offset(d_in))[g];
                children = ((v & 0xffu) != 0) + ((v & 0xff00u) != 0) + ((v & 0xff0000u) != 0) + ((v & 0xff000000u) != 0);
            }
        }
        nodes += children;                                                       // depth d_in
        if (d_in < 2) continue;                                                  // depth d_in-1 is the head
        const bool p1 = in && (children != 0 || flags[flags_offset(d_in - 1) + g] != 0);
        nodes += p1;                                                             // depth d_in-1
        const uint32_t m1 = __ballot_sync(0xffffffffu, p1);
        if (d_in < 3) continue;
        const uint64_t g2 = ((g - lane) >> 2) + lane;                            // lanes 0..7: the warp's 8 nodes at depth d_in-2
        const bool p2 = lane < 8 && g2 < (n1 >> 2) && (((m1 >> (4 * lane)) & 0xFu) != 0 || flags[flags_offset(d_in - 2) + g2] != 0);
        nodes += p2;                                                             // depth d_in-2
        const uint32_t m2 = __ballot_sync(0xffffffffu, p2);
        if (d_in < 4) continue;
        const uint64_t g3 = ((g - lane) >> 4) + lane;                            // lanes 0..1: the warp's 2 nodes at depth d_in-3
        if (lane < 2 && g3 < (n1 >> 4) && ((m2 >> (4 * lane)) & 0xFu) != 0) flags[flags_offset(d_in - 3) + g3] = 1;  // next launch's input
    }
    // block-level reduction: one set of global atomics per CTA instead of per warp (they all hit the same 10 words)
    constexpr int NV = LEAF ? 10 : 1;
    unsigned long long vals[10] = {tot, last[0], last[1], last[2], last[3], first[0], first[1], first[2], first[3], nodes};
    if constexpr (!LEAF) vals[0] = nodes;
    __shared__ unsigned long long red[kThreads / 32][10];
#pragma unroll
    for (int i = 0; i < NV; ++i) vals[i] = warp_sum(vals[i]);
    if (lane == 0)
        for (int i = 0; i < NV; ++i) red[threadIdx.x >> 5][i] = vals[i];
    __syncthreads();
    if (threadIdx.x < NV) {
        unsigned long long v = 0;
        for (int w = 0; w < kThreads / 32; ++w) v += red[w][threadIdx.x];
        if (v) atomicAdd(&scratch[LEAF ? threadIdx.x : 9], v);
    }
    if (is_last) {  // single block (<= 16 nodes): every earlier launch has completed, this block's own atomics are ordered by the barrier
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) finalize_counts(scratch, P, stream_bytes, out);
    }
}

// ------------------------------------------------------------------------------------------------
// synthetic FASTA generator: bit-identical twin of findkmer_b200/synth.py::render()
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint64_t splitmix64(uint64_t x)
{
    uint64_t z = x + 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    return z ^ (z >> 31);
}

__constant__ uint32_t c_run_quantiles[16] = {32, 98, 170, 247, 330, 421, 520, 629, 752, 891, 1052, 1242, 1474, 1773, 2197, 3466};

__device__ __forceinline__ uint8_t synth_letter(uint64_t g, uint64_t seed, int n_runs, int soft_mask)
{
    uint8_t out = (uint8_t)((0x54474341u >> (8 * (uint32_t)(splitmix64(seed + g) >> 62))) & 0xffu);
    if (soft_mask) {
        uint64_t slot = g / 3000ull;
        uint64_t h = splitmix64((seed ^ 0xC2B2AE3D27D4EB4Full) + slot);
        uint64_t ln = ((uint64_t)c_run_quantiles[h & 15] * 3ull) / 10ull;
        uint64_t start = (h >> 8) % (3000ull - ln);
        if ((g % 3000ull) - start < ln) out |= 0x20;
    }
    if (n_runs) {
        uint64_t slot = g / 20000ull;
        uint64_t h = splitmix64((seed ^ 0x9E3779B97F4A7C15ull) + slot);
        uint64_t ln = c_run_quantiles[h & 15];
        uint64_t start = (h >> 8) % (20000ull - ln);
        if ((g % 20000ull) - start < ln) out = 'N';
    }
    return out;
}

// Renders bytes [first_byte, first_byte + n_bytes) of the virtual file into out[0, n_bytes).
// line_width > 0: wrapped lines; == 0: one unwrapped line per record; < 0: the STRIPPED layout (no newline at
// all: a record is its header bytes followed by its bases -- with header_len == 1 this is the stream contract).
__global__ void __launch_bounds__(kThreads) synth_fasta_kernel(uint8_t *__restrict__ out, uint64_t first_byte, uint64_t n_bytes, int n_records,
                                                               const uint64_t *__restrict__ rec_off, const uint64_t *__restrict__ rec_base0,
                                                               const uint8_t *__restrict__ headers, int header_len, int line_width,
                                                               uint64_t seed, int n_runs, int soft_mask)
{
    // each thread renders 16 consecutive bytes and stores them as one 128-bit word
    const uint64_t n_groups = (n_bytes + 15) / 16;
    for (uint64_t gi = (uint64_t)blockIdx.x * kThreads + threadIdx.x; gi < n_groups; gi += (uint64_t)gridDim.x * kThreads) {
        uint32_t w[4] = {0, 0, 0, 0};
        uint64_t p = first_byte + gi * 16;
        const uint64_t p_end = first_byte + n_bytes;
        // record of the first byte by binary search; later bytes advance linearly
        int lo = 0, hi = n_records;  // rec_off[lo] <= p < rec_off[hi]
        while (hi - lo > 1) {
            int mid = (lo + hi) >> 1;
            if (rec_off[mid] <= p) lo = mid; else hi = mid;
        }
        int r = lo;
        for (int i = 0; i < 16; ++i, ++p) {
            if (p >= p_end) break;
            while (r + 1 < n_records && p >= rec_off[r + 1]) ++r;
            uint64_t o = p - rec_off[r];
            uint8_t ch;
            if (o < (uint64_t)header_len) {
                ch = headers[(uint64_t)r * header_len + o];
            } else {
                o -= header_len;
                uint64_t nb = rec_base0[r + 1] - rec_base0[r];
                if (line_width < 0) {
                    ch = synth_letter(rec_base0[r] + o, seed, n_runs, soft_mask);
                } else if (line_width == 0) {
                    ch = (o == nb) ? (uint8_t)'\n' : synth_letter(rec_base0[r] + o, seed, n_runs, soft_mask);
                } else {
                    uint64_t q = o / (uint64_t)(line_width + 1), col = o % (uint64_t)(line_width + 1);
                    uint64_t bi = q * line_width + col;
                    ch = (col == (uint64_t)line_width || bi >= nb) ? (uint8_t)'\n' : synth_letter(rec_base0[r] + bi, seed, n_runs, soft_mask);
                }
            }
            w[i >> 2] |= (uint32_t)ch << (8 * (i & 3));
        }
        if (gi * 16 + 16 <= n_bytes) {
            reinterpret_cast<uint4 *>(out)[gi] = make_uint4(w[0], w[1], w[2], w[3]);
        } else {
            for (uint64_t i = 0; gi * 16 + i < n_bytes; ++i) out[gi * 16 + i] = (uint8_t)(w[i >> 2] >> (8 * (i & 3)));
        }
    }
}

inline int grid_for(uint64_t items_per_thread_units, int sm_count, int ctas_per_sm)
{
    uint64_t need = (items_per_thread_units + kThreads - 1) / kThreads;
    uint64_t cap = (uint64_t)sm_count * ctas_per_sm;  // a whole number of waves: multiple of the SM count
    if (need < 1) need = 1;
    return (int)(need < cap ? need : cap);
}

}  // namespace

static cudaError_t launch_direct(const LaunchInfo &li, const uint8_t *d_stream, uint64_t begin, uint64_t end, uint64_t begin2, uint64_t end2,
                                 int k, uint32_t *d_table, uint8_t *d_flags, fkb_partials *d_partials, cudaStream_t st, int *launches,
                                 bool slivers = false)
{
    const uint64_t len1 = end > begin ? end - (begin & ~15ull) : 0, len2 = end2 > begin2 ? end2 - (begin2 & ~15ull) : 0;
    const uint64_t longest = len1 > len2 ? len1 : len2;
    if (!longest) return cudaSuccess;
    if (slivers) {
        // the two edge slivers of a bucketed interior (< 2 KiB each): 64-thread blocks, so that they fit on an SM next to a pass-2
        // CTA (1024 threads x 56 registers) and run on the side stream while pass 2 counts
        dim3 grid((unsigned)((longest + kChunk * 64 - 1) / (kChunk * 64)), len2 ? 2 : 1);
        count_direct_kernel<false><<<grid, 64, 0, st>>>(d_stream, begin, end, begin2, end2, k, d_table, d_flags, d_partials);
        if (launches) ++*launches;
        return cudaGetLastError();
    }
    if (k <= 7 && longest >= (256u << 10)) {  // private counters pay from a few hundred KiB on (every CTA adds its 4^k bins at the end)
        if (k == 7) {  // 64 KiB of dynamic shared memory needs the opt-in (per device, so not cached in a static)
            cudaError_t e = cudaFuncSetAttribute(count_direct_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
            if (e != cudaSuccess) return e;
        }
        dim3 grid((unsigned)grid_for((longest + kChunk - 1) / kChunk, li.sm_count, k == 7 ? 2 : 4), len2 ? 2 : 1);
        count_direct_kernel<true><<<grid, kThreads, sizeof(uint32_t) << (2 * k), st>>>(d_stream, begin, end, begin2, end2, k, d_table, d_flags, d_partials);
    } else {
        dim3 grid((unsigned)grid_for((longest + kChunk - 1) / kChunk, li.sm_count, 8), len2 ? 2 : 1);
        count_direct_kernel<false><<<grid, kThreads, 0, st>>>(d_stream, begin, end, begin2, end2, k, d_table, d_flags, d_partials);
    }
    if (launches) ++*launches;
    return cudaGetLastError();
}

// Below this many interior bytes the fixed costs of the bucketed path (1024 buckets to zero / fold / write out, fold kernels
// for k <= 8) lose against one red per window.  Measured crossovers on B200 (profiles/r01_crossover_direct_bucketed.txt):
// the direct kernel runs at ~180 Gbases/s from k = 9 up (one global red per window) and, with its counters privatised in shared
// memory, at ~430 / ~340 Gbases/s at k = 6 / 7.
uint64_t bucket_min_bytes(int k)
{
    if (k <= 6) return 150ull << 20;
    if (k == 7) return 110ull << 20;
    if (k == 8) return 44ull << 20;
    if (k == 9) return 36ull << 20;
    if (k == 10) return 30ull << 20;
    return 22ull << 20;
}

// The two edge slivers of an interior [lo, hi) go through the direct kernel (it owns every ragged / unaligned / short-halo
// case), in one launch -- on a side stream forked here and joined at the end, so that its ~25 us of latency-bound work
// hides behind the big kernels instead of standing in front of them (2 % of a 3.1 Gbp step, 10 % of a 1/8 shard).
template <typename Interior>
static cudaError_t launch_with_slivers(const LaunchInfo &li, const uint8_t *d_stream, uint64_t begin, uint64_t lo, uint64_t hi, uint64_t end, int k,
                                       uint32_t *d_table, uint8_t *d_flags, fkb_partials *d_partials, cudaStream_t st, int *launches, Interior interior)
{
    if (li.edge_stream) {
        cudaError_t e = cudaEventRecord(li.edge_fork, st);  // the slivers depend on what precedes this call (zeroing), not on the big kernels
        if (e != cudaSuccess) return e;
        e = interior();
        if (e != cudaSuccess) return e;
        e = cudaStreamWaitEvent(li.edge_stream, li.edge_fork, 0);
        if (e != cudaSuccess) return e;
        e = launch_direct(li, d_stream, begin, lo, hi, end, k, d_table, d_flags, d_partials, li.edge_stream, launches, true);
        if (e != cudaSuccess) return e;
        e = cudaEventRecord(li.edge_join, li.edge_stream);
        if (e != cudaSuccess) return e;
        return cudaStreamWaitEvent(st, li.edge_join, 0);
    }
    cudaError_t e = launch_direct(li, d_stream, begin, lo, hi, end, k, d_table, d_flags, d_partials, st, launches, true);
    if (e != cudaSuccess) return e;
    return interior();
}

// Below this many interior bytes the single-pass shared-memory path (k <= 8) loses against the direct kernel: every CTA zeroes
// and finally adds a whole 4^k table.
uint64_t smallk_min_bytes(int k) { return k <= 6 ? (256ull << 10) : (512ull << 10); }

cudaError_t launch_count(const LaunchInfo &li, const uint8_t *d_stream, uint64_t begin, uint64_t end, int k, uint32_t *d_table,
                         uint8_t *d_flags, fkb_partials *d_partials, cudaStream_t st, int *launches)
{
    if (end <= begin) return cudaSuccess;
    // k <= 8: single pass with the whole table in shared memory (AUTO), unless the bucketed / direct kernels are forced
    if (const uint64_t unit = smallk_unit_bytes(k); unit && (li.variant == VARIANT_AUTO || li.variant == VARIANT_SMEM)) {
        const uint64_t lo = (begin + 16 + unit - 1) / unit * unit, hi = end / unit * unit;
        if (hi > lo && hi - lo >= (li.variant == VARIANT_SMEM ? unit : smallk_min_bytes(k)))
            return launch_with_slivers(li, d_stream, begin, lo, hi, end, k, d_table, d_flags, d_partials, st, launches,
                                       [&] { return launch_count_smallk(li, d_stream, lo, hi, k, d_table, d_flags, d_partials, st, launches); });
        return launch_direct(li, d_stream, begin, end, 0, 0, k, d_table, d_flags, d_partials, st, launches);
    }
    // k = 11 (AUTO or forced): 16-mer items counted as two 13-mers (fkb_bucket2.cu); VARIANT_BUCKET keeps the 13-mer path of fkb_bucket.cu
    const bool fmt16 = bucket16_unit_bytes(k) != 0 && (li.variant == VARIANT_AUTO || li.variant == VARIANT_BUCKET16);
    const uint64_t unit = fmt16 ? bucket16_unit_bytes(k) : bucket_unit_bytes(k);
    bool bucket = unit && li.bucket.gbuf && li.variant != VARIANT_DIRECT && li.variant != VARIANT_SMEM && (fmt16 || li.variant != VARIANT_BUCKET16);
    if (bucket && !fmt16 && !bucket_folds_in_shared(k) && !(li.bucket.table_w && li.bucket.fold)) bucket = false;  // k <= 8 needs the W-mer table and the fold scratch
    uint64_t lo = 0, hi = 0;
    if (bucket) {
        // interior [lo, hi): whole units, aligned in absolute stream coordinates, 16 readable bytes on both sides
        lo = (begin + 16 + unit - 1) / unit * unit;
        hi = end >= 16 ? (end - 16) / unit * unit : 0;
        const uint64_t min_bytes = (li.variant == VARIANT_BUCKET || li.variant == VARIANT_BUCKET16) ? unit : bucket_min_bytes(k);
        bucket = hi > lo && hi - lo >= min_bytes;
    }
    if (!bucket) return launch_direct(li, d_stream, begin, end, 0, 0, k, d_table, d_flags, d_partials, st, launches);
    return launch_with_slivers(li, d_stream, begin, lo, hi, end, k, d_table, d_flags, d_partials, st, launches, [&] {
        return fmt16 ? launch_count_bucketed16(li, li.bucket, d_stream, lo, hi, d_table, d_flags, d_partials, st, launches)
                     : launch_count_bucketed(li, li.bucket, d_stream, lo, hi, k, d_table, d_flags, d_partials, st, launches);
    });
}

cudaError_t launch_finalize(const LaunchInfo &li, int k, const uint32_t *d_table, uint8_t *d_flags, const fkb_partials *d_partials,
                            uint64_t stream_bytes, fkb_counts *d_counts, unsigned long long *d_scratch, cudaStream_t st, int *launches)
{
    cudaError_t e = cudaMemsetAsync(d_scratch, 0, sizeof(unsigned long long) * kFinalizeScratchWords, st);
    if (e != cudaSuccess) return e;
    for (int d_in = k; d_in >= 1; d_in -= 3) {  // three trie levels per launch
        const int is_last = d_in <= 3;          // <= 16 threads of work: one block, which also assembles fkb_counts
        const int grid = grid_for(1ull << (2 * (d_in - 1)), li.sm_count, 8);
        if (d_in == k)
            finalize_levels_kernel<true><<<grid, kThreads, 0, st>>>(d_table, k, d_in, d_flags, d_scratch, is_last, d_partials, stream_bytes, d_counts);
        else
            finalize_levels_kernel<false><<<grid, kThreads, 0, st>>>(d_table, k, d_in, d_flags, d_scratch, is_last, d_partials, stream_bytes, d_counts);
        if (launches) ++*launches;
    }
    return cudaGetLastError();
}

// dst += src for the three accumulators of a shard (table: sums; prefix flags: OR; partials: sums).  src may be PEER memory
// (another GPU of the box, mapped over NVLink): the multi-GPU exchange step of fkb_count_fasta_host_gpus.
__global__ void __launch_bounds__(kThreads) accumulate_kernel(uint32_t *__restrict__ dst_table, const uint32_t *__restrict__ src_table, uint64_t n_table,
                                                              uint8_t *__restrict__ dst_flags, const uint8_t *__restrict__ src_flags, uint64_t n_flags,
                                                              fkb_partials *__restrict__ dst_p, const fkb_partials *__restrict__ src_p)
{
    const uint64_t tid = (uint64_t)blockIdx.x * kThreads + threadIdx.x, nth = (uint64_t)gridDim.x * kThreads;
    const uint64_t n4 = n_table / 4;
    for (uint64_t i = tid; i < n4; i += nth) {
        const uint4 a = reinterpret_cast<const uint4 *>(src_table)[i];
        uint4 d = reinterpret_cast<uint4 *>(dst_table)[i];
        d.x += a.x; d.y += a.y; d.z += a.z; d.w += a.w;
        reinterpret_cast<uint4 *>(dst_table)[i] = d;
    }
    for (uint64_t i = n4 * 4 + tid; i < n_table; i += nth) dst_table[i] += src_table[i];
    const uint64_t f16 = n_flags / 16;  // both flag arrays come from cudaMalloc: 16-byte aligned
    for (uint64_t i = tid; i < f16; i += nth) {
        const uint4 a = reinterpret_cast<const uint4 *>(src_flags)[i];
        uint4 d = reinterpret_cast<uint4 *>(dst_flags)[i];
        d.x |= a.x; d.y |= a.y; d.z |= a.z; d.w |= a.w;
        reinterpret_cast<uint4 *>(dst_flags)[i] = d;
    }
    for (uint64_t i = f16 * 16 + tid; i < n_flags; i += nth) dst_flags[i] |= src_flags[i];
    if (tid < sizeof(fkb_partials) / sizeof(unsigned long long))
        reinterpret_cast<unsigned long long *>(dst_p)[tid] += reinterpret_cast<const unsigned long long *>(src_p)[tid];
}

cudaError_t launch_accumulate(const LaunchInfo &li, int k, uint32_t *dst_table, const uint32_t *src_table, uint8_t *dst_flags, const uint8_t *src_flags,
                              fkb_partials *dst_p, const fkb_partials *src_p, cudaStream_t st, int *launches)
{
    const uint64_t n_table = 1ull << (2 * k);
    uint64_t n_flags = ((1ull << (2 * k)) - 4) / 3;
    if (n_flags < 16) n_flags = 16;
    accumulate_kernel<<<grid_for((n_table + 3) / 4, li.sm_count, 8), kThreads, 0, st>>>(dst_table, src_table, n_table, dst_flags, src_flags, n_flags, dst_p, src_p);
    if (launches) ++*launches;
    return cudaGetLastError();
}

cudaError_t launch_synth(const LaunchInfo &li, uint8_t *d_out, uint64_t first_byte, uint64_t n_bytes, int n_records,
                         const uint64_t *d_rec_offsets, const uint64_t *d_rec_base0, const uint8_t *d_headers, int header_len, int line_width,
                         uint64_t seed, int n_runs, int soft_mask, cudaStream_t st, int *launches)
{
    if (n_bytes == 0) return cudaSuccess;
    synth_fasta_kernel<<<grid_for((n_bytes + 15) / 16, li.sm_count, 8), kThreads, 0, st>>>(
        d_out, first_byte, n_bytes, n_records, d_rec_offsets, d_rec_base0, d_headers, header_len, line_width, seed, n_runs, soft_mask);
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace fkb
