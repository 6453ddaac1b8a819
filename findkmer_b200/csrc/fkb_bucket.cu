// fkb_bucket.cu -- VARIANT_BUCKET: the count path for large inputs, built around what profiles/
// r01_ubench_atomics measured on B200: random global `red`s stop at ~190 Gop/s, shared-memory atomics run at
// ~2600 Gop/s.  So the k-mers are ROUTED to a CTA that owns their bins in shared memory, and the number of
// routed items is cut by counting longer words:
//
//   Instead of one update per k-mer, count W-mers (W = 13 >= k) at every S-th stream position (S = W - k + 1,
//   positions p == 0 mod S).  A fully valid W-window at p contains the S k-mers that start at p .. p+S-1, so the
//   k-mer table is recovered EXACTLY by a fold:  T_k[x] = sum over offsets t < S and all (t-base prefixes a,
//   (S-1-t)-base suffixes b) of T_W[a x b],  plus the few "leftover" k-mers whose covering W-window is broken
//   by a reset or a range edge (those go straight to T_k with a global red).  At k = 11 that is 3x fewer items.
//
//   pass 1  bucketize : every warp streams a contiguous region of the stream, software-pipelined (128-bit loads).
//                       Fused ASCII->2-bit encode + validity (SIMD in 32-bit registers, validity accumulated over
//                       the lane's chunk and tested once) + window extraction; item = 26-bit W-mer code split into
//                       a 10-bit bucket and a 16-bit payload (see "item format").  Items are staged per bucket in
//                       shared memory (the shared atomicAdd returns the slot) and appended, in 16-byte chunks, to a
//                       region of HBM that is private to (bucket, CTA) -- no global atomics on the routing path.
//   pass 2  count     : one CTA at a time owns a bucket's 65536 16-bit counters (128 KiB of shared memory), streams
//                       the bucket's payloads with 128-bit loads and counts with shared-memory atomics; a counter
//                       that passes 0x8000 drains exactly to T_k.  k >= 9 (core buckets): the counters are folded
//                       to k-mer sums in shared memory and added to T_k.  k <= 8: the sub-table goes to T_W and
//   pass 3  fold      : T_k[x] += gathered sums of T_W (coalesced reads, no atomics).
//
// Semantics are those of the reference's scan (findKmer/src/findKmer.cpp:962-1069) exactly; the rare per-run
// events (seqSize == k, :1044-1057; seqSize < k, :1059-1062) are derived from the validity bit masks.
// profiles/r01_ncu_bucketize_v1_by_line.txt is the instruction profile of the first version of pass 1 that this
// layout answers (38 % of instructions in the flush copy loop, 15 % in run-mask loops, a barrier per iteration);
// profiles/r01_ncu_full_summary.txt is where it stands: both passes bound by the shared-memory data pipe (bank
// conflicts of random 32-lane accesses, 3.5 wavefronts per atomic / 16-bit store).
#include <atomic>
#include <type_traits>

#include "fkb_kernels.cuh"
#include "fkb_stream.cuh"

#ifndef FKB_P1_PREFETCH
#define FKB_P1_PREFETCH 2   // L2 prefetch distance of pass 1 in iterations beyond the register pipeline (profiles/r02_prefetch.txt)
#endif

namespace fkb {

namespace {

constexpr int kW = 13;                       // counted word length
constexpr int kBucketBits = 2 * kW - 16;     // 10
constexpr int kNB = 1 << kBucketBits;        // 1024 buckets
#ifndef FKB_STAGE_CAP
#define FKB_STAGE_CAP 104
#endif
#ifndef FKB_P1_MINBLOCKS
#define FKB_P1_MINBLOCKS 1
#endif
constexpr int kStageCap = FKB_STAGE_CAP;       // staged items per bucket (13 chunks of 8): 1024 rows of 208 B = 208 KiB
#ifndef FKB_TILE_AVG
#define FKB_TILE_AVG 64
#endif
#ifndef FKB_P1_THREADS_LO
#define FKB_P1_THREADS_LO 512  // measured: 512 > 640 > 768 > 1024 at S = 3 (registers beat warps: 80-register builds spill)
#endif
#ifndef FKB_P1_THREADS_HI
#define FKB_P1_THREADS_HI 448  // S = 4, 5
#endif
#ifndef FKB_P2_THREADS
#define FKB_P2_THREADS 1024
#endif
constexpr int kTileItems = FKB_TILE_AVG * kNB;         // items a CTA stages between flushes: ~64 per bucket on average (+ <= 7 carried) of 104
constexpr int kP2Threads = FKB_P2_THREADS;

// pass-1 CTA size (1 CTA per SM: the staging rows take 208 KiB), tuned on B200 with build.build_variant()
#ifndef FKB_P1_THREADS_XL
#define FKB_P1_THREADS_XL 384  // S >= 6 (k <= 8): the pipeline state is 10*S registers; 384 threads may use 168 each (448/512: 128 -> spills): k = 6 -14 %, k = 7 -9 %
#endif
template <int S> struct P1Cfg { static constexpr int kThreads = (S <= 3) ? FKB_P1_THREADS_LO : (S <= 5 ? FKB_P1_THREADS_HI : FKB_P1_THREADS_XL); };

// Staging rows are 208 bytes (13 chunks of 16 bytes) apart: 52 words == 20 banks, so consecutive rows start on 8
// different bank offsets and a warp's stores (all rows fill at about the same rate) spread over all 32 banks, while
// every chunk stays 16-byte aligned for the 128-bit flush copies.
__device__ __forceinline__ uint32_t stage_slot(uint32_t bucket, uint32_t pos) { return bucket * kStageCap + pos; }

// ---- item format -----------------------------------------------------------------------------------------------
// A W-mer is 26 bits  [ a : A = 2(S-1) bits | core : 10 bits | r : R = 16 - A bits ].  With S <= kCoreMaxS the BUCKET is the
// core -- the 5 bases that every one of the S k-mers inside the W-mer contains -- and the payload is (a, r).  All W-mers that
// contribute to one k-mer x at offset t then sit in ONE bucket (core = x's bases S-1-t .. S+3-t), so pass 2 folds its 65536
// counters to k-mer counts in shared memory and adds S * 4^(K-5) sums to T_k: the 4^W-entry W-mer table and the fold
// kernels over it disappear.  For larger S (k <= 8) the k-mers of a W-mer share fewer than 5 bases: bucket = top 10 bits,
// payload = low 16 bits, and the W-mer table is folded by the fold kernels below.
#ifndef FKB_CORE_MAX_S
#define FKB_CORE_MAX_S 5
#endif
constexpr int kCoreMaxS = FKB_CORE_MAX_S;
template <int S> struct ItemFmt {
    static constexpr bool kCore = (S <= kCoreMaxS) && (S <= 5);
    __host__ __device__ static constexpr int a_bits() { return kCore ? 2 * (S - 1) : 0; }  // bits of the W-mer above the bucket bits
    __host__ __device__ static constexpr int r_bits() { return 16 - a_bits(); }             // bits below them
    __host__ __device__ static constexpr uint32_t rmask() { return (1u << r_bits()) - 1u; }
    __device__ __forceinline__ static uint32_t wmer(uint32_t bucket, uint32_t payload)  // (bucket, payload) -> 26-bit W-mer code
    {
        if constexpr (kCore) return ((payload >> r_bits()) << (r_bits() + 10)) | (bucket << r_bits()) | (payload & rmask());
        else return (bucket << 16) | payload;
    }
};

// S k-mers of one W-mer: every exact escape (staging overflow, region overflow, counter drain) goes through here
template <int S>
__device__ __noinline__ void red_kmers_of_word(uint32_t wcode, uint32_t *table_k, uint32_t amount)
{
    constexpr int K = kW - S + 1;
    constexpr uint32_t kmask = (1u << (2 * K)) - 1u;
#pragma unroll
    for (int t = 0; t < S; ++t) red_add_u32(table_k + ((wcode >> (2 * (S - 1 - t))) & kmask), amount);
}

// ------------------------------------------------------------------------------------------------
// pass 1: bucketize.  The interior starts at `lo` (a multiple of 16*S in absolute stream coordinates) and is
// n_witers warp-iterations of 32 * 16*S bytes long; 16 readable bytes exist on both sides of it.
// ------------------------------------------------------------------------------------------------
struct P1Smem {
    uint16_t stage[kNB * kStageCap];
    uint32_t cursor[kNB];  // items staged in the row
    uint32_t junk[32];     // directly behind cursor[]: one junk cursor per lane, the target of the slot atomic of an item that is NOT emitted
    uint32_t goff[kNB];    // items already appended to the FRONT part of this CTA's region of the bucket (multiple of 8 until the end)
    uint32_t gback[kNB];   // items written straight to the BACK part of the region because the staging row was full (skewed input)
    uint32_t ev[16];       // per-run event counters (see warp_group_events), flushed to the partials once per tile
};

// shared-memory atomic / predicated store on 32-bit shared addresses: no branches, so the 8 atomics of a batch issue
// back to back and their latencies overlap (the compiler otherwise serialises `if (p) pos = atomicAdd(...)` chains)
// a * b + c as ONE integer multiply-add (FMA pipe); written in PTX so that the address arithmetic of the staging code is not
// re-associated into shift + mask pairs on the ALU pipe, which is the busier one in pass 1
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}
__device__ __forceinline__ uint32_t atoms_inc(uint32_t saddr)
{
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(saddr) : "memory");
    return old;
}
__device__ __forceinline__ void sts16_if(uint32_t saddr, uint32_t value, uint32_t pred)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u16 [%0], %1;\n\t}" ::"r"(saddr), "h"((uint16_t)value), "r"(pred) : "memory");
}

// staging-row overflow of one chunk's items (bit n of `ovf` = n-th static slot): exact escape, off the hot path
template <int S>
__device__ __noinline__ void escape_slots(uint32_t ovf, const uint32_t *codes /* S+1 words: group before + own groups */, uint32_t *table_k)
{
    constexpr uint32_t WMASK = (1u << (2 * kW)) - 1u;
    int n = 0;
    for (int g = 1; g <= S; ++g)
        for (int i = 0; i < 16; ++i) {
            if ((16 * (g - 1) + i) % S != (kW - 1) % S) continue;
            if (ovf & (1u << n)) red_kmers_of_word<S>(__funnelshift_r(codes[g], codes[g - 1], 2 * (15 - i)) & WMASK, table_k, 1u);
            ++n;
        }
}

template <int S>
__device__ __noinline__ void escape_chunk(uint32_t bucket, uint4 v, uint32_t *table_k)
{
    const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
    for (int e = 0; e < 8; ++e) red_kmers_of_word<S>(ItemFmt<S>::wmer(bucket, (w4[e >> 1] >> (16 * (e & 1))) & 0xFFFFu), table_k, 1u);
}

template <int S>
__global__ void __launch_bounds__(P1Cfg<S>::kThreads, FKB_P1_MINBLOCKS)
bucketize_kernel(const uint8_t *__restrict__ s, uint64_t lo, uint64_t n_witers, uint16_t *__restrict__ gbuf, uint32_t cap_cb, uint32_t cap_front,
                 uint32_t *__restrict__ gcount, uint32_t *__restrict__ bucket_total, uint32_t *__restrict__ table_k, uint8_t *__restrict__ flags,
                 fkb_partials *__restrict__ P)
{
    static_assert(S >= 1 && S <= 8, "stride");
    constexpr int kP1Threads = P1Cfg<S>::kThreads, kP1Warps = kP1Threads / 32;
    constexpr int kTileIters = kTileItems / (kP1Threads * 16);  // warp iterations between flushes
    constexpr int K = kW - S + 1;
    constexpr int CH = 16 * S;                     // bytes per lane per iteration: exactly 16 items
    constexpr int J = S - 1;
    constexpr uint64_t WSPAN = 32ull * CH;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    P1Smem &sm = *reinterpret_cast<P1Smem *>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t cursor_sa = (uint32_t)__cvta_generic_to_shared(sm.cursor), stage_sa = (uint32_t)__cvta_generic_to_shared(sm.stage);

    for (int b = threadIdx.x; b < kNB; b += kP1Threads) { sm.cursor[b] = 0; sm.goff[b] = 0; sm.gback[b] = 0; }
    if (threadIdx.x < 16) sm.ev[threadIdx.x] = 0;
    __syncthreads();

    // this warp's contiguous share of the interior (the host keeps n_witers below 2^32: run_bucketed)
    const uint64_t n_warps = (uint64_t)gridDim.x * kP1Warps, gw = (uint64_t)blockIdx.x * kP1Warps + warp;
    const uint64_t q = n_witers / n_warps, rem = n_witers % n_warps;
    const uint32_t my_iters = (uint32_t)(q + (gw < rem ? 1 : 0));
    const uint64_t my_first = gw * q + (gw < rem ? gw : rem);
    const uint32_t max_iters = (uint32_t)(q + (rem ? 1 : 0));  // CTA-uniform trip count (barriers)
    const uint64_t region = lo + my_first * WSPAN, hi = lo + n_witers * WSPAN;
    uint16_t *const my_gbuf = gbuf + (uint64_t)blockIdx.x * cap_cb;  // + bucket * gridDim.x * cap_cb
    // the two chunks of the interior whose windows meet a range edge: lane 0 of the first iteration, lane 31 of the last
    const bool edge_first = (lane == 0) && (region == lo), edge_last = (lane == 31) && (region + (uint64_t)my_iters * WSPAN == hi);

    uint32_t t_unknown = 0, t_dummy = 0, n_fast = 0;
    unsigned long long t_windows = 0, t_valid = 0;

    // ---- software pipeline: `cur` = packed groups of iteration it, `nxt` = of it+1, `raw` = loads of it+2 in flight ----
    // Group g of iteration `it` of this lane lies at region + it*WSPAN + lane*CH + 16*g.  Iteration my_iters consists of lane 0's
    // first group only (the right halo of the region); beyond it there is nothing (zero bytes: not bases).
    Group cur[S], nxt[S], carry;   // carry (lane 0 only): the 16 bytes in front of lane 0's chunk = previous iteration's lane 31, last group
    uint4 raw[S];
    const uint8_t *const lane_base = s + region + (uint64_t)lane * CH;
    auto load_group = [&](uint32_t it, int g) -> uint4 {
        const bool p = (it < my_iters) || (g == 0 && it == my_iters && lane == 0);
        return ldg128_if(lane_base + (uint64_t)it * WSPAN + 16 * g, p);
    };
    {
        carry = pack_group(ldg128(s + region - 16), t_dummy);  // left halo of the region (same address for all lanes)
#pragma unroll
        for (int g = 0; g < S; ++g) cur[g] = pack_group(load_group(0, g), my_iters > 0 ? t_unknown : t_dummy);
#pragma unroll
        for (int g = 0; g < S; ++g) nxt[g] = pack_group(load_group(1, g), my_iters > 1 ? t_unknown : t_dummy);
#pragma unroll
        for (int g = 0; g < S; ++g) raw[g] = load_group(2, g);
    }

    // Tiles (the iterations between two flushes) are kTileIters long, except the first one: CTA i starts with 1 + i mod kTileIters
    // iterations, so that the 148 CTAs -- which all have the same amount of work -- do not flush at the same moment (a flush is
    // a burst of 128 KiB of writes per CTA; in lockstep the bursts collide in L2/HBM while the write path idles in between).
#ifndef FKB_STAGGER
#define FKB_STAGGER 1
#endif
    for (uint32_t tile0 = 0, tile_len = FKB_STAGGER ? 1u + blockIdx.x % (uint32_t)kTileIters : (uint32_t)kTileIters; tile0 < max_iters;
         tile0 += tile_len, tile_len = kTileIters) {
        const uint32_t tile_end = min(tile0 + tile_len, my_iters);
        for (uint32_t it = tile0; it < tile_end; ++it) {
            // ---- neighbours: grp[0] = 16 bytes before my chunk, grp[1..S] = my chunk, grp[S+1] = 16 bytes after it.
            //      Two rotations: lane L takes the last group of lane L-1 (lane 0 receives lane 31's, which is the NEXT iteration's
            //      carry) and the first group of lane L+1 (lane 31 receives lane 0's first group of the next iteration). ----
            Group grp[S + 2];
#pragma unroll
            for (int g = 0; g < S; ++g) grp[g + 1] = cur[g];
            const uint32_t up_c = __shfl_sync(0xffffffffu, cur[S - 1].code, (lane + 31) & 31);
            const uint32_t dn_c = __shfl_sync(0xffffffffu, lane == 0 ? nxt[0].code : cur[0].code, (lane + 1) & 31);
            grp[0].code = lane == 0 ? carry.code : up_c;
            grp[S + 1].code = dn_c;
            const bool is_lo = edge_first && it == 0, is_hi = edge_last && it + 1 == my_iters;

            uint32_t own = cur[0].valid;
#pragma unroll
            for (int g = 1; g < S; ++g) own &= cur[g].valid;
            if (lane == 0) own &= carry.valid & nxt[0].valid;  // the two neighbours that are not another lane's own groups
            uint32_t emit[S + 2];
            // warp-uniform choice: if any lane needs the general path the whole warp takes it (it is correct for clean
            // chunks too); a per-lane branch would execute both paths back to back on soft-masked / N-rich input
            const bool all_emit = __all_sync(0xffffffffu, own == 0xFFFFu && !(is_lo || is_hi));
            if (all_emit) {
                // ---- fast path: 16*(S+2) valid bases around every lane: every anchored window exists, nothing is left over ----
#pragma unroll
                for (int g = 0; g < S + 2; ++g) emit[g] = 0;
                ++n_fast;
                carry.code = up_c;  // lane 0: lane 31's last group
                carry.valid = 0xFFFFu;
            } else {
                // ---- general path: window masks per group; emit = fully valid W-window ending at an anchored position ----
                const uint32_t up_v = __shfl_sync(0xffffffffu, cur[S - 1].valid, (lane + 31) & 31);
                const uint32_t dn_v = __shfl_sync(0xffffffffu, lane == 0 ? nxt[0].valid : cur[0].valid, (lane + 1) & 31);
                grp[0].valid = lane == 0 ? carry.valid : up_v;
                grp[S + 1].valid = dn_v;
                carry.code = up_c;
                carry.valid = up_v;
#pragma unroll
                for (int g = 1; g <= S + 1; ++g) {
                    const uint32_t m = (grp[g - 1].valid << 16) | grp[g].valid;
                    uint32_t phase = 0;  // anchored ends: (16*(g-1) + i) == W-1 (mod S), i = byte in group, bit = 15 - i
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if ((16 * (g - 1) + i) % S == (kW - 1) % S) phase |= 1u << (15 - i);
                    emit[g] = runs_of<kW>(m) & phase;
                }
                if (is_hi) emit[S + 1] = 0;               // windows ending beyond hi belong to nobody here: their k-mers are leftovers
                if (is_lo) emit[1] &= (0xFFFFu >> J);     // a window whose first covered k-mer ends before lo is not ours
#pragma unroll
                for (int g = 1; g <= S; ++g) {
                    const uint32_t m = (grp[g - 1].valid << 16) | grp[g].valid;
                    const uint32_t rk = runs_of<K>(m) & 0xFFFFu;
                    t_windows += __popc(rk);
                    t_valid += __popc(grp[g].valid);
                    const uint32_t E = (emit[g] << 16) | emit[g + 1];  // k-mers covered by an emitted window: it ends 0..J bytes later
                    uint32_t cov = E;
#pragma unroll
                    for (int d = 1; d <= J; ++d) cov |= (E << d);
                    const uint32_t left = rk & ~(cov >> 16);
                    const uint32_t first_k = rk & ~(m >> K) & 0xFFFFu;   // empty when m is all ones
                    const uint32_t shorts = grp[g].valid & ~rk;
                    const uint32_t who = __ballot_sync(0xffffffffu, (left | first_k | shorts) != 0);
                    if (who) warp_group_events(who, left, first_k, shorts, m, grp[g - 1].code, grp[g].code, K, flags, sm.ev, table_k);
                }
            }

            // ---- the 16 items of this chunk: static slots (slot n ends at chunk offset (W-1)%S + n*S), two batches of 8:
            //      8 shared atomics back to back (slot in the bucket's staging row), then 8 stores.  The W-mer is LEFT-aligned in
            //      the funnel-shifted word f (bits 31..6; the low 6 bits are the next bases): the bucket is f >> 22 with no mask,
            //      the two shared addresses are one IMAD each, and st.shared.u16 of f >> 6 stores the payload.
            //      BETWEEN the atomics and the stores that wait for their results, the lane encodes its groups of iteration it+2
            //      (ALU/FMA work that hides the shared-memory latency) and re-issues their loads for it+3. ----
            uint32_t ovf = 0;
            ValidAcc va;
            Group enc[S];
            const bool ld_full = it + 3 < my_iters, ld_halo = (it + 3 == my_iters) && lane == 0;
            const uint8_t *const p3 = lane_base + (uint64_t)(it + 3) * WSPAN;
#if FKB_P1_PREFETCH > 0
            // the lines of iteration it + 3 + D go to L2 now (no register, no scoreboard): see fkb_bucket2.cu FKB2_PREFETCH
            // (measured per stride, 3.1 Gbp: k = 13 / 12 (S = 1, 2) 5.91 -> 5.38 / 3.09 -> 3.00 ms; k = 10 / 9 (S = 4, 5) 1.89 -> 1.95 / 1.68 -> 1.66 ms:
            // the long iterations of S >= 4 already cover the latency, so the prefetch is on for the short ones only)
            if (S <= 2 && lane < (int)((WSPAN + 127) / 128) && it + 3u + FKB_P1_PREFETCH < my_iters)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(s + region + (uint64_t)(it + 3u + FKB_P1_PREFETCH) * WSPAN + (uint64_t)lane * 128u));
#endif
            auto stage_items = [&](auto all_t) {
                constexpr bool ALL = decltype(all_t)::value;  // every slot is emitted: no predicates at all
                constexpr int o0 = (kW - 1) % S;
#pragma unroll
                for (int h = 0; h < 2; ++h) {
                    uint32_t f[8], bk[8], pos[8];
#pragma unroll
                    for (int n = 8 * h; n < 8 * h + 8; ++n) {
                        const int o = o0 + n * S, g = o / 16 + 1, i = o % 16;
                        const int sh = 2 * (15 - i);  // the W-mer ending at byte i of group g is bits [sh+25 : sh] of (grp[g-1].code : grp[g].code)
                        if constexpr (ItemFmt<S>::kCore) {  // right-aligned: bits 25..0; the core starts A bits below the top
                            f[n - 8 * h] = __funnelshift_r(grp[g].code, grp[g - 1].code, sh);
                            bk[n - 8 * h] = (f[n - 8 * h] << (6 + ItemFmt<S>::a_bits())) >> 22;
                        } else {                            // left-aligned: bits 31..6; the bucket is the top 10 bits
                            f[n - 8 * h] = (sh >= 6) ? __funnelshift_r(grp[g].code, grp[g - 1].code, sh - 6) : (grp[g].code << (6 - sh));
                            bk[n - 8 * h] = f[n - 8 * h] >> 22;
                        }
                        const uint32_t ca = mad_u32(bk[n - 8 * h], 4u, cursor_sa);  // &cursor[bucket]
                        if constexpr (ALL) pos[n - 8 * h] = atoms_inc(ca);
                        else {
                            // An item that is not emitted still performs its atomic -- on this lane's junk cursor -- so that the 8 atomics
                            // of the batch stay unconditional and issue back to back (ptxas turns a predicated atomic that returns a value
                            // into a branch, which serialises the batch on soft-masked / N-rich input).
                            const uint32_t e = emit[g] & (1u << (15 - i));
                            const uint32_t ps = atoms_inc(e ? ca : cursor_sa + 4u * (uint32_t)(kNB + lane));
                            pos[n - 8 * h] = e ? ps : 0xFFFFFFFFu;  // not emitted: no slot
                        }
                    }
#pragma unroll
                    for (int g = 0; g < S; ++g)
                        if (g * 2 / S == h) {
                            enc[g].code = pack_codes_fast(raw[g], va);
                            enc[g].valid = 0xFFFFu;
                            raw[g] = ldg128_if(p3 + 16 * g, ld_full || (g == 0 && ld_halo));
                        }
                    uint32_t top = 0;  // max over the batch of (slot + 1); a slot that was not emitted contributes 0
#pragma unroll
                    for (int n = 0; n < 8; ++n) {
                        const uint32_t ps = pos[n];
                        const uint32_t sa = mad_u32(ps, 2u, mad_u32(bk[n], 2u * kStageCap, stage_sa));
                        uint32_t pay;  // only the low 16 bits are stored
                        if constexpr (ItemFmt<S>::kCore) pay = (f[n] & ItemFmt<S>::rmask()) | ((f[n] >> 10) & ~ItemFmt<S>::rmask());  // (a, r)
                        else pay = f[n] >> 6;
                        sts16_if(sa, pay, ps < (uint32_t)kStageCap);
                        top = max(top, ps + 1u);
                    }
                    if (top > (uint32_t)kStageCap) {
                        // Some staging row is full (a bucket that is more popular than the tile size allows for: skewed base
                        // composition, repeats).  The item goes straight to the BACK part of this CTA's region of the bucket:
                        // one more shared atomic and a 2-byte global store instead of a staged slot.  Only when that part is
                        // full too is the item escaped exactly with global reds (ovf).
#pragma unroll
                        for (int n = 0; n < 8; ++n) {
                            if (pos[n] != 0xFFFFFFFFu && pos[n] >= (uint32_t)kStageCap) {
                                const uint32_t q = atomicAdd(&sm.gback[bk[n]], 1u);
                                uint32_t pay;
                                if constexpr (ItemFmt<S>::kCore) pay = (f[n] & ItemFmt<S>::rmask()) | ((f[n] >> 10) & ~ItemFmt<S>::rmask());
                                else pay = f[n] >> 6;
                                if (q < cap_cb - cap_front) my_gbuf[(uint64_t)bk[n] * gridDim.x * cap_cb + cap_front + q] = (uint16_t)pay;
                                else ovf |= 1u << (n + 8 * h);
                            }
                        }
                    }
                }
            };
            if (all_emit) stage_items(std::true_type{});
            else stage_items(std::false_type{});
            if (ovf) {  // exact escape of the items that found their staging row full
                uint32_t codes[S + 1];
#pragma unroll
                for (int g = 0; g <= S; ++g) codes[g] = grp[g].code;
                escape_slots<S>(ovf, codes, table_k);
            }
            if (va.bad()) {  // some byte of this lane's 16*S is not a base: exact per-byte masks (and the unknown-character count) from a re-read
#pragma unroll
                for (int g = 0; g < S; ++g) enc[g] = pack_group(load_group(it + 2, g), (it + 2 < my_iters) ? t_unknown : t_dummy);
            }

            // ---- advance the pipeline ----
#pragma unroll
            for (int g = 0; g < S; ++g) { cur[g] = nxt[g]; nxt[g] = enc[g]; }
        }

        // ---- flush: append whole 16-byte chunks of every staged row to this CTA's region of the bucket ----
        __syncthreads();
        if (threadIdx.x < 9) {  // per-run event counters of this tile (32-bit in shared memory) -> 64-bit partials
            const uint32_t v = sm.ev[threadIdx.x];
            if (v) {
                sm.ev[threadIdx.x] = 0;
                unsigned long long *dst = threadIdx.x < 4 ? &P->head_base[threadIdx.x] : (threadIdx.x < 8 ? &P->short_first[threadIdx.x - 4] : &P->runs_ge_k);
                atomicAdd(dst, (unsigned long long)v);
            }
        }
        {
            // A warp takes 32 buckets at a time.  (A) one lane per bucket: fill, region offset and the <= 7 items that stay
            // staged; (B) 8 lanes per bucket, 4 buckets per step: 128-bit copies of the whole chunks (fill and offset come
            // from the owning lane in one shuffle); (C) one lane per bucket again: the tail goes to the front of the row.
            // Scalar work costs one shared-memory wavefront per 32 buckets this way instead of one per 4.
            const uint32_t sub = lane >> 3, c = lane & 7;
            uint32_t n_seg;  // gridDim.x, read here so that the flush's invariants do not occupy registers across the streaming loop
            asm volatile("mov.u32 %0, %%nctaid.x;" : "=r"(n_seg));
            const uint64_t bstride = (uint64_t)n_seg * cap_cb;
            const uint32_t cap8 = cap_front & ~7u;  // the flushed chunks fill the front part of the region
            for (uint32_t b0 = warp * 32; b0 < (uint32_t)kNB; b0 += kP1Warps * 32) {
                const uint32_t bl = b0 + lane;
                const uint32_t cnt = min(sm.cursor[bl], (uint32_t)kStageCap);
                const uint32_t n8 = cnt & ~7u, off = sm.goff[bl];  // off is a multiple of 8, <= cap8 < 2^31
                const uint32_t ncp = min(n8, cap8 - off);          // items of whole chunks that still fit this CTA's region
                const bool has_tail = n8 && cnt > n8;
                uint4 tail = make_uint4(0, 0, 0, 0);
                if (has_tail) tail = *reinterpret_cast<const uint4 *>(&sm.stage[bl * kStageCap + n8]);
                const uint32_t packed = (off >> 3) | ((ncp >> 3) << 28);
#pragma unroll
                for (int jj = 0; jj < 8; jj += 4) {  // 4 steps of 4 buckets at a time: shuffles, then all loads, then all stores
                    uint32_t pk[4];
                    uint4 v[8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) pk[j] = __shfl_sync(0xffffffffu, packed, 4 * (jj + j) + sub);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t ncb = (pk[j] >> 28) << 3;
                        const uint32_t row_sa = stage_sa + ((b0 + 4 * (jj + j) + sub) * kStageCap + c * 8u) * 2u;  // this lane's chunk of the row
                        v[2 * j] = lds128_if(row_sa, c * 8u < ncb);
                        v[2 * j + 1] = lds128_if(row_sa + 128u, c * 8u + 64u < ncb);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t ncb = (pk[j] >> 28) << 3, offb = (pk[j] & 0x0FFFFFFFu) << 3;
                        uint16_t *dst = my_gbuf + (uint64_t)(b0 + 4 * (jj + j) + sub) * bstride + offb + c * 8u;
                        stg128_if(dst, v[2 * j], c * 8u < ncb);
                        stg128_if(dst + 64, v[2 * j + 1], c * 8u + 64u < ncb);
                    }
                }
                __syncwarp();
                if (ncp < n8) {  // this CTA's region of the bucket is full (heavily skewed input): exact escape, off the copy loop
                    for (uint32_t i0 = ncp; i0 < n8; i0 += 8) escape_chunk<S>(bl, *reinterpret_cast<const uint4 *>(&sm.stage[bl * kStageCap + i0]), table_k);
                }
                if (has_tail) *reinterpret_cast<uint4 *>(&sm.stage[bl * kStageCap]) = tail;
                sm.cursor[bl] = cnt - n8;
                sm.goff[bl] = off + ncp;
            }
        }
        __syncthreads();
    }

    // ---- end: the <= 7 items still staged per bucket, then this CTA's fill of every bucket ----
    for (int b = threadIdx.x; b < kNB; b += kP1Threads) {
        const uint32_t cnt = sm.cursor[b];
        uint32_t off = sm.goff[b];
        uint16_t *dst = my_gbuf + (uint64_t)b * gridDim.x * cap_cb;
        for (uint32_t i = 0; i < cnt; ++i) {
            const uint16_t item = sm.stage[stage_slot(b, i)];
            if (off < cap_front) dst[off++] = item;
            else red_kmers_of_word<S>(ItemFmt<S>::wmer((uint32_t)b, item), table_k, 1u);
        }
        gcount[2 * ((uint64_t)b * gridDim.x + blockIdx.x)] = off;                                          // front part: [0, off)
        const uint32_t back = min(sm.gback[b], cap_cb - cap_front);
        gcount[2 * ((uint64_t)b * gridDim.x + blockIdx.x) + 1] = back;                                     // back part: [cap_front, cap_front + back)
        if (off + back) atomicAdd(&bucket_total[b], off + back);  // pass 2 takes the buckets largest first
    }
    t_windows += (unsigned long long)n_fast * (16 * S);
    t_valid += (unsigned long long)n_fast * (16 * S);
    t_windows = warp_sum(t_windows);
    t_valid = warp_sum(t_valid);
    unsigned long long t_unk = warp_sum((unsigned long long)t_unknown);
    if (lane == 0) {
        if (t_windows) atomicAdd(&P->n_windows, t_windows);
        if (t_valid) atomicAdd(&P->valid_bases, t_valid);
        if (t_unk) atomicAdd(&P->unknown_chars, t_unk);
    }
}

// ------------------------------------------------------------------------------------------------
// pass 2: per bucket, 65536 16-bit counters in shared memory; one shared atomicAdd per item.
// The bucket's items lie in n_seg CTA-private segments; warps take whole segments.
// ------------------------------------------------------------------------------------------------
// ---- pass 2 epilogue with core buckets: fold the bucket's 65536 W-mer counters to k-mer counts in shared memory ----
// Counter index = payload = (a : A bits | r : R bits); the bucket b is the core.  The k-mer at offset t of the W-mer
// [a | core | r] drops the top 2t bits of a and the low A-2t bits of r:  x_t = [a_lo | core | r_hi], so
//   T_k[x_t] += sum over a_hi (4^t slabs) and r_lo (an aligned run of 4^(S-1-t) consecutive counters).
// Every counter is read S times (vector loads for the runs, dp2a to add the 16-bit halves); consecutive threads produce
// consecutive k-mers, so the S * 2^R global reds of a bucket are coalesced (32 lanes per 128-byte line).
template <int NU16>
__device__ __forceinline__ uint32_t sum_u16_run(const uint16_t *p)  // NU16 consecutive counters, aligned to their size
{
    if constexpr (NU16 == 1) {
        return *p;
    } else if constexpr (NU16 == 4) {
        const uint2 v = *reinterpret_cast<const uint2 *>(p);
        return __dp2a_lo(v.y, 0x0101u, __dp2a_lo(v.x, 0x0101u, 0u));
    } else {
        uint32_t acc = 0;
#pragma unroll
        for (int i = 0; i < NU16 / 8; ++i) {
            const uint4 v = reinterpret_cast<const uint4 *>(p)[i];
            acc = __dp2a_lo(v.x, 0x0101u, acc);
            acc = __dp2a_lo(v.y, 0x0101u, acc);
            acc = __dp2a_lo(v.z, 0x0101u, acc);
            acc = __dp2a_lo(v.w, 0x0101u, acc);
        }
        return acc;
    }
}
template <int S, int T>
__device__ __forceinline__ void fold_core_offset(const uint16_t *cnt, uint32_t b, uint32_t *table_k)
{
    constexpr int A = ItemFmt<S>::a_bits(), R = ItemFmt<S>::r_bits();
    constexpr int LO = A - 2 * T;   // bits of r that are summed over
    constexpr int HI = 2 * T;       // bits of a that are summed over
    constexpr int RH = R - LO;      // bits of r that stay
    for (uint32_t o = threadIdx.x; o < (1u << R); o += kP2Threads) {
        const uint32_t r_hi = o & ((1u << RH) - 1u), a_lo = o >> RH;
        uint32_t sum = 0;
#pragma unroll 4
        for (uint32_t a_hi = 0; a_hi < (1u << HI); ++a_hi) {
            const uint32_t a = (a_hi << LO) | a_lo;  // A bits: a_hi on top (HI bits), a_lo below (A - HI = LO bits)
            sum += sum_u16_run<(1 << LO)>(cnt + ((a << R) | (r_hi << LO)));
        }
        if (sum) red_add_u32(table_k + ((a_lo << (10 + RH)) | (b << RH) | r_hi), sum);
    }
}
template <int S, int T = 0>
__device__ __forceinline__ void fold_core_bucket(const uint16_t *cnt, uint32_t b, uint32_t *table_k)
{
    if constexpr (T < S) {
        fold_core_offset<S, T>(cnt, b, table_k);
        fold_core_bucket<S, T + 1>(cnt, b, table_k);
    }
}

template <int S>
__device__ __noinline__ void drain_counter(uint32_t *word, uint32_t hi_half, uint32_t wcode, uint32_t *table_k)
{
    atomicSub(word, hi_half ? 0x80000000u : 0x8000u);
    red_kmers_of_word<S>(wcode, table_k, 32768u);
}

template <int S>
__global__ void __launch_bounds__(kP2Threads, 1)
count_buckets_kernel(const uint16_t *__restrict__ gbuf, uint32_t cap_cb, uint32_t cap_front, const uint32_t *__restrict__ gcount, int n_seg,
                     uint16_t *__restrict__ table_w, uint32_t *__restrict__ table_k, uint32_t *__restrict__ work)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t *sub = reinterpret_cast<uint32_t *>(smem_raw);  // 32768 words = 65536 packed 16-bit counters
    __shared__ uint32_t s_bucket, s_next_seg;
    __shared__ uint32_t s_key[kNB];
    __shared__ uint32_t s_max;
    const int lane = threadIdx.x & 31;
    const uint32_t sub_sa = (uint32_t)__cvta_generic_to_shared(sub);
    // Largest buckets first: real genomes load the buckets unevenly (base composition, repeats), and a bucket several times the
    // average that is started last would be the tail of the kernel.  Sizes are compared in 16 coarse classes (relative to the
    // largest bucket) and ties keep the index order, so that on evenly loaded input the CTAs still sweep the bucket regions in
    // address order (a random order costs ~3 % in pass 2: TLB / DRAM page locality).  Every CTA sorts the 1024 keys for itself
    // (bitonic, 55 compare-exchange steps, of which 30 stay inside a warp).
    const uint32_t *bucket_total = work + 16;
    // When the largest bucket is within 25 % of the mean (uniform sequence) the order does not matter: no sort, index order.
    __shared__ uint32_t s_sum;
    if (threadIdx.x == 0) { s_max = 0; s_sum = 0; }
    __syncthreads();
    for (int i = threadIdx.x; i < kNB; i += kP2Threads) {
        const uint32_t t = bucket_total[i];
        atomicMax(&s_max, t);
        atomicAdd(&s_sum, t >> 10);  // the mean, to within kNB
    }
    __syncthreads();
    const bool sorted = s_max > s_sum + (s_sum >> 2) + 1024u;
    {
        const uint32_t mx = s_max;
        const int shift = (sorted && mx >= 16) ? (32 - __clz(mx)) - 4 : 31;  // the largest bucket falls in class 8..15 (not sorted: one class)
        for (int i = threadIdx.x; i < kNB; i += kP2Threads) s_key[i] = (((bucket_total[i] >> shift) >> (sorted ? 0 : 1)) << 10) | (uint32_t)(kNB - 1 - i);
    }
    __syncthreads();
    for (int k2 = 2; sorted && k2 <= kNB; k2 <<= 1)
        for (int j = k2 >> 1; j > 0; j >>= 1) {
            for (int i = threadIdx.x; i < kNB; i += kP2Threads) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    const uint32_t a = s_key[i], c = s_key[ixj];
                    if (((i & k2) == 0) ? (a < c) : (a > c)) { s_key[i] = c; s_key[ixj] = a; }  // descending overall
                }
            }
            // thread i owns element i (kP2Threads == kNB): exchanges at distance < 32 stay inside a warp's 32 elements
            if (kP2Threads == kNB && j < 32 && (j >> 1) >= 1) __syncwarp();
            else __syncthreads();
        }
    constexpr int kFillCache = 512;            // segments per bucket whose fills are staged in shared memory (148 on B200)
    __shared__ uint2 s_fill[kFillCache];
    for (;;) {
        if (threadIdx.x == 0) { s_bucket = atomicAdd(work, 1u); s_next_seg = 0; }
        __syncthreads();
        if (s_bucket >= (uint32_t)kNB) break;
        const uint32_t b = (uint32_t)(kNB - 1) - (s_key[s_bucket] & (uint32_t)(kNB - 1));
        // zero the counters and stage the fills of the bucket's segments (one coalesced read instead of a dependent global
        // load in front of every segment: with short segments -- small ranges, many GPUs -- that latency was most of a bucket)
        for (int i = threadIdx.x; i < 32768 / 4; i += kP2Threads) reinterpret_cast<uint4 *>(sub)[i] = make_uint4(0, 0, 0, 0);
        for (int i = threadIdx.x; i < n_seg && i < kFillCache; i += kP2Threads) s_fill[i] = reinterpret_cast<const uint2 *>(gcount)[(uint64_t)b * n_seg + i];
        __syncthreads();
        // Two 16-bit counters per word.  A counter is drained around 0x8000 (not at 0xFFFF) so that a carry can never
        // cross into its neighbour, whatever the interleaving of the other threads' updates (fewer than 32768 increments
        // can be in flight).  Eight items (one 128-bit load) at a time, branch-free: 8 shared atomics back to back, then 8 checks.
        // The drain is triggered by the increment that SEES 0x8000 (unique: the half then reads 0x8001 and the trigger takes
        // 32768 out again), so the hot test is one AND per item: bit 15 of the half in the old value.  Whoever sees a half
        // in 0x8001..0xFFFE while a drain is pending takes the slow path and finds nothing to do.
        auto add8 = [&](const uint4 &v) {
            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
            uint32_t old[8], top[8], any = 0;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
                const uint32_t w = w4[e >> 1];
                uint32_t sa, inc;
                if (e & 1) {  // item in the high half-word
                    sa = sub_sa + ((w >> 15) & 0x1FFFCu);
                    inc = max(w & 0x10000u, 1u);           // odd index: the word's high counter
                } else {
                    sa = sub_sa + ((w << 1) & 0x1FFFCu);
                    inc = (w & 1u) * 0xFFFFu + 1u;
                }
                top[e] = inc << 15;                         // bit 15 of the counter this item increments
                asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old[e]) : "r"(sa), "r"(inc) : "memory");
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) any |= old[e] & top[e];
            if (any) {
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const uint32_t idx = (e & 1) ? (w4[e >> 1] >> 16) : (w4[e >> 1] & 0xFFFFu);
                    const uint32_t half = (idx & 1u) ? (old[e] >> 16) : (old[e] & 0xFFFFu);
                    if (half == 0x8000u) drain_counter<S>(sub + (idx >> 1), idx & 1u, ItemFmt<S>::wmer(b, idx), table_k);
                }
            }
        };
        auto add_item = [&](uint32_t idx) {
            const uint32_t hi_half = idx & 1u;
            uint32_t *word = sub + (idx >> 1);
            const uint32_t old = atomicAdd(word, hi_half ? 0x10000u : 1u);
            const uint32_t half = hi_half ? (old >> 16) : (old & 0xFFFFu);
            if (half == 0x8000u) drain_counter<S>(word, hi_half, ItemFmt<S>::wmer(b, idx), table_k);
        };
        for (;;) {
            uint32_t seg = 0;
            if (lane == 0) seg = atomicAdd(&s_next_seg, 1u);
            seg = __shfl_sync(0xffffffffu, seg, 0);
            if (seg >= (uint32_t)n_seg) break;
            const uint2 fill = seg < (uint32_t)kFillCache ? s_fill[seg] : reinterpret_cast<const uint2 *>(gcount)[(uint64_t)b * n_seg + seg];
            for (int part = 0; part < 2; ++part) {  // the flushed chunks at the front of the region, the direct appends behind cap_front
                const uint32_t n = part ? min(fill.y, cap_cb - cap_front) : min(fill.x, cap_front);
                if (!n) continue;
                const uint16_t *items = gbuf + ((uint64_t)b * n_seg + seg) * cap_cb + (part ? cap_front : 0u);  // 16-byte aligned: both caps are multiples of 8
                const uint32_t n8 = n & ~7u;
                for (uint32_t i = lane * 8u; i < n8; i += 1024u) {  // four 128-bit loads in flight per lane
                    uint4 v[4];
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        v[u] = (i + 256u * u < n8) ? *reinterpret_cast<const uint4 *>(items + i + 256u * u) : make_uint4(0, 0, 0, 0);
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (i + 256u * u < n8) add8(v[u]);
                }
                if ((uint32_t)lane < n - n8) add_item(items[n8 + lane]);
            }
        }
        __syncthreads();
        if constexpr (ItemFmt<S>::kCore) {
            fold_core_bucket<S>(reinterpret_cast<const uint16_t *>(sub), b, table_k);
        } else {
            uint4 *out = reinterpret_cast<uint4 *>(table_w + ((uint64_t)b << 16));
            for (int i = threadIdx.x; i < 32768 / 4; i += kP2Threads) out[i] = reinterpret_cast<const uint4 *>(sub)[i];
        }
        __syncthreads();
    }
}

// ------------------------------------------------------------------------------------------------
// pass 3: fold T_W (16-bit W-mers at stride S) into T_k, one base per level.
//   suf_m[y] = occurrences of y as the LAST m bases of a sampled window        suf_W = T_W,  suf_{m-1}[y] = sum_c suf_m[c y]
//   all_m[y] = occurrences of y at ANY offset inside a sampled window          all_W = T_W,  all_{m-1}[y] = sum_b all_m[y b] + suf_{m-1}[y]
// (an (m-1)-mer inside a window is either the prefix of an m-mer in it, or the window's last m-1 bases.)
// all_k is the count of every k-mer covered by a sampled window: T_k += all_k.  Work: ~2 * 4^W reads, no atomics.
// ------------------------------------------------------------------------------------------------
template <typename TIn>
__global__ void __launch_bounds__(256) fold_level_kernel(const TIn *__restrict__ all_in, const TIn *__restrict__ suf_in, int m,  // inputs are 4^m
                                                         uint32_t *__restrict__ all_out, uint32_t *__restrict__ suf_out,
                                                         uint32_t *__restrict__ table_k /* last level: add instead of store */)
{
    const uint64_t n = 1ull << (2 * (m - 1));
    for (uint64_t y = (uint64_t)blockIdx.x * 256 + threadIdx.x; y < n; y += (uint64_t)gridDim.x * 256) {
        const uint32_t suf = (uint32_t)suf_in[y] + suf_in[n + y] + suf_in[2 * n + y] + suf_in[3 * n + y];
        uint32_t all;
        if constexpr (sizeof(TIn) == 2) {
            const uint2 v = *reinterpret_cast<const uint2 *>(all_in + 4 * y);  // four consecutive 16-bit entries
            all = (v.x & 0xFFFFu) + (v.x >> 16) + (v.y & 0xFFFFu) + (v.y >> 16);
        } else {
            const uint4 v = *reinterpret_cast<const uint4 *>(all_in + 4 * y);
            all = v.x + v.y + v.z + v.w;
        }
        all += suf;
        if (table_k) {
            if (all) red_add_u32(table_k + y, all);  // T_k is live: the edge slivers add to it concurrently on the side stream
        } else {
            all_out[y] = all;
            suf_out[y] = suf;
        }
    }
}

__global__ void __launch_bounds__(256) fold_copy_kernel(const uint16_t *__restrict__ table_w, uint32_t *__restrict__ table_k, uint64_t n)
{
    for (uint64_t x = (uint64_t)blockIdx.x * 256 + threadIdx.x; x < n; x += (uint64_t)gridDim.x * 256) {
        const uint32_t v = table_w[x];
        if (v) red_add_u32(table_k + x, v);  // T_k is live (edge slivers on the side stream)
    }
}

// scratch: for every level m-1 in [k .. W-1) an `all` and a `suf` array of 4^(m-1) uint32
inline uint32_t *fold_level_ptr(uint32_t *scratch, int level /* entries 4^level */, int which)
{
    uint64_t off = 0;
    for (int l = kW - 1; l > level; --l) off += 2ull << (2 * l);
    return scratch + off + (which ? (1ull << (2 * level)) : 0);
}

template <int S>
cudaError_t run_fold(const LaunchInfo &li, const BucketScratch &bs, uint32_t *d_table, cudaStream_t st, int *launches)
{
    constexpr int K = kW - S + 1;
    const uint64_t cap = (uint64_t)li.sm_count * 8;
    auto grid = [&](uint64_t n) { uint64_t b = (n + 255) / 256; return (unsigned)(b < cap ? (b ? b : 1) : cap); };
    if constexpr (S == 1) {
        fold_copy_kernel<<<grid(1ull << (2 * kW)), 256, 0, st>>>(bs.table_w, d_table, 1ull << (2 * kW));
        if (launches) ++*launches;
        return cudaGetLastError();
    } else {
    for (int m = kW; m > K; --m) {  // 4^m -> 4^(m-1)
        const bool last = (m - 1 == K);
        uint32_t *all_out = last ? nullptr : fold_level_ptr(bs.fold, m - 1, 0), *suf_out = last ? nullptr : fold_level_ptr(bs.fold, m - 1, 1);
        const uint64_t n = 1ull << (2 * (m - 1));
        if (m == kW)
            fold_level_kernel<uint16_t><<<grid(n), 256, 0, st>>>(bs.table_w, bs.table_w, m, all_out, suf_out, last ? d_table : nullptr);
        else
            fold_level_kernel<uint32_t><<<grid(n), 256, 0, st>>>(fold_level_ptr(bs.fold, m, 0), fold_level_ptr(bs.fold, m, 1), m, all_out, suf_out,
                                                                 last ? d_table : nullptr);
        if (launches) ++*launches;
    }
    return cudaGetLastError();
    }
}

template <int S>
cudaError_t run_bucketed(const LaunchInfo &li, const BucketScratch &bs, const uint8_t *d_stream, uint64_t lo, uint64_t hi,
                         uint32_t *d_table, uint8_t *d_flags, fkb_partials *d_partials, cudaStream_t st, int *launches)
{
    // the dynamic shared-memory opt-in is a PER-DEVICE function attribute: remembered per device (one bit each), so that
    // contexts on several GPUs of one process (findKmer -g N) all get it
    static std::atomic<uint64_t> attr_done{0};
    int dev = 0;
    cudaError_t ed = cudaGetDevice(&dev);
    if (ed != cudaSuccess) return ed;
    const uint64_t dev_bit = 1ull << (dev & 63);
    if (!(attr_done.load(std::memory_order_acquire) & dev_bit)) {
        cudaError_t e = cudaFuncSetAttribute(bucketize_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(P1Smem));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(count_buckets_kernel<S>, cudaFuncAttributeMaxDynamicSharedMemorySize, 128 * 1024);
        if (e != cudaSuccess) return e;
        attr_done.fetch_or(dev_bit, std::memory_order_release);
    }
    const uint64_t wspan = 32ull * 16 * S;
    const uint64_t n_witers = (hi - lo) / wspan;
    if (n_witers >> 32) return cudaErrorInvalidValue;  // the kernel counts a warp's iterations in 32 bits (a range of >= 2 TiB: not on this device)
    cudaError_t e = cudaMemsetAsync(bs.work, 0, 64 + kNB * sizeof(uint32_t), st);  // pass-2 work counter + per-bucket item totals
    if (e != cudaSuccess) return e;
    const bool timed = bs.phase_ev[0] != nullptr;  // option "phase_events" (bench.py: live per-kernel durations)
    if (timed) cudaEventRecord(bs.phase_ev[0], st);
    bucketize_kernel<S><<<bs.n_cta * FKB_P1_MINBLOCKS, P1Cfg<S>::kThreads, sizeof(P1Smem), st>>>(d_stream, lo, n_witers, bs.gbuf, bs.cap_cb, bs.cap_front, bs.gcount, bs.work + 16, d_table, d_flags, d_partials);
    if (timed) cudaEventRecord(bs.phase_ev[1], st);
    count_buckets_kernel<S><<<li.sm_count, kP2Threads, 128 * 1024, st>>>(bs.gbuf, bs.cap_cb, bs.cap_front, bs.gcount, bs.n_cta * FKB_P1_MINBLOCKS, bs.table_w, d_table, bs.work);
    if (launches) *launches += 2;
    cudaError_t e2 = cudaGetLastError();
    if (e2 != cudaSuccess) return e2;
    if (timed) cudaEventRecord(bs.phase_ev[2], st);
    cudaError_t e3 = cudaSuccess;
    if constexpr (!ItemFmt<S>::kCore) e3 = run_fold<S>(li, bs, d_table, st, launches);  // (core buckets: pass 2 has already folded into T_k)
    if (timed) cudaEventRecord(bs.phase_ev[3], st);
    return e3;
}

}  // namespace

int bucket_stride_for(int k) { return (k >= 6 && k <= kW) ? kW - k + 1 : 0; }

uint64_t bucket_unit_bytes(int k)
{
    int S = bucket_stride_for(k);
    return S ? 32ull * 16 * S : 0;  // one warp iteration
}

size_t bucket_table_w_bytes() { return ((size_t)1 << (2 * kW)) * sizeof(uint16_t); }
size_t bucket_fold_bytes()
{
    size_t words = 0;
    for (int l = kW - 1; l >= 6; --l) words += (size_t)2 << (2 * l);
    return words * sizeof(uint32_t);
}
int bucket_count() { return kNB; }
bool bucket_folds_in_shared(int k) { const int S = bucket_stride_for(k); return S >= 1 && S <= kCoreMaxS && S <= 5; }  // ItemFmt<S>::kCore
int bucket_segments_per_sm() { return FKB_P1_MINBLOCKS; }

cudaError_t launch_count_bucketed(const LaunchInfo &li, const BucketScratch &bs, const uint8_t *d_stream, uint64_t lo, uint64_t hi, int k,
                                  uint32_t *d_table, uint8_t *d_flags, fkb_partials *d_partials, cudaStream_t st, int *launches)
{
    switch (bucket_stride_for(k)) {
    case 1: return run_bucketed<1>(li, bs, d_stream, lo, hi, d_table, d_flags, d_partials, st, launches);
    case 2: return run_bucketed<2>(li, bs, d_stream, lo, hi, d_table, d_flags, d_partials, st, launches);
    case 3: return run_bucketed<3>(li, bs, d_stream, lo, hi, d_table, d_flags, d_partials, st, launches);
    case 4: return run_bucketed<4>(li, bs, d_stream, lo, hi, d_table, d_flags, d_partials, st, launches);
    case 5: return run_bucketed<5>(li, bs, d_stream, lo, hi, d_table, d_flags, d_partials, st, launches);
    case 6: return run_bucketed<6>(li, bs, d_stream, lo, hi, d_table, d_flags, d_partials, st, launches);
    case 7: return run_bucketed<7>(li, bs, d_stream, lo, hi, d_table, d_flags, d_partials, st, launches);
    case 8: return run_bucketed<8>(li, bs, d_stream, lo, hi, d_table, d_flags, d_partials, st, launches);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace fkb
