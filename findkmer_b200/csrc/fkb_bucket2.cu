// fkb_bucket2.cu -- VARIANT_BUCKET16: the k = 11 count path for large inputs ("double 13-mers").
//
// fkb_bucket.cu routes one 13-mer per 3 bases (16-bit payload, 1024 core buckets) and is bound by the shared-memory data
// pipe: three scattered shared-memory operations per routed item (slot atomic + payload store in pass 1, counter atomic in
// pass 2), 3.5 wavefronts each (profiles/r01_ncu_full_summary.txt, profiles/r02_ceiling_model.md).  This path halves the
// scattered operations of pass 1 by routing LONGER items that pass 2 counts TWICE:
//
//   item   = the 16-mer (32 bits = one register, one 4-byte slot) that ends at every 6th stream position (p == 3 mod 6 in
//            absolute coordinates).  A fully valid 16-mer contains the six 11-mers that end at p-5 .. p, i.e. it is exactly
//            two 13-mers at stride 3:  A = bases 0..12 (11-mers at offsets 0,1,2) and B = bases 3..15 (offsets 3,4,5).
//   bucket = bases 5..9 of the 16-mer -- five of the six bases that ALL six 11-mers contain -- so every 16-mer that
//            contributes to an 11-mer x at a given offset lies in one bucket, and pass 2 can fold in shared memory.
//   pass 1 : one slot atomic + one 4-byte store per SIX bases (fkb_bucket.cu: per three); same staging geometry (1024 rows of
//            208 bytes, 16-byte chunks appended to a region of HBM private to (bucket, CTA)); same bytes to HBM (4 B / 6 bases).
//   pass 2 : two shared-memory atomics per item into two arrays of 65536 EIGHT-bit counters (A and B, 64 KiB each) -- the same
//            number of counter atomics per base as before.  A counter that is seen at 0x80 is drained by exactly that
//            increment (unique trigger; the 128 counts go to a small list that is applied to T_k at the end of the bucket).
//            Eight bits leave room for only 126 increments in flight between trigger and drain, which a CTA of 1024 threads
//            can exceed on pathological input (a homopolymer run of a kilobase): the increment that would wrap a byte SEES
//            0xFF, raises a flag, and the whole bucket is then recounted exactly with global reds -- nothing was committed to
//            T_k before the flag is known.  Epilogue: both arrays are folded to 6 x 4096 11-mer sums and added to T_k.
//
// Exactly the semantics of the reference's scan (findKmer/src/findKmer.cpp:962-1069): 11-mers whose covering 16-mer is broken
// by a reset or a range edge are "leftovers" and go straight to T_k; the rare per-run events (:1044-1057, :1059-1062) are
// handled as in fkb_bucket.cu.
#include <atomic>
#include <type_traits>

#include "fkb_kernels.cuh"
#include "fkb_stream.cuh"

namespace fkb {

namespace {

constexpr int kK = 11;                 // k-mer length this path counts
constexpr int kL = 16;                 // item length (bases)
constexpr int kStep = 6;               // stride of the items = number of k-mers per item
constexpr int kJ = kStep - 1;
constexpr int kG = 3;                  // 16-byte groups per lane per iteration: 48 bytes = 8 items
constexpr int kCH = 16 * kG;
constexpr int kItems = kCH / kStep;    // 8
constexpr uint64_t kWSpan = 32ull * kCH;
constexpr int kNB = 1024;              // buckets
constexpr int kCap = 52;               // staged items per bucket: 13 chunks of 4 items = 208 bytes per row, as in fkb_bucket.cu
// pass-1 CTA size and tile length, measured on B200 (config 4).  Without the L2 prefetch below the kernel was latency-bound and wanted
// warps (profiles/r02_b16_variants.txt: 512 threads x 7 iterations 1.44 ms, 576 x 6 1.42, 640 x 6 1.39 -- the last size without
// spills --, 768 x 5 1.55, 1024 x 3 2.27).  With it the ALU pipe and the flush bound it and fewer, fatter warps are as good or better
// (profiles/r02_prefetch.txt: 384 x 9 1.33, 448 x 8 1.32, 512 x 6 / 7 / 8 1.32 / 1.30 / 1.30, 576 x 7 1.31, 640 x 6 1.32, 768 x 5 1.50)
#ifndef FKB2_P1_THREADS
#define FKB2_P1_THREADS 512
#endif
#ifndef FKB2_TILE_ITERS
#define FKB2_TILE_ITERS 7              // warp iterations between flushes: 7 * 8 * 512 / 1024 = 28 items per row on average (+ <= 3 carried) of 52
#endif
// L2 prefetch distance of pass 1 in iterations beyond the register pipeline.  Measured (profiles/r02_prefetch.txt, 3.1 Gbp): off 1.42 ms,
// 1 / 2 / 3 / 4 / 8 iterations 1.32 / 1.32 / 1.33 / 1.40 / 1.62 ms.  The load-to-use distance of the register pipeline is one iteration
// (~1.5 us), about the HBM latency under load: 26 % of the kernel's stall samples sat on the first use of the loaded bytes.
#ifndef FKB2_PREFETCH
#define FKB2_PREFETCH 2
#endif
#ifndef FKB2_P2_THREADS
#define FKB2_P2_THREADS 1024
#endif
constexpr int kP1Threads = FKB2_P1_THREADS, kP1Warps = kP1Threads / 32;
constexpr int kTileIters = FKB2_TILE_ITERS;
constexpr int kP2Threads = FKB2_P2_THREADS;
constexpr uint32_t kKmask = (1u << (2 * kK)) - 1u;
constexpr int kDrainCap = 2048;        // drained (array, counter) pairs a bucket may collect before it is recounted exactly

__device__ __forceinline__ uint32_t bucket_of(uint32_t x) { return (x << 10) >> 22; }  // bases 5..9 = bits 21..12

// layout of the routed items in HBM: one segment of cap_cb items per (bucket, pass-1 CTA), [bucket][cta][cap] -- a bucket's segments are
// neighbours (pass 2 reads one region) and a CTA's 1024 flush targets lie 12 MB apart.  The transposed layout [cta][bucket][cap] (a
// CTA's targets within 120 MB) was measured: no difference in either pass (1.40 / 0.57 ms both ways), so TLB reach is not what
// bounds the flush.
__device__ __forceinline__ uint64_t seg_base(uint32_t bucket, uint32_t cta, uint32_t n_cta, uint32_t cap_cb) { return ((uint64_t)bucket * n_cta + cta) * cap_cb; }
__device__ __forceinline__ uint64_t bucket_stride(uint32_t n_cta, uint32_t cap_cb) { return (uint64_t)n_cta * cap_cb; }

// the six 11-mers of one item: every exact escape goes through here
__device__ __noinline__ void red_kmers_of_item(uint32_t x, uint32_t *table_k, uint32_t amount)
{
#pragma unroll
    for (int j = 0; j < kStep; ++j) red_add_u32(table_k + ((x >> (2 * j)) & kKmask), amount);
}

struct P1Smem {
    uint32_t stage[kNB * kCap];
    uint32_t cursor[kNB];  // items staged in the row
    uint32_t junk[32];     // directly behind cursor[]: one junk cursor per lane (slot atomic of an item that is not emitted)
    uint32_t goff[kNB];    // items already appended to the FRONT part of this CTA's region of the bucket (multiple of 4 until the end)
    uint32_t gback[kNB];   // items written straight to the BACK part of the region because the staging row was full
    uint32_t ev[16];
};

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
__device__ __forceinline__ uint32_t atoms_add(uint32_t saddr, uint32_t v)
{
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(saddr), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ void sts32_if(uint32_t saddr, uint32_t value, uint32_t pred)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0;\n\t@p st.shared.u32 [%0], %1;\n\t}" ::"r"(saddr), "r"(value), "r"(pred) : "memory");
}

__device__ __forceinline__ void sts32_if_neg(uint32_t saddr, uint32_t value, uint32_t t)
{
    asm volatile("{\n\t.reg .pred p;\n\tsetp.lt.s32 p, %2, 0;\n\t@p st.shared.u32 [%0], %1;\n\t}" ::"r"(saddr), "r"(value), "r"(t) : "memory");
}

// the item that ends at byte i of group g: bits [2*(15-i)+31 : 2*(15-i)] of (grp[g-1].code : grp[g].code)
template <int N>
__device__ __forceinline__ uint32_t item_at(const Group *grp)
{
    constexpr int o = (kL - 1) % kStep + N * kStep, g = o / 16 + 1, i = o % 16;
    if constexpr (i == 15) return grp[g].code;
    else return __funnelshift_r(grp[g].code, grp[g - 1].code, 2 * (15 - i));
}
template <int N> struct ItemPos {
    static constexpr int o = (kL - 1) % kStep + N * kStep, g = o / 16 + 1, i = o % 16;
};

// staging-row overflow of one chunk's items (bit n of `ovf` = n-th static slot): exact escape, off the hot path
__device__ __noinline__ void escape_slots(uint32_t ovf, const uint32_t *codes /* G+1 words: group before + own groups */, uint32_t *table_k)
{
    for (int n = 0; n < kItems; ++n) {
        if (!(ovf & (1u << n))) continue;
        const int o = (kL - 1) % kStep + n * kStep, g = o / 16 + 1, i = o % 16;
        const uint32_t x = (i == 15) ? codes[g] : __funnelshift_r(codes[g], codes[g - 1], 2 * (15 - i));
        red_kmers_of_item(x, table_k, 1u);
    }
}
__device__ __noinline__ void escape_chunk(uint4 v, uint32_t *table_k)
{
    red_kmers_of_item(v.x, table_k, 1u);
    red_kmers_of_item(v.y, table_k, 1u);
    red_kmers_of_item(v.z, table_k, 1u);
    red_kmers_of_item(v.w, table_k, 1u);
}

// ------------------------------------------------------------------------------------------------
// pass 1: bucketize.  The interior starts at `lo` (a multiple of 1536 in absolute stream coordinates) and is n_witers
// warp-iterations of 32 * 48 bytes long; 16 readable bytes exist on both sides of it.
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kP1Threads, 1)
bucketize16_kernel(const uint8_t *__restrict__ s, uint64_t lo, uint64_t n_witers, uint32_t *__restrict__ gbuf, uint32_t cap_cb, uint32_t cap_front,
                   uint32_t *__restrict__ gcount, uint32_t *__restrict__ bucket_total, uint32_t *__restrict__ table_k, uint8_t *__restrict__ flags,
                   fkb_partials *__restrict__ P)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    P1Smem &sm = *reinterpret_cast<P1Smem *>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t cursor_sa = (uint32_t)__cvta_generic_to_shared(sm.cursor), stage_sa = (uint32_t)__cvta_generic_to_shared(sm.stage);

    for (int b = threadIdx.x; b < kNB; b += kP1Threads) { sm.cursor[b] = 0; sm.goff[b] = 0; sm.gback[b] = 0; }
    if (threadIdx.x < 16) sm.ev[threadIdx.x] = 0;
    __syncthreads();

    const uint64_t n_warps = (uint64_t)gridDim.x * kP1Warps, gw = (uint64_t)blockIdx.x * kP1Warps + warp;
    const uint64_t q = n_witers / n_warps, rem = n_witers % n_warps;
    const uint32_t my_iters = (uint32_t)(q + (gw < rem ? 1 : 0));
    const uint64_t my_first = gw * q + (gw < rem ? gw : rem);
    const uint32_t max_iters = (uint32_t)(q + (rem ? 1 : 0));  // CTA-uniform trip count (barriers)
    const uint64_t region = lo + my_first * kWSpan, hi = lo + n_witers * kWSpan;
    uint32_t *const my_gbuf = gbuf + seg_base(0, blockIdx.x, gridDim.x, cap_cb);  // + bucket * bucket_stride
    const bool edge_first = (lane == 0) && (region == lo), edge_last = (lane == 31) && (region + (uint64_t)my_iters * kWSpan == hi);

    uint32_t t_unknown = 0, t_dummy = 0, n_fast = 0;
    unsigned long long t_windows = 0, t_valid = 0;

    // software pipeline: `cur` = packed groups of iteration it, `nxt` = of it+1, `raw` = loads of it+2 in flight (see fkb_bucket.cu)
    Group cur[kG], nxt[kG], carry;
    uint4 raw[kG];
    const uint8_t *const lane_base = s + region + (uint64_t)lane * kCH;
    auto load_group = [&](uint32_t it, int g) -> uint4 {
        const bool p = (it < my_iters) || (g == 0 && it == my_iters && lane == 0);
        return ldg128_if(lane_base + (uint64_t)it * kWSpan + 16 * g, p);
    };
    {
        carry = pack_group(ldg128(s + region - 16), t_dummy);
#pragma unroll
        for (int g = 0; g < kG; ++g) cur[g] = pack_group(load_group(0, g), my_iters > 0 ? t_unknown : t_dummy);
#pragma unroll
        for (int g = 0; g < kG; ++g) nxt[g] = pack_group(load_group(1, g), my_iters > 1 ? t_unknown : t_dummy);
#pragma unroll
        for (int g = 0; g < kG; ++g) raw[g] = load_group(2, g);
    }

    for (uint32_t tile0 = 0, tile_len = 1u + blockIdx.x % (uint32_t)kTileIters; tile0 < max_iters; tile0 += tile_len, tile_len = kTileIters) {
        const uint32_t tile_end = min(tile0 + tile_len, my_iters);
        for (uint32_t it = tile0; it < tile_end; ++it) {
            Group grp[kG + 2];
#pragma unroll
            for (int g = 0; g < kG; ++g) grp[g + 1] = cur[g];
            const uint32_t up_c = __shfl_sync(0xffffffffu, cur[kG - 1].code, (lane + 31) & 31);
            const uint32_t dn_c = __shfl_sync(0xffffffffu, lane == 0 ? nxt[0].code : cur[0].code, (lane + 1) & 31);
            grp[0].code = lane == 0 ? carry.code : up_c;
            grp[kG + 1].code = dn_c;
            const bool is_lo = edge_first && it == 0, is_hi = edge_last && it + 1 == my_iters;

            uint32_t own = cur[0].valid;
#pragma unroll
            for (int g = 1; g < kG; ++g) own &= cur[g].valid;
            if (lane == 0) own &= carry.valid & nxt[0].valid;
            uint32_t emit[kG + 2];
            const bool all_emit = __all_sync(0xffffffffu, own == 0xFFFFu && !(is_lo || is_hi));
            if (all_emit) {
#pragma unroll
                for (int g = 0; g < kG + 2; ++g) emit[g] = 0;
                ++n_fast;
                carry.code = up_c;
                carry.valid = 0xFFFFu;
            } else {
                const uint32_t up_v = __shfl_sync(0xffffffffu, cur[kG - 1].valid, (lane + 31) & 31);
                const uint32_t dn_v = __shfl_sync(0xffffffffu, lane == 0 ? nxt[0].valid : cur[0].valid, (lane + 1) & 31);
                grp[0].valid = lane == 0 ? carry.valid : up_v;
                grp[kG + 1].valid = dn_v;
                carry.code = up_c;
                carry.valid = up_v;
#pragma unroll
                for (int g = 1; g <= kG + 1; ++g) {
                    const uint32_t m = (grp[g - 1].valid << 16) | grp[g].valid;
                    uint32_t phase = 0;  // anchored ends: (16*(g-1) + i) == L-1 (mod 6), i = byte in group, bit = 15 - i
#pragma unroll
                    for (int i = 0; i < 16; ++i)
                        if ((16 * (g - 1) + i) % kStep == (kL - 1) % kStep) phase |= 1u << (15 - i);
                    emit[g] = runs_of<kL>(m) & phase;
                }
                if (is_hi) emit[kG + 1] = 0;
                if (is_lo) emit[1] &= (0xFFFFu >> kJ);
#pragma unroll
                for (int g = 1; g <= kG; ++g) {
                    const uint32_t m = (grp[g - 1].valid << 16) | grp[g].valid;
                    const uint32_t rk = runs_of<kK>(m) & 0xFFFFu;
                    t_windows += __popc(rk);
                    t_valid += __popc(grp[g].valid);
                    const uint32_t E = (emit[g] << 16) | emit[g + 1];  // k-mers covered by an emitted item: it ends 0..5 bytes later
                    uint32_t cov = E;
#pragma unroll
                    for (int d = 1; d <= kJ; ++d) cov |= (E << d);
                    const uint32_t left = rk & ~(cov >> 16);
                    const uint32_t first_k = rk & ~(m >> kK) & 0xFFFFu;
                    const uint32_t shorts = grp[g].valid & ~rk;
                    const uint32_t who = __ballot_sync(0xffffffffu, (left | first_k | shorts) != 0);
                    if (who) warp_group_events(who, left, first_k, shorts, m, grp[g - 1].code, grp[g].code, kK, flags, sm.ev, table_k);
                }
            }

            // ---- the 8 items of this chunk: 8 shared atomics back to back (slot in the bucket's staging row), the encode of
            //      iteration it+2 between them and the 8 stores that wait for their results ----
            uint32_t ovf = 0;
            ValidAcc va;
            Group enc[kG];
            const bool ld_full = it + 3 < my_iters, ld_halo = (it + 3 == my_iters) && lane == 0;
            const uint8_t *const p3 = lane_base + (uint64_t)(it + 3) * kWSpan;
#if FKB2_PREFETCH > 0
            // the 12 lines of iteration it + 3 + D go to L2 now (no register, no scoreboard): the load issued in iteration it + D then
            // has one iteration (~1.5 us) to come back from L2 instead of from HBM
            if (lane < (int)(kWSpan / 128) && it + 3u + FKB2_PREFETCH < my_iters)
                asm volatile("prefetch.global.L2 [%0];" ::"l"(s + region + (uint64_t)(it + 3u + FKB2_PREFETCH) * kWSpan + (uint64_t)lane * 128u));
#endif
            auto stage_items = [&](auto all_t) {
                constexpr bool ALL = decltype(all_t)::value;
                uint32_t f[kItems], bk[kItems], pos[kItems];
                auto slot = [&](auto n_t) {
                    constexpr int n = decltype(n_t)::value;
                    f[n] = item_at<n>(grp);
                    bk[n] = bucket_of(f[n]);
                    const uint32_t ca = mad_u32(bk[n], 4u, cursor_sa);
#ifdef FKB2_EXP_NOSTAGE  // (profiling experiment: no slot atomics, no payload stores -- loads + encode + loop overhead alone)
                    if constexpr (ALL) pos[n] = (ca & 1u) ? atoms_inc(ca) : 0xFFFFFFFFu;
#else
                    if constexpr (ALL) pos[n] = atoms_inc(ca);
#endif
                    else {
                        const uint32_t e = emit[ItemPos<n>::g] & (1u << (15 - ItemPos<n>::i));
                        const uint32_t ps = atoms_inc(e ? ca : cursor_sa + 4u * (uint32_t)(kNB + lane));
                        pos[n] = e ? ps : 0xFFFFFFFFu;
                    }
                };
                slot(std::integral_constant<int, 0>{}); slot(std::integral_constant<int, 1>{}); slot(std::integral_constant<int, 2>{});
                slot(std::integral_constant<int, 3>{}); slot(std::integral_constant<int, 4>{}); slot(std::integral_constant<int, 5>{});
                slot(std::integral_constant<int, 6>{}); slot(std::integral_constant<int, 7>{});
#pragma unroll
                for (int g = 0; g < kG; ++g) {
                    enc[g].code = pack_codes_fast(raw[g], va);
                    enc[g].valid = 0xFFFFu;
                    // (no zero fill when off: what stays in the registers beyond the end of the region is never consumed -- 1.30 -> 1.29 ms)
                    ldg128_keep(raw[g], p3 + 16 * g, ld_full || (g == 0 && ld_halo));
                }
                uint32_t top = 0;
#pragma unroll
                for (int n = 0; n < kItems; ++n) {
                    const uint32_t ps = pos[n];
                    const uint32_t sa = mad_u32(ps, 4u, mad_u32(bk[n], 4u * kCap, stage_sa));
                    sts32_if(sa, f[n], ps < (uint32_t)kCap);
                    top = max(top, ps + 1u);
                }
                if (top > (uint32_t)kCap) {
                    // a staging row is full (a bucket more popular than the tile allows for): straight to the BACK part of this CTA's
                    // region of the bucket; only when that is full too is the item escaped exactly with global reds
#pragma unroll
                    for (int n = 0; n < kItems; ++n) {
                        if (pos[n] != 0xFFFFFFFFu && pos[n] >= (uint32_t)kCap) {
                            const uint32_t qb = atomicAdd(&sm.gback[bk[n]], 1u);
                            if (qb < cap_cb - cap_front) my_gbuf[(uint64_t)bk[n] * bucket_stride(gridDim.x, cap_cb) + cap_front + qb] = f[n];
                            else ovf |= 1u << n;
                        }
                    }
                }
            };
            if (all_emit) stage_items(std::true_type{});
            else stage_items(std::false_type{});
            if (ovf) {
                uint32_t codes[kG + 1];
#pragma unroll
                for (int g = 0; g <= kG; ++g) codes[g] = grp[g].code;
                escape_slots(ovf, codes, table_k);
            }
            if (va.bad()) {
#pragma unroll
                for (int g = 0; g < kG; ++g) enc[g] = pack_group(load_group(it + 2, g), (it + 2 < my_iters) ? t_unknown : t_dummy);
            }
#pragma unroll
            for (int g = 0; g < kG; ++g) { cur[g] = nxt[g]; nxt[g] = enc[g]; }
        }

        // ---- flush: append whole 16-byte chunks (4 items) of every staged row to this CTA's region of the bucket ----
        __syncthreads();
        if (threadIdx.x < 9) {
            const uint32_t v = sm.ev[threadIdx.x];
            if (v) {
                sm.ev[threadIdx.x] = 0;
                unsigned long long *dst = threadIdx.x < 4 ? &P->head_base[threadIdx.x] : (threadIdx.x < 8 ? &P->short_first[threadIdx.x - 4] : &P->runs_ge_k);
                atomicAdd(dst, (unsigned long long)v);
            }
        }
        {
            const uint32_t sub = lane >> 3, c = lane & 7;
            uint32_t n_seg;
            asm volatile("mov.u32 %0, %%nctaid.x;" : "=r"(n_seg));
            const uint64_t bstride = bucket_stride(n_seg, cap_cb);
            const uint32_t cap4 = cap_front & ~3u;
            for (uint32_t b0 = warp * 32; b0 < (uint32_t)kNB; b0 += kP1Warps * 32) {
                const uint32_t bl = b0 + lane;
                const uint32_t cnt = min(sm.cursor[bl], (uint32_t)kCap);
                const uint32_t n4 = cnt & ~3u, off = sm.goff[bl];
                const uint32_t ncp = min(n4, cap4 - off);
                const bool has_tail = n4 && cnt > n4;
                uint4 tail = make_uint4(0, 0, 0, 0);
                if (has_tail) tail = *reinterpret_cast<const uint4 *>(&sm.stage[bl * kCap + n4]);
                const uint32_t packed = (off >> 2) | ((ncp >> 2) << 28);
#pragma unroll
                for (int jj = 0; jj < 8; jj += 4) {
                    uint32_t pk[4];
                    uint4 v[8];
#pragma unroll
                    for (int j = 0; j < 4; ++j) pk[j] = __shfl_sync(0xffffffffu, packed, 4 * (jj + j) + sub);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t ncb = (pk[j] >> 28) << 2;
                        const uint32_t row_sa = stage_sa + ((b0 + 4 * (jj + j) + sub) * kCap + c * 4u) * 4u;
                        v[2 * j] = lds128_if(row_sa, c * 4u < ncb);
                        v[2 * j + 1] = lds128_if(row_sa + 128u, c * 4u + 32u < ncb);
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const uint32_t ncb = (pk[j] >> 28) << 2, offb = (pk[j] & 0x0FFFFFFFu) << 2;
                        uint32_t *dst = my_gbuf + (uint64_t)(b0 + 4 * (jj + j) + sub) * bstride + offb + c * 4u;
#ifndef FKB2_EXP_NOFLUSH  // (profiling experiment, profiles/r02_b16_experiments.txt: without the flush stores the results are wrong and the time is the compute phase alone)
                        stg128_if(dst, v[2 * j], c * 4u < ncb);
                        stg128_if(dst + 32, v[2 * j + 1], c * 4u + 32u < ncb);
#else
                        if (v[2 * j].x == 0x12345u && v[2 * j + 1].y == 0x54321u) stg128_if(dst, v[2 * j], c * 4u < ncb);
#endif
                    }
                }
                __syncwarp();
                if (ncp < n4) {  // this CTA's region of the bucket is full (heavily skewed input): exact escape
                    for (uint32_t i0 = ncp; i0 < n4; i0 += 4) escape_chunk(*reinterpret_cast<const uint4 *>(&sm.stage[bl * kCap + i0]), table_k);
                }
                if (has_tail) *reinterpret_cast<uint4 *>(&sm.stage[bl * kCap]) = tail;
                sm.cursor[bl] = cnt - n4;
                sm.goff[bl] = off + ncp;
            }
        }
        __syncthreads();
    }

    // ---- end: the <= 3 items still staged per bucket, then this CTA's fill of every bucket ----
    for (int b = threadIdx.x; b < kNB; b += kP1Threads) {
        const uint32_t cnt = min(sm.cursor[b], (uint32_t)kCap);
        uint32_t off = sm.goff[b];
        uint32_t *dst = my_gbuf + (uint64_t)b * bucket_stride(gridDim.x, cap_cb);
        for (uint32_t i = 0; i < cnt; ++i) {
            const uint32_t item = sm.stage[b * kCap + i];
            if (off < cap_front) dst[off++] = item;
            else red_kmers_of_item(item, table_k, 1u);
        }
        gcount[2 * ((uint64_t)b * gridDim.x + blockIdx.x)] = off;
        const uint32_t back = min(sm.gback[b], cap_cb - cap_front);
        gcount[2 * ((uint64_t)b * gridDim.x + blockIdx.x) + 1] = back;
        if (off + back) atomicAdd(&bucket_total[b], off + back);
    }
    t_windows += (unsigned long long)n_fast * kCH;
    t_valid += (unsigned long long)n_fast * kCH;
    t_windows = warp_sum(t_windows);
    t_valid = warp_sum(t_valid);
    unsigned long long t_unk = warp_sum((unsigned long long)t_unknown);
    if (lane == 0) {
        if (t_windows) atomicAdd(&P->n_windows, t_windows);
        if (t_valid) atomicAdd(&P->valid_bases, t_valid);
        if (t_unk) atomicAdd(&P->unknown_chars, t_unk);
    }
}

#include "fkb_bucket2_ring.cuh"  // bucketize16_ring_kernel: pass 1 with ring rows and an asynchronous, warp-distributed flush (option "p1_ring")

// ------------------------------------------------------------------------------------------------
// pass 2: per bucket, two arrays of 65536 8-bit counters in shared memory; two shared atomics per item.
//   z = [bases 0..4 | bases 10..15] of the item (22 bits: the core taken out);  A index = z >> 6 (bases 0..4, 10..12),
//   B index = z & 0xFFFF (bases 3,4, 10..15).  Counter byte n of an array lives in word n >> 2, byte lane n & 3.
// ------------------------------------------------------------------------------------------------
struct P2Smem {
    uint32_t cnt[2][16384];   // A, B
    uint32_t drain[kDrainCap];  // (array << 16) | index of every drained counter (128 counts each)
    uint2 fill[512];
    uint32_t key[kNB];
    uint32_t bucket, bucket_next, next_seg, max_total, sum_total, n_drain, overflow;
};

// fold one array: out index o (12 bits) at offset t (0..2) of the array's 13-mer; R = number of low bits of o that lie below the core
//   t = 0: 16 consecutive counters at o << 4;  t = 1: 4 slabs (stride 2^14) of 4 consecutive counters at o << 2;  t = 2: 16 slabs (stride 2^12)
template <int T, int R>
__device__ __forceinline__ void fold_array(const uint8_t *cnt, uint32_t core, uint32_t *table_k)
{
    if constexpr (T == 0) {
        for (uint32_t o = threadIdx.x; o < 4096u; o += kP2Threads) {
            const uint4 v = *reinterpret_cast<const uint4 *>(cnt + (o << 4));
            const uint32_t sum = __dp4a(v.x, 0x01010101u, __dp4a(v.y, 0x01010101u, __dp4a(v.z, 0x01010101u, __dp4a(v.w, 0x01010101u, 0u))));
            if (sum) red_add_u32(table_k + (((o >> R) << (10 + R)) | (core << R) | (o & ((1u << R) - 1u))), sum);
        }
    } else if constexpr (T == 1) {
        for (uint32_t o = threadIdx.x; o < 4096u; o += kP2Threads) {
            uint32_t sum = 0;
#pragma unroll
            for (uint32_t ht = 0; ht < 4; ++ht) sum = __dp4a(*reinterpret_cast<const uint32_t *>(cnt + (ht << 14) + (o << 2)), 0x01010101u, sum);
            if (sum) red_add_u32(table_k + (((o >> R) << (10 + R)) | (core << R) | (o & ((1u << R) - 1u))), sum);
        }
    } else {
        for (uint32_t o4 = threadIdx.x; o4 < 1024u; o4 += kP2Threads) {  // four consecutive outputs per thread
            uint32_t even = 0, odd = 0;  // 16-bit lanes: bytes 0,2 and bytes 1,3 (sums of 16 bytes stay below 4096)
#pragma unroll
            for (uint32_t ht = 0; ht < 16; ++ht) {
                const uint32_t v = *reinterpret_cast<const uint32_t *>(cnt + (ht << 12) + (o4 << 2));
                even += v & 0x00FF00FFu;
                odd += (v >> 8) & 0x00FF00FFu;
            }
            const uint32_t sums[4] = {even & 0xFFFFu, odd & 0xFFFFu, even >> 16, odd >> 16};
#pragma unroll
            for (uint32_t e = 0; e < 4; ++e) {
                const uint32_t o = (o4 << 2) | e;
                if (sums[e]) red_add_u32(table_k + (((o >> R) << (10 + R)) | (core << R) | (o & ((1u << R) - 1u))), sums[e]);
            }
        }
    }
}

// the three 11-mers of a drained counter (array 0 = A: offsets 0..2 of the item, array 1 = B: offsets 3..5), `amount` counts each
__device__ __forceinline__ void red_kmers_of_counter(uint32_t array, uint32_t idx, uint32_t core, uint32_t *table_k, uint32_t amount)
{
    // rebuild the 13-mer [hi | core | lo] from the index: A = [bases 0..4 (10 bits) | bases 10..12 (6 bits)], B = [bases 3,4 (4 bits) | bases 10..15 (12 bits)]
    const uint32_t L = array ? 12u : 6u;
    const uint32_t hi = idx >> L, lo = idx & ((1u << L) - 1u);
    const uint32_t w13 = (hi << (10 + L)) | (core << L) | lo;  // A: item bases 0..12 = [hi10 | core | lo6];  B: item bases 3..15 = [hi4 | core | lo12]
#pragma unroll
    for (int t = 0; t < 3; ++t) red_add_u32(table_k + ((w13 >> (2 * (2 - t))) & kKmask), amount);
}

__global__ void __launch_bounds__(kP2Threads, 1)
count_buckets16_kernel(const uint32_t *__restrict__ gbuf, uint32_t cap_cb, uint32_t cap_front, const uint32_t *__restrict__ gcount, int n_seg,
                       uint32_t *__restrict__ table_k, uint32_t *__restrict__ work)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    P2Smem &sm = *reinterpret_cast<P2Smem *>(smem_raw);
    const int lane = threadIdx.x & 31;
    const uint32_t cntA_sa = (uint32_t)__cvta_generic_to_shared(sm.cnt[0]), cntB_sa = cntA_sa + 65536u;
    // largest buckets first (see fkb_bucket.cu) -- unless the buckets are about equally full (uniform sequence: the largest within
    // 25 % of the mean), where the order does not matter and the 55 barrier stages of the sort are 16 us of every CTA's time
    const uint32_t *bucket_total = work + 16;
    if (threadIdx.x == 0) { sm.max_total = 0; sm.sum_total = 0; }
    __syncthreads();
    for (int i = threadIdx.x; i < kNB; i += kP2Threads) {
        const uint32_t t = bucket_total[i];
        atomicMax(&sm.max_total, t);
        atomicAdd(&sm.sum_total, t >> 10);  // mean, to within kNB (no overflow)
    }
    __syncthreads();
    const bool sorted = sm.max_total > sm.sum_total + (sm.sum_total >> 2) + 1024u;
    if (sorted) {
        const uint32_t mx = sm.max_total;
        const int shift = mx >= 16 ? (32 - __clz(mx)) - 4 : 0;
        for (int i = threadIdx.x; i < kNB; i += kP2Threads) sm.key[i] = ((bucket_total[i] >> shift) << 10) | (uint32_t)(kNB - 1 - i);
        __syncthreads();
        for (int k2 = 2; k2 <= kNB; k2 <<= 1)
            for (int j = k2 >> 1; j > 0; j >>= 1) {
                for (int i = threadIdx.x; i < kNB; i += kP2Threads) {
                    const int ixj = i ^ j;
                    if (ixj > i) {
                        const uint32_t a = sm.key[i], c = sm.key[ixj];
                        if (((i & k2) == 0) ? (a < c) : (a > c)) { sm.key[i] = c; sm.key[ixj] = a; }
                    }
                }
                __syncthreads();
            }
    } else {
        for (int i = threadIdx.x; i < kNB; i += kP2Threads) sm.key[i] = (uint32_t)(kNB - 1 - i);  // bucket i at position i
    }
    // the next bucket's index is fetched by thread 0 while everybody folds the current one (a global atomic round trip and one CTA
    // barrier per bucket less: the per-bucket fixed costs are what a short shard pays for, profiles/r02_pass2_variants.txt)
    if (threadIdx.x == 0) { sm.bucket = atomicAdd(work, 1u); sm.next_seg = 0; sm.n_drain = 0; sm.overflow = 0; }
    __syncthreads();
    for (;;) {
        const uint32_t cur_bucket = sm.bucket;
        if (cur_bucket >= (uint32_t)kNB) break;
        const uint32_t b = (uint32_t)(kNB - 1) - (sm.key[cur_bucket] & (uint32_t)(kNB - 1));
        for (int i = threadIdx.x; i < 32768 / 4; i += kP2Threads) reinterpret_cast<uint4 *>(sm.cnt)[i] = make_uint4(0, 0, 0, 0);
        for (int i = threadIdx.x; i < n_seg && i < 512; i += kP2Threads) sm.fill[i] = reinterpret_cast<const uint2 *>(gcount)[(uint64_t)b * n_seg + i];
        __syncthreads();

        // one item: z, the two word addresses and byte-lane increments, two atomics
        auto add_slow = [&](uint32_t array, uint32_t idx, uint32_t old, uint32_t shift, uint32_t word_sa) {
            const uint32_t byte = (old >> shift) & 0xFFu;
            if (byte == 0x80u) {  // unique trigger: take 128 out, remember the counter
                asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(word_sa), "r"(0u - (0x80u << shift)) : "memory");
                const uint32_t n = atomicAdd(&sm.n_drain, 1u);
                if (n < (uint32_t)kDrainCap) sm.drain[n] = (array << 16) | idx;
                else sm.overflow = 1;
            } else if (byte == 0xFFu) {
                sm.overflow = 1;  // this increment wrapped the byte: the bucket is recounted exactly
            }
        };
        auto add4 = [&](const uint4 &v) {
            const uint32_t x4[4] = {v.x, v.y, v.z, v.w};
            uint32_t old[8], sh[8], wa[8], any = 0;
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const uint32_t z = mad_u32(x4[e] >> 22, 4096u, x4[e] & 0xFFFu);  // bits 21..0 = [bases 0..4 | bases 10..15]: the core taken out
                const uint32_t ia = z >> 6;                          // A index in bits 15..0 (garbage above)
                sh[2 * e] = (ia << 3) & 0x18u;
                wa[2 * e] = cntA_sa + (ia & 0xFFFCu);
                sh[2 * e + 1] = (z << 3) & 0x18u;
                wa[2 * e + 1] = cntB_sa + (z & 0xFFFCu);
            }
#pragma unroll
            for (int e = 0; e < 8; ++e) old[e] = atoms_add(wa[e], 1u << sh[e]);
#pragma unroll
            for (int e = 0; e < 8; ++e) any |= old[e];
            if (any & 0x80808080u) {  // some byte of a touched word is at or above 0x80: look closer (rare)
#pragma unroll
                for (int e = 0; e < 8; ++e) {
                    const uint32_t z = mad_u32(x4[e >> 1] >> 22, 4096u, x4[e >> 1] & 0xFFFu);
                    const uint32_t idx = (e & 1) ? (z & 0xFFFFu) : ((z >> 6) & 0xFFFFu);
                    add_slow(e & 1, idx, old[e], sh[e], wa[e]);
                }
            }
        };
        auto add_item = [&](uint32_t x) {
            const uint32_t z = mad_u32(x >> 22, 4096u, x & 0xFFFu);
            const uint32_t ia = (z >> 6) & 0xFFFFu, ib = z & 0xFFFFu;
            const uint32_t sa = (ia << 3) & 0x18u, sb = (ib << 3) & 0x18u;
            const uint32_t oa = atoms_add(cntA_sa + (ia & 0xFFFCu), 1u << sa);
            add_slow(0, ia, oa, sa, cntA_sa + (ia & 0xFFFCu));
            const uint32_t ob = atoms_add(cntB_sa + (ib & 0xFFFCu), 1u << sb);
            add_slow(1, ib, ob, sb, cntB_sa + (ib & 0xFFFCu));
        };
        for (;;) {
            uint32_t seg = 0;
            if (lane == 0) seg = atomicAdd(&sm.next_seg, 1u);
            seg = __shfl_sync(0xffffffffu, seg, 0);
            if (seg >= (uint32_t)n_seg) break;
            const uint2 fill = seg < 512u ? sm.fill[seg] : reinterpret_cast<const uint2 *>(gcount)[(uint64_t)b * n_seg + seg];
            for (int part = 0; part < 2; ++part) {
                const uint32_t n = part ? min(fill.y, cap_cb - cap_front) : min(fill.x, cap_front);
                if (!n) continue;
                const uint32_t *items = gbuf + seg_base(b, seg, (uint32_t)n_seg, cap_cb) + (part ? cap_front : 0u);
                // (tried: reloading a register set as soon as it has been counted, and taking the tail from an over-read last chunk instead
                // of a dependent load -- 0.57 -> 0.73 ms at 3.1 Gbp, no gain on 1/8 shards: profiles/r02_pass2_variants.txt)
                const uint32_t n4 = n & ~3u;
                for (uint32_t i = lane * 4u; i < n4; i += 512u) {  // four 128-bit loads in flight per lane
                    uint4 v[4];
                    // (an L2 prefetch of the warp's next 2 KiB steps was measured here too: 0.570 -> 0.577..0.591 ms, not kept)
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        v[u] = (i + 128u * u < n4) ? *reinterpret_cast<const uint4 *>(items + i + 128u * u) : make_uint4(0, 0, 0, 0);
#pragma unroll
                    for (int u = 0; u < 4; ++u)
                        if (i + 128u * u < n4) add4(v[u]);
                }
                if ((uint32_t)lane < n - n4) add_item(items[n4 + lane]);
            }
        }
        __syncthreads();
        const uint32_t was_overflow = sm.overflow, n_drained = sm.n_drain;
        if (threadIdx.x == 0) sm.bucket_next = atomicAdd(work, 1u);
        if (was_overflow) {
            // A byte wrapped (or the drain list ran over): the counters are meaningless.  Nothing of this bucket has reached T_k yet:
            // recount it exactly, six global reds per item.
            for (uint32_t seg = threadIdx.x >> 5; seg < (uint32_t)n_seg; seg += kP2Threads / 32) {
                const uint2 fill = reinterpret_cast<const uint2 *>(gcount)[(uint64_t)b * n_seg + seg];
                for (int part = 0; part < 2; ++part) {
                    const uint32_t n = part ? min(fill.y, cap_cb - cap_front) : min(fill.x, cap_front);
                    const uint32_t *items = gbuf + seg_base(b, seg, (uint32_t)n_seg, cap_cb) + (part ? cap_front : 0u);
                    for (uint32_t i = lane; i < n; i += 32) red_kmers_of_item(items[i], table_k, 1u);
                }
            }
        } else {
            const uint8_t *ca = reinterpret_cast<const uint8_t *>(sm.cnt[0]), *cb = reinterpret_cast<const uint8_t *>(sm.cnt[1]);
            fold_array<0, 2>(ca, b, table_k);
            fold_array<1, 4>(ca, b, table_k);
            fold_array<2, 6>(ca, b, table_k);
            fold_array<0, 8>(cb, b, table_k);
            fold_array<1, 10>(cb, b, table_k);
            fold_array<2, 12>(cb, b, table_k);
            const uint32_t nd = min(n_drained, (uint32_t)kDrainCap);
            for (uint32_t i = threadIdx.x; i < nd; i += kP2Threads) red_kmers_of_counter(sm.drain[i] >> 16, sm.drain[i] & 0xFFFFu, b, table_k, 128u);
        }
        if (threadIdx.x == 0) { sm.bucket = sm.bucket_next; sm.next_seg = 0; sm.n_drain = 0; sm.overflow = 0; }
        __syncthreads();
    }
}

}  // namespace

uint64_t bucket16_unit_bytes(int k) { return k == kK ? kWSpan : 0; }

cudaError_t launch_count_bucketed16(const LaunchInfo &li, const BucketScratch &bs, const uint8_t *d_stream, uint64_t lo, uint64_t hi, uint32_t *d_table,
                                    uint8_t *d_flags, fkb_partials *d_partials, cudaStream_t st, int *launches)
{
    static std::atomic<uint64_t> attr_done{0};  // the dynamic shared-memory opt-in is a per-device function attribute
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const uint64_t dev_bit = 1ull << (dev & 63);
    if (!(attr_done.load(std::memory_order_acquire) & dev_bit)) {
        e = cudaFuncSetAttribute(bucketize16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(P1Smem));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(count_buckets16_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(P2Smem));
        if (e != cudaSuccess) return e;
        e = cudaFuncSetAttribute(bucketize16_ring_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(P1Ring));
        if (e != cudaSuccess) return e;
        attr_done.fetch_or(dev_bit, std::memory_order_release);
    }
    const uint64_t n_witers = (hi - lo) / kWSpan;
    if (n_witers >> 32) return cudaErrorInvalidValue;
    // the scratch is sized in 16-bit items (fkb_bucket.cu); this path stores 32-bit items in the same bytes
    const uint32_t cap_cb = (bs.cap_cb / 2) & ~3u, cap_front = (bs.cap_front / 2) & ~3u;
    e = cudaMemsetAsync(bs.work, 0, 64 + kNB * sizeof(uint32_t), st);
    if (e != cudaSuccess) return e;
    const bool timed = bs.phase_ev[0] != nullptr;
    if (timed) cudaEventRecord(bs.phase_ev[0], st);
    // option "p1_ring": the asynchronous-flush pass 1 (bit-exact, measured 10 % slower than the synchronous flush: profiles/r02_ring_flush.md).
    // Its flush packs a segment's chunk offset into 20 bits and addresses gbuf by a 32-bit chunk index.
    if (bs.p1_ring && cap_front < (1u << 22) && (((uint64_t)kNB * bs.n_cta * cap_cb) >> 2) < (1ull << 32))
        bucketize16_ring_kernel<<<bs.n_cta, kRingThreads, sizeof(P1Ring), st>>>(d_stream, lo, n_witers, reinterpret_cast<uint32_t *>(bs.gbuf), cap_cb, cap_front,
                                                                               bs.gcount, bs.work + 16, d_table, d_flags, d_partials);
    else
        bucketize16_kernel<<<bs.n_cta, kP1Threads, sizeof(P1Smem), st>>>(d_stream, lo, n_witers, reinterpret_cast<uint32_t *>(bs.gbuf), cap_cb, cap_front, bs.gcount,
                                                                        bs.work + 16, d_table, d_flags, d_partials);
    if (timed) cudaEventRecord(bs.phase_ev[1], st);
    count_buckets16_kernel<<<li.sm_count, kP2Threads, sizeof(P2Smem), st>>>(reinterpret_cast<const uint32_t *>(bs.gbuf), cap_cb, cap_front, bs.gcount, bs.n_cta,
                                                                           d_table, bs.work);
    if (launches) *launches += 2;
    e = cudaGetLastError();
    if (timed) { cudaEventRecord(bs.phase_ev[2], st); cudaEventRecord(bs.phase_ev[3], st); }
    return e;
}

}  // namespace fkb
