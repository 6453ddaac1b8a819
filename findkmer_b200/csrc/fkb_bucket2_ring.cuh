// fkb_bucket2_ring.cuh -- the asynchronous-flush form of pass 1 of the k = 11 path (option "p1_ring"; NOT the default).
// Included by fkb_bucket2.cu inside its anonymous namespace, after the item / staging helpers it shares with bucketize16_kernel.
// The experiment record is profiles/r02_ring_flush.md.
#pragma once

// ------------------------------------------------------------------------------------------------
// pass 1, asynchronous flush ("ring") -- option "p1_ring"; NOT the default: bit-exact, but measured 1.52 ms against 1.38 ms
// for the kernel above (profiles/r02_ring_flush.md).  bucketize16_kernel stops the whole CTA every 6 iterations
// to copy the staging rows out: 0.40 ms of its 1.39 ms during which the SM computes nothing, plus the tile bookkeeping and
// two CTA barriers per tile (profiles/r02_b16_experiments.txt).  Here the rows are RINGS, there is no CTA barrier in the
// main loop, and every warp copies the completed 16-byte chunks of the 52 rows it OWNS to HBM every few iterations, at its
// own pace, while the other warps keep staging into those rows:
//
//   cursor word of a row = (claimed << 16) | limit.  A producer claims with ONE atomic add of 0x10000 (as before); the claim
//   is good iff claimed < limit (limit = flushed + 52), its slot is claimed mod 52.  A claim on a full row goes to the back
//   region / exact escape exactly as before and never touches the ring -- and because only the row's owner moves the limit,
//   bad claims are always a SUFFIX [limit, claimed) of the row's sequence: when the owner raises the limit (one CAS per row
//   and round) it also takes `claimed` back to the old limit, so no slot is ever copied out that was not written.
//   Which claimed slots are already WRITTEN: a warp publishes `progress = it + 1` after the ring stores of iteration it.
//   The owner snapshots its rows' cursors and then every warp's progress; one iteration later it checks that every warp
//   has published more than the snapshot saw (else it looks again an iteration later; a warp busy with bad claims has
//   published before it went there) and copies [flushed, min(claimed, limit)) of the snapshot, whole chunks only, then
//   raises the limits.  Both 16-bit fields are kept small by the owner (it subtracts 416 = 8 * 52 from both when
//   `flushed` passes 416), which also keeps `claimed mod 52` a single multiply-high.
//   (A dedicated flusher warp issuing cp.async.bulk was built first: cp.async.bulk is a uniform-datapath instruction, 32
//   lanes with 32 different rows go through a serialising R2UR loop at ~250 cycles per copy, and even with LSU copies one
//   warp needs longer for 1024 rows than the rings can absorb -- profiles/r02_ring_flush.txt.)
// ------------------------------------------------------------------------------------------------
#ifndef FKB2_PROD_WARPS
#define FKB2_PROD_WARPS 20             // 5 per scheduler: 96 registers (a 21st warp would cap the kernel at 80)
#endif
#ifndef FKB2_FLUSH_EVERY
#define FKB2_FLUSH_EVERY 6             // a warp's iterations between two flushes of its rows: ~5 items per row and iteration, 52 slots
#endif
#ifndef FKB2_FLUSH_WARPS
#define FKB2_FLUSH_WARPS 0             // > 0: that many DEDICATED flusher warps beside the producers, which then never flush (an experiment)
#endif
constexpr int kProdWarps = FKB2_PROD_WARPS, kFlushEvery = FKB2_FLUSH_EVERY, kFlushWarps = FKB2_FLUSH_WARPS;
constexpr int kRingThreads = 32 * (kProdWarps + kFlushWarps);
constexpr int kOwn = kFlushWarps ? 52 : (((kNB + kProdWarps - 1) / kProdWarps + 3) & ~3);  // rows per flush duty (52); the last group is what is left
constexpr int kOwners = (kNB + kOwn - 1) / kOwn;
constexpr uint32_t kRebase = kCap;  // `flushed` stays below 52, good claims below 104: slot = claimed - (claimed >= 52 ? 52 : 0)
constexpr uint32_t kDone = 0xFFFFFFFEu;  // even: nothing in flight
constexpr uint32_t kNotEmitted = 0x7FFF0000u;  // a "claim" that is neither good nor bad: claimed - limit = 0x7FFF
static_assert(kProdWarps <= 32 && kOwn <= 64 && kFlushEvery >= 1, "flush geometry");

struct P1Ring {
    uint32_t stage[kNB * kCap];
    uint32_t cursor[kNB];    // (claimed << 16) | limit
    uint32_t junk[32];       // directly behind cursor[]: one junk cursor per lane
    uint32_t goff[kNB];      // owner-private: items already appended to the FRONT part of this CTA's region of the bucket
    uint32_t gback[kNB];     // items written straight to the BACK part of the region because the ring was full
    uint32_t progress[32];   // per warp: 2 * it + 1 while the claims of iteration it are in flight, 2 * it + 2 once they are stored; kDone at the end
    uint32_t ev[16];
};
static_assert(sizeof(P1Ring) <= 232448, "shared memory");

__device__ __forceinline__ uint32_t lds_volatile(uint32_t saddr)
{
    uint32_t v;
    asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr) : "memory");
    return v;
}
__device__ __forceinline__ void sts_publish(uint32_t saddr, uint32_t v)
{
#ifdef FKB2_RING_RELEASE  // (experiment: a CTA-scope release -- MEMBAR.ALL.CTA -- in front of the store)
    asm volatile("st.release.cta.shared.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
#else
    // the shared-memory operations of one warp are performed in program order, and this store is issued after the warp's
    // ring stores of the iteration (a __syncwarp in between): a reader that sees it sees them
    asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(saddr), "r"(v) : "memory");
#endif
}
__device__ __forceinline__ uint32_t mod_cap(uint32_t n) { return n - kCap * ((n * 1261u) >> 16); }  // n < 1700
// rare: this CTA's region of the bucket is full (heavily skewed input) -- ring items [from, to) of a row are escaped exactly
__device__ __noinline__ void ring_escape(const uint32_t *row, uint32_t from, uint32_t to, uint32_t *table_k)
{
#pragma unroll 1
    for (uint32_t i = from; i < to; ++i) red_kmers_of_item(row[mod_cap(i)], table_k, 1u);
}

// flush duty of one warp: copy the completed chunks of the rows it owns to HBM and raise their limits.  Returns true when every
// warp had finished before the snapshot (so it held everything).  Copies are LSU copies, 8 lanes per row and 4 rows per
// instruction as in the synchronous flush.
__device__ __forceinline__ uint4 lds128(uint32_t saddr)
{
    uint4 r;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "r"(saddr) : "memory");
    return r;
}
__device__ __forceinline__ bool ring_flush(P1Ring &sm, int warp, int lane, uint32_t *gbuf0, uint32_t *my_gbuf, uint64_t bstride, uint32_t cap_front,
                                           uint32_t *table_k)
{
    constexpr int kSlots = (kOwn + 31) / 32;
    const uint32_t cursor_sa = (uint32_t)__cvta_generic_to_shared(sm.cursor), stage_sa = (uint32_t)__cvta_generic_to_shared(sm.stage);
    const uint32_t row0 = (uint32_t)warp * kOwn;
    // the cursors first, THEN every warp's progress: an even value = no claim of that warp was in flight (what the cursors show of
    // it is stored); an odd value = wait until it changes (the warp is between its claims and its stores: a few hundred cycles)
    uint32_t w[kSlots];
#pragma unroll
    for (int sl = 0; sl < kSlots; ++sl) {
        const uint32_t r = 32u * sl + lane, bl = row0 + r;
        w[sl] = (r < (uint32_t)kOwn && bl < (uint32_t)kNB) ? lds_volatile(cursor_sa + 4u * bl) : (uint32_t)kCap;
    }
    const uint32_t progress_sa = (uint32_t)__cvta_generic_to_shared(sm.progress) + 4u * (uint32_t)(lane < kProdWarps ? lane : 0);
    const uint32_t p = lds_volatile(progress_sa);
    const bool all_done = __all_sync(0xffffffffu, p == kDone);
    if (__any_sync(0xffffffffu, p & 1u)) {
        while (__any_sync(0xffffffffu, (p & 1u) && lds_volatile(progress_sa) == p)) { }
    }
    const uint32_t cap4 = cap_front & ~3u;
    const uint32_t sub = lane >> 3, c = lane & 7;
    uint32_t packed[kSlots], n4[kSlots];
#pragma unroll
    for (int sl = 0; sl < kSlots; ++sl) {
        const uint32_t bl = row0 + 32u * sl + lane;
        const uint32_t lim = w[sl] & 0xFFFFu, f = lim - (uint32_t)kCap;  // f < 52: the ring position of the first item to copy
        n4[sl] = min((min(w[sl] >> 16, lim) - f) & ~3u, 32u);  // at most 8 chunks per row and round: one per lane of its group of 8
        packed[sl] = 0;
        if (n4[sl]) {
            const uint32_t off = sm.goff[bl], ncp = min(n4[sl], cap4 - off);
            packed[sl] = (off >> 2) | ((f >> 2) << 20) | ((ncp >> 2) << 24);  // chunks: offset (20 bits), ring start (4), count (4)
            if (ncp < n4[sl]) ring_escape(&sm.stage[bl * kCap], f + ncp, f + n4[sl], table_k);
            sm.goff[bl] = off + ncp;
        }
    }
    // lane (sub, c) copies chunk c of row 4 * step + sub.  Loads are unconditional (an off lane reads its row's first chunk:
    // ptxas turns a predicated 128-bit load with a defined "off" value into a load + four predicated moves + two clears);
    // the destination is one 32-bit chunk index (gbuf is below 2^32 chunks of 16 bytes, checked by the launcher) and one IMAD.WIDE
    const uint32_t lane_sa = stage_sa + (row0 + sub) * (4u * kCap);
    const uint32_t row_chunks = (uint32_t)(bstride >> 2);  // chunks between the regions of two consecutive buckets
    uint32_t lane_chunk = (uint32_t)(((uint64_t)(my_gbuf - gbuf0)) >> 2) + (row0 + sub) * row_chunks + c;
    uint4 *const gchunks = reinterpret_cast<uint4 *>(gbuf0);
#pragma unroll
    for (int st0 = 0; st0 < kOwn / 4; st0 += 8) {
        uint32_t pk[8];
        uint4 v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            const int st = st0 + j;  // step: rows row0 + 4 * st .. + 3
            pk[j] = st < kOwn / 4 ? __shfl_sync(0xffffffffu, packed[st / 8], 4 * (st & 7) + sub) : 0u;
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (st0 + j >= kOwn / 4) continue;
            const uint32_t nc = pk[j] >> 24, t0 = ((pk[j] >> 20) & 15u) + c;
            const uint32_t c0 = min(t0, t0 - 13u);  // ring chunk of copy chunk c (13 chunks per row)
            const uint32_t row_sa = lane_sa + (uint32_t)(st0 + j) * (16u * kCap);
            v[j] = lds128(c < nc ? row_sa + 16u * c0 : row_sa);
        }
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            if (st0 + j >= kOwn / 4) continue;
            stg128_if(gchunks + (lane_chunk + (pk[j] & 0xFFFFFu)), v[j], c < (pk[j] >> 24));
            lane_chunk += 4u * row_chunks;
        }
    }
    // raise the limits (the copies above have READ their slots: the global stores depend on the loaded registers)
#pragma unroll
    for (int sl = 0; sl < kSlots; ++sl) {
        if (n4[sl]) {
            const uint32_t lim = w[sl] & 0xFFFFu, nf = lim - (uint32_t)kCap + n4[sl], rb = nf >= kRebase ? kRebase : 0u;
            uint32_t cur = w[sl];
            for (;;) {
                const uint32_t neww = ((min(cur >> 16, lim) - rb) << 16) | (nf + (uint32_t)kCap - rb);
                const uint32_t old = atomicCAS(&sm.cursor[row0 + 32u * sl + lane], cur, neww);
                if (old == cur) break;
                cur = old;
            }
        }
    }
    return all_done;
}

__global__ void __launch_bounds__(kRingThreads, 1)
bucketize16_ring_kernel(const uint8_t *__restrict__ s, uint64_t lo, uint64_t n_witers, uint32_t *__restrict__ gbuf, uint32_t cap_cb, uint32_t cap_front,
                        uint32_t *__restrict__ gcount, uint32_t *__restrict__ bucket_total, uint32_t *__restrict__ table_k, uint8_t *__restrict__ flags,
                        fkb_partials *__restrict__ P)
{
    extern __shared__ __align__(16) uint8_t smem_raw[];
    P1Ring &sm = *reinterpret_cast<P1Ring *>(smem_raw);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t cursor_sa = (uint32_t)__cvta_generic_to_shared(sm.cursor), stage_sa = (uint32_t)__cvta_generic_to_shared(sm.stage);

    for (int b = threadIdx.x; b < kNB; b += kRingThreads) { sm.cursor[b] = (uint32_t)kCap; sm.goff[b] = 0; sm.gback[b] = 0; }
    if (threadIdx.x < 32) { sm.progress[threadIdx.x] = 0; sm.junk[threadIdx.x] = 0; }
    if (threadIdx.x < 16) sm.ev[threadIdx.x] = 0;
    __syncthreads();

    uint32_t *const my_gbuf = gbuf + seg_base(0, blockIdx.x, gridDim.x, cap_cb);  // + bucket * bucket_stride
    uint32_t t_unknown = 0, t_dummy = 0, n_fast = 0;
    unsigned long long t_windows = 0, t_valid = 0;

    if (warp >= kProdWarps) {  // dedicated flusher warps (FKB2_FLUSH_WARPS > 0): walk the row groups until every producer has finished
        const uint64_t bstride = bucket_stride(gridDim.x, cap_cb);
        for (;;) {
            bool done = true;
            for (int o = warp - kProdWarps; o < kOwners; o += (kFlushWarps ? kFlushWarps : 1))
                done = ring_flush(sm, o, lane, gbuf, my_gbuf, bstride, cap_front, table_k) && done;
            if (done) break;
        }
    } else {
        const uint64_t bstride = bucket_stride(gridDim.x, cap_cb);
        const uint64_t n_warps = (uint64_t)gridDim.x * kProdWarps, gw = (uint64_t)blockIdx.x * kProdWarps + warp;
        const uint64_t q = n_witers / n_warps, rem = n_witers % n_warps;
        const uint32_t my_iters = (uint32_t)(q + (gw < rem ? 1 : 0));
        const uint64_t my_first = gw * q + (gw < rem ? gw : rem);
        const uint64_t region = lo + my_first * kWSpan, hi = lo + n_witers * kWSpan;
        const bool edge_first = (lane == 0) && (region == lo), edge_last = (lane == 31) && (region + (uint64_t)my_iters * kWSpan == hi);
        const uint32_t progress_sa = (uint32_t)__cvta_generic_to_shared(sm.progress) + 4u * (uint32_t)warp;

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

        uint32_t fc = 1u + (uint32_t)(warp % kFlushEvery);  // flush duty when it reaches 0
        for (uint32_t it = 0; it < my_iters; ++it) {
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

            // ---- the 8 items of this chunk: 8 shared atomics back to back (the claim in the bucket's ring), the encode of
            //      iteration it+2 between them and the 8 stores that wait for their results ----
            uint32_t ovf = 0;
            ValidAcc va;
            Group enc[kG];
            const bool ld_full = it + 3 < my_iters, ld_halo = (it + 3 == my_iters) && lane == 0;
            const uint8_t *const p3 = lane_base + (uint64_t)(it + 3) * kWSpan;
#if FKB2_PREFETCH > 0
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
                    if constexpr (ALL) pos[n] = atoms_add(ca, 0x10000u);
                    else {
                        const uint32_t e = emit[ItemPos<n>::g] & (1u << (15 - ItemPos<n>::i));
                        const uint32_t ps = atoms_add(e ? ca : cursor_sa + 4u * (uint32_t)(kNB + lane), 0x10000u);
                        pos[n] = e ? ps : kNotEmitted;
                    }
                };
                if (lane == 0) sts_publish(progress_sa, 2u * it + 1u);  // claims in flight
                slot(std::integral_constant<int, 0>{}); slot(std::integral_constant<int, 1>{}); slot(std::integral_constant<int, 2>{});
                slot(std::integral_constant<int, 3>{}); slot(std::integral_constant<int, 4>{}); slot(std::integral_constant<int, 5>{});
                slot(std::integral_constant<int, 6>{}); slot(std::integral_constant<int, 7>{});
#pragma unroll
                for (int g = 0; g < kG; ++g) {
                    enc[g].code = pack_codes_fast(raw[g], va);
                    enc[g].valid = 0xFFFFu;
                    raw[g] = ldg128_if(p3 + 16 * g, ld_full || (g == 0 && ld_halo));
                }
                uint32_t good = 0xFFFFFFFFu;  // sign bit survives iff every real claim of this lane was good
#pragma unroll
                for (int n = 0; n < kItems; ++n) {
                    const uint32_t ps = pos[n];
                    const uint32_t t = ps * 0xFFFF0001u;  // bits 31..16 = claimed - limit: negative = the slot is ours
                    const uint32_t cl = ps >> 16, sl = min(cl, cl - (uint32_t)kCap);  // claimed mod 52 (good claims are below 104)
                    const uint32_t sa = mad_u32(sl, 4u, mad_u32(bk[n], 4u * kCap, stage_sa));
                    sts32_if_neg(sa, f[n], t);
                    if constexpr (ALL) good &= t;
                    else good &= (ps == kNotEmitted) ? 0xFFFFFFFFu : t;
                }
                // every ring store of this iteration is issued: publish, THEN deal with bad claims (the flusher must never wait for them)
                __syncwarp();
                if (lane == 0) sts_publish(progress_sa, 2u * it + 2u);
                if (!(good >> 31)) {
                    // a ring is full (a bucket more popular than the flusher keeps up with): straight to the BACK part of this CTA's
                    // region of the bucket; only when that is full too is the item escaped exactly with global reds
#pragma unroll
                    for (int n = 0; n < kItems; ++n) {
                        const uint32_t t = pos[n] * 0xFFFF0001u;
                        if (pos[n] != kNotEmitted && !(t >> 31)) {
                            if ((t >> 16) > 0x6000u) {  // far past the limit: let the flusher reset the row before the 16-bit field can wrap
                                const uint32_t ca = mad_u32(bk[n], 4u, cursor_sa);
                                while ((((lds_volatile(ca) * 0xFFFF0001u) >> 16) & 0xFFFFu) - 0x1000u < 0x7000u) __nanosleep(200);
                            }
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

            // ---- flush duty for the rows this warp owns ----
            if (kFlushWarps == 0 && --fc == 0u) {
                ring_flush(sm, warp, lane, gbuf, my_gbuf, bstride, cap_front, table_k);
                fc = (uint32_t)kFlushEvery;
            }
        }
        __syncwarp();
        if (lane == 0) sts_publish(progress_sa, kDone);
        // keep serving my rows until every warp has finished; the last round starts after that and takes all whole chunks
        if (kFlushWarps == 0)
            while (!ring_flush(sm, warp, lane, gbuf, my_gbuf, bstride, cap_front, table_k)) __nanosleep(3000);
    }
    __syncthreads();

    // ---- end: the <= 3 items still in every ring, then this CTA's fill of every bucket ----
    if (threadIdx.x < 9) {
        const uint32_t v = sm.ev[threadIdx.x];
        if (v) {
            unsigned long long *dst = threadIdx.x < 4 ? &P->head_base[threadIdx.x] : (threadIdx.x < 8 ? &P->short_first[threadIdx.x - 4] : &P->runs_ge_k);
            atomicAdd(dst, (unsigned long long)v);
        }
    }
    for (int b = threadIdx.x; b < kNB; b += kRingThreads) {
        const uint32_t w = sm.cursor[b], lim = w & 0xFFFFu, fl = lim - (uint32_t)kCap, cnt = min(w >> 16, lim) - fl;
        uint32_t off = sm.goff[b];
        uint32_t *dst = my_gbuf + (uint64_t)b * bucket_stride(gridDim.x, cap_cb);
        for (uint32_t i = 0; i < cnt; ++i) {
            const uint32_t item = sm.stage[b * kCap + mod_cap(fl + i)];
            if (off < cap_front) dst[off++] = item;
            else red_kmers_of_item(item, table_k, 1u);
        }
        gcount[2 * ((uint64_t)b * gridDim.x + blockIdx.x)] = off;
        const uint32_t back = min(sm.gback[b], cap_cb - cap_front);
        gcount[2 * ((uint64_t)b * gridDim.x + blockIdx.x) + 1] = back;
        if (off + back) atomicAdd(&bucket_total[b], off + back);
    }
    {
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
}

