// fkb_smallk.cu -- VARIANT_SMEM: the single-pass count path for k <= 8.
//
// For k <= 8 the whole 4^k table fits into ONE SM's shared memory (k <= 7: 4^k x u32 <= 64 KiB; k = 8: 65536 x u16 =
// 128 KiB), so nothing has to be routed: every CTA streams its share of the stripped stream once, counts every window
// with one shared-memory atomic into its private copy of the table and adds the copy to the global table at the end.
// Compared with the bucketed path (fkb_bucket.cu: W-mers at stride S routed through HBM, folded afterwards) this has no
// W-windows, hence no left-over k-mers at run boundaries and no second pass: dirty sequence (N runs, soft-masked lower
// case -- BASELINE.json's config 5) costs one predicate per position instead of the general path of pass 1, which is
// what made that input run at 8.5 % of the roofline in round 1 (profiles/r01_sweep_configs_k6_11.txt).
//
// Semantics are those of the reference's scan (findKmer/src/findKmer.cpp:962-1069): a window is counted iff its k bytes
// are all in {A,C,G,T} (base2int :567-589; anything else resets, :1019-1024); the per-run events -- run length reaches k
// (:1044-1057), run still shorter than k (:1059-1062) -- go through the same warp-cooperative handler as in pass 1.
// Bound: the shared-memory data pipe, one atomic per base (32 random lanes = ~3.5 wavefronts).
#include <atomic>
#include <type_traits>

#include "fkb_kernels.cuh"
#include "fkb_stream.cuh"

namespace fkb {

namespace {

constexpr uint64_t kUnit = 1024;      // interior granularity: a whole number of warp iterations for every G below
constexpr uint32_t kEvTile = 2048;    // warp iterations between two flushes of the 32-bit event counters

template <int K> struct SmCfg {
    static constexpr bool kPacked = (K == 8);                     // 16-bit counters, two per word, drained at 0x8000
#ifndef FKB_SK8_THREADS
#define FKB_SK8_THREADS 1024
#endif
    static constexpr int kThreads = kPacked ? FKB_SK8_THREADS : 512;
    static constexpr int kMinBlocks = kPacked ? 1 : 2;
    static constexpr uint32_t kTableBytes = kPacked ? (2u << (2 * K)) : (4u << (2 * K));
    // 16-byte groups per lane per iteration.  Measured on B200 (profiles/r02_sweep_smallk.txt): k = 8 is faster with one group
    // (config 4 1.44 -> 1.42 ms, config 5 4.21 -> 3.86 ms: the packed path keeps 8 old values per batch live), k <= 7 with two
    static constexpr int kGroups = kPacked ? 1 : 2;
};

// shared-memory reduction (no return value).  ptxas turns a PREDICATED shared atomic into a branch around it (BSSY / BRA /
// ATOMS / BSYNC per item), so the general path redirects an item that is not a window to a junk word instead of skipping it.
__device__ __forceinline__ void reds_inc(uint32_t saddr) { asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(saddr) : "memory"); }
__device__ __forceinline__ uint32_t atoms_add(uint32_t saddr, uint32_t v)
{
    uint32_t old;
    asm volatile("atom.shared.add.u32 %0, [%1], %2;" : "=r"(old) : "r"(saddr), "r"(v) : "memory");
    return old;
}
__device__ __forceinline__ uint32_t mad_u32(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
}

// a 16-bit counter of the packed table saw 0x8000: take 32768 out and hand them to the global table (exact: the trigger is
// unique -- only the increment that READS 0x8000 comes here -- and fewer than 32768 increments can be in flight)
__device__ __noinline__ void drain_packed(uint32_t word_sa, uint32_t hi_half, uint32_t kmer, uint32_t *table_k)
{
    asm volatile("red.shared.add.u32 [%0], %1;" ::"r"(word_sa), "r"(hi_half ? 0x80000000u : 0xFFFF8000u) : "memory");  // - 0x8000 in that half
    red_add_u32(table_k + kmer, 32768u);
}

template <int K>
__global__ void __launch_bounds__(SmCfg<K>::kThreads, SmCfg<K>::kMinBlocks)
count_smem_kernel(const uint8_t *__restrict__ s, uint64_t lo, uint64_t n_witers, uint32_t *__restrict__ table_k, uint8_t *__restrict__ flags,
                  fkb_partials *__restrict__ P)
{
    constexpr int kThreads = SmCfg<K>::kThreads, kWarps = kThreads / 32;
    constexpr bool PACKED = SmCfg<K>::kPacked;
    constexpr int kG = SmCfg<K>::kGroups, kCH = 16 * kG;      // bytes per lane per iteration
    constexpr uint64_t kWSpan = 32ull * kCH;                   // bytes per warp iteration
    constexpr uint32_t MASK = (1u << (2 * K)) - 1u;
    extern __shared__ __align__(16) uint8_t smem_raw[];
    uint32_t *cnt = reinterpret_cast<uint32_t *>(smem_raw);
    __shared__ uint32_t ev[16];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t cnt_sa = (uint32_t)__cvta_generic_to_shared(cnt);

    for (uint32_t i = threadIdx.x; i < SmCfg<K>::kTableBytes / 16; i += kThreads) reinterpret_cast<uint4 *>(cnt)[i] = make_uint4(0, 0, 0, 0);
    if (threadIdx.x < 16) ev[threadIdx.x] = 0;
    __syncthreads();

    // this warp's contiguous share of the interior
    const uint64_t n_warps = (uint64_t)gridDim.x * kWarps, gw = (uint64_t)blockIdx.x * kWarps + warp;
    const uint64_t q = n_witers / n_warps, rem = n_witers % n_warps;
    const uint32_t my_iters = (uint32_t)(q + (gw < rem ? 1 : 0));
    const uint64_t my_first = gw * q + (gw < rem ? gw : rem);
    const uint32_t max_iters = (uint32_t)(q + (rem ? 1 : 0));  // CTA-uniform trip count (barriers of the event flush)
    const uint8_t *const lane_base = s + lo + my_first * kWSpan + (uint64_t)lane * kCH;

    uint32_t t_unknown = 0, t_dummy = 0, n_fast = 0;
    unsigned long long t_windows = 0, t_valid = 0;

    // `r0` = raw bytes of the iteration about to be processed; the loads of iteration it + 1 are issued as soon as iteration
    // it's bytes have been encoded (a whole iteration -- several thousand cycles -- before they are needed; no second buffer:
    // its registers were spilled at 1024 threads x 64 registers)
    uint4 r0[kG];
    Group carry = pack_group(ldg128(lane_base - (uint64_t)lane * kCH - 16), t_dummy);  // the 16 bytes in front of the region (same address for all lanes)
#pragma unroll
    for (int g = 0; g < kG; ++g) r0[g] = ldg128_if(lane_base + 16 * g, my_iters > 0);

    for (uint32_t tile0 = 0; tile0 < max_iters; tile0 += kEvTile) {
        const uint32_t tile_end = min(tile0 + kEvTile, my_iters);
        for (uint32_t it = tile0; it < tile_end; ++it) {
            Group grp[kG + 1];  // grp[0] = the 16 bytes in front of my chunk, grp[1..G] = my chunk
            {
                // fast encode: codes only, validity accumulated over the lane's 16*G bytes and tested once; exact per-byte masks
                // (and the unknown-character count) only for a lane that holds a byte that is not a base
                ValidAcc va;
#pragma unroll
                for (int g = 0; g < kG; ++g) {
                    grp[g + 1].code = pack_codes_fast(r0[g], va);
                    grp[g + 1].valid = 0xFFFFu;
                }
                if (va.bad()) {
#pragma unroll
                    for (int g = 0; g < kG; ++g) grp[g + 1].valid = exact_valid16(r0[g], t_unknown);
                }
            }
            const uint32_t up_c = __shfl_sync(0xffffffffu, grp[kG].code, (lane + 31) & 31);
            const uint32_t up_v = __shfl_sync(0xffffffffu, grp[kG].valid, (lane + 31) & 31);
            grp[0].code = lane == 0 ? carry.code : up_c;
            grp[0].valid = lane == 0 ? carry.valid : up_v;
            carry.code = up_c;   // lane 0: lane 31's last group = the bytes in front of the next iteration
            carry.valid = up_v;

            // lane state.  WHOLE: my bytes and the K bytes in front of them are bases -- every window that ends in my chunk exists and
            // none of them is the first of its run: the predicate-free count block below, no events.  Otherwise (a run boundary
            // within reach, or no base at all) the lane skips that block; its windows and its per-run events go through the
            // warp-cooperative handler, which adds a window straight to T_k with a global red (a few lanes per run boundary: ~1 %
            // of the windows of config 5).
            uint32_t own = 0xFFFFu;
#pragma unroll
            for (int g = 1; g <= kG; ++g) own &= grp[g].valid;
            constexpr uint32_t kCtx = (1u << K) - 1u;
            const bool whole = own == 0xFFFFu && (grp[0].valid & kCtx) == kCtx;
            if (__all_sync(0xffffffffu, whole)) {
                ++n_fast;
            } else {
                if (whole) {
                    t_windows += kCH;
                    t_valid += kCH;
                }
#pragma unroll
                for (int g = 1; g <= kG; ++g) {
                    const uint32_t m = (grp[g - 1].valid << 16) | grp[g].valid;
                    const uint32_t rk0 = runs_of<K>(m) & 0xFFFFu;               // bit 15-i: a full window ends at byte i of the group
                    const uint32_t rk = whole ? 0u : rk0;
                    const uint32_t first_k = rk & ~(m >> K) & 0xFFFFu;          // run length is exactly k here
                    const uint32_t shorts = whole ? 0u : (grp[g].valid & ~rk0);  // valid base whose run is still shorter than k
                    t_windows += __popc(rk);
                    if (!whole) t_valid += __popc(grp[g].valid);
                    const uint32_t who = __ballot_sync(0xffffffffu, (rk | shorts) != 0);
                    if (who) warp_group_events(who, rk, first_k, shorts, m, grp[g - 1].code, grp[g].code, K, flags, ev, table_k);
                }
            }

            // next iteration's bytes: issued HERE, behind the event handler -- a call waits for every load in flight
#pragma unroll
            for (int g = 0; g < kG; ++g) r0[g] = ldg128_if(lane_base + (uint64_t)(it + 1) * kWSpan + 16 * g, it + 1 < my_iters);

            // the 16 windows of every group of a WHOLE lane: one shared-memory atomic each, 8 back to back, no predicates
            if (whole) {
#pragma unroll
                for (int g = 1; g <= kG; ++g) {
#pragma unroll
                    for (int h = 0; h < 2; ++h) {
                        if constexpr (!PACKED) {
#pragma unroll
                            for (int i = 8 * h; i < 8 * h + 8; ++i) {
                                const uint32_t f = (i == 15) ? grp[g].code : __funnelshift_r(grp[g].code, grp[g - 1].code, 2 * (15 - i));
                                reds_inc(mad_u32(f & MASK, 4u, cnt_sa));
                            }
                        } else {
                            uint32_t f[8], old[8], any = 0;
#pragma unroll
                            for (int i = 8 * h; i < 8 * h + 8; ++i) {
                                const int j = i - 8 * h;
                                f[j] = (i == 15) ? grp[g].code : __funnelshift_r(grp[g].code, grp[g - 1].code, 2 * (15 - i));
                                const uint32_t sa = mad_u32(f[j] & 0xFFFEu, 2u, cnt_sa);   // word of counters 2n, 2n+1
                                const uint32_t inc = mad_u32(f[j] & 1u, 0xFFFFu, 1u);      // odd k-mer: the high half
                                old[j] = atoms_add(sa, inc);
                            }
#pragma unroll
                            for (int j = 0; j < 8; ++j) any |= old[j];
                            if (any & 0x80008000u) {  // some touched word has a half at or above 0x8000 (this item's or its neighbour's): look closer
#pragma unroll
                                for (int j = 0; j < 8; ++j) {
                                    const uint32_t x = f[j] & MASK;
                                    const uint32_t half = (x & 1u) ? (old[j] >> 16) : (old[j] & 0xFFFFu);
                                    if (half == 0x8000u) drain_packed(cnt_sa + (x & 0xFFFEu) * 2u, x & 1u, x, table_k);
                                }
                            }
                        }
                    }
                }
            }
        }
        // 32-bit event counters of this tile -> 64-bit partials
        __syncthreads();
        if (threadIdx.x < 9) {
            const uint32_t v = ev[threadIdx.x];
            if (v) {
                ev[threadIdx.x] = 0;
                unsigned long long *dst = threadIdx.x < 4 ? &P->head_base[threadIdx.x] : (threadIdx.x < 8 ? &P->short_first[threadIdx.x - 4] : &P->runs_ge_k);
                atomicAdd(dst, (unsigned long long)v);
            }
        }
        __syncthreads();
    }

    // the CTA's private table -> the global table (coalesced reds; zero counters are skipped)
    if constexpr (!PACKED) {
        for (uint32_t i = threadIdx.x; i < (1u << (2 * K)); i += kThreads) {
            const uint32_t v = cnt[i];
            if (v) red_add_u32(table_k + i, v);
        }
    } else {
        for (uint32_t i = threadIdx.x; i < (1u << (2 * K)); i += kThreads) {
            const uint32_t v = reinterpret_cast<const uint16_t *>(cnt)[i];
            if (v) red_add_u32(table_k + i, v);
        }
    }
    t_windows += (unsigned long long)n_fast * kCH;
    t_valid += (unsigned long long)n_fast * kCH;
    t_windows = warp_sum(t_windows);
    t_valid = warp_sum(t_valid);
    const unsigned long long t_unk = warp_sum((unsigned long long)t_unknown);
    if (lane == 0) {
        if (t_windows) atomicAdd(&P->n_windows, t_windows);
        if (t_valid) atomicAdd(&P->valid_bases, t_valid);
        if (t_unk) atomicAdd(&P->unknown_chars, t_unk);
    }
}

template <int K>
cudaError_t run_smallk(const LaunchInfo &li, const uint8_t *d_stream, uint64_t lo, uint64_t hi, uint32_t *d_table, uint8_t *d_flags,
                       fkb_partials *d_partials, cudaStream_t st, int *launches)
{
    static std::atomic<uint64_t> attr_done{0};  // the dynamic shared-memory opt-in is per device
    int dev = 0;
    cudaError_t e = cudaGetDevice(&dev);
    if (e != cudaSuccess) return e;
    const uint64_t dev_bit = 1ull << (dev & 63);
    if (SmCfg<K>::kTableBytes + 128 > 48 * 1024 && !(attr_done.load(std::memory_order_acquire) & dev_bit)) {
        e = cudaFuncSetAttribute(count_smem_kernel<K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)SmCfg<K>::kTableBytes + 128);
        if (e != cudaSuccess) return e;
        attr_done.fetch_or(dev_bit, std::memory_order_release);
    }
    constexpr uint64_t kWSpan = 32ull * 16 * SmCfg<K>::kGroups;
    const uint64_t n_witers = (hi - lo) / kWSpan;
    constexpr int kWarps = SmCfg<K>::kThreads / 32;
    if ((n_witers / kWarps) >> 32) return cudaErrorInvalidValue;
    // every CTA pays 4^k reds at the end: use fewer CTAs on short ranges (>= 16 KiB per warp)
    uint64_t ctas = (uint64_t)li.sm_count * SmCfg<K>::kMinBlocks;
    const uint64_t want = ((hi - lo) / 1024 + 16ull * kWarps - 1) / (16ull * kWarps);
    if (want < ctas) ctas = want ? want : 1;
    count_smem_kernel<K><<<(unsigned)ctas, SmCfg<K>::kThreads, SmCfg<K>::kTableBytes + 128, st>>>(d_stream, lo, n_witers, d_table, d_flags, d_partials);
    if (launches) ++*launches;
    return cudaGetLastError();
}

}  // namespace

uint64_t smallk_unit_bytes(int k) { return (k >= 1 && k <= 8) ? kUnit : 0; }

cudaError_t launch_count_smallk(const LaunchInfo &li, const uint8_t *d_stream, uint64_t lo, uint64_t hi, int k, uint32_t *d_table, uint8_t *d_flags,
                                fkb_partials *d_partials, cudaStream_t st, int *launches)
{
    switch (k) {
    case 1: return run_smallk<1>(li, d_stream, lo, hi, d_table, d_flags, d_partials, st, launches);
    case 2: return run_smallk<2>(li, d_stream, lo, hi, d_table, d_flags, d_partials, st, launches);
    case 3: return run_smallk<3>(li, d_stream, lo, hi, d_table, d_flags, d_partials, st, launches);
    case 4: return run_smallk<4>(li, d_stream, lo, hi, d_table, d_flags, d_partials, st, launches);
    case 5: return run_smallk<5>(li, d_stream, lo, hi, d_table, d_flags, d_partials, st, launches);
    case 6: return run_smallk<6>(li, d_stream, lo, hi, d_table, d_flags, d_partials, st, launches);
    case 7: return run_smallk<7>(li, d_stream, lo, hi, d_table, d_flags, d_partials, st, launches);
    case 8: return run_smallk<8>(li, d_stream, lo, hi, d_table, d_flags, d_partials, st, launches);
    default: return cudaErrorInvalidValue;
    }
}

}  // namespace fkb
