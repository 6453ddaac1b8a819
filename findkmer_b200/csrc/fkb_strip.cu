// fkb_strip.cu -- the stream contract on the DEVICE: record and line stripping as a stream compaction.
//
// Same rule as the host loader (fkb_loader.cpp), i.e. what the reference's scan loop does with three bytes
// (findKmer/src/findKmer.cpp:988-1011): drop every '\n'; a '>' that is not already inside a header line is kept as
// ONE sentinel byte and everything after it through the next '\n' is dropped; the first byte 0xFF outside a header ends
// the scan (char c = fgetc() aliases EOF, :975,:988).  Used when the caller's file image is in pinned memory: the raw bytes
// then cross PCIe once, with no host pass at all (profiles/r01_e2e_breakdown.txt: 16 host cores strip slower than PCIe copies).
//
// "Inside a header" is a 1-bit state carried along the file.  A piece of the file acts on it as  h -> a | (b & h)
// (a: the piece ends inside a header it started itself; b: the piece holds no '\n'), which composes associatively, so the
// state is resolved hierarchically: 16 bytes per thread (log-step segmented OR in a register), 32 threads by ballot,
// warps through shared memory, 4 KiB tiles by a single-block scan, chunks through a 32-byte state record in HBM.
//   pass A  strip_summarize : per tile (a, b, kept bytes before / after its first '\n', first kept 0xFF before / after)
//   pass B  strip_scan      : incoming state and output offset of every tile; the stop position
//   pass C  strip_compact   : recompute the keep masks with the true state, compact through shared memory, write
#include "fkb_kernels.cuh"

namespace fkb {

namespace {

constexpr int kTileThreads = 256;
constexpr int kTileBytes = kTileThreads * 16;  // 4 KiB
constexpr uint32_t kNone = 0xFFFFFFFFu;

struct TileSum {
    uint32_t cnt_pre, cnt_post;  // kept bytes before / after the tile's first '\n', assuming the tile starts outside a header
    uint32_t ff_pre, ff_post;    // tile-relative position of the first KEPT byte 0xFF before / after the first '\n' (kNone: none)
    uint32_t a, b;               // state transfer  h_out = a | (b & h_in)
};
struct TileIn {
    unsigned long long out_off;  // output offset of the tile's first kept byte
    uint32_t h_in, pad;
};

__device__ __forceinline__ uint4 ldg128(const uint8_t *p)
{
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// 4-bit mask (byte 0 -> bit 0) of the bytes of w equal to the splatted byte
__device__ __forceinline__ uint32_t eq_nibble(uint32_t w, uint32_t splat)
{
    const uint32_t u = w ^ splat;
    const uint32_t nz = (((u & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | u) & 0x80808080u;
    return (((nz ^ 0x80808080u) >> 7) * 0x01020408u) >> 24;
}
__device__ __forceinline__ uint32_t eq_mask16(const uint4 &g, uint32_t splat)
{
    return eq_nibble(g.x, splat) | (eq_nibble(g.y, splat) << 4) | (eq_nibble(g.z, splat) << 8) | (eq_nibble(g.w, splat) << 12);
}

// One thread's 16 bytes (bit i = byte i).  Given the state at byte 0, which bytes lie inside a header, and the state after byte 15.
struct Piece {
    uint32_t nl, gt, ff;  // 16-bit masks
    uint32_t inside0;     // 17 bits: inside-header mask assuming h_in = 0; bit 16 = state after the piece
    uint32_t pre;         // bits before the first '\n' (all 17 bits when there is none): these turn "inside" when h_in = 1
};
__device__ __forceinline__ Piece make_piece(const uint4 &g, uint32_t valid_bytes /* bytes beyond the chunk read as '\n'-free padding */)
{
    Piece p;
    const uint32_t vm = valid_bytes >= 16 ? 0xFFFFu : ((1u << valid_bytes) - 1u);
    p.nl = eq_mask16(g, 0x0A0A0A0Au) & vm;
    p.gt = eq_mask16(g, 0x3E3E3E3Eu) & vm;
    p.ff = eq_mask16(g, 0xFFFFFFFFu) & vm;
    // inside_i = gt_{i-1} | (!nl_{i-1} & inside_{i-1}), inside_0 = h_in: Kogge-Stone on (generate, propagate) bit vectors
    uint32_t gen = (p.gt << 1) & 0x1FFFFu, prop = ((~p.nl) << 1) & 0x1FFFFu;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        gen |= prop & (gen << d);
        prop &= (prop << d);
    }
    p.inside0 = gen & 0x1FFFFu;
    p.pre = p.nl ? ((p.nl & (0u - p.nl)) - 1u) : 0x1FFFFu;
    return p;
}

// state entering lane `lane`, given the per-lane (a, b) ballots and the state entering lane 0
__device__ __forceinline__ uint32_t state_before_lane(uint32_t A, uint32_t B, uint32_t h0, int lane)
{
    const uint32_t below = (1u << lane) - 1u;
    const uint32_t resets = ~B & below;  // earlier lanes that hold a '\n'
    if (resets) {
        const int i = 31 - __clz(resets);
        return (A & below & ~((1u << i) - 1u)) != 0;
    }
    return h0 | ((A & below) != 0);
}

// Block-wide resolution of the state entering every thread, given the state entering the tile.
// Returns h_in of this thread; *tile_a / *tile_b (valid in all threads) describe the whole tile.
__device__ __forceinline__ uint32_t resolve_states(uint32_t a, uint32_t b, uint32_t h_tile, uint32_t *tile_a, uint32_t *tile_b)
{
    __shared__ uint32_t s_a[kTileThreads / 32], s_b[kTileThreads / 32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const uint32_t A = __ballot_sync(0xffffffffu, a), B = __ballot_sync(0xffffffffu, b);
    if (lane == 0) {
        // the warp as one piece: with h0 = 0 its out state is "a of the last reset onwards"
        const uint32_t resets = ~B;
        uint32_t wa;
        if (resets) {
            const int i = 31 - __clz(resets);
            wa = (A & ~((1u << i) - 1u)) != 0;
        } else {
            wa = A != 0;
        }
        s_a[warp] = wa;
        s_b[warp] = (B == 0xffffffffu);
    }
    __syncthreads();
    uint32_t h = h_tile, ta = 0, tb = 1;  // h: state entering my warp; (ta, tb): the whole tile as one piece
    for (int w = 0; w < kTileThreads / 32; ++w) {
        const uint32_t wa = s_a[w], wb = s_b[w];
        if (w < warp) h = wa | (wb & h);
        ta = wa | (wb & ta);
        tb = wb & tb;
    }
    *tile_a = ta;
    *tile_b = tb;
    __syncthreads();
    return state_before_lane(A, B, h, lane);
}

__device__ __forceinline__ uint32_t block_sum(uint32_t v, uint32_t *scratch)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t t = 0;
    for (int w = 0; w < kTileThreads / 32; ++w) t += scratch[w];
    __syncthreads();
    return t;
}
__device__ __forceinline__ uint32_t block_min(uint32_t v, uint32_t *scratch)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = min(v, __shfl_xor_sync(0xffffffffu, v, o));
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    uint32_t t = kNone;
    for (int w = 0; w < kTileThreads / 32; ++w) t = min(t, scratch[w]);
    __syncthreads();
    return t;
}

__device__ __forceinline__ uint4 load_piece(const uint8_t *raw, uint64_t pos, uint64_t n)
{
    if (pos + 16 <= n) return ldg128(raw + pos);
    uint32_t w[4] = {0, 0, 0, 0};
    for (int i = 0; i < 16; ++i)
        if (pos + i < n) w[i >> 2] |= (uint32_t)raw[pos + i] << (8 * (i & 3));
    return make_uint4(w[0], w[1], w[2], w[3]);
}

// ---- pass A ----
__global__ void __launch_bounds__(kTileThreads) strip_summarize_kernel(const uint8_t *__restrict__ raw, uint64_t n, TileSum *__restrict__ sums)
{
    __shared__ uint32_t scratch[kTileThreads / 32];
    __shared__ uint32_t s_first_nl;
    const uint64_t tile = blockIdx.x, pos = tile * kTileBytes + (uint64_t)threadIdx.x * 16;
    const uint32_t valid = pos >= n ? 0u : (n - pos >= 16 ? 16u : (uint32_t)(n - pos));
    const Piece p = make_piece(load_piece(raw, pos, n), valid);
    const uint32_t a = (p.inside0 >> 16) & 1u, b = (p.nl == 0);
    uint32_t ta, tb;
    const uint32_t h_in = resolve_states(a, b, 0u, &ta, &tb);
    const uint32_t vm = valid >= 16 ? 0xFFFFu : ((1u << valid) - 1u);
    const uint32_t inside = (p.inside0 | (h_in ? p.pre : 0u)) & 0xFFFFu;
    const uint32_t keep = ~p.nl & ~inside & vm;
    // position of the tile's first '\n'
    if (threadIdx.x == 0) s_first_nl = kNone;
    __syncthreads();
    if (p.nl) atomicMin(&s_first_nl, threadIdx.x * 16u + (uint32_t)__ffs(p.nl) - 1u);
    __syncthreads();
    const uint32_t first_nl = s_first_nl, base = threadIdx.x * 16u;
    // bits of this piece that lie before the tile's first '\n'
    uint32_t pre_bits;
    if (first_nl == kNone || base + 16 <= first_nl) pre_bits = 0xFFFFu;
    else if (base > first_nl) pre_bits = 0;
    else pre_bits = (1u << (first_nl - base)) - 1u;
    const uint32_t cnt_pre = block_sum(__popc(keep & pre_bits), scratch);
    const uint32_t cnt_post = block_sum(__popc(keep & ~pre_bits), scratch);
    const uint32_t kf = keep & p.ff;
    const uint32_t ff_pre = block_min((kf & pre_bits) ? base + __ffs(kf & pre_bits) - 1u : kNone, scratch);
    const uint32_t ff_post = block_min((kf & ~pre_bits) ? base + __ffs(kf & ~pre_bits) - 1u : kNone, scratch);
    if (threadIdx.x == 0) sums[tile] = TileSum{cnt_pre, cnt_post, ff_pre, ff_post, ta, tb};
}

// ---- pass B: one block; a thread owns a contiguous run of tiles ----
struct Agg {  // a run of tiles as a function of the state entering it
    unsigned long long c0, c1;  // kept bytes if h_in = 0 / 1
    unsigned long long s0, s1;  // absolute position of the first kept 0xFF if h_in = 0 / 1 (~0: none)
    uint32_t a, b;
};
__device__ __forceinline__ Agg agg_identity() { return Agg{0, 0, ~0ull, ~0ull, 0u, 1u}; }
__device__ __forceinline__ Agg agg_tile(const TileSum &t, uint64_t tile_pos)
{
    Agg g;
    g.c0 = (unsigned long long)t.cnt_pre + t.cnt_post;
    g.c1 = t.cnt_post;
    const unsigned long long post = t.ff_post == kNone ? ~0ull : tile_pos + t.ff_post;
    const unsigned long long pre = t.ff_pre == kNone ? ~0ull : tile_pos + t.ff_pre;
    g.s0 = pre < post ? pre : post;
    g.s1 = post;
    g.a = t.a;
    g.b = t.b;
    return g;
}
__device__ __forceinline__ Agg agg_then(const Agg &x, const Agg &y)  // x first, then y
{
    Agg r;
    const uint32_t h0 = x.a, h1 = x.a | x.b;  // state entering y when x was entered with 0 / 1
    r.c0 = x.c0 + (h0 ? y.c1 : y.c0);
    r.c1 = x.c1 + (h1 ? y.c1 : y.c0);
    const unsigned long long y0 = h0 ? y.s1 : y.s0, y1 = h1 ? y.s1 : y.s0;
    r.s0 = x.s0 < y0 ? x.s0 : y0;
    r.s1 = x.s1 < y1 ? x.s1 : y1;
    r.a = y.a | (y.b & x.a);
    r.b = y.b & x.b;
    return r;
}

struct StripState {              // carried from chunk to chunk in HBM
    unsigned long long out_off;  // stripped bytes written so far
    unsigned long long stop_pos; // absolute raw position of the terminating 0xFF (~0: none yet)
    uint32_t h;                  // inside a header at the end of the previous chunk
    uint32_t pad;
};

constexpr int kScanThreads = 512;
__global__ void __launch_bounds__(kScanThreads) strip_scan_kernel(const TileSum *__restrict__ sums, uint64_t n_tiles, uint64_t chunk_pos,
                                                          StripState *__restrict__ st, TileIn *__restrict__ tin)
{
    __shared__ Agg s_agg[kScanThreads];
    __shared__ unsigned long long s_off[kScanThreads];
    __shared__ uint32_t s_h[kScanThreads];
    const uint64_t per = (n_tiles + kScanThreads - 1) / kScanThreads;
    const uint64_t t0 = threadIdx.x * per, t1 = t0 + per < n_tiles ? t0 + per : n_tiles;
    Agg mine = agg_identity();
    for (uint64_t t = t0; t < t1; ++t) mine = agg_then(mine, agg_tile(sums[t], chunk_pos + t * kTileBytes));
    s_agg[threadIdx.x] = mine;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t h = st->h;
        unsigned long long off = st->out_off, stop = st->stop_pos;
        const bool already_stopped = stop != ~0ull;  // an earlier chunk ended the scan: this one contributes nothing
        for (int i = 0; i < kScanThreads; ++i) {
            s_h[i] = h;
            s_off[i] = off;
            const Agg &g = s_agg[i];
            const unsigned long long s = h ? g.s1 : g.s0;
            if (s < stop) stop = s;
            off += h ? g.c1 : g.c0;
            h = g.a | (g.b & h);
        }
        if (!already_stopped) {
            st->h = h;
            st->stop_pos = stop;
            st->out_off = off;  // corrected by pass C when the stop lies in this chunk
        }
    }
    __syncthreads();
    uint32_t h = s_h[threadIdx.x];
    unsigned long long off = s_off[threadIdx.x];
    for (uint64_t t = t0; t < t1; ++t) {
        tin[t] = TileIn{off, h, 0u};
        const TileSum ts = sums[t];
        off += (h ? 0u : ts.cnt_pre) + ts.cnt_post;
        h = ts.a | (ts.b & h);
    }
}

// ---- pass C ----
__global__ void __launch_bounds__(kTileThreads) strip_compact_kernel(const uint8_t *__restrict__ raw, uint64_t n, uint64_t chunk_pos,
                                                                     const TileIn *__restrict__ tin, StripState *__restrict__ st,
                                                                     uint8_t *__restrict__ out)
{
    __shared__ uint32_t scratch[kTileThreads / 32];
    __shared__ uint32_t s_warp_off[kTileThreads / 32];
    __shared__ __align__(16) uint8_t s_bytes[kTileBytes];
    const uint64_t tile = blockIdx.x, rel = tile * kTileBytes + (uint64_t)threadIdx.x * 16, pos = chunk_pos + rel;
    const TileIn ti = tin[tile];
    const unsigned long long stop = st->stop_pos;
    if (chunk_pos + tile * kTileBytes >= stop) return;  // everything from the stop position on is dropped
    const uint32_t valid = rel >= n ? 0u : (n - rel >= 16 ? 16u : (uint32_t)(n - rel));
    const uint4 g = load_piece(raw, rel, n);
    const Piece p = make_piece(g, valid);
    const uint32_t a = (p.inside0 >> 16) & 1u, b = (p.nl == 0);
    uint32_t ta, tb;
    const uint32_t h_in = resolve_states(a, b, ti.h_in, &ta, &tb);
    const uint32_t vm = valid >= 16 ? 0xFFFFu : ((1u << valid) - 1u);
    const uint32_t inside = (p.inside0 | (h_in ? p.pre : 0u)) & 0xFFFFu;
    uint32_t keep = ~p.nl & ~inside & vm;
    if (pos + 16 > stop) keep &= (pos >= stop) ? 0u : ((1u << (uint32_t)(stop - pos)) - 1u);
    // exclusive prefix of the kept counts inside the tile
    const uint32_t cnt = __popc(keep);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
        if (lane >= o) incl += v;
    }
    if (lane == 31) scratch[warp] = incl;
    __syncthreads();
    if (threadIdx.x == 0) {
        uint32_t run = 0;
        for (int w = 0; w < kTileThreads / 32; ++w) { s_warp_off[w] = run; run += scratch[w]; }
        scratch[0] = run;  // kept bytes of the tile
    }
    __syncthreads();
    const uint32_t total = scratch[0];
    uint32_t r = s_warp_off[warp] + incl - cnt;
    // the thread that holds the stop byte knows the final length of the stream
    if (stop >= pos && stop < pos + 16) st->out_off = ti.out_off + r + cnt;
    const uint32_t w4[4] = {g.x, g.y, g.z, g.w};
    uint32_t k = keep;
    while (k) {
        const int i = __ffs(k) - 1;
        k &= k - 1;
        s_bytes[r++] = (uint8_t)(w4[i >> 2] >> (8 * (i & 3)));
    }
    __syncthreads();
    uint8_t *dst = out + ti.out_off;
    // coalesced copy out: bytes up to the first 16-byte boundary of the destination, then 128-bit stores
    const uint32_t head = min(total, (uint32_t)((16 - ((uintptr_t)dst & 15)) & 15));
    if (threadIdx.x < head) dst[threadIdx.x] = s_bytes[threadIdx.x];
    const uint32_t body = (total - head) & ~15u;
    for (uint32_t i = threadIdx.x * 16u; i < body; i += kTileThreads * 16u) {
        uint32_t v[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            const uint32_t o = head + i + 4 * j;  // unaligned in shared memory: assemble from bytes
            v[j] = s_bytes[o] | (s_bytes[o + 1] << 8) | (s_bytes[o + 2] << 16) | ((uint32_t)s_bytes[o + 3] << 24);
        }
        *reinterpret_cast<uint4 *>(dst + head + i) = make_uint4(v[0], v[1], v[2], v[3]);
    }
    const uint32_t tail0 = head + body;
    if (tail0 + threadIdx.x < total) dst[tail0 + threadIdx.x] = s_bytes[tail0 + threadIdx.x];
}

}  // namespace

size_t strip_scratch_bytes(uint64_t chunk_bytes)
{
    const uint64_t tiles = (chunk_bytes + kTileBytes - 1) / kTileBytes;
    return tiles * (sizeof(TileSum) + sizeof(TileIn)) + 256;
}
size_t strip_state_bytes() { return sizeof(StripState); }

// Strip raw[0, n) -- a chunk that starts at absolute file position chunk_pos -- and append to `out` at state->out_off.
// `state` (HBM) chains consecutive chunks; initialise it with strip_state_init.
cudaError_t launch_strip_chunk(const uint8_t *d_raw, uint64_t n, uint64_t chunk_pos, void *d_state, void *d_scratch, uint8_t *d_out,
                               cudaStream_t st, int *launches)
{
    if (n == 0) return cudaSuccess;
    const uint64_t tiles = (n + kTileBytes - 1) / kTileBytes;
    TileSum *sums = reinterpret_cast<TileSum *>(d_scratch);
    TileIn *tin = reinterpret_cast<TileIn *>(reinterpret_cast<uint8_t *>(d_scratch) + ((tiles * sizeof(TileSum) + 127) & ~127ull));
    strip_summarize_kernel<<<(unsigned)tiles, kTileThreads, 0, st>>>(d_raw, n, sums);
    strip_scan_kernel<<<1, kScanThreads, 0, st>>>(sums, tiles, chunk_pos, reinterpret_cast<StripState *>(d_state), tin);
    strip_compact_kernel<<<(unsigned)tiles, kTileThreads, 0, st>>>(d_raw, n, chunk_pos, tin, reinterpret_cast<StripState *>(d_state), d_out);
    if (launches) *launches += 3;
    return cudaGetLastError();
}

void strip_state_init(void *host_state, uint64_t out_off, int in_header)
{
    StripState s;
    s.out_off = out_off;
    s.stop_pos = ~0ull;
    s.h = in_header ? 1u : 0u;
    s.pad = 0;
    *reinterpret_cast<StripState *>(host_state) = s;
}
void strip_state_read(const void *host_state, uint64_t *out_off, uint64_t *stop_pos, int *in_header)
{
    const StripState &s = *reinterpret_cast<const StripState *>(host_state);
    *out_off = s.out_off;
    *stop_pos = s.stop_pos;
    *in_header = (int)s.h;
}

}  // namespace fkb
