// fkb_stream.cuh -- device helpers shared by the streaming count kernels (fkb_bucket.cu, fkb_bucket2.cu, fkb_smallk.cu):
// 128-bit loads, SIMD-in-register ASCII -> 2-bit encode with validity (base2int, findKmer/src/findKmer.cpp:567-589), run
// masks, and the warp-cooperative handler of the rare per-run events of the reference's scan (seqSize == k, :1044-1057;
// seqSize < k, :1059-1062).  Internal; everything is `static`/inline so that each translation unit keeps its own copy.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "findkmer_b200.h"

namespace fkb {
namespace {

// sum_{e<d} 4^e = (4^d - 4) / 3; (4^d - 1) / 3 is the bit pattern 0101..01 (d ones): no division on the event path
__host__ __device__ inline uint64_t flags_offset(int d) { return (0x5555555555555555ull & ((1ull << (2 * d)) - 1ull)) - 1ull; }

__device__ __forceinline__ uint4 ldg128(const uint8_t *p)
{
    uint4 r;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
// predicated 128-bit load (zeros when off): no branch, so the loads stay inside the basic block of the staging code
__device__ __forceinline__ uint4 ldg128_if(const uint8_t *p, bool pred)
{
    uint4 r = make_uint4(0, 0, 0, 0);
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];\n\t}"
                 : "+r"(r.x), "+r"(r.y), "+r"(r.z), "+r"(r.w) : "l"(p), "r"((uint32_t)pred));
    return r;
}
// predicated 128-bit load that KEEPS the register's previous content when off (no zero fill: two CS2R less per load) -- for the
// pipeline registers of the streaming kernels, whose values beyond the end of a warp's region are never consumed
__device__ __forceinline__ void ldg128_keep(uint4 &r, const uint8_t *p, bool pred)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];\n\t}"
                 : "+r"(r.x), "+r"(r.y), "+r"(r.z), "+r"(r.w) : "l"(p), "r"((uint32_t)pred));
}
// predicated 128-bit shared load / global store (pass-1 flush): no branches, loads and stores in separate statements so
// that a step's loads are all in flight before the first store waits for its data
__device__ __forceinline__ uint4 lds128_if(uint32_t saddr, bool pred)
{
    uint4 r = make_uint4(0, 0, 0, 0);
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];\n\t}"
                 : "+r"(r.x), "+r"(r.y), "+r"(r.z), "+r"(r.w) : "r"(saddr), "r"((uint32_t)pred));
    return r;
}
__device__ __forceinline__ void stg128_if(void *gptr, const uint4 &v, bool pred)
{
    asm volatile("{\n\t.reg .pred q;\n\tsetp.ne.u32 q, %5, 0;\n\t@q st.global.v4.u32 [%0], {%1,%2,%3,%4};\n\t}"
                 ::"l"(gptr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w), "r"((uint32_t)pred) : "memory");
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

// logical right shift by a constant (issuing it as IMAD.HI on the FMA pipe was measured: +2 %, not kept)
template <int N>
__device__ __forceinline__ uint32_t shr_c(uint32_t x) { return x >> N; }

// ---- SIMD-in-register encode of 4 ASCII bytes (one 32-bit word, byte 0 = lowest address) ----------------
// codes: 8 bits, byte 0's 2-bit code in bits 7..6 (A0 C1 G2 T3 = base2int, findKmer.cpp:569-576)
// bad  : a word whose byte i is ZERO iff input byte i is one of A,C,G,T.
//   A 0x41, C 0x43, G 0x47, T 0x54: bits 7,6,5,3 must read 0,1,0,0; with q = b2 & ~b1 the rest must satisfy
//   b4 == q and b0 == ~q (A,C,G: q = 0 for A and C, 1 for G ... T: b2 = 1, b1 = 0 -> q = 1, b4 = 1, b0 = 0; G: b2 = b1 = 1 -> q = 0).
__device__ __forceinline__ void encode_word(uint32_t w, uint32_t &codes, uint32_t &bad)
{
    const uint32_t p1 = shr_c<1>(w), p2 = shr_c<2>(w), p4 = shr_c<4>(w);
    const uint32_t t = (p1 ^ p2) & 0x03030303u;
    codes = shr_c<24>(t * 0x40100401u);
    const uint32_t q = p2 & ~p1;
    const uint32_t lo_bad = ((p4 ^ q) | ~(w ^ q)) & 0x01010101u;
    bad = ((w & 0xE8E8E8E8u) ^ 0x40404040u) | lo_bad;
}
// 4-bit mask (byte 0 -> bit 3) of the ZERO bytes of u
__device__ __forceinline__ uint32_t zero_bytes_nibble(uint32_t u)
{
    uint32_t nz = (((u & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | u) & 0x80808080u;  // bit 7 of every non-zero byte
    uint32_t f = (nz ^ 0x80808080u) >> 7;                                  // bit 0 of every zero byte
    return (f * 0x08040201u) >> 24;                                        // byte0->bit3 ... byte3->bit0
}
__device__ __forceinline__ uint32_t count_n_or_gt(uint32_t w)
{
    uint32_t u = w ^ 0x4E4E4E4Eu;            // 'N' -> 0x00, '>' -> 0x70
    u ^= (u & 0x10101010u) * 7u;             // 0x70 -> 0x00; no other byte value becomes zero
    const uint32_t nz = (((u & 0x7F7F7F7Fu) + 0x7F7F7F7Fu) | u) & 0x80808080u;
    return 4u - __popc(nz);
}

struct Group {
    uint32_t code;   // 16 bases x 2 bits, first byte in bits 31..30
    uint32_t valid;  // 16 bits, first byte in bit 15
};

// pack 16 bytes; `unknown` accumulates the bytes outside {A,C,G,T,N,'>'} (one stderr line each in the reference, :581-585)
__device__ __forceinline__ Group pack_group(const uint4 &g, uint32_t &unknown)
{
    Group r;
    uint32_t c0, c1, c2, c3, b0, b1, b2, b3;
    encode_word(g.x, c0, b0);
    encode_word(g.y, c1, b1);
    encode_word(g.z, c2, b2);
    encode_word(g.w, c3, b3);
    r.code = (c0 << 24) | (c1 << 16) | (c2 << 8) | c3;
    r.valid = 0xFFFFu;
    if ((b0 | b1 | b2 | b3) != 0) {  // rare in sequence data: some byte is not a base
        r.valid = (zero_bytes_nibble(b0) << 12) | (zero_bytes_nibble(b1) << 8) | (zero_bytes_nibble(b2) << 4) | zero_bytes_nibble(b3);
        // bytes that are 'N' (0x4E) or '>' (0x3E) reset silently; both map to a zero byte under u -> u ^ 7*(u & 0x10), u = w ^ 'N'
        uint32_t nN = count_n_or_gt(g.x) + count_n_or_gt(g.y) + count_n_or_gt(g.z) + count_n_or_gt(g.w), nH = 0;
        unknown += 16 - __popc(r.valid) - nN - nH;
    }
    return r;
}

// exact validity mask of 16 bytes (first byte = bit 15) WITHOUT the codes (the fast encode below already has them for every byte
// that is a base); `unknown` accumulates the bytes outside {A,C,G,T,N,'>'} as pack_group does
__device__ __forceinline__ uint32_t bad_bytes_word(uint32_t w)  // byte i is ZERO iff input byte i is one of A,C,G,T (see encode_word)
{
    const uint32_t p1 = shr_c<1>(w), p2 = shr_c<2>(w), p4 = shr_c<4>(w);
    const uint32_t q = p2 & ~p1;
    return ((w & 0xE8E8E8E8u) ^ 0x40404040u) | (((p4 ^ q) | ~(w ^ q)) & 0x01010101u);
}
__device__ __forceinline__ uint32_t exact_valid16(const uint4 &g, uint32_t &unknown)
{
    const uint32_t valid = (zero_bytes_nibble(bad_bytes_word(g.x)) << 12) | (zero_bytes_nibble(bad_bytes_word(g.y)) << 8) |
                           (zero_bytes_nibble(bad_bytes_word(g.z)) << 4) | zero_bytes_nibble(bad_bytes_word(g.w));
    const uint32_t nN = count_n_or_gt(g.x) + count_n_or_gt(g.y) + count_n_or_gt(g.z) + count_n_or_gt(g.w);
    unknown += 16 - __popc(valid) - nN;
    return valid;
}

// ---- fast encode: codes only; validity is ACCUMULATED over many words and tested once ----------------
// A byte is one of A,C,G,T  <=>  b7 b6 b5 b3 = 0 1 0 0,  b4 == q,  b0 == ~q  with q = b2 & ~b1 (see encode_word).
// The two relations are evaluated at bit 4 of every byte from LEFT-shifted copies of the word (left shifts are IMAD.SHL on
// the otherwise idle FMA pipe, and shifting up never drags a neighbour byte's bits into bit 4); the four constant bits go
// through an AND and an OR accumulator.  Per word: 1 shift + 1 mask + 1 multiply-gather for the codes, 3 IMAD.SHL + 4
// LOP3 for the validity.
struct ValidAcc {
    uint32_t rel = 0;              // bit 4 of a byte set: b4 != q or b0 == q somewhere
    uint32_t all = 0xFFFFFFFFu;    // AND of the words
    uint32_t any = 0;              // OR of the words
    __device__ __forceinline__ bool bad() const { return ((rel & 0x10101010u) | (any & 0xA8A8A8A8u) | (~all & 0x40404040u)) != 0; }
};
__device__ __forceinline__ uint32_t raw_index_gather(uint32_t w)  // top byte: (b2 b1) of byte 0 in bits 31..30, ... byte 3 in 25..24
{
    return ((w >> 1) & 0x03030303u) * 0x40100401u;
}
__device__ __forceinline__ void valid_acc_word(uint32_t w, ValidAcc &a)
{
    const uint32_t l2 = w << 2, l3 = w << 3, l4 = w << 4;
    const uint32_t q = l2 & ~l3;
    a.rel |= (w ^ q) | ~(l4 ^ q);
}
__device__ __forceinline__ uint32_t pack_codes_fast(const uint4 &g, ValidAcc &a)
{
    const uint32_t p0 = raw_index_gather(g.x), p1 = raw_index_gather(g.y), p2 = raw_index_gather(g.z), p3 = raw_index_gather(g.w);
    valid_acc_word(g.x, a);
    valid_acc_word(g.y, a);
    valid_acc_word(g.z, a);
    valid_acc_word(g.w, a);
    a.all &= g.x & g.y;
    a.all &= g.z & g.w;
    a.any |= g.x | g.y;
    a.any |= g.z | g.w;
    const uint32_t t01 = __byte_perm(p1, p0, 0x0073), t23 = __byte_perm(p3, p2, 0x0073);  // byte 0 = later word's top byte, byte 1 = earlier word's
    const uint32_t raw = __byte_perm(t23, t01, 0x5410);                                    // 16 raw indices: A0 C1 T2 G3
    return raw ^ ((raw >> 1) & 0x55555555u);                                               // -> A0 C1 G2 T3
}

// bits b of m (earlier bytes in higher bits) such that bits b .. b+LEN-1 are all set, by doubling
template <int LEN>
__device__ __forceinline__ uint32_t runs_of(uint32_t m)
{
    if constexpr (LEN == 1) {
        return m;
    } else {
        constexpr int H = LEN / 2;
        const uint32_t h = runs_of<H>(m);
        uint32_t r = h & (h >> H);
        if constexpr (LEN & 1) r &= (m >> (LEN - 1));
        return r;
    }
}

// A short run's prefix flag.  Flags only ever go from 0 to 1, and EVERY run start writes the same few bytes (depth 1..3 are 84
// bytes: one line): plain stores from 148 SMs onto one L2 line serialise -- config 5 ran at 0.86 Tbases/s with them and at 1.70
// without (profiles/r02_flag_stores.txt).  So look first: the load is served by the SM's L1 after the first miss; a stale 0 only
// costs one redundant store (which drops the stale line), never a wrong flag.
__device__ __forceinline__ void set_flag(uint8_t *flag)
{
    uint32_t v;
    asm volatile("ld.global.ca.u8 %0, [%1];" : "=r"(v) : "l"(flag) : "memory");
    if (v == 0) *flag = 1;
}

// rare per-run events of one group (positions given as bits of 16-bit masks, first byte = bit 15).
// ev[0..3] head_base, ev[4..7] short_first, ev[8] runs_ge_k: CTA-private counters in shared memory -- soft-masked /
// N-rich genomes have millions of run boundaries, and global atomics on nine fixed addresses would serialise them.
// left: valid k-mers whose covering W-window is broken (run edges / range edges) -> one global red each on T_k.
// Done by the WHOLE warp: `who` = ballot of the lanes whose group has
// events.  Two source lanes are served per round (half-warp each); inside a half-warp lane j handles bit j of the three
// 16-bit event masks, so a run start with its k-1 short positions costs one round instead of a 10-trip loop executed by one
// lane while 31 wait (soft-masked / N-rich input has a run boundary in most warp iterations).
__device__ __noinline__ void warp_group_events(uint32_t who, uint32_t left, uint32_t first_k, uint32_t shorts, uint32_t m32, uint32_t code_hi,
                                               uint32_t code_lo, int k, uint8_t *flags, uint32_t *ev, uint32_t *table_k)
{
    const uint32_t lane = threadIdx.x & 31, b = lane & 15, half = lane >> 4;
    const uint32_t kmask = (k == 16) ? 0xffffffffu : ((1u << (2 * k)) - 1u);
    while (who) {
        const int s0 = __ffs(who) - 1;
        who &= who - 1;
        const int s1 = who ? __ffs(who) - 1 : -1;
        if (who) who &= who - 1;
        const int src = half ? (s1 < 0 ? s0 : s1) : s0;
        const bool live = !(half && s1 < 0);
        const uint32_t L = __shfl_sync(0xffffffffu, left, src), F = __shfl_sync(0xffffffffu, first_k, src), Sh = __shfl_sync(0xffffffffu, shorts, src);
        const uint32_t M = __shfl_sync(0xffffffffu, m32, src), Chi = __shfl_sync(0xffffffffu, code_hi, src), Clo = __shfl_sync(0xffffffffu, code_lo, src);
        if (!live) continue;
        const uint32_t win = __funnelshift_r(Clo, Chi, 2 * b);  // the 16 bases that end at bit b of the group (first byte = bit 15)
        if ((L >> b) & 1u) red_add_u32(table_k + (win & kmask), 1u);  // a k-mer whose covering W-window is broken: straight to T_k
        if ((F >> b) & 1u) {  // run length reached exactly k here: first k-1 bases of the window go to head_base (:1050-1056)
            const uint32_t head = (win & kmask) >> 2;
            const uint32_t lo = head & 0x55555555u, hi = (head >> 1) & 0x55555555u;
            const uint32_t cT = __popc(lo & hi), cG = __popc(hi & ~lo), cC = __popc(lo & ~hi), cA = (uint32_t)(k - 1) - cT - cG - cC;
            if (cA) atomicAdd(&ev[0], cA);
            if (cC) atomicAdd(&ev[1], cC);
            if (cG) atomicAdd(&ev[2], cG);
            if (cT) atomicAdd(&ev[3], cT);
            atomicAdd(&ev[8], 1u);
        }
        if ((Sh >> b) & 1u) {  // valid base whose run is still shorter than k: the reference inserts a short path (:1059-1062)
            const int run = __ffs(~(M >> b)) - 1;  // consecutive valid bytes ending here (1 .. k-1)
            const uint32_t prefix = win & ((1u << (2 * run)) - 1u);
            set_flag(flags + flags_offset(run) + prefix);
            atomicAdd(&ev[4 + (prefix >> (2 * (run - 1)))], 1u);
        }
    }
}

}  // namespace
}  // namespace fkb
