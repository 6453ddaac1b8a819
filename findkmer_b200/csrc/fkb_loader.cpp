// fkb_loader.cpp -- host sequence loader: record and line stripping (the "stream contract").
//
// Restates, as a block-parallel filter, what the reference's scan loop does with three bytes
// (findKmer/src/findKmer.cpp:988-1011):
//   '\n'  -> dropped (windows span line breaks, :1011)
//   '>'   -> ONE '>' byte is emitted (it resets the window, :994) and everything through the next '\n'
//            is dropped (:999/:1005)
//   0xFF  -> outside a header it ends the scan (fgetc() stored in a `char` aliases EOF, :975,:988)
// Every other byte is copied verbatim; the device decides what is a base.
#include "fkb_loader.h"

#include <stdlib.h>
#include <string.h>
#include <thread>

#include <immintrin.h>

namespace fkb {

int default_host_threads()
{
    const char *env = getenv("FKB_HOST_THREADS");
    if (env && atoi(env) > 0) return atoi(env);
    unsigned hc = std::thread::hardware_concurrency();
    if (hc == 0) hc = 4;
    return (int)(hc > 32 ? 32 : hc);
}

bool in_header_at(const uint8_t *buf, size_t pos)
{
    if (pos == 0) return false;
    const uint8_t *nl = (const uint8_t *)memrchr(buf, '\n', pos);
    size_t line_start = nl ? (size_t)(nl - buf) + 1 : 0;
    return memchr(buf + line_start, '>', pos - line_start) != nullptr;
}

// Header state at every block start first + i * block_bytes (i < n_blocks) of buf[first, len).
// One look-back for block 0; every later state follows from a summary of the block in front of it -- does it hold a '\n',
// and is there a '>' after its last '\n' (or anywhere, when it has none) -- combined in order.  Each byte is looked at by at
// most one memrchr and one memchr: an unwrapped record (one line of 100 MB and more) costs one pass over the file here
// instead of one look-back over the whole line per block.
std::vector<uint8_t> header_states(const uint8_t *buf, size_t first, size_t len, size_t block_bytes, int n_threads)
{
    const size_t n_blocks = len > first ? (len - first + block_bytes - 1) / block_bytes : 0;
    std::vector<uint8_t> state(n_blocks ? n_blocks : 1, 0);
    if (!n_blocks) return state;
    std::vector<uint8_t> summary(n_blocks, 0);  // bit 0: the block holds a '\n'; bit 1: a '>' after its last '\n' (or anywhere, when none)
    auto summarise = [&](size_t i) {
        const size_t a = first + i * block_bytes, b = a + block_bytes < len ? a + block_bytes : len;
        const uint8_t *nl = (const uint8_t *)memrchr(buf + a, '\n', b - a);
        const size_t from = nl ? (size_t)(nl - buf) + 1 : a;
        summary[i] = (uint8_t)((nl ? 1 : 0) | (memchr(buf + from, '>', b - from) ? 2 : 0));
    };
    if (n_threads > 1 && n_blocks >= 64) {
        std::vector<std::thread> pool;
        const int nt = (int)((size_t)n_threads < n_blocks / 16 ? (size_t)n_threads : n_blocks / 16);
        for (int t = 0; t < nt; ++t)
            pool.emplace_back([&, t] { for (size_t i = (size_t)t; i + 1 < n_blocks; i += (size_t)nt) summarise(i); });
        for (auto &th : pool) th.join();
    } else {
        for (size_t i = 0; i + 1 < n_blocks; ++i) summarise(i);
    }
    state[0] = in_header_at(buf, first) ? 1 : 0;
    for (size_t i = 0; i + 1 < n_blocks; ++i) state[i + 1] = (summary[i] & 1) ? (summary[i] >> 1) : (uint8_t)(state[i] | (summary[i] >> 1));
    return state;
}

namespace {

// portable scan: index of the first byte in [p,b) that is '\n', '>' or 0xFF
inline size_t next_special_scalar(const uint8_t *buf, size_t p, size_t b)
{
    while (p < b) {
        uint8_t c = buf[p];
        if (c == '\n' || c == '>' || c == 0xFF) break;
        ++p;
    }
    return p;
}

template <bool WRITE>
StripResult strip_scalar(const uint8_t *buf, size_t a, size_t b, bool in_header, uint8_t *out)
{
    size_t o = 0, p = a;
    while (p < b) {
        if (in_header) {
            const uint8_t *nl = (const uint8_t *)memchr(buf + p, '\n', b - p);
            if (!nl) return {o, SIZE_MAX, true};
            p = (size_t)(nl - buf) + 1;
            in_header = false;
            continue;
        }
        size_t q = next_special_scalar(buf, p, b);
        if (WRITE) memcpy(out + o, buf + p, q - p);
        o += q - p;
        if (q == b) break;
        uint8_t c = buf[q];
        p = q + 1;
        if (c == '>') {
            if (WRITE) out[o] = '>';
            ++o;
            in_header = true;
        } else if (c == 0xFF) {
            return {o, q, false};
        }
    }
    return {o, SIZE_MAX, in_header};
}

// AVX2 scan: 32 bytes per step; the common 60-column FASTA line costs two steps.
template <bool WRITE>
__attribute__((target("avx2"))) StripResult strip_avx2(const uint8_t *buf, size_t a, size_t b, bool in_header, uint8_t *out)
{
    const __m256i v_nl = _mm256_set1_epi8('\n'), v_gt = _mm256_set1_epi8('>'), v_ff = _mm256_set1_epi8((char)0xFF);
    size_t o = 0, p = a;
    while (p < b) {
        if (in_header) {
            const uint8_t *nl = (const uint8_t *)memchr(buf + p, '\n', b - p);
            if (!nl) return {o, SIZE_MAX, true};
            p = (size_t)(nl - buf) + 1;
            in_header = false;
            continue;
        }
        if (p + 32 <= b) {
            __m256i v = _mm256_loadu_si256((const __m256i *)(buf + p));
            __m256i sp = _mm256_or_si256(_mm256_or_si256(_mm256_cmpeq_epi8(v, v_nl), _mm256_cmpeq_epi8(v, v_gt)), _mm256_cmpeq_epi8(v, v_ff));
            uint32_t m = (uint32_t)_mm256_movemask_epi8(sp);
            unsigned keep = m ? (unsigned)__builtin_ctz(m) : 32u;
            if (WRITE) {
                // the store may run past the kept bytes; the caller guarantees 32 bytes of slack at `out`,
                // and whatever follows is overwritten by the next step of this same block
                _mm256_storeu_si256((__m256i *)(out + o), v);
            }
            o += keep;
            if (!m) {
                p += 32;
                continue;
            }
            uint8_t c = buf[p + keep];
            size_t q = p + keep;
            p = q + 1;
            if (c == '>') {
                if (WRITE) out[o] = '>';
                ++o;
                in_header = true;
            } else if (c == 0xFF) {
                return {o, q, false};
            }
            continue;
        }
        // tail shorter than one vector
        size_t q = next_special_scalar(buf, p, b);
        if (WRITE) memcpy(out + o, buf + p, q - p);
        o += q - p;
        if (q == b) break;
        uint8_t c = buf[q];
        p = q + 1;
        if (c == '>') {
            if (WRITE) out[o] = '>';
            ++o;
            in_header = true;
        } else if (c == 0xFF) {
            return {o, q, false};
        }
    }
    return {o, SIZE_MAX, in_header};
}

// AVX-512 VBMI2 scan: 64 bytes per step, line breaks squeezed out in registers (vpcompressb), so a step costs the
// same whether the 64 bytes hold zero, one or two '\n'; only '>' and 0xFF leave the vector loop.
template <bool WRITE>
__attribute__((target("avx512f,avx512bw,avx512vbmi2,popcnt,bmi")))
StripResult strip_avx512(const uint8_t *buf, size_t a, size_t b, bool in_header, uint8_t *out)
{
    __m512i v_nl = _mm512_set1_epi8('\n'), v_gt = _mm512_set1_epi8('>'), v_ff = _mm512_set1_epi8((char)0xFF);
    asm("" : "+v"(v_nl), "+v"(v_gt), "+v"(v_ff));  // keep the three constants in registers (gcc -O2 re-broadcasts them per step)
    size_t o = 0, p = a;
    while (p < b) {
        if (in_header) {
            const uint8_t *nl = (const uint8_t *)memchr(buf + p, '\n', b - p);
            if (!nl) return {o, SIZE_MAX, true};
            p = (size_t)(nl - buf) + 1;
            in_header = false;
            continue;
        }
        // two vectors per step while nothing but line breaks shows up (the body of a record)
        while (p + 128 <= b) {
            __m512i v0 = _mm512_loadu_si512((const void *)(buf + p)), v1 = _mm512_loadu_si512((const void *)(buf + p + 64));
            // '>' = 0x3E and 0xFF both have bits 0x3E set: one compare finds a superset (also '?', '~', DEL, some high bytes)
            __m512i miss = _mm512_min_epu8(_mm512_andnot_si512(v0, v_gt), _mm512_andnot_si512(v1, v_gt));  // 0 where all six bits are set
            if (_mm512_testn_epi8_mask(miss, miss)) break;
            uint64_t k0 = _mm512_cmpneq_epi8_mask(v0, v_nl), k1 = _mm512_cmpneq_epi8_mask(v1, v_nl);
            size_t n0 = (size_t)__builtin_popcountll(k0);
            if (WRITE) {
                _mm512_storeu_si512((void *)(out + o), _mm512_maskz_compress_epi8((__mmask64)k0, v0));
                _mm512_storeu_si512((void *)(out + o + n0), _mm512_maskz_compress_epi8((__mmask64)k1, v1));
            }
            o += n0 + (size_t)__builtin_popcountll(k1);
            p += 128;
        }
        if (p + 64 <= b) {
            __m512i v = _mm512_loadu_si512((const void *)(buf + p));
            uint64_t keep = ~_mm512_cmpeq_epi8_mask(v, v_nl);
            uint64_t stop = _mm512_cmpeq_epi8_mask(v, v_gt) | _mm512_cmpeq_epi8_mask(v, v_ff);
            if (__builtin_expect(stop == 0, 1)) {
                // whole-vector store: o + 64 <= (p - a) + 64 <= b - a, and the bytes past the kept ones are
                // overwritten by the next step of this same block
                if (WRITE) _mm512_storeu_si512((void *)(out + o), _mm512_maskz_compress_epi8((__mmask64)keep, v));
                o += (size_t)__builtin_popcountll(keep);
                p += 64;
                continue;
            }
            unsigned t = (unsigned)__builtin_ctzll(stop);
            keep &= (1ull << t) - 1;
            if (WRITE) _mm512_storeu_si512((void *)(out + o), _mm512_maskz_compress_epi8((__mmask64)keep, v));
            o += (size_t)__builtin_popcountll(keep);
            uint8_t c = buf[p + t];
            size_t q = p + t;
            p = q + 1;
            if (c == '>') {
                if (WRITE) out[o] = '>';
                ++o;
                in_header = true;
            } else {
                return {o, q, false};
            }
            continue;
        }
        // tail shorter than one vector
        size_t q = next_special_scalar(buf, p, b);
        if (WRITE) memcpy(out + o, buf + p, q - p);
        o += q - p;
        if (q == b) break;
        uint8_t c = buf[q];
        p = q + 1;
        if (c == '>') {
            if (WRITE) out[o] = '>';
            ++o;
            in_header = true;
        } else if (c == 0xFF) {
            return {o, q, false};
        }
    }
    return {o, SIZE_MAX, in_header};
}

// 0 = scalar, 1 = AVX2, 2 = AVX-512 VBMI2.  FKB_STRIP_ISA=scalar|avx2|avx512 lowers it; it is read per block (4 MiB of
// work), so the tests can walk all three paths inside one process.
int strip_isa()
{
    static const int best = [] {
        int v = 0;
        if (__builtin_cpu_supports("avx2")) v = 1;
        if (v == 1 && __builtin_cpu_supports("avx512bw") && __builtin_cpu_supports("avx512vbmi2")) v = 2;
        return v;
    }();
    if (const char *env = getenv("FKB_STRIP_ISA")) {
        int want = !strcmp(env, "scalar") ? 0 : !strcmp(env, "avx2") ? 1 : best;
        return want < best ? want : best;
    }
    return best;
}

}  // namespace

// `out` must have room for (b - a) + 32 bytes: the vector paths store whole 32- or 64-byte words.
StripResult strip_block(const uint8_t *buf, size_t a, size_t b, bool in_header, uint8_t *out)
{
    switch (strip_isa()) {
    case 2: return strip_avx512<true>(buf, a, b, in_header, out);
    case 1: return strip_avx2<true>(buf, a, b, in_header, out);
    default: return strip_scalar<true>(buf, a, b, in_header, out);
    }
}

StripResult strip_block_count(const uint8_t *buf, size_t a, size_t b, bool in_header)
{
    switch (strip_isa()) {
    case 2: return strip_avx512<false>(buf, a, b, in_header, nullptr);
    case 1: return strip_avx2<false>(buf, a, b, in_header, nullptr);
    default: return strip_scalar<false>(buf, a, b, in_header, nullptr);
    }
}

}  // namespace fkb
