// fkb_ubench_stage.cu -- input staging A/B for the streaming count kernels (profiling aid, not product path).
//
// BASELINE.json's north star asks for "input tiles staged with TMA or cp.async, each choice justified by counters".  The count
// kernels read the stripped stream exactly once, 1 byte per base, with 128-bit ld.global.nc into a register pipeline
// (fkb_stream.cuh: ldg128 / ldg128_if).  This program measures the same streaming read three ways, with the real per-byte
// work of the kernels' clean path behind it (SIMD encode + validity accumulation, fkb_stream.cuh: pack_codes_fast) or with no
// work at all:
//   mode 0  ld.global.nc.v4 into registers, next tile's loads issued before this tile is processed (what the kernels do)
//   mode 1  cp.async.cg.shared.global 16 B (LDGSTS) into a 3-stage shared-memory ring, cp.async.wait_group, ld.shared.v4
//   mode 2  cp.async.bulk.shared::cluster.global (1-D TMA) of a whole 16 KiB tile by ONE thread into the ring, completion on an
//           mbarrier (expect_tx), consumers spin on mbarrier.try_wait.parity, ld.shared.v4
//   mode 3  as mode 0, but with the count kernels' work distribution: one contiguous region per WARP (thousands of concurrent
//           streams) instead of 16 KiB tiles handed out round-robin to the CTAs (one moving window over the stream)
// Output: one JSON line per (mode, work, CTAs per SM): {"bench":"stage","mode":..,"work":..,"ctas_per_sm":..,"gbs":..}
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include "fkb_stream.cuh"

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e = (x);                                                                   \
        if (e != cudaSuccess) {                                                                \
            fprintf(stderr, "%s: %s (%s:%d)\n", #x, cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(2);                                                                           \
        }                                                                                      \
    } while (0)

using namespace fkb;

constexpr int kThreads = 256;
constexpr int kPerThread = 64;                        // bytes per thread per tile (4 x 128 bit)
constexpr uint32_t kTile = kThreads * kPerThread;     // 16 KiB
constexpr int kStages = 3;

template <bool WORK>
__device__ __forceinline__ void consume(const uint4 (&v)[4], uint32_t &acc, ValidAcc &va)
{
    if constexpr (WORK) {
#pragma unroll
        for (int g = 0; g < 4; ++g) acc ^= pack_codes_fast(v[g], va);
    } else {
#pragma unroll
        for (int g = 0; g < 4; ++g) acc ^= v[g].x ^ v[g].y ^ v[g].z ^ v[g].w;
    }
}

template <int MODE, bool WORK>
__global__ void __launch_bounds__(kThreads) stream_kernel(const uint8_t *__restrict__ s, uint64_t n_tiles, uint32_t *__restrict__ out)
{
    extern __shared__ __align__(128) uint8_t ring[];  // kStages tiles (modes 1, 2)
    __shared__ __align__(8) unsigned long long mbar[kStages];
    uint32_t acc = 0;
    ValidAcc va;
    const uint32_t ring_sa = (uint32_t)__cvta_generic_to_shared(ring);
    if constexpr (MODE == 0) {
        uint4 cur[4], nxt[4];
        uint64_t t = blockIdx.x;
        if (t < n_tiles)
#pragma unroll
            for (int g = 0; g < 4; ++g) cur[g] = ldg128(s + t * kTile + threadIdx.x * kPerThread + 16 * g);
        for (; t < n_tiles; t += gridDim.x) {
            const uint64_t tn = t + gridDim.x;
#pragma unroll
            for (int g = 0; g < 4; ++g) nxt[g] = ldg128_if(s + tn * kTile + threadIdx.x * kPerThread + 16 * g, tn < n_tiles);
            consume<WORK>(cur, acc, va);
#pragma unroll
            for (int g = 0; g < 4; ++g) cur[g] = nxt[g];
        }
    } else if constexpr (MODE == 3) {
        // the count kernels' work distribution: every WARP streams its own contiguous region (lane = 64 consecutive bytes, warp
        // iteration = 2 KiB), so thousands of separate streams are open at the same time
        const uint64_t n_wit = n_tiles * (kTile / 2048), n_warps = (uint64_t)gridDim.x * (kThreads / 32);
        const uint64_t gw = (uint64_t)blockIdx.x * (kThreads / 32) + (threadIdx.x >> 5);
        const uint64_t q = n_wit / n_warps, first = gw * q, iters = (gw + 1 == n_warps) ? n_wit - first : q;
        const uint8_t *base = s + first * 2048 + (threadIdx.x & 31) * kPerThread;
        uint4 cur[4], nxt[4];
        if (iters)
#pragma unroll
            for (int g = 0; g < 4; ++g) cur[g] = ldg128(base + 16 * g);
        for (uint64_t it = 0; it < iters; ++it) {
#pragma unroll
            for (int g = 0; g < 4; ++g) nxt[g] = ldg128_if(base + (it + 1) * 2048 + 16 * g, it + 1 < iters);
            consume<WORK>(cur, acc, va);
#pragma unroll
            for (int g = 0; g < 4; ++g) cur[g] = nxt[g];
        }
    } else if constexpr (MODE == 1) {
        auto issue = [&](uint64_t t, int stage) {
            if (t < n_tiles) {
#pragma unroll
                for (int g = 0; g < 4; ++g) {  // consecutive threads copy consecutive 16-byte pieces: coalesced, conflict-free
                    const uint32_t off = (g * kThreads + threadIdx.x) * 16;
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(ring_sa + stage * kTile + off), "l"(s + t * kTile + off) : "memory");
                }
            }
            asm volatile("cp.async.commit_group;" ::: "memory");
        };
        uint64_t t = blockIdx.x;
        for (int st = 0; st < kStages - 1; ++st) issue(t + (uint64_t)st * gridDim.x, st);
        int stage = 0;
        for (; t < n_tiles; t += gridDim.x) {
            issue(t + (uint64_t)(kStages - 1) * gridDim.x, (stage + kStages - 1) % kStages);
            asm volatile("cp.async.wait_group %0;" ::"n"(kStages - 1) : "memory");
            __syncthreads();
            uint4 v[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) v[g] = *reinterpret_cast<const uint4 *>(ring + stage * kTile + threadIdx.x * kPerThread + 16 * g);
            consume<WORK>(v, acc, va);
            __syncthreads();  // the stage is refilled by the next iteration's issue
            stage = (stage + 1) % kStages;
        }
        asm volatile("cp.async.wait_group 0;" ::: "memory");
    } else {
        const uint32_t mbar_sa = (uint32_t)__cvta_generic_to_shared(mbar);
        if (threadIdx.x == 0) {
            for (int i = 0; i < kStages; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(mbar_sa + 8 * i) : "memory");
            asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        }
        __syncthreads();
        auto issue = [&](uint64_t t, int stage) {  // thread 0 only
            if (t < n_tiles) {
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(mbar_sa + 8 * stage), "r"(kTile) : "memory");
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(ring_sa + stage * kTile),
                             "l"(s + t * kTile), "r"(kTile), "r"(mbar_sa + 8 * stage)
                             : "memory");
            }
        };
        uint64_t t = blockIdx.x;
        if (threadIdx.x == 0)
            for (int st = 0; st < kStages - 1; ++st) issue(t + (uint64_t)st * gridDim.x, st);
        int stage = 0;
        uint32_t phase = 0;
        for (; t < n_tiles; t += gridDim.x) {
            if (threadIdx.x == 0) issue(t + (uint64_t)(kStages - 1) * gridDim.x, (stage + kStages - 1) % kStages);
            uint32_t done = 0;
            while (!done)
                asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                             : "=r"(done) : "r"(mbar_sa + 8 * stage), "r"(phase) : "memory");
            uint4 v[4];
#pragma unroll
            for (int g = 0; g < 4; ++g) v[g] = *reinterpret_cast<const uint4 *>(ring + stage * kTile + threadIdx.x * kPerThread + 16 * g);
            consume<WORK>(v, acc, va);
            __syncthreads();  // every thread has read the stage before thread 0 refills it
            if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
    }
    acc ^= va.rel ^ va.all ^ va.any;
    if (acc == 0x12345678u) out[blockIdx.x * kThreads + threadIdx.x] = acc;  // keeps the work alive; practically never taken
}

template <int MODE, bool WORK>
static void run(const uint8_t *d, uint64_t n_bytes, uint32_t *out, int sm_count, int ctas_per_sm)
{
    const uint64_t n_tiles = n_bytes / kTile;
    const int smem = (MODE == 0 || MODE == 3) ? 0 : kStages * (int)kTile;
    if (smem > 32 * 1024) CK(cudaFuncSetAttribute(stream_kernel<MODE, WORK>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    float best = 1e30f;
    for (int it = 0; it < 6; ++it) {
        CK(cudaEventRecord(e0));
        stream_kernel<MODE, WORK><<<sm_count * ctas_per_sm, kThreads, smem>>>(d, n_tiles, out);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        CK(cudaGetLastError());
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (it >= 2 && ms < best) best = ms;
    }
    printf("{\"bench\":\"stage\",\"mode\":%d,\"work\":%d,\"ctas_per_sm\":%d,\"bytes\":%llu,\"ms\":%.4f,\"gbs\":%.1f}\n", MODE, (int)WORK, ctas_per_sm,
           (unsigned long long)(n_tiles * kTile), best, n_tiles * (double)kTile / best / 1e6);
    fflush(stdout);
}

int main(int argc, char **argv)
{
    const uint64_t n_bytes = argc > 1 ? strtoull(argv[1], nullptr, 10) : (3ull << 30);
    const int only_mode = argc > 2 ? atoi(argv[2]) : -1, only_ctas = argc > 3 ? atoi(argv[3]) : 0;
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    uint8_t *d;
    uint32_t *out;
    CK(cudaMalloc(&d, n_bytes + 4096));
    CK(cudaMalloc(&out, (size_t)prop.multiProcessorCount * 8 * kThreads * 4));
    CK(cudaMemset(d, 'A', n_bytes + 4096));
    for (int ctas : {1, 2, 4, 8}) {
        if (only_ctas && ctas != only_ctas) continue;
        if (ctas * kStages * (int)kTile > 200 * 1024) {  // the ring of modes 1 and 2 limits residency: 4 CTAs of 48 KiB per SM
            if (only_mode <= 0) { run<0, true>(d, n_bytes, out, prop.multiProcessorCount, ctas); run<0, false>(d, n_bytes, out, prop.multiProcessorCount, ctas); }
            if (only_mode < 0 || only_mode == 3) { run<3, true>(d, n_bytes, out, prop.multiProcessorCount, ctas); run<3, false>(d, n_bytes, out, prop.multiProcessorCount, ctas); }
            continue;
        }
        if (only_mode < 0 || only_mode == 0) { run<0, true>(d, n_bytes, out, prop.multiProcessorCount, ctas); run<0, false>(d, n_bytes, out, prop.multiProcessorCount, ctas); }
        if (only_mode < 0 || only_mode == 1) { run<1, true>(d, n_bytes, out, prop.multiProcessorCount, ctas); run<1, false>(d, n_bytes, out, prop.multiProcessorCount, ctas); }
        if (only_mode < 0 || only_mode == 2) { run<2, true>(d, n_bytes, out, prop.multiProcessorCount, ctas); run<2, false>(d, n_bytes, out, prop.multiProcessorCount, ctas); }
        if (only_mode < 0 || only_mode == 3) { run<3, true>(d, n_bytes, out, prop.multiProcessorCount, ctas); run<3, false>(d, n_bytes, out, prop.multiProcessorCount, ctas); }
    }
    return 0;
}
