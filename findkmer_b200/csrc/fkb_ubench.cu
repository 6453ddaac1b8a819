// fkb_ubench.cu -- scatter/atomic throughput micro-benchmarks on B200 (profiling aid, not product path).
// The k = 11 count kernel is bound by scattered read-modify-write throughput, not by HBM; these loops
// measure the ceilings the count-kernel variants are designed against:
//   red_global  : red.global.add.u32 on uniformly random bins of a 2^b-entry table (L2 resident up to ~2^24)
//   red_global64: red.global.add.u64 (two packed 32-bit counters per operation)
//   atoms       : shared-memory atomicAdd (no return) on random bins of a CTA-private table
//   atoms_ret   : same, using the returned value
//   sts         : plain random 4-byte stores into shared memory (a non-atomic scatter, for partitioning)
// Output: one JSON line per measurement: {"bench":..., "bins_log2":..., "gops":...}
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e = (x);                                                                   \
        if (e != cudaSuccess) {                                                                \
            fprintf(stderr, "%s: %s (%s:%d)\n", #x, cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(2);                                                                           \
        }                                                                                      \
    } while (0)

__device__ __forceinline__ uint32_t mix(uint32_t x)
{
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

__global__ void __launch_bounds__(256) k_red_global(uint32_t *table, int bits, int iters, int mode)
{
    uint32_t x = mix(blockIdx.x * 256u + threadIdx.x + 1u);
    const uint32_t lane = threadIdx.x & 31;
    for (int i = 0; i < iters; ++i) {
        x = x * 1664525u + 1013904223u;
        uint32_t idx = x >> (32 - bits);
        if (mode == 1) {  // warp-local: all lanes of a warp land in one 128-byte line (coalescing upper bound)
            uint32_t wx = __shfl_sync(0xffffffffu, x, 0);
            idx = ((wx >> (32 - bits)) & ~31u) | lane;
        } else if (mode == 2) {  // groups of 8 lanes share a 32-byte sector
            uint32_t gx = __shfl_sync(0xffffffffu, x, lane & ~7u);
            idx = ((gx >> (32 - bits)) & ~7u) | (lane & 7u);
        }
        asm volatile("red.global.add.u32 [%0], %1;" ::"l"(table + idx), "r"(1u) : "memory");
    }
}

__global__ void __launch_bounds__(256) k_red_global64(unsigned long long *table, int bits, int iters)
{
    uint32_t x = mix(blockIdx.x * 256u + threadIdx.x + 1u);
    for (int i = 0; i < iters; ++i) {
        x = x * 1664525u + 1013904223u;
        uint32_t idx = x >> (32 - bits);
        asm volatile("red.global.add.u64 [%0], %1;" ::"l"(table + idx), "l"(0x100000001ull) : "memory");
    }
}

// two independent 32-bit reds per thread iteration issued back to back (ILP on the LSU)
__global__ void __launch_bounds__(256) k_red_global_x2(uint32_t *table, int bits, int iters)
{
    uint32_t x = mix(blockIdx.x * 256u + threadIdx.x + 1u), y = mix(x);
    for (int i = 0; i < iters; i += 2) {
        x = x * 1664525u + 1013904223u;
        y = y * 22695477u + 1u;
        asm volatile("red.global.add.u32 [%0], %1;" ::"l"(table + (x >> (32 - bits))), "r"(1u) : "memory");
        asm volatile("red.global.add.u32 [%0], %1;" ::"l"(table + (y >> (32 - bits))), "r"(1u) : "memory");
    }
}

template <int MODE>  // 0: atomicAdd no return, 1: with return, 2: plain store, 3: conflict-free atomicAdd
__global__ void k_smem(uint32_t *sink, int bits, int iters)
{
    extern __shared__ uint32_t sm[];
    const uint32_t n = 1u << bits;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) sm[i] = 0;
    __syncthreads();
    uint32_t x = mix(blockIdx.x * blockDim.x + threadIdx.x + 1u), acc = 0;
    const uint32_t lane = threadIdx.x & 31;
    for (int i = 0; i < iters; ++i) {
        x = x * 1664525u + 1013904223u;
        uint32_t idx = x >> (32 - bits);
        if (MODE == 0) atomicAdd(&sm[idx], 1u);
        if (MODE == 1) acc += atomicAdd(&sm[idx], 1u);
        if (MODE == 2) sm[idx] = x;
        if (MODE == 3) atomicAdd(&sm[(idx & ~31u) | lane], 1u);
    }
    __syncthreads();
    uint32_t s = acc;
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) s += sm[i];
    if (s == 0xdeadbeefu) sink[0] = s;
}

static float time_ms(cudaEvent_t a, cudaEvent_t b)
{
    float ms;
    CK(cudaEventElapsedTime(&ms, a, b));
    return ms;
}

int main(int argc, char **argv)
{
    int dev = 0;
    CK(cudaSetDevice(dev));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, dev));
    const int sms = prop.multiProcessorCount;
    printf("{\"device\": \"%s\", \"sms\": %d, \"clock_khz\": %d, \"l2_bytes\": %d}\n", prop.name, sms, prop.clockRate, prop.l2CacheSize);
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const bool quick = argc > 1 && !strcmp(argv[1], "--quick");

    // ---- global reds ------------------------------------------------------------------------------
    const size_t max_bins = (size_t)1 << 28;
    uint32_t *table;
    CK(cudaMalloc(&table, max_bins * sizeof(uint32_t)));
    CK(cudaMemset(table, 0, max_bins * sizeof(uint32_t)));
    const int iters = quick ? 256 : 1024;
    const int ctas_list[] = {4, 8};
    for (int ci = 0; ci < 2; ++ci) {
        const int grid = sms * ctas_list[ci];
        const double ops = (double)grid * 256 * iters;
        for (int bits = 12; bits <= 28; bits += 2) {
            for (int mode = 0; mode < 3; ++mode) {
                if (mode && bits != 22 && bits != 24) continue;
                k_red_global<<<grid, 256>>>(table, bits, iters, mode);  // warm-up
                CK(cudaEventRecord(e0));
                k_red_global<<<grid, 256>>>(table, bits, iters, mode);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                printf("{\"bench\": \"red_global_u32\", \"mode\": %d, \"ctas_per_sm\": %d, \"bins_log2\": %d, \"table_mib\": %.3f, \"gops\": %.2f}\n",
                       mode, ctas_list[ci], bits, (double)(4ull << bits) / 1048576.0, ops / time_ms(e0, e1) * 1e-6);
            }
            if (bits <= 26) {
                k_red_global64<<<grid, 256>>>((unsigned long long *)table, bits, iters);
                CK(cudaEventRecord(e0));
                k_red_global64<<<grid, 256>>>((unsigned long long *)table, bits, iters);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                printf("{\"bench\": \"red_global_u64\", \"ctas_per_sm\": %d, \"bins_log2\": %d, \"table_mib\": %.3f, \"gops\": %.2f}\n",
                       ctas_list[ci], bits, (double)(8ull << bits) / 1048576.0, ops / time_ms(e0, e1) * 1e-6);
            }
            k_red_global_x2<<<grid, 256>>>(table, bits, iters);
            CK(cudaEventRecord(e0));
            k_red_global_x2<<<grid, 256>>>(table, bits, iters);
            CK(cudaEventRecord(e1));
            CK(cudaEventSynchronize(e1));
            printf("{\"bench\": \"red_global_u32_x2\", \"ctas_per_sm\": %d, \"bins_log2\": %d, \"gops\": %.2f}\n", ctas_list[ci], bits,
                   ops / time_ms(e0, e1) * 1e-6);
        }
    }
    fflush(stdout);

    // ---- shared memory ----------------------------------------------------------------------------
    uint32_t *sink;
    CK(cudaMalloc(&sink, 64));
    CK(cudaFuncSetAttribute(k_smem<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    CK(cudaFuncSetAttribute(k_smem<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    CK(cudaFuncSetAttribute(k_smem<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    CK(cudaFuncSetAttribute(k_smem<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 224 * 1024));
    const int s_iters = quick ? 1024 : 8192;
    const int threads_list[] = {256, 512, 1024};
    for (int ti = 0; ti < 3; ++ti) {
        const int threads = threads_list[ti];
        for (int bits = 10; bits <= 15; ++bits) {
            const size_t smem = (size_t)4 << bits;
            // CTAs per SM that fit: limited by smem (227 KiB) and 2048 threads
            int per_sm = (int)((227 * 1024) / (smem + 1024));
            if (per_sm > 2048 / threads) per_sm = 2048 / threads;
            if (per_sm < 1) continue;
            const int grid = sms * per_sm;
            const double ops = (double)grid * threads * s_iters;
            for (int mode = 0; mode < 4; ++mode) {
                void (*fn)(uint32_t *, int, int) = mode == 0 ? k_smem<0> : mode == 1 ? k_smem<1> : mode == 2 ? k_smem<2> : k_smem<3>;
                fn<<<grid, threads, smem>>>(sink, bits, s_iters);
                CK(cudaEventRecord(e0));
                fn<<<grid, threads, smem>>>(sink, bits, s_iters);
                CK(cudaEventRecord(e1));
                CK(cudaEventSynchronize(e1));
                CK(cudaGetLastError());
                const char *names[] = {"atoms", "atoms_ret", "sts", "atoms_conflict_free"};
                printf("{\"bench\": \"%s\", \"threads\": %d, \"ctas_per_sm\": %d, \"bins_log2\": %d, \"gops\": %.2f}\n", names[mode], threads,
                       per_sm, bits, ops / time_ms(e0, e1) * 1e-6);
            }
        }
    }
    CK(cudaDeviceSynchronize());
    return 0;
}
