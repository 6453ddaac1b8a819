// fkb_ubench_bulk.cu -- how fast can one CTA push many SMALL rows from shared memory to global memory?
// (profiling aid, not product path).  Pass 1 of the bucketed count path flushes 1024 staging rows of ~128 bytes per tile;
// this measures that flush done (a) by the LSU (ld.shared.v4 + st.global.v4, 8 lanes per row) and (b) by the bulk-copy
// engine (cp.async.bulk.global.shared::cta, one instruction per row), with every SM busy.
// Output: one JSON line per measurement: {"bench":..., "row_bytes":..., "us_per_tile":..., "clk_per_row":...}
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#define CK(x)                                                                                  \
    do {                                                                                       \
        cudaError_t e = (x);                                                                   \
        if (e != cudaSuccess) {                                                                \
            fprintf(stderr, "%s: %s (%s:%d)\n", #x, cudaGetErrorString(e), __FILE__, __LINE__); \
            exit(2);                                                                           \
        }                                                                                      \
    } while (0)

constexpr int kRows = 1024, kRowStride = 208, kThreads = 512;

// dst layout: [row][cta][slot ring]: like the bucket segments; `ring` tiles deep so the footprint stays L2-sized
__global__ void __launch_bounds__(kThreads, 1) k_flush_lsu(uint8_t *dst, int row_bytes, int tiles, int ring)
{
    extern __shared__ __align__(16) uint8_t sm[];
    for (int i = threadIdx.x; i < kRows * kRowStride / 4; i += kThreads) reinterpret_cast<uint32_t *>(sm)[i] = i;
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane >> 3, c = lane & 7;
    const size_t seg = (size_t)ring * row_bytes;
    for (int t = 0; t < tiles; ++t) {
        for (int b = warp * 4 + sub; b < kRows; b += (kThreads / 32) * 4) {
            uint8_t *d = dst + ((size_t)b * gridDim.x + blockIdx.x) * seg + (size_t)(t % ring) * row_bytes;
            for (int i0 = c * 16; i0 < row_bytes; i0 += 128)
                *reinterpret_cast<uint4 *>(d + i0) = *reinterpret_cast<const uint4 *>(sm + b * kRowStride + i0);
        }
        __syncthreads();
    }
}

__global__ void __launch_bounds__(kThreads, 1) k_flush_bulk(uint8_t *dst, int row_bytes, int tiles, int ring)
{
    extern __shared__ __align__(16) uint8_t sm[];
    for (int i = threadIdx.x; i < kRows * kRowStride / 4; i += kThreads) reinterpret_cast<uint32_t *>(sm)[i] = i;
    __syncthreads();
    const uint32_t sm_sa = (uint32_t)__cvta_generic_to_shared(sm);
    const size_t seg = (size_t)ring * row_bytes;
    for (int t = 0; t < tiles; ++t) {
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int b = threadIdx.x; b < kRows; b += kThreads) {
            uint8_t *d = dst + ((size_t)b * gridDim.x + blockIdx.x) * seg + (size_t)(t % ring) * row_bytes;
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d), "r"(sm_sa + b * kRowStride), "r"(row_bytes) : "memory");
        }
        asm volatile("cp.async.bulk.commit_group;" ::: "memory");
        asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        __syncthreads();
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

// half of the rows by the LSU, half by the bulk-copy engine, at the same time: do the two paths add up?
__global__ void __launch_bounds__(kThreads, 1) k_flush_mixed(uint8_t *dst, int row_bytes, int tiles, int ring)
{
    extern __shared__ __align__(16) uint8_t sm[];
    for (int i = threadIdx.x; i < kRows * kRowStride / 4; i += kThreads) reinterpret_cast<uint32_t *>(sm)[i] = i;
    __syncthreads();
    const uint32_t sm_sa = (uint32_t)__cvta_generic_to_shared(sm);
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, sub = lane >> 3, c = lane & 7;
    const size_t seg = (size_t)ring * row_bytes;
    for (int t = 0; t < tiles; ++t) {
        if (warp < kThreads / 64) {  // rows [0, kRows/2): LSU copies
            for (int b = warp * 4 + sub; b < kRows / 2; b += (kThreads / 64) * 4) {
                uint8_t *d = dst + ((size_t)b * gridDim.x + blockIdx.x) * seg + (size_t)(t % ring) * row_bytes;
                for (int i0 = c * 16; i0 < row_bytes; i0 += 128)
                    *reinterpret_cast<uint4 *>(d + i0) = *reinterpret_cast<const uint4 *>(sm + b * kRowStride + i0);
            }
        } else {                     // rows [kRows/2, kRows): bulk copies
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            for (int b = kRows / 2 + (threadIdx.x - kThreads / 2); b < kRows; b += kThreads / 2) {
                uint8_t *d = dst + ((size_t)b * gridDim.x + blockIdx.x) * seg + (size_t)(t % ring) * row_bytes;
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(d), "r"(sm_sa + b * kRowStride), "r"(row_bytes) : "memory");
            }
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        }
        __syncthreads();
    }
    asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
}

int main()
{
    CK(cudaSetDevice(0));
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    const int smem = kRows * kRowStride;
    CK(cudaFuncSetAttribute(k_flush_lsu, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_flush_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    CK(cudaFuncSetAttribute(k_flush_mixed, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0));
    CK(cudaEventCreate(&e1));
    const int tiles = 200;
    const int rings[] = {1, 64};  // 1: every tile rewrites the same 19-30 MB (stays in L2); 64: streams to HBM like the real kernel
    const int grids[] = {sms, (sms + 7) / 8};  // every SM flushing at once vs one SM in eight (is the limit per SM or chip-wide?)
    for (int gi = 0; gi < 2; ++gi)
    for (int ri = 0; ri < 2; ++ri) {
        const int ring = rings[ri], grid = grids[gi];
        uint8_t *dst;
        CK(cudaMalloc(&dst, (size_t)kRows * sms * ring * 208));
        for (int row_bytes = 64; row_bytes <= 208; row_bytes += (row_bytes == 128 ? 80 : 64)) {
            for (int mode = 0; mode < 3; ++mode) {
                for (int rep = 0; rep < 2; ++rep) {
                    CK(cudaEventRecord(e0));
                    if (mode == 0) k_flush_lsu<<<grid, kThreads, smem>>>(dst, row_bytes, tiles, ring);
                    else if (mode == 1) k_flush_bulk<<<grid, kThreads, smem>>>(dst, row_bytes, tiles, ring);
                    else k_flush_mixed<<<grid, kThreads, smem>>>(dst, row_bytes, tiles, ring);
                    CK(cudaEventRecord(e1));
                    CK(cudaEventSynchronize(e1));
                    CK(cudaGetLastError());
                }
                float ms;
                CK(cudaEventElapsedTime(&ms, e0, e1));
                const double us_tile = ms * 1e3 / tiles;
                printf("{\"bench\": \"%s\", \"ctas\": %d, \"ring\": %d, \"row_bytes\": %d, \"us_per_tile\": %.3f, \"clk_per_row\": %.2f, \"chip_gbs\": %.0f}\n",
                       mode == 0 ? "flush_lsu" : (mode == 1 ? "flush_bulk" : "flush_half_lsu_half_bulk"), grid, ring, row_bytes, us_tile, us_tile * 1e-6 * prop.clockRate * 1e3 / kRows,
                       (double)grid * kRows * row_bytes / us_tile * 1e-3);
            }
        }
        CK(cudaFree(dst));
    }
    return 0;
}
