// fkb_kernels.cuh -- launchers of the sm_100a kernels (internal; the public ABI is include/findkmer_b200.h)
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "findkmer_b200.h"

namespace fkb {

// Which count kernel to run.  AUTO picks by k and stream length (see pick_variant()).
enum CountVariant : int {
    VARIANT_AUTO = 0,
    VARIANT_DIRECT = 1,   // fused encode + one red.global per window into the L2-resident table
    VARIANT_BUCKET = 2,   // W-mers at stride S routed through shared memory to per-bucket smem counters, then folded
    VARIANT_SMEM = 3,     // k <= 8: single pass, the whole 4^k table privatised per CTA in shared memory (fkb_smallk.cu)
    VARIANT_BUCKET16 = 4, // k = 11: 16-mers at stride 6 routed as 32-bit items, counted twice (two 13-mers) in 8-bit counters (fkb_bucket2.cu)
};

// device scratch of VARIANT_BUCKET (owned by the context)
struct BucketScratch {
    uint16_t *gbuf = nullptr;      // [n_buckets][n_cta][cap_cb] routed payloads: one private segment per (bucket, pass-1 CTA)
    uint32_t cap_cb = 0;           // items per segment (multiple of 8): [0, cap_front) flushed chunks, [cap_front, cap_cb) direct appends
    uint32_t cap_front = 0;        // multiple of 8
    int n_cta = 0;                 // pass-1 grid size the layout was built for
    uint32_t *gcount = nullptr;    // [n_buckets][n_cta][2] items in the front / back part of each segment
    uint32_t *work = nullptr;      // pass-2 work counter (word 0) and, from word 16 on, one item total per bucket
    uint16_t *table_w = nullptr;   // [4^W] 16-bit W-mer counts
    uint32_t *fold = nullptr;      // per-level (all, suf) arrays of the fold
    int p1_ring = 0;               // option "p1_ring": k = 11 pass 1 with ring rows and an asynchronous, warp-distributed flush (fkb_bucket2.cu)
    cudaEvent_t phase_ev[4] = {nullptr, nullptr, nullptr, nullptr};  // option "phase_events": before pass 1 / after pass 1 / after pass 2 / after the fold
};

struct LaunchInfo {
    int sm_count;
    int variant;          // forced variant (VARIANT_AUTO = choose)
    BucketScratch bucket; // valid when bucket.gbuf != nullptr
    cudaStream_t edge_stream = nullptr;                      // side stream for the edge slivers of a bucketed range (nullptr: same stream)
    cudaEvent_t edge_fork = nullptr, edge_join = nullptr;
};

int bucket_stride_for(int k);        // S = W - k + 1 when the bucketed path supports k, else 0
uint64_t bucket_unit_bytes(int k);   // interior granularity of the bucketed path
size_t bucket_table_w_bytes();
size_t bucket_fold_bytes();
int bucket_count();
uint64_t bucket_min_bytes(int k);    // interior bytes from which the bucketed path beats the direct kernel (measured crossovers)
bool bucket_folds_in_shared(int k);  // core buckets (k >= 9): pass 2 folds in shared memory, no W-mer table / fold scratch in HBM
int bucket_segments_per_sm();  // pass-1 CTAs per SM (each owns one segment per bucket)
cudaError_t launch_count_bucketed(const LaunchInfo &li, const BucketScratch &bs, const uint8_t *d_stream, uint64_t lo, uint64_t hi, int k,
                                  uint32_t *d_table, uint8_t *d_flags, fkb_partials *d_partials, cudaStream_t st, int *launches);

// k = 11, "double 13-mer" items (fkb_bucket2.cu); uses the same scratch as launch_count_bucketed
uint64_t bucket16_unit_bytes(int k);  // interior granularity, 0 when the path does not support k
cudaError_t launch_count_bucketed16(const LaunchInfo &li, const BucketScratch &bs, const uint8_t *d_stream, uint64_t lo, uint64_t hi, uint32_t *d_table,
                                    uint8_t *d_flags, fkb_partials *d_partials, cudaStream_t st, int *launches);

// single-pass shared-memory path for k <= 8 (fkb_smallk.cu)
uint64_t smallk_unit_bytes(int k);   // interior granularity (one warp iteration), 0 when k > 8
cudaError_t launch_count_smallk(const LaunchInfo &li, const uint8_t *d_stream, uint64_t lo, uint64_t hi, int k, uint32_t *d_table, uint8_t *d_flags,
                                fkb_partials *d_partials, cudaStream_t st, int *launches);

cudaError_t launch_count(const LaunchInfo &li, const uint8_t *d_stream, uint64_t begin, uint64_t end, int k,
                         uint32_t *d_table, uint8_t *d_flags, fkb_partials *d_partials, cudaStream_t st,
                         int *launches);

cudaError_t launch_finalize(const LaunchInfo &li, int k, const uint32_t *d_table, uint8_t *d_flags,
                            const fkb_partials *d_partials, uint64_t stream_bytes, fkb_counts *d_counts,
                            unsigned long long *d_scratch, cudaStream_t st, int *launches);

// dst accumulators += src accumulators (src may be peer memory of another GPU): table sum, flags OR, partials sum
cudaError_t launch_accumulate(const LaunchInfo &li, int k, uint32_t *dst_table, const uint32_t *src_table, uint8_t *dst_flags, const uint8_t *src_flags,
                              fkb_partials *dst_p, const fkb_partials *src_p, cudaStream_t st, int *launches);

cudaError_t launch_synth(const LaunchInfo &li, uint8_t *d_out, uint64_t first_byte, uint64_t n_bytes, int n_records,
                         const uint64_t *d_rec_offsets, const uint64_t *d_rec_base0, const uint8_t *d_headers,
                         int header_len, int line_width, uint64_t seed, int n_runs, int soft_mask,
                         cudaStream_t st, int *launches);

// device-side strip (fkb_strip.cu): the stream contract as a chunked stream compaction
size_t strip_scratch_bytes(uint64_t chunk_bytes);
size_t strip_state_bytes();
void strip_state_init(void *host_state, uint64_t out_off, int in_header);
void strip_state_read(const void *host_state, uint64_t *out_off, uint64_t *stop_pos, int *in_header);
cudaError_t launch_strip_chunk(const uint8_t *d_raw, uint64_t n, uint64_t chunk_pos, void *d_state, void *d_scratch, uint8_t *d_out,
                               cudaStream_t st, int *launches);

// number of unsigned long long words launch_finalize needs in d_scratch
constexpr int kFinalizeScratchWords = 16;

}  // namespace fkb
