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
    VARIANT_SMEM = 2,     // fused encode + CTA-private shared-memory table, merged at the end (small k)
};

struct LaunchInfo {
    int sm_count;
    int variant;          // forced variant (VARIANT_AUTO = choose)
};

cudaError_t launch_count(const LaunchInfo &li, const uint8_t *d_stream, uint64_t begin, uint64_t end, int k,
                         uint32_t *d_table, uint8_t *d_flags, fkb_partials *d_partials, cudaStream_t st,
                         int *launches);

cudaError_t launch_finalize(const LaunchInfo &li, int k, const uint32_t *d_table, uint8_t *d_flags,
                            const fkb_partials *d_partials, uint64_t stream_bytes, fkb_counts *d_counts,
                            unsigned long long *d_scratch, cudaStream_t st, int *launches);

cudaError_t launch_synth(const LaunchInfo &li, uint8_t *d_out, uint64_t first_byte, uint64_t n_bytes, int n_records,
                         const uint64_t *d_rec_offsets, const uint64_t *d_rec_base0, const uint8_t *d_headers,
                         int header_len, int line_width, uint64_t seed, int n_runs, int soft_mask,
                         cudaStream_t st, int *launches);

// number of unsigned long long words launch_finalize needs in d_scratch
constexpr int kFinalizeScratchWords = 16;

}  // namespace fkb
