// fkb_api.cu -- the extern "C" boundary of include/findkmer_b200.h: context, device entry points and
// the host-buffer (end-to-end) pipeline.  Replaces the call `headNode = findKmer(headNode, &baseCounter,
// baseStatistics, &TotalNumSequencesN)` of the reference's main() (findKmer/src/findKmer.cpp:1320).
#include <cuda_runtime.h>

#include <errno.h>
#include <fcntl.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <condition_variable>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "findkmer_b200.h"
#include "fkb_kernels.cuh"
#include "fkb_loader.h"

struct fkb_context {
    int device = 0;
    int sm_count = 0;
    int cc_major = 0, cc_minor = 0;
    size_t hbm_bytes = 0;
    std::string err;
    uint64_t launches = 0;
    int forced_variant = 0;

    cudaStream_t s_pipe = nullptr;  // host pipeline: copies and kernels, in order
    cudaStream_t s_aux = nullptr;   // zeroing that may overlap the first copies
    cudaStream_t s_edge = nullptr;  // the edge slivers of a bucketed range run here, next to pass 2
    cudaEvent_t ev_edge_fork = nullptr, ev_edge_join = nullptr;

    // finalize scratch
    unsigned long long *d_scratch = nullptr;
    fkb_counts *d_counts = nullptr;
    fkb_counts *h_counts = nullptr;  // pinned

    // host pipeline buffers (grown on demand, kept for reuse)
    uint8_t *d_stream = nullptr;
    size_t d_stream_cap = 0;
    uint32_t *d_table = nullptr;
    uint8_t *d_flags = nullptr;
    fkb_partials *d_partials = nullptr;
    int table_k = 0;

    fkb::BucketScratch bucket;  // scratch of the bucketed count path (grown on demand)

    // device-side loader (pinned inputs): two raw chunk buffers, tile scratch, the chained strip state
    int loader_mode = 0;  // 0 = auto (pinned input -> device strip, pageable -> host strip), 1 = host strip, 2 = device strip
    size_t loader_chunk = 0;  // raw bytes per device-loader chunk (0 = kRawChunkBytes); tests shrink it to cross chunk edges
    uint8_t *d_raw[2] = {nullptr, nullptr};
    void *d_strip_scratch = nullptr;
    void *d_strip_state = nullptr;
    void *h_strip_state = nullptr;  // pinned
    cudaEvent_t ev_copy[2] = {nullptr, nullptr}, ev_strip[2] = {nullptr, nullptr};

    // pinned staging ring for the loader
    static constexpr int kMaxSlots = 64;
    int n_slots = 24;  // ring size in use (option "loader_slots"); slots are allocated on first need
    uint8_t *slots[kMaxSlots] = {nullptr};
    cudaEvent_t slot_free[kMaxSlots] = {nullptr};
};

namespace {

constexpr size_t kBlockBytes = 4u << 20;  // loader block: 4 MiB of raw file per strip task
constexpr size_t kRawChunkBytes = 128u << 20;  // device loader: raw bytes per H2D copy / strip launch

int fail(fkb_context *ctx, int status, const char *fmt, ...)
{
    if (ctx) {
        char buf[512];
        va_list ap;
        va_start(ap, fmt);
        vsnprintf(buf, sizeof buf, fmt, ap);
        va_end(ap);
        ctx->err = buf;
    }
    return status;
}

#define FKB_CUDA(ctx, expr)                                                                             \
    do {                                                                                                \
        cudaError_t e__ = (expr);                                                                       \
        if (e__ != cudaSuccess)                                                                         \
            return fail((ctx), FKB_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(e__), __FILE__, __LINE__); \
    } while (0)

fkb::LaunchInfo launch_info(const fkb_context *ctx)
{
    return fkb::LaunchInfo{ctx->sm_count, ctx->forced_variant, ctx->bucket, ctx->s_edge, ctx->ev_edge_fork, ctx->ev_edge_join};
}

// Scratch of the bucketed path, sized for a range of `range_bytes` at word stride S: one private segment per
// (bucket, pass-1 CTA) of 4x the average fill (real genomes are skewed; anything beyond escapes exactly through
// global reds), the segment fills and the 16-bit W-mer table.  HBM is plentiful: 3.1 Gbp at k = 11 takes ~8.4 GB.
// FKB_TIMING=1: wall-clock marks of the host path on stderr (profiling aid, see profiles/tools/cli_probe.py)
void phase_mark(const char *what)
{
    static const bool on = getenv("FKB_TIMING") != nullptr;
    static auto last = std::chrono::steady_clock::now();
    if (!on) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[timing]   lib: %-21s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - last).count());
    last = now;
}

int ensure_bucket_scratch(fkb_context *ctx, int k, uint64_t range_bytes)
{
    const int S = fkb::bucket_stride_for(k);
    if (!S || ctx->forced_variant == fkb::VARIANT_DIRECT || ctx->forced_variant == fkb::VARIANT_SMEM) return FKB_OK;
    if (k <= 8 && ctx->forced_variant != fkb::VARIANT_BUCKET) return FKB_OK;  // k <= 8 takes the single-pass shared-memory path (fkb_smallk.cu)
    if (range_bytes < fkb::bucket_min_bytes(k) && ctx->forced_variant != fkb::VARIANT_BUCKET && ctx->forced_variant != fkb::VARIANT_BUCKET16) return FKB_OK;  // the direct kernel will run
    const uint64_t items = range_bytes / S + 1;
    const uint64_t nb = (uint64_t)fkb::bucket_count(), n_cta = (uint64_t)ctx->sm_count;
    const uint64_t n_seg = n_cta * (uint64_t)fkb::bucket_segments_per_sm();
    // front part: 4x the average fill (flushed 16-byte chunks); back part for items whose staging row was full
    // (12.6 GB of the 180 GB in all for a 3.1 Gbp range at k = 11)
    uint64_t cap_front = (4 * (items / (nb * n_seg)) + 64 + 7) & ~7ull;
    if (cap_front > 0x3FFFFFF8ull) cap_front = 0x3FFFFFF8ull;
    // back part = 50 % of the front part (2x the average fill: enough for buckets up to ~3.5x as popular as the average; beyond that
    // the surplus is escaped exactly).  A larger one costs ~1 % on evenly loaded input (regions further apart: TLB reach).
    static const uint64_t back_percent = [] {  // FKB_BACK_PERCENT overrides (tuning knob, read once)
        const char *e = getenv("FKB_BACK_PERCENT");
        return e ? (uint64_t)atoi(e) : 50ull;
    }();
    uint64_t cap = cap_front + ((cap_front * back_percent / 100 + 7) & ~7ull);
    if (!fkb::bucket_folds_in_shared(k)) {  // k <= 8 only: the 16-bit 13-mer table (128 MiB) and the fold levels
        if (!ctx->bucket.table_w) FKB_CUDA(ctx, cudaMalloc(&ctx->bucket.table_w, fkb::bucket_table_w_bytes()));
        if (!ctx->bucket.fold) FKB_CUDA(ctx, cudaMalloc(&ctx->bucket.fold, fkb::bucket_fold_bytes()));
    }
    if (ctx->bucket.gbuf && ctx->bucket.cap_cb >= cap) return FKB_OK;
    if (ctx->bucket.gbuf) cudaFree(ctx->bucket.gbuf);
    ctx->bucket.gbuf = nullptr;
    ctx->bucket.cap_cb = 0;
    ctx->bucket.n_cta = (int)n_cta;
    if (!ctx->bucket.gcount) {
        FKB_CUDA(ctx, cudaMalloc(&ctx->bucket.gcount, 2 * nb * n_seg * sizeof(uint32_t)));
        FKB_CUDA(ctx, cudaMalloc(&ctx->bucket.work, 64 + nb * sizeof(uint32_t)));  // pass-2 work counter, then one item total per bucket
    }
    FKB_CUDA(ctx, cudaMalloc(&ctx->bucket.gbuf, nb * n_seg * cap * sizeof(uint16_t)));
    ctx->bucket.cap_cb = (uint32_t)cap;
    ctx->bucket.cap_front = (uint32_t)cap_front;
    return FKB_OK;
}

int check_k(fkb_context *ctx, int k)
{
    if (k < 1 || k > FKB_MAX_K) return fail(ctx, FKB_ERR_BAD_K, "%d is not a valid value for k (dense engine supports 1..%d)", k, FKB_MAX_K);
    return FKB_OK;
}

int ensure_table(fkb_context *ctx, int k)
{
    if (ctx->table_k == k && ctx->d_table) return FKB_OK;
    if (ctx->d_table) cudaFree(ctx->d_table);
    if (ctx->d_flags) cudaFree(ctx->d_flags);
    ctx->d_table = nullptr;
    ctx->d_flags = nullptr;
    ctx->table_k = 0;
    FKB_CUDA(ctx, cudaMalloc(&ctx->d_table, fkb_table_entries(k) * sizeof(uint32_t)));
    FKB_CUDA(ctx, cudaMalloc(&ctx->d_flags, fkb_prefix_flags_bytes(k)));
    if (!ctx->d_partials) FKB_CUDA(ctx, cudaMalloc(&ctx->d_partials, sizeof(fkb_partials)));
    ctx->table_k = k;
    return FKB_OK;
}

int ensure_stream(fkb_context *ctx, size_t bytes)
{
    bytes = (bytes + 255) & ~(size_t)255;
    if (bytes <= ctx->d_stream_cap) return FKB_OK;
    if (ctx->d_stream) cudaFree(ctx->d_stream);
    ctx->d_stream = nullptr;
    ctx->d_stream_cap = 0;
    FKB_CUDA(ctx, cudaMalloc(&ctx->d_stream, bytes));
    ctx->d_stream_cap = bytes;
    return FKB_OK;
}

// the first `want` slots of the ring (a small file pins only what it uses).  Allocating them lazily from the worker
// threads was tried and is slower: cudaHostAlloc calls next to the running copies stall both.
int ensure_slots(fkb_context *ctx, int want)
{
    for (int i = 0; i < want; ++i) {
        if (ctx->slots[i]) continue;
        FKB_CUDA(ctx, cudaHostAlloc((void **)&ctx->slots[i], kBlockBytes + 64, cudaHostAllocDefault));
        FKB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->slot_free[i], cudaEventDisableTiming));
    }
    return FKB_OK;
}

int status_from_counts(fkb_context *ctx, const fkb_counts &c)
{
    if (c.rollover)
        return fail(ctx, FKB_ERR_COUNTER_ROLLOVER,
                    "COUNTER ROLLOVER DETECTED: a 32-bit k-mer/prefix counter would have wrapped (reference findKmer.cpp:640-648)");
    return FKB_OK;
}

// finalize + copy back table and counts, then drain the pipeline stream
int finish_host(fkb_context *ctx, int k, uint64_t stream_bytes, uint32_t *table, fkb_counts *counts)
{
    int launches = 0;
    FKB_CUDA(ctx, fkb::launch_finalize(launch_info(ctx), k, ctx->d_table, ctx->d_flags, ctx->d_partials, stream_bytes, ctx->d_counts,
                                       ctx->d_scratch, ctx->s_pipe, &launches));
    ctx->launches += launches;
    FKB_CUDA(ctx, cudaMemcpyAsync(table, ctx->d_table, fkb_table_entries(k) * sizeof(uint32_t), cudaMemcpyDeviceToHost, ctx->s_pipe));
    FKB_CUDA(ctx, cudaMemcpyAsync(ctx->h_counts, ctx->d_counts, sizeof(fkb_counts), cudaMemcpyDeviceToHost, ctx->s_pipe));
    FKB_CUDA(ctx, cudaStreamSynchronize(ctx->s_pipe));
    *counts = *ctx->h_counts;
    phase_mark("finalize + D2H + sync");
    return status_from_counts(ctx, *counts);
}

}  // namespace

// left halo of a shard: the last 16 stripped bytes in front of buf[own_offset] (zeros = reset bytes at the start of a file).
// Returns false (and the stop offset) when a byte 0xFF outside a header already ended the scan inside the look-back context.
static bool left_halo(const uint8_t *fasta, size_t own_offset, uint8_t halo[16], uint64_t *stop_offset)
{
    memset(halo, 0, 16);
    if (own_offset == 0) return true;
    size_t look = 256;
    std::vector<uint8_t> tmp;
    for (;;) {
        size_t lb = own_offset > look ? own_offset - look : 0;
        tmp.assign(own_offset - lb + 64, 0);
        fkb::StripResult r = fkb::strip_block(fasta, lb, own_offset, fkb::in_header_at(fasta, lb), tmp.data());
        if (r.stop_pos != SIZE_MAX) {
            *stop_offset = r.stop_pos;
            return false;
        }
        if (r.n_out >= 16 || lb == 0) {
            size_t n = r.n_out < 16 ? r.n_out : 16;
            memcpy(halo + 16 - n, tmp.data() + r.n_out - n, n);
            return true;
        }
        look *= 8;
    }
}

// Device loader over buf[own_offset, len) for PINNED input: the raw bytes cross PCIe in 128 MiB chunks (two device
// buffers, copy engine on its own stream), each chunk is stripped on the GPU (fkb_strip.cu) straight into the device
// stream and the newly completed byte range is counted -- the host only launches, it never touches the data.
static int pipeline_range_device(fkb_context *ctx, const uint8_t *fasta, size_t len, size_t own_offset, int k, uint32_t *d_table,
                                 uint8_t *d_flags, fkb_partials *d_partials, uint64_t *stream_bytes, uint64_t *stop_offset, int *ends_in_header_out)
{
    *stream_bytes = 0;
    *stop_offset = UINT64_MAX;
    *ends_in_header_out = 0;
    if (int s = ensure_stream(ctx, (len - own_offset) + 16 + 64)) return s;
    if (k)  // pieces of up to 96 MiB + one chunk are counted at a time
        if (int s = ensure_bucket_scratch(ctx, k, std::min<uint64_t>(len - own_offset, (96ull << 20) + kRawChunkBytes))) return s;
    if (!ctx->d_raw[0]) {
        for (int i = 0; i < 2; ++i) {
            FKB_CUDA(ctx, cudaMalloc(&ctx->d_raw[i], kRawChunkBytes + 64));
            FKB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_copy[i], cudaEventDisableTiming));
            FKB_CUDA(ctx, cudaEventCreateWithFlags(&ctx->ev_strip[i], cudaEventDisableTiming));
        }
        FKB_CUDA(ctx, cudaMalloc(&ctx->d_strip_scratch, fkb::strip_scratch_bytes(kRawChunkBytes)));
        FKB_CUDA(ctx, cudaMalloc(&ctx->d_strip_state, 64));
        FKB_CUDA(ctx, cudaHostAlloc(&ctx->h_strip_state, 64, cudaHostAllocDefault));
    }
    uint8_t halo[16];
    if (!left_halo(fasta, own_offset, halo, stop_offset)) return FKB_OK;  // the scan ended before this shard begins
    fkb::strip_state_init(ctx->h_strip_state, 16, fkb::in_header_at(fasta, own_offset) ? 1 : 0);
    FKB_CUDA(ctx, cudaMemcpyAsync(ctx->d_stream, halo, 16, cudaMemcpyHostToDevice, ctx->s_pipe));
    FKB_CUDA(ctx, cudaMemcpyAsync(ctx->d_strip_state, ctx->h_strip_state, fkb::strip_state_bytes(), cudaMemcpyHostToDevice, ctx->s_pipe));
    FKB_CUDA(ctx, cudaStreamSynchronize(ctx->s_pipe));  // halo[] is a stack buffer; also orders the state before the copy stream's work

    const size_t span = len - own_offset;
    const size_t chunk = ctx->loader_chunk ? ctx->loader_chunk : kRawChunkBytes;
    const size_t n_chunks = (span + chunk - 1) / chunk;
    auto issue_copy = [&](size_t c) -> cudaError_t {
        const int slot = (int)(c & 1);
        const size_t a = own_offset + c * chunk, n = (a + chunk <= len) ? chunk : len - a;
        cudaError_t e = cudaSuccess;
        if (c >= 2) e = cudaStreamWaitEvent(ctx->s_aux, ctx->ev_strip[slot], 0);  // the buffer's previous chunk has been consumed
        if (e == cudaSuccess) e = cudaMemcpyAsync(ctx->d_raw[slot], fasta + a, n, cudaMemcpyHostToDevice, ctx->s_aux);
        if (e == cudaSuccess) e = cudaEventRecord(ctx->ev_copy[slot], ctx->s_aux);
        return e;
    };
    uint64_t counted = 16, out_off = 16, stop = UINT64_MAX;
    int in_header = 0;
    FKB_CUDA(ctx, issue_copy(0));
    for (size_t c = 0; c < n_chunks; ++c) {
        if (c + 1 < n_chunks) FKB_CUDA(ctx, issue_copy(c + 1));  // keep PCIe busy while this chunk is stripped and counted
        const int slot = (int)(c & 1);
        const size_t a = own_offset + c * chunk, n = (a + chunk <= len) ? chunk : len - a;
        FKB_CUDA(ctx, cudaStreamWaitEvent(ctx->s_pipe, ctx->ev_copy[slot], 0));
        int launches = 0;
        FKB_CUDA(ctx, fkb::launch_strip_chunk(ctx->d_raw[slot], n, a, ctx->d_strip_state, ctx->d_strip_scratch, ctx->d_stream, ctx->s_pipe, &launches));
        FKB_CUDA(ctx, cudaEventRecord(ctx->ev_strip[slot], ctx->s_pipe));
        FKB_CUDA(ctx, cudaMemcpyAsync(ctx->h_strip_state, ctx->d_strip_state, fkb::strip_state_bytes(), cudaMemcpyDeviceToHost, ctx->s_pipe));
        FKB_CUDA(ctx, cudaStreamSynchronize(ctx->s_pipe));
        fkb::strip_state_read(ctx->h_strip_state, &out_off, &stop, &in_header);
        const bool last = (c + 1 == n_chunks) || stop != UINT64_MAX;
        if (k && out_off > counted && (last || out_off - counted >= (96ull << 20))) {  // k == 0: load only
            cudaError_t e = fkb::launch_count(launch_info(ctx), ctx->d_stream, counted, out_off, k, d_table, d_flags, d_partials, ctx->s_pipe, &launches);
            counted = out_off;
            FKB_CUDA(ctx, e);
        }
        ctx->launches += launches;
        if (stop != UINT64_MAX) break;
    }
    FKB_CUDA(ctx, cudaStreamSynchronize(ctx->s_aux));  // a copy issued ahead of a stop must not outlive the call
    *stream_bytes = out_off - 16;
    if (stop != UINT64_MAX) *stop_offset = stop;
    else *ends_in_header_out = in_header;
    return FKB_OK;
}

static bool is_pinned_host(const void *p)
{
    cudaPointerAttributes attr;
    if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return attr.type == cudaMemoryTypeHost;
}

// Loader pipeline over buf[own_offset, len): host threads strip 4 MiB blocks into pinned slots; this thread
// commits the blocks IN ORDER -- H2D copy to the running offset in the device stream, then a count kernel over
// the newly arrived byte range (its left halo is already there) -- and recycles a slot when its copy is done.
// Device stream layout: [16 halo bytes][stripped bytes of the owned range]; the halo is the last 16 stripped
// bytes in front of the shard (zeros = reset bytes at the start of a file).
static int pipeline_range(fkb_context *ctx, const uint8_t *fasta, size_t len, size_t own_offset, int k, uint32_t *d_table, uint8_t *d_flags,
                          fkb_partials *d_partials, uint64_t *stream_bytes, uint64_t *stop_offset, int *ends_in_header_out)
{
    // pinned file image: let the GPU strip it (no host pass); pageable memory has to be touched by the CPU anyway
    if (ctx->loader_mode == 2 || (ctx->loader_mode == 0 && is_pinned_host(fasta)))
        return pipeline_range_device(ctx, fasta, len, own_offset, k, d_table, d_flags, d_partials, stream_bytes, stop_offset, ends_in_header_out);
    *stream_bytes = 0;
    *stop_offset = UINT64_MAX;
    *ends_in_header_out = 0;
    phase_mark("(enter pipeline)");
    if (int s = ensure_stream(ctx, (len - own_offset) + 16 + 64)) return s;
    if (k)  // pieces of 32 MiB (+ one block) are counted at a time
        if (int s = ensure_bucket_scratch(ctx, k, std::min<uint64_t>(len - own_offset, (32ull << 20) + kBlockBytes))) return s;
    phase_mark("device stream alloc");
    const size_t span = len - own_offset;
    const size_t n_blocks = (span + kBlockBytes - 1) / kBlockBytes;
    const int n_slots = (size_t)ctx->n_slots < n_blocks ? ctx->n_slots : (int)n_blocks;
    if (int s = ensure_slots(ctx, n_slots)) return s;
    phase_mark("pinned slots");

    uint8_t halo[16];
    if (!left_halo(fasta, own_offset, halo, stop_offset)) return FKB_OK;  // the scan ended before this shard begins: it owns nothing
    FKB_CUDA(ctx, cudaMemcpyAsync(ctx->d_stream, halo, 16, cudaMemcpyHostToDevice, ctx->s_pipe));
    FKB_CUDA(ctx, cudaStreamSynchronize(ctx->s_pipe));  // halo[] is a stack buffer

    int n_threads = fkb::default_host_threads();
    if (n_threads >= 8 && !getenv("FKB_HOST_THREADS")) --n_threads;  // one core stays with the committing thread
    if ((size_t)n_threads > n_blocks) n_threads = (int)n_blocks;

    const std::vector<uint8_t> hdr_state = fkb::header_states(fasta, own_offset, len, kBlockBytes, n_threads);
    struct Task { fkb::StripResult r; bool done = false; };
    std::vector<Task> tasks(n_blocks);
    std::mutex mu;
    std::condition_variable cv_done, cv_slot;
    std::atomic<size_t> next{0};
    size_t released = 0;  // blocks whose H2D copy has completed, in order: their slots may be reused (guarded by mu)
    bool abort_all = false;

    auto worker = [&] {
        for (;;) {
            size_t i = next.fetch_add(1);
            if (i >= n_blocks) return;
            {   // slot i % n_slots last held block i - n_slots
                std::unique_lock<std::mutex> lk(mu);
                cv_slot.wait(lk, [&] { return abort_all || i < released + (size_t)n_slots; });
                if (abort_all) return;
            }
            size_t a = own_offset + i * kBlockBytes, b = a + kBlockBytes < len ? a + kBlockBytes : len;
            fkb::StripResult r = fkb::strip_block(fasta, a, b, hdr_state[i] != 0, ctx->slots[i % n_slots]);
            {
                std::lock_guard<std::mutex> lk(mu);
                tasks[i].r = r;
                tasks[i].done = true;
            }
            cv_done.notify_all();
        }
    };
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; ++t) pool.emplace_back(worker);

    size_t dev_off = 16, counted = 16;
    bool stopped = false, ends_in_header = false;
    cudaError_t cuda_err = cudaSuccess;
    const size_t count_every = 32u << 20;  // one count kernel per >= 32 MiB of newly arrived stream
    auto release_one = [&] {
        std::lock_guard<std::mutex> lk(mu);
        ++released;
        cv_slot.notify_all();
    };
    for (size_t i = 0; i < n_blocks && cuda_err == cudaSuccess && !stopped; ++i) {
        // wait for block i; while waiting, retire finished copies so the workers get their slots back
        for (;;) {
            size_t rel;
            {
                std::unique_lock<std::mutex> lk(mu);
                if (tasks[i].done) break;
                rel = released;
                if (rel >= i) {  // nothing outstanding: the workers cannot be starved of slots
                    cv_done.wait(lk, [&] { return tasks[i].done; });
                    break;
                }
            }
            cuda_err = cudaEventSynchronize(ctx->slot_free[rel % n_slots]);
            if (cuda_err != cudaSuccess) break;
            release_one();
        }
        if (cuda_err != cudaSuccess) break;
        const fkb::StripResult r = tasks[i].r;
        const int slot = (int)(i % n_slots);
        if (r.n_out) {
            cuda_err = cudaMemcpyAsync(ctx->d_stream + dev_off, ctx->slots[slot], r.n_out, cudaMemcpyHostToDevice, ctx->s_pipe);
            dev_off += r.n_out;
        }
        if (cuda_err == cudaSuccess) cuda_err = cudaEventRecord(ctx->slot_free[slot], ctx->s_pipe);
        ends_in_header = r.ends_in_header;
        if (r.stop_pos != SIZE_MAX) {  // byte 0xFF outside a header: the reference's loop ends here
            stopped = true;
            ends_in_header = false;
            *stop_offset = r.stop_pos;
        }
        const bool last = stopped || i + 1 == n_blocks;
        if (k && cuda_err == cudaSuccess && dev_off > counted && (dev_off - counted >= count_every || last)) {  // k == 0: load only
            int launches = 0;
            cuda_err = fkb::launch_count(launch_info(ctx), ctx->d_stream, counted, dev_off, k, d_table, d_flags, d_partials,
                                         ctx->s_pipe, &launches);
            ctx->launches += launches;
            counted = dev_off;
        }
        while (cuda_err == cudaSuccess) {  // opportunistic retire
            size_t rel;
            {
                std::lock_guard<std::mutex> lk(mu);
                rel = released;
            }
            if (rel > i || cudaEventQuery(ctx->slot_free[rel % n_slots]) != cudaSuccess) break;
            release_one();
        }
    }
    {
        std::lock_guard<std::mutex> lk(mu);
        abort_all = true;
    }
    cv_slot.notify_all();
    for (auto &th : pool) th.join();
    if (cuda_err != cudaSuccess) {
        cudaStreamSynchronize(ctx->s_pipe);
        return fail(ctx, FKB_ERR_CUDA, "loader pipeline: %s", cudaGetErrorString(cuda_err));
    }
    *stream_bytes = dev_off - 16;
    *ends_in_header_out = ends_in_header ? 1 : 0;
    phase_mark("strip + H2D + count issue");
    return FKB_OK;
}

extern "C" {

const char *fkb_version(void) { return "findkmer_b200 0.1 (sm_100a)"; }

const char *fkb_status_string(int status)
{
    switch (status) {
    case FKB_OK: return "ok";
    case FKB_ERR_EMPTY_INPUT: return "Sequence File Is Empty";
    case FKB_ERR_UNTERMINATED_HEADER: return "'>' header line without a terminating newline (the reference never returns on this input)";
    case FKB_ERR_COUNTER_ROLLOVER: return "COUNTER ROLLOVER DETECTED";
    case FKB_ERR_BAD_K: return "not a valid value for k";
    case FKB_ERR_NOMEM: return "host memory allocation failed";
    case FKB_ERR_CUDA: return "CUDA failure";
    case FKB_ERR_BAD_ARG: return "bad argument";
    case FKB_ERR_IO: return "I/O failure";
    case FKB_ERR_ZERO_BASE_PROBABILITY: return "Division overflow detected in statistics.";
    default: return "unknown status";
    }
}

size_t fkb_table_entries(int k) { return (k >= 1 && k <= FKB_MAX_K) ? ((size_t)1 << (2 * k)) : 0; }

size_t fkb_prefix_flags_bytes(int k)
{
    if (k < 1 || k > FKB_MAX_K) return 0;
    size_t n = (((size_t)1 << (2 * k)) - 4) / 3;  // sum_{d=1..k-1} 4^d
    return n < 16 ? 16 : n;
}

int fkb_create(int device, fkb_context **out)
{
    if (!out) return FKB_ERR_BAD_ARG;
    *out = nullptr;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess || n <= 0 || device < 0 || device >= n) return FKB_ERR_CUDA;  // no CPU fallback exists
    fkb_context *ctx = new (std::nothrow) fkb_context();
    if (!ctx) return FKB_ERR_NOMEM;
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        delete ctx;
        return FKB_ERR_CUDA;
    }
    ctx->sm_count = prop.multiProcessorCount;
    ctx->cc_major = prop.major;
    ctx->cc_minor = prop.minor;
    ctx->hbm_bytes = prop.totalGlobalMem;
    if (prop.major != 10) {  // the kernels are built for sm_100a only
        delete ctx;
        return FKB_ERR_CUDA;
    }
    const char *v = getenv("FKB_VARIANT");
    ctx->forced_variant = v ? atoi(v) : 0;
    if (const char *r = getenv("FKB_P1_RING")) ctx->bucket.p1_ring = r[0] == '1';
    const char *l = getenv("FKB_LOADER");
    ctx->loader_mode = l ? (!strcmp(l, "host") ? 1 : (!strcmp(l, "device") ? 2 : 0)) : 0;
    if (const char *ns = getenv("FKB_LOADER_SLOTS"))
        if (atoi(ns) >= 2 && atoi(ns) <= fkb_context::kMaxSlots) ctx->n_slots = atoi(ns);
    bool ok = cudaStreamCreateWithFlags(&ctx->s_pipe, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->s_aux, cudaStreamNonBlocking) == cudaSuccess &&
              cudaStreamCreateWithFlags(&ctx->s_edge, cudaStreamNonBlocking) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_edge_fork, cudaEventDisableTiming) == cudaSuccess &&
              cudaEventCreateWithFlags(&ctx->ev_edge_join, cudaEventDisableTiming) == cudaSuccess &&
              cudaMalloc(&ctx->d_scratch, sizeof(unsigned long long) * fkb::kFinalizeScratchWords) == cudaSuccess &&
              cudaMalloc(&ctx->d_counts, sizeof(fkb_counts)) == cudaSuccess &&
              cudaHostAlloc((void **)&ctx->h_counts, sizeof(fkb_counts), cudaHostAllocDefault) == cudaSuccess;
    if (!ok) {
        fkb_destroy(ctx);
        return FKB_ERR_CUDA;
    }
    *out = ctx;
    return FKB_OK;
}

void fkb_destroy(fkb_context *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (int i = 0; i < fkb_context::kMaxSlots; ++i) {
        if (ctx->slots[i]) cudaFreeHost(ctx->slots[i]);
        if (ctx->slot_free[i]) cudaEventDestroy(ctx->slot_free[i]);
    }
    for (int i = 0; i < 2; ++i) {
        if (ctx->d_raw[i]) cudaFree(ctx->d_raw[i]);
        if (ctx->ev_copy[i]) cudaEventDestroy(ctx->ev_copy[i]);
        if (ctx->ev_strip[i]) cudaEventDestroy(ctx->ev_strip[i]);
    }
    if (ctx->d_strip_scratch) cudaFree(ctx->d_strip_scratch);
    if (ctx->d_strip_state) cudaFree(ctx->d_strip_state);
    if (ctx->h_strip_state) cudaFreeHost(ctx->h_strip_state);
    if (ctx->bucket.gbuf) cudaFree(ctx->bucket.gbuf);
    if (ctx->bucket.gcount) cudaFree(ctx->bucket.gcount);
    if (ctx->bucket.work) cudaFree(ctx->bucket.work);
    for (int i = 0; i < 4; ++i)
        if (ctx->bucket.phase_ev[i]) cudaEventDestroy(ctx->bucket.phase_ev[i]);
    if (ctx->bucket.table_w) cudaFree(ctx->bucket.table_w);
    if (ctx->bucket.fold) cudaFree(ctx->bucket.fold);
    if (ctx->d_stream) cudaFree(ctx->d_stream);
    if (ctx->d_table) cudaFree(ctx->d_table);
    if (ctx->d_flags) cudaFree(ctx->d_flags);
    if (ctx->d_partials) cudaFree(ctx->d_partials);
    if (ctx->d_scratch) cudaFree(ctx->d_scratch);
    if (ctx->d_counts) cudaFree(ctx->d_counts);
    if (ctx->h_counts) cudaFreeHost(ctx->h_counts);
    if (ctx->s_pipe) cudaStreamDestroy(ctx->s_pipe);
    if (ctx->s_aux) cudaStreamDestroy(ctx->s_aux);
    if (ctx->s_edge) cudaStreamDestroy(ctx->s_edge);
    if (ctx->ev_edge_fork) cudaEventDestroy(ctx->ev_edge_fork);
    if (ctx->ev_edge_join) cudaEventDestroy(ctx->ev_edge_join);
    delete ctx;
}

const char *fkb_last_error(const fkb_context *ctx) { return ctx ? ctx->err.c_str() : "no context"; }

uint64_t fkb_launch_count(const fkb_context *ctx) { return ctx ? ctx->launches : 0; }

int fkb_device_info(fkb_context *ctx, int *sm_count, int *cc_major, int *cc_minor, size_t *hbm_bytes)
{
    if (!ctx) return FKB_ERR_BAD_ARG;
    if (sm_count) *sm_count = ctx->sm_count;
    if (cc_major) *cc_major = ctx->cc_major;
    if (cc_minor) *cc_minor = ctx->cc_minor;
    if (hbm_bytes) *hbm_bytes = ctx->hbm_bytes;
    return FKB_OK;
}

int fkb_set_option(fkb_context *ctx, const char *name, long value)
{
    if (!ctx || !name) return FKB_ERR_BAD_ARG;
    if (!strcmp(name, "variant") && value >= 0 && value <= 4) {
        ctx->forced_variant = (int)value;
        return FKB_OK;
    }
    if (!strcmp(name, "loader") && value >= 0 && value <= 2) {
        ctx->loader_mode = (int)value;
        return FKB_OK;
    }
    if (!strcmp(name, "loader_chunk") && value >= 0 && (size_t)value <= kRawChunkBytes) {
        ctx->loader_chunk = (size_t)value;
        return FKB_OK;
    }
    if (!strcmp(name, "loader_slots") && value >= 2 && value <= fkb_context::kMaxSlots) {
        ctx->n_slots = (int)value;
        return FKB_OK;
    }
    if (!strcmp(name, "p1_ring") && (value == 0 || value == 1)) {
        ctx->bucket.p1_ring = (int)value;
        return FKB_OK;
    }
    if (!strcmp(name, "phase_events") && (value == 0 || value == 1)) {
        FKB_CUDA(ctx, cudaSetDevice(ctx->device));
        for (int i = 0; i < 4; ++i) {
            if (value && !ctx->bucket.phase_ev[i]) FKB_CUDA(ctx, cudaEventCreate(&ctx->bucket.phase_ev[i]));
            if (!value && ctx->bucket.phase_ev[i]) {
                cudaEventDestroy(ctx->bucket.phase_ev[i]);
                ctx->bucket.phase_ev[i] = nullptr;
            }
        }
        return FKB_OK;
    }
    return fail(ctx, FKB_ERR_BAD_ARG, "unknown option %s=%ld", name, value);
}

int fkb_phase_times(fkb_context *ctx, double ms[3])
{
    if (!ctx || !ms) return FKB_ERR_BAD_ARG;
    if (!ctx->bucket.phase_ev[0]) return fail(ctx, FKB_ERR_BAD_ARG, "fkb_phase_times: option phase_events is off");
    FKB_CUDA(ctx, cudaEventSynchronize(ctx->bucket.phase_ev[3]));
    for (int i = 0; i < 3; ++i) {
        float t = 0;
        FKB_CUDA(ctx, cudaEventElapsedTime(&t, ctx->bucket.phase_ev[i], ctx->bucket.phase_ev[i + 1]));
        ms[i] = t;
    }
    return FKB_OK;
}

int fkb_alloc_pinned(fkb_context *ctx, size_t bytes, void **ptr)
{
    if (!ctx || !ptr) return FKB_ERR_BAD_ARG;
    FKB_CUDA(ctx, cudaSetDevice(ctx->device));
    FKB_CUDA(ctx, cudaHostAlloc(ptr, bytes ? bytes : 1, cudaHostAllocDefault));
    return FKB_OK;
}

int fkb_free_pinned(fkb_context *ctx, void *ptr)
{
    if (!ctx) return FKB_ERR_BAD_ARG;
    FKB_CUDA(ctx, cudaFreeHost(ptr));
    return FKB_OK;
}

// ---- host loader ---------------------------------------------------------------------------------
int fkb_strip_fasta(const uint8_t *fasta, size_t len, uint8_t *stream, size_t *stream_len, int n_threads)
{
    if (!stream_len || (len && (!fasta || !stream))) return FKB_ERR_BAD_ARG;
    *stream_len = 0;
    if (len == 0) return FKB_ERR_EMPTY_INPUT;
    if (n_threads <= 0) n_threads = fkb::default_host_threads();
    const size_t n_blocks = (len + kBlockBytes - 1) / kBlockBytes;
    if ((size_t)n_threads > n_blocks) n_threads = (int)n_blocks;

    struct Block { fkb::StripResult r; bool hdr; size_t off; };
    std::vector<Block> blocks(n_blocks);
    auto run_pass = [&](auto &&fn) {
        std::atomic<size_t> next{0};
        std::vector<std::thread> pool;
        auto worker = [&] {
            for (;;) {
                size_t i = next.fetch_add(1);
                if (i >= n_blocks) break;
                fn(i);
            }
        };
        for (int t = 1; t < n_threads; ++t) pool.emplace_back(worker);
        worker();
        for (auto &th : pool) th.join();
    };
    // pass A: sizes
    const std::vector<uint8_t> hdr_state = fkb::header_states(fasta, 0, len, kBlockBytes, n_threads);
    run_pass([&](size_t i) {
        size_t a = i * kBlockBytes, b = a + kBlockBytes < len ? a + kBlockBytes : len;
        blocks[i].hdr = hdr_state[i] != 0;
        blocks[i].r = fkb::strip_block_count(fasta, a, b, blocks[i].hdr);
    });
    size_t total = 0, last = n_blocks;  // `last`: block that holds the terminating 0xFF, if any
    for (size_t i = 0; i < n_blocks; ++i) {
        blocks[i].off = total;
        total += blocks[i].r.n_out;
        if (blocks[i].r.stop_pos != SIZE_MAX) { last = i; break; }
    }
    const size_t used = last == n_blocks ? n_blocks : last + 1;
    // pass B: write (private scratch per task keeps the vector stores' slack out of the neighbours' output)
    std::atomic<bool> nomem{false};
    run_pass([&](size_t i) {
        if (i >= used) return;
        size_t a = i * kBlockBytes, b = a + kBlockBytes < len ? a + kBlockBytes : len;
        thread_local std::vector<uint8_t> scratch;
        try { scratch.resize(kBlockBytes + 64); } catch (...) { nomem = true; return; }
        fkb::StripResult r = fkb::strip_block(fasta, a, b, blocks[i].hdr, scratch.data());
        memcpy(stream + blocks[i].off, scratch.data(), r.n_out);
    });
    if (nomem) return FKB_ERR_NOMEM;
    *stream_len = total;
    if (last == n_blocks && blocks[n_blocks - 1].r.ends_in_header) return FKB_ERR_UNTERMINATED_HEADER;
    return FKB_OK;
}

// ---- device entry points -------------------------------------------------------------------------
int fkb_zero_device(fkb_context *ctx, int k, uint32_t *d_table, uint8_t *d_flags, fkb_partials *d_partials, void *cuda_stream)
{
    if (!ctx) return FKB_ERR_BAD_ARG;
    if (int s = check_k(ctx, k)) return s;
    cudaStream_t st = (cudaStream_t)cuda_stream;
    if (d_table) FKB_CUDA(ctx, cudaMemsetAsync(d_table, 0, fkb_table_entries(k) * sizeof(uint32_t), st));
    if (d_flags) FKB_CUDA(ctx, cudaMemsetAsync(d_flags, 0, fkb_prefix_flags_bytes(k), st));
    if (d_partials) FKB_CUDA(ctx, cudaMemsetAsync(d_partials, 0, sizeof(fkb_partials), st));
    return FKB_OK;
}

int fkb_count_stream_device(fkb_context *ctx, const uint8_t *d_stream, uint64_t begin, uint64_t end, int k, uint32_t *d_table,
                            uint8_t *d_flags, fkb_partials *d_partials, void *cuda_stream)
{
    if (!ctx || !d_table || !d_flags || !d_partials) return FKB_ERR_BAD_ARG;
    if (int s = check_k(ctx, k)) return s;
    if (end < begin) return fail(ctx, FKB_ERR_BAD_ARG, "end < begin");
    if (end == begin) return FKB_OK;
    if (!d_stream || ((uintptr_t)d_stream & 15)) return fail(ctx, FKB_ERR_BAD_ARG, "d_stream must be a 16-byte aligned device pointer");
    if (int s = ensure_bucket_scratch(ctx, k, end - begin)) return s;
    int launches = 0;
    FKB_CUDA(ctx, fkb::launch_count(launch_info(ctx), d_stream, begin, end, k, d_table, d_flags, d_partials, (cudaStream_t)cuda_stream, &launches));
    ctx->launches += launches;
    return FKB_OK;
}

int fkb_finalize_device(fkb_context *ctx, int k, const uint32_t *d_table, uint8_t *d_flags, const fkb_partials *d_partials,
                        uint64_t stream_bytes, fkb_counts *d_counts, void *cuda_stream)
{
    if (!ctx || !d_table || !d_flags || !d_partials || !d_counts) return FKB_ERR_BAD_ARG;
    if (int s = check_k(ctx, k)) return s;
    int launches = 0;
    FKB_CUDA(ctx, fkb::launch_finalize(launch_info(ctx), k, d_table, d_flags, d_partials, stream_bytes, d_counts, ctx->d_scratch,
                                       (cudaStream_t)cuda_stream, &launches));
    ctx->launches += launches;
    return FKB_OK;
}

// ---- host-buffer pipelines -----------------------------------------------------------------------
int fkb_count_stream_host(fkb_context *ctx, const uint8_t *stream, size_t len, int k, uint32_t *table, fkb_counts *counts)
{
    if (!ctx || !table || !counts || (len && !stream)) return FKB_ERR_BAD_ARG;
    if (int s = check_k(ctx, k)) return s;
    FKB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int s = ensure_table(ctx, k)) return s;
    if (int s = ensure_stream(ctx, len + 64)) return s;
    if (int s = ensure_bucket_scratch(ctx, k, std::min<uint64_t>(len, 64ull << 20))) return s;
    if (int s = fkb_zero_device(ctx, k, ctx->d_table, ctx->d_flags, ctx->d_partials, ctx->s_pipe)) return s;
    // chunked so the count of chunk i overlaps the copy of chunk i+1 (copy engine vs SMs, one ordered stream
    // is enough: a kernel on [a,b) only needs bytes < b, and the copies are issued in order on a second stream)
    const size_t chunk = 64u << 20;
    cudaEvent_t ev;
    FKB_CUDA(ctx, cudaEventCreateWithFlags(&ev, cudaEventDisableTiming));
    for (size_t a = 0; a < len; a += chunk) {
        size_t b = a + chunk < len ? a + chunk : len;
        cudaError_t e = cudaMemcpyAsync(ctx->d_stream + a, stream + a, b - a, cudaMemcpyHostToDevice, ctx->s_aux);
        if (e == cudaSuccess) e = cudaEventRecord(ev, ctx->s_aux);
        if (e == cudaSuccess) e = cudaStreamWaitEvent(ctx->s_pipe, ev, 0);
        if (e != cudaSuccess) { cudaEventDestroy(ev); FKB_CUDA(ctx, e); }
        int launches = 0;
        e = fkb::launch_count(launch_info(ctx), ctx->d_stream, a, b, k, ctx->d_table, ctx->d_flags, ctx->d_partials, ctx->s_pipe, &launches);
        ctx->launches += launches;
        if (e != cudaSuccess) { cudaEventDestroy(ev); FKB_CUDA(ctx, e); }
    }
    cudaEventDestroy(ev);
    return finish_host(ctx, k, len, table, counts);
}

int fkb_count_fasta_host_range(fkb_context *ctx, const uint8_t *buf, size_t len, size_t own_offset, int k, uint32_t *d_table,
                               uint8_t *d_flags, fkb_partials *d_partials, uint64_t *stream_bytes, uint64_t *stop_offset, int *ends_in_header)
{
    if (!ctx || !buf || !d_table || !d_flags || !d_partials || !stream_bytes || !stop_offset || !ends_in_header || own_offset > len)
        return FKB_ERR_BAD_ARG;
    if (int s = check_k(ctx, k)) return s;
    FKB_CUDA(ctx, cudaSetDevice(ctx->device));
    *stream_bytes = 0;
    *stop_offset = UINT64_MAX;
    *ends_in_header = 0;
    if (own_offset == len) return FKB_OK;
    int s = pipeline_range(ctx, buf, len, own_offset, k, d_table, d_flags, d_partials, stream_bytes, stop_offset, ends_in_header);
    cudaError_t e = cudaStreamSynchronize(ctx->s_pipe);  // results are handed to the caller's stream
    if (s == FKB_OK && e != cudaSuccess) return fail(ctx, FKB_ERR_CUDA, "cudaStreamSynchronize: %s", cudaGetErrorString(e));
    return s;
}

int fkb_count_fasta_host(fkb_context *ctx, const uint8_t *fasta, size_t len, int k, uint32_t *table, fkb_counts *counts)
{
    if (!ctx || !table || !counts || (len && !fasta)) return FKB_ERR_BAD_ARG;
    if (int s = check_k(ctx, k)) return s;
    if (len == 0) return fail(ctx, FKB_ERR_EMPTY_INPUT, "Sequence File Is Empty");
    FKB_CUDA(ctx, cudaSetDevice(ctx->device));
    if (int s = ensure_table(ctx, k)) return s;
    if (int s = fkb_zero_device(ctx, k, ctx->d_table, ctx->d_flags, ctx->d_partials, ctx->s_pipe)) return s;
    uint64_t stream_bytes = 0, stop = 0;
    int ends_in_header = 0;
    if (int s = pipeline_range(ctx, fasta, len, 0, k, ctx->d_table, ctx->d_flags, ctx->d_partials, &stream_bytes, &stop, &ends_in_header)) {
        cudaStreamSynchronize(ctx->s_pipe);
        return s;
    }
    if (ends_in_header) {
        cudaStreamSynchronize(ctx->s_pipe);
        return fail(ctx, FKB_ERR_UNTERMINATED_HEADER, "%s", fkb_status_string(FKB_ERR_UNTERMINATED_HEADER));
    }
    return finish_host(ctx, k, stream_bytes, table, counts);
}

// The launcher's shape (k6thru11fullANDupstream.sh:16-24 runs one process per k over the same file) as ONE call: the file
// crosses PCIe and is stripped once, then every k is counted over the device-resident stream.
int fkb_count_fasta_host_multi(fkb_context *ctx, const uint8_t *fasta, size_t len, const int *ks, int n_k, uint32_t *const *tables,
                               fkb_counts *counts)
{
    if (!ctx || !ks || n_k < 1 || !tables || !counts || (len && !fasta)) return FKB_ERR_BAD_ARG;
    for (int i = 0; i < n_k; ++i) {
        if (int s = check_k(ctx, ks[i])) return s;
        if (!tables[i]) return FKB_ERR_BAD_ARG;
    }
    if (len == 0) return fail(ctx, FKB_ERR_EMPTY_INPUT, "Sequence File Is Empty");
    FKB_CUDA(ctx, cudaSetDevice(ctx->device));
    uint64_t stream_bytes = 0, stop = 0;
    int ends_in_header = 0;
    if (int s = pipeline_range(ctx, fasta, len, 0, /*k=*/0, nullptr, nullptr, nullptr, &stream_bytes, &stop, &ends_in_header)) {  // load only
        cudaStreamSynchronize(ctx->s_pipe);
        return s;
    }
    if (ends_in_header) {
        cudaStreamSynchronize(ctx->s_pipe);
        return fail(ctx, FKB_ERR_UNTERMINATED_HEADER, "%s", fkb_status_string(FKB_ERR_UNTERMINATED_HEADER));
    }
    // size the bucket regions once, for the k that needs the most (the largest k: smallest stride, most items),
    // instead of growing them -- cudaFree + cudaMalloc of up to 12.6 GB -- at every step of an ascending sweep
    {
        int k_max = 0;
        for (int i = 0; i < n_k; ++i) k_max = ks[i] > k_max && fkb::bucket_stride_for(ks[i]) ? ks[i] : k_max;
        if (k_max)
            if (int s = ensure_bucket_scratch(ctx, k_max, stream_bytes)) return s;
    }
    int worst = FKB_OK;
    for (int i = 0; i < n_k; ++i) {
        const int k = ks[i];
        if (int s = ensure_table(ctx, k)) return s;
        if (int s = ensure_bucket_scratch(ctx, k, stream_bytes)) return s;
        if (int s = fkb_zero_device(ctx, k, ctx->d_table, ctx->d_flags, ctx->d_partials, ctx->s_pipe)) return s;
        int launches = 0;
        FKB_CUDA(ctx, fkb::launch_count(launch_info(ctx), ctx->d_stream, 16, 16 + stream_bytes, k, ctx->d_table, ctx->d_flags, ctx->d_partials,
                                        ctx->s_pipe, &launches));
        ctx->launches += launches;
        int s = finish_host(ctx, k, stream_bytes, tables[i], &counts[i]);
        if (s != FKB_OK && s != FKB_ERR_COUNTER_ROLLOVER) return s;
        if (s != FKB_OK) worst = s;
    }
    return worst;
}

static int with_mapped_file(fkb_context *ctx, const char *path, const uint8_t **map_out, size_t *len_out)
{
    int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(ctx, FKB_ERR_IO, "Sequence file failed to open: %s: %s", path, strerror(errno));
    struct stat st;
    if (fstat(fd, &st) != 0) {
        close(fd);
        return fail(ctx, FKB_ERR_IO, "fstat(%s): %s", path, strerror(errno));
    }
    if (st.st_size == 0) {
        close(fd);
        return fail(ctx, FKB_ERR_EMPTY_INPUT, "Sequence File Is Empty");
    }
    void *map = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) return fail(ctx, FKB_ERR_IO, "mmap(%s): %s", path, strerror(errno));
    madvise(map, (size_t)st.st_size, MADV_SEQUENTIAL | MADV_WILLNEED);
    *map_out = (const uint8_t *)map;
    *len_out = (size_t)st.st_size;
    return FKB_OK;
}

int fkb_count_file_multi(fkb_context *ctx, const char *path, const int *ks, int n_k, uint32_t *const *tables, fkb_counts *counts)
{
    if (!ctx || !path) return FKB_ERR_BAD_ARG;
    const uint8_t *map = nullptr;
    size_t len = 0;
    if (int s = with_mapped_file(ctx, path, &map, &len)) return s;
    int s = fkb_count_fasta_host_multi(ctx, map, len, ks, n_k, tables, counts);
    munmap((void *)map, len);
    return s;
}

int fkb_count_file(fkb_context *ctx, const char *path, int k, uint32_t *table, fkb_counts *counts)
{
    if (!ctx || !path) return FKB_ERR_BAD_ARG;
    int fd = open(path, O_RDONLY);
    if (fd < 0) return fail(ctx, FKB_ERR_IO, "Sequence file failed to open: %s: %s", path, strerror(errno));
    struct stat st;
    if (fstat(fd, &st) != 0) {
        close(fd);
        return fail(ctx, FKB_ERR_IO, "fstat(%s): %s", path, strerror(errno));
    }
    if (st.st_size == 0) {
        close(fd);
        return fail(ctx, FKB_ERR_EMPTY_INPUT, "Sequence File Is Empty");
    }
    void *map = mmap(nullptr, (size_t)st.st_size, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (map == MAP_FAILED) return fail(ctx, FKB_ERR_IO, "mmap(%s): %s", path, strerror(errno));
    madvise(map, (size_t)st.st_size, MADV_SEQUENTIAL | MADV_WILLNEED);
    int s = fkb_count_fasta_host(ctx, (const uint8_t *)map, (size_t)st.st_size, k, table, counts);
    munmap(map, (size_t)st.st_size);
    return s;
}

// ---- several GPUs of one box behind the C boundary ------------------------------------------------
// The reference's only parallelism is one OS process per k (k6thru11fullANDupstream.sh:16-24).  Here ONE scan is sharded:
// the file image is cut into n contiguous byte ranges, one host thread + one context per GPU counts its range (a window
// belongs to the shard that owns its LAST byte; the 16-byte left halo comes from the look-back context), and the exchange
// step is one pass on devices[0] that adds every other GPU's table / flags / partials -- read directly from peer memory over
// NVLink when the devices can map each other, through a staging copy otherwise -- before finalize.
int fkb_count_fasta_host_gpus(const int *devices, int n_devices, const uint8_t *fasta, size_t len, int k, uint32_t *table, fkb_counts *counts,
                              char *err, size_t err_len)
{
    auto say = [&](const char *msg) { if (err && err_len) snprintf(err, err_len, "%s", msg); };
    if (!devices || n_devices < 1 || n_devices > 64 || !table || !counts || (len && !fasta)) return FKB_ERR_BAD_ARG;
    if (k < 1 || k > FKB_MAX_K) { say("not a valid value for k"); return FKB_ERR_BAD_K; }
    if (len == 0) { say("Sequence File Is Empty"); return FKB_ERR_EMPTY_INPUT; }
    const int n = n_devices;
    std::vector<fkb_context *> ctx(n, nullptr);
    struct Shard { size_t a = 0, b = 0; uint64_t stream_bytes = 0, stop = UINT64_MAX; int ends_in_header = 0; int status = FKB_OK; };
    std::vector<Shard> sh(n);
    for (int i = 0; i < n; ++i) {  // raw byte ranges; any cut is exact (the shard derives its halo and header state from the look-back)
        sh[i].a = (size_t)((unsigned __int128)len * i / n);
        sh[i].b = (size_t)((unsigned __int128)len * (i + 1) / n);
    }
    auto cleanup = [&] { for (auto *c : ctx) if (c) fkb_destroy(c); };
    {
        std::vector<std::thread> pool;
        for (int i = 0; i < n; ++i)
            pool.emplace_back([&, i] {
                Shard &s = sh[i];
                s.status = fkb_create(devices[i], &ctx[i]);
                if (s.status != FKB_OK) return;
                fkb_context *c = ctx[i];
                if ((s.status = ensure_table(c, k)) != FKB_OK) return;
                if ((s.status = fkb_zero_device(c, k, c->d_table, c->d_flags, c->d_partials, c->s_pipe)) != FKB_OK) return;
                if (s.b > s.a)
                    s.status = fkb_count_fasta_host_range(c, fasta, s.b, s.a, k, c->d_table, c->d_flags, c->d_partials, &s.stream_bytes, &s.stop, &s.ends_in_header);
                cudaStreamSynchronize(c->s_pipe);
            });
        for (auto &t : pool) t.join();
    }
    for (int i = 0; i < n; ++i)
        if (sh[i].status != FKB_OK) {
            say(ctx[i] ? fkb_last_error(ctx[i]) : "fkb_create failed: no usable sm_100 GPU for a shard (no CPU fallback exists)");
            const int st = sh[i].status;
            cleanup();
            return st;
        }
    // a byte 0xFF outside a header ends the reference's scan (:975,:988): shards that begin behind the first one count nothing
    uint64_t stop = UINT64_MAX;
    for (int i = 0; i < n; ++i) stop = sh[i].stop < stop ? sh[i].stop : stop;
    int last_live = n - 1;
    if (stop != UINT64_MAX)
        for (int i = 0; i < n; ++i)
            if (sh[i].a <= stop && stop < sh[i].b) last_live = i;
    if (stop == UINT64_MAX) {
        int tail = n - 1;
        while (tail > 0 && sh[tail].b == sh[tail].a) --tail;
        if (sh[tail].ends_in_header) {
            say(fkb_status_string(FKB_ERR_UNTERMINATED_HEADER));
            cleanup();
            return FKB_ERR_UNTERMINATED_HEADER;
        }
    }
    fkb_context *c0 = ctx[0];
    int status = FKB_OK;
    uint64_t stream_bytes = sh[0].stream_bytes;
    cudaSetDevice(c0->device);
    uint32_t *tmp_table = nullptr;
    uint8_t *tmp_flags = nullptr;
    fkb_partials *tmp_partials = nullptr;
    for (int i = 1; i <= last_live && status == FKB_OK; ++i) {
        fkb_context *ci = ctx[i];
        stream_bytes += sh[i].stream_bytes;
        int can = 0;
        cudaDeviceCanAccessPeer(&can, c0->device, ci->device);
        if (can) {
            cudaError_t e = cudaDeviceEnablePeerAccess(ci->device, 0);
            if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); e = cudaSuccess; }
            if (e != cudaSuccess) { cudaGetLastError(); can = 0; }
        }
        const uint32_t *src_t = ci->d_table;
        const uint8_t *src_f = ci->d_flags;
        const fkb_partials *src_p = ci->d_partials;
        cudaError_t e = cudaSuccess;
        if (!can) {  // no peer mapping: stage the shard's accumulators on devices[0]
            const size_t tb = fkb_table_entries(k) * sizeof(uint32_t), fb = fkb_prefix_flags_bytes(k);
            if (!tmp_table) {
                e = cudaMalloc(&tmp_table, tb);
                if (e == cudaSuccess) e = cudaMalloc(&tmp_flags, fb);
                if (e == cudaSuccess) e = cudaMalloc(&tmp_partials, sizeof(fkb_partials));
            }
            if (e == cudaSuccess) e = cudaMemcpyPeerAsync(tmp_table, c0->device, ci->d_table, ci->device, tb, c0->s_pipe);
            if (e == cudaSuccess) e = cudaMemcpyPeerAsync(tmp_flags, c0->device, ci->d_flags, ci->device, fb, c0->s_pipe);
            if (e == cudaSuccess) e = cudaMemcpyPeerAsync(tmp_partials, c0->device, ci->d_partials, ci->device, sizeof(fkb_partials), c0->s_pipe);
            src_t = tmp_table; src_f = tmp_flags; src_p = tmp_partials;
        }
        int launches = 0;
        if (e == cudaSuccess) e = fkb::launch_accumulate(launch_info(c0), k, c0->d_table, src_t, c0->d_flags, src_f, c0->d_partials, src_p, c0->s_pipe, &launches);
        c0->launches += launches;
        if (e == cudaSuccess && !can) e = cudaStreamSynchronize(c0->s_pipe);  // the staging buffers are reused by the next shard
        if (e != cudaSuccess) {
            say(cudaGetErrorString(e));
            status = FKB_ERR_CUDA;
        }
    }
    if (status == FKB_OK) {
        status = finish_host(c0, k, stream_bytes, table, counts);
        if (status != FKB_OK) say(fkb_last_error(c0));
    }
    if (tmp_table) cudaFree(tmp_table);
    if (tmp_flags) cudaFree(tmp_flags);
    if (tmp_partials) cudaFree(tmp_partials);
    cleanup();
    return status;
}

int fkb_count_file_gpus(const int *devices, int n_devices, const char *path, int k, uint32_t *table, fkb_counts *counts, char *err, size_t err_len)
{
    if (!path) return FKB_ERR_BAD_ARG;
    const uint8_t *map = nullptr;
    size_t len = 0;
    fkb_context tmp;  // only its error string is used
    if (int s = with_mapped_file(&tmp, path, &map, &len)) {
        if (err && err_len) snprintf(err, err_len, "%s", tmp.err.c_str());
        return s;
    }
    int s = fkb_count_fasta_host_gpus(devices, n_devices, map, len, k, table, counts, err, err_len);
    munmap((void *)map, len);
    return s;
}

int fkb_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

// ---- synthetic inputs ----------------------------------------------------------------------------
int fkb_synth_fasta_device(fkb_context *ctx, uint8_t *d_out, uint64_t first_byte, uint64_t n_bytes, int n_records,
                           const uint64_t *rec_offsets, const uint64_t *rec_base0, const uint8_t *headers, int header_len, int line_width,
                           uint64_t seed, int n_runs, int soft_mask, void *cuda_stream)
{
    if (!ctx || !d_out || !rec_offsets || !rec_base0 || !headers || n_records < 1 || header_len < 1) return FKB_ERR_BAD_ARG;
    if ((uintptr_t)d_out & 15) return fail(ctx, FKB_ERR_BAD_ARG, "d_out must be 16-byte aligned");
    cudaStream_t st = (cudaStream_t)cuda_stream;
    uint64_t *d_off = nullptr, *d_b0 = nullptr;
    uint8_t *d_hdr = nullptr;
    size_t nb = sizeof(uint64_t) * (size_t)(n_records + 1), nh = (size_t)n_records * header_len;
    FKB_CUDA(ctx, cudaMalloc(&d_off, nb));
    FKB_CUDA(ctx, cudaMalloc(&d_b0, nb));
    FKB_CUDA(ctx, cudaMalloc(&d_hdr, nh));
    FKB_CUDA(ctx, cudaMemcpyAsync(d_off, rec_offsets, nb, cudaMemcpyHostToDevice, st));
    FKB_CUDA(ctx, cudaMemcpyAsync(d_b0, rec_base0, nb, cudaMemcpyHostToDevice, st));
    FKB_CUDA(ctx, cudaMemcpyAsync(d_hdr, headers, nh, cudaMemcpyHostToDevice, st));
    int launches = 0;
    cudaError_t e = fkb::launch_synth(launch_info(ctx), d_out, first_byte, n_bytes, n_records, d_off, d_b0, d_hdr, header_len, line_width, seed,
                                      n_runs, soft_mask, st, &launches);
    ctx->launches += launches;
    cudaError_t e2 = cudaStreamSynchronize(st);  // the temporaries are freed below
    cudaFree(d_off);
    cudaFree(d_b0);
    cudaFree(d_hdr);
    FKB_CUDA(ctx, e);
    FKB_CUDA(ctx, e2);
    return FKB_OK;
}

}  // extern "C"
