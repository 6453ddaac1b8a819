/*
 * findkmer_b200.h -- C ABI of the B200-native k-mer counting engine.
 *
 * Drop-in for ONE path of soundude462/findKmer: the counting path
 *     node_t* findKmer(node_t*, unsigned long long* baseCounter,
 *                      statistics_t* baseStatistics, unsigned long long* TotalNumSequencesN)
 *     (reference: findKmer/src/findKmer.cpp:962-1069, trie insert :612-690)
 * whose only caller is main() (:1320) and whose only consumers are statistics() (:1323) and
 * histo_recursive() (:1342).  The reference has no FFI of its own; this header IS the seam a
 * maintainer would bind (INTEGRATION.md shows the patch to the reference's main()).
 *
 * What replaces what:
 *   trie root + depth-k node frequencies (node_t, :107-111)  -> dense table uint32[4^k], index
 *        = sum b_i * 4^(k-1-i), A0 C1 G2 T3 (base2int, :567-589), i.e. row order of histo_recursive (:719-724)
 *   *baseCounter (:1040,:1056)                               -> fkb_counts.base_total
 *   baseStatistics[b].Count (:1041,:1051)                    -> fkb_counts.base_count[b]
 *   *TotalNumSequencesN (:1042,:1057)                        -> fkb_counts.n_kmers
 *   global nodeCounter (:128,:620)                           -> fkb_counts.node_count
 *   one stderr line per unknown byte (base2int :581-585)     -> fkb_counts.unknown_chars (a count)
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function returns an FKB_*
 * status (0 = ok) and never calls exit(); the caller owns input and result buffers, the library
 * owns its device scratch; one host thread per context.  There is NO CPU fallback: every
 * counting entry point fails with FKB_ERR_CUDA when no sm_100 device is usable.
 */
#ifndef FINDKMER_B200_H
#define FINDKMER_B200_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>

#ifdef __cplusplus
extern "C" {
#endif

#define FKB_MAX_K 16 /* dense uint32 tables: 4^16 * 4 B = 16 GiB.  The reference's trie accepts k <= 20 (:337,:438). */

/* status codes */
#define FKB_OK 0
#define FKB_ERR_EMPTY_INPUT 1         /* reference: "Sequence File Is Empty" + exit (:982-985) */
#define FKB_ERR_UNTERMINATED_HEADER 2 /* reference: '>' line without '\n' before EOF spins forever (:999,:1005) */
#define FKB_ERR_COUNTER_ROLLOVER 3    /* reference: "COUNTER ROLLOVER DETECTED" + exit (:640-648) */
#define FKB_ERR_BAD_K 4               /* reference: "%d is not a valid value for k" + exit (:337-342, :438-443) */
#define FKB_ERR_NOMEM 5
#define FKB_ERR_CUDA 6                /* any CUDA runtime failure, or no usable device */
#define FKB_ERR_BAD_ARG 7
#define FKB_ERR_IO 8
#define FKB_ERR_ZERO_BASE_PROBABILITY 9 /* reference: "Division overflow detected in statistics." + exit (:522-525) */

typedef struct fkb_context fkb_context;

/* Host-side result: everything findKmer() hands to statistics()/histo_recursive() besides the table. */
typedef struct fkb_counts {
    uint64_t n_kmers;       /* TotalNumSequencesN: windows of k consecutive valid bases */
    uint64_t base_total;    /* baseCounter: valid bases lying in runs of length >= k */
    uint64_t base_count[4]; /* per-base split of base_total, A C G T (reference keeps 32 bits; no wrap below 2^32) */
    uint64_t node_count;    /* nodeCounter: 1 + distinct trie prefixes the reference would have allocated */
    uint64_t unknown_chars; /* bytes outside {A,C,G,T,N} that reached base2int (one stderr line each in the reference) */
    uint64_t stream_bytes;  /* length of the stripped stream that was counted */
    uint64_t runs_ge_k;     /* number of maximal valid runs of length >= k */
    uint64_t valid_bases;   /* A/C/G/T bytes in the stream (>= base_total) */
    uint64_t rollover;      /* nonzero: a trie node of the reference would have been visited 2^32 times (:640-648) */
} fkb_counts;

/*
 * Device-side accumulators a shard produces besides its table.  All fields are sums (or, for the
 * prefix flags, maxima) over shards, so a multi-GPU run combines them with one reduce each.
 */
typedef struct fkb_partials {
    unsigned long long head_base[4];   /* bases in the first k-1 positions of every run of length >= k */
    unsigned long long short_first[4]; /* positions with run length < k, by the first base of their run */
    unsigned long long runs_ge_k;
    unsigned long long unknown_chars;
    unsigned long long valid_bases;
    unsigned long long n_windows;      /* windows counted (independent of the table; rollover cross-check) */
} fkb_partials;

/* ---- lifetime -------------------------------------------------------------------------------- */
int fkb_create(int device, fkb_context **ctx);        /* binds the context to one GPU (cudaSetDevice) */
void fkb_destroy(fkb_context *ctx);
const char *fkb_last_error(const fkb_context *ctx);   /* text of the last failure on this context */
const char *fkb_status_string(int status);
const char *fkb_version(void);
int fkb_device_info(fkb_context *ctx, int *sm_count, int *cc_major, int *cc_minor, size_t *hbm_bytes);
/* Tuning knobs (tests and profiling).  "variant": 0 = choose by k and range length (default), 1 = always the
 * direct kernel (one global red per window), 2 = the bucketed kernels whenever k and the range allow.
 * The environment variable FKB_VARIANT sets the same knob at fkb_create time.
 * "loader": 0 = pinned host input is stripped on the GPU, pageable input by the host loader threads (default),
 * 1 = always the host loader, 2 = always the device loader (FKB_LOADER=host|device).
 * "loader_chunk": raw bytes per device-loader chunk (0 = 128 MiB).
 * "loader_slots": 4 MiB pinned slots of the host loader's ring, 2..64 (default 24; FKB_LOADER_SLOTS).
 * "phase_events": 1 = record CUDA events around the kernels of the bucketed path (see fkb_phase_times).
 * "p1_ring": 1 = the k = 11 path routes through ring-shaped staging rows that the warps flush asynchronously (an experiment
 * kept for measurement: same counts, slower than the default synchronous flush; FKB_P1_RING=1 sets it at fkb_create time). */
int fkb_set_option(fkb_context *ctx, const char *name, long value);
/* Profiling aid (no counterpart in the reference): with option "phase_events" = 1 the bucketed count path records CUDA events
 * around its kernels; this returns the durations of the LAST such count call in milliseconds -- ms[0] pass 1 (bucketize),
 * ms[1] pass 2 (count_buckets incl. the shared-memory fold), ms[2] fold kernels (k <= 8; 0 otherwise).  Synchronises. */
int fkb_phase_times(fkb_context *ctx, double ms[3]);

/* ---- host loader: the "stream contract" of findKmer()'s outer loop (:988-1011) ----------------
 * Drops every '\n'; turns each '>'...'\n' header into ONE '>' byte; stops at the first byte 0xFF
 * outside a header (the reference stores fgetc() in a char, so 0xFF aliases EOF, :975,:988);
 * copies every other byte verbatim.  Multi-threaded (n_threads <= 0: hardware concurrency).
 * stream must have room for len bytes.  Pure host code; needs no GPU. */
int fkb_strip_fasta(const uint8_t *fasta, size_t len, uint8_t *stream, size_t *stream_len, int n_threads);

/* pinned host buffers for the loader / result tables (cudaHostAlloc / cudaFreeHost) */
int fkb_alloc_pinned(fkb_context *ctx, size_t bytes, void **ptr);
int fkb_free_pinned(fkb_context *ctx, void *ptr);

/* ---- counting, device-resident (the hot path; what bench.py's `value` times) ------------------
 * d_stream : stripped stream in HBM, 16-byte aligned, readable over [0, end)
 * [begin,end): the byte range this call OWNS -- a window is counted here iff its LAST byte lies in
 *            the range; bytes [max(0,begin-k), begin) are the shard's left halo.
 * d_table  : uint32[4^k] in HBM, ACCUMULATED into (caller zeroes it, see fkb_zero_table_device)
 * d_flags  : uint8[fkb_prefix_flags_bytes(k)] short-run prefix flags, accumulated (OR)
 * d_partials: fkb_partials in HBM, accumulated
 * cuda_stream: a cudaStream_t (0 = default stream).  Asynchronous.  */
size_t fkb_table_entries(int k);      /* 4^k */
size_t fkb_prefix_flags_bytes(int k); /* sum_{d=1..k-1} 4^d, at least 16 */
int fkb_zero_device(fkb_context *ctx, int k, uint32_t *d_table, uint8_t *d_flags, fkb_partials *d_partials,
                    void *cuda_stream);
int fkb_count_stream_device(fkb_context *ctx, const uint8_t *d_stream, uint64_t begin, uint64_t end, int k,
                            uint32_t *d_table, uint8_t *d_flags, fkb_partials *d_partials, void *cuda_stream);

/* Turn (table, flags, partials) -- after any cross-GPU reduction -- into fkb_counts.
 * d_counts: fkb_counts in HBM (asynchronous); copy it back when the stream has drained. */
int fkb_finalize_device(fkb_context *ctx, int k, const uint32_t *d_table, uint8_t *d_flags,
                        const fkb_partials *d_partials, uint64_t stream_bytes, fkb_counts *d_counts,
                        void *cuda_stream);

/* ---- counting, host buffers (the end-to-end path; what bench.py's `e2e` times) ----------------
 * fasta/len : raw file bytes exactly as the reference would fgetc() them (pinned memory is faster)
 * table     : caller-provided uint32[4^k] on the host, overwritten
 * Strips on host threads into pinned staging, overlaps H2D copies with the count kernels, copies
 * the table and counts back.  Returns FKB_ERR_EMPTY_INPUT / FKB_ERR_UNTERMINATED_HEADER /
 * FKB_ERR_COUNTER_ROLLOVER where the reference would have exited or hung. */
int fkb_count_fasta_host(fkb_context *ctx, const uint8_t *fasta, size_t len, int k, uint32_t *table,
                         fkb_counts *counts);
/* same, for an already stripped stream held on the host */
int fkb_count_stream_host(fkb_context *ctx, const uint8_t *stream, size_t len, int k, uint32_t *table,
                          fkb_counts *counts);
/* One shard of a file, for multi-GPU runs.  buf[0,len) holds raw file bytes; the shard OWNS buf[own_offset,len)
 * and buf[0,own_offset) is look-back context (it must reach back to a '\n' or to the start of the file, and
 * far enough to hold 16 sequence bytes).  Results are ACCUMULATED into the caller's device buffers (to be
 * reduced across GPUs and then passed to fkb_finalize_device).  stop_offset: offset in buf of a byte 0xFF that
 * ended the scan, or UINT64_MAX.  ends_in_header: the shard ends inside a '>' line (an error only at end of file). */
int fkb_count_fasta_host_range(fkb_context *ctx, const uint8_t *buf, size_t len, size_t own_offset, int k,
                               uint32_t *d_table, uint8_t *d_flags, fkb_partials *d_partials,
                               uint64_t *stream_bytes, uint64_t *stop_offset, int *ends_in_header);
/* Several k over the same file image in one call -- the launcher's shape (k6thru11fullANDupstream.sh:16-24 starts one
 * process per k on the same file): the bytes cross PCIe and are stripped ONCE, then every k is counted over the
 * device-resident stream.  tables[i]: uint32[4^ks[i]] on the host; counts[i] likewise.  Returns the first hard error, or
 * FKB_ERR_COUNTER_ROLLOVER if any k rolled over (the other results are valid). */
int fkb_count_fasta_host_multi(fkb_context *ctx, const uint8_t *fasta, size_t len, const int *ks, int n_k,
                               uint32_t *const *tables, fkb_counts *counts);
int fkb_count_file_multi(fkb_context *ctx, const char *path, const int *ks, int n_k, uint32_t *const *tables,
                         fkb_counts *counts);
/* same, reading the file itself: the reference's `-p <file>` (:416-428, :344) */
int fkb_count_file(fkb_context *ctx, const char *path, int k, uint32_t *table, fkb_counts *counts);

/* ---- several GPUs of one box (no context argument: one context per device is created and destroyed inside) ----
 * The file image is cut into n_devices contiguous byte ranges; one host thread per GPU counts its range through the host
 * pipeline (16-byte left halo from the look-back context, exactly as fkb_count_fasta_host_range); devices[0] then adds the
 * other GPUs' tables / prefix flags / partials -- reading them from peer memory over NVLink where the devices can map each
 * other, through a staging copy otherwise -- finalizes and returns table and counts as fkb_count_fasta_host does.
 * The reference has no counterpart (its parallelism is one process per k, k6thru11fullANDupstream.sh:16-24); results are
 * bit-identical to the single-GPU call.  err/err_len: optional buffer for the text of a failure. */
int fkb_count_fasta_host_gpus(const int *devices, int n_devices, const uint8_t *fasta, size_t len, int k, uint32_t *table,
                              fkb_counts *counts, char *err, size_t err_len);
int fkb_count_file_gpus(const int *devices, int n_devices, const char *path, int k, uint32_t *table, fkb_counts *counts,
                        char *err, size_t err_len);
int fkb_device_count(void); /* usable CUDA devices (0 when there is none) */

/* ---- writers: statistics() (:491-565) and histo_recursive() (:699-942) on the dense table -----
 * Byte-identical files to the reference's (x87 long double arithmetic, same expression order).
 * max_nodes = estimate_RAM_usage()'s return value, 1 + sum_{n=1..k} 4^n (:1256-1260). */
int fkb_write_base_stats(FILE *stats_out, FILE *console, int k, const fkb_counts *counts,
                         long double base_probability[4]);
int fkb_write_histogram(FILE *csv_out, int k, const uint32_t *table, const fkb_counts *counts,
                        const long double base_probability[4], int z_threshold_enable, long double z_threshold,
                        int n_threads, uint64_t *rows_written);
/* The same rows for mergeFile4GNUPLOT.pl (findKmer/mergeFile4GNUPLOT.pl:13-32 joins two TAB-separated tables on column 0 and
 * carries columns 1 and 2): one line per observed k-mer, `kmer \t h \t frequency \t H [\t z] \n`, no header line, same
 * row order, same number formats and same Z-score filter as fkb_write_histogram.  An extension: the reference writes CSV only. */
int fkb_write_histogram_tsv(FILE *tsv_out, int k, const uint32_t *table, const fkb_counts *counts,
                            const long double base_probability[4], int z_threshold_enable, long double z_threshold,
                            int n_threads, uint64_t *rows_written);
uint64_t fkb_max_nodes(int k);

/* ---- synthetic inputs of the BASELINE.json shapes, generated in HBM (bench only) --------------
 * Renders bytes [first_byte, first_byte + n_bytes) of the virtual file into d_out[0, n_bytes);
 * bit-identical to findkmer_b200/synth.py::render().  line_width > 0: wrapped lines, 0: one
 * unwrapped line per record, < 0: the stripped layout (no newlines; with header_len == 1 and
 * header ">" this IS the stream contract's output).  rec_offsets/rec_base0: uint64[n_records+1]
 * on the HOST; headers: n_records * header_len bytes on the HOST. */
int fkb_synth_fasta_device(fkb_context *ctx, uint8_t *d_out, uint64_t first_byte, uint64_t n_bytes, int n_records,
                           const uint64_t *rec_offsets, const uint64_t *rec_base0, const uint8_t *headers,
                           int header_len, int line_width, uint64_t seed, int n_runs, int soft_mask,
                           void *cuda_stream);

/* number of kernels this library has launched on this context since creation (bench: gpu_launches) */
uint64_t fkb_launch_count(const fkb_context *ctx);

#ifdef __cplusplus
}
#endif
#endif /* FINDKMER_B200_H */
