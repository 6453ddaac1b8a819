// findkmer_main.cpp -- the drop-in `findKmer` program: same flags, defaults, file names and stdout phrases as
// the reference's main() (findKmer/src/findKmer.cpp:1292-1379), with the counting path
// (`findKmer()`, :962-1069) replaced by one call into the sm_100a library (include/findkmer_b200.h).
//
//   reference                                   here
//   init_conf/parse_arguments (:246-255,:394-490)  Config + parse_arguments(): same grammar, same messages
//   set_default_conf/print_conf (:257-364)         finish_config()/print_config(): same file naming, same banner
//   estimate_RAM_usage (:1226-1291)                print_memory_estimate(): same sentence, the dense table's numbers
//   findKmer (:962-1069)                           fkb_count_file()  -> GPU
//   statistics (:491-565)                          fkb_write_base_stats()
//   histo_recursive (:699-942)                     fkb_write_histogram()
//
// Deliberate differences (DESIGN.md "boundary"): exit status 0 on success (the reference always dies in an
// invalid free() after closing its files, :1370-1371); a missing option value prints usage once and exits 1
// instead of looping forever (:1302-1303); unknown characters are reported once with a count instead of one
// stderr line each (:581-585); k = 17..20 is refused (dense 4^k tables); -K/--ksweep (several k in one run: the file is
// uploaded and stripped once), -g/--gpus (contiguous shards of the file on several GPUs of this box) and -t/--tsv (the
// histogram once more as the tab-separated table mergeFile4GNUPLOT.pl joins) are extensions.
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <string>
#include <thread>
#include <vector>

#include "findkmer_b200.h"

#define DEFAULT_SEQUENCE_FILE_NAME "test.txt"  // :76
#define DEFAULT_K_VALUE 7                      // :78
#define OUT_FILE_COLUMN_HEADERS "Sequence, Shannon Entropy h, Shannon Entropy H, Frequency, Z score"  // :79
#define DEFAULT_SUPPRESS_OUTPUT_VALUE 0        // :80
#define DEFAULT_Z_THRESHOLD_ENABLE 0           // :81
#define DEFAULT_Z_THRESHOLD 1000               // :82

namespace {

struct Config {
    std::string sequence_file;
    bool have_sequence_file = false;
    std::string out_file;
    bool have_out_file = false;
    int k = 0;
    int suppress_output = -1;  // tri-state: -1 = not given (:252)
    int z_enable = -1;
    long double z_threshold = -1;
    int gpus = 1;
    std::vector<int> sweep;  // -K/--ksweep a-b | a,b,c : several k in one run (extension; one upload, one strip)
    std::string tsv_file;    // -t/--tsv <file>: tab-separated copy of the histogram for mergeFile4GNUPLOT.pl (extension)
};

void usage()
{
    fprintf(stdout, "\n");
    fprintf(stdout, "Usage: findKmer [options]\n");
    fprintf(stdout, "             [--parse|-p <sequence_file.txt>] \n"
                    "               File with DNA sequence data.\n"
                    "               File must be in current directory.\n"
                    "               Parser follows .fas and .fa formats\n"
                    "                Default is %s.\n\n", DEFAULT_SEQUENCE_FILE_NAME);
    fprintf(stdout, "             [--export|-e  <out_file.csv>] \n"
                    "               File to output histogram data to.\n"
                    "                Default output file name is dynamic.\n\n");
    fprintf(stdout, "             [--ksize|-k  <k>] \n"
                    "               Size of sequence for histogram.\n"
                    "                Default is %d.\n\n", DEFAULT_K_VALUE);
    fprintf(stdout, "             [--quiet|-q  < 0 for FALSE | 1 for TRUE >] \n"
                    "               Suppress file read output and breaks.\n"
                    "                Default is %s.\n\n", DEFAULT_SUPPRESS_OUTPUT_VALUE ? "true" : "false");
    long double z = DEFAULT_Z_THRESHOLD;
    fprintf(stdout, "             [--zthreshold|-z  < Threshold_for_Z >] \n"
                    "               Suppress sequences with Z scores < threshold.\n"
                    "                Default is %s with a value of %LG.\n\n", DEFAULT_Z_THRESHOLD_ENABLE ? "enabled" : "disabled", z);
    fprintf(stdout, "\n");
}

void check_file(const char *name, const char *mode)  // :233-242
{
    FILE *f = fopen(name, mode);
    if (!f) {
        fprintf(stderr, "Unable to open file %s in %s mode\nFile MUST be in current directory.\n", name, mode);
        exit(EXIT_FAILURE);
    }
    fclose(f);
}

// returns false when an option value is missing (the reference then re-prints usage, :1302-1303)
bool parse_arguments(int argc, char **argv, Config &cfg)
{
    for (int i = 1; i < argc; ++i) {
        const char *a = argv[i];
        if (!strcmp(a, "-h") || !strcmp(a, "--help")) {
            exit(1);
        } else if (!strcmp(a, "-e") || !strcmp(a, "--export")) {
            if (++i == argc) { fprintf(stderr, "Export file name missing.\n"); return false; }
            check_file(argv[i], "w");
            cfg.out_file = argv[i];
            cfg.have_out_file = true;
        } else if (!strcmp(a, "-p") || !strcmp(a, "--parse")) {
            if (++i == argc) { fprintf(stderr, "Sequence data file name missing.\n"); return false; }
            check_file(argv[i], "r");
            cfg.sequence_file = argv[i];
            cfg.have_sequence_file = true;
        } else if (!strcmp(a, "-k") || !strcmp(a, "--ksize")) {
            if (++i == argc) { fprintf(stderr, "Number for size of k is missing.\n"); return false; }
            int k = atoi(argv[i]);
            if (k < 0 || k > 20) {
                fprintf(stderr, "%d is not a valid value for k.\nPlease select a number greater than zero and less than 21\n", k);
                exit(EXIT_FAILURE);
            }
            cfg.k = k;
        } else if (!strcmp(a, "-q") || !strcmp(a, "--quiet")) {
            if (++i == argc) {
                fprintf(stderr, "True/false value for quiet option is missing.\nUsage is \"-q 1\" for suppression OR \"-q 0\" for expansion\n");
                exit(EXIT_FAILURE);
            }
            int q = atoi(argv[i]);
            if (q == 0 || q == 1) {
                cfg.suppress_output = q;
            } else {
                fprintf(stderr, "%d is not a valid value for suppress Output Enable Option.\nPlease select either 0 for FALSE or a 1 for TRUE", q);
                exit(EXIT_FAILURE);
            }
        } else if (!strcmp(a, "-z") || !strcmp(a, "--zthreshold")) {
            if (++i == argc) {
                fprintf(stderr, "Z threshold number is missing\nUsage is \"-z 1000\".\n");
                exit(EXIT_FAILURE);
            }
            cfg.z_enable = 1;
            cfg.z_threshold = atoi(argv[i]);
        } else if (!strcmp(a, "-K") || !strcmp(a, "--ksweep")) {  // extension: the launcher's k loop in one process
            if (++i == argc) { fprintf(stderr, "k sweep is missing. Usage is \"-K 6-11\" or \"-K 6,8,11\".\n"); return false; }
            const char *p = argv[i];
            int lo = atoi(p);
            const char *dash = strchr(p, '-');
            if (dash) {
                int hi = atoi(dash + 1);
                for (int k = lo; k <= hi; ++k) cfg.sweep.push_back(k);
            } else {
                for (const char *q = p; q && *q; q = strchr(q, ',') ? strchr(q, ',') + 1 : nullptr) cfg.sweep.push_back(atoi(q));
            }
            for (int k : cfg.sweep)
                if (k < 1 || k > FKB_MAX_K) {
                    fprintf(stderr, "%d is not a valid value for k.\nPlease select a number greater than zero and less than %d\n", k, FKB_MAX_K + 1);
                    exit(EXIT_FAILURE);
                }
        } else if (!strcmp(a, "-t") || !strcmp(a, "--tsv")) {  // extension: the table mergeFile4GNUPLOT.pl joins
            if (++i == argc) { fprintf(stderr, "TSV file name missing.\n"); return false; }
            check_file(argv[i], "w");
            cfg.tsv_file = argv[i];
        } else if (!strcmp(a, "-g") || !strcmp(a, "--gpus")) {  // extension
            if (++i == argc) { fprintf(stderr, "Number of GPUs is missing.\n"); return false; }
            cfg.gpus = atoi(argv[i]) > 0 ? atoi(argv[i]) : 1;
        } else {
            fprintf(stderr, "Ignoring invalid option %s\n", a);
            if (cfg.suppress_output == 0) {
                fprintf(stderr, "Press enter to continue.\n");
                getchar();
            }
        }
    }
    return true;
}

void finish_config(Config &cfg)  // set_default_conf, :257-305
{
    if (!cfg.have_sequence_file) cfg.sequence_file = DEFAULT_SEQUENCE_FILE_NAME;
    if (!cfg.k) cfg.k = DEFAULT_K_VALUE;
    if (cfg.suppress_output < 0) cfg.suppress_output = DEFAULT_SUPPRESS_OUTPUT_VALUE;
    if (cfg.z_enable < 0) {
        cfg.z_enable = DEFAULT_Z_THRESHOLD_ENABLE;
        cfg.z_threshold = DEFAULT_Z_THRESHOLD;
    }
    if (!cfg.have_out_file) {
        cfg.out_file = std::to_string(cfg.k) + "mer_Historam_Of_" + cfg.sequence_file + (cfg.z_enable ? "zScoreFiltered" : "") + ".csv";
    }
}

void print_memory_estimate(const Config &cfg)  // estimate_RAM_usage, :1256-1289 (sentence kept, numbers are the dense engine's)
{
    const uint64_t max_nodes = fkb_max_nodes(cfg.k);
    const double disk = (double)(cfg.k + 10) * (double)max_nodes;
    const double gib = 1024.0 * 1024.0 * 1024.0, mib = 1024.0 * 1024.0;
    if (disk >= gib) fprintf(stdout, "%g gibibytes", disk / gib);
    else fprintf(stdout, "%g mibibytes", disk / mib);
    fprintf(stdout, " of disk usage and ");
    const double ram = (double)fkb_table_entries(cfg.k) * 4.0 + (double)fkb_prefix_flags_bytes(cfg.k);
    if (ram >= gib) {
        fprintf(stdout, "%g gibibytes of RAM usage likely\n", ram / gib);
        fprintf(stdout, "We are stopping here to make sure that is ok with you!\n");
        fprintf(stdout, "Hit enter to proceed or else abort the program.\n");
        if (cfg.suppress_output == 0) getchar();
    } else {
        fprintf(stdout, "%g mibibytes of RAM usage likely\n", ram / mib);
    }
}

// The per-record progress lines of a non-quiet run (findKmer() :994-1003): at every '>' outside a header the reference
// prints "Read <baseCounter> bases", then echoes the header line.  baseCounter at that moment is the number of bases in
// runs of length >= k seen so far (:1040, :1056) -- a running sum over the file that the one-shot GPU count does not
// produce, so this (and only this) is scanned on the host, and only when -q 0 asks for it; the k-mer table never is.
void print_progress_lines(const char *path, int k)
{
    int fd = open(path, O_RDONLY);
    if (fd < 0) return;
    struct stat st;
    if (fstat(fd, &st) != 0 || st.st_size == 0) { close(fd); return; }
    const size_t len = (size_t)st.st_size;
    const unsigned char *buf = (const unsigned char *)mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0);
    close(fd);
    if (buf == MAP_FAILED) return;
    unsigned long long base_counter = 0, run = 0;
    for (size_t p = 0; p < len; ++p) {
        const unsigned char c = buf[p];
        if (c == 0xFF) break;  // fgetc() stored in a char: EOF alias (:975, :988)
        if (c == '>') {
            run = 0;
            const unsigned char *nl = (const unsigned char *)memchr(buf + p, '\n', len - p);
            const size_t e = nl ? (size_t)(nl - buf) : len;
            fprintf(stdout, "Read %llu bases\n", base_counter);
            fwrite(buf + p, 1, e - p, stdout);
            fprintf(stdout, "\n");
            p = e;  // the '\n' itself is ignored by the scan (:1011)
            continue;
        }
        if (c == '\n') continue;
        if (c == 'A' || c == 'C' || c == 'G' || c == 'T') {
            ++run;
            if (run > (unsigned long long)k) ++base_counter;            // :1040
            else if (run == (unsigned long long)k) base_counter += k;   // :1056
        } else {
            run = 0;  // :1019-1024
        }
    }
    munmap((void *)buf, len);
}

}  // namespace

// FKB_TIMING=1: wall-clock phases of the program on stderr (profiling aid; nothing is printed otherwise)
static void phase_mark(const char *what)
{
    static const bool on = getenv("FKB_TIMING") != nullptr;
    static auto last = std::chrono::steady_clock::now();
    if (!on) return;
    auto now = std::chrono::steady_clock::now();
    fprintf(stderr, "[timing] %-28s %8.1f ms\n", what, std::chrono::duration<double, std::milli>(now - last).count());
    last = now;
}

int main(int argc, char **argv)
{
    phase_mark("start");
    Config cfg;
    usage();  // the reference prints usage on every run (:1301)
    if (!parse_arguments(argc, argv, cfg)) {
        usage();
        return EXIT_FAILURE;
    }

    // print_conf, :307-364
    fprintf(stdout, "\nATTEMPTING CONFIGURATION: \n");
    if (!cfg.sweep.empty()) cfg.k = cfg.sweep[0];  // the file opened below is then the first k's file, not a stray one
    finish_config(cfg);
    fprintf(stdout, "- sequence_file file: %s\n", cfg.sequence_file.c_str());
    fprintf(stdout, "- export file: %s\n", cfg.out_file.c_str());
    fprintf(stdout, "- k size: %d\n", cfg.k);
    fprintf(stdout, "- %s\n", cfg.suppress_output > 0 ? "Suppressing file read output and breaks." : "Showing DNA Sequence identifier and allowing breaks.");
    fprintf(stdout, "- Z score filtering is %s", cfg.z_enable ? "enabled" : "disabled");
    if (cfg.z_enable > 0) fprintf(stdout, "\n    with threshold of %LG", cfg.z_threshold);
    fprintf(stdout, ".\n");
    if (cfg.suppress_output == 0 && argc < 2) {
        fprintf(stdout, "Press enter to proceed with this configuration.");
        getchar();
    }
    if (cfg.k < 0 || cfg.k > 20) {
        fprintf(stderr, "%d is not a valid value for k. Please select a number greater than zero\n", cfg.k);
        return EXIT_FAILURE;
    }
    if (cfg.k > FKB_MAX_K) {
        fprintf(stderr, "%d is not a valid value for k here: the dense-table GPU engine supports k <= %d (4^k x 4 bytes of HBM)\n", cfg.k, FKB_MAX_K);
        return EXIT_FAILURE;
    }
    FILE *seq_probe = fopen(cfg.sequence_file.c_str(), "r");
    if (!seq_probe) {
        fprintf(stderr, "Sequence file failed to open\n\n");
        return EXIT_FAILURE;
    }
    fclose(seq_probe);
    FILE *csv = fopen(cfg.out_file.c_str(), "w");
    if (!csv) {
        fprintf(stderr, "Out file failed to open\nFile MUST be in current directory.\n");
        return EXIT_FAILURE;
    }
    fprintf(csv, OUT_FILE_COLUMN_HEADERS);  // no newline: every row starts with one (:354, :858)
    fprintf(stdout, "Sequence file and out file opened properly\n");
    fprintf(stdout, "\n");

    print_memory_estimate(cfg);

    fprintf(stdout, "!!!Find The KMER!!!\n");
    fprintf(stdout, "Reading sequence from file\n");
    fprintf(stdout, "     2858658142 bases in the reference genome FYI.\nThat is 2,858,658,142 by the way.\n");
    if (cfg.suppress_output == 0) print_progress_lines(cfg.sequence_file.c_str(), cfg.k);  // :996-1002
    fflush(stdout);

    // ---- the counting path: findKmer() -> GPU ----
    std::vector<int> ks = cfg.sweep.empty() ? std::vector<int>{cfg.k} : cfg.sweep;
    // -g N (extension): contiguous shards of the file on the first N GPUs of this box, tables summed on GPU 0 over NVLink
    const int n_gpus = cfg.gpus > 1 ? (cfg.gpus < fkb_device_count() ? cfg.gpus : fkb_device_count()) : 1;
    if (cfg.gpus > 1 && n_gpus < cfg.gpus)
        fprintf(stderr, "findKmer (B200 engine): -g %d asked for, %d GPU(s) present: using %d\n", cfg.gpus, fkb_device_count(), n_gpus > 0 ? n_gpus : 1);
    const bool multi_gpu = n_gpus > 1 && ks.size() == 1;
    if (n_gpus > 1 && !multi_gpu) fprintf(stderr, "findKmer (B200 engine): -K sweeps run on one GPU; -g ignored\n");
    fkb_context *ctx = nullptr;
    int status = FKB_OK;
    if (!multi_gpu) {
        status = fkb_create(0, &ctx);
        if (status != FKB_OK) {
            fprintf(stderr, "findKmer (B200 engine): no usable sm_100 GPU (%s); this build has no CPU fallback\n", fkb_status_string(status));
            fclose(csv);
            return EXIT_FAILURE;
        }
    }
    phase_mark("fkb_create (CUDA context)");
    std::vector<std::vector<uint32_t>> tables(ks.size());
    std::vector<uint32_t *> table_ptrs(ks.size());
    std::vector<fkb_counts> all_counts(ks.size());
    try {
        for (size_t i = 0; i < ks.size(); ++i) {
            tables[i].resize(fkb_table_entries(ks[i]));
            table_ptrs[i] = tables[i].data();
        }
    } catch (...) {
        fprintf(stderr, "allocate_array():: memory allocation failed\n");
        return EXIT_FAILURE;
    }
    memset(all_counts.data(), 0, sizeof(fkb_counts) * ks.size());
    char gpus_err[512] = "";
    if (multi_gpu) {
        std::vector<int> devices(n_gpus);
        for (int i = 0; i < n_gpus; ++i) devices[i] = i;
        status = fkb_count_file_gpus(devices.data(), n_gpus, cfg.sequence_file.c_str(), ks[0], table_ptrs[0], &all_counts[0], gpus_err, sizeof gpus_err);
    } else if (ks.size() == 1) status = fkb_count_file(ctx, cfg.sequence_file.c_str(), ks[0], table_ptrs[0], &all_counts[0]);
    else status = fkb_count_file_multi(ctx, cfg.sequence_file.c_str(), ks.data(), (int)ks.size(), table_ptrs.data(), all_counts.data());
    phase_mark("count (mmap, load, GPU)");
    if (status == FKB_ERR_EMPTY_INPUT) {
        fprintf(stderr, "Sequence File Is Empty, Ending Program");  // :983
        fclose(csv);
        fkb_destroy(ctx);
        return EXIT_FAILURE;
    }
    if (status == FKB_ERR_COUNTER_ROLLOVER) {  // :643-647, both streams
        const char *msg = "\n\n!!! COUNTER ROLLOVER DETECTED! \nIncrease the number of bits used for the counter variable if you have the source code, else use a smaller sequence file.\n\n";
        fprintf(stderr, "%s", msg);
        fprintf(stdout, "%s", msg);
        fclose(csv);
        fkb_destroy(ctx);
        return EXIT_FAILURE;
    }
    if (status != FKB_OK) {
        fprintf(stderr, "findKmer (B200 engine): %s\n", ctx ? fkb_last_error(ctx) : gpus_err);
        fclose(csv);
        fkb_destroy(ctx);
        return EXIT_FAILURE;
    }
    // the tables are on the host now: release the GPU (cudaFree of the stream and bucket regions takes 20-500 ms) next to
    // the file writing instead of in front of it
    struct Teardown {
        std::thread t;
        ~Teardown() { if (t.joinable()) t.join(); }
    } teardown{std::thread([ctx] { if (ctx) fkb_destroy(ctx); })};
    if (all_counts[0].unknown_chars)
        fprintf(stderr, "Unknown character processed! File may be corrupted. (%llu such characters; the reference prints one line each)\n",
                (unsigned long long)all_counts[0].unknown_chars);

    int exit_code = 0;
    for (size_t i = 0; i < ks.size(); ++i) {
        const int k = ks[i];
        const fkb_counts &counts = all_counts[i];
        std::string out_name = cfg.out_file;
        if (ks.size() > 1) {  // sweep: every k gets the file names a separate run would have produced
            Config one = cfg;
            one.k = k;
            one.have_out_file = false;
            finish_config(one);
            out_name = cfg.have_out_file ? std::to_string(k) + "mer_" + cfg.out_file : one.out_file;
            if (i > 0 || out_name != cfg.out_file) {
                if (i == 0) fclose(csv);
                csv = fopen(out_name.c_str(), "w");
                if (!csv) {
                    fprintf(stderr, "Out file failed to open\nFile MUST be in current directory.\n");
                    return EXIT_FAILURE;
                }
                fprintf(csv, OUT_FILE_COLUMN_HEADERS);
            }
            fprintf(stdout, "\n- k size: %d\n- export file: %s\n", k, out_name.c_str());
        }
        // ---- statistics(), :491-565 ----
        const std::string stats_name = std::to_string(k) + "mer_Base_Stats_Of_" + cfg.sequence_file + ".txt";
        FILE *stats = fopen(stats_name.c_str(), "w");
        if (!stats) {
            fprintf(stderr, "Out file failed to open\nFile MUST be in current directory.\n");
            fclose(csv);
            return EXIT_FAILURE;
        }
        long double base_probability[4];
        status = fkb_write_base_stats(stats, stdout, k, &counts, base_probability);
        fclose(stats);
        if (status == FKB_ERR_ZERO_BASE_PROBABILITY) {  // the reference exits here with a header-only CSV
            fclose(csv);
            csv = nullptr;
            exit_code = EXIT_FAILURE;
            if (ks.size() == 1) return exit_code;
            continue;
        }
        fprintf(stdout, "Now creating histogram.\n");
        fflush(stdout);
        uint64_t rows = 0;
        status = fkb_write_histogram(csv, k, tables[i].data(), &counts, base_probability, cfg.z_enable, cfg.z_threshold, 0, &rows);
        if (status != FKB_OK) {
            fprintf(stderr, "Out file write error! (%s)\n", fkb_status_string(status));
            fclose(csv);
            return EXIT_FAILURE;
        }
        if (!cfg.tsv_file.empty()) {  // extension: the same rows, tab-separated, for mergeFile4GNUPLOT.pl
            const std::string tsv_name = ks.size() > 1 ? std::to_string(k) + "mer_" + cfg.tsv_file : cfg.tsv_file;
            FILE *tsv = fopen(tsv_name.c_str(), "w");
            uint64_t tsv_rows = 0;
            if (!tsv || fkb_write_histogram_tsv(tsv, k, tables[i].data(), &counts, base_probability, cfg.z_enable, cfg.z_threshold, 0, &tsv_rows) != FKB_OK) {
                fprintf(stderr, "TSV file write error!\n");
                exit_code = EXIT_FAILURE;
            }
            if (tsv) fclose(tsv);
        }
        fprintf(stdout, "histogram creation finished.\n");
        if (fclose(csv) == EOF)
            fprintf(stderr, "Out file close error! This is not expected and might mean the data was not written to the file properly before the close.\n");
        csv = nullptr;
        fprintf(stdout, "Your file can be found in the current directory as: \n    %s\n", out_name.c_str());
    }
    phase_mark("statistics + histogram files");
    teardown.t.join();
    phase_mark("fkb_destroy (rest)");
    if (exit_code == 0) fprintf(stdout, "End of program was reached properly.\n\n");
    fprintf(stderr, " ");
    return exit_code;
}
