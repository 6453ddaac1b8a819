/*
 * oracle/ref_probe.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT.
 *
 * Compiles the UNTOUCHED reference translation unit (included from where it lies under
 * /root/reference -- nothing is copied) with its main() renamed, and wraps it so that the
 * two hand-over values the reference never prints can be observed:
 *   nodeCounter          (global, findKmer.cpp:128)  -> tree density / "All possible ..." line
 * The reference always dies in an invalid free() after closing its outputs
 * (findKmer.cpp:1370-1371), so the values are written from SIGSEGV/SIGABRT handlers as well
 * as from atexit (covers the early exit(EXIT_FAILURE) paths).
 *
 * Output: one line "nodeCounter=<n>\n" appended to the file named by $FINDKMER_PROBE_OUT.
 */
#ifndef FINDKMER_REF_SRC
#error "build with -DFINDKMER_REF_SRC='\"/root/reference/findKmer/src/findKmer.cpp\"'"
#endif

#define main findKmer_reference_main
#include FINDKMER_REF_SRC
#undef main

#include <signal.h>
#include <unistd.h>
#include <fcntl.h>

static void probe_dump(void)
{
    const char *path = getenv("FINDKMER_PROBE_OUT");
    if (!path) return;
    int fd = open(path, O_WRONLY | O_CREAT | O_APPEND, 0644);
    if (fd < 0) return;
    char line[64];
    int n = snprintf(line, sizeof line, "nodeCounter=%llu\n", nodeCounter);
    if (n > 0) (void)!write(fd, line, (size_t)n);
    close(fd);
}

static void probe_on_signal(int sig)
{
    fflush(stdout);
    probe_dump();
    _exit(128 + sig);
}

int main(int argc, char *argv[])
{
    signal(SIGSEGV, probe_on_signal);
    signal(SIGABRT, probe_on_signal);
    atexit(probe_dump);
    return findKmer_reference_main(argc, argv);
}
