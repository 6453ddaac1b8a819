// hostreg_probe.cu -- can the CLI page-lock its mmap'ed input in windows and let the GPU read it in place?
// (VERDICT r1 item 6 / DESIGN.md 9.5.)  Measures, on a file in page cache: cudaHostRegister / cudaHostUnregister of one
// window of the mapping (read-only), H2D from the registered window, and the same with N windows registered by N threads.
//   nvcc -O2 -o /tmp/hostreg_probe profiles/tools/hostreg_probe.cu && /tmp/hostreg_probe <file> [window MiB]
#include <cuda_runtime.h>
#include <fcntl.h>
#include <stdio.h>
#include <stdlib.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>

#include <chrono>
#include <thread>
#include <vector>

static double now_ms() { return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count(); }

int main(int argc, char **argv)
{
    if (argc < 2) return 2;
    const size_t win = (size_t)(argc > 2 ? atoi(argv[2]) : 128) << 20;
    int fd = open(argv[1], O_RDONLY);
    struct stat st;
    fstat(fd, &st);
    const size_t len = (size_t)st.st_size;
    unsigned char *map = (unsigned char *)mmap(nullptr, len, PROT_READ, MAP_PRIVATE, fd, 0);
    if (map == MAP_FAILED) { perror("mmap"); return 1; }
    cudaFree(0);
    unsigned char *d;
    cudaMalloc(&d, win * 2);
    const size_t n_win = len / win;
    printf("file %.2f GB, window %zu MiB, %zu windows\n", len / 1e9, win >> 20, n_win);
    for (unsigned flags : {(unsigned)cudaHostRegisterReadOnly, (unsigned)cudaHostRegisterDefault}) {
        double t0 = now_ms();
        cudaError_t e = cudaHostRegister(map, win, flags);
        double t1 = now_ms();
        printf("cudaHostRegister(flags=%u) of window 0: %s, %.1f ms (%.2f GB/s)\n", flags, cudaGetErrorString(e), t1 - t0, win / (t1 - t0) / 1e6);
        if (e != cudaSuccess) { cudaGetLastError(); continue; }
        double t2 = now_ms();
        cudaMemcpy(d, map, win, cudaMemcpyHostToDevice);
        double t3 = now_ms();
        printf("  H2D from the registered window: %.1f ms (%.2f GB/s)\n", t3 - t2, win / (t3 - t2) / 1e6);
        cudaHostUnregister(map);
        printf("  cudaHostUnregister: %.1f ms\n", now_ms() - t3);
        // N windows by N threads (does registration scale?)
        for (int nt : {2, 4, 8}) {
            if ((size_t)nt + 1 > n_win) break;
            std::vector<std::thread> pool;
            double a = now_ms();
            for (int t = 0; t < nt; ++t) pool.emplace_back([&, t] { cudaHostRegister(map + (size_t)(t + 1) * win, win, flags); });
            for (auto &th : pool) th.join();
            double b = now_ms();
            printf("  %d windows registered by %d threads: %.1f ms (%.2f GB/s aggregate)\n", nt, nt, b - a, nt * (double)win / (b - a) / 1e6);
            for (int t = 0; t < nt; ++t) cudaHostUnregister(map + (size_t)(t + 1) * win);
        }
        break;
    }
    // reference points: pageable H2D straight from the mapping, and pread into a pinned buffer
    double t4 = now_ms();
    cudaMemcpy(d, map + (n_win - 1) * win, win, cudaMemcpyHostToDevice);
    printf("pageable cudaMemcpy from the mapping (one window): %.1f ms (%.2f GB/s)\n", now_ms() - t4, win / (now_ms() - t4) / 1e6);
    unsigned char *pin;
    double t5 = now_ms();
    cudaHostAlloc((void **)&pin, win, cudaHostAllocDefault);
    printf("cudaHostAlloc of one window: %.1f ms\n", now_ms() - t5);
    double t6 = now_ms();
    {
        const int nt = 8;
        std::vector<std::thread> pool;
        for (int t = 0; t < nt; ++t)
            pool.emplace_back([&, t] { size_t part = win / nt; (void)!pread(fd, pin + t * part, part, (off_t)(t * part)); });
        for (auto &th : pool) th.join();
    }
    printf("pread of one window into pinned memory, 8 threads: %.1f ms (%.2f GB/s)\n", now_ms() - t6, win / (now_ms() - t6) / 1e6);
    return 0;
}
