// Host->device copy rate of the e2e path's pattern: 120 frames of 6.2 MB from pinned memory.
// Build: nvcc -O3 -o scripts/ubench_h2d scripts/ubench_h2d.cu
#include <cstdio>
#include <cstring>
#include <cuda_runtime.h>
int main()
{
    const size_t fb = 1080 * 1920 * 3, n = 120, total = fb * n;
    char *d; cudaMalloc(&d, total);
    cudaStream_t s[2]; cudaStreamCreate(&s[0]); cudaStreamCreate(&s[1]);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int flags : {0, (int)cudaHostAllocWriteCombined}) {
        char *h; if (cudaHostAlloc(&h, total, flags) != cudaSuccess) { printf("alloc failed\n"); return 1; }
        memset(h, 1, total);
        for (int mode = 0; mode < 3; mode++) {      // 0: one copy, 1: per frame one stream, 2: per frame two streams
            float best = 1e9;
            for (int rep = 0; rep < 4; rep++) {
                cudaDeviceSynchronize();
                cudaEventRecord(e0, s[0]);
                if (mode == 0) cudaMemcpyAsync(d, h, total, cudaMemcpyHostToDevice, s[0]);
                else for (size_t i = 0; i < n; i++) cudaMemcpyAsync(d + i * fb, h + i * fb, fb, cudaMemcpyHostToDevice, s[mode == 2 ? i & 1 : 0]);
                if (mode == 2) { cudaEvent_t j; cudaEventCreate(&j); cudaEventRecord(j, s[1]); cudaStreamWaitEvent(s[0], j, 0); }
                cudaEventRecord(e1, s[0]); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
            }
            printf("flags=%d mode=%d: %.3f ms  %.1f GB/s\n", flags, mode, best, total / best / 1e6);
        }
        cudaFreeHost(h);
    }
    return 0;
}
