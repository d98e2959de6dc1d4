// Issue-rate microbenchmark for the integer ops K1 is built from (which pipe does IDP share?).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/ubench_pipes scripts/ubench_pipes.cu
// Prints warp-instructions per cycle per SM for each op and for pairwise mixes.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define ITERS 32768
#define NACC 8

template <int OP> __device__ __forceinline__ uint32_t op1(uint32_t a, uint32_t b, uint32_t c)
{
    uint32_t d;
    if (OP == 0) asm volatile("dp2a.lo.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(b), "r"(a), "r"(c));
    else if (OP == 1) asm volatile("dp4a.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    else if (OP == 2) asm volatile("prmt.b32 %0, %1, %2, 0x5410;" : "=r"(d) : "r"(a), "r"(c));
    else if (OP == 3) asm volatile("{.reg .u32 t; add.u32 t, %1, %2; add.u32 %0, t, %3;}" : "=r"(d) : "r"(a), "r"(b), "r"(c));   // IADD3
    else if (OP == 4) asm volatile("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));  // IMAD
    else if (OP == 5) asm volatile("lop3.b32 %0, %1, %2, %3, 0x96;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    else if (OP == 6) asm volatile("shf.l.wrap.b32 %0, %1, %2, 5;" : "=r"(d) : "r"(a), "r"(c));
    else if (OP == 7) { float f; asm volatile("fma.rn.f32 %0, %1, %2, %3;" : "=f"(f) : "f"(__uint_as_float(a)), "f"(__uint_as_float(b)), "f"(__uint_as_float(c))); d = __float_as_uint(f); }
    else if (OP == 8) { float f; asm volatile("add.rn.f32 %0, %1, %2;" : "=f"(f) : "f"(__uint_as_float(a)), "f"(__uint_as_float(c))); d = __float_as_uint(f); }
    else if (OP == 9) asm volatile("vadd2.u32.u32.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));   // SIMD video add (emulated?)
    else if (OP == 12) asm volatile("mad.lo.u32 %0, %1, 1, %2;" : "=r"(d) : "r"(a), "r"(c));          // IMAD.IADD ?
    else if (OP == 13) asm volatile("mad.lo.u32 %0, %1, -4, %2;" : "=r"(d) : "r"(c), "r"(a));         // IMAD imm
    else if (OP == 10) asm volatile("{.reg .b32 t; add.u16x2 %0, %1, %2;}" : "=r"(d) : "r"(a), "r"(c));
    else d = a;
    return d;
}

template <int OPA, int OPB, int RA, int RB> __global__ void bench(uint32_t *out, uint32_t seed, long long *cyc)
{
    uint32_t acc[NACC], acc2[NACC];
#pragma unroll
    for (int i = 0; i < NACC; i++) { acc[i] = seed + i * 77 + threadIdx.x; acc2[i] = seed * 3 + i + threadIdx.x; }
    uint32_t a = seed ^ threadIdx.x, b = seed * 5 + 1;
    unsigned long long g0; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < ITERS; it++) {
#pragma unroll
        for (int i = 0; i < NACC; i++) {
#pragma unroll
            for (int r = 0; r < RA; r++) acc[i] = op1<OPA>(a, b, acc[i]);
#pragma unroll
            for (int r = 0; r < RB; r++) acc2[i] = op1<OPB>(a, b, acc2[i]);
        }
    }
    long long t1 = clock64();
    unsigned long long g1; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    uint32_t s = 0;
#pragma unroll
    for (int i = 0; i < NACC; i++) s += acc[i] ^ acc2[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) { cyc[0] = t1 - t0; cyc[1] = (long long)(g1 - g0); }
}

template <int OPA, int OPB, int RA, int RB> void run(const char *name, uint32_t *out, long long *cyc, int threads = 512)
{
    const int blocks = 148 * 2;      // 2 CTAs per SM
    bench<OPA, OPB, RA, RB><<<blocks, threads>>>(out, 12345u, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    bench<OPA, OPB, RA, RB><<<blocks, threads>>>(out, 12345u, cyc);
    cudaEventRecord(e1); cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long cc[2]; cudaMemcpy(cc, cyc, sizeof cc, cudaMemcpyDeviceToHost);
    long long c = cc[0];
    double winstr_per_sm = (double)ITERS * NACC * (RA + RB) * (2 * threads / 32);
    printf("%-22s w/SMSP=%d  %6.3f warp-instr/clk/SMSP (clock64)  %6.3f per ns/SMSP (event)  clk=%.0f MHz (clock64/globaltimer)  %.3f ms\n", name, 2 * threads / 128,
           winstr_per_sm / c / 4, winstr_per_sm / 4 / (ms * 1e6), 1e3 * c / (double)cc[1], ms);
}

int main()
{
    uint32_t *out; long long *cyc;
    cudaMalloc(&out, 148 * 2 * 512 * 4); cudaMalloc(&cyc, 16);
    for (int i = 0; i < 20; i++) bench<4, 7, 1, 1><<<296, 512>>>(out, 1u, cyc);   // warm the clocks
    cudaDeviceSynchronize();
    run<0, 11, 1, 0>("dp2a", out, cyc);
    run<1, 11, 1, 0>("dp4a", out, cyc);
    run<2, 11, 1, 0>("prmt", out, cyc);
    run<3, 11, 1, 0>("iadd", out, cyc);
    run<4, 11, 1, 0>("imad", out, cyc);
    run<5, 11, 1, 0>("lop3", out, cyc);
    run<6, 11, 1, 0>("shf", out, cyc);
    run<7, 11, 1, 0>("ffma", out, cyc);
    run<8, 11, 1, 0>("fadd", out, cyc);
    run<9, 11, 1, 0>("vadd2", out, cyc);
    run<0, 2, 1, 1>("dp2a+prmt", out, cyc);
    run<0, 3, 1, 1>("dp2a+iadd", out, cyc);
    run<0, 4, 1, 1>("dp2a+imad", out, cyc);
    run<0, 7, 1, 1>("dp2a+ffma", out, cyc);
    run<0, 8, 1, 1>("dp2a+fadd", out, cyc);
    run<1, 2, 1, 1>("dp4a+prmt", out, cyc);
    run<1, 4, 1, 1>("dp4a+imad", out, cyc);
    run<2, 3, 1, 1>("prmt+iadd", out, cyc);
    run<2, 4, 1, 1>("prmt+imad", out, cyc);
    run<3, 4, 1, 1>("iadd+imad", out, cyc);
    run<4, 7, 1, 1>("imad+ffma", out, cyc);
    run<3, 8, 1, 1>("iadd+fadd", out, cyc);
    run<2, 8, 1, 1>("prmt+fadd", out, cyc);
    run<0, 2, 2, 1>("2dp2a+prmt", out, cyc);
    run<0, 2, 1, 2>("dp2a+2prmt", out, cyc);
    run<4, 2, 1, 2>("imad+2prmt", out, cyc);
    run<12, 11, 1, 0>("imad.iadd", out, cyc);
    run<13, 11, 1, 0>("imad-4", out, cyc);
    run<0, 12, 1, 1>("dp2a+imad.iadd", out, cyc);
    run<0, 13, 1, 1>("dp2a+imad-4", out, cyc);
    run<2, 12, 1, 1>("prmt+imad.iadd", out, cyc);
    run<8, 12, 1, 1>("fadd+imad.iadd", out, cyc);
    run<0, 2, 1, 1>("dp2a+prmt", out, cyc, 256);
    run<0, 2, 1, 1>("dp2a+prmt", out, cyc, 128);
    run<0, 2, 1, 1>("dp2a+prmt", out, cyc, 64);
    run<0, 11, 1, 0>("dp2a", out, cyc, 128);
    run<0, 11, 1, 0>("dp2a", out, cyc, 64);
    run<2, 11, 1, 0>("prmt", out, cyc, 64);
    printf("err=%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
