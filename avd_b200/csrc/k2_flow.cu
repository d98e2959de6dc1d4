// K2: batched dense Farneback optical flow on 320x320 thumbnails and its magnitude statistics.
//
// Replaces cv2.calcOpticalFlowFarneback(prev, small, None, 0.5, 3, 15, 3, 5, 1.2, 0) (video.py:45) and
// np.mean/np.var of the magnitude (video.py:46-48).  Algorithm restated in SURVEY.md A.5:
// 4 pyramid levels (40,80,160,320), polynomial expansion n=5 sigma=1.2, 15x15 box window, 3 iterations.
//
// B200 design: the batch dimension is the image / pair index.  The per-IMAGE work (pyramid + polynomial
// expansion) is done once per thumbnail, not once per pair as the CPU library does (each thumbnail is
// "next" of one pair and "prev" of the following one).  All intermediates are planar float32 so every
// warp access is a contiguous 128-byte line.  At the finest level they are NOT L2-resident for a 119-pair batch (M alone is
// 244 MB against 126 MB of L2): measured DRAM traffic per kernel is in profiles/r02_k2_notes.md.
// Compiled with -fmad=false: the float sequences round exactly like the scalar library code.
#include "avd_internal.h"
#include <math.h>
#include <stdlib.h>
#include <stdio.h>
#include <string.h>

namespace {

constexpr int T = AVD_T;
constexpr int PYR_PX = 40 * 40 + 80 * 80 + 160 * 160 + 320 * 320;   // 136000 pixels over the 4 levels
__host__ __device__ constexpr int lvl_size(int k) { return T >> k; }
__host__ __device__ constexpr int lvl_off(int k) { return k == 3 ? 0 : k == 2 ? 1600 : k == 1 ? 8000 : 33600; }

struct ImgSrc {           // where image i of the batch lives
    const uint8_t *a, *b; // chain: a = thumbs, b = optional predecessor.  pairs: a = prev[], b = next[]
    size_t sa, sb;        // strides in bytes
    int mode;             // 0 chain without predecessor, 1 chain with predecessor, 2 explicit pairs
};
__device__ __forceinline__ const uint8_t *img_ptr(const ImgSrc &s, int i)
{
    if (s.mode == 0) return s.a + (size_t)i * s.sa;
    if (s.mode == 1) return i == 0 ? s.b : s.a + (size_t)(i - 1) * s.sa;
    return (i & 1) ? s.b + (size_t)(i >> 1) * s.sb : s.a + (size_t)(i >> 1) * s.sa;
}

struct FlowConst {
    float g[6], xg[6], xxg[6];          // polynomial-expansion Gaussian moments, index |k|
    double ig11, ig03, ig33, ig55;
    float k1[2], k2[5], k3[10];         // pyramid blur kernels: centre..edge for levels 1,2,3 (level 0 is .5/.25)
};
__constant__ FlowConst FC;

__device__ __forceinline__ int r101(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }
__device__ __forceinline__ int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// ---- pyramid -----------------------------------------------------------------------------------------
// GaussianBlur(ksize, sigma) of the full-resolution image (row pass then column pass, symmetric pairs
// summed first) evaluated only at the taps the following INTER_LINEAR / 2x2-area resize reads.
template <int R>
__device__ __forceinline__ float blur_row(const uint8_t *__restrict__ row, int x, const float *__restrict__ k)
{
    float a = k[0] * (float)row[x];
#pragma unroll
    for (int j = 1; j <= R; j++) a = a + k[j] * ((float)row[r101(x - j, T)] + (float)row[r101(x + j, T)]);
    return a;
}

// A CTA owns PYR_CW output columns of one image over the full height.  Phase A: the row-blurred values of the source
// columns those outputs read, for all 320 source rows, go to shared memory once (the previous version recomputed them for
// every output: 2-3x the work).  Phase B: column blur at the taps of each output and the resize arithmetic.
constexpr int PYR_CW = 8;
#ifndef PYR_MINB
#define PYR_MINB 1
#endif
template <int K>
__global__ void __launch_bounds__(256, PYR_MINB) pyr_kernel(ImgSrc src, float *__restrict__ pyr)
{
    constexpr int S = lvl_size(K), ST = 1 << K, R = K == 0 ? 1 : (K == 1 ? 1 : (K == 2 ? 4 : 9)), N = K == 0 ? 1 : 2;
    constexpr int OFF = K == 2 ? 1 : (K == 3 ? 3 : 0), NC = PYR_CW * N;
    __shared__ float t[T][NC];
    const int img_i = blockIdx.y, xo0 = blockIdx.x * PYR_CW;
    const uint8_t *img = img_ptr(src, img_i);
    float *out = pyr + (size_t)img_i * PYR_PX + lvl_off(K);
    const float k0[2] = {0.5f, 0.25f};
    const float *kk = K == 0 ? k0 : (K == 1 ? FC.k1 : (K == 2 ? FC.k2 : FC.k3));
#ifndef PYR_UNROLL
#define PYR_UNROLL 1
#endif
    constexpr int NT = 256;                              // == blockDim.x: with a compile-time trip count the loops are 20 % faster (155 -> 125 us)
    constexpr int PYU = PYR_UNROLL;
#pragma unroll PYU
    for (int i = threadIdx.x; i < T * NC; i += NT) {
        const int y = i / NC, ci = i - y * NC;
        const int xc = ST * (xo0 + ci / N) + OFF + ci % N;
        t[y][ci] = blur_row<R>(img + (size_t)y * T, xc, kk);
    }
    __syncthreads();
#pragma unroll PYU
    for (int i = threadIdx.x; i < S * PYR_CW; i += NT) {
        const int y = i / PYR_CW, cx = i - y * PYR_CW;
        const int yc = ST * y + OFF;
        float o[N][N];
#pragma unroll
        for (int oy = 0; oy < N; oy++)
#pragma unroll
            for (int ox = 0; ox < N; ox++) {
                const int c = cx * N + ox, r = yc + oy;
                float a = kk[0] * t[r][c];
#pragma unroll
                for (int j = 1; j <= R; j++) a = a + kk[j] * (t[r101(r - j, T)][c] + t[r101(r + j, T)][c]);
                o[oy][ox] = a;
            }
        float v;
        if (K == 0) v = o[0][0];
        else if (K == 1) v = (((o[0][0] + o[0][N - 1]) + o[N - 1][0]) + o[N - 1][N - 1]) * 0.25f;     // 320 -> 160: exact 2x => 2x2 area mean
        else {                                                                                        // bilinear at the cell centre (weights .5/.5)
            const float r0 = o[0][0] * 0.5f + o[0][N - 1] * 0.5f, r1 = o[N - 1][0] * 0.5f + o[N - 1][N - 1] * 0.5f;
            v = r0 * 0.5f + r1 * 0.5f;
        }
        out[(size_t)y * S + xo0 + cx] = v;
    }
}

// ---- polynomial expansion ----------------------------------------------------------------------------
// R[img][c][y][x], c = 0..4, planar.  Vertical pass in float32, horizontal pass with float64 accumulators, in the
// library's operation order.  A CTA owns PE_ROWS full-width rows of one image; a thread owns a column:
// it reads its 18 input values once (coalesced across the CTA), produces the three vertical moments of its 8 rows
// out of registers, parks them in shared memory (with the replicated border columns) and then runs the horizontal
// pass of its 8 pixels out of shared memory.
constexpr int PE_ROWS = 8, PE_HALO = 5;

#ifndef PE_MINB
#define PE_MINB 4      /* 48 registers, 4 CTAs of 320 threads per SM: at 3 CTAs (62 registers) the kernel takes 272 instead of 207 us */
#endif
template <int K>
__global__ void __launch_bounds__(T, PE_MINB) polyexp_kernel(const float *__restrict__ pyr, float *__restrict__ Rall)
{
    constexpr int k = K;
    extern __shared__ float pe_v[];                                    // [3][PE_ROWS][s + 2*PE_HALO]
    constexpr int s = lvl_size(k), vw = s + 2 * PE_HALO;       // level as a compile-time constant: offsets become immediates
    const float *I = pyr + (size_t)blockIdx.y * PYR_PX + lvl_off(k);
    float *R = Rall + (size_t)blockIdx.y * PYR_PX * 5 + (size_t)lvl_off(k) * 5;
    const int x = threadIdx.x, y0 = blockIdx.x * PE_ROWS;
    if (x < s) {
        float w[PE_ROWS + 2 * PE_HALO];
#pragma unroll
        for (int i = 0; i < PE_ROWS + 2 * PE_HALO; i++) w[i] = I[(size_t)clampi(y0 - PE_HALO + i, 0, s - 1) * s + x];
#pragma unroll
        for (int r = 0; r < PE_ROWS; r++) {
            const float c0 = w[r + PE_HALO];
            float r0 = c0 * FC.g[0], r1 = 0.f, r2 = 0.f;
#pragma unroll
            for (int j = 1; j <= 5; j++) {
                const float a = w[r + PE_HALO - j], b = w[r + PE_HALO + j];
                const float p = a + b;
                r0 = r0 + FC.g[j] * p;
                r1 = r1 + FC.xg[j] * (b - a);
                r2 = r2 + FC.xxg[j] * p;
            }
            float *v0 = pe_v + (size_t)r * vw + x + PE_HALO, *v1 = v0 + (size_t)PE_ROWS * vw, *v2 = v1 + (size_t)PE_ROWS * vw;
            *v0 = r0; *v1 = r1; *v2 = r2;
            if (x == 0) {
#pragma unroll
                for (int h = 1; h <= PE_HALO; h++) { v0[-h] = r0; v1[-h] = r1; v2[-h] = r2; }
            }
            if (x == s - 1) {
#pragma unroll
                for (int h = 1; h <= PE_HALO; h++) { v0[h] = r0; v1[h] = r1; v2[h] = r2; }
            }
        }
    }
    __syncthreads();
    if (x >= s) return;
    // Horizontal pass: a thread owns TWO adjacent pixels (columns 2cp, 2cp+1) of four of the block's eight rows, so that the
    // 12-value window both pixels read from a row of moments comes in as six 64-bit shared-memory loads (33 32-bit loads
    // per pixel before: the kernel was throttled by the shared-memory instruction queue).  Per accumulator the operations
    // and their order are unchanged (each accumulator only ever sees one plane of moments).
    constexpr int plane = s * s;
    constexpr int half = s >> 1;
    const int cp = x % half, r0 = (x / half) * (PE_ROWS / 2);
#pragma unroll 1
    for (int rr = 0; rr < PE_ROWS / 2; rr++) {
        const int r = r0 + rr;
        double b1[2], b2[2] = {0, 0}, b3[2], b4[2] = {0, 0}, b5[2], b6[2] = {0, 0};
        float w[12];
        const float2 *src = (const float2 *)(pe_v + (size_t)r * vw + 2 * cp);          // window columns 2cp-5 .. 2cp+6
#pragma unroll
        for (int i = 0; i < 6; i++) { const float2 q = src[i]; w[2 * i] = q.x; w[2 * i + 1] = q.y; }
#pragma unroll
        for (int e = 0; e < 2; e++) {
            b1[e] = (double)(w[5 + e] * FC.g[0]);
#pragma unroll
            for (int j = 1; j <= 5; j++) {
                const float p0 = w[5 + e + j], m0 = w[5 + e - j];
                const double tg = (double)(p0 + m0);
                // tg and the coefficient are float values held in doubles: their product is exact (48 bits), so one fused
                // multiply-add rounds exactly like the library's multiply followed by add
                b1[e] = __fma_rn(tg, (double)FC.g[j], b1[e]);
                b4[e] = __fma_rn(tg, (double)FC.xxg[j], b4[e]);
                b2[e] += (double)((p0 - m0) * FC.xg[j]);
            }
        }
        src += (size_t)PE_ROWS * vw / 2;
#pragma unroll
        for (int i = 0; i < 6; i++) { const float2 q = src[i]; w[2 * i] = q.x; w[2 * i + 1] = q.y; }
#pragma unroll
        for (int e = 0; e < 2; e++) {
            b3[e] = (double)(w[5 + e] * FC.g[0]);
#pragma unroll
            for (int j = 1; j <= 5; j++) {
                const float p1 = w[5 + e + j], m1 = w[5 + e - j];
                b3[e] += (double)((p1 + m1) * FC.g[j]);
                b6[e] += (double)((p1 - m1) * FC.xg[j]);
            }
        }
        src += (size_t)PE_ROWS * vw / 2;
#pragma unroll
        for (int i = 0; i < 6; i++) { const float2 q = src[i]; w[2 * i] = q.x; w[2 * i + 1] = q.y; }
#pragma unroll
        for (int e = 0; e < 2; e++) {
            b5[e] = (double)(w[5 + e] * FC.g[0]);
#pragma unroll
            for (int j = 1; j <= 5; j++) b5[e] += (double)((w[5 + e + j] + w[5 + e - j]) * FC.g[j]);
        }
        float *o = R + (size_t)(y0 + r) * s + 2 * cp;
        *(float2 *)o = make_float2((float)(b3[0] * FC.ig11), (float)(b3[1] * FC.ig11));
        *(float2 *)(o + plane) = make_float2((float)(b2[0] * FC.ig11), (float)(b2[1] * FC.ig11));
        *(float2 *)(o + 2 * plane) = make_float2((float)(b1[0] * FC.ig03 + b5[0] * FC.ig33), (float)(b1[1] * FC.ig03 + b5[1] * FC.ig33));
        *(float2 *)(o + 3 * plane) = make_float2((float)(b1[0] * FC.ig03 + b4[0] * FC.ig33), (float)(b1[1] * FC.ig03 + b4[1] * FC.ig33));
        *(float2 *)(o + 4 * plane) = make_float2((float)(b6[0] * FC.ig55), (float)(b6[1] * FC.ig55));
    }
}

// ---- update matrices ---------------------------------------------------------------------------------
// FLOWSRC: 0 = zero flow (coarsest level), 1 = flow of this level, 2 = 2x bilinear upsample of the
// coarser level's flow times 2 (the library's resize(prevFlow)*1/pyr_scale, fused here).
__device__ __forceinline__ float up2_axis_w(int d, int ssz, int &i0, int &i1, bool clamp_w)
{
    // INTER_LINEAR 2x upsampling tap of destination index d: f = d/2 - .25
    float f = (float)((d + 0.5) * 0.5 - 0.5);
    int s = (int)floorf(f);
    f -= (float)s;
    if (clamp_w) {                       // x axis: weight clamped at the ends
        if (s < 0) { f = 0.f; s = 0; }
        if (s >= ssz - 1) { f = 0.f; s = ssz - 1; }
        i0 = s; i1 = min(s + 1, ssz - 1);
    } else {                             // y axis: rows clipped, weight kept
        i0 = clampi(s, 0, ssz - 1); i1 = clampi(s + 1, 0, ssz - 1);
    }
    return f;
}

__device__ __forceinline__ float up2_sample(const float *__restrict__ c, int cs, int x, int y)
{
    int x0, x1, y0, y1;
    float fx = up2_axis_w(x, cs, x0, x1, true), fy = up2_axis_w(y, cs, y0, y1, false);
    float ax0 = 1.f - fx, ay0 = 1.f - fy;
    bool edge = (x0 == cs - 1);          // library: beyond xmax the row pass is S[sx]*1
    float r0 = edge ? c[(size_t)y0 * cs + x0] * 1.f : c[(size_t)y0 * cs + x0] * ax0 + c[(size_t)y0 * cs + x1] * fx;
    float r1 = edge ? c[(size_t)y1 * cs + x0] * 1.f : c[(size_t)y1 * cs + x0] * ax0 + c[(size_t)y1 * cs + x1] * fx;
    return (r0 * ay0 + r1 * fy) * 2.f;
}

// ---- 15x15 box sums (float64) + 2x2 solve -------------------------------------------------------------------
// Strip geometry shared by the kernel and the host: a warp covers 128 columns (4 per lane); lanes 2..29 produce the 112
// output columns, lanes 0,1,30,31 are the 8-column halo (>= 7) on each side.
constexpr int BX_COLS = 112, BS_M = 7;
constexpr int BX_MIN_SEG = 8;
__host__ __device__ constexpr int bx_strips(int s) { return (s + BX_COLS - 1) / BX_COLS; }
__host__ __device__ constexpr int bx_segs(int s, int seg) { return (s + seg - 1) / seg; }

// ---- update matrices ---------------------------------------------------------------------------------
// The five entries of the update matrix of one pixel (the library's FarnebackUpdateMatrices) for the
// displacement (dx, dy): bilinear warp of R1, border scaling.  The level's size is a compile-time constant: every plane /
// row offset is an immediate of the load instruction instead of 64-bit integer arithmetic (round 1's runtime-level kernel
// spent ~40 % of its issue slots on addresses: 226 -> 167 lane-instructions per pixel).
template <int S>
__device__ __forceinline__ void um_pixel_c(const float *__restrict__ R0, const float *__restrict__ R1, int x, int y, float dx, float dy, float *m)
{
    constexpr int plane = S * S;
    const int o = y * S + x;
    float fx = (float)x + dx, fy = (float)y + dy;
    int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
    fx -= (float)x1; fy -= (float)y1;
    float r2, r3, r4, r5, r6;
    const float q0 = R0[o], q1 = R0[plane + o], q2 = R0[2 * plane + o], q3 = R0[3 * plane + o], q4 = R0[4 * plane + o];
    if ((unsigned)x1 < (unsigned)(S - 1) && (unsigned)y1 < (unsigned)(S - 1)) {
        const float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
        const float *b = R1 + (y1 * S + x1);
#define BIL(c) (((a00 * b[(c) * plane] + a01 * b[(c) * plane + 1]) + a10 * b[(c) * plane + S]) + a11 * b[(c) * plane + S + 1])
        r2 = BIL(0); r3 = BIL(1); r4 = BIL(2); r5 = BIL(3); r6 = BIL(4);
#undef BIL
        r4 = (q2 + r4) * 0.5f; r5 = (q3 + r5) * 0.5f; r6 = (q4 + r6) * 0.25f;
    } else {
        r2 = r3 = 0.f; r4 = q2; r5 = q3; r6 = q4 * 0.5f;
    }
    r2 = (q0 - r2) * 0.5f;
    r3 = (q1 - r3) * 0.5f;
    r2 = r2 + (r4 * dy + r6 * dx);
    r3 = r3 + (r6 * dy + r5 * dx);
    constexpr int B = 5;
    if ((unsigned)(x - B) >= (unsigned)(S - 2 * B) || (unsigned)(y - B) >= (unsigned)(S - 2 * B)) {
        const float border[5] = {0.14f, 0.14f, 0.4472f, 0.4472f, 0.4472f};
        const float sc = (x < B ? border[x] : 1.f) * (x >= S - B ? border[S - x - 1] : 1.f) *
                         (y < B ? border[y] : 1.f) * (y >= S - B ? border[S - y - 1] : 1.f);
        r2 *= sc; r3 *= sc; r4 *= sc; r5 *= sc; r6 *= sc;
    }
    m[0] = r4 * r4 + r6 * r6;
    m[1] = (r4 + r5) * r6;
    m[2] = r5 * r5 + r6 * r6;
    m[3] = r4 * r2 + r6 * r3;
    m[4] = r6 * r2 + r5 * r3;
}

#ifndef UM_PPT
#define UM_PPT 2      /* pixels per thread (rows y and y + 8 of a 32 x 16 block): more loads in flight per warp, -1 % on the pass */
#endif
template <int K, int FLOWSRC>
__global__ void __launch_bounds__(256) um_kernel(const float *__restrict__ Rall, const float *__restrict__ flow, const float *__restrict__ coarse,
                                                 float *__restrict__ Mall, int pair_img_mode, int p_base)
{
    constexpr int S = lvl_size(K), plane = S * S;
    const int x = blockIdx.x * 32 + threadIdx.x, p = blockIdx.z + p_base;
    if (x >= S) return;
    const int ia = pair_img_mode == 2 ? 2 * p : p;
    const float *R0 = Rall + (size_t)ia * PYR_PX * 5 + lvl_off(K) * 5;
    const float *R1 = R0 + (size_t)PYR_PX * 5;
    int ys[UM_PPT];
    float dx[UM_PPT], dy[UM_PPT];
#pragma unroll
    for (int i = 0; i < UM_PPT; i++) {
        const int y = blockIdx.y * 8 * UM_PPT + threadIdx.y + 8 * i;
        ys[i] = y;
        dx[i] = dy[i] = 0.f;
        if (y < S) {
            const int o = y * S + x;
            if (FLOWSRC == 1) {
                const float *f = flow + (size_t)blockIdx.z * 2 * plane + o;      // flow buffers are indexed by the pair's position in its chain
                dx[i] = f[0]; dy[i] = f[plane];
            } else if (FLOWSRC == 2) {
                constexpr int cs = S >> 1;
                const float *cf = coarse + (size_t)blockIdx.z * 2 * cs * cs;
                dx[i] = up2_sample(cf, cs, x, y); dy[i] = up2_sample(cf + cs * cs, cs, x, y);
            }
        }
    }
    float m[UM_PPT][5];
#pragma unroll
    for (int i = 0; i < UM_PPT; i++)
        if (ys[i] < S) um_pixel_c<S>(R0, R1, x, ys[i], dx[i], dy[i], m[i]);
#pragma unroll
    for (int i = 0; i < UM_PPT; i++)
        if (ys[i] < S) {
            float *M = Mall + (size_t)blockIdx.z * 5 * T * T + ys[i] * S + x;      // M is indexed by the pair's position in the (sub-)batch
            M[0] = m[i][0]; M[plane] = m[i][1]; M[2 * plane] = m[i][2]; M[3 * plane] = m[i][3]; M[4 * plane] = m[i][4];
        }
}

// 15x15 box sums + 2x2 solve, one CHANNEL per warp.
// A CTA of five warps owns a strip of 128 columns (112 outputs + 8 halo columns on each side) and a segment of rows of one
// pair; warp c marches down the segment keeping the vertical 15-row running sums of channel c only in float64 registers
// (4 columns per lane; run += entering row - leaving row: 2 adds per value instead of 15), assembles the horizontal
// 15-column window from warp shuffles of per-lane partial sums -- window(j) = [T(t-1)+T(t)+T(t+1)] + suffix_{t-2}(j+1) +
// prefix_{t+2}(j) -- and parks the four window sums of each lane in shared memory.  After one block barrier per TWO rows
// (double-buffered) 112 threads solve one pixel each from the five sums and write the flow rows with coalesced stores.
// float64 sums of <= 225 float32 values are exact to ~1e-16, so any summation order agrees with the library's.
// Against round 1's one-warp-does-all-five-channels kernel: 5x the warps for the same pixels with no extra halo (the
// small pyramid levels and small batches fill the GPU), 64 instead of 126 registers, a fifth of the serial float64 chain
// per row; dense flow fields bit-identical (scripts/k2_fieldsha.py).
// STATS: the last iteration of the finest level emits per-task sums of |flow| and |flow|^2 instead of the field.
constexpr int BCS_THREADS = 160;
#ifndef BCS_MINB
#define BCS_MINB 6      /* 64 registers, 6 CTAs of 160 threads per SM; unbounded the compiler takes 94 (4 CTAs): 1.475 -> 1.398 ms per 119 pairs */
#endif
template <int K, bool STATS>
__global__ void __launch_bounds__(BCS_THREADS, BCS_MINB) box_cs_kernel(const float *__restrict__ Mall, float *__restrict__ flow, int seg_rows, int n_pairs,
                                                            double *__restrict__ task_stats, int p_base)
{
    constexpr int S = lvl_size(K), STRIPS = bx_strips(S), plane = S * S;
    __shared__ __align__(16) double asum[2][2][5][128];            // [phase][row of the pair][channel][column]
    __shared__ double red[2][4];
    const int lane = threadIdx.x & 31, c = threadIdx.x >> 5;
    const int tasks_per_pair = STRIPS * bx_segs(S, seg_rows);
    const int task = blockIdx.x;
    const int pl = task / tasks_per_pair, tp = task - pl * tasks_per_pair, p = pl + p_base;
    const int strip = tp % STRIPS, seg = tp / STRIPS;
    const int xb = strip * BX_COLS - 8 + 4 * lane, xa = clampi(xb, 0, S - 4);
    const int y0 = seg * seg_rows, y1 = min(y0 + seg_rows, S), nout = y1 - y0;
    const float *M = Mall + (size_t)pl * 5 * T * T + c * plane + xa;
    // columns outside the image replicate the edge column: a lane left of the image loads columns 0..3 and uses column 0
    // for all four, a lane right of it loads S-4..S-1 and uses S-1 (the running sums of the other columns are never used)
    const bool lo_edge = xb < 0, hi_edge = xb >= S;
    auto row = [&](int r) { return *(const float4 *)(M + clampi(r, 0, S - 1) * S); };

    double r0 = 0.0, r1 = 0.0, r2 = 0.0, r3 = 0.0;
#pragma unroll 5
    for (int j = -BS_M; j <= BS_M; j++) {
        const float4 q = row(y0 + j);
        r0 += (double)q.x; r1 += (double)q.y; r2 += (double)q.z; r3 += (double)q.w;
    }
    const double scale = 1.0 / 225.0;
    // window sums of the current row for this warp's channel -> shared memory
    auto emit = [&](double *dst4) {
        double v0 = r0, v1 = r1, v2 = r2, v3 = r3;
        if (lo_edge) { v1 = v0; v2 = v0; v3 = v0; }
        if (hi_edge) { v0 = v3; v1 = v3; v2 = v3; }
        const double tot = (v0 + v1) + (v2 + v3);
        const double mid = __shfl_up_sync(0xffffffffu, tot, 1) + tot + __shfl_down_sync(0xffffffffu, tot, 1);
        const double l1 = __shfl_up_sync(0xffffffffu, v1, 2), l2 = __shfl_up_sync(0xffffffffu, v2, 2), l3 = __shfl_up_sync(0xffffffffu, v3, 2);
        const double n0 = __shfl_down_sync(0xffffffffu, v0, 2), n1 = __shfl_down_sync(0xffffffffu, v1, 2), n2 = __shfl_down_sync(0xffffffffu, v2, 2);
        const double s3 = l3, s2 = l2 + l3, s1 = l1 + s2;       // suffix sums of lane t-2 from local index 3,2,1
        const double p1 = n0, p2 = n0 + n1, p3 = p2 + n2;       // prefix sums of lane t+2 of 1,2,3 columns
        ((double2 *)dst4)[0] = make_double2((mid + s1) * scale, (mid + s2 + p1) * scale);
        ((double2 *)dst4)[1] = make_double2((mid + s3 + p2) * scale, (mid + p3) * scale);
    };
    // the window slides one row down: + entering row, - leaving row
    auto slide = [&](const float4 &a, const float4 &b) {
        r0 += (double)a.x - (double)b.x; r1 += (double)a.y - (double)b.y;
        r2 += (double)a.z - (double)b.z; r3 += (double)a.w - (double)b.w;
    };
    const int t = threadIdx.x, px = strip * BX_COLS + t;
    const bool solver = t < BX_COLS && px < S;
    float *fo = flow + (size_t)pl * 2 * plane + px;
    double sm = 0, sq = 0;
    auto solve = [&](const double (*a)[128], int y) {
        const double g11 = a[0][8 + t], g12 = a[1][8 + t], g22 = a[2][8 + t], h1 = a[3][8 + t], h2 = a[4][8 + t];
        const double idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3);
        const float fx = (float)((g11 * h2 - g12 * h1) * idet);
        const float fy = (float)((g22 * h1 - g12 * h2) * idet);
        if (STATS) {
            const double mag = (double)sqrtf(fx * fx + fy * fy);     // float32 magnitude like np.sqrt(fx**2+fy**2)
            sm += mag; sq += mag * mag;
        } else {                                                      // the statistics launch is the last of the chain: nobody reads its flow field
            fo[y * S] = fx;
            fo[plane + y * S] = fy;
        }
    };
    // two rows per block barrier; the rows that enter / leave the window (y -> y+1: rows y+8 and y-7) are fetched one
    // iteration ahead, so their latency hides behind the barrier and the solve
    float4 an = row(y0 + BS_M + 1), ao = row(y0 - BS_M), bn = row(y0 + BS_M + 2), bo = row(y0 - BS_M + 1);
    int ph = 0;
    for (int j = 0; j < nout; j += 2) {
        const int y = y0 + j;
        emit(&asum[ph][0][c][4 * lane]);
        slide(an, ao);
        emit(&asum[ph][1][c][4 * lane]);                 // row y+1 (unused when it lies past the segment)
        slide(bn, bo);
        an = row(y + BS_M + 3); ao = row(y - BS_M + 2); bn = row(y + BS_M + 4); bo = row(y - BS_M + 3);
        __syncthreads();
        if (solver) {
            solve(asum[ph][0], y);
            if (y + 1 < y1) solve(asum[ph][1], y + 1);
        }
        ph ^= 1;
    }
    if (STATS) {
        // deterministic: fixed-order warp reduction, then the four solver warps in order; one partial per task
        for (int o = 16; o; o >>= 1) { sm += __shfl_down_sync(0xffffffffu, sm, o); sq += __shfl_down_sync(0xffffffffu, sq, o); }
        if (lane == 0 && c < 4) { red[0][c] = sm; red[1][c] = sq; }
        __syncthreads();
        if (t == 0) {
            const size_t ti = (size_t)p * tasks_per_pair + tp;
            task_stats[2 * ti] = ((red[0][0] + red[0][1]) + red[0][2]) + red[0][3];
            task_stats[2 * ti + 1] = ((red[1][0] + red[1][1]) + red[1][2]) + red[1][3];
        }
    }
}

// mean / variance of |flow| from the per-tile partial sums (fixed order => bit-reproducible)
__global__ void magstats_finalize_kernel(const double *__restrict__ tile_stats, int tiles, avd_sample_rec *rec, int n_pairs)
{
    int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= n_pairs) return;
    double a = 0, b = 0;
    for (int t = 0; t < tiles; t++) { a += tile_stats[2 * ((size_t)p * tiles + t)]; b += tile_stats[2 * ((size_t)p * tiles + t) + 1]; }
    const double n = (double)(T * T), mean = a / n;
    double var = b / n - mean * mean;
    rec[p].flow_mean = (float)mean;
    rec[p].flow_var = (float)(var > 0 ? var : 0.0);
    rec[p].flags |= AVD_REC_HAS_FLOW;
}

__global__ void interleave_kernel(const float *__restrict__ flow, float *__restrict__ out, int n_pairs)
{
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    size_t total = (size_t)n_pairs * T * T;
    if (i >= total) return;
    size_t p = i / (T * T), o = i % (T * T);
    out[2 * i] = flow[p * 2 * T * T + o];
    out[2 * i + 1] = flow[p * 2 * T * T + T * T + o];
}

FlowConst make_consts()
{
    FlowConst c{};
    const int n = 5; const double sigma = 1.2;
    float g[11], *gc = g + n;
    double sum = 0;
    for (int x = -n; x <= n; x++) { gc[x] = (float)exp(-x * x / (2 * sigma * sigma)); sum += gc[x]; }
    sum = 1. / sum;
    for (int x = -n; x <= n; x++) gc[x] = (float)(gc[x] * sum);
    for (int x = 0; x <= n; x++) { c.g[x] = gc[x]; c.xg[x] = (float)(x * gc[x]); c.xxg[x] = (float)(x * x * gc[x]); }
    double a = 0, b = 0, cc = 0, d = 0;
    for (int y = -n; y <= n; y++)
        for (int x = -n; x <= n; x++) {
            double w = (double)gc[y] * gc[x];
            a += w; b += w * x * x; cc += w * x * x * x * x; d += w * x * x * y * y;
        }
    c.ig11 = 1.0 / b; c.ig55 = 1.0 / d;
    double det = a * (cc * cc - d * d) - b * (b * cc - b * d) + b * (b * d - b * cc);
    c.ig03 = -(b * cc - b * d) / det;
    c.ig33 = (a * cc - b * b) / det;
    auto gk = [](int ksize, double s, float *out) {   // centre..edge
        double t[32], sm = 0;
        for (int i = 0; i < ksize; i++) { double x = i - (ksize - 1) * 0.5; t[i] = exp(-x * x / (2 * s * s)); sm += t[i]; }
        for (int i = 0; i <= ksize / 2; i++) out[i] = (float)(t[ksize / 2 + i] / sm);
    };
    gk(3, 0.5, c.k1); gk(9, 1.5, c.k2); gk(19, 3.5, c.k3);
    return c;
}

}  // namespace

// Workspace per batch of NI images / NP pairs (floats):
//   pyr  NI*136000 | R  NI*136000*5 | M  NP*5*320*320 | flowA, flowB  NP*2*320*320 each | tile stats NP*100*2 doubles
constexpr int STAT_TILES_MAX = bx_strips(T) * bx_segs(T, BX_MIN_SEG);    // most tasks per pair at the finest level
static size_t k2_ws_bytes(int ni, int np)
{
    return sizeof(float) * ((size_t)ni * PYR_PX * 6 + (size_t)np * T * T * 9) + sizeof(double) * 2 * STAT_TILES_MAX * (size_t)np;
}

// Rows per box segment, per level (AVD_K2_SEG="s0,s1,s2,s3"); fixed per level -- never chosen per batch -- so that a pair's
// flow does not depend on the batch it is computed in.
static const int *k2_seg_rows()
{
    static int seg_rows[4] = {32, 24, 16, 8};
    static bool init = false;
    if (!init) {
        if (const char *e = getenv("AVD_K2_SEG")) sscanf(e, "%d,%d,%d,%d", &seg_rows[0], &seg_rows[1], &seg_rows[2], &seg_rows[3]);
        for (int &v : seg_rows) v = v < BX_MIN_SEG ? BX_MIN_SEG : v;
        init = true;
    }
    return seg_rows;
}

struct ChainArgs {
    float *R, *M, *flowA, *flowB;
    double *tile_stats;
    int mode;
    bool want_stats;
    const cudaEvent_t *level_ready;     // level_ready[k] != nullptr: the polynomial expansion of level k runs on another stream
};

// The flow iterations of pairs [q0, q0 + nq) of the batch on stream st: levels 3..0, per level the update matrices
// (first launch: flow upsampled from the coarser level) and three box-sum/solve + update rounds.
static int run_chain(avd_ctx *ctx, cudaStream_t st, const ChainArgs &A, int q0, int nq, cudaEvent_t after_first)
{
    // AVD_K2_LEVEL_NAMES=1: the per-kernel timing log (avd_profile_read) keeps the pyramid levels apart
    static const bool lvl_names = getenv("AVD_K2_LEVEL_NAMES") && atoi(getenv("AVD_K2_LEVEL_NAMES"));
    static const char *um_name[4] = {"update_matrices_L0", "update_matrices_L1", "update_matrices_L2", "update_matrices_L3"};
    static const char *box_name[4] = {"box_solve_L0", "box_solve_L1", "box_solve_L2", "box_solve_L3"};
    const int *seg_rows = k2_seg_rows();
    // M and the flow buffers are indexed by the pair's position in THIS chain: the chains of a split batch are at different
    // pyramid levels at the same time, and a level-k flow of pair p sits at p * 2 * s_k^2 -- only chain-private regions
    // (at the finest level they coincide with the batch layout p * 2 * T^2) keep them apart
    float *R = A.R, *M = A.M + (size_t)q0 * 5 * T * T, *cur = A.flowA + (size_t)q0 * 2 * T * T, *prev = A.flowB + (size_t)q0 * 2 * T * T;
    double *tile_stats = A.tile_stats;
    const int mode = A.mode;
    const bool want_stats = A.want_stats;
    for (int k = 3; k >= 0; k--) {
        const int s = lvl_size(k);
        const int seg = seg_rows[k];
        if (A.level_ready && A.level_ready[k]) AVD_CUDA(ctx, cudaStreamWaitEvent(st, A.level_ready[k], 0));
        dim3 gu((s + 31) / 32, (s + 8 * UM_PPT - 1) / (8 * UM_PPT), nq), bu(32, 8);
        const int tasks = bx_strips(s) * bx_segs(s, seg) * nq;
#define UM_LAUNCH(K, SRC, FLOW, COARSE) AVD_LAUNCH(ctx, (lvl_names ? um_name[K] : "update_matrices"), st, um_kernel<K, SRC><<<gu, bu, 0, st>>>(R, FLOW, COARSE, M, mode, q0))
#define BOX_LAUNCH(K, STATS) AVD_LAUNCH(ctx, (lvl_names ? box_name[K] : "box_solve"), st, box_cs_kernel<K, STATS><<<tasks, BCS_THREADS, 0, st>>>(M, cur, seg, nq, STATS ? tile_stats : nullptr, q0))
#define LEVEL(K)                                                                                  \
    {                                                                                             \
        if (K == 3) UM_LAUNCH(K, 0, nullptr, nullptr); else UM_LAUNCH(K, 2, nullptr, prev);       \
        if (K == 3 && after_first) AVD_CUDA(ctx, cudaEventRecord(after_first, st));               \
        for (int it = 0; it < 3; it++) {                                                          \
            if (K == 0 && it == 2 && want_stats) BOX_LAUNCH(K, true); else BOX_LAUNCH(K, false);  \
            if (it < 2) UM_LAUNCH(K, 1, cur, nullptr);                                            \
        }                                                                                         \
    }
        if (k == 3) LEVEL(3) else if (k == 2) LEVEL(2) else if (k == 1) LEVEL(1) else LEVEL(0)
#undef LEVEL
#undef BOX_LAUNCH
#undef UM_LAUNCH
        float *t = cur; cur = prev; prev = t;   // prev now holds this level's flow
    }
    return AVD_OK;
}

int k2_reserve(avd_ctx *ctx, int n_pairs)
{
    if (n_pairs <= 0) return AVD_OK;
    const int cap_np = n_pairs < 128 ? n_pairs : 128;
    return avd_ensure(ctx, &ctx->k2_ws, &ctx->k2_ws_cap, k2_ws_bytes(cap_np + 1, cap_np));
}

int k2_flow(avd_ctx *ctx, const uint8_t *d_prev, const uint8_t *d_next, size_t prev_stride, size_t next_stride,
            int n_pairs, bool shared_chain, avd_sample_rec *d_rec_first, float *d_flow_out, cudaStream_t st)
{
    // shared_chain: images are one chain; pair p = (image p, image p+1); d_prev = optional predecessor
    // thumbnail (may be NULL), d_next = thumbs.  Otherwise explicit pairs (d_prev[p], d_next[p]).
    static bool consts_done[64] = {};
    if (!consts_done[ctx->device & 63]) {
        FlowConst c = make_consts();
        AVD_CUDA(ctx, cudaMemcpyToSymbol(FC, &c, sizeof(c)));
        consts_done[ctx->device & 63] = true;
    }
    if (n_pairs <= 0) return AVD_OK;
    if (d_rec_first && d_flow_out) { ctx->err = "k2_flow: statistics and dense fields are separate calls"; return AVD_ERR_ARG; }
    const int MAXP = 128;   // pairs per batch
    int rc;
    int cap_np = n_pairs < MAXP ? n_pairs : MAXP, cap_ni = shared_chain ? cap_np + 1 : 2 * cap_np;
    if ((rc = avd_ensure(ctx, &ctx->k2_ws, &ctx->k2_ws_cap, k2_ws_bytes(cap_ni, cap_np)))) return rc;
    for (int p0 = 0; p0 < n_pairs; p0 += MAXP) {
        int np = n_pairs - p0 < MAXP ? n_pairs - p0 : MAXP;
        int ni = shared_chain ? np + 1 : 2 * np;
        float *pyr = (float *)ctx->k2_ws, *R = pyr + (size_t)ni * PYR_PX, *M = R + (size_t)ni * PYR_PX * 5;
        float *flowA = M + (size_t)np * 5 * T * T, *flowB = flowA + (size_t)np * 2 * T * T;
        double *tile_stats = (double *)(flowB + (size_t)np * 2 * T * T);   // 8-byte aligned: all counts above are even
        ImgSrc src{};
        int mode;
        if (shared_chain) {
            // chain image j of this batch = global chain image p0 + j
            if (d_prev && p0 == 0) { src.mode = 1; src.a = d_next; src.b = d_prev; src.sa = next_stride; }
            else { src.mode = 0; src.a = d_next + (size_t)(p0 - (d_prev ? 1 : 0)) * next_stride; src.sa = next_stride; }
            mode = 0;
        } else {
            src.mode = 2; src.a = d_prev + (size_t)p0 * prev_stride; src.b = d_next + (size_t)p0 * next_stride;
            src.sa = prev_stride; src.sb = next_stride; mode = 2;
        }
        // Per image: pyramid level k, then its polynomial expansion.  The flow chain starts at the coarsest level and its first
        // levels are small, latency-bound launches, so only level 3 is prepared in front of it; levels 2, 1, 0 (three quarters of
        // the per-image work) run on a side stream beside the chain's coarse levels, which waits for level k right before it
        // needs it (AVD_K2_IMG_OVERLAP=0: everything in front, as before).  Same kernels, same arithmetic.
        static const bool img_overlap = !(getenv("AVD_K2_IMG_OVERLAP") && !atoi(getenv("AVD_K2_IMG_OVERLAP")));
        if (img_overlap && !ctx->k2_img) {
            AVD_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->k2_img, cudaStreamNonBlocking));
            for (int i = 0; i < 5; i++) AVD_CUDA(ctx, cudaEventCreateWithFlags(&ctx->k2_img_ev[i], cudaEventDisableTiming));
        }
        cudaEvent_t level_ready[4] = {nullptr, nullptr, nullptr, nullptr};
#define PYR_LAUNCH(K, S_) AVD_LAUNCH(ctx, "pyr", S_, pyr_kernel<K><<<dim3(lvl_size(K) / PYR_CW, ni), 256, 0, S_>>>(src, pyr))
#define PE_LAUNCH(K, S_) AVD_LAUNCH(ctx, "polyexp", S_, polyexp_kernel<K><<<dim3(lvl_size(K) / PE_ROWS, ni), (lvl_size(K) + 31) / 32 * 32, 3 * PE_ROWS * (lvl_size(K) + 2 * PE_HALO) * sizeof(float), S_>>>(pyr, R))
        if (img_overlap) {
            cudaStream_t is = ctx->k2_img;
            AVD_CUDA(ctx, cudaEventRecord(ctx->k2_img_ev[4], st));                // the thumbnails (and the workspace) are ready
            AVD_CUDA(ctx, cudaStreamWaitEvent(is, ctx->k2_img_ev[4], 0));
            PYR_LAUNCH(3, st); PE_LAUNCH(3, st);
            PYR_LAUNCH(2, is); PE_LAUNCH(2, is); AVD_CUDA(ctx, cudaEventRecord(ctx->k2_img_ev[2], is));
            PYR_LAUNCH(1, is); PE_LAUNCH(1, is); AVD_CUDA(ctx, cudaEventRecord(ctx->k2_img_ev[1], is));
            PYR_LAUNCH(0, is); PE_LAUNCH(0, is); AVD_CUDA(ctx, cudaEventRecord(ctx->k2_img_ev[0], is));
            for (int k = 0; k < 3; k++) level_ready[k] = ctx->k2_img_ev[k];
        } else {
            PYR_LAUNCH(3, st); PYR_LAUNCH(2, st); PYR_LAUNCH(1, st); PYR_LAUNCH(0, st);
            PE_LAUNCH(3, st); PE_LAUNCH(2, st); PE_LAUNCH(1, st); PE_LAUNCH(0, st);
        }
#undef PYR_LAUNCH
#undef PE_LAUNCH
        // Flow iterations (run_chain): um_kernel<K,.> + box_cs_kernel<K,.> per level, 3 iterations each.
        // AVD_K2_SPLIT=2: the pairs of the batch are split into two halves whose chains run on two streams, the second one
        // launch behind the first, so that the memory-bound update_matrices of one half shares the SMs with the issue-bound
        // box_solve of the other.  A pair's arithmetic does not depend on which half it is in.
        static const int split_env = getenv("AVD_K2_SPLIT") ? atoi(getenv("AVD_K2_SPLIT")) : 2;
        const int split = split_env < 1 ? 1 : (split_env > 4 ? 4 : split_env);
        ChainArgs A{R, M, flowA, flowB, tile_stats, mode, d_rec_first != nullptr, level_ready};
        if (split >= 2 && np >= 16 * split) {      // measured: 119 pairs 1.52 -> 1.45 ms, 80: 1.08 -> 1.03, 40: 0.64 -> 0.63, 16: slower
            if (!ctx->k2_side[0]) {
                for (int i = 0; i < 3; i++) AVD_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->k2_side[i], cudaStreamNonBlocking));
                for (int i = 0; i < 6; i++) AVD_CUDA(ctx, cudaEventCreateWithFlags(&ctx->k2_ev[i], cudaEventDisableTiming));
            }
            // chain c runs pairs [c*np/split, (c+1)*np/split) on its own stream, one launch behind chain c-1
            for (int c = 0; c < split; c++) {
                const int q0 = (int)((long long)np * c / split), q1 = (int)((long long)np * (c + 1) / split);
                cudaStream_t cs = c == 0 ? st : ctx->k2_side[c - 1];
                if (c > 0) AVD_CUDA(ctx, cudaStreamWaitEvent(cs, ctx->k2_ev[c - 1], 0));
                if ((rc = run_chain(ctx, cs, A, q0, q1 - q0, c + 1 < split ? ctx->k2_ev[c] : nullptr))) return rc;
                if (c > 0) {
                    AVD_CUDA(ctx, cudaEventRecord(ctx->k2_ev[3 + c - 1], cs));
                    AVD_CUDA(ctx, cudaStreamWaitEvent(st, ctx->k2_ev[3 + c - 1], 0));
                }
            }
        } else if ((rc = run_chain(ctx, st, A, 0, np, nullptr))) return rc;
        float *prev = flowB;                       // levels 3..0 write flowA, flowB, flowA, flowB: the finest flow ends in flowB
        if (d_rec_first) {
            // chain: pair p0+j -> record of its "next" image
            AVD_LAUNCH(ctx, "magstats", st, magstats_finalize_kernel<<<(np + 127) / 128, 128, 0, st>>>(tile_stats, bx_strips(T) * bx_segs(T, k2_seg_rows()[0]), d_rec_first + p0, np));
        }
        if (d_flow_out) {
            size_t total = (size_t)np * T * T;
            AVD_LAUNCH(ctx, "interleave", st, interleave_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(prev, d_flow_out + (size_t)p0 * T * T * 2, np));
        }
    }
    return AVD_OK;
}
