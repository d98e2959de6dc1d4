// K2: batched dense Farneback optical flow on 320x320 thumbnails and its magnitude statistics.
//
// Replaces cv2.calcOpticalFlowFarneback(prev, small, None, 0.5, 3, 15, 3, 5, 1.2, 0) (video.py:45) and
// np.mean/np.var of the magnitude (video.py:46-48).  Algorithm restated in SURVEY.md A.5:
// 4 pyramid levels (40,80,160,320), polynomial expansion n=5 sigma=1.2, 15x15 box window, 3 iterations.
//
// B200 design: the batch dimension is the image / pair index.  The per-IMAGE work (pyramid + polynomial
// expansion) is done once per thumbnail, not once per pair as the CPU library does (each thumbnail is
// "next" of one pair and "prev" of the following one).  All intermediates are planar float32 so every
// warp access is a contiguous 128-byte line; they are sized to stay L2-resident per batch.
// Compiled with -fmad=false: the float sequences round exactly like the scalar library code.
#include "avd_internal.h"
#include <math.h>
#include <stdlib.h>
#include <stdio.h>

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
template <int K>
__global__ void __launch_bounds__(256) pyr_kernel(ImgSrc src, float *__restrict__ pyr)
{
    constexpr int S = lvl_size(K), ST = 1 << K, R = K == 0 ? 1 : (K == 1 ? 1 : (K == 2 ? 4 : 9)), N = K == 0 ? 1 : 2;
    constexpr int OFF = K == 2 ? 1 : (K == 3 ? 3 : 0), NC = PYR_CW * N;
    __shared__ float t[T][NC];
    const int img_i = blockIdx.y, xo0 = blockIdx.x * PYR_CW;
    const uint8_t *img = img_ptr(src, img_i);
    float *out = pyr + (size_t)img_i * PYR_PX + lvl_off(K);
    const float k0[2] = {0.5f, 0.25f};
    const float *kk = K == 0 ? k0 : (K == 1 ? FC.k1 : (K == 2 ? FC.k2 : FC.k3));
    for (int i = threadIdx.x; i < T * NC; i += blockDim.x) {
        const int y = i / NC, ci = i - y * NC;
        const int xc = ST * (xo0 + ci / N) + OFF + ci % N;
        t[y][ci] = blur_row<R>(img + (size_t)y * T, xc, kk);
    }
    __syncthreads();
    for (int i = threadIdx.x; i < S * PYR_CW; i += blockDim.x) {
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

__global__ void __launch_bounds__(T) polyexp_kernel(const float *__restrict__ pyr, float *__restrict__ Rall, int k)
{
    extern __shared__ float pe_v[];                                    // [3][PE_ROWS][s + 2*PE_HALO]
    const int s = lvl_size(k), vw = s + 2 * PE_HALO;
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
    const size_t plane = (size_t)s * s;
#pragma unroll 1
    for (int r = 0; r < PE_ROWS; r++) {
        const float *v0 = pe_v + (size_t)r * vw + x + PE_HALO, *v1 = v0 + (size_t)PE_ROWS * vw, *v2 = v1 + (size_t)PE_ROWS * vw;
        double b1 = (double)(v0[0] * FC.g[0]), b2 = 0, b3 = (double)(v1[0] * FC.g[0]), b4 = 0, b5 = (double)(v2[0] * FC.g[0]), b6 = 0;
#pragma unroll
        for (int j = 1; j <= 5; j++) {
            const float p0 = v0[j], m0 = v0[-j], p1 = v1[j], m1 = v1[-j], p2 = v2[j], m2 = v2[-j];
            const double tg = (double)(p0 + m0);
            b1 += tg * (double)FC.g[j];
            b4 += tg * (double)FC.xxg[j];
            b2 += (double)((p0 - m0) * FC.xg[j]);
            b3 += (double)((p1 + m1) * FC.g[j]);
            b6 += (double)((p1 - m1) * FC.xg[j]);
            b5 += (double)((p2 + m2) * FC.g[j]);
        }
        const size_t o = (size_t)(y0 + r) * s + x;
        R[o] = (float)(b3 * FC.ig11);
        R[plane + o] = (float)(b2 * FC.ig11);
        R[2 * plane + o] = (float)(b1 * FC.ig03 + b5 * FC.ig33);
        R[3 * plane + o] = (float)(b1 * FC.ig03 + b4 * FC.ig33);
        R[4 * plane + o] = (float)(b6 * FC.ig55);
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

// The five entries of the update matrix of one pixel (the library's FarnebackUpdateMatrices) for the
// displacement (dx, dy): bilinear warp of R1, border scaling.  
__device__ __forceinline__ void um_pixel(const float *__restrict__ R0, const float *__restrict__ R1, int s, int x, int y, float dx, float dy, float *m)
{
    const size_t plane = (size_t)s * s, o = (size_t)y * s + x;
    float fx = (float)x + dx, fy = (float)y + dy;
    int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
    fx -= (float)x1; fy -= (float)y1;
    float r2, r3, r4, r5, r6;
    float q0 = R0[o], q1 = R0[plane + o], q2 = R0[2 * plane + o], q3 = R0[3 * plane + o], q4 = R0[4 * plane + o];
    if ((unsigned)x1 < (unsigned)(s - 1) && (unsigned)y1 < (unsigned)(s - 1)) {
        float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
        size_t b = (size_t)y1 * s + x1;
#define BIL(c) (((a00 * R1[(c) * plane + b] + a01 * R1[(c) * plane + b + 1]) + a10 * R1[(c) * plane + b + s]) + a11 * R1[(c) * plane + b + s + 1])
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
    const int B = 5;
    if ((unsigned)(x - B) >= (unsigned)(s - 2 * B) || (unsigned)(y - B) >= (unsigned)(s - 2 * B)) {
        const float border[5] = {0.14f, 0.14f, 0.4472f, 0.4472f, 0.4472f};
        float sc = (x < B ? border[x] : 1.f) * (x >= s - B ? border[s - x - 1] : 1.f) *
                   (y < B ? border[y] : 1.f) * (y >= s - B ? border[s - y - 1] : 1.f);
        r2 *= sc; r3 *= sc; r4 *= sc; r5 *= sc; r6 *= sc;
    }
    m[0] = r4 * r4 + r6 * r6;
    m[1] = (r4 + r5) * r6;
    m[2] = r5 * r5 + r6 * r6;
    m[3] = r4 * r2 + r6 * r3;
    m[4] = r6 * r2 + r5 * r3;
}

template <int FLOWSRC>
__global__ void __launch_bounds__(256) update_matrices_kernel(const float *__restrict__ Rall, const float *__restrict__ flow,
                                                              const float *__restrict__ coarse, float *__restrict__ Mall,
                                                              float *__restrict__ flow_store, int k, int pair_img_mode, int p_base)
{
    const int s = lvl_size(k);
    int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y, p = blockIdx.z + p_base;
    if (x >= s || y >= s) return;
    int ia = pair_img_mode == 2 ? 2 * p : p, ib = ia + 1;
    size_t plane = (size_t)s * s, o = (size_t)y * s + x;
    const float *R0 = Rall + (size_t)ia * PYR_PX * 5 + (size_t)lvl_off(k) * 5;
    const float *R1 = Rall + (size_t)ib * PYR_PX * 5 + (size_t)lvl_off(k) * 5;
    float dx = 0.f, dy = 0.f;
    if (FLOWSRC == 1) {
        dx = flow[(size_t)p * 2 * plane + o]; dy = flow[(size_t)p * 2 * plane + plane + o];
    } else if (FLOWSRC == 2) {
        int cs = s >> 1;
        const float *cf = coarse + (size_t)p * 2 * cs * cs;
        dx = up2_sample(cf, cs, x, y); dy = up2_sample(cf + (size_t)cs * cs, cs, x, y);
    }
    float m[5];
    um_pixel(R0, R1, s, x, y, dx, dy, m);
    float *M = Mall + (size_t)p * 5 * T * T;
    M[o] = m[0]; M[plane + o] = m[1]; M[2 * plane + o] = m[2]; M[3 * plane + o] = m[3]; M[4 * plane + o] = m[4];
    (void)flow_store;
}

// ---- 15x15 box sums (float64) + 2x2 solve -------------------------------------------------------------
// Warp-autonomous marching kernel: no shared memory, no block barriers.
// A warp owns a strip of 128 columns (4 per lane; lanes 2..29 produce the 112 output columns, lanes 0,1,30,31
// are the 8-column halo >= 7) and a segment of BX_SEG rows.  It marches down the segment keeping, per lane,
// the vertical 15-row running sums of its 4 columns x 5 channels in float64 registers (run += new row - old
// row: 2 adds per pixel-channel instead of 15).  The horizontal 15-column window is assembled from warp
// shuffles of per-lane partial sums: window(j) = [T(t-1)+T(t)+T(t+1)] + suffix_{t-2}(j+1) + prefix_{t+2}(j).
// float64 sums of <=225 float32 values are exact to ~1e-16, so any summation order agrees with the library's.
// STATS: the last iteration of the finest level also emits per-task sums of |flow| and |flow|^2.
// The segment length trades redundant window start-up (15 rows per segment) for parallelism: the small pyramid levels
// are latency-bound (a few hundred warps for the whole GPU), so they use short segments.
constexpr int BX_COLS = 112, BX_WARPS = 4, BS_M = 7;
constexpr int BX_MIN_SEG = 8;
__host__ __device__ constexpr int bx_strips(int s) { return (s + BX_COLS - 1) / BX_COLS; }
__host__ __device__ constexpr int bx_segs(int s, int seg) { return (s + seg - 1) / seg; }

__device__ __forceinline__ float4 ld_row4(const float *__restrict__ rowp, int xa, int xb, int s)
{
    // 4 consecutive columns starting at xb (a multiple of 4); columns outside [0,s) replicate the edge column
    float4 q;                                           // xa = clamp(xb, 0, s-4): always in range and 16-byte aligned
#ifndef K2_VEC_ASM
#define K2_VEC_ASM 0   /* measured: hand-placed ld.global.v4 doubles box_solve time (the compiler's own loads are software-pipelined) */
#endif
#if K2_VEC_ASM
    asm("ld.global.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(q.x), "=f"(q.y), "=f"(q.z), "=f"(q.w) : "l"(rowp + xa));
#else
    q = *(const float4 *)(rowp + xa);
#endif
    if (xb < 0) q = make_float4(q.x, q.x, q.x, q.x);
    else if (xb >= s) q = make_float4(q.w, q.w, q.w, q.w);
    return q;
}

template <bool STATS>
__global__ void __launch_bounds__(32 * BX_WARPS) box_solve_kernel(const float *__restrict__ Mall, float *__restrict__ flow, int k, int BX_SEG,
                                                                 int n_pairs, double *__restrict__ task_stats, int p_base)
{
    const int s = lvl_size(k);
    const int lane = threadIdx.x & 31;
    const int tasks_per_pair = bx_strips(s) * bx_segs(s, BX_SEG);
    const int task = blockIdx.x * BX_WARPS + (threadIdx.x >> 5);
    if (task >= tasks_per_pair * n_pairs) return;
    const int pl = task / tasks_per_pair, tp = task - pl * tasks_per_pair, p = pl + p_base;
    const int strip = tp % bx_strips(s), seg = tp / bx_strips(s);
    const int xb = strip * BX_COLS - 8 + 4 * lane, xa = clampi(xb, 0, s - 4);
    const int y0 = seg * BX_SEG, y1 = min(y0 + BX_SEG, s);
    const size_t plane = (size_t)s * s;
    const float *M = Mall + (size_t)p * 5 * T * T;
    const bool out_lane = lane >= 2 && lane < 30 && xb >= 0 && xb < s;

    double run[5][4];
#pragma unroll
    for (int c = 0; c < 5; c++) { run[c][0] = run[c][1] = run[c][2] = run[c][3] = 0.0; }
    for (int j = -BS_M; j <= BS_M; j++) {
        const float *rowp = M + (size_t)clampi(y0 + j, 0, s - 1) * s;
#pragma unroll
        for (int c = 0; c < 5; c++) {
            float4 q = ld_row4(rowp + c * plane, xa, xb, s);
            run[c][0] += (double)q.x; run[c][1] += (double)q.y; run[c][2] += (double)q.z; run[c][3] += (double)q.w;
        }
    }
    const double scale = 1.0 / 225.0;
    double sm = 0, sq = 0;
    for (int y = y0; y < y1; y++) {
        double a[5][4];
#pragma unroll
        for (int c = 0; c < 5; c++) {
            const double v0 = run[c][0], v1 = run[c][1], v2 = run[c][2], v3 = run[c][3];
            const double tot = (v0 + v1) + (v2 + v3);
            const double mid = __shfl_up_sync(0xffffffffu, tot, 1) + tot + __shfl_down_sync(0xffffffffu, tot, 1);
            const double l1 = __shfl_up_sync(0xffffffffu, v1, 2), l2 = __shfl_up_sync(0xffffffffu, v2, 2), l3 = __shfl_up_sync(0xffffffffu, v3, 2);
            const double r0 = __shfl_down_sync(0xffffffffu, v0, 2), r1 = __shfl_down_sync(0xffffffffu, v1, 2), r2 = __shfl_down_sync(0xffffffffu, v2, 2);
            const double s3 = l3, s2 = l2 + l3, s1 = l1 + s2;       // suffix sums of lane t-2 from local index 3,2,1
            const double p1 = r0, p2 = r0 + r1, p3 = p2 + r2;       // prefix sums of lane t+2 of 1,2,3 columns
            a[c][0] = (mid + s1) * scale;
            a[c][1] = (mid + s2 + p1) * scale;
            a[c][2] = (mid + s3 + p2) * scale;
            a[c][3] = (mid + p3) * scale;
        }
        if (out_lane) {
            float fx[4], fy[4];
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const double g11 = a[0][j], g12 = a[1][j], g22 = a[2][j], h1 = a[3][j], h2 = a[4][j];
                const double idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3);
                fx[j] = (float)((g11 * h2 - g12 * h1) * idet);
                fy[j] = (float)((g22 * h1 - g12 * h2) * idet);
                if (STATS) {
                    const double mag = (double)sqrtf(fx[j] * fx[j] + fy[j] * fy[j]);   // float32 magnitude like np.sqrt(fx**2+fy**2)
                    sm += mag; sq += mag * mag;
                }
            }
            if (!STATS) {        // the statistics launch is the last of the chain: nobody reads its flow field
                float *fo = flow + (size_t)p * 2 * plane + (size_t)y * s + xb;
                *(float4 *)fo = make_float4(fx[0], fx[1], fx[2], fx[3]);
                *(float4 *)(fo + plane) = make_float4(fy[0], fy[1], fy[2], fy[3]);
            }
        }
        // slide the vertical window: + row y+8, - row y-7 (replicated borders)
        const float *rn = M + (size_t)min(y + BS_M + 1, s - 1) * s, *ro = M + (size_t)max(y - BS_M, 0) * s;
#pragma unroll
        for (int c = 0; c < 5; c++) {
            const float4 qn = ld_row4(rn + c * plane, xa, xb, s), qo = ld_row4(ro + c * plane, xa, xb, s);
            run[c][0] += (double)qn.x - (double)qo.x; run[c][1] += (double)qn.y - (double)qo.y;
            run[c][2] += (double)qn.z - (double)qo.z; run[c][3] += (double)qn.w - (double)qo.w;
        }
    }
    if (STATS) {
        // deterministic: fixed-order warp reduction, one partial per task, summed in task order afterwards
        for (int o = 16; o; o >>= 1) { sm += __shfl_down_sync(0xffffffffu, sm, o); sq += __shfl_down_sync(0xffffffffu, sq, o); }
        if (lane == 0) { size_t t = (size_t)p * tasks_per_pair + tp; task_stats[2 * t] = sm; task_stats[2 * t + 1] = sq; }
    }
}

// ---- box_solve with an asynchronous row ring ---------------------------------------------------------------
// Same arithmetic as box_solve_kernel, different data movement: every lane copies the rows it is going to need into its
// own shared-memory ring with cp.async (LDGSTS), BXR-1 steps ahead, instead of holding them in registers; a lane only
// ever reads what it copied itself, so the only synchronisation is cp.async.wait_group.  Measured (119 pairs, all 12
// launches): direct loads 878 us, ring depth 2: 851 us, 3: 917 us, 4: 948 us, 6: 1379 us -- the kernel lives on resident
// warps, not on prefetch distance, so the shallowest ring (40 KB per CTA, 4 CTAs per SM) is the default.
// Step i < 15 brings window row y0-7+i; step 15+j brings the row that enters (y0+8+j) and the row that leaves (y0-7+j).
#ifndef K2_BXR
#define K2_BXR 2
#endif
constexpr int BXR = K2_BXR;                                   // ring depth (steps)
constexpr int BXR_FLOAT4_PER_WARP = 2 * BXR * 5 * 32;    // 2 streams x ring x 5 channels x 32 lanes
__device__ __forceinline__ void cp_async16(float4 *dst, const float *src)
{
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(dst)), "l"(src) : "memory");
}
__device__ __forceinline__ float4 edge_fix(float4 q, int xb, int s)
{
    if (xb < 0) q = make_float4(q.x, q.x, q.x, q.x);
    else if (xb >= s) q = make_float4(q.w, q.w, q.w, q.w);
    return q;
}

template <bool STATS>
__global__ void __launch_bounds__(32 * BX_WARPS, STATS ? 4 : 0) box_solve_ring_kernel(const float *__restrict__ Mall, float *__restrict__ flow, int k, int BX_SEG,
                                                                      int n_pairs, double *__restrict__ task_stats, int p_base)
{
    extern __shared__ float4 bx_ring[];
    const int s = lvl_size(k);
    const int lane = threadIdx.x & 31;
    const int tasks_per_pair = bx_strips(s) * bx_segs(s, BX_SEG);
    const int task = blockIdx.x * BX_WARPS + (threadIdx.x >> 5);
    if (task >= tasks_per_pair * n_pairs) return;
    const int pl = task / tasks_per_pair, tp = task - pl * tasks_per_pair, p = pl + p_base;
    const int strip = tp % bx_strips(s), seg = tp / bx_strips(s);
    const int xb = strip * BX_COLS - 8 + 4 * lane, xa = clampi(xb, 0, s - 4);
    const int y0 = seg * BX_SEG, y1 = min(y0 + BX_SEG, s), nout = y1 - y0;
    const size_t plane = (size_t)s * s;
    const float *M = Mall + (size_t)p * 5 * T * T + xa;
    const bool out_lane = lane >= 2 && lane < 30 && xb >= 0 && xb < s;
    float4 *ring = bx_ring + (threadIdx.x >> 5) * BXR_FLOAT4_PER_WARP + lane;
    const int nsteps = 15 + nout - 1;

    auto issue = [&](int i) {
        if (i < nsteps) {
            float4 *slot = ring + (i % BXR) * 5 * 32;
            const float *ra = M + (size_t)clampi(y0 - BS_M + i, 0, s - 1) * s;
#pragma unroll
            for (int c = 0; c < 5; c++) cp_async16(slot + c * 32, ra + c * plane);
            if (i >= 15) {
                const float *rs = M + (size_t)max(y0 - BS_M + i - 15, 0) * s;
#pragma unroll
                for (int c = 0; c < 5; c++) cp_async16(slot + (BXR * 5 + c) * 32, rs + c * plane);
            }
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int i = 0; i < BXR - 1; i++) issue(i);

    double run[5][4];
#pragma unroll
    for (int c = 0; c < 5; c++) { run[c][0] = run[c][1] = run[c][2] = run[c][3] = 0.0; }
    for (int i = 0; i < 15; i++) {
        issue(i + BXR - 1);
        asm volatile("cp.async.wait_group %0;" ::"n"(BXR - 1) : "memory");
        const float4 *slot = ring + (i % BXR) * 5 * 32;
#pragma unroll
        for (int c = 0; c < 5; c++) {
            const float4 q = edge_fix(slot[c * 32], xb, s);
            run[c][0] += (double)q.x; run[c][1] += (double)q.y; run[c][2] += (double)q.z; run[c][3] += (double)q.w;
        }
    }
    const double scale = 1.0 / 225.0;
    double sm = 0, sq = 0;
    for (int j = 0; j < nout; j++) {
        const int y = y0 + j;
        double a[5][4];
#pragma unroll
        for (int c = 0; c < 5; c++) {
            const double v0 = run[c][0], v1 = run[c][1], v2 = run[c][2], v3 = run[c][3];
            const double tot = (v0 + v1) + (v2 + v3);
            const double mid = __shfl_up_sync(0xffffffffu, tot, 1) + tot + __shfl_down_sync(0xffffffffu, tot, 1);
            const double l1 = __shfl_up_sync(0xffffffffu, v1, 2), l2 = __shfl_up_sync(0xffffffffu, v2, 2), l3 = __shfl_up_sync(0xffffffffu, v3, 2);
            const double r0 = __shfl_down_sync(0xffffffffu, v0, 2), r1 = __shfl_down_sync(0xffffffffu, v1, 2), r2 = __shfl_down_sync(0xffffffffu, v2, 2);
            const double s3 = l3, s2 = l2 + l3, s1 = l1 + s2;       // suffix sums of lane t-2 from local index 3,2,1
            const double p1 = r0, p2 = r0 + r1, p3 = p2 + r2;       // prefix sums of lane t+2 of 1,2,3 columns
            a[c][0] = (mid + s1) * scale;
            a[c][1] = (mid + s2 + p1) * scale;
            a[c][2] = (mid + s3 + p2) * scale;
            a[c][3] = (mid + p3) * scale;
        }
        if (out_lane) {
            float fx[4], fy[4];
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
                const double g11 = a[0][jj], g12 = a[1][jj], g22 = a[2][jj], h1 = a[3][jj], h2 = a[4][jj];
                const double idet = 1.0 / (g11 * g22 - g12 * g12 + 1e-3);
                fx[jj] = (float)((g11 * h2 - g12 * h1) * idet);
                fy[jj] = (float)((g22 * h1 - g12 * h2) * idet);
                if (STATS) {
                    const double mag = (double)sqrtf(fx[jj] * fx[jj] + fy[jj] * fy[jj]);   // float32 magnitude like np.sqrt(fx**2+fy**2)
                    sm += mag; sq += mag * mag;
                }
            }
            if (!STATS) {        // the statistics launch is the last of the chain: nobody reads its flow field
                float *fo = flow + (size_t)p * 2 * plane + (size_t)y * s + xb;
                *(float4 *)fo = make_float4(fx[0], fx[1], fx[2], fx[3]);
                *(float4 *)(fo + plane) = make_float4(fy[0], fy[1], fy[2], fy[3]);
            }
        }
        if (j + 1 < nout) {
            // slide the vertical window: + row y+8, - row y-7 (replicated borders)
            const int i = 15 + j;
            issue(i + BXR - 1);
            asm volatile("cp.async.wait_group %0;" ::"n"(BXR - 1) : "memory");
            const float4 *slot = ring + (i % BXR) * 5 * 32;
#pragma unroll
            for (int c = 0; c < 5; c++) {
                const float4 qn = edge_fix(slot[c * 32], xb, s), qo = edge_fix(slot[(BXR * 5 + c) * 32], xb, s);
                run[c][0] += (double)qn.x - (double)qo.x; run[c][1] += (double)qn.y - (double)qo.y;
                run[c][2] += (double)qn.z - (double)qo.z; run[c][3] += (double)qn.w - (double)qo.w;
            }
        }
    }
    if (STATS) {
        for (int o = 16; o; o >>= 1) { sm += __shfl_down_sync(0xffffffffu, sm, o); sq += __shfl_down_sync(0xffffffffu, sq, o); }
        if (lane == 0) { size_t t = (size_t)p * tasks_per_pair + tp; task_stats[2 * t] = sm; task_stats[2 * t + 1] = sq; }
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
        AVD_LAUNCH(ctx, "pyr", st, pyr_kernel<3><<<dim3(40 / PYR_CW, ni), 256, 0, st>>>(src, pyr));
        AVD_LAUNCH(ctx, "pyr", st, pyr_kernel<2><<<dim3(80 / PYR_CW, ni), 256, 0, st>>>(src, pyr));
        AVD_LAUNCH(ctx, "pyr", st, pyr_kernel<1><<<dim3(160 / PYR_CW, ni), 256, 0, st>>>(src, pyr));
        AVD_LAUNCH(ctx, "pyr", st, pyr_kernel<0><<<dim3(320 / PYR_CW, ni), 256, 0, st>>>(src, pyr));
        for (int k = 3; k >= 0; k--) {
            int s = lvl_size(k);
            AVD_LAUNCH(ctx, "polyexp", st, polyexp_kernel<<<dim3(s / PE_ROWS, ni), (s + 31) / 32 * 32, 3 * PE_ROWS * (s + 2 * PE_HALO) * sizeof(float), st>>>(pyr, R, k));
        }
        float *cur = flowA, *prev = flowB;
        // Optional L2-resident sub-batching of levels 0/1 (AVD_K2_SB0 / AVD_K2_SB1 = pairs per sub-batch, 0 = off).
        // Measured on B200 (profiles/r01_k2_notes.md): OFF is faster -- box_solve has only 30 warps of work per pair,
        // so a 16-pair sub-batch leaves most SMs idle (box_solve 1.06 -> 4.1 ms, update_matrices 0.69 -> 0.84 ms).
        // rows per box_solve segment, per level (AVD_K2_SEG="s0,s1,s2,s3" overrides; multiples of 8)
        // 56 / 24 rows at levels 0 / 1: with 119 pairs (BASELINE's 120-sample clip) the launch is 536 / 417 CTAs, one wave of the
        // 592 resident CTAs; 64 rows left a quarter of the slots empty at level 0 and 16 rows made level 1 595 CTAs, i.e. two waves.
        // Fixed per level (not chosen per batch) so that a pair's flow never depends on the batch it is computed in.
        static int seg_rows[4] = {56, 24, 8, 8};
        static bool seg_init = false;
        if (!seg_init) {
            if (const char *e = getenv("AVD_K2_SEG")) sscanf(e, "%d,%d,%d,%d", &seg_rows[0], &seg_rows[1], &seg_rows[2], &seg_rows[3]);
            for (int &v : seg_rows) v = v < BX_MIN_SEG ? BX_MIN_SEG : v;
            seg_init = true;
        }
        static const bool box_ring = !(getenv("AVD_K2_BOXRING") && atoi(getenv("AVD_K2_BOXRING")) == 0);
        static bool ring_attr[64] = {};
        if (box_ring && !ring_attr[ctx->device & 63]) {
            const int bytes = (int)(BX_WARPS * BXR_FLOAT4_PER_WARP * sizeof(float4));
            AVD_CUDA(ctx, cudaFuncSetAttribute(box_solve_ring_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
            AVD_CUDA(ctx, cudaFuncSetAttribute(box_solve_ring_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
            ring_attr[ctx->device & 63] = true;
        }
        static int sb0 = getenv("AVD_K2_SB0") ? atoi(getenv("AVD_K2_SB0")) : 0, sb1 = getenv("AVD_K2_SB1") ? atoi(getenv("AVD_K2_SB1")) : 0;
        for (int k = 3; k >= 0; k--) {
            const int s = lvl_size(k);
            const int SB = k == 0 ? (sb0 > 0 ? sb0 : np) : k == 1 ? (sb1 > 0 ? sb1 : np) : np;
            for (int q0 = 0; q0 < np; q0 += SB) {
                const int nq = np - q0 < SB ? np - q0 : SB;
                dim3 gu((s + 31) / 32, (s + 7) / 8, nq), bu(32, 8);
                const int seg = seg_rows[k];
                const int bx_blocks = (bx_strips(s) * bx_segs(s, seg) * nq + BX_WARPS - 1) / BX_WARPS;
                if (k == 3) {
                    AVD_LAUNCH(ctx, "update_matrices", st, update_matrices_kernel<0><<<gu, bu, 0, st>>>(R, nullptr, nullptr, M, nullptr, k, mode, q0));
                } else {
                    AVD_LAUNCH(ctx, "update_matrices", st, update_matrices_kernel<2><<<gu, bu, 0, st>>>(R, nullptr, prev, M, nullptr, k, mode, q0));
                }
                for (int it = 0; it < 3; it++) {
                    const size_t ring_bytes = (size_t)BX_WARPS * BXR_FLOAT4_PER_WARP * sizeof(float4);
                    if (box_ring) {
                        if (k == 0 && it == 2 && d_rec_first) {
                            AVD_LAUNCH(ctx, "box_solve", st, box_solve_ring_kernel<true><<<bx_blocks, 32 * BX_WARPS, ring_bytes, st>>>(M, cur, k, seg, nq, tile_stats, q0));
                        } else {
                            AVD_LAUNCH(ctx, "box_solve", st, box_solve_ring_kernel<false><<<bx_blocks, 32 * BX_WARPS, ring_bytes, st>>>(M, cur, k, seg, nq, nullptr, q0));
                        }
                    } else if (k == 0 && it == 2 && d_rec_first) {
                        AVD_LAUNCH(ctx, "box_solve", st, box_solve_kernel<true><<<bx_blocks, 32 * BX_WARPS, 0, st>>>(M, cur, k, seg, nq, tile_stats, q0));
                    } else {
                        AVD_LAUNCH(ctx, "box_solve", st, box_solve_kernel<false><<<bx_blocks, 32 * BX_WARPS, 0, st>>>(M, cur, k, seg, nq, nullptr, q0));
                    }
                    if (it < 2) {
                        AVD_LAUNCH(ctx, "update_matrices", st, update_matrices_kernel<1><<<gu, bu, 0, st>>>(R, cur, nullptr, M, nullptr, k, mode, q0));
                    }
                }
            }
            float *t = cur; cur = prev; prev = t;   // prev now holds this level's flow
        }
        if (d_rec_first) {
            // chain: pair p0+j -> record of its "next" image
            AVD_LAUNCH(ctx, "magstats", st, magstats_finalize_kernel<<<(np + 127) / 128, 128, 0, st>>>(tile_stats, bx_strips(T) * bx_segs(T, seg_rows[0]), d_rec_first + p0, np));
        }
        if (d_flow_out) {
            size_t total = (size_t)np * T * T;
            AVD_LAUNCH(ctx, "interleave", st, interleave_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(prev, d_flow_out + (size_t)p0 * T * T * 2, np));
        }
    }
    return AVD_OK;
}
