// K4: per-window features of the audio analyzer (include/avd.h: avd_audio_windows).
//
// The reference cuts the 16 kHz mono signal into 0.5 s windows (8000 samples, the last one ragged) and computes per
// window (audio.py:41-61): RMS and zero-crossing rate in float32, then in float64 the Hann-windowed real spectrum,
// its flatness exp(mean(log|X|))/mean|X|, the 85 % roll-off bin (a sequential scan) and the centroid on a [0,1] axis.
//
//   spectrum2_kernel  full windows whose length factors as N1 x N2 (8000 = 80 x 100): one CTA per window, two stages of short
//                     direct DFTs out of shared memory (N (N1 + N2) multiply-adds instead of N^2 / 2).
//   spectrum_kernel   the ragged last window (any length, also prime), or every window when win does not factor:
//                     CTA = (window, block of 512 bins).  The windowed samples (float64) and the table of the len-th
//                     roots of unity (one sincospi per entry, so every twiddle is correctly rounded) sit in shared
//                     memory; a thread owns two bins and walks the sample pairs (x[i], x[len-i]), advancing the angle by
//                     rotation and re-reading it from the table every 64 terms -- a direct DFT.  The last window has
//                     an arbitrary (possibly prime) length, which a direct DFT takes unchanged; 4001 x 4000 x 6 flops
//                     per window is ~0.1 Gflop of float64, a 60 s clip ~12 Gflop: a millisecond of this GPU, not worth
//                     a mixed-radix FFT plus a Bluestein path for the ragged window.
//   features_kernel   CTA = window; five warps each run one of the reference's reductions in the reference's order
//                     (NumPy pairwise summation restated leaf by leaf, float32 where the reference is float32), so RMS, ZCR and
//                     the roll-off index are exact and the float64 features differ only by the rounding of the DFT,
//                     log and exp (~1e-13 relative; the parity bar is 1e-4).
#include "avd_internal.h"
#include <math.h>

namespace {

constexpr int SPEC_THREADS = 256, SPEC_BPT = 2, SPEC_BINS = SPEC_THREADS * SPEC_BPT;
constexpr int MAX_WIN = 9000;                       // (8 + 16) bytes per sample of shared memory

// ---- NumPy add.reduce (pairwise summation) over f(0..n), in type T --------------------------------------------
template <class T> __device__ __forceinline__ T radd(T a, T b);
template <> __device__ __forceinline__ float radd<float>(float a, float b) { return __fadd_rn(a, b); }
template <> __device__ __forceinline__ double radd<double>(double a, double b) { return __dadd_rn(a, b); }

template <class T, class F>
__device__ T pw_block(F f, int lo, int n)          // n <= 128
{
    if (n < 8) {
        T r = (T)-0.0;
        for (int i = 0; i < n; i++) r = radd<T>(r, f(lo + i));
        return r;
    }
    T r[8];
#pragma unroll
    for (int j = 0; j < 8; j++) r[j] = f(lo + j);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] = radd<T>(r[j], f(lo + i + j));
    }
    T res = radd<T>(radd<T>(radd<T>(r[0], r[1]), radd<T>(r[2], r[3])), radd<T>(radd<T>(r[4], r[5]), radd<T>(r[6], r[7])));
    for (; i < n; i++) res = radd<T>(res, f(lo + i));
    return res;
}

template <class T, class F>
__device__ T pw_sum(F f, int n)
{
    if (n <= 128) return pw_block<T>(f, 0, n);
    struct Fr { int lo, n; T left; int st; } stk[24];
    int sp = 0;
    stk[0] = {0, n, (T)0, 0};
    T ret = (T)0;
    while (sp >= 0) {
        Fr &t = stk[sp];
        if (t.n <= 128) { ret = pw_block<T>(f, t.lo, t.n); sp--; continue; }
        int n2 = t.n / 2; n2 -= n2 % 8;
        if (t.st == 0) { t.st = 1; stk[sp + 1] = {t.lo, n2, (T)0, 0}; sp++; }
        else if (t.st == 1) { t.left = ret; t.st = 2; stk[sp + 1] = {t.lo + n2, t.n - n2, (T)0, 0}; sp++; }
        else { ret = radd<T>(t.left, ret); sp--; }
    }
    return ret;
}

// The same reduction by a whole warp: the tree of the library routine is fixed by n alone, so lane 0 lists its leaves
// (blocks of <= 128 elements), the lanes sum the leaves in parallel (each leaf exactly as pw_block does) and lane 0
// combines the leaf sums in the tree's order.  Bit-identical to pw_sum; ~30x shorter critical path for n = 8000.
constexpr int PW_MAX_LEAVES = 192;                 // n <= 9000 gives at most 9000 / 64 leaves
template <class T> struct PwScratch { int lo[PW_MAX_LEAVES], n[PW_MAX_LEAVES]; T sum[PW_MAX_LEAVES]; };

template <class T, class F>
__device__ T pw_sum_warp(F f, int n, PwScratch<T> &S, int lane)
{
    if (n <= 128) {
        T r = (T)0;
        if (lane == 0) r = pw_block<T>(f, 0, n);
        return __shfl_sync(0xffffffffu, r, 0);
    }
    struct Fr { int lo, n, st; T left; };
    int leaves = 0;
    if (lane == 0) {
        Fr stk[24];
        int sp = 0;
        stk[0] = {0, n, 0, (T)0};
        while (sp >= 0) {
            Fr &t = stk[sp];
            if (t.n <= 128) { S.lo[leaves] = t.lo; S.n[leaves] = t.n; leaves++; sp--; continue; }
            int n2 = t.n / 2; n2 -= n2 % 8;
            if (t.st == 0) { t.st = 1; stk[sp + 1] = {t.lo, n2, 0, (T)0}; sp++; }
            else if (t.st == 1) { t.st = 2; stk[sp + 1] = {t.lo + n2, t.n - n2, 0, (T)0}; sp++; }
            else sp--;
        }
    }
    leaves = __shfl_sync(0xffffffffu, leaves, 0);
    __syncwarp();
    for (int i = lane; i < leaves; i += 32) S.sum[i] = pw_block<T>(f, S.lo[i], S.n[i]);
    __syncwarp();
    T ret = (T)0;
    if (lane == 0) {
        Fr stk[24];
        int sp = 0, leaf = 0;
        stk[0] = {0, n, 0, (T)0};
        while (sp >= 0) {
            Fr &t = stk[sp];
            if (t.n <= 128) { ret = S.sum[leaf++]; sp--; continue; }
            int n2 = t.n / 2; n2 -= n2 % 8;
            if (t.st == 0) { t.st = 1; stk[sp + 1] = {t.lo, n2, 0, (T)0}; sp++; }
            else if (t.st == 1) { t.left = ret; t.st = 2; stk[sp + 1] = {t.lo + n2, t.n - n2, 0, (T)0}; sp++; }
            else { ret = radd<T>(t.left, ret); sp--; }
        }
    }
    __syncwarp();
    return __shfl_sync(0xffffffffu, ret, 0);
}

// ---- |rfft(seg * hanning(len))| + 1e-9 -------------------------------------------------------------------------
// Real input: X[k] = x[0] + (len even) x[len/2] (-1)^k + sum_{i=1..(len-1)/2} (xe[i] cos(2 pi k i/len) - j xo[i] sin(2 pi k i/len))
// with xe[i] = x[i] + x[len-i], xo[i] = x[i] - x[len-i]: half the terms.  The angle advances by rotation with the bin's
// own root (4 flops) and is re-read from the exact table every RB terms, so the rounding of the recurrence never
// accumulates beyond RB steps (~1e-14) and the table is touched once per RB terms instead of once per term (a table
// look-up per term is bound by shared-memory bandwidth: 4.9 ms per 120 windows measured, against 0.9 ms this way).
constexpr int RB = 64;

__global__ void __launch_bounds__(SPEC_THREADS) spectrum_kernel(const float *__restrict__ wav, long long n_samples, int win, double *__restrict__ mag, int mag_stride, int w0)
{
    extern __shared__ __align__(16) uint8_t sm[];
    const int w = w0 + blockIdx.y;
    const long long s0 = (long long)w * win;
    const int len = (int)min((long long)win, n_samples - s0);
    const int L = len / 2 + 1, half = (len - 1) / 2;
    const int k0 = blockIdx.x * SPEC_BINS;
    if (k0 >= L) return;
    double2 *tw = (double2 *)sm;                                  // [len]: (cos, sin)(2 pi j / len), correctly rounded
    double *x = (double *)(sm + (size_t)len * 16);                // [len]: windowed samples, then xe / xo in place
    for (int i = threadIdx.x; i < len; i += SPEC_THREADS) {
        // np.hanning(M): ones(1) for M == 1, else 0.5 + 0.5*cos(pi*n/(M-1)), n = 1-M, 3-M, ...      (audio.py:46)
        double h = 1.0;
        if (len > 1) h = __dadd_rn(0.5, __dmul_rn(0.5, cos(__ddiv_rn(__dmul_rn(M_PI, (double)(2 * i + 1 - len)), (double)(len - 1)))));
        x[i] = __dmul_rn((double)wav[s0 + i], h);
        double sn, cs;
        sincospi(__ddiv_rn((double)(2 * i), (double)len), &sn, &cs);
        tw[i] = make_double2(cs, sn);
    }
    __syncthreads();
    // fold: x[i] <- xe[i], x[len-i] <- xo[i] for i = 1..half (each pair is touched by one thread)
    for (int i = 1 + threadIdx.x; i <= half; i += SPEC_THREADS) {
        const double a = x[i], b = x[len - i];
        x[i] = __dadd_rn(a, b);
        x[len - i] = __dsub_rn(a, b);
    }
    __syncthreads();
    const double x0 = x[0], xm = (len % 2 == 0 && len > 1) ? x[len / 2] : 0.0;
#pragma unroll
    for (int j = 0; j < SPEC_BPT; j++) {
        const int k = k0 + threadIdx.x + j * SPEC_THREADS;
        if (k >= L) continue;
        const double2 root = tw[k];                               // the bin's own root of unity
        const int kstep = (int)(((long long)k * RB) % len);
        int idx = k;                                              // (k * i0) mod len, i0 = 1
        double re = __dadd_rn(x0, (k & 1) ? -xm : xm), im = 0.0;
        for (int i0 = 1; i0 <= half; i0 += RB) {
            const double2 t = tw[idx];
            double c = t.x, sn = t.y;
            const int iend = min(i0 + RB, half + 1);
            for (int i = i0; i < iend; i++) {
                re = __fma_rn(x[i], c, re);
                im = __fma_rn(x[len - i], sn, im);
                const double cn = __fma_rn(c, root.x, -__dmul_rn(sn, root.y));
                sn = __fma_rn(sn, root.x, __dmul_rn(c, root.y));
                c = cn;
            }
            idx += kstep;
            if (idx >= len) idx -= len;
        }
        mag[(size_t)w * mag_stride + k] = __dadd_rn(hypot(re, im), 1e-9);      // audio.py:48 (the sign of im does not matter)
    }
}

// ---- the same spectrum for full windows whose length factors as N1 x N2 (8000 = 80 x 100): two-stage DFT ----------
// X[k1 + N1 k2] = sum_{n2} W_N2^{n2 k2} * ( W_N^{n2 k1} * sum_{n1} x[N2 n1 + n2] W_N1^{n1 k1} )      (Cooley-Tukey, one split)
// Each stage is a set of direct DFTs of length N1 / N2 out of shared memory: N (N1 + N2) multiply-adds per window
// instead of N^2 / 2, and sums of at most 100 terms.  Real input: stage 1 computes k1 <= N1/2 and mirrors the rest.
// One CTA per window; the thread mapping keeps every shared-memory access of a warp either consecutive or a broadcast.
constexpr int SPEC2_THREADS = 512;

__global__ void __launch_bounds__(SPEC2_THREADS) spectrum2_kernel(const float *__restrict__ wav, int win, int N1, int N2, double *__restrict__ mag, int mag_stride)
{
    extern __shared__ __align__(16) uint8_t sm[];
    const int w = blockIdx.x, N = win, L = N / 2 + 1;
    const long long s0 = (long long)w * win;
    double2 *Y = (double2 *)sm;                                   // [N2][N1]
    double *x = (double *)(sm + (size_t)N * 16);                  // [N]
    double2 *T1 = (double2 *)(sm + (size_t)N * 24), *T2 = T1 + N1;   // exp(-2 pi i j / N1), exp(-2 pi i j / N2)
    for (int i = threadIdx.x; i < N; i += SPEC2_THREADS) {
        double h = 1.0;
        if (N > 1) h = __dadd_rn(0.5, __dmul_rn(0.5, cos(__ddiv_rn(__dmul_rn(M_PI, (double)(2 * i + 1 - N)), (double)(N - 1)))));
        x[i] = __dmul_rn((double)wav[s0 + i], h);
    }
    for (int j = threadIdx.x; j < N1 + N2; j += SPEC2_THREADS) {
        const int n = j < N1 ? N1 : N2, jj = j < N1 ? j : j - N1;
        double sn, cs;
        sincospi(__ddiv_rn((double)(2 * jj), (double)n), &sn, &cs);
        T1[j] = make_double2(cs, -sn);                            // T2 follows T1
    }
    __syncthreads();
    // stage 1: Y[n2][k1] = sum_{n1} x[N2 n1 + n2] W_N1^{n1 k1}, k1 = 0 .. N1/2, mirrored to N1 - k1
    const int H1 = N1 / 2 + 1;
    for (int idx = threadIdx.x; idx < H1 * N2; idx += SPEC2_THREADS) {
        const int k1 = idx / N2, n2 = idx - k1 * N2;
        double re = 0.0, im = 0.0;
        int t = 0;
        for (int n1 = 0; n1 < N1; n1++) {
            const double xv = x[N2 * n1 + n2];
            const double2 tw = T1[t];
            re = __fma_rn(xv, tw.x, re);
            im = __fma_rn(xv, tw.y, im);
            t += k1;
            if (t >= N1) t -= N1;
        }
        Y[n2 * N1 + k1] = make_double2(re, im);
        if (k1 != 0 && 2 * k1 != N1) Y[n2 * N1 + N1 - k1] = make_double2(re, -im);
    }
    __syncthreads();
    // twiddles W_N^{n2 k1}
    for (int idx = threadIdx.x; idx < N; idx += SPEC2_THREADS) {
        const int n2 = idx / N1, k1 = idx - n2 * N1;
        double sn, cs;
        sincospi(__ddiv_rn((double)(2 * ((n2 * k1) % N)), (double)N), &sn, &cs);
        const double2 y = Y[idx];
        Y[idx] = make_double2(__fma_rn(y.x, cs, __dmul_rn(y.y, sn)), __fma_rn(y.y, cs, -__dmul_rn(y.x, sn)));   // y * (cs - i sn)
    }
    __syncthreads();
    // stage 2: X[k1 + N1 k2] = sum_{n2} Y[n2][k1] W_N2^{n2 k2}, for the bins 0 .. N/2
    for (int k = threadIdx.x; k < L; k += SPEC2_THREADS) {
        const int k2 = k / N1, k1 = k - k2 * N1;
        double re = 0.0, im = 0.0;
        int t = 0;
        for (int n2 = 0; n2 < N2; n2++) {
            const double2 y = Y[n2 * N1 + k1];
            const double2 tw = T2[t];
            re = __fma_rn(y.x, tw.x, __fma_rn(-y.y, tw.y, re));
            im = __fma_rn(y.x, tw.y, __fma_rn(y.y, tw.x, im));
            t += k2;
            if (t >= N2) t -= N2;
        }
        mag[(size_t)w * mag_stride + k] = __dadd_rn(hypot(re, im), 1e-9);
    }
}

// ---- the five per-window features ---------------------------------------------------------------------------
__global__ void __launch_bounds__(160) features_kernel(const float *__restrict__ wav, long long n_samples, int win,
                                                       const double *__restrict__ mag, int mag_stride, double *__restrict__ feats)
{
    extern __shared__ __align__(16) uint8_t sm[];
    __shared__ unsigned int zc_count;
    const int w = blockIdx.x;
    const long long s0 = (long long)w * win;
    const int len = (int)min((long long)win, n_samples - s0);
    const int L = len / 2 + 1;
    const size_t Lp = ((size_t)L + 1) & ~(size_t)1;
    double *m = (double *)sm, *lg = m + Lp, *fm = lg + Lp;      // [L] each: mag, log(mag), linspace(0,1,L) * mag
    float *seg = (float *)(fm + Lp);                            // [len]: samples, then their float32 squares
    if (threadIdx.x == 0) zc_count = 0;
    // centroid axis np.linspace(0, 1, L): i * (1/(L-1)), last point exactly 1                             (audio.py:58)
    const double step = L > 1 ? __ddiv_rn(1.0, (double)(L - 1)) : 0.0;
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        const double v = mag[(size_t)w * mag_stride + i];
        m[i] = v;
        lg[i] = log(v);
        fm[i] = __dmul_rn(i == L - 1 && L > 1 ? 1.0 : __dmul_rn((double)i, step), v);
    }
    for (int i = threadIdx.x; i < len; i += blockDim.x) seg[i] = wav[s0 + i];
    __syncthreads();
    // zero crossings: sum |sign(x[i+1]) - sign(x[i])| is an exact small integer in any order            (audio.py:44)
    unsigned int zc = 0;
    for (int i = threadIdx.x; i + 1 < len; i += blockDim.x) {
        const int a = (seg[i] > 0.f) - (seg[i] < 0.f), b = (seg[i + 1] > 0.f) - (seg[i + 1] < 0.f);
        zc += (unsigned int)abs(b - a);
    }
    atomicAdd(&zc_count, zc);
    __syncthreads();
    for (int i = threadIdx.x; i < len; i += blockDim.x) seg[i] = __fmul_rn(seg[i], seg[i]);      // seg**2, float32
    __syncthreads();
    // the element-wise work is done; what is left are the reference's reductions, each by one warp in the reference's order
    __shared__ PwScratch<double> scr[5];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *out = feats + (size_t)w * 5;
    PwScratch<double> &S = scr[warp];
    if (warp == 0) {
        // float(np.sqrt((seg**2).mean())): float32 pairwise sum, float32 division and sqrt                    (audio.py:43)
        const float s = pw_sum_warp<float>([&](int i) { return seg[i]; }, len, *(PwScratch<float> *)&S, lane);
        if (lane == 0) out[0] = (double)__fsqrt_rn(__fdiv_rn(s, (float)len));
    } else if (warp == 1) {
        // np.mean(float32 array)/2.0: the sum is an exact integer, the division float32                          (audio.py:44-45)
        if (lane == 0) out[1] = len > 1 ? (double)__fdiv_rn((float)zc_count, (float)(len - 1)) / 2.0 : __longlong_as_double(0x7ff8000000000000ll);
    } else if (warp == 2) {
        // flatness                                                                                            (audio.py:49)
        const double ml = __ddiv_rn(pw_sum_warp<double>([&](int i) { return lg[i]; }, L, S, lane), (double)L);
        const double mm = __ddiv_rn(pw_sum_warp<double>([&](int i) { return m[i]; }, L, S, lane), (double)L);
        if (lane == 0) out[2] = __ddiv_rn(exp(ml), mm);
    } else if (warp == 3) {
        // roll-off: first bin where the running sum reaches 85 % of np.sum(mag); the scan is sequential by definition (audio.py:50-57)
        const double cutoff = __dmul_rn(0.85, pw_sum_warp<double>([&](int i) { return m[i]; }, L, S, lane));
        if (lane == 0) {
            // mag > 0, so the running sum only grows: test it once per 8 bins (the 8 loads are then independent of the
            // adds), and find the exact bin inside the block that crosses -- same additions, same order
            double s = 0.0;
            int idx = 0, i = 0;
            bool found = false;
            for (; i + 8 <= L && !found; i += 8) {
                double v[8];
#pragma unroll
                for (int j = 0; j < 8; j++) v[j] = m[i + j];
                double t = s;
#pragma unroll
                for (int j = 0; j < 8; j++) t = __dadd_rn(t, v[j]);
                if (t >= cutoff) {
                    for (int j = 0; j < 8; j++) {
                        s = __dadd_rn(s, v[j]);
                        if (s >= cutoff) { idx = i + j; found = true; break; }
                    }
                } else s = t;
            }
            if (!found)
                for (i = L & ~7; i < L; i++) {          // tail shorter than a block (only reached if no block crossed)
                    s = __dadd_rn(s, m[i]);
                    if (s >= cutoff) { idx = i; break; }
                }
            out[3] = __ddiv_rn((double)idx, fmax(1.0, (double)L));
        }
    } else {
        // centroid                                                                                            (audio.py:59-60)
        const double num = pw_sum_warp<double>([&](int i) { return fm[i]; }, L, S, lane);
        const double den = pw_sum_warp<double>([&](int i) { return m[i]; }, L, S, lane);
        if (lane == 0) out[4] = __ddiv_rn(num, den);
    }
}

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int avd_audio_windows(avd_ctx *ctx, const float *h_wav, long long n_samples, int win, double *h_feats, void *stream)
{
    if (!ctx) return AVD_ERR_ARG;
    if (n_samples < 0 || win < 1 || win > MAX_WIN) { ctx->err = "audio: n_samples < 0 or window length outside 1..9000"; return AVD_ERR_ARG; }
    if (n_samples == 0) return AVD_OK;
    if (!h_wav || !h_feats) { ctx->err = "null buffer"; return AVD_ERR_ARG; }
    const long long nwin = (n_samples + win - 1) / win;
    if (nwin > 65535) { ctx->err = "audio: more than 65535 windows in one call"; return AVD_ERR_ARG; }
    AVD_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;
    const int Lmax = win / 2 + 1, mag_stride = (Lmax + 1) & ~1;
    const size_t o_wav = 0, o_mag = ((size_t)n_samples * 4 + 255) & ~(size_t)255, o_ft = o_mag + (size_t)nwin * mag_stride * 8,
                 total = o_ft + (size_t)nwin * 5 * 8;
    int rc = avd_ensure(ctx, &ctx->k4_ws, &ctx->k4_ws_cap, total);
    if (rc) return rc;
    uint8_t *ws = (uint8_t *)ctx->k4_ws;
    float *d_wav = (float *)(ws + o_wav);
    double *d_mag = (double *)(ws + o_mag), *d_ft = (double *)(ws + o_ft);
    AVD_CUDA(ctx, cudaMemcpyAsync(d_wav, h_wav, (size_t)n_samples * 4, cudaMemcpyHostToDevice, st));
    // full windows: two-stage DFT if win = N1 x N2 with both factors <= 128; the ragged last window (and every window when
    // win does not factor that way) takes the direct DFT
    int N1 = 0, N2 = 0;
    for (int d = 1; d * d <= win; d++)
        if (win % d == 0 && win / d <= 128) { N1 = d; N2 = win / d; }
    const long long nfull = n_samples / win;
    const bool two_stage = N1 > 1 && nfull > 0;
    const size_t smem_spec = (size_t)win * 24, smem_spec2 = (size_t)win * 24 + (size_t)(N1 + N2) * 16;
    const size_t smem_feat = 3 * 8 * (((size_t)Lmax + 1) & ~(size_t)1) + (size_t)win * 4;
    static size_t set_spec[64] = {}, set_spec2[64] = {}, set_feat[64] = {};
    if (smem_spec > set_spec[ctx->device & 63]) {
        AVD_CUDA(ctx, cudaFuncSetAttribute(spectrum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_spec));
        set_spec[ctx->device & 63] = smem_spec;
    }
    if (two_stage && smem_spec2 > set_spec2[ctx->device & 63]) {
        AVD_CUDA(ctx, cudaFuncSetAttribute(spectrum2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_spec2));
        set_spec2[ctx->device & 63] = smem_spec2;
    }
    if (smem_feat > set_feat[ctx->device & 63]) {
        AVD_CUDA(ctx, cudaFuncSetAttribute(features_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_feat));
        set_feat[ctx->device & 63] = smem_feat;
    }
    const int w_direct0 = two_stage ? (int)nfull : 0, n_direct = (int)nwin - w_direct0;
    if (two_stage)
        AVD_LAUNCH(ctx, "k4_spectrum2", st, (spectrum2_kernel<<<(unsigned)nfull, SPEC2_THREADS, smem_spec2, st>>>(d_wav, win, N1, N2, d_mag, mag_stride)));
    if (n_direct > 0)
        AVD_LAUNCH(ctx, "k4_spectrum", st, (spectrum_kernel<<<dim3((Lmax + SPEC_BINS - 1) / SPEC_BINS, (unsigned)n_direct), SPEC_THREADS, smem_spec, st>>>(d_wav, n_samples, win, d_mag, mag_stride, w_direct0)));
    AVD_LAUNCH(ctx, "k4_features", st, (features_kernel<<<(unsigned)nwin, 160, smem_feat, st>>>(d_wav, n_samples, win, d_mag, mag_stride, d_ft)));
    AVD_CUDA(ctx, cudaMemcpyAsync(h_feats, d_ft, (size_t)nwin * 5 * 8, cudaMemcpyDeviceToHost, st));
    AVD_CUDA(ctx, cudaStreamSynchronize(st));
    return AVD_OK;
}

#pragma GCC visibility pop
}
