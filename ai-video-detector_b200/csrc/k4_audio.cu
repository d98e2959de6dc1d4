// K4: per-window features of the audio analyzer (include/avd.h: avd_audio_windows).
//
// The reference cuts the 16 kHz mono signal into 0.5 s windows (8000 samples, the last one ragged) and computes per
// window (audio.py:41-61): RMS and zero-crossing rate in float32, then in float64 the Hann-windowed real spectrum,
// its flatness exp(mean(log|X|))/mean|X|, the 85 % roll-off bin (a sequential scan) and the centroid on a [0,1] axis.
//
//   spectrum_kernel   CTA = (window, block of 512 bins).  The windowed samples (float64) and the table of the len-th
//                     roots of unity (one sincospi per entry, so every twiddle is correctly rounded) sit in shared
//                     memory; a thread owns two bins and walks the samples with table index (k*i) mod len kept
//                     incrementally -- a direct DFT.  The last window has an arbitrary (possibly prime) length, which a
//                     direct DFT takes unchanged; 4001 x 8000 x 4 flops per window is ~0.13 Gflop of float64, a 60 s
//                     clip ~15 Gflop: a millisecond of this GPU, not worth a mixed-radix FFT plus a Bluestein path.
//   features_kernel   CTA = window; five warps each run one of the reference's reductions in the reference's order
//                     (NumPy pairwise summation restated, float32 where the reference is float32), so RMS, ZCR and
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

// ---- |rfft(seg * hanning(len))| + 1e-9 -------------------------------------------------------------------------
__global__ void __launch_bounds__(SPEC_THREADS) spectrum_kernel(const float *__restrict__ wav, long long n_samples, int win, double *__restrict__ mag, int mag_stride)
{
    extern __shared__ __align__(16) uint8_t sm[];
    const int w = blockIdx.y;
    const long long s0 = (long long)w * win;
    const int len = (int)min((long long)win, n_samples - s0);
    const int L = len / 2 + 1;
    const int k0 = blockIdx.x * SPEC_BINS;
    if (k0 >= L) return;
    double *x = (double *)sm;                                   // [len]
    double2 *tw = (double2 *)(sm + (((size_t)len * 8 + 15) & ~(size_t)15));   // [len]: exp(-2 pi i j / len)
    for (int i = threadIdx.x; i < len; i += SPEC_THREADS) {
        // np.hanning(M): ones(1) for M == 1, else 0.5 + 0.5*cos(pi*n/(M-1)), n = 1-M, 3-M, ...      (audio.py:46)
        double h = 1.0;
        if (len > 1) h = __dadd_rn(0.5, __dmul_rn(0.5, cos(__ddiv_rn(__dmul_rn(M_PI, (double)(2 * i + 1 - len)), (double)(len - 1)))));
        x[i] = __dmul_rn((double)wav[s0 + i], h);
        double sn, cs;
        sincospi(__ddiv_rn((double)(2 * i), (double)len), &sn, &cs);
        tw[i] = make_double2(cs, -sn);
    }
    __syncthreads();
    int k[SPEC_BPT], idx[SPEC_BPT];
    double re[SPEC_BPT], im[SPEC_BPT];
#pragma unroll
    for (int j = 0; j < SPEC_BPT; j++) {
        k[j] = k0 + threadIdx.x + j * SPEC_THREADS;
        if (k[j] >= L) k[j] = 0;                                // idle bin: computes X[0] again, not stored
        idx[j] = 0; re[j] = 0.0; im[j] = 0.0;
    }
#pragma unroll 4
    for (int i = 0; i < len; i++) {
        const double xi = x[i];
#pragma unroll
        for (int j = 0; j < SPEC_BPT; j++) {
            const double2 t = tw[idx[j]];
            re[j] = fma(xi, t.x, re[j]);
            im[j] = fma(xi, t.y, im[j]);
            idx[j] += k[j];
            if (idx[j] >= len) idx[j] -= len;
        }
    }
#pragma unroll
    for (int j = 0; j < SPEC_BPT; j++) {
        const int kk = k0 + threadIdx.x + j * SPEC_THREADS;
        if (kk < L) mag[(size_t)w * mag_stride + kk] = __dadd_rn(hypot(re[j], im[j]), 1e-9);      // audio.py:48
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
    double *m = (double *)sm;                                   // [L]
    float *seg = (float *)(sm + (((size_t)L * 8 + 15) & ~(size_t)15));   // [len]
    if (threadIdx.x == 0) zc_count = 0;
    for (int i = threadIdx.x; i < L; i += blockDim.x) m[i] = mag[(size_t)w * mag_stride + i];
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
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    double *out = feats + (size_t)w * 5;
    if (lane != 0) return;
    if (warp == 0) {
        // float(np.sqrt((seg**2).mean())): float32 squares, float32 pairwise sum, float32 division and sqrt     (audio.py:43)
        const float s = pw_sum<float>([&](int i) { return __fmul_rn(seg[i], seg[i]); }, len);
        out[0] = (double)__fsqrt_rn(__fdiv_rn(s, (float)len));
    } else if (warp == 1) {
        // np.mean(float32 array)/2.0: the sum is an exact integer, the division float32                          (audio.py:44-45)
        out[1] = len > 1 ? (double)__fdiv_rn((float)zc_count, (float)(len - 1)) / 2.0 : __longlong_as_double(0x7ff8000000000000ll);
    } else if (warp == 2) {
        // flatness                                                                                            (audio.py:49)
        const double ml = __ddiv_rn(pw_sum<double>([&](int i) { return log(m[i]); }, L), (double)L);
        const double mm = __ddiv_rn(pw_sum<double>([&](int i) { return m[i]; }, L), (double)L);
        out[2] = __ddiv_rn(exp(ml), mm);
    } else if (warp == 3) {
        // roll-off: first bin where the running sum reaches 85 % of np.sum(mag)                             (audio.py:50-57)
        const double total = pw_sum<double>([&](int i) { return m[i]; }, L);
        const double cutoff = __dmul_rn(0.85, total);
        double s = 0.0;
        int idx = 0;
        for (int i = 0; i < L; i++) {
            s = __dadd_rn(s, m[i]);
            if (s >= cutoff) { idx = i; break; }
        }
        out[3] = __ddiv_rn((double)idx, fmax(1.0, (double)L));
    } else {
        // centroid on np.linspace(0, 1, L): i * (1/(L-1)), last point exactly 1                              (audio.py:58-60)
        const double step = L > 1 ? __ddiv_rn(1.0, (double)(L - 1)) : 0.0;
        const double num = pw_sum<double>([&](int i) { return __dmul_rn(i == L - 1 && L > 1 ? 1.0 : __dmul_rn((double)i, step), m[i]); }, L);
        const double den = pw_sum<double>([&](int i) { return m[i]; }, L);
        out[4] = __ddiv_rn(num, den);
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
    const size_t smem_spec = (((size_t)win * 8 + 15) & ~(size_t)15) + (size_t)win * 16;
    const size_t smem_feat = (((size_t)Lmax * 8 + 15) & ~(size_t)15) + (size_t)win * 4;
    static size_t set_spec[64] = {}, set_feat[64] = {};
    if (smem_spec > set_spec[ctx->device & 63]) {
        AVD_CUDA(ctx, cudaFuncSetAttribute(spectrum_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_spec));
        set_spec[ctx->device & 63] = smem_spec;
    }
    if (smem_feat > set_feat[ctx->device & 63]) {
        AVD_CUDA(ctx, cudaFuncSetAttribute(features_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem_feat));
        set_feat[ctx->device & 63] = smem_feat;
    }
    AVD_LAUNCH(ctx, "k4_spectrum", st, (spectrum_kernel<<<dim3((Lmax + SPEC_BINS - 1) / SPEC_BINS, (unsigned)nwin), SPEC_THREADS, smem_spec, st>>>(d_wav, n_samples, win, d_mag, mag_stride)));
    AVD_LAUNCH(ctx, "k4_features", st, (features_kernel<<<(unsigned)nwin, 160, smem_feat, st>>>(d_wav, n_samples, win, d_mag, mag_stride, d_ft)));
    AVD_CUDA(ctx, cudaMemcpyAsync(h_feats, d_ft, (size_t)nwin * 5 * 8, cudaMemcpyDeviceToHost, st));
    AVD_CUDA(ctx, cudaStreamSynchronize(st));
    return AVD_OK;
}

#pragma GCC visibility pop
}
