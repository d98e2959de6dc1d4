// K3: per-clip scoring tail on the device (include/avd.h: avd_score_clips).
//
// One thread per clip walks that clip's records and timelines in float64 in exactly the order the reference's
// NumPy calls do (video.py:54-81, fusion.py:7-14,16-109):
//   np.mean / np.sum        pairwise summation: < 8 elements sequential from -0.0; <= 128 elements eight strided
//                           partial sums combined as ((r0+r1)+(r2+r3))+((r4+r5)+(r6+r7)) then the tail one by one;
//                           longer inputs split at n/2 rounded down to a multiple of 8
//   np.var / np.std         mean first, then the pairwise sum of (x - mean)^2, divided by n; std = sqrt
//   np.convolve(t, 1/3, "same")  interior: (t[i-1]*w + t[i]*w) + t[i+1]*w with every product rounded;
//                           the two end points: fma(t[hi], w, t[lo]*w)  (what the library's 2-term dot does)
//   round(x, 2)             correctly rounded decimal rounding, ties to even on the exact binary value
// Every operation is an explicit round-to-nearest intrinsic, so no contraction can change a result.
// The work is O(samples) scalars per clip: a batch of 256 clips is 256 threads; the point of the kernel is that a
// clip batch's records never leave HBM and only the scores come back, not speed of the arithmetic.
#include "avd_internal.h"
#include <string.h>

namespace {

__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double ddiv(double a, double b) { return __ddiv_rn(a, b); }
__device__ __forceinline__ double dclip(double x, double lo, double hi) { return fmin(fmax(x, lo), hi); }
__device__ __forceinline__ int dsign(double x) { return (x > 0.0) - (x < 0.0); }

// a NumPy add.reduce over f(0..n): f is the element functor
template <class F>
__device__ double pw_block(F f, int lo, int n)          // n <= 128
{
    if (n < 8) {
        double r = -0.0;
        for (int i = 0; i < n; i++) r = dadd(r, f(lo + i));
        return r;
    }
    double r[8];
#pragma unroll
    for (int j = 0; j < 8; j++) r[j] = f(lo + j);
    int i = 8;
    for (; i < n - (n % 8); i += 8) {
#pragma unroll
        for (int j = 0; j < 8; j++) r[j] = dadd(r[j], f(lo + i + j));
    }
    double res = dadd(dadd(dadd(r[0], r[1]), dadd(r[2], r[3])), dadd(dadd(r[4], r[5]), dadd(r[6], r[7])));
    for (; i < n; i++) res = dadd(res, f(lo + i));
    return res;
}

// the recursion of the library routine, unrolled into an explicit stack of (lo, n, state) frames
template <class F>
__device__ double pw_sum(F f, int n)
{
    if (n <= 128) return pw_block(f, 0, n);
    struct Fr { int lo, n; double left; int st; } stk[24];
    int sp = 0;
    stk[0] = {0, n, 0.0, 0};
    double ret = 0.0;
    while (sp >= 0) {
        Fr &t = stk[sp];
        if (t.n <= 128) { ret = pw_block(f, t.lo, t.n); sp--; continue; }
        int n2 = t.n / 2; n2 -= n2 % 8;
        if (t.st == 0) { t.st = 1; stk[sp + 1] = {t.lo, n2, 0.0, 0}; sp++; }
        else if (t.st == 1) { t.left = ret; t.st = 2; stk[sp + 1] = {t.lo + n2, t.n - n2, 0.0, 0}; sp++; }
        else { ret = dadd(t.left, ret); sp--; }
    }
    return ret;
}

template <class F> __device__ double np_mean(F f, int n) { return ddiv(pw_sum(f, n), (double)n); }
template <class F>
__device__ double np_var(F f, int n)
{
    const double m = np_mean(f, n);
    return ddiv(pw_sum([&](int i) { double d = dsub(f(i), m); return dmul(d, d); }, n), (double)n);
}

// Python round(x, 2) for 0 <= x <= 1: q = x*100 rounded half-to-even on the exact value, result = q/100 correctly rounded
__device__ double round2(double x)
{
    if (!(x >= 0x1p-10)) return 0.0;                     // 100x < 0.5
    const unsigned long long bits = (unsigned long long)__double_as_longlong(x);
    const int e = (int)((bits >> 52) & 0x7ff) - 1075;    // x = m * 2^e, normal numbers only here
    const unsigned long long m = (bits & 0xfffffffffffffull) | (1ull << 52);
    if (e >= 0) return x;                                // integers >= 2^52 (not reachable for scores)
    const int s = -e;                                    // 1 .. 62
    const unsigned long long p = m * 100ull;             // < 2^60
    unsigned long long q = p >> s;
    const unsigned long long rem = p & ((1ull << s) - 1ull), half = 1ull << (s - 1);
    if (rem > half || (rem == half && (q & 1ull))) q++;
    return ddiv((double)q, 100.0);
}

struct K3Args {
    const avd_sample_rec *rec;
    const avd_clip_job *jobs;
    const double *tl_in;
    avd_clip_score *scores;
    double *vt, *fused, *binned;
    uint8_t *peaks;
    double *scratch;           // [2 * n_rec]: textures, compacted flow means
    int n_rec, n_clips;
};

__global__ void __launch_bounds__(32) k3_score_kernel(K3Args A)
{
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= A.n_clips) return;
    const avd_clip_job J = A.jobs[c];
    avd_clip_score S;
    memset(&S, 0, sizeof(S));
    double *V = A.vt + J.out_off;
    int vlen;
    double flow_mean, texture_var, sc_rate, dup_density;

    if (!J.video_given) {
        // ---- video.py:54-81 ----
        const int n = J.rec_count;
        const avd_sample_rec *R = A.rec + J.rec_first;
        double *T = A.scratch + J.rec_first, *FM = A.scratch + A.n_rec + J.rec_first;
        const double N = (double)J.npix;
        int k = 0, dup = 0, cuts = 0;
        double last = 0.5;
        for (int i = 0; i < n; i++) {
            const double s = (double)R[i].lap_sum, ss = (double)R[i].lap_sumsq;
            const double tex = ddiv(dsub(ss, ddiv(dmul(s, s), N)), N);          // float(lap.var()), video.py:52
            T[i] = tex;
            if (i > 0 && (R[i].flags & AVD_REC_HAS_FLOW)) {
                FM[k++] = (double)R[i].flow_mean;
                cuts += (double)R[i].flow_var > 0.5;                            // video.py:62
            }
            if (i > 0 && (R[i].flags & AVD_REC_HAS_PREV) && R[i].ham_prev == 0) dup++;   // video.py:39-40
            const double mot = k ? FM[k - 1] : 0.0;                             // video.py:55
            last = dclip(dsub(1.0, dmul(ddiv(tex, dadd(tex, 1000.0)), dadd(1.0, mot))), 0.0, 1.0);   // video.py:56
            if (i < J.tlen) V[i] = last;
        }
        for (int i = n; i < J.tlen; i++) V[i] = last;                           // video.py:74-79 (0.5 if no samples)
        vlen = J.tlen;
        dup_density = ddiv((double)dup, (double)(n - 1 > 1 ? n - 1 : 1));       // video.py:61
        sc_rate = k ? ddiv((double)cuts, (double)k) : 0.0;
        flow_mean = k ? np_mean([&](int i) { return FM[i]; }, k) : 0.0;
        S.flow_var = k ? np_var([&](int i) { return FM[i]; }, k) : 0.0;
        texture_var = n ? np_var([&](int i) { return T[i]; }, n) : 0.0;
    } else {
        vlen = J.vgiven_len;
        for (int i = 0; i < vlen; i++) V[i] = A.tl_in[J.vgiven_off + i];
        flow_mean = J.g_flow_mean; texture_var = J.g_texture_var; sc_rate = J.g_sc_rate; dup_density = J.g_dup_density;
    }
    S.dup_density = dup_density; S.scene_change_rate = sc_rate; S.flow_mean = flow_mean; S.texture_var = texture_var;
    S.vlen = vlen;
    S.L = vlen;

    if (J.do_fuse) {
        // ---- fusion.py:16-109 ----
        const int alen = J.audio_len;
        const int L = max(max(alen, vlen), 1);
        const double *AU = A.tl_in + J.audio_off;
        const double vpad = vlen ? V[vlen - 1] : 0.5, apad = alen ? AU[alen - 1] : 0.5;
        for (int i = vlen; i < L; i++) V[i] = vpad;                             // fusion.py:21 (in place)
        auto a_at = [&](int i) { return i < alen ? AU[i] : apad; };             // fusion.py:20
        auto v_at = [&](int i) { return V[i]; };
        const double mean_a = np_mean(a_at, L), mean_v = np_mean(v_at, L);
        double w_a = 0.65, w_v = 0.25;
        const double agree = dsign(dsub(mean_a, 0.5)) == dsign(dsub(mean_v, 0.5)) ? 0.10 : 0.0;
        if (J.speech_ratio < 0.25) {
            w_a = dmul(w_a, 0.6);
            w_v = fmax(0.2, dsub(dsub(1.0, w_a), agree));
        }
        double bonus = 0.0;
        if (flow_mean > 5.0 && texture_var > 200.0 && dup_density < 0.05) bonus = dsub(bonus, 0.10);
        if (sc_rate > 0.7) bonus = dsub(bonus, 0.05);
        if (sc_rate >= 0.9 && texture_var > 300.0 && dup_density < 0.02) bonus = dsub(bonus, 0.08);
        if (J.tts_like >= 0.95 && flow_mean > 8.0 && texture_var > 300.0 && dup_density < 0.05) {
            w_a = dmul(w_a, 0.55);
            w_v = fmax(0.25, dsub(dsub(1.0, w_a), agree));
        }
        double *TL = A.fused + J.out_off, *B = A.binned + J.out_off;
        uint8_t *PK = A.peaks + J.out_off;
        int npk = 0;
        for (int i = 0; i < L; i++) {
            const double a = a_at(i), v = V[i];
            const double mix = dadd(dadd(dmul(w_a, a), dmul(w_v, v)), ddiv(dmul(agree, dadd(a, v)), 2.0));
            const double x = dclip(dadd(dsub(mix, J.pen), bonus), 0.0, 1.0);
            TL[i] = x;
            const int pk = x <= 0.25 || x >= 0.75;
            PK[i] = (uint8_t)pk;
            npk += pk;
        }
        auto t_at = [&](int i) { return TL[i]; };
        S.score = np_mean(t_at, L);
        S.spread = __dsqrt_rn(np_var(t_at, L));
        const double disagree = fabs(dsub(mean_a, mean_v));
        S.conf = dclip(dsub(dsub(dadd(0.20, dmul(2.2, S.spread)), J.pen), dmul(0.5, fmax(0.0, dsub(0.3, disagree)))), 0.10, 0.99);
        S.label = S.score <= J.thresh_real_max ? 0 : (S.score >= J.thresh_ai_min ? 2 : 1);
        const double w = 1.0 / 3.0;                                            // np.ones(3) / 3.0
        if (L >= 3) {
            for (int i = 0; i < L; i++) {
                double s;
                if (i == 0) s = __fma_rn(TL[1], w, dmul(TL[0], w));
                else if (i == L - 1) s = __fma_rn(TL[L - 1], w, dmul(TL[L - 2], w));
                else s = dadd(dadd(dmul(TL[i - 1], w), dmul(TL[i], w)), dmul(TL[i + 1], w));
                B[i] = dclip(s, 0.0, 1.0);
            }
        } else {
            for (int i = 0; i < L; i++) B[i] = dclip(TL[i], 0.0, 1.0);
        }
        S.ai_score = round2(S.score);
        S.confidence = round2(S.conf);
        S.L = L;
        S.n_peaks = npk;
    }
    A.scores[c] = S;
}

inline size_t al16(size_t x) { return (x + 15) & ~(size_t)15; }

}  // namespace

extern "C" {
#pragma GCC visibility push(default)

int avd_score_clips(avd_ctx *ctx, const avd_sample_rec *d_rec, int n_rec, const avd_clip_job *h_jobs, int n_clips,
                    const double *h_tl_in, int n_tl_in, avd_clip_score *h_scores, double *h_vtimeline,
                    double *h_fused, double *h_binned, uint8_t *h_peaks, int n_out, void *stream)
{
    if (!ctx) return AVD_ERR_ARG;
    if (n_clips < 0 || n_rec < 0 || n_tl_in < 0 || n_out < 0) { ctx->err = "negative count"; return AVD_ERR_ARG; }
    if (n_clips == 0) return AVD_OK;
    if (!h_jobs || !h_scores || !h_vtimeline) { ctx->err = "null buffer"; return AVD_ERR_ARG; }
    if (n_tl_in > 0 && !h_tl_in) { ctx->err = "null timeline input"; return AVD_ERR_ARG; }
    if (!d_rec) {
        d_rec = ctx->recs;
        if (n_rec > ctx->recs_n) { ctx->err = "n_rec exceeds the records of the last avd_analyze_host run"; return AVD_ERR_ARG; }
    }
    bool any_fuse = false;
    for (int c = 0; c < n_clips; c++) {
        const avd_clip_job &j = h_jobs[c];
        if (j.tlen < 1 || j.out_off < 0) { ctx->err = "clip job: tlen < 1 or negative output offset"; return AVD_ERR_ARG; }
        int vlen;
        if (j.video_given) {
            if (!j.do_fuse) { ctx->err = "clip job: video_given without do_fuse has nothing to compute"; return AVD_ERR_ARG; }
            if (j.vgiven_len < 0 || j.vgiven_off < 0 || (long long)j.vgiven_off + j.vgiven_len > n_tl_in) { ctx->err = "clip job: given video timeline out of range"; return AVD_ERR_ARG; }
            vlen = j.vgiven_len;
        } else {
            if (j.rec_count < 0 || j.rec_first < 0 || (long long)j.rec_first + j.rec_count > n_rec) { ctx->err = "clip job: records out of range"; return AVD_ERR_ARG; }
            if (j.rec_count > 0 && (j.npix < 1 || !d_rec)) { ctx->err = "clip job: npix < 1 or no records"; return AVD_ERR_ARG; }
            vlen = j.tlen;
        }
        int L = vlen;
        if (j.do_fuse) {
            any_fuse = true;
            if (j.audio_len < 0 || j.audio_off < 0 || (long long)j.audio_off + j.audio_len > n_tl_in) { ctx->err = "clip job: audio timeline out of range"; return AVD_ERR_ARG; }
            L = j.audio_len > L ? j.audio_len : L;
            if (L < 1) L = 1;
        }
        if ((long long)j.out_off + L > n_out) { ctx->err = "clip job: outputs exceed n_out"; return AVD_ERR_ARG; }
    }
    if (any_fuse && (!h_fused || !h_binned || !h_peaks)) { ctx->err = "null fuse output"; return AVD_ERR_ARG; }
    AVD_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;

    // one workspace: jobs | timeline inputs | scores | vt | fused | binned | scratch | peaks
    const size_t o_jobs = 0, o_in = al16(o_jobs + sizeof(avd_clip_job) * n_clips), o_sc = al16(o_in + 8 * (size_t)n_tl_in),
                 o_vt = al16(o_sc + sizeof(avd_clip_score) * n_clips), o_fu = al16(o_vt + 8 * (size_t)n_out),
                 o_bi = al16(o_fu + 8 * (size_t)n_out), o_scr = al16(o_bi + 8 * (size_t)n_out),
                 o_pk = al16(o_scr + 16 * (size_t)n_rec), total = al16(o_pk + (size_t)n_out);
    int rc = avd_ensure(ctx, &ctx->k3_ws, &ctx->k3_ws_cap, total);
    if (rc) return rc;
    uint8_t *ws = (uint8_t *)ctx->k3_ws;
    AVD_CUDA(ctx, cudaMemcpyAsync(ws + o_jobs, h_jobs, sizeof(avd_clip_job) * n_clips, cudaMemcpyHostToDevice, st));
    if (n_tl_in) AVD_CUDA(ctx, cudaMemcpyAsync(ws + o_in, h_tl_in, 8 * (size_t)n_tl_in, cudaMemcpyHostToDevice, st));
    AVD_CUDA(ctx, cudaMemsetAsync(ws + o_sc, 0, total - o_sc, st));      // positions no clip writes read back as zeros
    K3Args A;
    A.rec = d_rec; A.jobs = (const avd_clip_job *)(ws + o_jobs); A.tl_in = (const double *)(ws + o_in);
    A.scores = (avd_clip_score *)(ws + o_sc); A.vt = (double *)(ws + o_vt); A.fused = (double *)(ws + o_fu);
    A.binned = (double *)(ws + o_bi); A.peaks = ws + o_pk; A.scratch = (double *)(ws + o_scr);
    A.n_rec = n_rec; A.n_clips = n_clips;
    AVD_LAUNCH(ctx, "k3_score", st, (k3_score_kernel<<<(n_clips + 31) / 32, 32, 0, st>>>(A)));
    AVD_CUDA(ctx, cudaMemcpyAsync(h_scores, ws + o_sc, sizeof(avd_clip_score) * n_clips, cudaMemcpyDeviceToHost, st));
    if (n_out) {
        AVD_CUDA(ctx, cudaMemcpyAsync(h_vtimeline, ws + o_vt, 8 * (size_t)n_out, cudaMemcpyDeviceToHost, st));
        if (any_fuse) {
            AVD_CUDA(ctx, cudaMemcpyAsync(h_fused, ws + o_fu, 8 * (size_t)n_out, cudaMemcpyDeviceToHost, st));
            AVD_CUDA(ctx, cudaMemcpyAsync(h_binned, ws + o_bi, 8 * (size_t)n_out, cudaMemcpyDeviceToHost, st));
            AVD_CUDA(ctx, cudaMemcpyAsync(h_peaks, ws + o_pk, (size_t)n_out, cudaMemcpyDeviceToHost, st));
        }
    }
    AVD_CUDA(ctx, cudaStreamSynchronize(st));
    return AVD_OK;
}

#pragma GCC visibility pop
}
