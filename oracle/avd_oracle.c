/*
 * avd_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Plain-C CPU restatement of the arithmetic on the per-frame video-analyzer hot path of
 * backtato/ai-video-detector (app/analyzers/video.py).  The reference itself holds no
 * arithmetic: every number is produced by a third-party call (opencv-python-headless,
 * pinned >=4.10,<4.11 in requirements.txt:8; numpy >=1.26,<2.0 in requirements.txt:7).
 * Neither library's source is under /root/reference, so each routine below restates the
 * library's published algorithm for exactly the call the reference makes, citing the call
 * site.  The restatement is pinned against outputs of the reference's own code executed in
 * the build container (cv2 4.13.0 / numpy 2.3.5): see tests/golden/make_golden.py and
 * tests/test_oracle_golden.py.
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may load this file's
 * shared object.  The product (libavd.so) never does.
 *
 * Build: make -C oracle      (gcc -O2 -ffp-contract=off: no FMA contraction, the float
 *                             sequences below must round exactly as written)
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define AVD_O_API __attribute__((visibility("default")))

/* ------------------------------------------------------------------------------------------
 * C1  cv2.cvtColor(img, COLOR_BGR2GRAY) on uint8            video.py:5, :43, :51
 * 15-bit fixed point: Y = (3735 B + 19235 G + 9798 R + 16384) >> 15
 * ---------------------------------------------------------------------------------------- */
AVD_O_API void avd_o_gray(const uint8_t *bgr, int H, int W, size_t pitch, uint8_t *gray)
{
    for (int y = 0; y < H; y++) {
        const uint8_t *s = bgr + (size_t)y * pitch;
        uint8_t *d = gray + (size_t)y * W;
        for (int x = 0; x < W; x++)
            d[x] = (uint8_t)((3735u * s[3 * x] + 19235u * s[3 * x + 1] + 9798u * s[3 * x + 2] + 16384u) >> 15);
    }
}

/* round-half-to-even of a float, saturated to uint8 (cv::saturate_cast<uchar>(float)) */
static uint8_t sat_u8_rne(float v)
{
    long r = lrintf(v); /* default rounding mode: to nearest even */
    return (uint8_t)(r < 0 ? 0 : (r > 255 ? 255 : r));
}

/* ------------------------------------------------------------------------------------------
 * C2  cv2.resize(gray, (32,32), interpolation=INTER_AREA)                     video.py:6
 * Two library code paths:
 *   both scale factors integer -> block sum (int) * float(1/area), round half even
 *   otherwise                  -> decimation table (fractional first/last taps), float32
 *                                 weights, strictly sequential float32 accumulation,
 *                                 horizontal pass per source row then vertical pass
 * ---------------------------------------------------------------------------------------- */
typedef struct { int si, di; float alpha; } area_tap;

static int area_tab(int ssize, int dsize, double scale, area_tap *tab)
{
    int k = 0;
    for (int d = 0; d < dsize; d++) {
        double fs1 = d * scale, fs2 = fs1 + scale;
        double cell = scale < (ssize - fs1) ? scale : (ssize - fs1);
        int s1 = (int)ceil(fs1), s2 = (int)floor(fs2);
        if (s2 > ssize - 1) s2 = ssize - 1;
        if (s1 > s2) s1 = s2;
        if (s1 - fs1 > 1e-3) { tab[k].di = d; tab[k].si = s1 - 1; tab[k++].alpha = (float)((s1 - fs1) / cell); }
        for (int s = s1; s < s2; s++) { tab[k].di = d; tab[k].si = s; tab[k++].alpha = (float)(1.0 / cell); }
        if (fs2 - s2 > 1e-3) {
            double t = fs2 - s2; if (t > 1.0) t = 1.0; if (t > cell) t = cell;
            tab[k].di = d; tab[k].si = s2; tab[k++].alpha = (float)(t / cell);
        }
    }
    return k;
}

AVD_O_API int avd_o_area(const uint8_t *gray, int H, int W, int D, uint8_t *out)
{
    if (H < D || W < D) return -1; /* upsampling INTER_AREA is a different library path; not on the hot path */
    double sx = (double)W / D, sy = (double)H / D;
    int isx = (int)sx, isy = (int)sy;
    if (fabs(sx - isx) < 2.220446049250313e-16 && fabs(sy - isy) < 2.220446049250313e-16) {
        float scale = 1.f / (float)(isx * isy);
        for (int dy = 0; dy < D; dy++)
            for (int dx = 0; dx < D; dx++) {
                int sum = 0;
                for (int yy = 0; yy < isy; yy++)
                    for (int xx = 0; xx < isx; xx++)
                        sum += gray[(size_t)(dy * isy + yy) * W + dx * isx + xx];
                out[dy * D + dx] = sat_u8_rne((float)sum * scale);
            }
        return 0;
    }
    area_tap *xt = (area_tap *)malloc(sizeof(area_tap) * (size_t)W * 2);
    area_tap *yt = (area_tap *)malloc(sizeof(area_tap) * (size_t)H * 2);
    float *buf = (float *)malloc(sizeof(float) * D), *sum = (float *)malloc(sizeof(float) * D);
    int nx = area_tab(W, D, sx, xt), ny = area_tab(H, D, sy, yt);
    int prev_dy = yt[0].di;
    for (int d = 0; d < D; d++) sum[d] = 0.f;
    for (int j = 0; j < ny; j++) {
        float beta = yt[j].alpha;
        int dy = yt[j].di;
        const uint8_t *S = gray + (size_t)yt[j].si * W;
        for (int d = 0; d < D; d++) buf[d] = 0.f;
        for (int k = 0; k < nx; k++) {
            float p = (float)S[xt[k].si] * xt[k].alpha; /* separate multiply ... */
            buf[xt[k].di] = buf[xt[k].di] + p;          /* ... and add: no FMA */
        }
        if (dy != prev_dy) {
            for (int d = 0; d < D; d++) { out[prev_dy * D + d] = sat_u8_rne(sum[d]); sum[d] = beta * buf[d]; }
            prev_dy = dy;
        } else {
            for (int d = 0; d < D; d++) { float p = beta * buf[d]; sum[d] = sum[d] + p; }
        }
    }
    for (int d = 0; d < D; d++) out[prev_dy * D + d] = sat_u8_rne(sum[d]);
    free(xt); free(yt); free(buf); free(sum);
    return 0;
}

/* ------------------------------------------------------------------------------------------
 * C3  mean = g.mean(); bits = (g >= mean)                                   video.py:7-8
 * g.mean() is the exact integer sum / 1024 in float64, so g >= mean <=> 1024 g >= sum.
 * Packed little-endian: bit (i & 31) of word (i >> 5), i = row*32 + col.
 * ---------------------------------------------------------------------------------------- */
AVD_O_API void avd_o_hash(const uint8_t *a32, uint32_t *words32)
{
    unsigned sum = 0;
    for (int i = 0; i < 1024; i++) sum += a32[i];
    for (int w = 0; w < 32; w++) {
        uint32_t v = 0;
        for (int b = 0; b < 32; b++) if (1024u * a32[w * 32 + b] >= sum) v |= 1u << b;
        words32[w] = v;
    }
}

/* int(np.sum(h ^ prev))                                                     video.py:38 */
AVD_O_API int avd_o_hamming(const uint32_t *a, const uint32_t *b)
{
    int n = 0;
    for (int w = 0; w < 32; w++) n += __builtin_popcount(a[w] ^ b[w]);
    return n;
}

/* ------------------------------------------------------------------------------------------
 * C4  cv2.resize(gray, (320,320))  (INTER_LINEAR, uint8)                       video.py:43
 * 11-bit fixed-point bilinear, no antialiasing.  640x640 sources take the library's 2x2
 * area shortcut.
 * ---------------------------------------------------------------------------------------- */
static int rne_i(float v) { return (int)lrintf(v); }

AVD_O_API void avd_o_linear_tabs(int ssize, int dsize, int *ofs, int16_t *a0, int16_t *a1, int is_y)
{
    double scale = 1.0 / ((double)dsize / ssize);
    for (int d = 0; d < dsize; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (!is_y) {
            if (s < 0) { f = 0.f; s = 0; }
            if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
        }
        ofs[d] = s;
        a0[d] = (int16_t)rne_i((1.f - f) * 2048.f);
        a1[d] = (int16_t)rne_i(f * 2048.f);
    }
}

AVD_O_API void avd_o_thumb(const uint8_t *gray, int H, int W, int D, uint8_t *out)
{
    if (W == 2 * D && H == 2 * D) {
        for (int y = 0; y < D; y++)
            for (int x = 0; x < D; x++) {
                const uint8_t *p = gray + (size_t)(2 * y) * W + 2 * x;
                out[y * D + x] = (uint8_t)((p[0] + p[1] + p[W] + p[W + 1] + 2) >> 2);
            }
        return;
    }
    int *xo = (int *)malloc(sizeof(int) * D), *yo = (int *)malloc(sizeof(int) * D);
    int16_t *xa0 = (int16_t *)malloc(2 * D), *xa1 = (int16_t *)malloc(2 * D);
    int16_t *ya0 = (int16_t *)malloc(2 * D), *ya1 = (int16_t *)malloc(2 * D);
    avd_o_linear_tabs(W, D, xo, xa0, xa1, 0);
    avd_o_linear_tabs(H, D, yo, ya0, ya1, 1);
    for (int y = 0; y < D; y++) {
        int y0 = yo[y], y1 = yo[y] + 1;
        y0 = y0 < 0 ? 0 : (y0 >= H ? H - 1 : y0);
        y1 = y1 < 0 ? 0 : (y1 >= H ? H - 1 : y1);
        const uint8_t *r0 = gray + (size_t)y0 * W, *r1 = gray + (size_t)y1 * W;
        for (int x = 0; x < D; x++) {
            int x0 = xo[x], x1 = x0 + 1 < W ? x0 + 1 : W - 1;
            int S0 = r0[x0] * xa0[x] + r0[x1] * xa1[x];
            int S1 = r1[x0] * xa0[x] + r1[x1] * xa1[x];
            out[y * D + x] = (uint8_t)((((ya0[y] * (S0 >> 4)) >> 16) + ((ya1[y] * (S1 >> 4)) >> 16) + 2) >> 2);
        }
    }
    free(xo); free(yo); free(xa0); free(xa1); free(ya0); free(ya1);
}

/* ------------------------------------------------------------------------------------------
 * C5  cv2.Laplacian(gray, CV_64F).var()                                        video.py:52
 * ksize=1 aperture [0 1 0; 1 -4 1; 0 1 0], BORDER_REFLECT_101.  The response is an exact
 * integer in [-1020,1020]; the population variance follows from the int64 sums.
 * ---------------------------------------------------------------------------------------- */
static inline int r101(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

AVD_O_API void avd_o_lapstats(const uint8_t *gray, int H, int W, int64_t *sum, int64_t *sumsq)
{
    int64_t s = 0, ss = 0;
    for (int y = 0; y < H; y++) {
        const uint8_t *c = gray + (size_t)y * W;
        const uint8_t *u = gray + (size_t)r101(y - 1, H) * W;
        const uint8_t *d = gray + (size_t)r101(y + 1, H) * W;
        for (int x = 0; x < W; x++) {
            int L = u[x] + d[x] + c[r101(x - 1, W)] + c[r101(x + 1, W)] - 4 * c[x];
            s += L; ss += (int64_t)L * L;
        }
    }
    *sum = s; *sumsq = ss;
}

/* ------------------------------------------------------------------------------------------
 * C6  cv2.calcOpticalFlowFarneback(prev, next, None, 0.5, 3, 15, 3, 5, 1.2, 0)  video.py:45
 * Dense Farneback flow: pyramid (pyr_scale .5, 3 extra levels), polynomial expansion
 * n=5 sigma=1.2, 15x15 box window (flags=0), 3 iterations per level.
 * ---------------------------------------------------------------------------------------- */
static void gauss_kernel(int ksize, double sigma, float *k)
{
    if (sigma <= 0 && ksize == 3) { k[0] = 0.25f; k[1] = 0.5f; k[2] = 0.25f; return; }
    double s = 0, t[64];
    for (int i = 0; i < ksize; i++) { double x = i - (ksize - 1) * 0.5; t[i] = exp(-x * x / (2 * sigma * sigma)); s += t[i]; }
    for (int i = 0; i < ksize; i++) k[i] = (float)(t[i] / s);
}

/* cv2.GaussianBlur on float32, separable symmetric kernel, REFLECT_101: row pass then column pass */
static void gauss_blur(const float *src, int H, int W, int ksize, const float *k, float *tmp, float *dst)
{
    int r = ksize / 2;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const float *s = src + (size_t)y * W;
            float a = k[r] * s[x];
            for (int j = 1; j <= r; j++) { float p = k[r + j] * (s[r101(x - j, W)] + s[r101(x + j, W)]); a = a + p; }
            tmp[(size_t)y * W + x] = a;
        }
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            float a = k[r] * tmp[(size_t)y * W + x];
            for (int j = 1; j <= r; j++) {
                float p = k[r + j] * (tmp[(size_t)r101(y - j, H) * W + x] + tmp[(size_t)r101(y + j, H) * W + x]);
                a = a + p;
            }
            dst[(size_t)y * W + x] = a;
        }
}

/* cv2.resize float32 INTER_LINEAR (cn channels interleaved); exact 2x decimation takes the 2x2 area mean */
static void resize_linear_f32(const float *src, int sH, int sW, int cn, float *dst, int dH, int dW)
{
    if (sH == dH && sW == dW) { memcpy(dst, src, sizeof(float) * (size_t)sH * sW * cn); return; }
    if (sW == 2 * dW && sH == 2 * dH) {
        for (int y = 0; y < dH; y++)
            for (int x = 0; x < dW; x++)
                for (int c = 0; c < cn; c++) {
                    const float *p = src + ((size_t)(2 * y) * sW + 2 * x) * cn + c;
                    dst[((size_t)y * dW + x) * cn + c] = (p[0] + p[cn] + p[(size_t)sW * cn] + p[(size_t)sW * cn + cn]) * 0.25f;
                }
        return;
    }
    double scx = 1.0 / ((double)dW / sW), scy = 1.0 / ((double)dH / sH);
    int *xo = (int *)malloc(sizeof(int) * dW);
    float *xa = (float *)malloc(sizeof(float) * dW * 2);
    int xmax = dW;
    for (int d = 0; d < dW; d++) {
        float f = (float)((d + 0.5) * scx - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        if (s < 0) { f = 0.f; s = 0; }
        if (s + 1 >= sW) { if (d < xmax) xmax = d; if (s >= sW - 1) { f = 0.f; s = sW - 1; } }
        xo[d] = s; xa[2 * d] = 1.f - f; xa[2 * d + 1] = f;
    }
    float *row0 = (float *)malloc(sizeof(float) * (size_t)dW * cn), *row1 = (float *)malloc(sizeof(float) * (size_t)dW * cn);
    for (int y = 0; y < dH; y++) {
        float f = (float)((y + 0.5) * scy - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        float b0 = 1.f - f, b1 = f;
        int y0 = s < 0 ? 0 : (s >= sH ? sH - 1 : s);
        int y1 = s + 1 < 0 ? 0 : (s + 1 >= sH ? sH - 1 : s + 1);
        const float *S[2] = { src + (size_t)y0 * sW * cn, src + (size_t)y1 * sW * cn };
        float *R[2] = { row0, row1 };
        for (int t = 0; t < 2; t++)
            for (int d = 0; d < dW; d++)
                for (int c = 0; c < cn; c++) {
                    if (d < xmax) {
                        float p0 = S[t][xo[d] * cn + c] * xa[2 * d], p1 = S[t][(xo[d] + 1) * cn + c] * xa[2 * d + 1];
                        R[t][d * cn + c] = p0 + p1;
                    } else
                        R[t][d * cn + c] = S[t][xo[d] * cn + c] * 1.f;
                }
        for (int i = 0; i < dW * cn; i++) { float p0 = row0[i] * b0, p1 = row1[i] * b1; dst[(size_t)y * dW * cn + i] = p0 + p1; }
    }
    free(xo); free(xa); free(row0); free(row1);
}

/* Gaussian moments for the polynomial expansion (n=5, sigma=1.2) and the inverse-moment constants */
static void poly_prepare(int n, double sigma, float *g, float *xg, float *xxg, double *ig11, double *ig03, double *ig33, double *ig55)
{
    double s = 0;
    for (int x = -n; x <= n; x++) { g[x] = (float)exp(-x * x / (2 * sigma * sigma)); s += g[x]; }
    s = 1. / s;
    for (int x = -n; x <= n; x++) { g[x] = (float)(g[x] * s); xg[x] = (float)(x * g[x]); xxg[x] = (float)(x * x * g[x]); }
    /* G = sum g(y)g(x) [1 x y x^2 y^2 xy]^T[...]; only four distinct moments are non-zero */
    double a = 0, b = 0, c = 0, d = 0;
    for (int y = -n; y <= n; y++)
        for (int x = -n; x <= n; x++) {
            double w = (double)g[y] * g[x];
            a += w; b += w * x * x; c += w * x * x * x * x; d += w * x * x * y * y;
        }
    /* closed-form inverse of the sparse 6x6: rows/cols {0,3,4} couple, {1},{2},{5} are diagonal */
    *ig11 = 1.0 / b;
    *ig55 = 1.0 / d;
    /* block [[a,b,b],[b,c,d],[b,d,c]] inverse entries (0,1) and (1,1) */
    double det = a * (c * c - d * d) - b * (b * c - b * d) + b * (b * d - b * c);
    *ig03 = -(b * c - b * d) / det;
    *ig33 = (a * c - b * b) / det;
}

static void poly_exp(const float *src, int H, int W, float *dst /* H*W*5 interleaved */)
{
    enum { N = 5 };
    float kb[3 * (2 * N + 1)];
    float *g = kb + N, *xg = g + 2 * N + 1, *xxg = xg + 2 * N + 1;
    double ig11, ig03, ig33, ig55;
    poly_prepare(N, 1.2, g, xg, xxg, &ig11, &ig03, &ig33, &ig55);
    float *rowbuf = (float *)malloc(sizeof(float) * (size_t)(W + 2 * N) * 3);
    float *row = rowbuf + N * 3;
    for (int y = 0; y < H; y++) {
        const float *s0 = src + (size_t)y * W;
        for (int x = 0; x < W; x++) { row[x * 3] = s0[x] * g[0]; row[x * 3 + 1] = row[x * 3 + 2] = 0.f; }
        for (int k = 1; k <= N; k++) {
            const float *a = src + (size_t)(y - k > 0 ? y - k : 0) * W;
            const float *b = src + (size_t)(y + k < H - 1 ? y + k : H - 1) * W;
            for (int x = 0; x < W; x++) {
                float p = a[x] + b[x];
                float m0 = g[k] * p, m1 = xg[k] * (b[x] - a[x]), m2 = xxg[k] * p;
                row[x * 3] = row[x * 3] + m0; row[x * 3 + 1] = row[x * 3 + 1] + m1; row[x * 3 + 2] = row[x * 3 + 2] + m2;
            }
        }
        for (int x = 0; x < N * 3; x++) { row[-1 - x] = row[2 - x]; row[W * 3 + x] = row[W * 3 + x - 3]; }
        float *d = dst + (size_t)y * W * 5;
        for (int x = 0; x < W; x++) {
            double b1 = (double)(row[x * 3] * g[0]), b2 = 0, b3 = (double)(row[x * 3 + 1] * g[0]), b4 = 0,
                   b5 = (double)(row[x * 3 + 2] * g[0]), b6 = 0;
            for (int k = 1; k <= N; k++) {
                double tg = (double)(row[(x + k) * 3] + row[(x - k) * 3]);
                b1 += tg * g[k];
                b4 += tg * xxg[k];
                b2 += (double)((row[(x + k) * 3] - row[(x - k) * 3]) * xg[k]);
                b3 += (double)((row[(x + k) * 3 + 1] + row[(x - k) * 3 + 1]) * g[k]);
                b6 += (double)((row[(x + k) * 3 + 1] - row[(x - k) * 3 + 1]) * xg[k]);
                b5 += (double)((row[(x + k) * 3 + 2] + row[(x - k) * 3 + 2]) * g[k]);
            }
            d[x * 5 + 1] = (float)(b2 * ig11);
            d[x * 5] = (float)(b3 * ig11);
            d[x * 5 + 3] = (float)(b1 * ig03 + b4 * ig33);
            d[x * 5 + 2] = (float)(b1 * ig03 + b5 * ig33);
            d[x * 5 + 4] = (float)(b6 * ig55);
        }
    }
    free(rowbuf);
}

static void update_matrices(const float *R0, const float *R1, const float *flow, float *M, int H, int W)
{
    static const float border[5] = { 0.14f, 0.14f, 0.4472f, 0.4472f, 0.4472f };
    const int B = 5;
    for (int y = 0; y < H; y++)
        for (int x = 0; x < W; x++) {
            const float *r0 = R0 + ((size_t)y * W + x) * 5;
            float dx = flow[((size_t)y * W + x) * 2], dy = flow[((size_t)y * W + x) * 2 + 1];
            float fx = (float)x + dx, fy = (float)y + dy;
            int x1 = (int)floorf(fx), y1 = (int)floorf(fy);
            float r2, r3, r4, r5, r6;
            fx -= (float)x1; fy -= (float)y1;
            if ((unsigned)x1 < (unsigned)(W - 1) && (unsigned)y1 < (unsigned)(H - 1)) {
                const float *p = R1 + ((size_t)y1 * W + x1) * 5;
                size_t st = (size_t)W * 5;
                float a00 = (1.f - fx) * (1.f - fy), a01 = fx * (1.f - fy), a10 = (1.f - fx) * fy, a11 = fx * fy;
#define BIL(c) (((a00 * p[c] + a01 * p[5 + c]) + a10 * p[st + c]) + a11 * p[st + 5 + c])
                r2 = BIL(0); r3 = BIL(1); r4 = BIL(2); r5 = BIL(3); r6 = BIL(4);
#undef BIL
                r4 = (r0[2] + r4) * 0.5f; r5 = (r0[3] + r5) * 0.5f; r6 = (r0[4] + r6) * 0.25f;
            } else {
                r2 = r3 = 0.f; r4 = r0[2]; r5 = r0[3]; r6 = r0[4] * 0.5f;
            }
            r2 = (r0[0] - r2) * 0.5f;
            r3 = (r0[1] - r3) * 0.5f;
            r2 = r2 + (r4 * dy + r6 * dx);
            r3 = r3 + (r6 * dy + r5 * dx);
            if ((unsigned)(x - B) >= (unsigned)(W - 2 * B) || (unsigned)(y - B) >= (unsigned)(H - 2 * B)) {
                float sc = (x < B ? border[x] : 1.f) * (x >= W - B ? border[W - x - 1] : 1.f) *
                           (y < B ? border[y] : 1.f) * (y >= H - B ? border[H - y - 1] : 1.f);
                r2 *= sc; r3 *= sc; r4 *= sc; r5 *= sc; r6 *= sc;
            }
            float *m = M + ((size_t)y * W + x) * 5;
            m[0] = r4 * r4 + r6 * r6;
            m[1] = (r4 + r5) * r6;
            m[2] = r5 * r5 + r6 * r6;
            m[3] = r4 * r2 + r6 * r3;
            m[4] = r6 * r2 + r5 * r3;
        }
}

/* 15x15 box window with replicated borders, float64 sums, then the 2x2 solve */
static void blur_solve(const float *M, float *flow, int H, int W, int bs)
{
    int m = bs / 2;
    double scale = 1.0 / (bs * bs);
    double *vs = (double *)malloc(sizeof(double) * (size_t)W * 5);
    for (int y = 0; y < H; y++) {
        for (int i = 0; i < W * 5; i++) vs[i] = 0;
        for (int k = -m; k <= m; k++) {
            int yy = y + k; yy = yy < 0 ? 0 : (yy > H - 1 ? H - 1 : yy);
            const float *r = M + (size_t)yy * W * 5;
            for (int i = 0; i < W * 5; i++) vs[i] += r[i];
        }
        for (int x = 0; x < W; x++) {
            double a[5] = { 0, 0, 0, 0, 0 };
            for (int k = -m; k <= m; k++) {
                int xx = x + k; xx = xx < 0 ? 0 : (xx > W - 1 ? W - 1 : xx);
                for (int c = 0; c < 5; c++) a[c] += vs[xx * 5 + c];
            }
            double g11 = a[0] * scale, g12 = a[1] * scale, g22 = a[2] * scale, h1 = a[3] * scale, h2 = a[4] * scale;
            double idet = 1. / (g11 * g22 - g12 * g12 + 1e-3);
            flow[((size_t)y * W + x) * 2] = (float)((g11 * h2 - g12 * h1) * idet);
            flow[((size_t)y * W + x) * 2 + 1] = (float)((g22 * h1 - g12 * h2) * idet);
        }
    }
    free(vs);
}

static int cv_round(double v) { return (int)lrint(v); }

/* prev/next: H x W uint8; flow out: H x W x 2 float32.  Fixed parameters of video.py:45. */
AVD_O_API void avd_o_farneback(const uint8_t *prev, const uint8_t *next, int H, int W, float *flow_out)
{
    const double pyr_scale = 0.5; const int winsize = 15, iters = 3; int levels = 3;
    const int min_size = 32;
    int k; double scale;
    for (k = 0, scale = 1; k < levels; k++) { scale *= pyr_scale; if (W * scale < min_size || H * scale < min_size) break; }
    levels = k;
    size_t N = (size_t)H * W;
    float *fimg = (float *)malloc(sizeof(float) * N), *blur = (float *)malloc(sizeof(float) * N), *tmp = (float *)malloc(sizeof(float) * N);
    float *I = (float *)malloc(sizeof(float) * N);
    float *R[2] = { (float *)malloc(sizeof(float) * N * 5), (float *)malloc(sizeof(float) * N * 5) };
    float *M = (float *)malloc(sizeof(float) * N * 5);
    float *flow = (float *)malloc(sizeof(float) * N * 2), *pflow = (float *)malloc(sizeof(float) * N * 2);
    int pw = 0, ph = 0;
    const uint8_t *img[2] = { prev, next };
    for (k = levels; k >= 0; k--) {
        int i;
        for (i = 0, scale = 1; i < k; i++) scale *= pyr_scale;
        double sigma = (1. / scale - 1) * 0.5;
        int ksz = cv_round(sigma * 5) | 1; if (ksz < 3) ksz = 3;
        int w = cv_round(W * scale), h = cv_round(H * scale);
        if (pw == 0) memset(flow, 0, sizeof(float) * (size_t)w * h * 2);
        else {
            resize_linear_f32(pflow, ph, pw, 2, flow, h, w);
            float mul = (float)(1. / pyr_scale);
            for (size_t j = 0; j < (size_t)w * h * 2; j++) flow[j] *= mul;
        }
        float kern[64];
        gauss_kernel(ksz, sigma, kern);
        for (i = 0; i < 2; i++) {
            for (size_t j = 0; j < N; j++) fimg[j] = (float)img[i][j];
            gauss_blur(fimg, H, W, ksz, kern, tmp, blur);
            resize_linear_f32(blur, H, W, 1, I, h, w);
            poly_exp(I, h, w, R[i]);
        }
        update_matrices(R[0], R[1], flow, M, h, w);
        for (i = 0; i < iters; i++) {
            blur_solve(M, flow, h, w, winsize);
            if (i < iters - 1) update_matrices(R[0], R[1], flow, M, h, w);
        }
        memcpy(pflow, flow, sizeof(float) * (size_t)w * h * 2); pw = w; ph = h;
    }
    memcpy(flow_out, flow, sizeof(float) * N * 2);
    free(fimg); free(blur); free(tmp); free(I); free(R[0]); free(R[1]); free(M); free(flow); free(pflow);
}

/* ------------------------------------------------------------------------------------------
 * C7  mag = np.sqrt(fx**2 + fy**2); np.mean(mag); np.var(mag)               video.py:46-48
 * float32 throughout, NumPy pairwise summation (8-lane blocks of <=128, recursive halving).
 * ---------------------------------------------------------------------------------------- */
static float pairwise_f32(const float *a, size_t n)
{
    if (n < 8) { float r = 0.f; for (size_t i = 0; i < n; i++) r += a[i]; return r; }
    if (n <= 128) {
        float r[8]; size_t i;
        for (int j = 0; j < 8; j++) r[j] = a[j];
        for (i = 8; i < n - (n % 8); i += 8) for (int j = 0; j < 8; j++) r[j] += a[i + j];
        float res = ((r[0] + r[1]) + (r[2] + r[3])) + ((r[4] + r[5]) + (r[6] + r[7]));
        for (; i < n; i++) res += a[i];
        return res;
    }
    size_t n2 = n / 2; n2 -= n2 % 8;
    return pairwise_f32(a, n2) + pairwise_f32(a + n2, n - n2);
}

AVD_O_API void avd_o_magstats(const float *flow, size_t npix, float *mean, float *var)
{
    float *mag = (float *)malloc(sizeof(float) * npix);
    for (size_t i = 0; i < npix; i++) { float a = flow[2 * i] * flow[2 * i], b = flow[2 * i + 1] * flow[2 * i + 1]; mag[i] = sqrtf(a + b); }
    float mu = pairwise_f32(mag, npix) / (float)npix;
    *mean = mu;
    for (size_t i = 0; i < npix; i++) { float d = mag[i] - mu; mag[i] = d * d; }
    *var = pairwise_f32(mag, npix) / (float)npix;
    free(mag);
}

/* ------------------------------------------------------------------------------------------
 * One sampled frame, all per-frame features (video.py:36,43,51-52).  thumb may be NULL.
 * ---------------------------------------------------------------------------------------- */
AVD_O_API int avd_o_frame_features(const uint8_t *bgr, int H, int W, size_t pitch, uint8_t *gray_scratch,
                                   uint8_t *area32, uint32_t *hash32, uint8_t *thumb320, int64_t *lap_sum, int64_t *lap_sumsq)
{
    avd_o_gray(bgr, H, W, pitch, gray_scratch);
    if (avd_o_area(gray_scratch, H, W, 32, area32)) return -1;
    avd_o_hash(area32, hash32);
    if (thumb320) avd_o_thumb(gray_scratch, H, W, 320, thumb320);
    avd_o_lapstats(gray_scratch, H, W, lap_sum, lap_sumsq);
    return 0;
}
