/* Plain-C client of the C ABI (include/avd.h): proves the boundary needs nothing but a C compiler and libavd.so, and is
 * the small, Python-free program that compute-sanitizer runs on (scripts/gpu_sanitize.sh).
 *   gcc -std=c11 -O1 -Iinclude tests/native/abi_smoke.c -Lavd_b200/csrc -lavd -Wl,-rpath,$PWD/avd_b200/csrc -o tests/native/abi_smoke
 * Exercises: pinned allocation, the BGR host path on a ragged shape (any-width kernel), a W%16 shape (banded kernel) and
 * 720p (streaming kernel) incl. Farneback, the 4:2:0 host path, determinism (two runs, identical records), duplicates. */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include "avd.h"

static unsigned lcg(unsigned *s) { *s = *s * 1664525u + 1013904223u; return *s >> 8; }

static int run_shape(avd_ctx *ctx, int H, int W, int n, int yuv)
{
    const size_t fb = yuv ? (size_t)H * W * 3 / 2 : (size_t)H * W * 3;
    uint8_t *buf = (uint8_t *)avd_host_alloc(ctx, fb * n);
    if (!buf) { fprintf(stderr, "host alloc failed\n"); return 1; }
    unsigned seed = 12345u + (unsigned)(H * 31 + W);
    /* smooth-ish content: a pattern that shifts by 3 pixels per frame, plus noise; frame 2 duplicates frame 1 */
    for (int f = 0; f < n; f++) {
        uint8_t *p = buf + fb * f;
        if (f == 2) { memcpy(p, p - fb, fb); continue; }
        for (size_t i = 0; i < fb; i++) {
            size_t px = yuv ? i : i / 3;
            int x = (int)(px % (size_t)W), y = (int)(px / (size_t)W);
            p[i] = (uint8_t)(((x + 3 * f) * 5 + y * 3 + ((x + 3 * f) / 16 + y / 16) * 37 + (lcg(&seed) & 7)) & 255);
        }
    }
    const uint8_t **frames = (const uint8_t **)malloc(sizeof(*frames) * n);
    for (int f = 0; f < n; f++) frames[f] = buf + fb * f;
    avd_sample_rec *r1 = (avd_sample_rec *)calloc(n, sizeof(*r1)), *r2 = (avd_sample_rec *)calloc(n, sizeof(*r2));
    int rc = 0;
    for (int rep = 0; rep < 2 && !rc; rep++) {
        avd_sample_rec *r = rep ? r2 : r1;
        rc = yuv ? avd_analyze_host_yuv420(ctx, frames, n, H, W, 0, r) : avd_analyze_host(ctx, frames, n, H, W, (size_t)W * 3, r);
        if (rc) fprintf(stderr, "%dx%d %s: status %d: %s\n", W, H, yuv ? "yuv" : "bgr", rc, avd_last_error(ctx));
    }
    if (!rc && memcmp(r1, r2, sizeof(*r1) * n)) { fprintf(stderr, "%dx%d: two runs differ\n", W, H); rc = 1; }
    if (!rc && (r1[0].ham_prev != -1 || r1[2].ham_prev != 0 || !(r1[1].flags & AVD_REC_HAS_FLOW) || r1[0].flags != 0)) {
        fprintf(stderr, "%dx%d: record invariants violated (ham0 %d ham2 %d)\n", W, H, r1[0].ham_prev, r1[2].ham_prev); rc = 1;
    }
    if (!rc) printf("%4dx%-4d %s n=%d  lap_sumsq[0]=%lld  flow_mean[1]=%.6f  ham[1]=%d  ok\n", W, H, yuv ? "i420" : "bgr ", n,
                    (long long)r1[0].lap_sumsq, r1[1].flow_mean, r1[1].ham_prev);
    free(r1); free(r2); free(frames);
    avd_host_free(ctx, buf);
    return rc;
}

int main(void)
{
    avd_ctx *ctx = NULL;
    int rc = avd_create(0, &ctx);
    if (rc) { fprintf(stderr, "avd_create: status %d (needs an sm_100 GPU; there is no CPU fallback)\n", rc); return 2; }
    printf("%s\n", avd_version());
    int bad = 0;
    bad |= run_shape(ctx, 120, 203, 5, 0);      /* any-width kernel */
    bad |= run_shape(ctx, 272, 480, 5, 0);      /* banded kernel */
    bad |= run_shape(ctx, 720, 1280, 4, 0);     /* streaming kernel */
    bad |= run_shape(ctx, 720, 1280, 4, 1);     /* 4:2:0 upload + device colour conversion */
    bad |= run_shape(ctx, 66, 34, 3, 1);        /* 4:2:0, byte kernel, tiny */
    if (avd_host_alloc_count(ctx) != 0) { fprintf(stderr, "pinned regions leaked\n"); bad = 1; }
    if (avd_check_guards(ctx) != 0) { fprintf(stderr, "guard zones: %s\n", avd_last_error(ctx)); bad = 1; }
    avd_destroy(ctx);
    printf(bad ? "FAILED\n" : "abi_smoke ok\n");
    return bad;
}
