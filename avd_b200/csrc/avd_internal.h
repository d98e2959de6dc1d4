// Internal declarations shared by the translation units of libavd.so (not part of the C ABI).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>
#include <map>
#include <set>
#include <utility>
#include <vector>
#include "../../include/avd.h"

#define AVD_D 32            // aHash side
#define AVD_T 320           // thumbnail side
#define AVD_FMT_BGR24 0
#define AVD_FMT_I420 1         // yuv420p, limited range (H.264 / HEVC / MPEG-4 decoders)
#define AVD_FMT_J420 2         // yuvj420p, full range (MJPEG)
#define AVD_LEVELS 4        // Farneback pyramid: 40, 80, 160, 320

// One INTER_AREA destination cell along one axis (OpenCV computeResizeAreaTab, restated):
// taps in order: optional head (index s1-1, weight ah), s in [s1,s2) weight am, optional tail (index s2, weight at).
struct AreaCell {
    int s1, s2;
    float ah, am, at;
    int has_head, has_tail;
};

// INTER_LINEAR (8-bit, 11-bit fixed point) tap of one destination index along one axis.
struct LinTap {
    int i0, i1;      // source indices (already clamped / clipped)
    int a0, a1;      // weights, sum 2048
};

// Per-(H,W) resampling tables, resident in device memory.
struct ShapeTables {
    AreaCell ax[AVD_D], ay[AVD_D];
    LinTap lx[AVD_T], ly[AVD_T];
    int area_int;            // 1: both scales integer -> block-sum path
    int isx, isy;            // integer scales (area_int)
    float inv_area;          // float(1/(isx*isy))
    int thumb_2x2;           // source is exactly 640x640 -> 2x2 mean
    int row_owner_begin[AVD_D + 1];  // first source row owned by band b (fused kernel); [32] = H
    int thumb_row_begin[AVD_D + 1];  // first thumbnail row whose top tap lies in band b; [32] = 320
    // streaming kernel: thumbnail columns whose left tap lies in the 16-pixel strip s (first column, count)
    uint16_t strip_tx0[256];
    uint8_t strip_ntx[256];
    int strip_ntx_max;
    // streaming kernel: role of source row r as a thumbnail tap row (0 = none): bit0 tap, bit1 bottom tap of its
    // pair, bits 2-13 vertical weight, bits 16-24 thumbnail row.  Filled only when every row has at most one role.
    uint32_t rowinfo[4096];
};

struct avd_ctx {
    int device = 0;
    int sm_count = 0;
    cudaStream_t compute = nullptr, copy = nullptr;
    std::string err;
    uint64_t launches = 0;
    uint64_t ws_generation = 0;                            // bumped whenever a grow-only device buffer moves (avd_ensure)
    struct Guarded { void *base, *user; size_t cap; };
    std::vector<Guarded> guarded;                          // grow-only device buffers with their guard zones (avd_ensure)
    std::set<void *> host_allocs;                          // avd_host_alloc regions not yet freed (released by avd_destroy)
    int k1_mode = 0;
    std::map<std::pair<int, int>, ShapeTables *> tables;   // device pointers
    std::map<std::pair<int, int>, ShapeTables> tables_host;
    // generic K1 scratch
    uint8_t *gray = nullptr; size_t gray_cap = 0;
    float *area_rows = nullptr; size_t area_rows_cap = 0;
    // K2 workspace (see k2_flow.cu)
    void *k2_ws = nullptr; size_t k2_ws_cap = 0; int k2_batch = 0;
    cudaStream_t k2_side[3] = {}; cudaEvent_t k2_ev[6] = {};     // further chains of a split batch (AVD_K2_SPLIT)
    cudaStream_t k2_img = nullptr; cudaEvent_t k2_img_ev[5] = {};  // pyramid + polynomial expansion of the finer levels beside the coarse flow levels
    // host streaming ring (avd_analyze_host)
    uint8_t *ring = nullptr; size_t ring_cap = 0;
    uint8_t *bgr_slot = nullptr; size_t bgr_slot_cap = 0;   // YUV input: the converted frames of the slot in flight
    cudaEvent_t ev_copied[4] = {}, ev_free[4] = {};
    uint8_t *thumbs = nullptr; size_t thumbs_cap = 0;
    avd_sample_rec *recs = nullptr; size_t recs_cap = 0; int recs_n = 0;   // recs_n: records of the last host run
    void *k3_ws = nullptr; size_t k3_ws_cap = 0;               // avd_score_clips workspace
    void *k4_ws = nullptr; size_t k4_ws_cap = 0;               // avd_audio_windows workspace
    uint8_t *area32 = nullptr; size_t area32_cap = 0;
    uint32_t *fixlist = nullptr; size_t fixlist_cap = 0;   // streaming K1: [0] = count, then uncertified cells
    // optional per-kernel timing (avd_profile_enable): CUDA events around every launch, on its stream
    bool profiling = false;
    struct ProfRec { const char *name; cudaEvent_t a, b; };
    std::vector<ProfRec> prof;
    std::vector<cudaEvent_t> prof_pool;
};

#define AVD_CUDA(ctx, call)                                                                   \
    do {                                                                                      \
        cudaError_t e__ = (call);                                                             \
        if (e__ != cudaSuccess) {                                                             \
            (ctx)->err = std::string(#call) + ": " + cudaGetErrorString(e__);                 \
            cudaGetLastError();                                                               \
            return AVD_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

#define AVD_LAUNCH_CHECK(ctx, name)                                                           \
    do {                                                                                      \
        (ctx)->launches++;                                                                    \
        cudaError_t e__ = cudaGetLastError();                                                 \
        if (e__ != cudaSuccess) {                                                             \
            (ctx)->err = std::string("launch ") + name + ": " + cudaGetErrorString(e__);      \
            return AVD_ERR_CUDA;                                                              \
        }                                                                                     \
    } while (0)

void avd_prof_begin(avd_ctx *ctx, const char *name, cudaStream_t st);
void avd_prof_end(avd_ctx *ctx, cudaStream_t st);

// AVD_LAUNCH(ctx, "name", stream, kernel<<<grid, block, smem, stream>>>(args...));
#define AVD_LAUNCH(ctx, name, st, ...)                                                        \
    do {                                                                                      \
        if ((ctx)->profiling) avd_prof_begin(ctx, name, st);                                  \
        __VA_ARGS__;                                                                          \
        if ((ctx)->profiling) avd_prof_end(ctx, st);                                          \
        AVD_LAUNCH_CHECK(ctx, name);                                                          \
    } while (0)

int avd_ensure(avd_ctx *ctx, void **p, size_t *cap, size_t need);   // grow-only device buffer
int avd_get_tables(avd_ctx *ctx, int H, int W, const ShapeTables **dev, const ShapeTables **host);

// k1_generic.cu
int k1_rec_init(avd_ctx *ctx, avd_sample_rec *d_rec, int n, uint32_t *d_fixcount, cudaStream_t st);
int k1_generic(avd_ctx *ctx, const uint8_t *d_bgr, int n, int H, int W, size_t pitch, size_t fstride,
               const ShapeTables *tab, const ShapeTables *htab, avd_sample_rec *d_rec, uint8_t *d_thumbs,
               uint8_t *d_area32, cudaStream_t st);
int k1_gray_only(avd_ctx *ctx, const uint8_t *d_bgr, int H, int W, size_t pitch, uint8_t *d_gray, cudaStream_t st);
// fix != NULL: the cells listed by the streaming kernel are first recomputed from the frames (exact float32 order)
struct K1FixArgs { const uint8_t *bgr; size_t pitch, fstride; const ShapeTables *tab; const uint32_t *list; };
int k1_hash_finalize(avd_ctx *ctx, uint8_t *d_area32, int n, const uint32_t *d_prev_hash,
                     avd_sample_rec *d_rec, const K1FixArgs *fix, cudaStream_t st);
// k1_fused.cu (streaming kernel: W % 64 == 0, 640 <= W <= 4096, 640 <= H <= 4096) and k1_banded.cu (W % 16 == 0)
bool k1_fused_supported(int H, int W, size_t pitch, size_t fstride, const void *base, const ShapeTables *htab);
bool k1_banded_supported(int H, int W, size_t pitch, size_t fstride, const void *base);
int k1_banded(avd_ctx *ctx, const uint8_t *d_bgr, int n, int H, int W, size_t pitch, size_t fstride,
              const ShapeTables *tab, const ShapeTables *htab, avd_sample_rec *d_rec, uint8_t *d_thumbs,
              uint8_t *d_area32, cudaStream_t st);
int k1_fused(avd_ctx *ctx, const uint8_t *d_bgr, int n, int H, int W, size_t pitch, size_t fstride,
             const ShapeTables *tab, const ShapeTables *htab, avd_sample_rec *d_rec, uint8_t *d_thumbs,
             uint8_t *d_area32, uint32_t *d_fixlist, cudaStream_t st);
// k0_yuv.cu: dense I420 frames (Y HxW, U, V (H/2)x(W/2)) -> packed BGR, swscale's unscaled yuv420p->bgr24 arithmetic
int k0_yuv420_to_bgr(avd_ctx *ctx, const uint8_t *d_yuv, size_t yuv_stride, int n, int H, int W, bool full_range,
                     uint8_t *d_bgr, size_t pitch, size_t fstride, cudaStream_t st);
// k1_rows.cu (any width / alignment; the band's gray rows in shared memory)
bool k1_rows_supported(int H, int W, const ShapeTables *htab);
int k1_rows(avd_ctx *ctx, const uint8_t *d_bgr, int n, int H, int W, size_t pitch, size_t fstride,
            const ShapeTables *tab, const ShapeTables *htab, avd_sample_rec *d_rec, uint8_t *d_thumbs,
            uint8_t *d_area32, uint32_t *d_fixlist, cudaStream_t st);
// k2_flow.cu
int k2_reserve(avd_ctx *ctx, int n_pairs);   // size the workspace for runs of up to n_pairs pairs now
int k2_flow(avd_ctx *ctx, const uint8_t *d_prev, const uint8_t *d_next, size_t prev_stride, size_t next_stride,
            int n_pairs, bool shared_chain, avd_sample_rec *d_rec_first, float *d_flow_out, cudaStream_t st);
