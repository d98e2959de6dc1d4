/*
 * avd.h -- C ABI of libavd.so, the B200 (sm_100a) implementation of the per-frame analyzer hot path of
 * backtato/ai-video-detector.
 *
 * The reference has no FFI: its seam is the Python call  video_an.analyze(path, meta)  (api.py:133,
 * defined at app/analyzers/video.py:10).  Everything that call computes per sampled frame is a library
 * call into OpenCV/NumPy; the entry points below replace exactly those calls and are what a maintainer
 * binds with ctypes (INTEGRATION.md shows the stub).  Each entry cites the reference lines it replaces.
 *
 * Conventions: plain pointers and sizes, no C++/torch types; every function returns 0 on success or a
 * negative avd_status, never throws, never aborts; the message for the last failure on a context is
 * avd_last_error(ctx).  The caller owns every buffer it passes; the library owns its streams, staging
 * rings and workspaces.  "d_" pointers are device memory on the context's device, "h_" pointers are host
 * memory (pinned for full PCIe rate: avd_host_alloc).  `stream` is a cudaStream_t passed as void* (NULL =
 * CUDA's default stream, as usual); device-pointer entry points are asynchronous on that stream.
 * A context may be used by one thread at a time; create one per (thread, device) (api.py:133 runs the
 * analyzer on pool threads).
 */
#ifndef AVD_H
#define AVD_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AVD_THUMB 320      /* video.py:43  resize(..., (320,320)) */
#define AVD_HASH_SIDE 32   /* video.py:36  _average_hash(frame, size=32) */

typedef enum avd_status {
    AVD_OK = 0,
    AVD_ERR_ARG = -1,       /* bad argument (NULL, n < 0, H or W < 32, pitch too small, ...) */
    AVD_ERR_CUDA = -2,      /* a CUDA runtime call or kernel failed; see avd_last_error */
    AVD_ERR_NOMEM = -3,     /* device or pinned allocation failed */
    AVD_ERR_NODEVICE = -4   /* no usable sm_100 device */
} avd_status;

typedef struct avd_ctx avd_ctx;

/* One record per sampled frame: everything video.py:36-52 derives from that frame (and its predecessor).
 * 160 bytes; also the element of the multi-GPU all-gather. */
typedef struct avd_sample_rec {
    int64_t  lap_sum;      /* sum   of cv2.Laplacian(gray, CV_64F)            video.py:52 */
    int64_t  lap_sumsq;    /* sum of squares of the same; var = (ss - s*s/N)/N video.py:52 */
    float    flow_mean;    /* float(np.mean(mag)) vs previous sample           video.py:47 ; 0 if none */
    float    flow_var;     /* float(np.var(mag))                               video.py:48 ; 0 if none */
    uint32_t hash[32];     /* (g >= g.mean()) bit i&31 of word i>>5, i=row*32+col  video.py:4-8 */
    int32_t  ham_prev;     /* int(np.sum(hsh ^ prev_hash)); -1 if no predecessor   video.py:38 */
    uint32_t flags;        /* bit0: has predecessor hash; bit1: has flow stats */
} avd_sample_rec;

#define AVD_REC_HAS_PREV 1u
#define AVD_REC_HAS_FLOW 2u

const char *avd_version(void);
int  avd_device_count(void);

int  avd_create(int device, avd_ctx **out);
void avd_destroy(avd_ctx *ctx);
const char *avd_last_error(const avd_ctx *ctx);
int  avd_sync(avd_ctx *ctx);                         /* waits for all work of the context */

/* pinned host memory for frame batches (cudaHostAlloc).  Regions still allocated when the context is destroyed
 * are released by avd_destroy; avd_host_alloc_count is the number of live regions (leak checks). */
void *avd_host_alloc(avd_ctx *ctx, size_t bytes);
void  avd_host_free(avd_ctx *ctx, void *p);
size_t avd_host_alloc_count(const avd_ctx *ctx);

/*
 * K1 -- per-frame features of n BGR frames resident in HBM.   Replaces, per frame:
 *   cv2.cvtColor(BGR2GRAY)                       video.py:5,43,51
 *   cv2.resize(g,(32,32),INTER_AREA) + g>=mean   video.py:6-8      -> rec.hash   (d_area32: the 32x32 u8, optional)
 *   int(np.sum(hsh ^ prev_hash))                 video.py:38       -> rec.ham_prev
 *   cv2.resize(gray,(320,320))                   video.py:43       -> d_thumbs[i] (320*320 u8)
 *   cv2.Laplacian(gray,CV_64F) sums              video.py:52       -> rec.lap_sum / lap_sumsq
 * Frame i starts at d_bgr + i*frame_stride, row y at + y*row_pitch, pixel = B,G,R bytes.
 * d_prev_hash: 32 words of the sample before frame 0 (chunk halo), or NULL.
 * Writes all fields of d_rec[0..n) (flow fields zeroed).
 */
int avd_frames_features(avd_ctx *ctx, const uint8_t *d_bgr, int n, int H, int W, size_t row_pitch,
                        size_t frame_stride, const uint32_t *d_prev_hash, avd_sample_rec *d_rec,
                        uint8_t *d_thumbs, uint8_t *d_area32, void *stream);

/*
 * K2 -- Farneback flow magnitude statistics between consecutive thumbnails.   Replaces
 *   cv2.calcOpticalFlowFarneback(prev, small, None, 0.5, 3, 15, 3, 5, 1.2, 0)   video.py:45
 *   mag = sqrt(fx^2+fy^2); np.mean(mag); np.var(mag)                              video.py:46-48
 * for the pairs (d_thumbs[i-1], d_thumbs[i]), i = 1..n-1, and, if d_prev_thumb != NULL, the pair
 * (d_prev_thumb, d_thumbs[0]).  Results go to d_rec[i].flow_mean / flow_var and flags |= HAS_FLOW.
 */
int avd_flow_stats(avd_ctx *ctx, const uint8_t *d_thumbs, int n, const uint8_t *d_prev_thumb,
                   avd_sample_rec *d_rec, void *stream);

/* Parity/debug: the dense flow fields themselves.  Pair p = (d_prev + p*102400, d_next + p*102400);
 * d_flow[p] is 320 x 320 x 2 float32 (x,y interleaved, the layout cv2 returns). */
int avd_flow_fields(avd_ctx *ctx, const uint8_t *d_prev, const uint8_t *d_next, int n_pairs,
                    float *d_flow, void *stream);

/* Parity/debug: the gray plane of one device-resident frame (H*W u8, dense). */
int avd_gray(avd_ctx *ctx, const uint8_t *d_bgr, int H, int W, size_t row_pitch, uint8_t *d_gray, void *stream);

/*
 * Whole hot path from host memory -- what video.analyze's sampled-frame loop (video.py:31-52) does for
 * a run of n sampled frames.  h_frames[i] points at frame i (H rows of row_pitch bytes).  Frames are
 * streamed to the device in chunks on a copy stream while K1/K2 run on the compute stream; records come
 * back to h_rec[0..n).  For a chunk of a longer clip (multi-GPU frame-chunk sharding, or host-side
 * streaming of a long clip) the caller passes the sample *before* the chunk as frame 0 (one-sample halo)
 * and drops h_rec[0].  Synchronous: returns when h_rec is complete.  h_rec = NULL leaves the records on the device
 * (for avd_score_clips).
 */
int avd_analyze_host(avd_ctx *ctx, const uint8_t *const *h_frames, int n, int H, int W, size_t row_pitch,
                     avd_sample_rec *h_rec);

/*
 * Decode -> device, first step: the decoder's own planar 4:2:0 pictures cross PCIe (1.5 bytes
 * per pixel instead of 3) and the colour conversion cv2.VideoCapture.retrieve() performs on the host (video.py:32;
 * sws_scale yuv420p/yuvj420p -> bgr24, unscaled) runs on the device, bit for bit: libswscale's 16-bit fixed-point
 * yuv2rgb arithmetic with the ITU-601 default coefficients, chroma replicated 2x2.  A frame is dense I420:
 * H rows of W luma bytes, then (H/2) x (W/2) U, then (H/2) x (W/2) V; W and H even.  full_range = 1 for yuvj420p (MJPEG).
 *
 * avd_yuv420_to_bgr: n device-resident frames (frame i at d_yuv + i*yuv_frame_stride) -> packed BGR frames.
 * avd_analyze_host_yuv420: avd_analyze_host for 4:2:0 host frames; records equal those of avd_analyze_host on the BGR
 * frames the reference's decoder would have produced from the same pictures.
 */
int avd_yuv420_to_bgr(avd_ctx *ctx, const uint8_t *d_yuv, size_t yuv_frame_stride, int n, int H, int W, int full_range,
                      uint8_t *d_bgr, size_t row_pitch, size_t frame_stride, void *stream);
int avd_analyze_host_yuv420(avd_ctx *ctx, const uint8_t *const *h_frames, int n, int H, int W, int full_range,
                            avd_sample_rec *h_rec);

/*
 * K3 -- the per-clip scoring tail on the device, for batches of clips whose records are already in HBM (SURVEY.md
 * section 8(f)-3).  One call scores n_clips clips; nothing but the compact scores and timelines crosses PCIe.  Replaces
 *   ai_susp per sample, dup_density, scene_change_rate, flow_mean, flow_var, texture_var, timeline pad/truncate   video.py:54-81
 *   fuse: weights, agreement, bonuses, fused timeline, peaks, score, spread, confidence, label, 3-tap binning,
 *         round(score, 2), round(conf, 2)                                                                       fusion.py:7-14,16-109
 * in float64 with NumPy's operation order (pairwise sums of np.mean / np.var / np.std, the two-pass variance, the
 * accumulation order of np.convolve(..., "same")), so results are equal to the reference's, not merely close.
 * What stays with the caller because it is string work: hints -> `pen` (fusion.py:37-41) and label/conditions -> `reason`.
 *
 * A job names the clip's records (the first is treated as having no predecessor, video.py:20,26), its timeline
 * length tlen = int(max(1, round(duration))) (Python round, video.py:73) and -- if do_fuse -- the audio side:
 * audio timeline = h_tl_in[audio_off .. audio_off+audio_len), flags speech_ratio / tts_like, penalty `pen`.
 * video_given = 1 scores a video timeline and summary supplied by the caller instead of records (any dict fuse()
 * accepts).  Outputs per clip start at out_off and are L = max(audio_len, video timeline length, 1) long (tlen when
 * do_fuse = 0): h_vtimeline = the video timeline as fuse leaves it (padded in place, fusion.py:20-21; its first
 * score.vlen entries are video.analyze's timeline), h_fused = the clipped fused timeline, h_binned, h_peaks (1 = peak).
 * d_rec = NULL scores the records of the last avd_analyze_host run on this context (which may be given h_rec = NULL
 * to keep the records on the device).  Synchronous on `stream`.
 */
typedef struct avd_clip_job {
    int32_t rec_first, rec_count;      /* records of the clip inside d_rec, sample order (count >= 0) */
    int32_t npix;                      /* H*W of its frames */
    int32_t tlen;                      /* >= 1 */
    int32_t do_fuse;
    int32_t audio_off, audio_len;
    int32_t video_given, vgiven_off, vgiven_len;
    int32_t out_off, pad_;
    double  speech_ratio, tts_like, pen;
    double  thresh_real_max, thresh_ai_min;                        /* fusion.py:4-5 */
    double  g_flow_mean, g_texture_var, g_sc_rate, g_dup_density;  /* video_given: the summary fields fuse reads */
} avd_clip_job;

typedef struct avd_clip_score {
    double  dup_density, scene_change_rate, flow_mean, flow_var, texture_var;   /* video.py:64-69 */
    double  score, spread, conf;         /* before rounding */
    double  ai_score, confidence;        /* round(score, 2), round(conf, 2)   fusion.py:103-104 */
    int32_t label;                       /* 0 real, 1 uncertain, 2 ai          fusion.py:82-100 */
    int32_t vlen, L, n_peaks;
} avd_clip_score;

int avd_score_clips(avd_ctx *ctx, const avd_sample_rec *d_rec, int n_rec, const avd_clip_job *h_jobs, int n_clips,
                    const double *h_tl_in, int n_tl_in, avd_clip_score *h_scores, double *h_vtimeline,
                    double *h_fused, double *h_binned, uint8_t *h_peaks, int n_out, void *stream);

/*
 * K4 -- per-window features of the audio analyzer.  h_wav: the mono float32 samples the
 * reference's decode helper returns (audio.py:7-20, 16 kHz); win = max(1, int(sr * 0.5)) samples (audio.py:38; at most
 * 9000).  For every window w = 0 .. ceil(n_samples / win) - 1 (the last one ragged) writes five float64 values to
 * h_feats[5 w ..]: what audio.py:43-60 appends to rms, zcr, flat, roll, sc_cent:
 *   float(np.sqrt((seg**2).mean()))                      float32 arithmetic, exact
 *   float(np.mean(np.abs(np.diff(np.sign(seg)))) / 2.0)  exact (NaN for a one-sample window, like the reference)
 *   flatness  exp(mean(log(mag))) / mean(mag),  mag = |np.fft.rfft(seg * np.hanning(len))| + 1e-9
 *   roll-off  first bin whose running sum reaches 0.85 * sum(mag), divided by the number of bins
 *   centroid  sum(linspace(0, 1, bins) * mag) / sum(mag)
 * The spectrum is float64: a two-stage DFT for full windows whose length factors as N1 x N2 <= 128 x 128 (8000 = 80 x 100),
 * a direct DFT for the ragged last window (any length, also prime); twiddles are correctly rounded.  The
 * O(windows) tail (percentile, variances, tts_like, timeline; audio.py:63-111) is host arithmetic.  Synchronous.
 */
int avd_audio_windows(avd_ctx *ctx, const float *h_wav, long long n_samples, int win, double *h_feats, void *stream);

/* Counters for bench.py: number of kernel launches issued by this context since creation. */
uint64_t avd_launch_count(const avd_ctx *ctx);

/* The library's scratch buffers (tables, Farneback workspace, fix-up lists) only ever grow; when one has to move, this
 * counter is bumped.  A caller that captured device-pointer entry points into a CUDA graph must re-capture when the
 * value differs from the one it saw at capture time (the graph holds the old addresses). */
uint64_t avd_ws_generation(const avd_ctx *ctx);

/* Every device buffer the library owns sits between two 256-byte guard zones filled with a canary.  Returns the number of
 * guard bytes that no longer hold it (0 = no kernel wrote outside a workspace), or a negative avd_status.  Synchronises. */
long long avd_check_guards(avd_ctx *ctx);

/* Per-kernel device timing for bench.py: when enabled, every launch is bracketed by CUDA events on its
 * stream.  avd_profile_read synchronises, writes JSON {"kernel": {"ms": total, "launches": n}, ...} to buf
 * and clears the log. */
int avd_profile_enable(avd_ctx *ctx, int on);
int avd_profile_read(avd_ctx *ctx, char *buf, size_t cap);

/* Select the per-frame kernel: 0 = auto (streaming fused kernel when W % 64 == 0, W >= 1280, 640 <= H <= 4096 and the
 * buffers are 16-byte aligned; else the any-width fused kernel when a band of gray rows fits in shared memory; else the
 * banded fused kernel when W % 16 == 0; generic multi-kernel path otherwise), 1 = force the generic path, 2 = the banded
 * kernel where it applies (else any-width, else generic), 3 = the any-width kernel where it applies (else generic).
 * All four produce identical bytes.  For tests. */
int avd_set_k1_mode(avd_ctx *ctx, int mode);

/*
 * Host-only view of the per-shape plan the frame kernels run from (no device needed): band ownership of source
 * rows, INTER_LINEAR taps (video.py:43), thumbnail columns per 16-pixel strip, thumbnail role of every source row
 * (bit0 tap row, bit1 bottom tap, bits 2-13 vertical weight, bits 16-24 thumbnail row) and the kernel a dense
 * aligned buffer of this shape is given (0 streaming fused, 2 banded fused, 3 any-width fused, 1 generic).  For tests and debugging.
 */
typedef struct avd_shape_plan_t {
    int area_int, thumb_2x2, strip_ntx_max, k1_path;
    int row_owner_begin[AVD_HASH_SIDE + 1], thumb_row_begin[AVD_HASH_SIDE + 1];
    int lx_i0[AVD_THUMB], lx_a0[AVD_THUMB], lx_a1[AVD_THUMB];
    int ly_i0[AVD_THUMB], ly_a0[AVD_THUMB], ly_a1[AVD_THUMB];
    int strip_tx0[256], strip_ntx[256];
    uint32_t rowinfo[4096];
} avd_shape_plan_t;
int avd_shape_plan(int H, int W, avd_shape_plan_t *out);

#ifdef __cplusplus
}
#endif
#endif /* AVD_H */
