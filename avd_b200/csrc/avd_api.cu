// C ABI of libavd.so (include/avd.h): context, resampling tables, entry points, host streaming pipeline.
#include "avd_internal.h"
#include <math.h>
#include <string.h>
#include <stdio.h>
#include <stdlib.h>
#include <new>
#include <memory>
#include <vector>

static_assert(sizeof(avd_sample_rec) == 160, "record layout is part of the ABI");

// ---- resampling tables (host) ------------------------------------------------------------------------
// INTER_AREA decimation cells (video.py:6): restated from the library's table builder, SURVEY.md A.2.
static void build_area_cells(int ssize, int dsize, AreaCell *cells)
{
    double scale = (double)ssize / dsize;
    for (int d = 0; d < dsize; d++) {
        double fs1 = d * scale, fs2 = fs1 + scale;
        double cell = scale < (ssize - fs1) ? scale : (ssize - fs1);
        int s1 = (int)ceil(fs1), s2 = (int)floor(fs2);
        if (s2 > ssize - 1) s2 = ssize - 1;
        if (s1 > s2) s1 = s2;
        AreaCell c{};
        c.s1 = s1; c.s2 = s2;
        c.has_head = (s1 - fs1 > 1e-3);
        c.ah = c.has_head ? (float)((s1 - fs1) / cell) : 0.f;
        c.am = (float)(1.0 / cell);
        c.has_tail = (fs2 - s2 > 1e-3);
        if (c.has_tail) {
            double t = fs2 - s2;
            if (t > 1.0) t = 1.0;
            if (t > cell) t = cell;
            c.at = (float)(t / cell);
        }
        cells[d] = c;
    }
}

// INTER_LINEAR 8-bit taps (video.py:43): SURVEY.md A.3.
static void build_lin_taps(int ssize, int dsize, LinTap *taps, bool is_y)
{
    double scale = 1.0 / ((double)dsize / ssize);
    for (int d = 0; d < dsize; d++) {
        float f = (float)((d + 0.5) * scale - 0.5);
        int s = (int)floorf(f);
        f -= (float)s;
        LinTap t{};
        if (!is_y) {
            if (s < 0) { f = 0.f; s = 0; }
            if (s >= ssize - 1) { f = 0.f; s = ssize - 1; }
            t.i0 = s; t.i1 = s + 1 < ssize ? s + 1 : ssize - 1;
        } else {
            t.i0 = s < 0 ? 0 : (s >= ssize ? ssize - 1 : s);
            t.i1 = s + 1 < 0 ? 0 : (s + 1 >= ssize ? ssize - 1 : s + 1);
        }
        t.a0 = (int)lrintf((1.f - f) * 2048.f);
        t.a1 = (int)lrintf(f * 2048.f);
        taps[d] = t;
    }
}

// Everything the kernels need to know about a frame shape; pure host code (also exported for the CPU tests).
static void build_shape_tables(int H, int W, ShapeTables &t)
{
    build_area_cells(W, AVD_D, t.ax);
    build_area_cells(H, AVD_D, t.ay);
    build_lin_taps(W, AVD_T, t.lx, false);
    build_lin_taps(H, AVD_T, t.ly, true);
    double sx = (double)W / AVD_D, sy = (double)H / AVD_D;
    t.isx = (int)sx; t.isy = (int)sy;
    t.area_int = fabs(sx - t.isx) < 2.220446049250313e-16 && fabs(sy - t.isy) < 2.220446049250313e-16;
    t.inv_area = 1.f / (float)(t.isx * t.isy);
    t.thumb_2x2 = (W == 2 * AVD_T && H == 2 * AVD_T);
    // band b of the fused kernels owns source rows [ceil(b*H/32), ceil((b+1)*H/32))
    for (int b = 0; b <= AVD_D; b++) t.row_owner_begin[b] = b == AVD_D ? H : (int)ceil(b * sy);
    // thumbnail row ty belongs to the band that owns its top tap row
    int ty = 0;
    for (int b = 0; b < AVD_D; b++) {
        t.thumb_row_begin[b] = ty;
        while (ty < AVD_T && (t.thumb_2x2 ? 2 * ty : t.ly[ty].i0) < t.row_owner_begin[b + 1]) ty++;
    }
    t.thumb_row_begin[AVD_D] = AVD_T;
    // thumbnail columns per 16-pixel strip (streaming kernel)
    t.strip_ntx_max = 0;
    for (int s = 0, tx = 0; s < 256; s++) {
        while (tx < AVD_T && t.lx[tx].i0 < 16 * s) tx++;
        int cnt = 0;
        while (tx + cnt < AVD_T && t.lx[tx + cnt].i0 < 16 * s + 16) cnt++;
        t.strip_tx0[s] = (uint16_t)tx; t.strip_ntx[s] = (uint8_t)(cnt > 255 ? 255 : cnt);
        if (16 * s < W && cnt > t.strip_ntx_max) t.strip_ntx_max = cnt;
    }
    if (H <= 4096) {
        bool single_role = true;
        for (int y = 0; y < AVD_T && single_role; y++) {
            const LinTap &v = t.ly[y];
            if (v.i1 != v.i0 + 1 || t.rowinfo[v.i0] || t.rowinfo[v.i1]) { single_role = false; break; }
            t.rowinfo[v.i0] = 1u | ((uint32_t)v.a0 << 2) | ((uint32_t)y << 16);
            t.rowinfo[v.i1] = 3u | ((uint32_t)v.a1 << 2) | ((uint32_t)y << 16);
        }
        if (!single_role) memset(t.rowinfo, 0, sizeof t.rowinfo);
    }
}

int avd_get_tables(avd_ctx *ctx, int H, int W, const ShapeTables **dev, const ShapeTables **host)
{
    auto key = std::make_pair(H, W);
    auto it = ctx->tables.find(key);
    if (it == ctx->tables.end()) {
        std::unique_ptr<ShapeTables> tp(new ShapeTables());   // 21 KB, value-initialised: keep it off the stack
        ShapeTables &t = *tp;
        build_shape_tables(H, W, t);
        ShapeTables *d = nullptr;
        AVD_CUDA(ctx, cudaMalloc(&d, sizeof(ShapeTables)));
        AVD_CUDA(ctx, cudaMemcpy(d, &t, sizeof(ShapeTables), cudaMemcpyHostToDevice));
        ctx->tables[key] = d;
        ctx->tables_host[key] = t;
        it = ctx->tables.find(key);
    }
    *dev = it->second;
    *host = &ctx->tables_host[key];
    return AVD_OK;
}

// Grow-only device buffers of the library.  Every buffer sits between two guard zones filled with a canary pattern;
// avd_check_guards() verifies them (compute-sanitizer is not available on every pool: an out-of-bounds WRITE by any kernel
// into the neighbourhood of a workspace shows up here).
constexpr size_t AVD_GUARD = 256;            // bytes on each side; keeps the 256-byte alignment of cudaMalloc
constexpr int AVD_CANARY = 0xA5;

static void guarded_free(avd_ctx *ctx, void *user)
{
    for (size_t i = 0; i < ctx->guarded.size(); i++)
        if (ctx->guarded[i].user == user) {
            cudaFree(ctx->guarded[i].base);
            ctx->guarded.erase(ctx->guarded.begin() + i);
            return;
        }
    cudaFree(user);
}

int avd_ensure(avd_ctx *ctx, void **p, size_t *cap, size_t need)
{
    if (*cap >= need) return AVD_OK;
    if (*p) { guarded_free(ctx, *p); *p = nullptr; *cap = 0; }
    ctx->ws_generation++;            // cached CUDA graphs hold the old address: callers key their caches on this
    size_t want = need + need / 8;
    want = (want + 255) & ~(size_t)255;
    void *base = nullptr;
    if (cudaMalloc(&base, want + 2 * AVD_GUARD) != cudaSuccess) {
        cudaGetLastError();
        want = (need + 255) & ~(size_t)255;
        if (cudaMalloc(&base, want + 2 * AVD_GUARD) != cudaSuccess) {
            cudaGetLastError();
            ctx->err = "cudaMalloc of " + std::to_string(need) + " bytes failed";
            return AVD_ERR_NOMEM;
        }
    }
    cudaMemset(base, AVD_CANARY, AVD_GUARD);
    cudaMemset((char *)base + AVD_GUARD + want, AVD_CANARY, AVD_GUARD);
    *p = (char *)base + AVD_GUARD;
    *cap = want;
    ctx->guarded.push_back({base, *p, want});
    return AVD_OK;
}

void avd_prof_begin(avd_ctx *ctx, const char *name, cudaStream_t st)
{
    avd_ctx::ProfRec r{name, nullptr, nullptr};
    for (cudaEvent_t *e : {&r.a, &r.b}) {
        if (!ctx->prof_pool.empty()) { *e = ctx->prof_pool.back(); ctx->prof_pool.pop_back(); }
        else cudaEventCreate(e);
    }
    cudaEventRecord(r.a, st);
    ctx->prof.push_back(r);
}

void avd_prof_end(avd_ctx *ctx, cudaStream_t st) { cudaEventRecord(ctx->prof.back().b, st); }

// ---- C ABI -------------------------------------------------------------------------------------------
extern "C" {
#pragma GCC visibility push(default)

const char *avd_version(void) { return "avd-b200 0.1.0 (sm_100a)"; }

int avd_device_count(void)
{
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) { cudaGetLastError(); return 0; }
    return n;
}

int avd_create(int device, avd_ctx **out)
{
    if (!out) return AVD_ERR_ARG;
    *out = nullptr;
    int n = avd_device_count();
    if (n <= 0 || device < 0 || device >= n) return AVD_ERR_NODEVICE;
    avd_ctx *ctx = new (std::nothrow) avd_ctx();
    if (!ctx) return AVD_ERR_NOMEM;
    ctx->device = device;
    cudaDeviceProp prop;
    if (cudaSetDevice(device) != cudaSuccess || cudaGetDeviceProperties(&prop, device) != cudaSuccess) {
        cudaGetLastError(); delete ctx; return AVD_ERR_NODEVICE;
    }
    if (prop.major < 10) { delete ctx; return AVD_ERR_NODEVICE; }   // sm_100a SASS only: fail loudly elsewhere
    ctx->sm_count = prop.multiProcessorCount;
    if (cudaStreamCreateWithFlags(&ctx->compute, cudaStreamNonBlocking) != cudaSuccess ||
        cudaStreamCreateWithFlags(&ctx->copy, cudaStreamNonBlocking) != cudaSuccess) {
        cudaGetLastError(); delete ctx; return AVD_ERR_CUDA;
    }
    for (int i = 0; i < 4; i++) {
        cudaEventCreateWithFlags(&ctx->ev_copied[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&ctx->ev_free[i], cudaEventDisableTiming);
    }
    *out = ctx;
    return AVD_OK;
}

void avd_destroy(avd_ctx *ctx)
{
    if (!ctx) return;
    cudaSetDevice(ctx->device);
    cudaDeviceSynchronize();
    for (auto &kv : ctx->tables) cudaFree(kv.second);
    for (auto &g : ctx->guarded) cudaFree(g.base);          // every grow-only workspace (avd_ensure)
    ctx->guarded.clear();
    for (void *h : ctx->host_allocs) cudaFreeHost(h);
    ctx->host_allocs.clear();
    for (int i = 0; i < 4; i++) { cudaEventDestroy(ctx->ev_copied[i]); cudaEventDestroy(ctx->ev_free[i]); }
    if (ctx->k2_img) {
        cudaStreamDestroy(ctx->k2_img);
        for (int i = 0; i < 5; i++) cudaEventDestroy(ctx->k2_img_ev[i]);
    }
    if (ctx->k2_side[0]) {
        for (int i = 0; i < 3; i++) cudaStreamDestroy(ctx->k2_side[i]);
        for (int i = 0; i < 6; i++) cudaEventDestroy(ctx->k2_ev[i]);
    }
    cudaStreamDestroy(ctx->compute); cudaStreamDestroy(ctx->copy);
    cudaGetLastError();
    delete ctx;
}

const char *avd_last_error(const avd_ctx *ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int avd_sync(avd_ctx *ctx)
{
    if (!ctx) return AVD_ERR_ARG;
    AVD_CUDA(ctx, cudaSetDevice(ctx->device));
    AVD_CUDA(ctx, cudaDeviceSynchronize());
    return AVD_OK;
}

void *avd_host_alloc(avd_ctx *ctx, size_t bytes)
{
    void *p = nullptr;
    if (!ctx || cudaSetDevice(ctx->device) != cudaSuccess) return nullptr;
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) { cudaGetLastError(); ctx->err = "cudaHostAlloc failed"; return nullptr; }
    ctx->host_allocs.insert(p);
    return p;
}

void avd_host_free(avd_ctx *ctx, void *p)
{
    if (!p) return;
    if (ctx) ctx->host_allocs.erase(p);
    cudaFreeHost(p);
}

uint64_t avd_launch_count(const avd_ctx *ctx) { return ctx ? ctx->launches : 0; }
uint64_t avd_ws_generation(const avd_ctx *ctx) { return ctx ? ctx->ws_generation : 0; }

// Number of corrupted guard bytes around the library's device buffers (0 = intact), or a negative status.
long long avd_check_guards(avd_ctx *ctx)
{
    if (!ctx) return AVD_ERR_ARG;
    if (cudaSetDevice(ctx->device) != cudaSuccess || cudaDeviceSynchronize() != cudaSuccess) {
        ctx->err = std::string("avd_check_guards: ") + cudaGetErrorString(cudaGetLastError());
        return AVD_ERR_CUDA;
    }
    long long bad = 0;
    unsigned char h[2 * AVD_GUARD];
    for (size_t i = 0; i < ctx->guarded.size(); i++) {
        const auto &g = ctx->guarded[i];
        if (cudaMemcpy(h, g.base, AVD_GUARD, cudaMemcpyDeviceToHost) != cudaSuccess ||
            cudaMemcpy(h + AVD_GUARD, (char *)g.base + AVD_GUARD + g.cap, AVD_GUARD, cudaMemcpyDeviceToHost) != cudaSuccess) {
            cudaGetLastError();
            return AVD_ERR_CUDA;
        }
        long long b = 0;
        for (size_t k = 0; k < 2 * AVD_GUARD; k++) b += h[k] != AVD_CANARY;
        if (b) ctx->err = "guard zone of workspace #" + std::to_string(i) + " (" + std::to_string(g.cap) + " bytes) overwritten: " + std::to_string(b) + " bytes";
        bad += b;
    }
    return bad;
}
size_t avd_host_alloc_count(const avd_ctx *ctx) { return ctx ? ctx->host_allocs.size() : 0; }

int avd_profile_enable(avd_ctx *ctx, int on)
{
    if (!ctx) return AVD_ERR_ARG;
    ctx->profiling = on != 0;
    return AVD_OK;
}

// JSON: {"<kernel>": {"ms": <sum of launch durations>, "launches": <n>}, ...}; clears the log.
int avd_profile_read(avd_ctx *ctx, char *buf, size_t cap)
{
    if (!ctx || !buf || cap < 3) return AVD_ERR_ARG;
    AVD_CUDA(ctx, cudaSetDevice(ctx->device));
    AVD_CUDA(ctx, cudaDeviceSynchronize());
    std::map<std::string, std::pair<double, long>> acc;
    for (auto &r : ctx->prof) {
        float ms = 0.f;
        if (cudaEventElapsedTime(&ms, r.a, r.b) == cudaSuccess) { acc[r.name].first += ms; acc[r.name].second++; }
        else cudaGetLastError();
        ctx->prof_pool.push_back(r.a); ctx->prof_pool.push_back(r.b);
    }
    ctx->prof.clear();
    std::string js = "{";
    for (auto &kv : acc) {
        char tmp[160];
        snprintf(tmp, sizeof tmp, "%s\"%s\": {\"ms\": %.6f, \"launches\": %ld}", js.size() > 1 ? ", " : "", kv.first.c_str(), kv.second.first, kv.second.second);
        js += tmp;
    }
    js += "}";
    if (js.size() + 1 > cap) { ctx->err = "profile buffer too small"; return AVD_ERR_ARG; }
    memcpy(buf, js.c_str(), js.size() + 1);
    return AVD_OK;
}

// Host-only (no device, no context): the per-shape plan of the frame kernels, for tests and integration debugging.
int avd_shape_plan(int H, int W, avd_shape_plan_t *out)
{
    if (!out || H < AVD_D || W < AVD_D) return AVD_ERR_ARG;
    std::unique_ptr<ShapeTables> tp(new ShapeTables());
    ShapeTables &t = *tp;
    build_shape_tables(H, W, t);
    memset(out, 0, sizeof *out);
    out->area_int = t.area_int; out->thumb_2x2 = t.thumb_2x2; out->strip_ntx_max = t.strip_ntx_max;
    for (int b = 0; b <= AVD_D; b++) { out->row_owner_begin[b] = t.row_owner_begin[b]; out->thumb_row_begin[b] = t.thumb_row_begin[b]; }
    for (int i = 0; i < AVD_T; i++) {
        out->lx_i0[i] = t.lx[i].i0; out->lx_a0[i] = t.lx[i].a0; out->lx_a1[i] = t.lx[i].a1;
        out->ly_i0[i] = t.ly[i].i0; out->ly_a0[i] = t.ly[i].a0; out->ly_a1[i] = t.ly[i].a1;
    }
    for (int s = 0; s < 256; s++) { out->strip_tx0[s] = t.strip_tx0[s]; out->strip_ntx[s] = t.strip_ntx[s]; }
    for (int r = 0; r < H && r < 4096; r++) out->rowinfo[r] = t.rowinfo[r];
    // which per-frame kernel a dense, 16-byte aligned buffer of this shape gets (0 streaming, 2 banded, 3 any-width fused, 1 generic)
    const size_t pitch = (size_t)W * 3;
    out->k1_path = k1_fused_supported(H, W, pitch, pitch * H, nullptr, &t) ? 0 :
                   (k1_rows_supported(H, W, &t) ? 3 : (k1_banded_supported(H, W, pitch, pitch * H, nullptr) ? 2 : 1));
    return AVD_OK;
}

int avd_set_k1_mode(avd_ctx *ctx, int mode)
{
    if (!ctx || mode < 0 || mode > 3) return AVD_ERR_ARG;
    ctx->k1_mode = mode;
    return AVD_OK;
}

static int check_shape(avd_ctx *ctx, int n, int H, int W, size_t pitch)
{
    if (n < 0 || H < AVD_D || W < AVD_D || pitch < (size_t)W * 3) {
        ctx->err = "bad shape: need n >= 0, H >= 32, W >= 32, row_pitch >= 3*W";
        return AVD_ERR_ARG;
    }
    return AVD_OK;
}

int avd_frames_features(avd_ctx *ctx, const uint8_t *d_bgr, int n, int H, int W, size_t row_pitch,
                        size_t frame_stride, const uint32_t *d_prev_hash, avd_sample_rec *d_rec,
                        uint8_t *d_thumbs, uint8_t *d_area32, void *stream)
{
    if (!ctx) return AVD_ERR_ARG;
    int rc = check_shape(ctx, n, H, W, row_pitch);
    if (rc) return rc;
    if (n == 0) return AVD_OK;
    if (!d_bgr || !d_rec || !d_thumbs || frame_stride < row_pitch * (size_t)H) { ctx->err = "null buffer or frame_stride < H*row_pitch"; return AVD_ERR_ARG; }
    AVD_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;   // NULL is CUDA's default stream, as for any cudaStream_t
    const ShapeTables *tab, *htab;
    if ((rc = avd_get_tables(ctx, H, W, &tab, &htab))) return rc;
    if (!d_area32) {
        if ((rc = avd_ensure(ctx, (void **)&ctx->area32, &ctx->area32_cap, (size_t)n * 1024))) return rc;
        d_area32 = ctx->area32;
    }
    K1FixArgs fix{d_bgr, row_pitch, frame_stride, tab, nullptr};
    // the streaming kernel writes the thumbnail rows with 128-bit stores
    // 0 streaming (TMA) kernel, 2 banded kernel, 3 any-width kernel, 1 generic multi-kernel path
    int path = 1;
    if (ctx->k1_mode == 0 && ((uintptr_t)d_thumbs % 16) == 0 && k1_fused_supported(H, W, row_pitch, frame_stride, d_bgr, htab)) path = 0;
    else if (ctx->k1_mode == 2 && k1_banded_supported(H, W, row_pitch, frame_stride, d_bgr)) path = 2;
    else if (ctx->k1_mode != 1 && k1_rows_supported(H, W, htab) && (unsigned long long)row_pitch * (unsigned long long)H < (1ull << 32)) path = 3;
    else if (ctx->k1_mode == 0 && k1_banded_supported(H, W, row_pitch, frame_stride, d_bgr)) path = 2;
    // the streaming and the any-width kernel list the INTER_AREA cells they cannot certify; area_fix_kernel recomputes them
    const bool lists = path == 0 || path == 3;
    if (lists) {
        if ((rc = avd_ensure(ctx, (void **)&ctx->fixlist, &ctx->fixlist_cap, ((size_t)n * 1024 + 1) * sizeof(uint32_t)))) return rc;
        fix.list = ctx->fixlist;
    }
    if ((rc = k1_rec_init(ctx, d_rec, n, lists ? ctx->fixlist : nullptr, st))) return rc;
    if (path == 0)
        rc = k1_fused(ctx, d_bgr, n, H, W, row_pitch, frame_stride, tab, htab, d_rec, d_thumbs, d_area32, ctx->fixlist, st);
    else if (path == 2)
        rc = k1_banded(ctx, d_bgr, n, H, W, row_pitch, frame_stride, tab, htab, d_rec, d_thumbs, d_area32, st);
    else if (path == 3)
        rc = k1_rows(ctx, d_bgr, n, H, W, row_pitch, frame_stride, tab, htab, d_rec, d_thumbs, d_area32, ctx->fixlist, st);
    else
        rc = k1_generic(ctx, d_bgr, n, H, W, row_pitch, frame_stride, tab, htab, d_rec, d_thumbs, d_area32, st);
    if (rc) return rc;
    return k1_hash_finalize(ctx, d_area32, n, d_prev_hash, d_rec, fix.list ? &fix : nullptr, st);
}

int avd_flow_stats(avd_ctx *ctx, const uint8_t *d_thumbs, int n, const uint8_t *d_prev_thumb,
                   avd_sample_rec *d_rec, void *stream)
{
    if (!ctx || n < 0) return AVD_ERR_ARG;
    if (n == 0) return AVD_OK;
    if (!d_thumbs || !d_rec) { ctx->err = "null buffer"; return AVD_ERR_ARG; }
    AVD_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;   // NULL is CUDA's default stream, as for any cudaStream_t
    int n_pairs = d_prev_thumb ? n : n - 1;
    return k2_flow(ctx, d_prev_thumb, d_thumbs, 0, (size_t)AVD_T * AVD_T, n_pairs, true,
                   d_prev_thumb ? d_rec : d_rec + 1, nullptr, st);
}

int avd_flow_fields(avd_ctx *ctx, const uint8_t *d_prev, const uint8_t *d_next, int n_pairs, float *d_flow, void *stream)
{
    if (!ctx || n_pairs < 0) return AVD_ERR_ARG;
    if (n_pairs == 0) return AVD_OK;
    if (!d_prev || !d_next || !d_flow) { ctx->err = "null buffer"; return AVD_ERR_ARG; }
    AVD_CUDA(ctx, cudaSetDevice(ctx->device));
    cudaStream_t st = (cudaStream_t)stream;   // NULL is CUDA's default stream, as for any cudaStream_t
    return k2_flow(ctx, d_prev, d_next, (size_t)AVD_T * AVD_T, (size_t)AVD_T * AVD_T, n_pairs, false, nullptr, d_flow, st);
}

int avd_gray(avd_ctx *ctx, const uint8_t *d_bgr, int H, int W, size_t row_pitch, uint8_t *d_gray, void *stream)
{
    if (!ctx) return AVD_ERR_ARG;
    int rc = check_shape(ctx, 1, H, W, row_pitch);
    if (rc) return rc;
    if (!d_bgr || !d_gray) { ctx->err = "null buffer"; return AVD_ERR_ARG; }
    AVD_CUDA(ctx, cudaSetDevice(ctx->device));
    return k1_gray_only(ctx, d_bgr, H, W, row_pitch, d_gray, (cudaStream_t)stream);
}

// Host streaming pipeline.  Frames go to a ring of device slots on the copy stream; the compute stream converts
// (YUV input) and runs K1 on a slot as soon as its copy lands; K2 runs on the thumbnails accumulated so far whenever at
// least k2_min pairs are pending (and once more at the end), so PCIe transfer, the per-frame kernel and the flow
// kernels overlap while K2 still gets batches large enough to fill the GPU.
static int analyze_host_impl(avd_ctx *ctx, const uint8_t *const *h_frames, int n, int H, int W, size_t row_pitch, int fmt,
                             avd_sample_rec *h_rec)
{
    int rc;
    const bool yuv = fmt != AVD_FMT_BGR24;
    // device frame layout inside a BGR slot: dense rows padded to 16 bytes so the streaming kernel applies
    const size_t dpitch = ((size_t)W * 3 + 15) & ~(size_t)15;
    const size_t fbytes = dpitch * H;
    const size_t ybytes = (size_t)W * H * 3 / 2;                    // dense I420 frame
    const size_t ubytes = yuv ? ybytes : fbytes;                    // bytes uploaded per frame
    const int SLOTS = 3;
    int per_slot = (int)((112u << 20) / fbytes);
    if (per_slot < 1) per_slot = 1;
    static const int max_slot = getenv("AVD_HOST_CHUNK") ? atoi(getenv("AVD_HOST_CHUNK")) : 16;   // frames per device slot (tuning)
    static const int k2_min = getenv("AVD_HOST_K2_MIN") ? atoi(getenv("AVD_HOST_K2_MIN")) : 40;   // pairs per K2 batch (tuning)
    if (per_slot > max_slot) per_slot = max_slot < 1 ? 1 : max_slot;
    if (per_slot > n) per_slot = n;
    if ((rc = avd_ensure(ctx, (void **)&ctx->ring, &ctx->ring_cap, (size_t)SLOTS * per_slot * ubytes))) return rc;
    if (yuv && (rc = avd_ensure(ctx, (void **)&ctx->bgr_slot, &ctx->bgr_slot_cap, (size_t)per_slot * fbytes))) return rc;
    if ((rc = avd_ensure(ctx, (void **)&ctx->thumbs, &ctx->thumbs_cap, (size_t)n * AVD_T * AVD_T))) return rc;
    if ((rc = avd_ensure(ctx, (void **)&ctx->recs, &ctx->recs_cap, (size_t)n * sizeof(avd_sample_rec)))) return rc;
    if ((rc = avd_ensure(ctx, (void **)&ctx->area32, &ctx->area32_cap, (size_t)n * 1024))) return rc;
    if ((rc = k2_reserve(ctx, n - 1))) return rc;                   // no workspace growth while the pipeline is in flight
    const ShapeTables *tab, *htab;
    if ((rc = avd_get_tables(ctx, H, W, &tab, &htab))) return rc;
    int chunk = 0, k2_done = 0;                                     // k2_done: thumbnails whose flow statistics are issued
    for (int f0 = 0; f0 < n; f0 += per_slot, chunk++) {
        const int nb = n - f0 < per_slot ? n - f0 : per_slot;
        const int slot = chunk % SLOTS;
        uint8_t *up = ctx->ring + (size_t)slot * per_slot * ubytes;
        if (chunk >= SLOTS) AVD_CUDA(ctx, cudaStreamWaitEvent(ctx->copy, ctx->ev_free[slot], 0));
        for (int j = 0; j < nb;) {
            if (yuv || dpitch == row_pitch) {
                // frames that sit back to back in host memory (one decoded batch) go as one copy: fewer, larger DMA requests
                int run = 1;
                while (j + run < nb && h_frames[f0 + j + run] == h_frames[f0 + j + run - 1] + ubytes) run++;
                AVD_CUDA(ctx, cudaMemcpyAsync(up + (size_t)j * ubytes, h_frames[f0 + j], ubytes * run, cudaMemcpyHostToDevice, ctx->copy));
                j += run;
            } else {
                AVD_CUDA(ctx, cudaMemcpy2DAsync(up + (size_t)j * fbytes, dpitch, h_frames[f0 + j], row_pitch, (size_t)W * 3, H,
                                               cudaMemcpyHostToDevice, ctx->copy));
                j++;
            }
        }
        AVD_CUDA(ctx, cudaEventRecord(ctx->ev_copied[slot], ctx->copy));
        AVD_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, ctx->ev_copied[slot], 0));
        const uint8_t *bgr = up;
        if (yuv) {
            if ((rc = k0_yuv420_to_bgr(ctx, up, ybytes, nb, H, W, fmt == AVD_FMT_J420, ctx->bgr_slot, dpitch, fbytes, ctx->compute))) return rc;
            AVD_CUDA(ctx, cudaEventRecord(ctx->ev_free[slot], ctx->compute));    // the upload slot is free once converted
            bgr = ctx->bgr_slot;
        }
        avd_sample_rec *rec = ctx->recs + f0;
        uint8_t *th = ctx->thumbs + (size_t)f0 * AVD_T * AVD_T;
        const uint32_t *prev_hash = f0 > 0 ? ctx->recs[f0 - 1].hash : nullptr;
        if ((rc = avd_frames_features(ctx, bgr, nb, H, W, dpitch, fbytes, prev_hash, rec, th, ctx->area32 + (size_t)f0 * 1024, ctx->compute))) return rc;
        if (!yuv) AVD_CUDA(ctx, cudaEventRecord(ctx->ev_free[slot], ctx->compute));
        const int ready = f0 + nb, last = ready == n;
        const int pending = ready - k2_done - (k2_done == 0 ? 1 : 0);            // pairs not yet issued
        // the batch issued after the LAST copy is the pipeline's un-hideable tail: flush before it, so that it only holds the
        // last slot's pairs (120 samples of 1080p: 24 pairs = 0.45 ms -> 8 pairs = 0.3 ms after the last copy)
        const bool before_last = !last && n - ready <= per_slot && pending >= 8;
        if (pending > 0 && (last || pending >= k2_min || before_last)) {
            const uint8_t *t0 = ctx->thumbs + (size_t)k2_done * AVD_T * AVD_T;
            if ((rc = avd_flow_stats(ctx, t0, ready - k2_done, k2_done > 0 ? t0 - (size_t)AVD_T * AVD_T : nullptr, ctx->recs + k2_done, ctx->compute))) return rc;
            k2_done = ready;
        }
    }
    if (h_rec) AVD_CUDA(ctx, cudaMemcpyAsync(h_rec, ctx->recs, (size_t)n * sizeof(avd_sample_rec), cudaMemcpyDeviceToHost, ctx->compute));
    AVD_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
    ctx->recs_n = n;
    return AVD_OK;
}

static int analyze_host_checked(avd_ctx *ctx, const uint8_t *const *h_frames, int n, int H, int W, size_t row_pitch, int fmt,
                                avd_sample_rec *h_rec)
{
    if (!ctx) return AVD_ERR_ARG;
    int rc = check_shape(ctx, n, H, W, row_pitch);
    if (rc) return rc;
    if (fmt != AVD_FMT_BGR24 && fmt != AVD_FMT_I420 && fmt != AVD_FMT_J420) { ctx->err = "unknown frame format"; return AVD_ERR_ARG; }
    if (fmt != AVD_FMT_BGR24 && ((H | W) & 1)) { ctx->err = "4:2:0 input needs even width and height"; return AVD_ERR_ARG; }
    if (n == 0) return AVD_OK;
    if (!h_frames) { ctx->err = "null buffer"; return AVD_ERR_ARG; }
    for (int i = 0; i < n; i++)
        if (!h_frames[i]) { ctx->err = "null frame pointer"; return AVD_ERR_ARG; }
    ctx->recs_n = 0;
    AVD_CUDA(ctx, cudaSetDevice(ctx->device));
    rc = analyze_host_impl(ctx, h_frames, n, H, W, row_pitch, fmt, h_rec);
    if (rc) {
        // single error exit: copies from the caller's staging memory may still be in flight -- the caller is free to reuse
        // that memory as soon as we return, so drain both streams first (and clear a sticky launch error)
        cudaStreamSynchronize(ctx->copy);
        cudaStreamSynchronize(ctx->compute);
        cudaGetLastError();
    }
    return rc;
}

int avd_analyze_host(avd_ctx *ctx, const uint8_t *const *h_frames, int n, int H, int W, size_t row_pitch,
                     avd_sample_rec *h_rec)
{
    return analyze_host_checked(ctx, h_frames, n, H, W, row_pitch, AVD_FMT_BGR24, h_rec);
}

int avd_analyze_host_yuv420(avd_ctx *ctx, const uint8_t *const *h_frames, int n, int H, int W, int full_range,
                            avd_sample_rec *h_rec)
{
    return analyze_host_checked(ctx, h_frames, n, H, W, (size_t)W * 3, full_range ? AVD_FMT_J420 : AVD_FMT_I420, h_rec);
}

int avd_yuv420_to_bgr(avd_ctx *ctx, const uint8_t *d_yuv, size_t yuv_frame_stride, int n, int H, int W, int full_range,
                      uint8_t *d_bgr, size_t row_pitch, size_t frame_stride, void *stream)
{
    if (!ctx) return AVD_ERR_ARG;
    int rc = check_shape(ctx, n, H, W, row_pitch);
    if (rc) return rc;
    if ((H | W) & 1) { ctx->err = "4:2:0 input needs even width and height"; return AVD_ERR_ARG; }
    if (n == 0) return AVD_OK;
    if (!d_yuv || !d_bgr || yuv_frame_stride < (size_t)W * H * 3 / 2 || frame_stride < row_pitch * (size_t)H) { ctx->err = "null buffer or stride too small"; return AVD_ERR_ARG; }
    AVD_CUDA(ctx, cudaSetDevice(ctx->device));
    return k0_yuv420_to_bgr(ctx, d_yuv, yuv_frame_stride, n, H, W, full_range != 0, d_bgr, row_pitch, frame_stride, (cudaStream_t)stream);
}

#pragma GCC visibility pop
}  // extern "C"
