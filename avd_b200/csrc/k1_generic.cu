// K1, generic path: any H, W >= 32, any pitch/alignment.  Several simple kernels over a gray scratch
// plane.  Used when the fused TMA kernel (k1_fused.cu) cannot take the shape (W % 16 != 0 or unaligned
// buffers), and as the on-device cross-check of the fused kernel in tests.
//
// Arithmetic restated from the library calls the reference makes:
//   gray   video.py:5,43,51   Y = (3735 B + 19235 G + 9798 R + 16384) >> 15
//   area   video.py:6         INTER_AREA -> 32x32, float32 sequential accumulation, no FMA
//   thumb  video.py:43        INTER_LINEAR -> 320x320, 11-bit fixed point
//   lap    video.py:52        4-neighbour Laplacian, REFLECT_101, exact int64 sums
#include "avd_internal.h"

namespace {

__global__ void rec_init_kernel(avd_sample_rec *rec, int n, uint32_t *fixcount)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0 && fixcount) *fixcount = 0;
    if (i >= n) return;
    rec[i].lap_sum = 0; rec[i].lap_sumsq = 0; rec[i].flow_mean = 0.f; rec[i].flow_var = 0.f;
    rec[i].ham_prev = -1; rec[i].flags = 0;
}

__global__ void gray_kernel(const uint8_t *__restrict__ bgr, int H, int W, size_t pitch, size_t fstride,
                            uint8_t *__restrict__ gray)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
    if (x >= W) return;
    const uint8_t *p = bgr + (size_t)f * fstride + (size_t)y * pitch + 3 * (size_t)x;
    gray[((size_t)f * H + y) * W + x] = (uint8_t)((3735u * p[0] + 19235u * p[1] + 9798u * p[2] + 16384u) >> 15);
}

__device__ __forceinline__ int r101(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }

__global__ void lap_kernel(const uint8_t *__restrict__ gray, int H, int W, avd_sample_rec *rec)
{
    int f = blockIdx.z, y = blockIdx.y;
    const uint8_t *g = gray + (size_t)f * H * W;
    const uint8_t *c = g + (size_t)y * W, *u = g + (size_t)r101(y - 1, H) * W, *d = g + (size_t)r101(y + 1, H) * W;
    long long s = 0, ss = 0;
    for (int x = blockIdx.x * blockDim.x + threadIdx.x; x < W; x += gridDim.x * blockDim.x) {
        int L = (int)u[x] + d[x] + c[r101(x - 1, W)] + c[r101(x + 1, W)] - 4 * (int)c[x];
        s += L; ss += L * L;
    }
    for (int o = 16; o; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); ss += __shfl_down_sync(0xffffffffu, ss, o); }
    __shared__ long long sh[2][32];
    int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    if (l == 0) { sh[0][w] = s; sh[1][w] = ss; }
    __syncthreads();
    if (w == 0) {
        int nw = blockDim.x >> 5;
        s = l < nw ? sh[0][l] : 0; ss = l < nw ? sh[1][l] : 0;
        for (int o = 16; o; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); ss += __shfl_down_sync(0xffffffffu, ss, o); }
        if (l == 0) {
            atomicAdd((unsigned long long *)&rec[f].lap_sum, (unsigned long long)s);
            atomicAdd((unsigned long long *)&rec[f].lap_sumsq, (unsigned long long)ss);
        }
    }
}

// horizontal INTER_AREA pass: one thread per (frame,row,dx); strictly sequential float32 adds
__global__ void area_h_kernel(const uint8_t *__restrict__ gray, int H, int W, const ShapeTables *__restrict__ tab,
                              float *__restrict__ rows, int total)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    int dx = i & 31, fy = i >> 5;
    const uint8_t *S = gray + (size_t)fy * W;
    AreaCell c = tab->ax[dx];
    float buf = 0.f;
    if (c.has_head) buf = __fadd_rn(buf, __fmul_rn((float)S[c.s1 - 1], c.ah));
    for (int s = c.s1; s < c.s2; s++) buf = __fadd_rn(buf, __fmul_rn((float)S[s], c.am));
    if (c.has_tail) buf = __fadd_rn(buf, __fmul_rn((float)S[c.s2], c.at));
    rows[i] = buf;
}

__device__ __forceinline__ uint8_t sat_rne(float v)
{
    int r = __float2int_rn(v);
    return (uint8_t)min(max(r, 0), 255);
}

__global__ void area_v_kernel(const float *__restrict__ rows, int H, const ShapeTables *__restrict__ tab,
                              uint8_t *__restrict__ area32, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 1024) return;
    int dx = i & 31, dy = (i >> 5) & 31, f = i >> 10;
    const float *R = rows + (size_t)f * H * 32 + dx;
    AreaCell c = tab->ay[dy];
    float sum = 0.f;
    if (c.has_head) sum = __fadd_rn(sum, __fmul_rn(c.ah, R[(size_t)(c.s1 - 1) * 32]));
    for (int s = c.s1; s < c.s2; s++) sum = __fadd_rn(sum, __fmul_rn(c.am, R[(size_t)s * 32]));
    if (c.has_tail) sum = __fadd_rn(sum, __fmul_rn(c.at, R[(size_t)c.s2 * 32]));
    area32[i] = sat_rne(sum);
}

// integer-scale INTER_AREA: exact integer block sum, one float multiply, round half even
__global__ void area_int_kernel(const uint8_t *__restrict__ gray, int H, int W, const ShapeTables *__restrict__ tab,
                                uint8_t *__restrict__ area32, int n)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * 1024) return;
    int dx = i & 31, dy = (i >> 5) & 31, f = i >> 10;
    int isx = tab->isx, isy = tab->isy;
    const uint8_t *g = gray + (size_t)f * H * W + (size_t)dy * isy * W + dx * isx;
    int sum = 0;
    for (int y = 0; y < isy; y++)
        for (int x = 0; x < isx; x++) sum += g[(size_t)y * W + x];
    area32[i] = sat_rne(__fmul_rn((float)sum, tab->inv_area));
}

__global__ void thumb_kernel(const uint8_t *__restrict__ gray, int H, int W, const ShapeTables *__restrict__ tab,
                             uint8_t *__restrict__ thumbs)
{
    int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, f = blockIdx.z;
    if (x >= AVD_T) return;
    const uint8_t *g = gray + (size_t)f * H * W;
    uint8_t out;
    if (tab->thumb_2x2) {
        const uint8_t *p = g + (size_t)(2 * y) * W + 2 * x;
        out = (uint8_t)((p[0] + p[1] + p[W] + p[W + 1] + 2) >> 2);
    } else {
        LinTap tx = tab->lx[x], ty = tab->ly[y];
        const uint8_t *r0 = g + (size_t)ty.i0 * W, *r1 = g + (size_t)ty.i1 * W;
        int S0 = r0[tx.i0] * tx.a0 + r0[tx.i1] * tx.a1;
        int S1 = r1[tx.i0] * tx.a0 + r1[tx.i1] * tx.a1;
        out = (uint8_t)((((ty.a0 * (S0 >> 4)) >> 16) + ((ty.a1 * (S1 >> 4)) >> 16) + 2) >> 2);
    }
    thumbs[((size_t)f * AVD_T + y) * AVD_T + x] = out;
}

// INTER_AREA cells the streaming kernel could not certify (exact mean too close to a rounding boundary, ~0.2 % of
// cells): recomputed here in the library's exact float32 order.  The CTAs share the list; all threads
// of a CTA produce the weighted pixels fl(g * alpha) of a cell in parallel, then only the float32 additions -- whose
// order is what matters -- run as sequential chains (one thread per source row, then one thread over the rows).
constexpr int FIX_FLOATS = 8192, FIX_THREADS = 256;
__global__ void __launch_bounds__(FIX_THREADS) area_fix_kernel(K1FixArgs A, uint8_t *__restrict__ area32)
{
    __shared__ float prod[FIX_FLOATS];
    __shared__ float tv[FIX_THREADS];
    const int tid = threadIdx.x;
    const uint32_t count = A.list[0];
    for (uint32_t e = blockIdx.x; e < count; e += gridDim.x) {
        const uint32_t ent = A.list[1 + e];
        const int f = (int)(ent >> 10), dy = (int)((ent >> 5) & 31u), dx = (int)(ent & 31u);
        const uint8_t *frame = A.bgr + (size_t)f * A.fstride;
        const AreaCell cy = A.tab->ay[dy], cx = A.tab->ax[dx];
        const int r0 = cy.has_head ? cy.s1 - 1 : cy.s1, r1 = cy.has_tail ? cy.s2 : cy.s2 - 1, nr = r1 - r0 + 1;
        const int x0 = cx.has_head ? cx.s1 - 1 : cx.s1, x1 = cx.has_tail ? cx.s2 : cx.s2 - 1, nc = x1 - x0 + 1;
        int ch = FIX_FLOATS / nc;
        if (ch > FIX_THREADS) ch = FIX_THREADS;
        float sum = 0.f;
        for (int rb = 0; rb < nr; rb += ch) {
            const int cnt = min(ch, nr - rb);
            for (int idx = tid; idx < cnt * nc; idx += FIX_THREADS) {
                const int r = idx / nc, c = idx - r * nc;
                const uint8_t *px = frame + (size_t)(r0 + rb + r) * A.pitch + 3 * (size_t)(x0 + c);
                const float g = (float)((3735u * px[0] + 19235u * px[1] + 9798u * px[2] + 16384u) >> 15);
                const float w = (cx.has_head && c == 0) ? cx.ah : ((cx.has_tail && c == nc - 1) ? cx.at : cx.am);
                prod[idx] = __fmul_rn(g, w);
            }
            __syncthreads();
            if (tid < cnt) {
                const float *pr = prod + tid * nc;
                float buf = 0.f;
#pragma unroll 4
                for (int i = 0; i < nc; i++) buf = __fadd_rn(buf, pr[i]);
                const int row = r0 + rb + tid;
                const float beta = (cy.has_head && row == r0) ? cy.ah : ((cy.has_tail && row == cy.s2) ? cy.at : cy.am);
                tv[tid] = __fmul_rn(beta, buf);
            }
            __syncthreads();
            if (tid == 0)
                for (int i = 0; i < cnt; i++) sum = __fadd_rn(sum, tv[i]);
            __syncthreads();
        }
        if (tid == 0) area32[(size_t)f * 1024 + dy * 32 + dx] = sat_rne(sum);
    }
}

// aHash of frame f (and of its predecessor, recomputed, for the Hamming distance): 1024 threads per frame
__device__ __forceinline__ uint32_t hash_word(const uint8_t *a, int tid, unsigned *sh_sum)
{
    unsigned v = a[tid];
    unsigned s = v;
    for (int o = 16; o; o >>= 1) s += __shfl_down_sync(0xffffffffu, s, o);
    __syncthreads();
    if (tid == 0) *sh_sum = 0;
    __syncthreads();
    if ((tid & 31) == 0) atomicAdd(sh_sum, s);
    __syncthreads();
    unsigned total = *sh_sum;
    return __ballot_sync(0xffffffffu, 1024u * v >= total);   // g >= mean  <=>  1024 g >= sum
}

__global__ void hash_kernel(const uint8_t *__restrict__ area32, const uint32_t *__restrict__ prev_hash,
                            avd_sample_rec *rec)
{
    __shared__ unsigned sh_sum;
    __shared__ int sh_ham;
    int f = blockIdx.x, tid = threadIdx.x;
    uint32_t w = hash_word(area32 + (size_t)f * 1024, tid, &sh_sum);
    if (tid == 0) sh_ham = 0;
    bool has_prev = f > 0 || prev_hash != nullptr;
    uint32_t pw = 0;
    if (f > 0) pw = hash_word(area32 + (size_t)(f - 1) * 1024, tid, &sh_sum);
    else if (prev_hash) pw = prev_hash[tid >> 5];
    __syncthreads();
    if ((tid & 31) == 0) {
        rec[f].hash[tid >> 5] = w;
        if (has_prev) atomicAdd(&sh_ham, __popc(w ^ pw));
    }
    __syncthreads();
    if (tid == 0) {
        rec[f].ham_prev = has_prev ? sh_ham : -1;
        rec[f].flags = (rec[f].flags & ~AVD_REC_HAS_PREV) | (has_prev ? AVD_REC_HAS_PREV : 0u);
    }
}

}  // namespace

int k1_rec_init(avd_ctx *ctx, avd_sample_rec *d_rec, int n, uint32_t *d_fixcount, cudaStream_t st)
{
    AVD_LAUNCH(ctx, "rec_init", st, rec_init_kernel<<<(n + 255) / 256, 256, 0, st>>>(d_rec, n, d_fixcount));
    return AVD_OK;
}

int k1_hash_finalize(avd_ctx *ctx, uint8_t *d_area32, int n, const uint32_t *d_prev_hash,
                     avd_sample_rec *d_rec, const K1FixArgs *fix, cudaStream_t st)
{
    if (fix) AVD_LAUNCH(ctx, "area_fix", st, area_fix_kernel<<<2 * ctx->sm_count, FIX_THREADS, 0, st>>>(*fix, d_area32));
    AVD_LAUNCH(ctx, "hash", st, hash_kernel<<<n, 1024, 0, st>>>(d_area32, d_prev_hash, d_rec));
    return AVD_OK;
}

int k1_gray_only(avd_ctx *ctx, const uint8_t *d_bgr, int H, int W, size_t pitch, uint8_t *d_gray, cudaStream_t st)
{
    AVD_LAUNCH(ctx, "gray", st, gray_kernel<<<dim3((W + 255) / 256, H, 1), 256, 0, st>>>(d_bgr, H, W, pitch, 0, d_gray));
    return AVD_OK;
}

int k1_generic(avd_ctx *ctx, const uint8_t *d_bgr, int n, int H, int W, size_t pitch, size_t fstride,
               const ShapeTables *tab, const ShapeTables *htab, avd_sample_rec *d_rec, uint8_t *d_thumbs,
               uint8_t *d_area32, cudaStream_t st)
{
    const int B = 32;   // frames per scratch batch
    int rc;
    if ((rc = avd_ensure(ctx, (void **)&ctx->gray, &ctx->gray_cap, (size_t)B * H * W))) return rc;
    if (!htab->area_int)
        if ((rc = avd_ensure(ctx, (void **)&ctx->area_rows, &ctx->area_rows_cap, (size_t)B * H * 32 * sizeof(float)))) return rc;
    for (int f0 = 0; f0 < n; f0 += B) {
        int nb = n - f0 < B ? n - f0 : B;
        const uint8_t *src = d_bgr + (size_t)f0 * fstride;
        AVD_LAUNCH(ctx, "gray", st, gray_kernel<<<dim3((W + 255) / 256, H, nb), 256, 0, st>>>(src, H, W, pitch, fstride, ctx->gray));
        AVD_LAUNCH(ctx, "lap", st, lap_kernel<<<dim3(1, H, nb), 256, 0, st>>>(ctx->gray, H, W, d_rec + f0));
        if (htab->area_int) {
            AVD_LAUNCH(ctx, "area_int", st, area_int_kernel<<<(nb * 1024 + 255) / 256, 256, 0, st>>>(ctx->gray, H, W, tab, d_area32 + (size_t)f0 * 1024, nb));
        } else {
            int total = nb * H * 32;
            AVD_LAUNCH(ctx, "area_h", st, area_h_kernel<<<(total + 255) / 256, 256, 0, st>>>(ctx->gray, H, W, tab, ctx->area_rows, total));
            AVD_LAUNCH(ctx, "area_v", st, area_v_kernel<<<(nb * 1024 + 255) / 256, 256, 0, st>>>(ctx->area_rows, H, tab, d_area32 + (size_t)f0 * 1024, nb));
        }
        AVD_LAUNCH(ctx, "thumb", st, thumb_kernel<<<dim3((AVD_T + 127) / 128, AVD_T, nb), 128, 0, st>>>(ctx->gray, H, W, tab, d_thumbs + (size_t)f0 * AVD_T * AVD_T));
    }
    return AVD_OK;
}
