// K1, any-width fused path: one pass over the BGR bytes for frames whose width is not a multiple of 16 (854x480, 1366x768,
// odd crops ...) or whose buffers are not 16-byte aligned -- shapes the TMA kernels (k1_fused.cu, k1_banded.cu) do not take
// and that used to fall to the multi-kernel generic path (5 launches over a gray scratch plane, ~5 % of the HBM roofline).
//
// Same decomposition as the fused kernels: CTA = (frame, band b), band b = the source rows owned by INTER_AREA
// destination row b plus one halo row above and below (REFLECT_101 rows at the frame edges are loaded as the halo).
// The band's gray rows live in shared memory (u8); everything else reads them from there:
//   gray        video.py:5,43,51   a warp per row, a lane per 4 pixels: aligned 32-bit loads + funnel shift by the row's misalignment
//   Laplacian   video.py:52        exact integer sums, one 64-bit atomic pair per CTA
//   INTER_AREA  video.py:6         the library's float32 arithmetic in the library's ORDER: horizontal pass,
//                                  one thread per (row, cell column), sequential adds; vertical pass, one thread per cell
//                                  column, sequential adds -- every row a cell needs is in the band
//   thumbnail   video.py:43        11-bit fixed point; the thumbnail rows whose top tap row the band owns
// Arithmetic per value is that of k1_generic.cu (which stays as the on-device cross-check and the fallback for bands that do
// not fit in shared memory): results are identical by construction and checked byte for byte in the tests.
#include "avd_internal.h"

namespace {

constexpr int KR_THREADS = 256;

struct KRParams {
    const uint8_t *bgr;
    size_t pitch, fstride;
    int H, W, Wp;                      // Wp: shared-memory row pitch (W rounded up to 4)
    const ShapeTables *tab;
    avd_sample_rec *rec;
    uint8_t *thumbs, *area32;
};

__device__ __forceinline__ uint32_t gray_of(uint32_t b, uint32_t g, uint32_t r) { return (3735u * b + 19235u * g + 9798u * r + 16384u) >> 15; }
__device__ __forceinline__ int r101(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }
__device__ __forceinline__ uint8_t sat_rne(float v)
{
    int r = __float2int_rn(v);
    return (uint8_t)min(max(r, 0), 255);
}

__global__ void __launch_bounds__(KR_THREADS) k1_rows_kernel(KRParams P)
{
    extern __shared__ __align__(16) uint8_t sm[];
    const ShapeTables *__restrict__ tab = P.tab;
    const int tid = threadIdx.x, b = blockIdx.x, f = blockIdx.y;
    const int H = P.H, W = P.W, Wp = P.Wp;
    const int own0 = tab->row_owner_begin[b], own1 = tab->row_owner_begin[b + 1];
    const int n = own1 - own0, nseq = n + 2;                       // sequence row q = source row own0 - 1 + q (reflected)
    uint8_t *G = sm;                                               // [nseq][Wp] gray
    float *rowsum = (float *)(sm + (((size_t)nseq * Wp + 15) & ~(size_t)15));   // [nseq][32] horizontal INTER_AREA sums
    __shared__ long long red[2][KR_THREADS / 32];
    const uint8_t *frame = P.bgr + (size_t)f * P.fstride;

    // ---- gray rows of the band: a warp per row, a lane per group of 4 pixels (12 bytes) -------------------------------------
    const int lane = tid & 31, wid = tid >> 5;
    const int groups = (W + 3) >> 2;
    for (int q = wid; q < nseq; q += KR_THREADS / 32) {
        const int r = r101(own0 - 1 + q, H);
        const uint8_t *rowp = frame + (size_t)r * P.pitch;
        uint32_t *grow = (uint32_t *)(G + (size_t)q * Wp);
        // a row of a dense ragged-width frame starts at any byte: aligned 32-bit loads + a funnel shift by the row's misalignment
        // (the same for every group of the row, since a group is 12 bytes)
        const uint32_t mis = (uint32_t)((uintptr_t)rowp & 3u), sh = 8u * mis;
        for (int g4 = lane; g4 < groups; g4 += 32) {
            const int x = 4 * g4;
            const uint8_t *src = rowp + 3 * x;
            uint32_t out;
            if (x + 4 <= W) {
                const uint32_t *al = (const uint32_t *)(src - mis);
                uint32_t w0 = al[0], w1 = al[1], w2 = al[2];
                if (mis) {                                       // warp-uniform
                    const uint32_t w3 = al[3];                   // holds bytes of this group: never past the row's last pixel
                    w0 = __funnelshift_r(w0, w1, sh); w1 = __funnelshift_r(w1, w2, sh); w2 = __funnelshift_r(w2, w3, sh);
                }
                out = gray_of(w0 & 255u, (w0 >> 8) & 255u, (w0 >> 16) & 255u)
                    | gray_of(w0 >> 24, w1 & 255u, (w1 >> 8) & 255u) << 8
                    | gray_of((w1 >> 16) & 255u, w1 >> 24, w2 & 255u) << 16
                    | gray_of((w2 >> 8) & 255u, (w2 >> 16) & 255u, w2 >> 24) << 24;
            } else {
                out = 0;
                for (int k = 0; k < 4 && x + k < W; k++) out |= gray_of(src[3 * k], src[3 * k + 1], src[3 * k + 2]) << (8 * k);
            }
            grow[g4] = out;
        }
    }
    __syncthreads();

    // ---- Laplacian sums over the owned rows (sequence rows 1 .. n): a warp per row, 4 pixels per lane and step -----------------
    long long s = 0;
    unsigned long long ss = 0;
    for (int q = 1 + wid; q <= n; q += KR_THREADS / 32) {
        const uint8_t *c = G + (size_t)q * Wp;
        int rs = 0;
        unsigned rss = 0;                                       // <= 128 pixels per lane and row at W <= 4096: < 2^31
        for (int x = 4 * lane; x < W; x += 128) {
            const uint32_t cw = *(const uint32_t *)(c + x), uw = *(const uint32_t *)(c + x - Wp), dw = *(const uint32_t *)(c + x + Wp);
            const int left = c[x == 0 ? 1 : x - 1];
            const int cnt = min(4, W - x);
            int pv[6];
            pv[0] = left;
#pragma unroll
            for (int k = 0; k < 4; k++) pv[1 + k] = (int)((cw >> (8 * k)) & 255u);
            pv[5] = x + 4 < W ? (int)c[x + 4] : 0;
            const bool lastg = x + cnt == W;
#pragma unroll
            for (int k = 0; k < 4; k++)
                if (k < cnt) {
                    const int right = (lastg && k + 1 == cnt) ? pv[k] : pv[k + 2];     // REFLECT_101: right of W-1 is W-2
                    const int L = (int)((uw >> (8 * k)) & 255u) + (int)((dw >> (8 * k)) & 255u) + pv[k] + right - 4 * pv[k + 1];
                    rs += L; rss += (unsigned)(L * L);
                }
        }
        s += rs; ss += rss;
    }
    {
        long long ss2 = (long long)ss;
        for (int o = 16; o; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); ss2 += __shfl_down_sync(0xffffffffu, ss2, o); }
        if (lane == 0) { red[0][wid] = s; red[1][wid] = ss2; }
    }

    // ---- thumbnail rows whose top tap row this band owns ---------------------------------------------------------------------
    {
        const int ty0 = tab->thumb_row_begin[b], nty = tab->thumb_row_begin[b + 1] - ty0;
        uint8_t *dst = P.thumbs + ((size_t)f * AVD_T + ty0) * AVD_T;
        for (int i = tid; i < nty * AVD_T; i += KR_THREADS) {
            const int ty = ty0 + i / AVD_T, tx = i % AVD_T;
            uint32_t out;
            if (tab->thumb_2x2) {
                const uint8_t *p = G + (size_t)(2 * ty - (own0 - 1)) * Wp + 2 * tx;
                out = (p[0] + p[1] + p[Wp] + p[Wp + 1] + 2u) >> 2;
            } else {
                const LinTap hx = tab->lx[tx], vy = tab->ly[ty];
                const uint8_t *r0 = G + (size_t)(vy.i0 - (own0 - 1)) * Wp, *r1 = G + (size_t)(vy.i1 - (own0 - 1)) * Wp;
                const int S0 = r0[hx.i0] * hx.a0 + r0[hx.i1] * hx.a1;
                const int S1 = r1[hx.i0] * hx.a0 + r1[hx.i1] * hx.a1;
                out = (uint32_t)((((vy.a0 * (S0 >> 4)) >> 16) + ((vy.a1 * (S1 >> 4)) >> 16) + 2) >> 2);
            }
            dst[i] = (uint8_t)out;
        }
    }

    // ---- INTER_AREA destination row b ----------------------------------------------------------------------------------------
    if (tab->area_int) {
        // integer scales: exact integer block sums (any order), one float multiply, round half to even
        int *isum = (int *)rowsum;
        const int isx = tab->isx;
        for (int i = tid; i < n * 32; i += KR_THREADS) {
            const int q = 1 + (i >> 5), dx = i & 31;
            const uint8_t *c = G + (size_t)q * Wp + dx * isx;
            int a = 0;
            for (int x = 0; x < isx; x++) a += c[x];
            isum[i] = a;
        }
        __syncthreads();
        if (tid < 32) {
            int a = 0;
            for (int q = 0; q < n; q++) a += isum[q * 32 + tid];
            P.area32[(size_t)f * 1024 + b * 32 + tid] = sat_rne(__fmul_rn((float)a, tab->inv_area));
        }
    } else {
        const AreaCell cy = tab->ay[b];
        const int ra = cy.has_head ? cy.s1 - 1 : cy.s1, rb = cy.has_tail ? cy.s2 : cy.s2 - 1;      // source rows of the cell
        const int nr = rb - ra + 1;
        for (int i = tid; i < nr * 32; i += KR_THREADS) {
            const int rr = i >> 5, dx = i & 31;
            const uint8_t *S = G + (size_t)(ra + rr - (own0 - 1)) * Wp;
            const AreaCell c = tab->ax[dx];
            float buf = 0.f;
            if (c.has_head) buf = __fadd_rn(buf, __fmul_rn((float)S[c.s1 - 1], c.ah));
            for (int x = c.s1; x < c.s2; x++) buf = __fadd_rn(buf, __fmul_rn((float)S[x], c.am));
            if (c.has_tail) buf = __fadd_rn(buf, __fmul_rn((float)S[c.s2], c.at));
            rowsum[i] = buf;
        }
        __syncthreads();
        if (tid < 32) {
            float sum = 0.f;
            int rr = 0;
            if (cy.has_head) { sum = __fadd_rn(sum, __fmul_rn(cy.ah, rowsum[tid])); rr = 1; }
            for (int r = cy.s1; r < cy.s2; r++, rr++) sum = __fadd_rn(sum, __fmul_rn(cy.am, rowsum[rr * 32 + tid]));
            if (cy.has_tail) sum = __fadd_rn(sum, __fmul_rn(cy.at, rowsum[rr * 32 + tid]));
            P.area32[(size_t)f * 1024 + b * 32 + tid] = sat_rne(sum);
        }
    }
    __syncthreads();
    if (tid == 0) {
        long long a = 0, c2 = 0;
        for (int w = 0; w < KR_THREADS / 32; w++) { a += red[0][w]; c2 += red[1][w]; }
        atomicAdd((unsigned long long *)&P.rec[f].lap_sum, (unsigned long long)a);
        atomicAdd((unsigned long long *)&P.rec[f].lap_sumsq, (unsigned long long)c2);
    }
}

size_t kr_smem_bytes(int H, int W, const ShapeTables *htab)
{
    int rows = 0;
    for (int b = 0; b < AVD_D; b++) rows = rows > htab->row_owner_begin[b + 1] - htab->row_owner_begin[b] ? rows : htab->row_owner_begin[b + 1] - htab->row_owner_begin[b];
    const size_t Wp = ((size_t)W + 3) & ~(size_t)3;
    return (((size_t)(rows + 2) * Wp + 15) & ~(size_t)15) + (size_t)(rows + 2) * 32 * sizeof(float);
}

}  // namespace

bool k1_rows_supported(int H, int W, const ShapeTables *htab)
{
    if (H < AVD_D || W < AVD_D || kr_smem_bytes(H, W, htab) > 200 * 1024) return false;
    // every row a band touches must be the band's own rows or its one-row halo
    for (int b = 0; b < AVD_D; b++) {
        const int own0 = htab->row_owner_begin[b], own1 = htab->row_owner_begin[b + 1];
        if (own1 <= own0) return false;
        if (!htab->area_int) {
            const AreaCell &c = htab->ay[b];
            const int ra = c.has_head ? c.s1 - 1 : c.s1, rb = c.has_tail ? c.s2 : c.s2 - 1;
            if (ra < own0 - 1 || rb > own1 || ra < 0 || rb >= H) return false;
        } else if (own1 - own0 != htab->isy) return false;
        for (int ty = htab->thumb_row_begin[b]; ty < htab->thumb_row_begin[b + 1]; ty++) {
            const int i0 = htab->thumb_2x2 ? 2 * ty : htab->ly[ty].i0, i1 = htab->thumb_2x2 ? 2 * ty + 1 : htab->ly[ty].i1;
            if (i0 < own0 - 1 || i1 > own1 || i0 < 0 || i1 >= H) return false;
            // the halo slots hold REFLECTED rows at the frame edges: a tap may only use a halo slot that holds the true row
            if ((i0 == own0 - 1 && own0 == 0) || (i1 == own1 && own1 == H)) return false;
        }
    }
    return true;
}

int k1_rows(avd_ctx *ctx, const uint8_t *d_bgr, int n, int H, int W, size_t pitch, size_t fstride,
            const ShapeTables *tab, const ShapeTables *htab, avd_sample_rec *d_rec, uint8_t *d_thumbs,
            uint8_t *d_area32, cudaStream_t st)
{
    KRParams P{};
    P.bgr = d_bgr; P.pitch = pitch; P.fstride = fstride; P.H = H; P.W = W; P.Wp = (W + 3) & ~3;
    P.tab = tab; P.rec = d_rec; P.thumbs = d_thumbs; P.area32 = d_area32;
    const size_t smem = kr_smem_bytes(H, W, htab);
    static size_t smem_set[64] = {};
    size_t &set = smem_set[ctx->device & 63];
    if (smem > set) {
        AVD_CUDA(ctx, cudaFuncSetAttribute(k1_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        set = smem;
    }
    for (int f0 = 0; f0 < n; f0 += 65535) {
        const int nb = n - f0 < 65535 ? n - f0 : 65535;
        KRParams Q = P;
        Q.bgr = P.bgr + (size_t)f0 * fstride; Q.rec = P.rec + f0;
        Q.thumbs = P.thumbs + (size_t)f0 * AVD_T * AVD_T; Q.area32 = P.area32 + (size_t)f0 * 1024;
        AVD_LAUNCH(ctx, "k1_rows", st, k1_rows_kernel<<<dim3(AVD_D, nb), KR_THREADS, smem, st>>>(Q));
    }
    return AVD_OK;
}
