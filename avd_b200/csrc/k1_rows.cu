// K1, any-width fused path: one pass over the BGR bytes for frames the TMA streaming kernel (k1_fused.cu) does not take:
// widths that are not a multiple of 64 or below 1280 (854x480, 640x480, 1366x768, 960x540, odd crops ...), frames lower than
// 640 rows, buffers and row pitches of any alignment.  Round 1 sent those to a multi-kernel path over a gray scratch plane
// (~5 % of the HBM roofline), the first version of this file to scalar byte loops (19 %).
//
// Same decomposition as the streaming kernel: CTA = (frame, band b), band b = the source rows owned by INTER_AREA destination
// row b plus one halo row above and below (REFLECT_101 rows at the frame edges are loaded as the halo).  The band's gray rows
// live in shared memory (u8, 16-byte left pad, reflected columns -1 and W written once); four phases read them from there:
//   gray        video.py:5,43,51   16 pixels per thread and step: four ALIGNED 128-bit global loads covering the group's 48
//                                  bytes, a byte funnel shift by the group's misalignment (rows of a dense ragged-width frame
//                                  start at any byte), two IDP.2A per pixel on doubled 15-bit weights (as in k1_fused.cu)
//   Laplacian   video.py:52        a thread owns a 16-pixel column strip of a row chunk and marches down it with three gray
//                                  rows in registers, two pixels per register in 16-bit lanes biased by +1020; sum and sum of
//                                  squares through IDP (exact integers); columns >= W of the last strip are forced to L = 0
//   thumbnail   video.py:43        11-bit fixed point INTER_LINEAR; a thread owns a thumbnail column (taps and weights in
//                                  registers) and walks the thumbnail rows whose top tap row the band owns
//   INTER_AREA  video.py:6         exact integer cell sums: the Laplacian march also accumulates the column sums of the owned
//                                  rows; weighted by the fractional head / tail row they give V[x], and a cell is the sum
//                                  of V over its columns with the fractional head / tail column (all weights in 1/32
//                                  units, M = mean * W * H < 2^32).  The library's float32 result can differ from the exactly
//                                  rounded mean only within its summation error of a rounding boundary: those cells are
//                                  listed and recomputed in the library's ORDER by area_fix_kernel -- the
//                                  certificate of k1_fused.cu, with both axes fractional.
// Gray, Laplacian and thumbnail arithmetic per value is that of k1_generic.cu (which stays as the on-device cross-check and the
// fallback for bands that do not fit in shared memory); results are checked byte for byte against it in the tests.
#include "avd_internal.h"
#include <stdlib.h>

namespace {

constexpr int KR_THREADS = 256;
constexpr int KR_MAXTY = 24;           // thumbnail rows a band can own

struct KRParams {
    const uint8_t *bgr;
    size_t pitch, fstride;
    int H, W, GP;                      // GP: shared-memory pitch of a gray row: 16 (left pad) + 16 * ng + 16
    int ng;                            // 16-pixel groups per row
    uint32_t magic;                    // ceil(2^32 / ng): i / ng == umulhi(i, magic) for i * ng < 2^32
    uint32_t magic_f;                  // the same for W / 16 (strips that lie inside the frame)
    int nchunk;                        // row chunks per strip in the Laplacian phase
    int maxrows;                       // owned rows of the tallest band (fixes the shared-memory layout)
    uint32_t *fixlist;                 // [0] = count (zeroed by k1_rec_init), then frame * 1024 + cell (k1_fused.cu)
    int frame0;
    // per-band tables in the parameter bank: the first global loads of a CTA are frame bytes, not table entries
    unsigned short own[AVD_D + 1];     // ShapeTables::row_owner_begin
    unsigned short tyb[AVD_D + 1];     // ShapeTables::thumb_row_begin
    const ShapeTables *tab;
    avd_sample_rec *rec;
    uint8_t *thumbs, *area32;
};

__device__ __forceinline__ int r101(int i, int n) { return i < 0 ? -i : (i >= n ? 2 * n - 2 - i : i); }
__device__ __forceinline__ uint8_t sat_rne(float v)
{
    int r = __float2int_rn(v);
    return (uint8_t)min(max(r, 0), 255);
}

// 2*(3735,19235,9798): doubled 15-bit weights; acc = 2*(B*3735 + G*19235 + R*9798) + 32768, gray = acc >> 16 = byte 2
constexpr uint32_t WB = 7470u, WG = 38470u, WR = 19596u;
constexpr uint32_t C_BG = WB | (WG << 16), C_R0 = WR, C_0B = WB << 16, C_GR = WG | (WR << 16), RND = 32768u;
constexpr uint32_t KBIAS = 1020u | (1020u << 16);

// gray accumulators of 4 consecutive pixels held in 3 consecutive words
__device__ __forceinline__ uint32_t gray4(uint32_t w0, uint32_t w1, uint32_t w2)
{
    const uint32_t a0 = __dp2a_hi(C_R0, w0, __dp2a_lo(C_BG, w0, RND));      // B=w0.0 G=w0.1 R=w0.2
    const uint32_t a1 = __dp2a_lo(C_GR, w1, __dp2a_hi(C_0B, w0, RND));      // B=w0.3 G=w1.0 R=w1.1
    const uint32_t a2 = __dp2a_lo(C_R0, w2, __dp2a_hi(C_BG, w1, RND));      // B=w1.2 G=w1.3 R=w2.0
    const uint32_t a3 = __dp2a_hi(C_GR, w2, __dp2a_lo(C_0B, w2, RND));      // B=w2.1 G=w2.2 R=w2.3
    const uint32_t t01 = __byte_perm(a0, a1, 0x0062), t23 = __byte_perm(a2, a3, 0x0062);
    return __byte_perm(t01, t23, 0x5410);
}

// v[k] = bytes [4*WS + 4k + bs/8, +4) of the 64-byte window w[0..15]
template <int WS>
__device__ __forceinline__ void window_shift(const uint32_t *w, uint32_t bs, uint32_t *v)
{
#pragma unroll
    for (int k = 0; k < 12; k++) v[k] = __funnelshift_r(w[WS + k], w[WS + k + 1], bs);
}

// one gray row of a 16-pixel strip: E[g] = (p[4g], p[4g+2]), O[g] = (p[4g+1], p[4g+3]) in 16-bit lanes,
// LR = (pixel left of the strip, pixel right of the strip)
struct RowSet { uint32_t E[4], O[4], LR; };

__device__ __forceinline__ void load_rowset(const uint8_t *p, RowSet &R)
{
    const uint4 w = *(const uint4 *)p;
    const uint32_t l = p[-1], r = p[16];
    R.E[0] = __byte_perm(w.x, 0u, 0x4240); R.O[0] = __byte_perm(w.x, 0u, 0x4341);
    R.E[1] = __byte_perm(w.y, 0u, 0x4240); R.O[1] = __byte_perm(w.y, 0u, 0x4341);
    R.E[2] = __byte_perm(w.z, 0u, 0x4240); R.O[2] = __byte_perm(w.z, 0u, 0x4341);
    R.E[3] = __byte_perm(w.w, 0u, 0x4240); R.O[3] = __byte_perm(w.w, 0u, 0x4341);
    R.LR = l | (r << 16);
}

struct LapAcc { uint32_t ll, hl, hh, sb, cs[8]; };   // byte-split sums of (L+1020)^2, sum of L+1020; column sums of the centre rows (E[0..3], O[0..3])

// lanes of a register pair that hold pixels < valid (valid = pixels of the strip inside the frame, minus 4g)
__device__ __forceinline__ uint32_t lane_mask(int k0, int k1) { return (k0 > 0 ? 0xffffu : 0u) | (k1 > 0 ? 0xffff0000u : 0u); }

template <bool PARTIAL>
__device__ __forceinline__ void lap_row(const RowSet &U, const RowSet &C, const RowSet &D, LapAcc &A, int valid)
{
    uint32_t tot = 0;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        // left neighbours of (p0,p2) are (pL,p1); right neighbours of (p1,p3) are (p2,pR)
        const uint32_t el = g == 0 ? __byte_perm(C.LR, C.O[0], 0x5410) : __byte_perm(C.O[g > 0 ? g - 1 : 0], C.O[g], 0x5432);
        const uint32_t orr = g == 3 ? __byte_perm(C.E[3], C.LR, 0x7632) : __byte_perm(C.E[g], C.E[g < 3 ? g + 1 : 3], 0x5432);
        const uint32_t me = KBIAS - 4u * C.E[g], mo = KBIAS - 4u * C.O[g];
        uint32_t le = (el + C.O[g] + me) + U.E[g] + D.E[g];      // L+1020 of (p0,p2), each in [0,2040]
        uint32_t lo = (C.E[g] + orr + mo) + U.O[g] + D.O[g];     // L+1020 of (p1,p3)
        if (PARTIAL) {                                           // columns >= W: L = 0
            const uint32_t mE = lane_mask(valid - 4 * g, valid - 4 * g - 2), mO = lane_mask(valid - 4 * g - 1, valid - 4 * g - 3);
            le = (le & mE) | (KBIAS & ~mE);
            lo = (lo & mO) | (KBIAS & ~mO);
        }
        const uint32_t pl = __byte_perm(le, lo, 0x6420), ph = __byte_perm(le, lo, 0x7531);
        A.ll = __dp4a(pl, pl, A.ll);
        A.hl = __dp4a(ph, pl, A.hl);
        A.hh = __dp4a(ph, ph, A.hh);
        tot += le + lo;                                          // per lane <= 8 * 2040
        A.cs[g] += C.E[g]; A.cs[4 + g] += C.O[g];                // per lane <= 255 * rows of the band < 2^16
    }
    A.sb = __dp2a_lo(tot, 0x0101u, A.sb);
}

template <bool PARTIAL>
__device__ __forceinline__ void lap_chunk(const uint8_t *col, int GP, int qa, int qb, LapAcc &A, int valid)
{
    RowSet R0, R1, R2;
    load_rowset(col + (size_t)(qa - 1) * GP, R0);
    load_rowset(col + (size_t)qa * GP, R1);
    const uint8_t *p = col + (size_t)(qa + 1) * GP;
    int q = qa;
    while (true) {
        load_rowset(p, R2); p += GP; lap_row<PARTIAL>(R0, R1, R2, A, valid); if (++q == qb) break;
        load_rowset(p, R0); p += GP; lap_row<PARTIAL>(R1, R2, R0, A, valid); if (++q == qb) break;
        load_rowset(p, R1); p += GP; lap_row<PARTIAL>(R2, R0, R1, A, valid); if (++q == qb) break;
    }
}

// Two builds of the kernel.  PF (bands that leave room for three CTAs per SM at most): a thread's next group is loaded while it
// converts the current one, 79 registers.  !PF (small bands: 480p-class frames and below): no load prefetch, 64 registers, four
// CTAs per SM -- measured: 854x480 32.5 -> 33.9 %, 960x540 37 -> 40 %, 640x360 23 -> 25 %; 1918x1080 the other way round (45 against 36 %).
constexpr int KR_MAXQ = 136;           // sequence rows of a band (H / 32 + 3, H <= 4096)
constexpr int KR_MAXCH = 64;           // row chunks per strip

// AL: every row of the launch starts on a 16-byte boundary (16-byte aligned buffer, row pitch and frame stride: the upload ring
// of the host path, dense frames whose width is a multiple of 16): three loads per group, no byte shift.
template <bool PF, bool AL>
__global__ void __launch_bounds__(KR_THREADS, PF ? 3 : 4) k1_rows_kernel(KRParams P)
{
    extern __shared__ __align__(16) uint8_t sm[];
    const ShapeTables *__restrict__ tab = P.tab;
    const int tid = threadIdx.x, b = blockIdx.x, f = blockIdx.y;
    const int H = P.H, W = P.W, GP = P.GP, ng = P.ng;
    const int own0 = P.own[b], own1 = P.own[b + 1];
    const int n = own1 - own0, nseq = n + 2;                       // sequence row q = source row own0 - 1 + q (reflected)
    uint8_t *G = sm;                                               // [nseq][GP] gray; pixel x of row q at q * GP + 16 + x
    uint32_t *CS = (uint32_t *)(sm + (size_t)(P.maxrows + 2) * GP); // [ng][8] column sums of the owned rows, packed like a RowSet
    int *V = (int *)(CS + 8 * ng);                                 // [16 ng] weighted column sums (INTER_AREA)
    __shared__ long long red[2][KR_THREADS / 32];
    __shared__ int4 trow[KR_MAXTY];                                // per owned thumbnail row: tap row offsets in G, weights
    __shared__ int qtab[KR_MAXCH + 1];                             // first sequence row of chunk c
    const uint8_t *frame = P.bgr + (size_t)f * P.fstride;
    const int lane = tid & 31, wid = tid >> 5;
#ifdef KR_TIMING
    long long tck[9]; int tn = 0;
#define KR_TICK() do { tck[tn++] = clock64(); } while (0)
#else
#define KR_TICK() do {} while (0)
#endif
    KR_TICK();

    // ---- gray rows of the band: a thread per 16-pixel group (48 bytes); the loads of a thread's next group are in flight
    // while it converts the current one ---------------------------------------------------------------------------------------
    const int ty0 = P.tyb[b], nty = P.tyb[b + 1] - ty0;
    const int t2x2 = tab->thumb_2x2;
    const int nchunk = P.nchunk;
    LinTap hx;
    {
        const uint32_t base_lo = (uint32_t)(uintptr_t)frame;
        const int nitems = nseq * ng;
        const int gr = (W - 2) >> 4, kr = (W - 2) & 15;             // group and byte that hold pixel W - 2
        struct Item { uint4 c0, c1, c2, c3; uint32_t mis; int q, g; };
        // aligned chunk j holds group bytes [16j - mis, 16j - mis + 16): only chunks that hold a byte of the group are read
        // (an aligned chunk that holds one valid byte lies inside the allocation's granule)
        auto issue = [&](int i, Item &it) {
            it.q = (int)__umulhi((uint32_t)i, P.magic); it.g = i - it.q * ng;
            const uint32_t off = (uint32_t)r101(own0 - 1 + it.q, H) * (uint32_t)P.pitch + 48u * (uint32_t)it.g;
            it.mis = AL ? 0u : (base_lo + off) & 15u;
            const uint4 *al = (const uint4 *)(frame + off - it.mis);
            const int nbytes = 3 * W - 48 * it.g;                     // bytes of the row from the group's first byte on (> 0)
            const int mis = (int)it.mis;
            const uint4 z = make_uint4(0u, 0u, 0u, 0u);
            it.c0 = al[0];
            if (AL) {
                it.c1 = (16 < nbytes) ? al[1] : z;
                it.c2 = (32 < nbytes) ? al[2] : z;
                return;
            }
            it.c1 = (16 - mis < nbytes) ? al[1] : z;
            it.c2 = (32 - mis < nbytes) ? al[2] : z;
            it.c3 = (mis != 0 && 48 - mis < nbytes) ? al[3] : z;
        };
        auto convert = [&](const Item &it) {
            const uint32_t w[16] = {it.c0.x, it.c0.y, it.c0.z, it.c0.w, it.c1.x, it.c1.y, it.c1.z, it.c1.w,
                                    it.c2.x, it.c2.y, it.c2.z, it.c2.w, it.c3.x, it.c3.y, it.c3.z, it.c3.w};
            uint32_t v[12];
            if (AL) {
#pragma unroll
                for (int k = 0; k < 12; k++) v[k] = w[k];
            } else {
                const uint32_t bs = 8u * (it.mis & 3u);
                switch (it.mis >> 2) {
                case 0: window_shift<0>(w, bs, v); break;
                case 1: window_shift<1>(w, bs, v); break;
                case 2: window_shift<2>(w, bs, v); break;
                default: window_shift<3>(w, bs, v); break;
                }
            }
            uint4 o;
            o.x = gray4(v[0], v[1], v[2]); o.y = gray4(v[3], v[4], v[5]); o.z = gray4(v[6], v[7], v[8]); o.w = gray4(v[9], v[10], v[11]);
            uint8_t *dst = G + (size_t)it.q * GP + 16 + 16 * it.g;
            *(uint4 *)dst = o;                                        // pixels >= W of the last group: never used unmasked
            // REFLECT_101 columns, written by the thread that holds the source pixel: x = -1 -> 1, x = W -> W - 2 (after the
            // vector store, which may cover position W with a pixel from beyond the row)
            if (it.g == 0) dst[-1] = (uint8_t)(o.x >> 8);
            if (it.g == gr && (W & 15) != 1) {
                const uint32_t ow = (kr >> 2) == 0 ? o.x : ((kr >> 2) == 1 ? o.y : ((kr >> 2) == 2 ? o.z : o.w));
                dst[kr + 2] = (uint8_t)(ow >> (8 * (kr & 3)));
            }
        };
        Item A, B;
        int i = tid;
        if (i < nitems) issue(i, A);
        // per-band tables while the first loads are in flight (first used after the barrier below)
        if (tid < nty) {
            const int ty = ty0 + tid;
            int4 v;
            if (t2x2) v = make_int4((2 * ty - (own0 - 1)) * GP, (2 * ty + 1 - (own0 - 1)) * GP, 0, 0);
            else { const LinTap vy = tab->ly[ty]; v = make_int4((vy.i0 - (own0 - 1)) * GP, (vy.i1 - (own0 - 1)) * GP, vy.a0, vy.a1); }
            trow[tid] = v;
        }
        if (tid <= nchunk) qtab[tid] = 1 + (tid * n) / nchunk;
        for (int k = tid; k < 8 * ng; k += KR_THREADS) CS[k] = 0u;
        hx = tab->lx[tid];                                          // column tid; columns 256..319 are reloaded in their own pass
        if constexpr (PF) {
            while (i < nitems) {
                int j = i + KR_THREADS;
                if (j < nitems) issue(j, B);
                convert(A);
                i = j;
                if (i >= nitems) break;
                j = i + KR_THREADS;
                if (j < nitems) issue(j, A);
                convert(B);
                i = j;
            }
        } else {
            while (i < nitems) {
                convert(A);
                i += KR_THREADS;
                if (i < nitems) issue(i, A);
            }
        }
    }
    __syncthreads();
    KR_TICK();
    if ((W & 15) == 1) {          // pixel W - 2 and position W belong to different threads' groups: patch after the barrier
        for (int q = tid; q < nseq; q += KR_THREADS) G[(size_t)q * GP + 16 + W] = G[(size_t)q * GP + 16 + W - 2];
        __syncthreads();
    }
    KR_TICK();

    // ---- Laplacian sums and column sums over the owned rows (sequence rows 1 .. n): a thread per (16-pixel strip, row
    // chunk); the strip that sticks out of the frame (W % 16 != 0) runs in its own pass so that no warp executes both forms ----
    long long s = 0, ss = 0;
    {
        auto finish = [&](const LapAcc &A, int t, int rows) {
            const long long npx = 16ll * rows;
            const long long sl = (long long)A.sb - 1020ll * npx;
            s += sl;
            ss += (long long)((unsigned long long)A.ll + 512ull * A.hl + 65536ull * A.hh) - 2040ll * sl - 1040400ll * npx;
#pragma unroll
            for (int k = 0; k < 8; k++) atomicAdd(&CS[8 * t + k], A.cs[k]);
        };
        const int ngf = W >> 4;                                      // strips that lie inside the frame
        for (int it = tid; it < ngf * nchunk; it += KR_THREADS) {
            const int c = (int)__umulhi((uint32_t)it, P.magic_f), t = it - c * ngf;
            const int qa = qtab[c], qb = qtab[c + 1];
            if (qa >= qb) continue;
            LapAcc A{};
            lap_chunk<false>(G + 16 + 16 * t, GP, qa, qb, A, 16);
            finish(A, t, qb - qa);
        }
        if ((W & 15) && tid >= KR_THREADS - nchunk) {                // last warp(s): the lightest loaded ones of the loop above
            const int c = KR_THREADS - 1 - tid;
            const int qa = qtab[c], qb = qtab[c + 1];
            if (qa < qb) {
                LapAcc A{};
                lap_chunk<true>(G + 16 + 16 * ngf, GP, qa, qb, A, W & 15);
                finish(A, ngf, qb - qa);
            }
        }
        for (int o = 16; o; o >>= 1) { s += __shfl_down_sync(0xffffffffu, s, o); ss += __shfl_down_sync(0xffffffffu, ss, o); }
        if (lane == 0) { red[0][wid] = s; red[1][wid] = ss; }
    }
    KR_TICK();

    // ---- thumbnail: column tid over the band's thumbnail rows, then columns 256..319 spread over all threads ---------------
    {
        const uint8_t *Gx = G + 16;
        uint8_t *dst = P.thumbs + ((size_t)f * AVD_T + ty0) * AVD_T;
        auto px = [&](const LinTap &h, int j) -> uint32_t {
            const int4 ri = trow[j];
            const uint8_t *r0 = Gx + ri.x, *r1 = Gx + ri.y;
            // the right tap is column i0 + 1, except at the clamped right edge where its weight a1 is 0 (build_lin_taps):
            // reading column i0 + 1 there (position W holds the reflected pixel) multiplies it by 0
            const int S0 = r0[h.i0] * h.a0 + r0[h.i0 + 1] * h.a1;
            const int S1 = r1[h.i0] * h.a0 + r1[h.i0 + 1] * h.a1;
            return (uint32_t)((((ri.z * (S0 >> 4)) >> 16) + ((ri.w * (S1 >> 4)) >> 16) + 2) >> 2);
        };
        auto px2 = [&](int tx, int j) -> uint32_t {                  // 640x640: the library's 2x2 mean
            const int4 ri = trow[j];
            const uint8_t *r0 = Gx + ri.x + 2 * tx, *r1 = Gx + ri.y + 2 * tx;
            return (r0[0] + r0[1] + r1[0] + r1[1] + 2u) >> 2;
        };
        uint8_t *d = dst + tid;
        if (t2x2) for (int j = 0; j < nty; j++, d += AVD_T) *d = (uint8_t)px2(tid, j);
        else for (int j = 0; j < nty; j++, d += AVD_T) *d = (uint8_t)px(hx, j);
        for (int i = tid; i < (AVD_T - KR_THREADS) * nty; i += KR_THREADS) {
            const int j = i >> 6, tx = KR_THREADS + (i & 63);
            dst[j * AVD_T + tx] = (uint8_t)(t2x2 ? px2(tx, j) : px(tab->lx[tx], j));
        }
    }
    static_assert(AVD_T - KR_THREADS == 64, "second thumbnail pass assumes 64 remaining columns");
    KR_TICK();
    __syncthreads();
    KR_TICK();

    // ---- INTER_AREA destination row b (video.py:6).  Exact integer cell sums; the library's float32 result can only differ from
    // the exactly rounded mean when the mean is within the summation error bound of a rounding boundary: those cells go to
    // the fix list and are recomputed in the library's order by area_fix_kernel (certificate: see k1_fused.cu) ------------------
    // weights in 1/32 units: head row (sequence row 0) wh, owned rows 32, last owned row wt
    const int wh = 32 * own0 - b * H, wt = (b + 1) * H - 32 * (own1 - 1);
    if (tid < ng) {
        const uint4 c0 = *(const uint4 *)(CS + 8 * tid), c1 = *(const uint4 *)(CS + 8 * tid + 4);      // E[0..3], O[0..3]
        const uint4 p0 = *(const uint4 *)(G + 16 + 16 * tid), pn = *(const uint4 *)(G + (size_t)n * GP + 16 + 16 * tid);
        const uint32_t e[4] = {c0.x, c0.y, c0.z, c0.w}, o[4] = {c1.x, c1.y, c1.z, c1.w};
        const uint32_t a[4] = {p0.x, p0.y, p0.z, p0.w}, z[4] = {pn.x, pn.y, pn.z, pn.w};
        int v[16];
#pragma unroll
        for (int g = 0; g < 4; g++) {
            const int cs[4] = {(int)(e[g] & 0xffffu), (int)(o[g] & 0xffffu), (int)(e[g] >> 16), (int)(o[g] >> 16)};
#pragma unroll
            for (int j = 0; j < 4; j++)
                v[4 * g + j] = 32 * cs[j] + wh * (int)((a[g] >> (8 * j)) & 255u) - (32 - wt) * (int)((z[g] >> (8 * j)) & 255u);
        }
#pragma unroll
        for (int g = 0; g < 4; g++) *(int4 *)(V + 16 * tid + 4 * g) = make_int4(v[4 * g], v[4 * g + 1], v[4 * g + 2], v[4 * g + 3]);
    }
    __syncthreads();
    KR_TICK();
    if (tid == KR_THREADS - 1) {                                   // the frame's Laplacian totals, while warp 0 finishes the cells
        long long a = 0, c2 = 0;
        for (int w = 0; w < KR_THREADS / 32; w++) { a += red[0][w]; c2 += red[1][w]; }
        atomicAdd((unsigned long long *)&P.rec[f].lap_sum, (unsigned long long)a);
        atomicAdd((unsigned long long *)&P.rec[f].lap_sumsq, (unsigned long long)c2);
    }
    if (tid < 32) {
        // cell column dx covers [dx W / 32, (dx + 1) W / 32): head pixel xs0 - 1 (weight oh), pixels xs0 .. xs1 - 1 (32), the
        // last one of them with weight ot (exact integers: any order)
        const int dx = tid;
        const int xs0 = (dx * W + 31) >> 5, xs1 = ((dx + 1) * W + 31) >> 5;
        const int oh = 32 * xs0 - dx * W, ot = (dx + 1) * W - 32 * (xs1 - 1);
        uint32_t a0 = 0, a1 = 0;                                     // V <= 32 * 255 * (rows + 1), <= 130 columns: < 2^32
        int x = xs0;
        for (; x + 2 <= xs1; x += 2) { a0 += (unsigned)V[x]; a1 += (unsigned)V[x + 1]; }
        if (x < xs1) a0 += (unsigned)V[x];
        const unsigned long long acc = 32ull * ((unsigned long long)a0 + a1) + (unsigned long long)oh * (unsigned)V[max(xs0 - 1, 0)]
                                       - (unsigned long long)(32 - ot) * (unsigned)V[xs1 - 1];
        const uint32_t M = (uint32_t)acc;                            // = mean * W * H <= 255 W H < 2^32
        uint32_t out;
        bool close = false;
        if (tab->area_int) {
            out = sat_rne(__fmul_rn((float)(M >> 10), tab->inv_area));
        } else {
            // exact mean v = M / (W H).  close <=> |frac(v) - 0.5| <= 17/16 (nx + ny + 4) 2^-24 v, in integers (times 2 W H)
            const uint32_t Dn = (uint32_t)H * (uint32_t)W;
            const uint32_t k = M / Dn, r = M - k * Dn;
            const long long d2 = 2ll * r - (long long)Dn;
            const uint32_t bound = (uint32_t)(((unsigned long long)M * (uint32_t)(17 * ((W >> 5) + 2 + n + 1 + 4))) >> 27) + 2u;
            close = (unsigned long long)(d2 < 0 ? -d2 : d2) <= bound;
            out = k + (d2 >= 0 ? 1u : 0u);
        }
        P.area32[(size_t)f * 1024 + b * 32 + dx] = (uint8_t)out;
        // uncertified cells go to the fix list
        const uint32_t mask = __ballot_sync(0xffffffffu, close);
        if (mask) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(P.fixlist, (uint32_t)__popc(mask));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (close) P.fixlist[1 + base + __popc(mask & ((1u << lane) - 1u))] = (uint32_t)(P.frame0 + f) * 1024u + (uint32_t)(b * 32 + dx);
        }
    }
#ifdef KR_TIMING
    KR_TICK();
    if ((tid == 0 || tid == 255) && (b == 5 || b == 20) && (f == 100 || f == 200))
        printf("cta(%d,%d) t%d: gray %lld reflect %lld lap %lld thumb %lld wait %lld V %lld cells %lld total %lld\n", b, f, tid,
               tck[1] - tck[0], tck[2] - tck[1], tck[3] - tck[2], tck[4] - tck[3], tck[5] - tck[4], tck[6] - tck[5], tck[7] - tck[6], tck[7] - tck[0]);
#endif
}

int kr_max_rows(const ShapeTables *htab)
{
    int rows = 0;
    for (int b = 0; b < AVD_D; b++) rows = rows > htab->row_owner_begin[b + 1] - htab->row_owner_begin[b] ? rows : htab->row_owner_begin[b + 1] - htab->row_owner_begin[b];
    return rows;
}
int kr_groups(int W) { return (W + 15) >> 4; }
size_t kr_smem_bytes(int W, const ShapeTables *htab)
{
    const size_t ng = (size_t)kr_groups(W), GP = 16 * ng + 32;
    return (size_t)(kr_max_rows(htab) + 2) * GP + ng * 8 * sizeof(uint32_t) + ng * 16 * sizeof(int);
}

// Row chunks per strip in the Laplacian phase.  The kernel is bound by instruction issue, so what counts is warp
// instructions: warps needed for the (strip, chunk) items x the length of a chunk (~84 per row + ~80 per chunk for the two
// priming rows, the column-sum atomics and the 64-bit totals).  Ties go to the shorter chunks (less serial latency).
int kr_chunks(int W, const ShapeTables *htab)
{
    const int ngf = W >> 4, n = kr_max_rows(htab);
    int best = 1;
    long best_cost = -1;
    for (int k = 1; k <= n && k <= KR_MAXCH; k++) {
        const int items = ngf * k, warps = (items + 31) / 32, rounds = (items + KR_THREADS - 1) / KR_THREADS;
        const long cost = (long)(warps + (rounds - 1)) * (84L * ((n + k - 1) / k) + 80L);
        if (best_cost < 0 || cost <= best_cost) { best_cost = cost; best = k; }
    }
    return best;
}

}  // namespace

bool k1_rows_supported(int H, int W, const ShapeTables *htab)
{
    if (H < AVD_D || W < AVD_D || W > 4096 || H > 4096 || kr_smem_bytes(W, htab) > 200 * 1024) return false;
    if (kr_max_rows(htab) + 2 > KR_MAXQ) return false;
    // every row a band's thumbnail rows touch must be one of the band's own rows or its one-row halo
    for (int b = 0; b < AVD_D; b++) {
        const int own0 = htab->row_owner_begin[b], own1 = htab->row_owner_begin[b + 1];
        if (own1 <= own0) return false;
        if (htab->thumb_row_begin[b + 1] - htab->thumb_row_begin[b] > KR_MAXTY) return false;
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
            uint8_t *d_area32, uint32_t *d_fixlist, cudaStream_t st)
{
    if ((unsigned long long)pitch * (unsigned long long)H >= (1ull << 32)) { ctx->err = "k1_rows: frame larger than 4 GB"; return AVD_ERR_ARG; }
    KRParams P{};
    P.bgr = d_bgr; P.pitch = pitch; P.fstride = fstride; P.H = H; P.W = W;
    P.ng = kr_groups(W); P.GP = 16 * P.ng + 32;
    P.magic = (uint32_t)(0xffffffffu / (uint32_t)P.ng) + 1u;
    P.magic_f = (uint32_t)(0xffffffffu / (uint32_t)(W >> 4)) + 1u;
    P.nchunk = kr_chunks(W, htab);
    P.maxrows = kr_max_rows(htab);
    P.tab = tab; P.rec = d_rec; P.thumbs = d_thumbs; P.area32 = d_area32; P.fixlist = d_fixlist;
    for (int b = 0; b <= AVD_D; b++) { P.own[b] = (unsigned short)htab->row_owner_begin[b]; P.tyb[b] = (unsigned short)htab->thumb_row_begin[b]; }
    const size_t smem = kr_smem_bytes(W, htab);
    // four CTAs per SM fit (shared memory): the 64-register build without load prefetch
    static const int pf_env = getenv("AVD_KR_PF") ? atoi(getenv("AVD_KR_PF")) : -1;
    const bool pf = pf_env >= 0 ? pf_env != 0 : (smem + 2048) * 4 > 227 * 1024;
    static const bool al_off = getenv("AVD_KR_NOALIGNED") && atoi(getenv("AVD_KR_NOALIGNED"));
    // (only with the small-band build: with the prefetching build the aligned form measured slower, 1936x1080 46 against 52 %)
    const bool al = !al_off && !pf && pitch % 16 == 0 && fstride % 16 == 0 && ((uintptr_t)d_bgr % 16) == 0;
    void (*kern)(KRParams) = pf ? k1_rows_kernel<true, false> : (al ? k1_rows_kernel<false, true> : k1_rows_kernel<false, false>);
    static size_t smem_set[4][64] = {};
    size_t &set = smem_set[2 * pf + al][ctx->device & 63];
    if (smem > set) {
        AVD_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        set = smem;
    }
    for (int f0 = 0; f0 < n; f0 += 65535) {
        const int nb = n - f0 < 65535 ? n - f0 : 65535;
        KRParams Q = P;
        Q.bgr = P.bgr + (size_t)f0 * fstride; Q.rec = P.rec + f0;
        Q.thumbs = P.thumbs + (size_t)f0 * AVD_T * AVD_T; Q.area32 = P.area32 + (size_t)f0 * 1024;
        Q.frame0 = f0;
        AVD_LAUNCH(ctx, "k1_rows", st, kern<<<dim3(AVD_D, nb), KR_THREADS, smem, st>>>(Q));
    }
    return AVD_OK;
}
