// K1, fused path: ONE pass over each BGR frame produces everything video.py:36-52 derives from it
//   gray (video.py:5,43,51; never written to memory), Laplacian sum / sum of squares (video.py:52),
//   the 320x320 INTER_LINEAR thumbnail (video.py:43) and the 32x32 INTER_AREA cell means (video.py:6).
// HBM traffic = H*W*3 bytes read once + 102400 + 1024 + 16 bytes written per frame.
//
// Decomposition: CTA = (frame, band b), band b = the source rows of INTER_AREA destination row b
// (H/32 rows plus one halo row above and below for the Laplacian / fractional cell taps).
//   TMA ring        the band's rows stream into a ring of single-row shared-memory stages with TMA bulk
//                   copies (cp.async.bulk -> UBLKCP) completing on per-stage mbarriers; the warp that is
//                   LAST to finish reading a stage (shared-memory arrival counter) refills it, so nobody
//                   ever waits for a stage to become free
//   all warps       each thread owns a 16-pixel column strip (48 bytes = 3 conflict-free LDS.128) and
//                   marches down the band two rows at a time, keeping the previous rows in registers
//                   (two alternating register sets, so no register shuffling), writes packed gray to a
//                   small ring; every 4 rows all consumers meet once (named barrier) and then run the
//                   resampling passes (INTER_AREA row partials, thumbnail rows) out of the gray ring.
//
// Instruction budget: at 6.5 TB/s an SM has ~16 issue slots per pixel, so the arithmetic is packed:
//   gray        2 IDP.2A per pixel on 16-bit weights doubled so that Y lands in byte 2 of the accumulator
//   Laplacian   two pixels per register in 16-bit lanes, biased by +1020 to stay non-negative
//   sum/sumsq   IDP.2A / IDP.4A on the split low/high bytes of the biased response
//   INTER_AREA  float(g)*alpha as one exact FFMA on the magic-number float 2^23+g, then the mandated
//               sequential FADD (bit-exact with the library's float32 accumulation order)
#include "avd_internal.h"

namespace {

// build-time tunables (scripts/build_variants.sh builds A/B libraries with -D overrides)
#ifndef K1_NRS
#define K1_NRS 4
#endif
#ifndef K1_PRODUCER_SLEEP
#define K1_PRODUCER_SLEEP 1
#endif
#ifndef K1_TELESCOPE
#define K1_TELESCOPE 0
#endif
#ifndef K1_MINB
#define K1_MINB 5
#endif
constexpr int NRS = K1_NRS;           // single-row TMA stages in flight per CTA
constexpr int GR = 4;                 // rows per resample group (two row pairs)
constexpr int RC = 2 * GR + 1;        // gray ring rows: previous group + carry row + group being written
constexpr int MAX_TY = 32;            // thumbnail rows one band can own

struct K1Params {
    const uint8_t *bgr;
    size_t pitch, fstride;
    int H, W, row_bytes;
    const ShapeTables *tab;
    avd_sample_rec *rec;
    uint8_t *thumbs, *area32;
    int area_aligned;                 // INTER_AREA cells are whole aligned words in x
    int max_cell_rows;
};

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}" ::"r"(bar), "r"(parity) : "memory");
}
// producer-side wait: it waits most of the time, so give the scheduler a long suspend-time hint
__device__ __forceinline__ void consumer_sync(int n) { asm volatile("bar.sync 1, %0;" ::"r"(n) : "memory"); }

// 2*(3735,19235,9798): the doubled 15-bit weights; acc = 2*(...)+32768 so gray = acc >> 16 = byte 2
constexpr uint32_t WB = 7470u, WG = 38470u, WR = 19596u;
constexpr uint32_t C_BG = WB | (WG << 16), C_R0 = WR, C_0B = WB << 16, C_GR = WG | (WR << 16);
constexpr uint32_t RND = 32768u;
constexpr uint32_t KBIAS = 1020u | (1020u << 16);

// gray accumulators of 4 consecutive pixels held in 3 consecutive words
__device__ __forceinline__ void gray4(uint32_t w0, uint32_t w1, uint32_t w2, uint32_t &a0, uint32_t &a1, uint32_t &a2, uint32_t &a3)
{
    a0 = __dp2a_hi(C_R0, w0, __dp2a_lo(C_BG, w0, RND));      // B=w0.0 G=w0.1 R=w0.2
    a1 = __dp2a_lo(C_GR, w1, __dp2a_hi(C_0B, w0, RND));      // B=w0.3 G=w1.0 R=w1.1
    a2 = __dp2a_lo(C_R0, w2, __dp2a_hi(C_BG, w1, RND));      // B=w1.2 G=w1.3 R=w2.0
    a3 = __dp2a_hi(C_GR, w2, __dp2a_lo(C_0B, w2, RND));      // B=w2.1 G=w2.2 R=w2.3
}

struct LapAcc { uint32_t s, ll, hl, hh; };
// one gray row of a 16-pixel strip: even/odd pixels in 16-bit lanes, and base = left+right-4*centre+1020
struct RowSet { uint32_t E[4], O[4], BE[4], BO[4]; };

// biased Laplacian (L+1020 in [0,2040] per 16-bit lane) of the centre row and its statistics
//
// Only the sum of squares is accumulated per pixel.  The plain sum telescopes under REFLECT_101:
//   sum L = sum_x (g[1][x] + g[H-2][x] - g[0][x] - g[H-1][x]) + sum_y (g[y][1] + g[y][W-2] - g[y][0] - g[y][W-1])
// so it only needs the first/last two rows (4 rows per frame) and the first/last two pixels of every row.
__device__ __forceinline__ void lap_row(const RowSet &c, const RowSet &u, const RowSet &d, LapAcc &A, int cr, int H,
                                        bool first, bool last)
{
#if K1_TELESCOPE
    if (first) A.s += (c.O[0] & 0xffffu) - (c.E[0] & 0xffffu);        // + g[y][1] - g[y][0]
    if (last) A.s += (c.E[3] >> 16) - (c.O[3] >> 16);                 // + g[y][W-2] - g[y][W-1]
    if (cr <= 1 || cr >= H - 2) {                                     // warp-uniform, 4 rows per frame
        uint32_t t = 0;
#pragma unroll
        for (int g = 0; g < 4; g++) { t = __dp2a_lo(c.E[g], 0x0101u, t); t = __dp2a_lo(c.O[g], 0x0101u, t); }
        A.s += (cr == 1 || cr == H - 2) ? t : 0u - t;
    }
#endif
#pragma unroll
    for (int g = 0; g < 4; g++) {
        uint32_t le = c.BE[g] + u.E[g] + d.E[g];
        uint32_t lo = c.BO[g] + u.O[g] + d.O[g];
#if !K1_TELESCOPE
        A.s = __dp2a_lo(le, 0x0101u, A.s);
        A.s = __dp2a_lo(lo, 0x0101u, A.s);
#endif
        uint32_t pl = __byte_perm(le, lo, 0x6420), ph = __byte_perm(le, lo, 0x7531);
        A.ll = __dp4a(pl, pl, A.ll);
        A.hl = __dp4a(ph, pl, A.hl);
        A.hh = __dp4a(ph, ph, A.hh);
    }
}

// convert one staged BGR row strip to gray: fills the row set and stores the 16 packed gray bytes
__device__ __forceinline__ void load_row(const uint8_t *sp, uint8_t *gdst, bool first, bool last, RowSet &R)
{
    const uint4 q0 = *(const uint4 *)sp, q1 = *(const uint4 *)(sp + 16), q2 = *(const uint4 *)(sp + 32);
    const uint32_t wl = *(const uint32_t *)(sp - 4), wr = *(const uint32_t *)(sp + 48);
    uint32_t a[16];
    gray4(q0.x, q0.y, q0.z, a[0], a[1], a[2], a[3]);
    gray4(q0.w, q1.x, q1.y, a[4], a[5], a[6], a[7]);
    gray4(q1.z, q1.w, q2.x, a[8], a[9], a[10], a[11]);
    gray4(q2.y, q2.z, q2.w, a[12], a[13], a[14], a[15]);
    // neighbour pixels x-1 and x+16 (REFLECT_101 at the frame edges: x=-1 -> 1, x=W -> W-2)
    uint32_t aL = __dp2a_hi(C_GR, wl, __dp2a_lo(C_0B, wl, RND));
    uint32_t aR = __dp2a_hi(C_R0, wr, __dp2a_lo(C_BG, wr, RND));
    aL = first ? a[1] : aL;
    aR = last ? a[14] : aR;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        R.E[g] = __byte_perm(a[4 * g], a[4 * g + 2], 0x7632);       // (g0, g2) in 16-bit lanes
        R.O[g] = __byte_perm(a[4 * g + 1], a[4 * g + 3], 0x7632);   // (g1, g3)
    }
    uint4 gw;
    gw.x = __byte_perm(R.E[0], R.O[0], 0x6240); gw.y = __byte_perm(R.E[1], R.O[1], 0x6240);
    gw.z = __byte_perm(R.E[2], R.O[2], 0x6240); gw.w = __byte_perm(R.E[3], R.O[3], 0x6240);
    *(uint4 *)gdst = gw;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        // left neighbours of (g0,g2) are (gL,g1); right neighbours of (g1,g3) are (g2,gR)
        uint32_t el = g == 0 ? __byte_perm(aL, R.O[0], 0x5432) : __byte_perm(R.O[g - 1], R.O[g], 0x5432);
        uint32_t orr = g == 3 ? __byte_perm(R.E[3], aR, 0x7632) : __byte_perm(R.E[g], R.E[g + 1], 0x5432);
        const uint32_t me = KBIAS - 4u * R.E[g], mo = KBIAS - 4u * R.O[g];   // one IMAD each, then one IADD3
        R.BE[g] = el + R.O[g] + me;
        R.BO[g] = R.E[g] + orr + mo;
    }
}

__device__ __forceinline__ uint8_t sat_rne(float v)
{
    int r = __float2int_rn(v);
    return (uint8_t)min(max(r, 0), 255);
}

// exact float32(g * alpha) for a byte g: fma((2^23+g), alpha, -2^23*alpha) rounds once
__device__ __forceinline__ float mulbyte(uint32_t word, uint32_t sel, float alpha, float nbias)
{
    return __fmaf_rn(__uint_as_float(__byte_perm(word, 0x4B000000u, sel)), alpha, nbias);
}

struct Band {
    int own0, own1, lo, hi, H, W, nrows, nwarps;
    uint32_t bar_full, cnt;               // shared addresses: full barriers [NRS], arrival counters [NRS]
    const uint8_t *stage;
    uint8_t *ring;
    int row_bytes;
    const uint8_t *src;                   // first loaded row of the band in global memory
    size_t pitch;
    uint64_t pol_keep, pol_stream;
    bool keep_first_rows;
};

__device__ __forceinline__ void issue_row(const Band &B, int i)
{
    const int s = i % NRS;
    mbar_expect_tx(B.bar_full + 8 * s, (uint32_t)B.row_bytes);
    // the first two rows of a band are read again, later, by the band above: keep them in L2
    tma_load_1d(smem_u32(B.stage + s * B.row_bytes), B.src + (size_t)i * B.pitch, (uint32_t)B.row_bytes, B.bar_full + 8 * s,
                (i < 2 && B.keep_first_rows) ? B.pol_keep : B.pol_stream);
}

// this warp is done reading stage s (row i): the last warp to say so refills the stage with row i+NRS
__device__ __forceinline__ void release_stage(const Band &B, int i)
{
    const int s = i % NRS;
    uint32_t old;
    asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(B.cnt + 4 * s) : "memory");
    if (old == (uint32_t)(B.nwarps - 1)) {
        asm volatile("st.relaxed.cta.shared::cta.u32 [%0], %1;" ::"r"(B.cnt + 4 * s), "r"(0u) : "memory");
        if (i + NRS < B.nrows) issue_row(B, i + NRS);
    }
}

// rows (r0, r0+1) of the band: p0/p1 hold rows r0-2, r0-1; c0/c1 receive rows r0, r0+1
__device__ __forceinline__ void pair_step(const Band &B, int i0, int r0, bool have1, int slot0, int slot1, int tid, bool active,
                                          bool first, bool last, const RowSet &p0, const RowSet &p1, RowSet &c0, RowSet &c1, LapAcc &A)
{
    const int s0 = i0 % NRS, s1 = (i0 + 1) % NRS;
    mbar_wait(B.bar_full + 8 * s0, (uint32_t)((i0 / NRS) & 1));
    if (active) load_row(B.stage + s0 * B.row_bytes + 48 * tid, B.ring + slot0 * B.W + 16 * tid, first, last, c0);
    if (have1) {
        mbar_wait(B.bar_full + 8 * s1, (uint32_t)(((i0 + 1) / NRS) & 1));
        if (active) load_row(B.stage + s1 * B.row_bytes + 48 * tid, B.ring + slot1 * B.W + 16 * tid, first, last, c1);
    }
    __syncwarp();
    if ((tid & 31) == 0) {                       // this warp is done with both stages
        release_stage(B, i0);
        if (have1) release_stage(B, i0 + 1);
    }
    if (!active) return;
    const int cr = r0 - 1;                       // row r0-1 is now complete (down neighbour = c0)
    if (cr >= B.own0 && cr < B.own1) lap_row(p1, p0, c0, A, cr, B.H, first, last);
    if (have1) {
        if (r0 >= B.own0 && r0 < B.own1) {
            if (r0 == 0) lap_row(c0, c1, c1, A, r0, B.H, first, last);  // top frame edge: row -1 reflects to row 1
            else lap_row(c0, p1, c1, A, r0, B.H, first, last);
        }
        if (r0 + 1 == B.H - 1 && r0 + 1 < B.own1) lap_row(c1, c0, c0, A, r0 + 1, B.H, first, last);   // bottom frame edge: row H reflects to H-2
    } else if (r0 == B.H - 1 && r0 < B.own1) {
        lap_row(c0, p1, p1, A, r0, B.H, first, last);
    }
}

template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k1_banded_kernel(K1Params P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x;
    const int cons = blockDim.x;                 // every warp is a consumer
    const int b = blockIdx.x, f = blockIdx.y;
    const ShapeTables *__restrict__ tab = P.tab;
    const int H = P.H, W = P.W;
    const int own0 = tab->row_owner_begin[b], own1 = tab->row_owner_begin[b + 1];
    const int lo = max(own0 - 1, 0), hi = min(own1, H - 1);
    const int nrows = hi - lo + 1;
    // shared memory carve-up
    uint8_t *stage = smem + 128;                                   // NRS rows of BGR
    uint8_t *ring = stage + NRS * P.row_bytes;                     // RC rows of gray
    float *rowbuf = (float *)(ring + RC * W);                      // INTER_AREA row partials, double-buffered [2][GR][32]
    uint2 *xt = (uint2 *)(rowbuf + 2 * GR * 32);                   // packed x taps of the thumbnail [320]
    int4 *yt = (int4 *)(xt + AVD_T);                               // y taps of this band's thumbnail rows
    const uint32_t bar_full = smem_u32(smem), cnt = bar_full + 8 * NRS;
    const int nwarps_c = cons >> 5;

    if (tid == 0) {
        for (int s = 0; s < NRS; s++) {
            mbar_init(bar_full + 8 * s, 1);
            asm volatile("st.relaxed.cta.shared::cta.u32 [%0], %1;" ::"r"(cnt + 4 * s), "r"(0u) : "memory");
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    const int ty_begin = tab->thumb_row_begin[b], nty = tab->thumb_row_begin[b + 1] - ty_begin;
    const int thumb_2x2 = tab->thumb_2x2;
    for (int i = tid; i < AVD_T; i += blockDim.x) {
        LinTap t = tab->lx[i];
        if (thumb_2x2) { t.i0 = 2 * i; t.i1 = 2 * i + 1; }
        xt[i] = make_uint2((uint32_t)t.i0 | ((uint32_t)t.i1 << 16), (uint32_t)t.a0 | ((uint32_t)t.a1 << 16));
    }
    for (int i = tid; i < nty; i += blockDim.x) {
        LinTap t = tab->ly[ty_begin + i];
        if (thumb_2x2) { t.i0 = 2 * (ty_begin + i); t.i1 = t.i0 + 1; }
        yt[i] = make_int4(t.i0, t.i1, t.a0, t.a1);
    }
    __syncthreads();

    const bool active = 16 * tid < W, first = tid == 0, last = 16 * tid + 16 == W;
    Band B;
    B.own0 = own0; B.own1 = own1; B.lo = lo; B.hi = hi; B.H = H; B.W = W; B.nrows = nrows; B.nwarps = nwarps_c;
    B.bar_full = bar_full; B.cnt = cnt; B.stage = stage; B.ring = ring; B.row_bytes = P.row_bytes;
    B.src = P.bgr + (size_t)f * P.fstride + (size_t)lo * P.pitch; B.pitch = P.pitch; B.keep_first_rows = b > 0;
    asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(B.pol_keep));
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(B.pol_stream));
    if (tid == 0)
        for (int i = 0; i < min(NRS, nrows); i++) issue_row(B, i);
    RowSet A0, A1, B0, B1;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        A0.E[g] = A0.O[g] = A0.BE[g] = A0.BO[g] = 0; A1.E[g] = A1.O[g] = A1.BE[g] = A1.BO[g] = 0;
        B0.E[g] = B0.O[g] = B0.BE[g] = B0.BO[g] = 0; B1.E[g] = B1.O[g] = B1.BE[g] = B1.BO[g] = 0;
    }
    LapAcc A = {0, 0, 0, 0};
    const AreaCell cy = tab->ay[b];
    const int cell_lo = cy.has_head ? cy.s1 - 1 : cy.s1, cell_hi = cy.has_tail ? cy.s2 : cy.s2 - 1;
    const int area_int = tab->area_int, isx = tab->isx;
    const int cw = W >> 5;
    const float alpha = tab->ax[0].am, nbias = -8388608.0f * alpha;
    int jty = 0;                                   // next thumbnail row of this band
    int gslot = 0;                                 // ring slot of the first row of the current group
    // INTER_AREA vertical pass, folded progressively by thread dx = tid < 32 (sequential over rows, as mandated)
    float vsum = 0.f;
    int vtot = 0, pg0 = 0, pgn = 0, gi = 0;
    auto fold = [&](int fg0, int fgn, const float *buf) {
        const int r0 = max(fg0, cell_lo), r1 = min(fg0 + fgn - 1, cell_hi);
        for (int r = r0; r <= r1; r++) {
            const float v = buf[(r - fg0) * 32 + tid];
            if (area_int) vtot += __float_as_int(v);
            else {
                const float beta = (cy.has_head && r == cell_lo) ? cy.ah : ((cy.has_tail && r == cy.s2) ? cy.at : cy.am);
                vsum = __fadd_rn(vsum, __fmul_rn(beta, v));
            }
        }
    };

    for (int i = 0; i < nrows; i += GR, gi++) {
        const int g0 = lo + i, gn = min(GR, nrows - i);
        // ---- two row pairs: BGR -> gray ring + Laplacian statistics ----
        {
            int sl0 = gslot, sl1 = gslot + 1; if (sl1 >= RC) sl1 -= RC;
            pair_step(B, i, g0, gn > 1, sl0, sl1, tid, active, first, last, B0, B1, A0, A1, A);
        }
        if (gn > 2) {
            int sl0 = gslot + 2; if (sl0 >= RC) sl0 -= RC;
            int sl1 = gslot + 3; if (sl1 >= RC) sl1 -= RC;
            pair_step(B, i + 2, g0 + 2, gn > 3, sl0, sl1, tid, active, first, last, A0, A1, B0, B1, A);
        } else {
            // the band ended on an odd pair: keep the register roles consistent (nothing follows)
        }
        consumer_sync(cons);                       // the group's gray rows are visible to every consumer
        const int g1 = g0 + gn;
        float *rb = rowbuf + (gi & 1) * GR * 32;
        if (tid < 32 && pgn) fold(pg0, pgn, rowbuf + ((gi & 1) ^ 1) * GR * 32);   // previous group's partials
        pg0 = g0; pgn = gn;

        // ---- INTER_AREA horizontal pass for the cell rows of this group: one task per (row, dx) ----
        const int ar0 = max(g0, cell_lo), ar1 = min(g1 - 1, cell_hi);
        for (int id = tid; id < (ar1 - ar0 + 1) * 32; id += cons) {
            const int row = ar0 + (id >> 5), dx = id & 31;
            int sl = gslot + (row - g0); if (sl >= RC) sl -= RC;
            const uint8_t *gr = ring + sl * W;
            float out;
            if (area_int) {
                unsigned sum = 0;
                if (P.area_aligned) {
                    const uint32_t *wp = (const uint32_t *)(gr + dx * isx);
                    for (int k = 0; k < (isx >> 2); k++) sum = __dp4a(wp[k], 0x01010101u, sum);
                } else {
                    for (int k = 0; k < isx; k++) sum += gr[dx * isx + k];
                }
                out = __int_as_float(sum);
            } else if (P.area_aligned) {
                const uint32_t *wp = (const uint32_t *)(gr + dx * cw);
                float buf = 0.f;
                for (int k = 0; k < (cw >> 2); k++) {
                    uint32_t w = wp[k];
                    buf = __fadd_rn(buf, mulbyte(w, 0x7650, alpha, nbias));
                    buf = __fadd_rn(buf, mulbyte(w, 0x7651, alpha, nbias));
                    buf = __fadd_rn(buf, mulbyte(w, 0x7652, alpha, nbias));
                    buf = __fadd_rn(buf, mulbyte(w, 0x7653, alpha, nbias));
                }
                out = buf;
            } else {
                const AreaCell cx = tab->ax[dx];
                float buf = 0.f;
                if (cx.has_head) buf = __fadd_rn(buf, __fmul_rn((float)gr[cx.s1 - 1], cx.ah));
                for (int k = cx.s1; k < cx.s2; k++) buf = __fadd_rn(buf, __fmul_rn((float)gr[k], cx.am));
                if (cx.has_tail) buf = __fadd_rn(buf, __fmul_rn((float)gr[cx.s2], cx.at));
                out = buf;
            }
            rb[(row - g0) * 32 + dx] = out;
        }
        // ---- thumbnail rows whose lower tap row is now in the ring ----
        while (jty < nty) {
            const int4 vy = yt[jty];
            if (vy.y >= g1) break;
            int sa = gslot + (vy.x - g0); if (sa < 0) sa += RC; if (sa >= RC) sa -= RC;
            int sb = gslot + (vy.y - g0); if (sb < 0) sb += RC; if (sb >= RC) sb -= RC;
            const uint8_t *p0 = ring + sa * W, *p1 = ring + sb * W;
            uint8_t *dst = P.thumbs + ((size_t)f * AVD_T + (ty_begin + jty)) * AVD_T;
            for (int tx = tid; tx < AVD_T; tx += cons) {
                const uint2 vx = xt[tx];
                const int x0 = vx.x & 0xffff, x1 = vx.x >> 16;
                uint8_t o;
                if (thumb_2x2) {
                    o = (uint8_t)((p0[x0] + p0[x1] + p1[x0] + p1[x1] + 2) >> 2);
                } else {
                    const int a0 = vx.y & 0xffff, a1 = vx.y >> 16;
                    int S0 = p0[x0] * a0 + p0[x1] * a1;
                    int S1 = p1[x0] * a0 + p1[x1] * a1;
                    o = (uint8_t)((((vy.z * (S0 >> 4)) >> 16) + ((vy.w * (S1 >> 4)) >> 16) + 2) >> 2);
                }
                dst[tx] = o;
            }
            jty++;
        }
        gslot += gn; if (gslot >= RC) gslot -= RC;
    }

    // ---- Laplacian sums: warp shuffle -> shared -> one 64-bit atomic pair per CTA ----
#if K1_TELESCOPE
    unsigned long long s1 = (unsigned long long)(long long)(int)A.s;
#else
    unsigned long long s1 = A.s;
#endif
    unsigned long long sq = (unsigned long long)A.ll + 512ull * A.hl + 65536ull * A.hh;
    for (int o = 16; o; o >>= 1) { s1 += __shfl_down_sync(0xffffffffu, s1, o); sq += __shfl_down_sync(0xffffffffu, sq, o); }
    __shared__ unsigned long long red[2][32];
    if ((tid & 31) == 0) { red[0][tid >> 5] = s1; red[1][tid >> 5] = sq; }
    consumer_sync(cons);                           // also orders the last group's row partials
    if (tid < 32) {
        s1 = tid < nwarps_c ? red[0][tid] : 0; sq = tid < nwarps_c ? red[1][tid] : 0;
        for (int o = 16; o; o >>= 1) { s1 += __shfl_down_sync(0xffffffffu, s1, o); sq += __shfl_down_sync(0xffffffffu, sq, o); }
        if (tid == 0) {
            const long long npx = (long long)(own1 - own0) * W;
#if K1_TELESCOPE
            const long long sumL = (long long)s1;                                   // telescoped border sums
#else
            const long long sumL = (long long)s1 - 1020ll * npx;                    // s1 = sum (L+1020)
#endif
            const long long sumL2 = (long long)sq - 2040ll * sumL - 1040400ll * npx;       // sq = sum (L+1020)^2
            atomicAdd((unsigned long long *)&P.rec[f].lap_sum, (unsigned long long)sumL);
            atomicAdd((unsigned long long *)&P.rec[f].lap_sumsq, (unsigned long long)sumL2);
        }
        // ---- INTER_AREA vertical pass: destination row b, column tid ----
        if (pgn) fold(pg0, pgn, rowbuf + ((gi & 1) ^ 1) * GR * 32);      // gi was incremented past the last group
        P.area32[(size_t)f * 1024 + b * 32 + tid] = area_int ? sat_rne(__fmul_rn((float)vtot, tab->inv_area)) : sat_rne(vsum);
    }
}

}  // namespace

static size_t k1b_smem_bytes(int W, int max_cell_rows)
{
    (void)max_cell_rows;
    return 128 + (size_t)NRS * 3 * W + (size_t)RC * W + (size_t)2 * GR * 32 * sizeof(float) + AVD_T * sizeof(uint2) +
           MAX_TY * sizeof(int4);
}

bool k1_banded_supported(int H, int W, size_t pitch, size_t fstride, const void *base)
{
    return W % 16 == 0 && W <= 4096 && pitch % 16 == 0 && fstride % 16 == 0 && ((uintptr_t)base % 16) == 0 && H >= 32 &&
           k1b_smem_bytes(W, H / 32 + 3) <= 227 * 1024;
}

int k1_banded(avd_ctx *ctx, const uint8_t *d_bgr, int n, int H, int W, size_t pitch, size_t fstride,
             const ShapeTables *tab, const ShapeTables *htab, avd_sample_rec *d_rec, uint8_t *d_thumbs,
             uint8_t *d_area32, cudaStream_t st)
{
    K1Params P{};
    P.bgr = d_bgr; P.pitch = pitch; P.fstride = fstride; P.H = H; P.W = W; P.row_bytes = 3 * W;
    P.tab = tab; P.rec = d_rec; P.thumbs = d_thumbs; P.area32 = d_area32;
    P.area_aligned = htab->area_int ? (htab->isx % 4 == 0) : (W % 128 == 0);
    int max_cell_rows = 0;
    for (int b = 0; b < AVD_D; b++) {
        const AreaCell &c = htab->ay[b];
        int rows = (c.has_tail ? c.s2 : c.s2 - 1) - (c.has_head ? c.s1 - 1 : c.s1) + 1;
        if (rows > max_cell_rows) max_cell_rows = rows;
        if (htab->thumb_row_begin[b + 1] - htab->thumb_row_begin[b] > MAX_TY) { ctx->err = "fused K1: thumbnail rows per band"; return AVD_ERR_ARG; }
    }
    P.max_cell_rows = max_cell_rows;
    size_t smem = k1b_smem_bytes(W, max_cell_rows);
    if (smem > 227 * 1024) { ctx->err = "fused K1: frame too wide for shared memory"; return AVD_ERR_ARG; }
    int threads = ((W / 16) + 31) / 32 * 32;
    const bool narrow = threads <= 128;                              // <= 2048 px wide: 5 CTAs / SM, <= 102 registers
    static size_t smem_set_dev[64][2] = {};          // per device: function attributes belong to the device's context
    size_t *smem_set = smem_set_dev[ctx->device & 63];
    if (smem > smem_set[narrow]) {
        if (narrow) AVD_CUDA(ctx, cudaFuncSetAttribute(k1_banded_kernel<128, K1_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else AVD_CUDA(ctx, cudaFuncSetAttribute(k1_banded_kernel<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set[narrow] = smem;
    }
    for (int f0 = 0; f0 < n; f0 += 65535) {                          // grid.y limit
        int nb = n - f0 < 65535 ? n - f0 : 65535;
        K1Params Q = P;
        Q.bgr = d_bgr + (size_t)f0 * fstride; Q.rec = d_rec + f0;
        Q.thumbs = d_thumbs + (size_t)f0 * AVD_T * AVD_T; Q.area32 = d_area32 + (size_t)f0 * 1024;
        if (narrow) AVD_LAUNCH(ctx, "k1_banded", st, (k1_banded_kernel<128, K1_MINB><<<dim3(AVD_D, nb), threads, smem, st>>>(Q)));
        else AVD_LAUNCH(ctx, "k1_banded", st, (k1_banded_kernel<256, 2><<<dim3(AVD_D, nb), threads, smem, st>>>(Q)));
    }
    return AVD_OK;
}
