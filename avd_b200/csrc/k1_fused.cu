// K1, fused streaming path: ONE pass over each BGR frame produces everything video.py:36-52 derives from it
//   gray (video.py:5,43,51; never written to memory), Laplacian sum / sum of squares (video.py:52),
//   the 320x320 INTER_LINEAR thumbnail (video.py:43) and the 32x32 INTER_AREA cell means (video.py:6).
// HBM traffic = H*W*3 bytes read once + 102400 + 1024 + 16 bytes written per frame.
//
// The kernel is bound by instruction issue, not by HBM, unless it spends fewer than ~17 lane-instructions per
// pixel (148 SMs x 128 lanes x 1.965 GHz / 2.18 Tpx/s at the measured copy bandwidth), and the two integer
// pipes (ALU: PRMT/IADD3/LOP3, FMA-heavy: IMAD/IDP) each take only one warp instruction every other cycle
// (scripts/ubench_pipes.cu).  Everything below is organised around that budget.
//
// Decomposition: CTA = (frame, band b); band b = the source rows owned by INTER_AREA destination row b plus one
// halo row above and below (REFLECT_101 rows at the frame edges are simply loaded as the halo).
//   TMA ring     rows stream into NRS single-row shared-memory stages (cp.async.bulk -> UBLKCP, one mbarrier per
//                stage); the warp that is LAST to finish reading a stage (mbarrier arrive + pending count) refills it
//                from registers, nobody waits for a free stage; the next stage's barrier is probed a row ahead
//   threads      a thread owns a 16-pixel column strip (48 bytes = 3 conflict-free LDS.128) and marches down the
//                band keeping three rows in registers (straight-line code, three rotating register sets):
//                gray        2 IDP.2A per pixel (15-bit weights doubled so that Y lands in byte 2)
//                Laplacian   two pixels per register in 16-bit lanes, biased by +1020; sum of squares through
//                            byte-split IDP.4A; the plain sum telescopes to strip/band borders
//                INTER_AREA  exact integer cell sums (one IADD3 per 4 pixels) -- see "certificate" below
//                thumbnail   the thread that owns a strip also owns the thumbnail columns whose left tap lies in
//                            it; tap values pass through a private 48-byte shared-memory slot (which register holds
//                            a given tap differs from thread to thread; the slot offset is per-thread data) or, for
//                            W = 1280 / 2560, come straight out of the registers.  Thread-private, so the warps never
//                            meet a block barrier inside the band; the band's thumbnail rows are staged in shared
//                            memory and written out by its last warp.
//
// INTER_AREA certificate.  The library accumulates float32 products sequentially; the result is
// rounded to u8 and only that byte is observable.  The kernel computes the exact rational cell mean from integer
// sums; float32 evaluation in the library's order differs from it by at most (nx+ny+4) * 2^-24 relative
// (standard forward error bound of recursive summation, nx/ny = taps per cell), so whenever the exact mean is
// farther than that bound (taken with a 17/16 margin) from a rounding boundary k+0.5 the library's byte is
// round(mean).  The few cells (~0.2 %) that are closer are listed and recomputed in the library's exact float32
// order by area_fix_kernel (k1_generic.cu), off this kernel's critical path.
#include "avd_internal.h"
#include <stdlib.h>

namespace {

#ifndef K1_NRS
#define K1_NRS 6
#endif
#ifndef K1_NOTHUMB
#define K1_NOTHUMB 0
#endif
#ifndef K1_MINB
#define K1_MINB 4
#endif
constexpr int NRS = K1_NRS;           // single-row TMA stages per CTA (multiple of 3: stage index is static per unrolled step)
static_assert(NRS % 3 == 0 && NRS <= 6, "stages come in groups of 3: the position inside a group is tied to the 3-way register rotation");
constexpr int NGRP = NRS / 3;
constexpr int MAXQ = 132;             // rows a band can load (H/32 + 3, H <= 4096)
constexpr int MAX_TROWS = 16;         // thumbnail rows a band can own (staged in shared memory, pair form)
constexpr int SLOT = 48;              // bytes of a thread's private gray slot (48: conflict-free 128-bit stores)

// shared-memory header (bytes)
constexpr int OFF_FULL = 0, OFF_CNT = 64, OFF_MISC = 192, OFF_CELL = 256, OFF_INFO = 512, HDR = 1040;
static_assert(OFF_INFO + 4 * MAXQ <= HDR, "header layout");

struct K1Params {
    const uint8_t *bgr;
    size_t pitch, fstride;
    int H, W, row_bytes;
    const ShapeTables *tab;
    avd_sample_rec *rec;
    uint8_t *thumbs, *area32;
    uint32_t *fixlist;                // [0] = count (zeroed by k1_rec_init), then frame * 1024 + cell
    int frame0;
    int nframes;                      // frames of this launch (PAIR form: the last CTA row may hold a single frame)
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
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity)
{
    // try_wait suspends the thread in hardware for up to the hinted time, so this loop turns over rarely;
    // a lost refill must fail loudly instead of hanging the GPU
#pragma unroll 1
    for (int it = 0; it < (1 << 22); it++) {
        uint32_t ok;
        asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2, %3;\nselp.u32 %0, 1, 0, P1;\n}"
                     : "=r"(ok) : "r"(bar), "r"(parity), "r"(20000u) : "memory");
        if (ok) return;
    }
    __trap();
}
// Non-blocking probe (test_wait, not try_wait: a failing try_wait suspends the warp), issued early so that its latency
// hides behind the previous row's arithmetic; the branch on the result comes a row later.
__device__ __forceinline__ uint32_t mbar_try(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n.reg .pred P1;\nmbarrier.test_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    return ok;
}
// lane 0 of a warp counts the warp off a stage; returns 1 in the lane that arrived last (pending count before its arrival)
// `dep` is a value computed from the last bytes loaded from the stage: the arrival cannot be issued before they landed
__device__ __forceinline__ uint32_t stage_arrive(uint32_t bar, uint32_t lane0, uint32_t dep)
{
    uint32_t pend = 0;
    asm volatile("{\n.reg .pred p;\n.reg .b64 st;\nsetp.ne.u32 p, %2, 0;\n@p mbarrier.arrive.shared::cta.b64 st, [%1];\n@p mbarrier.pending_count.b64 %0, st;\n}"
                 : "+r"(pend) : "r"(bar), "r"(lane0), "r"(dep) : "memory");
    return pend;
}

// 2*(3735,19235,9798): the doubled 15-bit weights; acc = 2*(...)+32768 so gray = acc >> 16 = byte 2
constexpr uint32_t WB = 7470u, WG = 38470u, WR = 19596u;
constexpr uint32_t C_BG = WB | (WG << 16), C_R0 = WR, C_0B = WB << 16, C_GR = WG | (WR << 16);
constexpr uint32_t RND = 32768u;
constexpr uint32_t KBIAS = 1020u | (1020u << 16);

// IDP takes its weights from registers: keep them resident (loaded once through an opaque move) instead of letting
// the compiler rematerialise five constants on every row
struct GrayK {
    uint32_t bg, r0, ob, gr, rnd;
    __device__ __forceinline__ void load()
    {
        asm volatile("mov.u32 %0, %1;" : "=r"(bg) : "n"(C_BG)); asm volatile("mov.u32 %0, %1;" : "=r"(r0) : "n"(C_R0));
        asm volatile("mov.u32 %0, %1;" : "=r"(ob) : "n"(C_0B)); asm volatile("mov.u32 %0, %1;" : "=r"(gr) : "n"(C_GR));
        asm volatile("mov.u32 %0, %1;" : "=r"(rnd) : "n"(RND));
    }
};

// gray accumulators of 4 consecutive pixels held in 3 consecutive words
__device__ __forceinline__ void gray4(const GrayK &K, uint32_t w0, uint32_t w1, uint32_t w2, uint32_t &a0, uint32_t &a1, uint32_t &a2, uint32_t &a3)
{
    a0 = __dp2a_hi(K.r0, w0, __dp2a_lo(K.bg, w0, K.rnd));      // B=w0.0 G=w0.1 R=w0.2
    a1 = __dp2a_lo(K.gr, w1, __dp2a_hi(K.ob, w0, K.rnd));      // B=w0.3 G=w1.0 R=w1.1
    a2 = __dp2a_lo(K.r0, w2, __dp2a_hi(K.bg, w1, K.rnd));      // B=w1.2 G=w1.3 R=w2.0
    a3 = __dp2a_hi(K.gr, w2, __dp2a_lo(K.ob, w2, K.rnd));      // B=w2.1 G=w2.2 R=w2.3
}

// one gray row of a 16-pixel strip: E[g] = (p[4g], p[4g+2]), O[g] = (p[4g+1], p[4g+3]) in 16-bit lanes,
// LR = (pixel left of the strip, pixel right of the strip), REFLECT_101 at the frame edges
struct RowSet { uint32_t E[4], O[4], LR; };

struct Acc {
    uint32_t ll, hl, hh;        // byte-split sums of (L+1020)^2: lo*lo, hi*lo, hi*hi
    uint32_t aLR, aE0, aO3;     // packed column sums for the telescoped horizontal part of sum L
    uint32_t cs[4];             // packed pair-column sums of the owned rows (INTER_AREA)
};

__device__ __forceinline__ uint32_t lanes_sum(uint32_t v) { return (v & 0xffffu) + (v >> 16); }
__device__ __forceinline__ uint32_t row_sum(const RowSet &R)
{
    return lanes_sum(R.E[0] + R.E[1] + R.E[2] + R.E[3] + R.O[0] + R.O[1] + R.O[2] + R.O[3]);
}

struct Band {
    uint32_t sbase;                       // shared address of the header
    uint32_t stage;                       // shared address of stage 0
    uint32_t slot;                        // shared address of this thread's private gray slot
    uint32_t sp0, rb1, rb2;               // this thread's strip in stage 0; one and two rows in bytes
    int row_bytes, nseq;                  // bytes of one frame row (one bulk copy)
    int H, own0;
    const uint8_t *frame, *frame2;        // PAIR form: the second frame's rows land right behind the first one's in a stage
    size_t pitch;
    uint64_t pol_stream, pol_keep;
    GrayK K;
    int todd;                             // pair-sum thumbnail: tap pairs start at odd offsets
};

// thumbnail columns owned by this thread (left tap inside its strip): shared address of the left tap byte in the
// thread's slot, the two horizontal weights, the vertically weighted first tap row carried to the second
template <int JT> struct Thumb {
    uint32_t addr[JT], a0[JT], a1[JT], u0[JT];
    int ntx;
    uint8_t *dst;                          // &thumbs[frame][0][first column of this thread] (general form)
    uint32_t th;                           // pair form: shared address of this thread's first column in the band's staged rows
};

// ---- thumbnail, pair-sum form (W a multiple of 640: every horizontal tap pair has weights 1024/1024, so the
// horizontal pass is 1024 * (p[x0] + p[x0+1]) and those pair sums already exist in 16-bit lanes: E+O holds the pairs
// starting at even offsets, O+orr the pairs starting at odd offsets).  A thumbnail row is produced when its bottom tap
// row is the centre row C: its top tap row is then still in registers as U.  Both rows' pair sums go to the thread's
// private slot and each thumbnail column of the thread reads its lanes back at its own (per-thread constant) offset;
// the loads are issued before the Laplacian arithmetic and consumed after it.
__device__ __forceinline__ void pair_sums(const RowSet &R, const uint32_t *orr, int todd, uint32_t *X)
{
#pragma unroll
    for (int g = 0; g < 4; g++) X[g] = todd ? R.O[g] + orr[g] : R.E[g] + R.O[g];
}
__device__ __forceinline__ void right_shifted(const RowSet &R, uint32_t *orr)      // (p2,p4) of each group
{
#pragma unroll
    for (int g = 0; g < 4; g++) orr[g] = g == 3 ? __byte_perm(R.E[3], R.LR, 0x7632) : __byte_perm(R.E[g], R.E[g + 1], 0x5432);
}
template <int JT>
__device__ __forceinline__ void thumb_pairs_begin(const Band &B, const Thumb<JT> &T, const uint32_t *XU, const uint32_t *XC, uint32_t *s0, uint32_t *s1)
{
    asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(B.slot), "r"(XU[0]), "r"(XU[1]), "r"(XU[2]), "r"(XU[3]) : "memory");
    asm volatile("st.shared.v4.u32 [%0+16], {%1,%2,%3,%4};" ::"r"(B.slot), "r"(XC[0]), "r"(XC[1]), "r"(XC[2]), "r"(XC[3]) : "memory");
#pragma unroll
    for (int k = 0; k < JT; k++) {
        asm volatile("ld.shared.u16 %0, [%1];" : "=r"(s0[k]) : "r"(T.addr[k]) : "memory");
        asm volatile("ld.shared.u16 %0, [%1+16];" : "=r"(s1[k]) : "r"(T.addr[k]) : "memory");
    }
}
// W = 1280 (TM 4) and W = 2560 (TM 8): the tap pairs sit at the same offsets in every strip (4g+1 resp. 3 and 11), so
// the lanes come straight out of the registers, no trip through the slot
template <int JT, int TM>
__device__ __forceinline__ void thumb_pairs_static(const uint32_t *XU, const uint32_t *XC, uint32_t *s0, uint32_t *s1)
{
#pragma unroll
    for (int k = 0; k < JT; k++) {
        if (TM == 4) { s0[k] = XU[k & 3] & 0xffffu; s1[k] = XC[k & 3] & 0xffffu; }            // pair (4k+1, 4k+2): low lane of group k
        else { s0[k] = XU[(2 * k) & 3] >> 16; s1[k] = XC[(2 * k) & 3] >> 16; }               // pair (8k+3, 8k+4): high lane of group 2k
    }
}

template <int JT>
__device__ __forceinline__ void thumb_pairs_end(const Thumb<JT> &T, uint32_t info, const uint32_t *s0, const uint32_t *s1)
{
    const uint32_t b1 = (info >> 2) & 0xfffu, b0 = 2048u - b1;
    const uint32_t row = T.th + (info >> 16) * AVD_T;          // the band's thumbnail rows are staged in shared memory
#pragma unroll
    for (int k = 0; k < JT; k++) {        // (b * ((1024 * sum) >> 4)) >> 16  ==  (b * sum) >> 10
        const uint32_t v = (((b0 * s0[k]) >> 10) + ((b1 * s1[k]) >> 10) + 2u) >> 2;
        if (k < T.ntx) asm volatile("st.shared.u8 [%0], %1;" ::"r"(row + k), "r"(v) : "memory");
    }
}

// Laplacian statistics of centre row C (rows above / below: U, D), its INTER_AREA column sums and, in the pair-sum
// form of the thumbnail, its thumbnail columns (infoC = thumbnail role of the centre row)
template <int JT, int TM>
__device__ __forceinline__ void lap_step(const Band &B, const RowSet &U, const RowSet &C, const RowSet &D, Acc &A, Thumb<JT> &T, uint32_t infoC)
{
    uint32_t el[4], orr[4], Pp[4];
#pragma unroll
    for (int g = 0; g < 4; g++) {
        // left neighbours of (p0,p2) are (pL,p1); right neighbours of (p1,p3) are (p2,pR)
        el[g] = g == 0 ? __byte_perm(C.LR, C.O[0], 0x5410) : __byte_perm(C.O[g - 1], C.O[g], 0x5432);
        orr[g] = g == 3 ? __byte_perm(C.E[3], C.LR, 0x7632) : __byte_perm(C.E[g], C.E[g + 1], 0x5432);
        Pp[g] = C.E[g] + C.O[g];
    }
    uint32_t s0[JT], s1[JT];
#if !K1_NOTHUMB
    if ((TM != 0) && (infoC & 2u)) {               // C is the bottom tap row of a thumbnail row, U its top tap row
        uint32_t orrU[4], XU[4], XC[4];
        right_shifted(U, orrU);
        pair_sums(U, orrU, B.todd, XU);
        pair_sums(C, orr, B.todd, XC);
        if constexpr (TM >= 4) thumb_pairs_static<JT, TM>(XU, XC, s0, s1);
        else thumb_pairs_begin<JT>(B, T, XU, XC, s0, s1);
    }
#endif
#pragma unroll
    for (int g = 0; g < 4; g++) {
        const uint32_t me = KBIAS - 4u * C.E[g], mo = KBIAS - 4u * C.O[g];
        const uint32_t le = (el[g] + C.O[g] + me) + U.E[g] + D.E[g];      // L+1020 of (p0,p2), each in [0,2040]
        const uint32_t lo = (C.E[g] + orr[g] + mo) + U.O[g] + D.O[g];     // L+1020 of (p1,p3)
        const uint32_t pl = __byte_perm(le, lo, 0x6420), ph = __byte_perm(le, lo, 0x7531);
        A.ll = __dp4a(pl, pl, A.ll);
        A.hl = __dp4a(ph, pl, A.hl);
        A.hh = __dp4a(ph, ph, A.hh);
        A.cs[g] += Pp[g];
    }
    A.aLR += C.LR; A.aE0 += C.E[0]; A.aO3 += C.O[3];
#if !K1_NOTHUMB
    if ((TM != 0) && (infoC & 2u)) thumb_pairs_end<JT>(T, infoC, s0, s1);
#endif
}

// sequence row q of the band = source row own0-1+q, reflected at the frame edges (BORDER_REFLECT_101)
__device__ __forceinline__ void issue_row(const Band &B, int q, uint32_t stage_off, uint32_t bar_off)
{
    int r = B.own0 - 1 + q;
    r = r < 0 ? 1 : (r >= B.H ? B.H - 2 : r);
    const uint32_t bar = B.sbase + OFF_FULL + bar_off;
    mbar_expect_tx(bar, (uint32_t)B.row_bytes * (B.frame2 ? 2u : 1u));
    // the two first and two last rows of a band are also read by the neighbouring bands: do not mark them evict-first
    const bool shared_row = q < 2 || q >= B.nseq - 2;
    tma_load_1d(B.stage + stage_off, B.frame + (size_t)r * B.pitch, (uint32_t)B.row_bytes, bar, shared_row ? B.pol_keep : B.pol_stream);
    if (B.frame2)
        tma_load_1d(B.stage + stage_off + (uint32_t)B.row_bytes, B.frame2 + (size_t)r * B.pitch, (uint32_t)B.row_bytes, bar, shared_row ? B.pol_keep : B.pol_stream);
}

// Stage s holds sequence row q: convert this thread's strip to gray (row set R); on tap rows of the thumbnail also
// advance this thread's thumbnail columns.  Returns 1 in lane 0 of the warp that was last to leave the stage
// (passed on to release_finish).
template <int POS, int JT, int TM>
__device__ __forceinline__ uint32_t load_row(const Band &B, int q, uint32_t goff, uint32_t gbar, uint32_t parity, uint32_t &ok, uint32_t nbar, uint32_t npar,
                                             int tid, bool first, bool last, RowSet &R, Thumb<JT> &T)
{
    // goff / gbar: byte offsets of the stage group (0 or 3 rows; 0 or 3 barriers), toggled by the caller: no per-row multiplies
    if (!ok) mbar_wait_slow(B.sbase + OFF_FULL + 8 * POS + gbar, parity);
    const uint32_t sp = B.sp0 + goff + (POS == 0 ? 0u : (POS == 1 ? B.rb1 : B.rb2));
    uint4 q0, q1, q2;
    uint32_t wl, wr;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q0.x), "=r"(q0.y), "=r"(q0.z), "=r"(q0.w) : "r"(sp));
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(q1.x), "=r"(q1.y), "=r"(q1.z), "=r"(q1.w) : "r"(sp));
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+32];" : "=r"(q2.x), "=r"(q2.y), "=r"(q2.z), "=r"(q2.w) : "r"(sp));
    asm volatile("ld.shared.u32 %0, [%1+-4];" : "=r"(wl) : "r"(sp));
    asm volatile("ld.shared.u32 %0, [%1+48];" : "=r"(wr) : "r"(sp));
    uint32_t a[16];
    const GrayK &K = B.K;
    gray4(K, q0.x, q0.y, q0.z, a[0], a[1], a[2], a[3]);
    gray4(K, q0.w, q1.x, q1.y, a[4], a[5], a[6], a[7]);
    gray4(K, q1.z, q1.w, q2.x, a[8], a[9], a[10], a[11]);
    gray4(K, q2.y, q2.z, q2.w, a[12], a[13], a[14], a[15]);
    // neighbour pixels x-1 and x+16 (REFLECT_101 at the frame edges: x=-1 -> 1, x=W -> W-2)
    uint32_t aL = __dp2a_hi(K.gr, wl, __dp2a_lo(K.ob, wl, K.rnd));
    uint32_t aR = __dp2a_hi(K.r0, wr, __dp2a_lo(K.bg, wr, K.rnd));
    // every lane of this warp has its bytes in registers: count the warp off the stage (result used much later)
    __syncwarp();
    const uint32_t ticket = stage_arrive(B.sbase + OFF_CNT + 8 * POS + gbar, (tid & 31) == 0 ? 1u : 0u, aL ^ aR);
    ok = mbar_try(nbar, npar);               // next row's stage: checked at the top of the next step
    aL = first ? a[1] : aL;
    aR = last ? a[14] : aR;
#pragma unroll
    for (int g = 0; g < 4; g++) {
        R.E[g] = __byte_perm(a[4 * g], a[4 * g + 2], 0x7632);
        R.O[g] = __byte_perm(a[4 * g + 1], a[4 * g + 3], 0x7632);
    }
    R.LR = __byte_perm(aL, aR, 0x7632);
    // general form of the thumbnail (any tap weights): handled when the row is loaded
    // info: bit0 tap, bit1 second row of its pair, bits 2-13 vertical weight, bits 16-24 thumbnail row
    uint32_t info = 0;
    if (!(TM != 0)) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(info) : "r"(B.sbase + OFF_INFO + 4 * q));
#if K1_NOTHUMB
    info = 0;                                 // timing experiment only
#endif
    if (!(TM != 0) && (info & 1u)) {
        uint4 gw;
        gw.x = __byte_perm(R.E[0], R.O[0], 0x6240); gw.y = __byte_perm(R.E[1], R.O[1], 0x6240);
        gw.z = __byte_perm(R.E[2], R.O[2], 0x6240); gw.w = __byte_perm(R.E[3], R.O[3], 0x6240);
        asm volatile("st.shared.v4.u32 [%0], {%1,%2,%3,%4};" ::"r"(B.slot), "r"(gw.x), "r"(gw.y), "r"(gw.z), "r"(gw.w) : "memory");
        asm volatile("st.shared.u8 [%0+16], %1;" ::"r"(B.slot), "r"(aR >> 16) : "memory");
        const uint32_t bw = (info >> 2) & 0xfffu;
        uint32_t u[JT];
#pragma unroll
        for (int k = 0; k < JT; k++) {
            uint32_t g0, g1;
            asm volatile("ld.shared.u8 %0, [%1];" : "=r"(g0) : "r"(T.addr[k]) : "memory");
            asm volatile("ld.shared.u8 %0, [%1+1];" : "=r"(g1) : "r"(T.addr[k]) : "memory");
            const uint32_t Sx = g0 * T.a0[k] + g1 * T.a1[k];
            u[k] = (bw * (Sx >> 4)) >> 16;
        }
        if (info & 2u) {
            uint8_t *dst = T.dst + (size_t)(info >> 16) * AVD_T;
#pragma unroll
            for (int k = 0; k < JT; k++)
                if (k < T.ntx) dst[k] = (uint8_t)((T.u0[k] + u[k] + 2u) >> 2);
        } else {
#pragma unroll
            for (int k = 0; k < JT; k++) T.u0[k] = u[k];
        }
    }
    return ticket;
}

// the warp that arrived last at stage S refills it with sequence row q + NRS
template <int POS>
__device__ __forceinline__ void release_finish(const Band &B, int q, uint32_t goff, uint32_t gbar, uint32_t ticket)
{
    if (ticket == 1u && q + NRS < B.nseq) issue_row(B, q + NRS, goff + POS * B.rb1, gbar + 8 * POS);
}

__device__ __forceinline__ uint8_t sat_rne(float v)
{
    int r = __float2int_rn(v);
    return (uint8_t)min(max(r, 0), 255);
}

// sum of the pair lanes selected by the byte masks m[g] (byte 0: low lane, byte 1: high lane)
__device__ __forceinline__ uint32_t masked_sum(const uint32_t *P, const uint32_t *m)
{
    uint32_t s = 0;
#pragma unroll
    for (int g = 0; g < 4; g++) s = __dp2a_lo(P[g], m[g], s);
    return s;
}

// PAIR: a CTA takes band b of TWO consecutive frames side by side (thread tid < W/16: frame 2y, the others: frame 2y+1).  At
// W = 1280 a frame has 80 strips = 2.5 warps; alone, a sixth of the CTA's lanes idles and the band's fixed cost (barrier set-up,
// first fill of the ring, epilogue) is spread over few bytes; two frames are 160 strips = 5 full warps.
template <int MAXT, int MINB, int JT, int TM, bool PAIR = false>
__global__ void __launch_bounds__(MAXT, MINB) k1_fused_kernel(K1Params P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int nthr = blockDim.x, nw = nthr >> 5;
    const int b = blockIdx.x;
    const ShapeTables *__restrict__ tab = P.tab;
    const int H = P.H, W = P.W;
    const int S = W >> 4;                                     // strips of a frame
    const int half = PAIR && tid >= S ? 1 : 0, lt = tid - half * S;       // which of the CTA's frames, strip inside it
    const int f = PAIR ? 2 * (int)blockIdx.y + half : (int)blockIdx.y;
    const bool have2 = PAIR && 2 * (int)blockIdx.y + 1 < P.nframes;
    const uint32_t stride = (PAIR ? 2u : 1u) * (uint32_t)P.row_bytes;      // bytes of a stage
    const int own0 = tab->row_owner_begin[b], own1 = tab->row_owner_begin[b + 1];
    const int n = own1 - own0, nseq = n + 2;
    uint32_t sbase;                         // opaque, or the compiler recomputes the shared window address at every use
    asm volatile("mov.u32 %0, %1;" : "=r"(sbase) : "r"(smem_u32(smem)));
    const uint8_t *frame0 = P.bgr + (size_t)(PAIR ? 2 * (int)blockIdx.y : (int)blockIdx.y) * P.fstride;

    Band B;
    B.sbase = sbase; B.stage = sbase + HDR; B.slot = sbase + HDR + NRS * stride + 64 + SLOT * tid;
    B.row_bytes = P.row_bytes; B.nseq = nseq;
    B.sp0 = sbase + HDR + 48 * tid; B.rb1 = stride; B.rb2 = 2u * stride;
    B.K.load();
    B.todd = tab->lx[0].i0 & 1;
    B.H = H; B.own0 = own0; B.frame = frame0; B.frame2 = have2 ? frame0 + P.fstride : nullptr; B.pitch = P.pitch;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(B.pol_stream));
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(B.pol_keep));

    if (tid == 0) {
        for (int s = 0; s < NRS; s++) {
            mbar_init(sbase + OFF_FULL + 8 * s, 1);
            mbar_init(sbase + OFF_CNT + 8 * s, nw);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int q = 0; q < NRS; q++) issue_row(B, q, q * stride, 8 * q);
    }
    if (tid < (PAIR ? 64 : 32)) ((int *)(smem + OFF_CELL))[tid] = 0;
    if (tid == 32) ((uint32_t *)(smem + OFF_MISC))[1] = 0;          // warps that finished the band
    // thumbnail role of the band's rows: the top tap row must be owned by this band, the bottom tap is the row below it
    for (int q = tid; q < nseq; q += nthr) {
        const int r = own0 - 1 + q;
        uint32_t v = (r >= 0 && r < H) ? tab->rowinfo[r] : 0u;
        if (q == 0 || (q == 1 && (v & 2u)) || (q == nseq - 1 && !(v & 2u))) v = 0;
        ((uint32_t *)(smem + OFF_INFO))[q] = v;
    }
    // this thread's thumbnail columns
    Thumb<JT> T;
    {
        const int tx0 = tab->strip_tx0[lt];
        T.ntx = (16 * lt < W && (half == 0 || have2)) ? tab->strip_ntx[lt] : 0;
        T.dst = P.thumbs + (size_t)f * AVD_T * AVD_T + tx0;
        T.th = B.slot - SLOT * tid + SLOT * nthr + 64 + tx0 - (uint32_t)(tab->thumb_row_begin[b] * AVD_T)     // row index comes as an absolute thumbnail row
               + (uint32_t)(half * MAX_TROWS * AVD_T);
#pragma unroll
        for (int k = 0; k < JT; k++) {
            const LinTap t = tab->lx[min(tx0 + k, AVD_T - 1)];
            const bool ok = k < T.ntx;
            const int rel = t.i0 - 16 * lt;                             // left tap, relative to the strip
            T.addr[k] = B.slot + (ok ? (uint32_t)((TM != 0) ? 2 * (rel >> 1) : rel) : 0u);   // pair form: 16-bit lane rel/2 (even) or (rel-1)/2 (odd)
            T.a0[k] = (uint32_t)t.a0; T.a1[k] = (uint32_t)t.a1; T.u0[k] = 0;
        }
    }
    __syncthreads();

    const bool active = 16 * lt < W && (half == 0 || have2), first = lt == 0, last = 16 * lt + 16 == W;
    RowSet R0, R1, R2;
    Acc A;
    A.ll = A.hl = A.hh = 0; A.aLR = A.aE0 = A.aO3 = 0; A.cs[0] = A.cs[1] = A.cs[2] = A.cs[3] = 0;

    // INTER_AREA: this strip's pair lanes split between cell cA (pairs < kA) and cell cA+1
    const int sx = W >> 5;
    const int cA = (16 * lt) / sx;
    const int kA = min(8, ((cA + 1) * sx - 16 * lt) >> 1);
    uint32_t mA[4];
#pragma unroll
    for (int g = 0; g < 4; g++) mA[g] = (2 * g < kA ? 1u : 0u) | (2 * g + 1 < kA ? 0x100u : 0u);
    const uint32_t mAll[4] = {0x0101u, 0x0101u, 0x0101u, 0x0101u};
    const int wh = 32 * own0 - b * H;                     // weight of the head row (sequence row 0), 1/32 units
    const int wt = (b + 1) * H - 32 * (own1 - 1);         // weight of the last owned row, 1/32 units (32 = full)

    // stage group state: byte offset of the group's first stage / first barrier (toggled, never multiplied), phase parity
    const uint32_t g3 = NGRP > 1 ? 3u * stride : 0u, gb3 = NGRP > 1 ? 24u : 0u;
    const uint32_t fbar = sbase + OFF_FULL;
    uint32_t par = 0, goff = 0, gbar = 0;
    uint32_t ok = mbar_try(fbar, 0u);
    uint32_t t0 = load_row<0, JT, TM>(B, 0, 0u, 0u, 0u, ok, fbar + 8, 0u, tid, first, last, R0, T);
    release_finish<0>(B, 0, 0u, 0u, t0);
    uint32_t t1 = load_row<1, JT, TM>(B, 1, 0u, 0u, 0u, ok, fbar + 16, 0u, tid, first, last, R1, T);
    release_finish<1>(B, 1, 0u, 0u, t1);
    int tv = (int)row_sum(R0) - (int)row_sum(R1);         // vertical part of sum L telescopes to 4 rows per band
    int mcA = 0, mcB = 0;                                  // cell sums in 1/32 units
    if (wh > 0) {
        uint32_t Pp[4];
#pragma unroll
        for (int g = 0; g < 4; g++) Pp[g] = R0.E[g] + R0.O[g];
        const uint32_t sa = masked_sum(Pp, mA), st = masked_sum(Pp, mAll);
        mcA = wh * (int)sa; mcB = wh * (int)(st - sa);
    }
    // at the last sequence row: C is the last owned row, D the bottom halo
    // thumbnail role of sequence row q (pair-sum form: looked up when the row is the centre row)
    auto row_info = [&](int qq) {
        uint32_t v = 0;
        if ((TM != 0)) asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(sbase + OFF_INFO + 4 * qq));
        return v;
    };
    auto band_end = [&](const RowSet &C, const RowSet &D) {
#if !K1_NOTHUMB
        const uint32_t infoD = row_info(nseq - 1);       // the bottom halo row is never a centre row
        if ((TM != 0) && (infoD & 2u)) {
            uint32_t orrC[4], orrD[4], XU[4], XC[4], s0[JT], s1[JT];
            right_shifted(C, orrC); right_shifted(D, orrD);
            pair_sums(C, orrC, B.todd, XU);
            pair_sums(D, orrD, B.todd, XC);
            if constexpr (TM >= 4) thumb_pairs_static<JT, TM>(XU, XC, s0, s1);
            else thumb_pairs_begin<JT>(B, T, XU, XC, s0, s1);
            thumb_pairs_end<JT>(T, infoD, s0, s1);
        }
#endif
        tv += (int)row_sum(D) - (int)row_sum(C);
        uint32_t Pp[4];
#pragma unroll
        for (int g = 0; g < 4; g++) Pp[g] = C.E[g] + C.O[g];
        const uint32_t sa = masked_sum(Pp, mA), st = masked_sum(Pp, mAll);
        mcA -= (32 - wt) * (int)sa; mcB -= (32 - wt) * (int)(st - sa);
    };

    int q = 2;
    while (true) {
        // the group after this one: the other group (NRS = 6) or the same one (NRS = 3); the parity flips when group 0 comes round again
        const uint32_t ngoff = goff ^ g3, ngbar = gbar ^ gb3;
        const uint32_t np = ngbar == 0 ? par ^ 1u : par;
        uint32_t tk = load_row<2, JT, TM>(B, q, goff, gbar, par, ok, fbar + ngbar, np, tid, first, last, R2, T);
        lap_step<JT, TM>(B, R0, R1, R2, A, T, row_info(q - 1));
        release_finish<2>(B, q, goff, gbar, tk);
        if (q == nseq - 1) { band_end(R1, R2); break; }
        q++; goff = ngoff; gbar = ngbar; par = np;
        tk = load_row<0, JT, TM>(B, q, goff, gbar, par, ok, fbar + gbar + 8, par, tid, first, last, R0, T);
        lap_step<JT, TM>(B, R1, R2, R0, A, T, row_info(q - 1));
        release_finish<0>(B, q, goff, gbar, tk);
        if (q == nseq - 1) { band_end(R2, R0); break; }
        q++;
        tk = load_row<1, JT, TM>(B, q, goff, gbar, par, ok, fbar + gbar + 16, par, tid, first, last, R1, T);
        lap_step<JT, TM>(B, R2, R0, R1, A, T, row_info(q - 1));
        release_finish<1>(B, q, goff, gbar, tk);
        if (q == nseq - 1) { band_end(R0, R1); break; }
        q++;
    }

    // ---- per-thread totals; no block barrier: every warp publishes its own sums, the last one finishes the cells ----
    {
        const uint32_t sa = masked_sum(A.cs, mA), st = masked_sum(A.cs, mAll);
        mcA += 32 * (int)sa; mcB += 32 * (int)(st - sa);
    }
    long long sumL = (long long)tv + (long long)lanes_sum(A.aLR) - (long long)(A.aE0 & 0xffffu) - (long long)(A.aO3 >> 16);
    // sum (L+1020)^2 -> sum L^2 over this thread's n x 16 pixels
    long long sumL2 = (long long)((unsigned long long)A.ll + 512ull * A.hl + 65536ull * A.hh) - 2040ll * sumL - 1040400ll * (16ll * n);
    int *cellM = (int *)(smem + OFF_CELL);
    if (active) {
        atomicAdd(&cellM[32 * half + cA], mcA);
        if (kA < 8) atomicAdd(&cellM[32 * half + cA + 1], mcB);
    } else {
        sumL = 0; sumL2 = 0;
    }
    if (PAIR) {
        // a warp may hold strips of both frames: one reduction per frame
        long long a0 = half ? 0 : sumL, a1 = half ? 0 : sumL2, b0 = half ? sumL : 0, b1 = half ? sumL2 : 0;
        for (int o = 16; o; o >>= 1) {
            a0 += __shfl_down_sync(0xffffffffu, a0, o); a1 += __shfl_down_sync(0xffffffffu, a1, o);
            b0 += __shfl_down_sync(0xffffffffu, b0, o); b1 += __shfl_down_sync(0xffffffffu, b1, o);
        }
        if (lane == 0) {
            const int fa = 2 * (int)blockIdx.y;
            if (a0 | a1) {
                atomicAdd((unsigned long long *)&P.rec[fa].lap_sum, (unsigned long long)a0);
                atomicAdd((unsigned long long *)&P.rec[fa].lap_sumsq, (unsigned long long)a1);
            }
            if (have2 && (b0 | b1)) {
                atomicAdd((unsigned long long *)&P.rec[fa + 1].lap_sum, (unsigned long long)b0);
                atomicAdd((unsigned long long *)&P.rec[fa + 1].lap_sumsq, (unsigned long long)b1);
            }
        }
    } else {
        for (int o = 16; o; o >>= 1) { sumL += __shfl_down_sync(0xffffffffu, sumL, o); sumL2 += __shfl_down_sync(0xffffffffu, sumL2, o); }
        if (lane == 0) {
            atomicAdd((unsigned long long *)&P.rec[f].lap_sum, (unsigned long long)sumL);
            atomicAdd((unsigned long long *)&P.rec[f].lap_sumsq, (unsigned long long)sumL2);
        }
    }
    uint32_t is_last = 0;
    __syncwarp();
    if (lane == 0) {
        uint32_t old;
        asm volatile("atom.acq_rel.cta.shared::cta.add.u32 %0, [%1], 1;" : "=r"(old) : "r"(sbase + OFF_MISC + 4) : "memory");
        is_last = old == (uint32_t)(nw - 1);
    }
    is_last = __shfl_sync(0xffffffffu, is_last, 0);
    if (!is_last) return;

    // ---- this warp is the last one of the band: per frame of the CTA, the staged thumbnail rows (contiguous in the thumbnail)
    // go out, then INTER_AREA destination row b, column `lane` ----
    for (int h = 0; h < (have2 ? 2 : 1); h++) {
        const int fh = PAIR ? 2 * (int)blockIdx.y + h : f;
        if ((TM != 0)) {
            const int ty0 = tab->thumb_row_begin[b], nty = tab->thumb_row_begin[b + 1] - ty0;
            const uint4 *src = (const uint4 *)(smem + HDR + NRS * stride + 64 + SLOT * nthr + 64 + h * MAX_TROWS * AVD_T);
            uint4 *dst = (uint4 *)(P.thumbs + ((size_t)fh * AVD_T + ty0) * AVD_T);
            for (int i = lane; i < nty * (AVD_T / 16); i += 32) dst[i] = src[i];
        }
        const int M = cellM[32 * h + lane];
        uint32_t out;
        bool close = false;
        if (tab->area_int) {
            out = sat_rne(__fmul_rn((float)(M >> 5), tab->inv_area));
        } else {
            // exact mean v = M / (H*sx).  close <=> |frac(v) - 0.5| <= 17/16 (sx+n+5) 2^-24 v, all in integers (times 2 H sx)
            const uint32_t Dn = (uint32_t)H * (uint32_t)sx;
            const uint32_t k = (uint32_t)M / Dn, r = (uint32_t)M - k * Dn;
            const int d2 = (int)(2u * r) - (int)Dn;
            const uint32_t bound = (uint32_t)(((unsigned long long)(uint32_t)M * (uint32_t)(17 * (sx + n + 5))) >> 27) + 2u;
            close = (uint32_t)abs(d2) <= bound;
            out = k + (d2 >= 0 ? 1u : 0u);
        }
        P.area32[(size_t)fh * 1024 + b * 32 + lane] = (uint8_t)out;
        // uncertified cells go to a list; k1_hash_finalize recomputes them in the library's float32 order
        const uint32_t mask = __ballot_sync(0xffffffffu, close);
        if (mask) {
            uint32_t base = 0;
            if (lane == 0) base = atomicAdd(P.fixlist, (uint32_t)__popc(mask));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (close) P.fixlist[1 + base + __popc(mask & ((1u << lane) - 1u))] = (uint32_t)(P.frame0 + fh) * 1024u + (uint32_t)(b * 32 + lane);
        }
    }
}

}  // namespace

// frames a CTA takes side by side: two where that fills the warps (W = 1280: 80 strips per frame)
static int k1_frames_per_cta(int W)
{
    static const bool off = getenv("AVD_K1_NOPAIR") && atoi(getenv("AVD_K1_NOPAIR"));
    return (W == 1280 && !off) ? 2 : 1;
}
static size_t k1_smem_bytes(int W)
{
    const int fpc = k1_frames_per_cta(W);
    const int nthr = (fpc * (W / 16) + 31) / 32 * 32;
    return HDR + (size_t)NRS * 3 * W * fpc + 64 + (size_t)SLOT * nthr + 64 + (size_t)MAX_TROWS * AVD_T * fpc;
}

bool k1_fused_supported(int H, int W, size_t pitch, size_t fstride, const void *base, const ShapeTables *htab)
{
    if (W % 64 != 0 || W < 1280 || W > 4096 || H < 640 || H > 4096 || htab->thumb_2x2) return false;
    if (pitch % 16 != 0 || fstride % 16 != 0 || ((uintptr_t)base % 16) != 0) return false;
    if (k1_smem_bytes(W) > 227 * 1024 || htab->strip_ntx_max > 4) return false;
    for (int b = 0; b < AVD_D; b++)
        if (htab->thumb_row_begin[b + 1] - htab->thumb_row_begin[b] > MAX_TROWS) return false;
    // plain two-tap decimation: taps (i, i+1), every tap row used by exactly one thumbnail row
    for (int i = 0; i < AVD_T; i++) {
        if (htab->lx[i].i1 != htab->lx[i].i0 + 1 || htab->ly[i].i1 != htab->ly[i].i0 + 1) return false;
        if (i > 0 && htab->ly[i].i0 < htab->ly[i - 1].i0 + 2) return false;
    }
    return true;
}

template <int MAXT, int MINB, int JT, int TM, bool PAIR = false>
static int k1_launch(avd_ctx *ctx, const K1Params &P, int n, size_t smem, int threads, cudaStream_t st)
{
    static size_t smem_set[64] = {};                 // per device: function attributes belong to the device's context
    size_t &set = smem_set[ctx->device & 63];
    if (smem > set) {
        AVD_CUDA(ctx, cudaFuncSetAttribute(k1_fused_kernel<MAXT, MINB, JT, TM, PAIR>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        set = smem;
    }
    for (int f0 = 0; f0 < n; f0 += 65534) {                          // grid.y limit (even, so that pairs never straddle launches)
        int nb = n - f0 < 65534 ? n - f0 : 65534;
        K1Params Q = P;
        Q.bgr = P.bgr + (size_t)f0 * P.fstride; Q.rec = P.rec + f0;
        Q.thumbs = P.thumbs + (size_t)f0 * AVD_T * AVD_T; Q.area32 = P.area32 + (size_t)f0 * 1024;
        Q.frame0 = f0; Q.nframes = nb;
        AVD_LAUNCH(ctx, "k1_fused", st, (k1_fused_kernel<MAXT, MINB, JT, TM, PAIR><<<dim3(AVD_D, PAIR ? (nb + 1) / 2 : nb), threads, smem, st>>>(Q)));
    }
    return AVD_OK;
}

int k1_fused(avd_ctx *ctx, const uint8_t *d_bgr, int n, int H, int W, size_t pitch, size_t fstride,
             const ShapeTables *tab, const ShapeTables *htab, avd_sample_rec *d_rec, uint8_t *d_thumbs,
             uint8_t *d_area32, uint32_t *d_fixlist, cudaStream_t st)
{
    K1Params P{};
    P.bgr = d_bgr; P.pitch = pitch; P.fstride = fstride; P.H = H; P.W = W; P.row_bytes = 3 * W;
    P.tab = tab; P.rec = d_rec; P.thumbs = d_thumbs; P.area32 = d_area32; P.fixlist = d_fixlist;
    const size_t smem = k1_smem_bytes(W);
    const int fpc = k1_frames_per_cta(W);
    const int threads = (fpc * (W / 16) + 31) / 32 * 32;
    const int jt = htab->strip_ntx_max;
    // pair-sum thumbnail: every horizontal tap pair is (1024, 1024) and all left taps have the same parity
    bool pair = true;
    for (int i = 0; i < AVD_T; i++)
        if (htab->lx[i].a0 != 1024 || htab->lx[i].a1 != 1024 || ((htab->lx[i].i0 ^ htab->lx[0].i0) & 1) ||
            htab->ly[i].a0 + htab->ly[i].a1 != 2048) pair = false;
    if (fpc == 2 && pair) return k1_launch<160, 3, 4, 4, true>(ctx, P, n, smem, threads, st);      // W = 1280: two frames per CTA
    if (threads <= 128) {                                            // <= 2048 px wide
        if (pair) {
            if (W == 1280) return k1_launch<128, K1_MINB, 4, 4>(ctx, P, n, smem, threads, st);
            if (jt <= 3) return k1_launch<128, K1_MINB, 3, 1>(ctx, P, n, smem, threads, st);
            return k1_launch<128, K1_MINB, 4, 1>(ctx, P, n, smem, threads, st);
        }
        if (jt <= 3) return k1_launch<128, K1_MINB, 3, 0>(ctx, P, n, smem, threads, st);
        return k1_launch<128, K1_MINB, 4, 0>(ctx, P, n, smem, threads, st);
    }
    if (pair) {
        if (W == 2560) return k1_launch<256, 2, 2, 8>(ctx, P, n, smem, threads, st);
        if (jt <= 2) return k1_launch<256, 2, 2, 1>(ctx, P, n, smem, threads, st);
        return k1_launch<256, 2, 3, 1>(ctx, P, n, smem, threads, st);
    }
    if (jt <= 2) return k1_launch<256, 2, 2, 0>(ctx, P, n, smem, threads, st);
    return k1_launch<256, 2, 3, 0>(ctx, P, n, smem, threads, st);
}
