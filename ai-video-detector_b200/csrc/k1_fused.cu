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
//                stage); the warp that is LAST to finish reading a stage refills it, nobody waits for a free stage
//   main warps   a thread owns a 16-pixel column strip (48 bytes = 3 conflict-free LDS.128) and marches down the
//                band keeping three rows in registers (straight-line code, three rotating register sets):
//                gray        2 IDP.2A per pixel (15-bit weights doubled so that Y lands in byte 2)
//                Laplacian   two pixels per register in 16-bit lanes, biased by +1020; sum of squares through
//                            byte-split IDP.4A; the plain sum telescopes to strip/band borders
//                INTER_AREA  exact integer cell sums (one IADD3 per 4 pixels) -- see "certificate" below
//   thumb warp   one extra warp turns the tap rows of the thumbnail (written by the main warps into a small gray
//                ring, handed over with mbarriers) into thumbnail rows, so the main warps never meet a block barrier
//
// INTER_AREA certificate.  The library accumulates float32 products sequentially; the result is
// rounded to u8 and only that byte is observable.  The kernel computes the exact rational cell mean from integer
// sums; float32 evaluation in the library's order differs from it by at most (nx+ny+4) * 2^-24 relative
// (standard forward error bound of recursive summation, nx/ny = taps per cell), so whenever the exact mean is
// farther than twice that bound from a rounding boundary k+0.5 the library's byte is round(mean).  The few cells
// (~0.5 %) that are closer are recomputed at the end of the band in the library's exact float32 order.
#include "avd_internal.h"

namespace {

#ifndef K1_NRS
#define K1_NRS 3
#endif
#ifndef K1_MINB
#define K1_MINB 4
#endif
constexpr int NRS = K1_NRS;           // single-row TMA stages per CTA (multiple of 3: stage index is static per unrolled step)
static_assert(NRS == 3, "stage index is tied to the 3-way register rotation");
constexpr int RP = 4;                 // tap-row pairs in the gray ring
constexpr int MAXQ = 160;             // rows a band can load (H/32 + 3)
constexpr int NTX = AVD_T / 32;       // thumbnail columns per lane of the thumb warp

// shared-memory header (bytes)
constexpr int OFF_FULL = 0, OFF_RFULL = 64, OFF_REMPTY = 96, OFF_CNT = 128, OFF_MISC = 160, OFF_CELL = 256, OFF_RED = 384,
              OFF_INFO = 576, HDR = 768;
static_assert(OFF_INFO + MAXQ <= HDR, "header layout");

struct K1Params {
    const uint8_t *bgr;
    size_t pitch, fstride;
    int H, W, row_bytes;
    const ShapeTables *tab;
    avd_sample_rec *rec;
    uint8_t *thumbs, *area32;
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
__device__ __forceinline__ void mbar_arrive(uint32_t bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar, uint64_t policy)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar), "l"(policy) : "memory");
}
__device__ __noinline__ void mbar_wait_slow(uint32_t bar, uint32_t parity)
{
    const long long t0 = clock64();
    while (true) {
        uint32_t ok;
        asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
        if (ok) return;
        if (clock64() - t0 > 4000000000ll) __trap();     // ~2 s: a lost refill must fail loudly, not hang the GPU
    }
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
    uint32_t ok;
    asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], %2;\nselp.u32 %0, 1, 0, P1;\n}" : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    if (!ok) mbar_wait_slow(bar, parity);
}
// lane 0 of a warp counts the warp off a stage; returns 1 in the lane that arrived last (pending count before its arrival)
__device__ __forceinline__ uint32_t stage_arrive(uint32_t bar, uint32_t lane0)
{
    uint32_t pend = 0;
    asm volatile("{\n.reg .pred p;\n.reg .b64 st;\nsetp.ne.u32 p, %2, 0;\n@p mbarrier.arrive.shared::cta.b64 st, [%1];\n@p mbarrier.pending_count.b64 %0, st;\n}"
                 : "+r"(pend) : "r"(bar), "r"(lane0) : "memory");
    return pend;
}
__device__ __forceinline__ void main_sync(int n) { asm volatile("bar.sync 1, %0;" ::"r"(n) : "memory"); }

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

// Laplacian statistics of centre row C (rows above / below: U, D) and its INTER_AREA column sums
__device__ __forceinline__ void lap_step(const RowSet &U, const RowSet &C, const RowSet &D, Acc &A)
{
#pragma unroll
    for (int g = 0; g < 4; g++) {
        // left neighbours of (p0,p2) are (pL,p1); right neighbours of (p1,p3) are (p2,pR)
        const uint32_t el = g == 0 ? __byte_perm(C.LR, C.O[0], 0x5410) : __byte_perm(C.O[g - 1], C.O[g], 0x5432);
        const uint32_t orr = g == 3 ? __byte_perm(C.E[3], C.LR, 0x7632) : __byte_perm(C.E[g], C.E[g + 1], 0x5432);
        const uint32_t me = KBIAS - 4u * C.E[g], mo = KBIAS - 4u * C.O[g];
        const uint32_t le = (el + C.O[g] + me) + U.E[g] + D.E[g];      // L+1020 of (p0,p2), each in [0,2040]
        const uint32_t lo = (C.E[g] + orr + mo) + U.O[g] + D.O[g];     // L+1020 of (p1,p3)
        const uint32_t pl = __byte_perm(le, lo, 0x6420), ph = __byte_perm(le, lo, 0x7531);
        A.ll = __dp4a(pl, pl, A.ll);
        A.hl = __dp4a(ph, pl, A.hl);
        A.hh = __dp4a(ph, ph, A.hh);
        A.cs[g] += C.E[g] + C.O[g];
    }
    A.aLR += C.LR; A.aE0 += C.E[0]; A.aO3 += C.O[3];
}

struct Band {
    uint32_t sbase;                       // shared address of the header
    uint32_t stage;                       // shared address of stage 0
    uint8_t *ring;                        // gray ring (generic pointer)
    int row_bytes, W, H, own0, nseq, nw;
    const uint8_t *frame;
    size_t pitch;
    uint64_t pol_stream, pol_keep;
};

// sequence row q of the band = source row own0-1+q, reflected at the frame edges (BORDER_REFLECT_101)
__device__ __forceinline__ void issue_row(const Band &B, int q, int s)
{
    int r = B.own0 - 1 + q;
    r = r < 0 ? 1 : (r >= B.H ? B.H - 2 : r);
    const uint32_t bar = B.sbase + OFF_FULL + 8 * s;
    mbar_expect_tx(bar, (uint32_t)B.row_bytes);
    // the two first and two last rows of a band are also read by the neighbouring bands: do not mark them evict-first
    const bool shared_row = q < 2 || q >= B.nseq - 2;
    tma_load_1d(B.stage + s * B.row_bytes, B.frame + (size_t)r * B.pitch, (uint32_t)B.row_bytes, bar, shared_row ? B.pol_keep : B.pol_stream);
}

// Stage s holds sequence row q: convert this thread's strip to gray (row set R); tap rows of the thumbnail also go
// to the gray ring.  Returns the stage's arrival ticket of this warp (lane 0), to be passed to release_finish.
template <int S>
__device__ __forceinline__ uint32_t load_row(const Band &B, int q, uint32_t parity, int tid, bool active, bool first, bool last, RowSet &R)
{
    mbar_wait(B.sbase + OFF_FULL + 8 * S, parity);
    const uint32_t sp = B.stage + S * B.row_bytes + 48 * tid;
    uint4 q0, q1, q2;
    uint32_t wl, wr;
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(q0.x), "=r"(q0.y), "=r"(q0.z), "=r"(q0.w) : "r"(sp));
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+16];" : "=r"(q1.x), "=r"(q1.y), "=r"(q1.z), "=r"(q1.w) : "r"(sp));
    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4+32];" : "=r"(q2.x), "=r"(q2.y), "=r"(q2.z), "=r"(q2.w) : "r"(sp));
    asm volatile("ld.shared.u32 %0, [%1+-4];" : "=r"(wl) : "r"(sp));
    asm volatile("ld.shared.u32 %0, [%1+48];" : "=r"(wr) : "r"(sp));
    // every lane of this warp has its bytes in registers: count the warp off the stage (result used much later)
    __syncwarp();
    const uint32_t ticket = stage_arrive(B.sbase + OFF_CNT + 8 * S, (tid & 31) == 0 ? 1u : 0u);
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
        R.E[g] = __byte_perm(a[4 * g], a[4 * g + 2], 0x7632);
        R.O[g] = __byte_perm(a[4 * g + 1], a[4 * g + 3], 0x7632);
    }
    R.LR = __byte_perm(aL, aR, 0x7632);
    // thumbnail tap row?  info: bit0 tap, bit1 second row of its pair, bits 2-3 pair slot, bit4 wait for the slot, bit5 its parity
    uint32_t info;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(info) : "r"(B.sbase + OFF_INFO + q));
    if (info & 1u) {
        const uint32_t slot = (info >> 2) & 3u;
        if (info & 16u) mbar_wait(B.sbase + OFF_REMPTY + 8 * slot, (info >> 5) & 1u);
        uint4 gw;
        gw.x = __byte_perm(R.E[0], R.O[0], 0x6240); gw.y = __byte_perm(R.E[1], R.O[1], 0x6240);
        gw.z = __byte_perm(R.E[2], R.O[2], 0x6240); gw.w = __byte_perm(R.E[3], R.O[3], 0x6240);
        if (active) *(uint4 *)(B.ring + (2 * slot + ((info >> 1) & 1u)) * B.W + 16 * tid) = gw;
        if (info & 2u) {
            __syncwarp();
            if ((tid & 31) == 0) mbar_arrive(B.sbase + OFF_RFULL + 8 * slot);
        }
    }
    return ticket;
}

// the warp that arrived last at stage S refills it with sequence row q + NRS
template <int S>
__device__ __forceinline__ void release_finish(const Band &B, int q, uint32_t ticket, int tid)
{
    if (ticket == 1u && q + NRS < B.nseq) issue_row(B, q + NRS, S);
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

// INTER_AREA cell (dy,dx) of frame `src` in the library's exact float32 order, one warp per cell
__device__ float area_cell_exact(const uint8_t *src, size_t pitch, const ShapeTables *tab, int dy, int dx, int lane)
{
    const AreaCell cy = tab->ay[dy], cx = tab->ax[dx];
    const int r0 = cy.has_head ? cy.s1 - 1 : cy.s1, r1 = cy.has_tail ? cy.s2 : cy.s2 - 1;
    float sum = 0.f;
    for (int rb = r0; rb <= r1; rb += 32) {
        const int r = rb + lane;
        float term = 0.f;
        if (r <= r1) {
            const uint8_t *p = src + (size_t)r * pitch;
            auto gray = [&](int x) { const uint8_t *q = p + 3 * x; return (float)((3735u * q[0] + 19235u * q[1] + 9798u * q[2] + 16384u) >> 15); };
            float buf = 0.f;
            if (cx.has_head) buf = __fadd_rn(buf, __fmul_rn(gray(cx.s1 - 1), cx.ah));
            for (int x = cx.s1; x < cx.s2; x++) buf = __fadd_rn(buf, __fmul_rn(gray(x), cx.am));
            if (cx.has_tail) buf = __fadd_rn(buf, __fmul_rn(gray(cx.s2), cx.at));
            const float beta = (cy.has_head && r == r0) ? cy.ah : ((cy.has_tail && r == cy.s2) ? cy.at : cy.am);
            term = __fmul_rn(beta, buf);
        }
        const int cnt = min(32, r1 - rb + 1);
        for (int i = 0; i < cnt; i++) sum = __fadd_rn(sum, __shfl_sync(0xffffffffu, term, i));
    }
    return sum;
}

template <int MAXT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) k1_fused_kernel(K1Params P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, lane = tid & 31;
    const int nmain = blockDim.x - 32, nw = nmain >> 5;
    const int b = blockIdx.x, f = blockIdx.y;
    const ShapeTables *__restrict__ tab = P.tab;
    const int H = P.H, W = P.W;
    const int own0 = tab->row_owner_begin[b], own1 = tab->row_owner_begin[b + 1];
    const int n = own1 - own0, nseq = n + 2;
    const int ty_begin = tab->thumb_row_begin[b], nty = tab->thumb_row_begin[b + 1] - ty_begin;
    const uint32_t sbase = smem_u32(smem);
    uint8_t *ring = smem + HDR + NRS * P.row_bytes;
    const uint8_t *frame = P.bgr + (size_t)f * P.fstride;

    Band B;
    B.sbase = sbase; B.stage = sbase + HDR; B.ring = ring; B.row_bytes = P.row_bytes; B.W = W; B.H = H; B.own0 = own0;
    B.nseq = nseq; B.nw = nw; B.frame = frame; B.pitch = P.pitch;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(B.pol_stream));
    asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(B.pol_keep));

    if (tid == 0) {
        for (int s = 0; s < NRS; s++) {
            mbar_init(sbase + OFF_FULL + 8 * s, 1);
            mbar_init(sbase + OFF_CNT + 8 * s, nw);
        }
        for (int p = 0; p < RP; p++) { mbar_init(sbase + OFF_RFULL + 8 * p, nw); mbar_init(sbase + OFF_REMPTY + 8 * p, 1); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
        for (int q = 0; q < NRS; q++) issue_row(B, q, q);
    }
    if (tid < 32) ((int *)(smem + OFF_CELL))[tid] = 0;

    if (tid >= nmain) {
        // ================= thumb warp =================
        uint8_t *info = smem + OFF_INFO;
        for (int q = lane; q < nseq; q += 32) info[q] = 0;
        __syncwarp();
        for (int j = lane; j < nty; j += 32) {
            const int q0 = tab->ly[ty_begin + j].i0 - (own0 - 1);
            const int slot = j % RP, rnd = j / RP;
            info[q0] = (uint8_t)(1 | (slot << 2) | (rnd > 0 ? 16 : 0) | (((rnd - 1) & 1) << 5));
            info[q0 + 1] = (uint8_t)(1 | 2 | (slot << 2));
        }
        int x0[NTX], a0[NTX], a1[NTX];
#pragma unroll
        for (int k = 0; k < NTX; k++) {
            const LinTap t = tab->lx[lane + 32 * k];
            x0[k] = t.i0; a0[k] = t.a0; a1[k] = t.a1;
        }
        __syncthreads();
        for (int j = 0; j < nty; j++) {
            const int slot = j % RP;
            const LinTap vy = tab->ly[ty_begin + j];
            mbar_wait(sbase + OFF_RFULL + 8 * slot, (uint32_t)((j / RP) & 1));
            const uint8_t *p0 = ring + (2 * slot) * W, *p1 = p0 + W;
            uint8_t *dst = P.thumbs + ((size_t)f * AVD_T + (ty_begin + j)) * AVD_T + lane;
#pragma unroll
            for (int k = 0; k < NTX; k++) {
                const int S0 = p0[x0[k]] * a0[k] + p0[x0[k] + 1] * a1[k];
                const int S1 = p1[x0[k]] * a0[k] + p1[x0[k] + 1] * a1[k];
                dst[32 * k] = (uint8_t)((((vy.a0 * (S0 >> 4)) >> 16) + ((vy.a1 * (S1 >> 4)) >> 16) + 2) >> 2);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(sbase + OFF_REMPTY + 8 * slot);
        }
        return;
    }

    // ================= main warps =================
    __syncthreads();
    const bool active = 16 * tid < W, first = tid == 0, last = 16 * tid + 16 == W;
    RowSet R0, R1, R2;
    Acc A;
    A.ll = A.hl = A.hh = 0; A.aLR = A.aE0 = A.aO3 = 0; A.cs[0] = A.cs[1] = A.cs[2] = A.cs[3] = 0;

    // INTER_AREA: this strip's pair lanes split between cell cA (pairs < kA) and cell cA+1
    const int sx = W >> 5;
    const int cA = (16 * tid) / sx;
    const int kA = min(8, ((cA + 1) * sx - 16 * tid) >> 1);
    uint32_t mA[4];
#pragma unroll
    for (int g = 0; g < 4; g++) mA[g] = (2 * g < kA ? 1u : 0u) | (2 * g + 1 < kA ? 0x100u : 0u);
    const uint32_t mAll[4] = {0x0101u, 0x0101u, 0x0101u, 0x0101u};
    const int wh = 32 * own0 - b * H;                     // weight of the head row (sequence row 0), 1/32 units
    const int wt = (b + 1) * H - 32 * (own1 - 1);         // weight of the last owned row, 1/32 units (32 = full)

    uint32_t t0 = load_row<0>(B, 0, 0u, tid, active, first, last, R0);
    release_finish<0>(B, 0, t0, tid);
    uint32_t t1 = load_row<1>(B, 1, 0u, tid, active, first, last, R1);
    release_finish<1>(B, 1, t1, tid);
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
    auto band_end = [&](const RowSet &C, const RowSet &D) {
        tv += (int)row_sum(D) - (int)row_sum(C);
        uint32_t Pp[4];
#pragma unroll
        for (int g = 0; g < 4; g++) Pp[g] = C.E[g] + C.O[g];
        const uint32_t sa = masked_sum(Pp, mA), st = masked_sum(Pp, mAll);
        mcA -= (32 - wt) * (int)sa; mcB -= (32 - wt) * (int)(st - sa);
    };

    uint32_t par = 0;
    int q = 2;
    while (true) {
        uint32_t tk = load_row<2>(B, q, par, tid, active, first, last, R2);
        lap_step(R0, R1, R2, A);
        release_finish<2>(B, q, tk, tid);
        if (q == nseq - 1) { band_end(R1, R2); break; }
        q++; par ^= 1u;
        tk = load_row<0>(B, q, par, tid, active, first, last, R0);
        lap_step(R1, R2, R0, A);
        release_finish<0>(B, q, tk, tid);
        if (q == nseq - 1) { band_end(R2, R0); break; }
        q++;
        tk = load_row<1>(B, q, par, tid, active, first, last, R1);
        lap_step(R2, R0, R1, A);
        release_finish<1>(B, q, tk, tid);
        if (q == nseq - 1) { band_end(R0, R1); break; }
        q++;
    }

    // ---- per-thread totals ----
    {
        const uint32_t sa = masked_sum(A.cs, mA), st = masked_sum(A.cs, mAll);
        mcA += 32 * (int)sa; mcB += 32 * (int)(st - sa);
    }
    long long sumL = (long long)tv + (long long)lanes_sum(A.aLR) - (long long)(A.aE0 & 0xffffu) - (long long)(A.aO3 >> 16);
    unsigned long long sq = (unsigned long long)A.ll + 512ull * A.hl + 65536ull * A.hh;
    if (!active) { sumL = 0; sq = 0; mcA = 0; mcB = 0; }
    int *cellM = (int *)(smem + OFF_CELL);
    if (active) {
        atomicAdd(&cellM[cA], mcA);
        if (kA < 8) atomicAdd(&cellM[cA + 1], mcB);
    }
    for (int o = 16; o; o >>= 1) { sumL += __shfl_down_sync(0xffffffffu, sumL, o); sq += __shfl_down_sync(0xffffffffu, sq, o); }
    unsigned long long *red = (unsigned long long *)(smem + OFF_RED);
    if (lane == 0) { red[tid >> 5] = (unsigned long long)sumL; red[8 + (tid >> 5)] = sq; }
    main_sync(nmain);
    uint32_t *misc = (uint32_t *)(smem + OFF_MISC);
    if (tid < 32) {
        long long s1 = tid < nw ? (long long)red[tid] : 0;
        unsigned long long s2 = tid < nw ? red[8 + tid] : 0;
        for (int o = 16; o; o >>= 1) { s1 += __shfl_down_sync(0xffffffffu, s1, o); s2 += __shfl_down_sync(0xffffffffu, s2, o); }
        if (tid == 0) {
            const long long npx = (long long)n * W;
            const long long sumL2 = (long long)s2 - 2040ll * s1 - 1040400ll * npx;       // s2 = sum (L+1020)^2
            atomicAdd((unsigned long long *)&P.rec[f].lap_sum, (unsigned long long)s1);
            atomicAdd((unsigned long long *)&P.rec[f].lap_sumsq, (unsigned long long)sumL2);
        }
        // ---- INTER_AREA destination row b, column tid ----
        const int M = cellM[tid];
        uint8_t out;
        bool close = false;
        if (tab->area_int) {
            out = sat_rne(__fmul_rn((float)(M >> 5), tab->inv_area));
        } else {
            const double v = (double)M / ((double)H * (double)sx);
            const double fl = floor(v);
            const double bound = v * (2.0 * (double)(sx + n + 5) * 5.9604644775390625e-08) + 1e-9;
            close = fabs(v - fl - 0.5) <= bound;
            out = (uint8_t)min(255, (int)floor(v + 0.5));
        }
        P.area32[(size_t)f * 1024 + b * 32 + tid] = out;
        const uint32_t mask = __ballot_sync(0xffffffffu, close);
        if (tid == 0) misc[0] = mask;
    }
    main_sync(nmain);
    uint32_t mask = misc[0];
    while (mask) {                                      // rare: cells too close to a rounding boundary
        const int dx = __ffs(mask) - 1;
        mask &= mask - 1;
        if ((dx % nw) != (tid >> 5)) continue;
        const float v = area_cell_exact(frame, P.pitch, tab, b, dx, lane);
        if (lane == 0) P.area32[(size_t)f * 1024 + b * 32 + dx] = sat_rne(v);
    }
}

}  // namespace

static size_t k1_smem_bytes(int W) { return HDR + (size_t)NRS * 3 * W + (size_t)2 * RP * W + 64; }

bool k1_fused_supported(int H, int W, size_t pitch, size_t fstride, const void *base, const ShapeTables *htab)
{
    if (W % 64 != 0 || W < 640 || W > 4096 || H < 640 || H > 4096 || htab->thumb_2x2) return false;
    if (pitch % 16 != 0 || fstride % 16 != 0 || ((uintptr_t)base % 16) != 0) return false;
    if (k1_smem_bytes(W) > 227 * 1024) return false;
    // the thumb warp assumes plain two-tap decimation: taps (i, i+1), every tap row used by one thumbnail row
    for (int i = 0; i < AVD_T; i++) {
        if (htab->lx[i].i1 != htab->lx[i].i0 + 1 || htab->ly[i].i1 != htab->ly[i].i0 + 1) return false;
        if (i > 0 && htab->ly[i].i0 < htab->ly[i - 1].i0 + 2) return false;
    }
    return true;
}

int k1_fused(avd_ctx *ctx, const uint8_t *d_bgr, int n, int H, int W, size_t pitch, size_t fstride,
             const ShapeTables *tab, const ShapeTables *htab, avd_sample_rec *d_rec, uint8_t *d_thumbs,
             uint8_t *d_area32, cudaStream_t st)
{
    (void)htab;
    K1Params P{};
    P.bgr = d_bgr; P.pitch = pitch; P.fstride = fstride; P.H = H; P.W = W; P.row_bytes = 3 * W;
    P.tab = tab; P.rec = d_rec; P.thumbs = d_thumbs; P.area32 = d_area32;
    const size_t smem = k1_smem_bytes(W);
    const int nw = (W / 16 + 31) / 32;
    const int threads = (nw + 1) * 32;
    const bool narrow = nw <= 4;                                     // <= 2048 px wide
    static size_t smem_set[2] = {0, 0};
    if (smem > smem_set[narrow]) {
        if (narrow) AVD_CUDA(ctx, cudaFuncSetAttribute(k1_fused_kernel<160, K1_MINB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        else AVD_CUDA(ctx, cudaFuncSetAttribute(k1_fused_kernel<288, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set[narrow] = smem;
    }
    for (int f0 = 0; f0 < n; f0 += 65535) {                          // grid.y limit
        int nb = n - f0 < 65535 ? n - f0 : 65535;
        K1Params Q = P;
        Q.bgr = d_bgr + (size_t)f0 * fstride; Q.rec = d_rec + f0;
        Q.thumbs = d_thumbs + (size_t)f0 * AVD_T * AVD_T; Q.area32 = d_area32 + (size_t)f0 * 1024;
        if (narrow) AVD_LAUNCH(ctx, "k1_fused", st, (k1_fused_kernel<160, K1_MINB><<<dim3(AVD_D, nb), threads, smem, st>>>(Q)));
        else AVD_LAUNCH(ctx, "k1_fused", st, (k1_fused_kernel<288, 2><<<dim3(AVD_D, nb), threads, smem, st>>>(Q)));
    }
    return AVD_OK;
}
