// K1, fused path: ONE pass over each BGR frame produces everything video.py:36-52 derives from it
//   gray (video.py:5,43,51; never written to memory), Laplacian sum / sum of squares (video.py:52),
//   the 320x320 INTER_LINEAR thumbnail (video.py:43) and the 32x32 INTER_AREA cell means (video.py:6).
// HBM traffic = H*W*3 bytes read once + 102400 + 1024 + 16 bytes written per frame.
//
// Decomposition: CTA = (frame, band b), band b = the source rows of INTER_AREA destination row b
// (H/32 rows plus one halo row above and below for the Laplacian / fractional cell taps).  Rows stream
// through shared memory with TMA bulk copies (cp.async.bulk -> UBLKCP) on an mbarrier ring; each thread
// owns a 16-pixel column strip (48 bytes = 3 conflict-free LDS.128) and marches down the band keeping
// the previous two gray rows in registers, so the Laplacian needs no second read of anything.
//
// Instruction budget: at 6.5 TB/s an SM has ~16 issue slots per pixel, so the arithmetic is packed:
//   gray        2 IDP.2A per pixel on 16-bit weights doubled so that Y lands in byte 2 of the accumulator
//   Laplacian   two pixels per register in 16-bit lanes, biased by +1020 to stay non-negative
//   sum/sumsq   IDP.2A / IDP.4A on the split low/high bytes of the biased response
//   INTER_AREA  float(g)*alpha as one exact FFMA on the magic-number float 2^23+g, then the mandated
//               sequential FADD (bit-exact with the library's float32 accumulation order)
#include "avd_internal.h"

namespace {

constexpr int CH = 2;                 // rows per TMA stage
constexpr int NSTAGE = 2;             // stages in flight per CTA (several CTAs per SM overlap)
constexpr int GR = 4;                 // rows per resample group (multiple of CH)
constexpr int RC = GR + CH + 1;       // gray ring rows: group + carry row + first chunk of the next group

struct K1Params {
    const uint8_t *bgr;
    size_t pitch, fstride;
    int H, W, row_bytes;
    const ShapeTables *tab;
    avd_sample_rec *rec;
    uint8_t *thumbs, *area32;
    int area_aligned;                 // W % 128 == 0: INTER_AREA cells are whole aligned words
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
__device__ __forceinline__ void tma_load_1d(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
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

// biased Laplacian of 4 pixels (two 16-bit-lane registers) and its statistics
__device__ __forceinline__ void lap4(uint32_t he, uint32_t ho, uint32_t ec, uint32_t oc, uint32_t eu, uint32_t ou,
                                     uint32_t ed, uint32_t od, LapAcc &A)
{
    uint32_t le = he + eu + ed + KBIAS - 4u * ec;            // lanes: L + 1020 in [0, 2040]
    uint32_t lo = ho + ou + od + KBIAS - 4u * oc;
    A.s = __dp2a_lo(le, 0x0101u, A.s);
    A.s = __dp2a_lo(lo, 0x0101u, A.s);
    uint32_t pl = __byte_perm(le, lo, 0x6420), ph = __byte_perm(le, lo, 0x7531);
    A.ll = __dp4a(pl, pl, A.ll);
    A.hl = __dp4a(ph, pl, A.hl);
    A.hh = __dp4a(ph, ph, A.hh);
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

__global__ void __launch_bounds__(256) k1_fused_kernel(K1Params P)
{
    extern __shared__ __align__(128) uint8_t smem[];
    const int tid = threadIdx.x, nt = blockDim.x;
    const int b = blockIdx.x, f = blockIdx.y;
    const ShapeTables *__restrict__ tab = P.tab;
    const int H = P.H, W = P.W;
    const int own0 = tab->row_owner_begin[b], own1 = tab->row_owner_begin[b + 1];
    const int lo = max(own0 - 1, 0), hi = min(own1, H - 1);
    const int nchunks = (hi - lo + 1 + CH - 1) / CH;
    const int stage_bytes = CH * P.row_bytes;
    uint8_t *stage = smem + 128;
    uint8_t *ring = stage + NSTAGE * stage_bytes;
    float *rowbuf = (float *)(ring + RC * W);
    const uint32_t bar0 = smem_u32(smem);
    const uint8_t *src = P.bgr + (size_t)f * P.fstride;
    const AreaCell cy = tab->ay[b];
    const int cell_lo = cy.has_head ? cy.s1 - 1 : cy.s1, cell_hi = cy.has_tail ? cy.s2 : cy.s2 - 1;
    const int area_int = tab->area_int;

    if (tid == 0) {
        for (int s = 0; s < NSTAGE; s++) mbar_init(bar0 + 8 * s, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    auto issue = [&](int c) {
        int s = c % NSTAGE, r0 = lo + c * CH, nr = min(CH, hi + 1 - r0);
        mbar_expect_tx(bar0 + 8 * s, (uint32_t)(nr * P.row_bytes));
        for (int j = 0; j < nr; j++)
            tma_load_1d(smem_u32(stage + s * stage_bytes + j * P.row_bytes), src + (size_t)(r0 + j) * P.pitch,
                        (uint32_t)P.row_bytes, bar0 + 8 * s);
    };
    if (tid == 0)
        for (int c = 0; c < min(NSTAGE, nchunks); c++) issue(c);

    const bool active = 16 * tid < W, first = tid == 0, last = 16 * tid + 16 == W;
    uint32_t Eu[4] = {0, 0, 0, 0}, Ou[4] = {0, 0, 0, 0};            // gray row r-2 (16-bit lanes: even / odd pixels)
    uint32_t Ec[4] = {0, 0, 0, 0}, Oc[4] = {0, 0, 0, 0};            // gray row r-1
    uint32_t HEc[4] = {0, 0, 0, 0}, HOc[4] = {0, 0, 0, 0};          // its horizontal neighbour sums
    LapAcc A = {0, 0, 0, 0};
    int ty_next = tab->thumb_row_begin[b];
    const int ty_end = tab->thumb_row_begin[b + 1];
    const int thumb_2x2 = tab->thumb_2x2;
    int g0 = lo;                                                    // first row of the current resample group

    for (int c = 0; c < nchunks; c++) {
        const int s = c % NSTAGE, r0 = lo + c * CH, nr = min(CH, hi + 1 - r0);
        mbar_wait(bar0 + 8 * s, (uint32_t)((c / NSTAGE) & 1));
        if (active) {
            for (int j = 0; j < nr; j++) {
                const int r = r0 + j;
                const uint8_t *sp = stage + s * stage_bytes + j * P.row_bytes + 48 * tid;
                const uint4 q0 = *(const uint4 *)sp, q1 = *(const uint4 *)(sp + 16), q2 = *(const uint4 *)(sp + 32);
                uint32_t a[16];
                gray4(q0.x, q0.y, q0.z, a[0], a[1], a[2], a[3]);
                gray4(q0.w, q1.x, q1.y, a[4], a[5], a[6], a[7]);
                gray4(q1.z, q1.w, q2.x, a[8], a[9], a[10], a[11]);
                gray4(q2.y, q2.z, q2.w, a[12], a[13], a[14], a[15]);
                // neighbour pixels x-1 and x+16 (REFLECT_101 at the frame edges: x=-1 -> 1, x=W -> W-2)
                uint32_t aL, aR;
                if (first) aL = a[1];
                else { uint32_t w = *(const uint32_t *)(sp - 4); aL = __dp2a_hi(C_GR, w, __dp2a_lo(C_0B, w, RND)); }
                if (last) aR = a[14];
                else { uint32_t w = *(const uint32_t *)(sp + 48); aR = __dp2a_hi(C_R0, w, __dp2a_lo(C_BG, w, RND)); }
                uint32_t E[4], O[4], HE[4], HO[4];
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    E[g] = __byte_perm(a[4 * g], a[4 * g + 2], 0x7632);       // (g0, g2) in 16-bit lanes
                    O[g] = __byte_perm(a[4 * g + 1], a[4 * g + 3], 0x7632);   // (g1, g3)
                }
                uint4 gw;
                gw.x = __byte_perm(E[0], O[0], 0x6240); gw.y = __byte_perm(E[1], O[1], 0x6240);
                gw.z = __byte_perm(E[2], O[2], 0x6240); gw.w = __byte_perm(E[3], O[3], 0x6240);
                *(uint4 *)(ring + ((r - lo) % RC) * W + 16 * tid) = gw;
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    // left neighbours of (g0,g2) are (gL,g1); right neighbours of (g1,g3) are (g2,gR)
                    uint32_t el = g == 0 ? __byte_perm(aL, O[0], 0x5432) : __byte_perm(O[g - 1], O[g], 0x5432);
                    uint32_t orr = g == 3 ? __byte_perm(E[3], aR, 0x7632) : __byte_perm(E[g], E[g + 1], 0x5432);
                    HE[g] = el + O[g];
                    HO[g] = E[g] + orr;
                }
                const int cr = r - 1;                                // centre row whose Laplacian is now complete
                if (cr >= own0 && cr < own1) {
                    if (cr == 0) {
#pragma unroll
                        for (int g = 0; g < 4; g++) lap4(HEc[g], HOc[g], Ec[g], Oc[g], E[g], O[g], E[g], O[g], A);
                    } else {
#pragma unroll
                        for (int g = 0; g < 4; g++) lap4(HEc[g], HOc[g], Ec[g], Oc[g], Eu[g], Ou[g], E[g], O[g], A);
                    }
                }
#pragma unroll
                for (int g = 0; g < 4; g++) {
                    Eu[g] = Ec[g]; Ou[g] = Oc[g];
                    Ec[g] = E[g]; Oc[g] = O[g]; HEc[g] = HE[g]; HOc[g] = HO[g];
                }
            }
        }
        __syncthreads();                                             // gray rows visible; stage s is free again
        if (tid == 0 && c + NSTAGE < nchunks) issue(c + NSTAGE);

        const int g1 = r0 + nr;                                      // one past the last row converted so far
        if (g1 - g0 >= GR || c == nchunks - 1) {
            // ---- INTER_AREA horizontal pass for the cell rows of this group (one task per (row, dx)) ----
            const int ar0 = max(g0, cell_lo), ar1 = min(g1 - 1, cell_hi);
            for (int id = tid; id < (ar1 - ar0 + 1) * 32; id += nt) {
                const int row = ar0 + (id >> 5), dx = id & 31;
                const uint8_t *gr = ring + ((row - lo) % RC) * W;
                float out;
                if (area_int) {
                    const int isx = tab->isx;
                    unsigned sum = 0;
                    if (P.area_aligned) {
                        const uint32_t *wp = (const uint32_t *)(gr + dx * isx);
                        for (int k = 0; k < (isx >> 2); k++) sum = __dp4a(wp[k], 0x01010101u, sum);
                    } else {
                        for (int k = 0; k < isx; k++) sum += gr[dx * isx + k];
                    }
                    out = __int_as_float(sum);
                } else if (P.area_aligned) {
                    const int cw = W >> 5;
                    const float alpha = tab->ax[0].am, nb = -8388608.0f * alpha;
                    const uint32_t *wp = (const uint32_t *)(gr + dx * cw);
                    float buf = 0.f;
                    for (int k = 0; k < (cw >> 2); k++) {
                        uint32_t w = wp[k];
                        buf = __fadd_rn(buf, mulbyte(w, 0x7650, alpha, nb));
                        buf = __fadd_rn(buf, mulbyte(w, 0x7651, alpha, nb));
                        buf = __fadd_rn(buf, mulbyte(w, 0x7652, alpha, nb));
                        buf = __fadd_rn(buf, mulbyte(w, 0x7653, alpha, nb));
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
                rowbuf[(row - cell_lo) * 32 + dx] = out;
            }
            // ---- thumbnail rows whose lower tap row is now available ----
            int tcount = 0;
            while (ty_next + tcount < ty_end && (thumb_2x2 ? 2 * (ty_next + tcount) + 1 : tab->ly[ty_next + tcount].i1) < g1) tcount++;
            for (int id = tid; id < tcount * AVD_T; id += nt) {
                const int ty = ty_next + id / AVD_T, tx = id % AVD_T;
                uint8_t o;
                if (thumb_2x2) {
                    const uint8_t *p0 = ring + ((2 * ty - lo) % RC) * W + 2 * tx, *p1 = ring + ((2 * ty + 1 - lo) % RC) * W + 2 * tx;
                    o = (uint8_t)((p0[0] + p0[1] + p1[0] + p1[1] + 2) >> 2);
                } else {
                    const LinTap vx = tab->lx[tx], vy = tab->ly[ty];
                    const uint8_t *p0 = ring + ((vy.i0 - lo) % RC) * W, *p1 = ring + ((vy.i1 - lo) % RC) * W;
                    int S0 = p0[vx.i0] * vx.a0 + p0[vx.i1] * vx.a1;
                    int S1 = p1[vx.i0] * vx.a0 + p1[vx.i1] * vx.a1;
                    o = (uint8_t)((((vy.a0 * (S0 >> 4)) >> 16) + ((vy.a1 * (S1 >> 4)) >> 16) + 2) >> 2);
                }
                P.thumbs[((size_t)f * AVD_T + ty) * AVD_T + tx] = o;
            }
            ty_next += tcount;
            g0 = g1;
        }
    }
    // bottom frame edge: Laplacian of row H-1 with REFLECT_101 (row H -> H-2)
    if (active && own1 == H) {
#pragma unroll
        for (int g = 0; g < 4; g++) lap4(HEc[g], HOc[g], Ec[g], Oc[g], Eu[g], Ou[g], Eu[g], Ou[g], A);
    }

    // ---- Laplacian sums: warp shuffle -> shared -> one 64-bit atomic pair per CTA ----
    unsigned long long s1 = A.s, sq = (unsigned long long)A.ll + 512ull * A.hl + 65536ull * A.hh;
    for (int o = 16; o; o >>= 1) { s1 += __shfl_down_sync(0xffffffffu, s1, o); sq += __shfl_down_sync(0xffffffffu, sq, o); }
    __shared__ unsigned long long red[2][32];
    if ((tid & 31) == 0) { red[0][tid >> 5] = s1; red[1][tid >> 5] = sq; }
    __syncthreads();                                                 // also orders rowbuf for the vertical pass
    if (tid < 32) {
        const int nw = (nt + 31) >> 5;
        s1 = tid < nw ? red[0][tid] : 0; sq = tid < nw ? red[1][tid] : 0;
        for (int o = 16; o; o >>= 1) { s1 += __shfl_down_sync(0xffffffffu, s1, o); sq += __shfl_down_sync(0xffffffffu, sq, o); }
        if (tid == 0) {
            const long long npx = (long long)(own1 - own0) * W;
            const long long sumL = (long long)s1 - 1020ll * npx;
            const long long sumL2 = (long long)sq - 2040ll * (long long)s1 + 1040400ll * npx;
            atomicAdd((unsigned long long *)&P.rec[f].lap_sum, (unsigned long long)sumL);
            atomicAdd((unsigned long long *)&P.rec[f].lap_sumsq, (unsigned long long)sumL2);
        }
        // ---- INTER_AREA vertical pass: destination row b, column tid ----
        uint8_t o;
        if (area_int) {
            int tot = 0;
            for (int r = cell_lo; r <= cell_hi; r++) tot += __float_as_int(rowbuf[(r - cell_lo) * 32 + tid]);
            o = sat_rne(__fmul_rn((float)tot, tab->inv_area));
        } else {
            float sum = 0.f;
            if (cy.has_head) sum = __fadd_rn(sum, __fmul_rn(cy.ah, rowbuf[tid]));
            for (int r = cy.s1; r < cy.s2; r++) sum = __fadd_rn(sum, __fmul_rn(cy.am, rowbuf[(r - cell_lo) * 32 + tid]));
            if (cy.has_tail) sum = __fadd_rn(sum, __fmul_rn(cy.at, rowbuf[(cy.s2 - cell_lo) * 32 + tid]));
            o = sat_rne(sum);
        }
        P.area32[(size_t)f * 1024 + b * 32 + tid] = o;
    }
}

}  // namespace

bool k1_fused_supported(int H, int W, size_t pitch, size_t fstride, const void *base)
{
    return W % 16 == 0 && W <= 4096 && pitch % 16 == 0 && fstride % 16 == 0 && ((uintptr_t)base % 16) == 0 && H >= 32;
}

int k1_fused(avd_ctx *ctx, const uint8_t *d_bgr, int n, int H, int W, size_t pitch, size_t fstride,
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
    }
    size_t smem = 128 + (size_t)NSTAGE * CH * 3 * W + (size_t)RC * W + (size_t)max_cell_rows * 32 * sizeof(float);
    if (smem > 227 * 1024) { ctx->err = "fused K1: frame too wide for shared memory"; return AVD_ERR_ARG; }
    static size_t smem_set = 0;
    if (smem > smem_set) {
        AVD_CUDA(ctx, cudaFuncSetAttribute(k1_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        smem_set = smem;
    }
    int threads = ((W / 16) + 31) / 32 * 32;
    for (int f0 = 0; f0 < n; f0 += 65535) {                          // grid.y limit
        int nb = n - f0 < 65535 ? n - f0 : 65535;
        K1Params Q = P;
        Q.bgr = d_bgr + (size_t)f0 * fstride; Q.rec = d_rec + f0;
        Q.thumbs = d_thumbs + (size_t)f0 * AVD_T * AVD_T; Q.area32 = d_area32 + (size_t)f0 * 1024;
        AVD_LAUNCH(ctx, "k1_fused", st, k1_fused_kernel<<<dim3(AVD_D, nb), threads, smem, st>>>(Q));
    }
    return AVD_OK;
}
