// K0: planar 4:2:0 frames as the host decoder emits them -> the packed BGR frames cv2.VideoCapture.retrieve() returns
// (video.py:32), on the device, so that only 1.5 bytes per pixel cross PCIe instead of 3 (first step).
//
// OpenCV converts a decoded yuv420p / yuvj420p picture with sws_scale(SWS_BICUBIC, same size) -> AV_PIX_FMT_BGR24.  For an
// unscaled 4:2:0 -> 24-bit RGB conversion of even height libswscale takes its dedicated yuv2rgb routine, which on every x86
// CPU with SSSE3 is the 16-bit fixed-point SIMD kernel (not the C table kernel): chroma is replicated 2x2 (no
// interpolation), and per pixel, in signed 16-bit lanes,
//     yc = mulhi((Y << 3) - yOffset, yCoeff)          mulhi(a, b) = (a * b) >> 16, arithmetic shift (PMULHW)
//     ub = mulhi((U << 3) - 1024, ubCoeff)   ug = mulhi((U << 3) - 1024, ugCoeff)
//     vr = mulhi((V << 3) - 1024, vrCoeff)   vg = mulhi((V << 3) - 1024, vgCoeff)
//     B = sat_u8(yc + ub)   G = sat_u8(yc + (ug + vg))   R = sat_u8(yc + vr)
// with the coefficients of the ITU-601 default table (OpenCV never sets colourspace details) in 3.13 fixed point;
// limited range for yuv420p, full range for yuvj420p.  None of the 16-bit saturating adds can saturate for 8-bit input, so
// plain 32-bit integer arithmetic reproduces it.  Pinned against libswscale itself (tests/golden/make_golden_yuv.py calls
// sws_scale of the library cv2 ships) and against cv2.VideoCapture on real MPEG-4 and MJPEG files.
#include "avd_internal.h"

namespace {

struct YuvCoef { int y, vr, ub, vg, ug, yo; };

// libswscale's coefficient set-up (ff_yuv2rgb_c_init_tables) for inv_table = ITU601, brightness 0, contrast = saturation = 1.0
int round_to_int16(long long f)
{
    long long r = (f + (1 << 15)) >> 16;
    return (int)(r < -0x7FFF ? -0x7FFF : (r > 0x7FFF ? 0x7FFF : r));
}

YuvCoef make_coef(bool full_range)
{
    long long crv = 104597, cbu = 132201, cgu = -25675, cgv = -53279, cy = 1 << 16, oy = 0;
    if (!full_range) { cy = (cy * 255) / 219; oy = 16 << 16; }
    else { crv = (crv * 224) / 255; cbu = (cbu * 224) / 255; cgu = (cgu * 224) / 255; cgv = (cgv * 224) / 255; }   // C division: toward zero
    YuvCoef k;
    k.y = round_to_int16(cy * (1 << 13)); k.vr = round_to_int16(crv * (1 << 13)); k.ub = round_to_int16(cbu * (1 << 13));
    k.vg = round_to_int16(cgv * (1 << 13)); k.ug = round_to_int16(cgu * (1 << 13)); k.yo = round_to_int16(oy * 8);
    return k;
}

__device__ __forceinline__ int sat8(int v) { return min(max(v, 0), 255); }

// 16 pixels x 2 rows per thread: 128-bit loads of Y, 64-bit loads of U and V, 128-bit stores of BGR.
__global__ void __launch_bounds__(256) yuv420_bgr_vec_kernel(const uint8_t *__restrict__ yuv, size_t yuv_stride, int H, int W, YuvCoef k,
                                                             uint8_t *__restrict__ bgr, size_t pitch, size_t fstride)
{
    const int strips = W >> 4, sx = blockIdx.x * blockDim.x + threadIdx.x, yp = blockIdx.y, f = blockIdx.z;
    if (sx >= strips) return;
    const uint8_t *Y = yuv + (size_t)f * yuv_stride, *U = Y + (size_t)H * W, *V = U + (size_t)(H >> 1) * (W >> 1);
    const uint2 u8 = *(const uint2 *)(U + (size_t)yp * (W >> 1) + 8 * sx), v8 = *(const uint2 *)(V + (size_t)yp * (W >> 1) + 8 * sx);
    int ub[8], cg[8], vr[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        const int u = (int)(((i < 4 ? u8.x : u8.y) >> (8 * (i & 3))) & 255u), v = (int)(((i < 4 ? v8.x : v8.y) >> (8 * (i & 3))) & 255u);
        const int uw = (u << 3) - 1024, vw = (v << 3) - 1024;
        ub[i] = (uw * k.ub) >> 16;
        cg[i] = ((uw * k.ug) >> 16) + ((vw * k.vg) >> 16);
        vr[i] = (vw * k.vr) >> 16;
    }
#pragma unroll
    for (int r = 0; r < 2; r++) {
        const uint4 y16 = *(const uint4 *)(Y + (size_t)(2 * yp + r) * W + 16 * sx);
        const uint32_t yw[4] = {y16.x, y16.y, y16.z, y16.w};
        uint32_t out[12];
#pragma unroll
        for (int w = 0; w < 12; w++) out[w] = 0;
#pragma unroll
        for (int i = 0; i < 16; i++) {
            const int yv = (int)((yw[i >> 2] >> (8 * (i & 3))) & 255u);
            const int yc = (((yv << 3) - k.yo) * k.y) >> 16;
            const int c = i >> 1;
            const uint32_t b = sat8(yc + ub[c]), g = sat8(yc + cg[c]), rr = sat8(yc + vr[c]);
            out[(3 * i) >> 2] |= b << (8 * ((3 * i) & 3));
            out[(3 * i + 1) >> 2] |= g << (8 * ((3 * i + 1) & 3));
            out[(3 * i + 2) >> 2] |= rr << (8 * ((3 * i + 2) & 3));
        }
        uint4 *dst = (uint4 *)(bgr + (size_t)f * fstride + (size_t)(2 * yp + r) * pitch + 48 * sx);
        dst[0] = make_uint4(out[0], out[1], out[2], out[3]);
        dst[1] = make_uint4(out[4], out[5], out[6], out[7]);
        dst[2] = make_uint4(out[8], out[9], out[10], out[11]);
    }
}

// any even W, H and any alignment: one 2x2 block per thread
__global__ void __launch_bounds__(256) yuv420_bgr_any_kernel(const uint8_t *__restrict__ yuv, size_t yuv_stride, int H, int W, YuvCoef k,
                                                             uint8_t *__restrict__ bgr, size_t pitch, size_t fstride)
{
    const int cx = blockIdx.x * blockDim.x + threadIdx.x, yp = blockIdx.y, f = blockIdx.z;
    if (cx >= (W >> 1)) return;
    const uint8_t *Y = yuv + (size_t)f * yuv_stride, *U = Y + (size_t)H * W, *V = U + (size_t)(H >> 1) * (W >> 1);
    const int u = U[(size_t)yp * (W >> 1) + cx], v = V[(size_t)yp * (W >> 1) + cx];
    const int uw = (u << 3) - 1024, vw = (v << 3) - 1024;
    const int ub = (uw * k.ub) >> 16, cg = ((uw * k.ug) >> 16) + ((vw * k.vg) >> 16), vr = (vw * k.vr) >> 16;
#pragma unroll
    for (int r = 0; r < 2; r++)
#pragma unroll
        for (int c = 0; c < 2; c++) {
            const int x = 2 * cx + c, y = 2 * yp + r;
            const int yc = ((((int)Y[(size_t)y * W + x] << 3) - k.yo) * k.y) >> 16;
            uint8_t *d = bgr + (size_t)f * fstride + (size_t)y * pitch + 3 * x;
            d[0] = (uint8_t)sat8(yc + ub); d[1] = (uint8_t)sat8(yc + cg); d[2] = (uint8_t)sat8(yc + vr);
        }
}

}  // namespace

int k0_yuv420_to_bgr(avd_ctx *ctx, const uint8_t *d_yuv, size_t yuv_stride, int n, int H, int W, bool full_range,
                     uint8_t *d_bgr, size_t pitch, size_t fstride, cudaStream_t st)
{
    static const YuvCoef coef[2] = {make_coef(false), make_coef(true)};
    const YuvCoef k = coef[full_range ? 1 : 0];
    const bool vec = W % 16 == 0 && ((uintptr_t)d_yuv % 16) == 0 && yuv_stride % 16 == 0 && ((uintptr_t)d_bgr % 16) == 0 &&
                     pitch % 16 == 0 && fstride % 16 == 0;
    for (int f0 = 0; f0 < n; f0 += 65535) {     // grid.z limit
        const int nf = n - f0 < 65535 ? n - f0 : 65535;
        const uint8_t *src = d_yuv + (size_t)f0 * yuv_stride;
        uint8_t *dst = d_bgr + (size_t)f0 * fstride;
        if (vec) {
            const int strips = W >> 4, bx = strips < 128 ? ((strips + 31) / 32) * 32 : 128;
            AVD_LAUNCH(ctx, "k0_yuv420_bgr", st, yuv420_bgr_vec_kernel<<<dim3((strips + bx - 1) / bx, H >> 1, nf), bx, 0, st>>>(src, yuv_stride, H, W, k, dst, pitch, fstride));
        } else {
            AVD_LAUNCH(ctx, "k0_yuv420_bgr", st, yuv420_bgr_any_kernel<<<dim3(((W >> 1) + 127) / 128, H >> 1, nf), 128, 0, st>>>(src, yuv_stride, H, W, k, dst, pitch, fstride));
        }
    }
    return AVD_OK;
}
