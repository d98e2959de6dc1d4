// K1 fused path placeholder (filled in next): see k1_generic.cu for the arithmetic.
#include "avd_internal.h"
bool k1_fused_supported(int, int, size_t, size_t, const void *) { return false; }
int k1_fused(avd_ctx *ctx, const uint8_t *, int, int, int, size_t, size_t, const ShapeTables *, const ShapeTables *,
             avd_sample_rec *, uint8_t *, uint8_t *, cudaStream_t)
{
    ctx->err = "fused K1 not built";
    return AVD_ERR_ARG;
}
