// conv_tc.cu — tcgen05 implicit-GEMM convolution (placeholder until the kernel lands).
#include "common.cuh"
int gnb_conv_tc_init(gnb_ctx* ctx) { GNB_SET_ERR(ctx, "tcgen05 conv not built"); return GNB_E_INVALID; }
int gnb_conv_tc_layer(gnb_ctx* ctx, const ConvLayer&, const bf16*, int, int, int, bf16*, float*, int, int) {
    GNB_SET_ERR(ctx, "tcgen05 conv not built");
    return GNB_E_INVALID;
}
