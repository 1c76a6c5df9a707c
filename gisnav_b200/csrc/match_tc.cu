// match_tc.cu — tcgen05 descriptor GEMM (placeholder until the kernel lands).
#include "common.cuh"
int gnb_match_tc_init(gnb_ctx* ctx) { GNB_SET_ERR(ctx, "tcgen05 matcher not built"); return GNB_E_INVALID; }
int gnb_match_tc_rowpass(gnb_ctx* ctx, int, int, int, int) { GNB_SET_ERR(ctx, "tcgen05 matcher not built"); return GNB_E_INVALID; }
