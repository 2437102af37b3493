// k_gemm_tc.cu -- tcgen05 engine (placeholder until the TMA/TMEM kernel lands): reports "not eligible".
#include "rau_model.cuh"
int tc_gemm_try(rau_ctx* ctx, const SimtGemm& g) { (void)ctx; (void)g; return 0; }
