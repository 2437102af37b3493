// rau_kernels.cuh -- host-callable wrappers of the pointwise / reduction kernels (k_pointwise.cu,
// k_attention.cu, k_optim.cu).  Every producer can emit a float32 tensor, a bf16 tensor or both:
// bf16 copies are the tcgen05 operands of the fast mode, float32 ones feed pointwise consumers.
#pragma once
#include "rau_common.cuh"

// ---- dropout masks (packed keep bits, bit i of word i>>5)
// nHop > 1: the masks of hops 0..nHop-1 in one launch (stream ids stream_id ^ hop, bit words hop_stride apart)
int k_mask_gen(rau_ctx* ctx, uint32_t* bits, int64_t n, float p, uint64_t seed, uint64_t stream_id, int nHop = 1,
               int64_t hop_stride = 0);
int k_mask_pack(rau_ctx* ctx, uint32_t* bits, const uint8_t* bytes, int64_t n);
int k_mask_unpack(rau_ctx* ctx, const uint32_t* bits, int64_t n, uint8_t* bytes);
// keep decisions of the rows pack kernels' inline 16-bit Philox draws as bytes [nHop, B, C, S]
int k_xmask16_bytes(rau_ctx* ctx, uint8_t* out, int B, int C, int S, int nHop, float p_drop, uint64_t stream_id);
static inline int64_t mask_words(int64_t n) { return (n + 31) / 32 + 4; }

// ---- word embedding (F:203-206)
int k_embed_fwd(rau_ctx* ctx, const float* ids, int n, int D, int V, const float* E,
                const uint32_t* bits, float scale, float* out_f, bf16* out_b, int ldb, bf16* out_lo = nullptr);   // out_b / out_lo: packed (hi, lo)
int k_embed_bwd(rau_ctx* ctx, const float* ids, int n, int D, int V, const float* out,
                const uint32_t* bits, float scale, const float* dout, int lddout, float* gE);

// ---- LSTM pointwise (A:12-25, D:47-61). saved = [5][B][H] planes i,f,o,g,tanh(c)
int k_lstm_fwd(rau_ctx* ctx, int B, int H, int order, const float* G, int ldg,
               const float* c_prev, int ldcp, float* c, int ldc, float* h, int ldh,
               bf16* h_b, int ldhb, float* saved);
// dc_out/dh_out may be NULL (zeros). inject: when lengths != NULL and lengths[b] == t the pair is
// REPLACED by dq_c/dq_h rows (F:604-610). dh_extra (NULL ok) is added to dh_out after injection.
int k_lstm_bwd(rau_ctx* ctx, int B, int H, int order,
               const float* dc_out, int lddc, const float* dh_out, int lddh,
               const float* dh_extra, int ldhe,
               const float* lengths, int t, const float* dq_c, const float* dq_h, int lddq,
               const float* c_prev, int ldcp, const float* saved,
               float* dG, bf16* dG_b, float* dc_prev, int lddcp, bf16* dG_lo = nullptr,   // dG_b / dG_lo: packed (hi, lo) copy
               float* zero_out = nullptr,    // optional [B, H] buffer cleared on the way (target of the next split-K dgrad)
               const uint32_t* extra_bits = nullptr, int64_t extra_bit0 = 0, float extra_scale = 1.0f);   // keep mask of dh_extra

// ---- generic elementwise
// y = x * keep(bits) * scale over a [rows, cols] matrix (mask indexed by row*cols+col); padded output pitch
int k_dropout(rau_ctx* ctx, const float* x, int64_t rows, int cols, int ldx, const uint32_t* bits, float scale,
              float* y_f, int ldyf, bf16* y_b, int ldyb, int cols_pad, bf16* y_lo = nullptr);   // y_b / y_lo: packed (hi, lo)
// y = x*keep*scale written only as packed bf16 (hi, lo) rows of pitch cols_pad: a tcgen05 operand (cols % 4 == 0)
int k_dropout_pack(rau_ctx* ctx, const float* x, int64_t rows, int cols, const uint32_t* bits, float scale, bf16* hi, bf16* lo,
                   int cols_pad, int64_t ldx = 0);   // ldx: row pitch of x (0 = cols)
// y (+)= x*keep*scale  (used for dq accumulation over hops and dX)
int k_dropout_bwd_acc(rau_ctx* ctx, const float* dx, int64_t n, const uint32_t* bits, float scale, float* y, int accumulate,
                      bf16* y_hi = nullptr, bf16* y_lo = nullptr);
// batched over the hops of the unroll (hop h's keep bits start bits_stride words after hop h-1's; bits == NULL = keep all)
int k_dropout_hops(rau_ctx* ctx, const float* x, int64_t n, int nHop, const uint32_t* bits, int64_t bits_stride, float scale,
                   float* y, bf16* y_hi, bf16* y_lo);
int k_dropout_bwd_hops(rau_ctx* ctx, float* y, int64_t n, int nHop, const uint32_t* bits, int64_t bits_stride, float scale,
                       bf16* y_hi, bf16* y_lo);
int k_dropout_bwd_sum_hops(rau_ctx* ctx, const float* dx, int64_t n, int nHop, const uint32_t* bits, int64_t bits_stride,
                           float scale, float* out);
int k_tanh_bwd(rau_ctx* ctx, const float* dy, const float* y, int64_t n, float* dx_f, bf16* dx_b, bf16* dx_lo = nullptr);
int k_add(rau_ctx* ctx, const float* a, const float* b, int64_t n, float* y);              // y = a + b
int k_axpy(rau_ctx* ctx, float alpha, const float* x, int64_t n, float* y);                // y += alpha x
int k_fill(rau_ctx* ctx, float* x, int64_t n, float v);
int k_l2_evict(rau_ctx* ctx, const void* buf, size_t bytes, float* sink);   // read `bytes` (> L2) so that L2 holds clean lines only
int k_to_bf16(rau_ctx* ctx, const float* x, int64_t rows, int cols, int ldx, bf16* y, int ldy, int cols_pad);
int k_rowdot_sigmoid(rau_ctx* ctx, const float* x, int B, int K, const float* w, const float* b, float* y);
// wd/bd gradient and dm contribution of the do_pred head (zero in training, kept for the module API)
int k_dopred_bwd(rau_ctx* ctx, const float* ddo, const float* dop, const float* m, const float* wd, int B, int K,
                 float* dm_acc, float* gwd, float* gbd);

// ---- reductions
// out[c] (+)= sum_r x[r*ld + c]
int k_colsum(rau_ctx* ctx, const float* x, int64_t rows, int cols, int ld, float* out, int accumulate, float* out2 = nullptr);
// out[m] += sum_b sum_s x[(b*M + m)*Sp + s], s < S   (bias grads of the 1x1 convolutions)
template <typename T>
int k_rowsum_bms(rau_ctx* ctx, const T* x, int B, int M, int S, int Sp, float* out);
int k_sum_all(rau_ctx* ctx, const float* x, int64_t n, float* out, int accumulate);

// ---- encoder glue
int k_select_state(rau_ctx* ctx, const float* S_all, int T, int B, int Q, const float* lengths, float* out);

// ---- attention proper (a7 tail, a8, a9) on saved projections, one CTA per image
template <typename T>
int k_attn_fwd(rau_ctx* ctx, int B, int M, int A, int S, int Sp, const T* E, const T* I,
               const float* ws, const float* mem, float* p, bf16* p_b, int ldpb, float* a, bf16* a_b);
template <typename T>
int k_attn_bwd(rau_ctx* ctx, int B, int M, int A, int S, int Sp, const T* E, const T* I,
               const float* ws, const float* p, const float* dp_in, const float* da,
               float* ds, bf16* ds_b, int lddsb, T* dZ, float* dqa, bf16* dqa_b, float* gws_part,
               bf16* dZ_hi = nullptr, bf16* dZ_lo = nullptr);   // dZ may be NULL when only the packed form is wanted
// dY = (dI + da[b,m] p[b,s]) * (1 - I^2), pad columns zeroed; dI may alias dY when T == float
template <typename T>
int k_iembed_bwd_pw(rau_ctx* ctx, int B, int M, int S, int Sp, const float* dI, const T* I,
                    const float* da, const float* p, T* dY);
// same, one warp per (image, channel) row: writes dY as fp32 (dY may be NULL) and/or packed bf16 (hi, lo) and adds
// the row sums into gbi[m] (the bias gradient of the 1x1 convolution)
int k_iembed_bwd_rows(rau_ctx* ctx, int B, int M, int S, int Sp, const float* dI, const float* I, const float* da,
                      const float* p, float* dY, bf16* dY_hi, bf16* dY_lo, float* gbi);

// ---- criteria (a12)
int k_softmax_ce(rau_ctx* ctx, int B, int N, const float* score, const float* labels, float loss_scale, float grad_scale,
                 float* loss_sum, float* dscore_f, bf16* dscore_b, int lddb, float* answers, bf16* dscore_lo = nullptr);
// logging-only merged predictions (F:539-574) over all hops; also writes uni/select when requested (predict)
int k_merge_preds(rau_ctx* ctx, int nHop, int B, int N, int S, const float* scores, const float* do_pred,
                  const float* attprob, const float* labels, const float* answers_hop, int force_last, float inv_bglobal,
                  float* loss_uni_sel /*[2]*/, float* loss_do_pred /*[nHop]*/, float* answers_uni_sel /*[2,B]*/,
                  float* pred_uni, float* pred_sel, float* att_uni, float* att_sel);

// open-ended / multiple-choice answers of `rows` = (nHop+2)*B prediction rows (F:903-918); mc [B, nmc] 1-based ids, 0 = empty
int k_answers(rau_ctx* ctx, int rows, int B, int N, const float* pred, const float* mc, int nmc, float* oe_out, float* mc_out);

// ---- noise / clip / optimizers (a13, a14)
int k_noise_norm(rau_ctx* ctx, float* g, int64_t n, float std, const float* noise_override,
                 uint64_t seed, uint64_t stream_id, double* norm2_out);
int k_clip_optim(rau_ctx* ctx, int optim, int64_t n, float* x, float* g, const double* norm2, float clip,
                 float lr, float h0, float h1, float h2, float* s0, float* s1, int64_t t, float* norm_out, int group = -1);
