// rau_rows.cuh -- host interface of the rows-layout tcgen05 engine (k_rows_tc.cu).
#pragma once
#include "rau_common.cuh"

enum { ROWS_EPI_PLAIN = 0, ROWS_EPI_RED = 1, ROWS_EPI_TANH = 2, ROWS_EPI_ATT = 3, ROWS_EPI_DY = 4, ROWS_EPI_LINEAR = 5, ROWS_EPI_LSTM = 6 };

// one operand: bf16 hi (and lo in bf16x3 mode); mn = 0: stored [rows, ld >= K] (K-major); mn = 1: stored [K, ld >= rows]
struct RowsOperand { const bf16* hi = nullptr; const bf16* lo = nullptr; int mn = 0; int64_t ld = 0; };

struct RowsGemm {
  int M = 0, N = 0, K = 0;          // D[M,N] = sum_k A[m,k] B[n,k]
  RowsOperand A, B;
  RowsOperand A2, B2; int K2 = 0;   // optional second K segment with the same major-ness (x Wi^T + h Wh^T)
  int epi = ROWS_EPI_PLAIN;
  int BN = 0;                       // accumulator columns per tile (64 / 128 / 256), 0 = chosen from the shape
  bf16 *out_hi = nullptr, *out_lo = nullptr;   // EPI_TANH / EPI_DY; optional extra outputs of EPI_LINEAR (pitch ldo_b)
  float* out_f = nullptr;                      // EPI_PLAIN / EPI_RED / EPI_ATT / EPI_LINEAR
  int64_t ldo = 0, ldo_b = 0;
  const float* bias = nullptr;      // [N]
  const float* bias2 = nullptr;     // [N]           (EPI_LINEAR)
  const float *addend = nullptr, *addend2 = nullptr; int64_t ldadd = 0;   // [M, ldadd] (EPI_LINEAR)
  int act = 0;                      // EPI_LINEAR: 0 none, 1 tanh, 2 sigmoid, 3 tanh BACKWARD: result * (1 - addend2^2)
  // EPI_LSTM (N = 4H permuted gate columns, see rows_pack_lstm): bias/addend as above, plus
  const float* c_prev = nullptr; int64_t ldcp = 0;
  float* c_out = nullptr; int64_t ldc = 0;
  float* h_out = nullptr; int64_t ldh = 0;
  float* lsaved = nullptr;          // [5][M][H] planes i, f, o, g, tanh(c)
  bf16 *hpk_hi = nullptr, *hpk_lo = nullptr; int64_t ldhp = 0;
  const float* rowvec = nullptr;    // [rows / S, N]
  const float* colw = nullptr;      // [N]
  float* rowout = nullptr;          // [M]
  const float* rowscale = nullptr;  // [M]
  const bf16 *aux_hi = nullptr, *aux_lo = nullptr; int64_t ldaux = 0;
  float* colsum = nullptr;          // [N]
  int S = 0;                        // rows per image
  float alpha = 1.0f;
  int f16 = 0;                      // the operands are single fp16 planes (lo must be NULL)
  int of16 = 0;                     // EPI_TANH / EPI_DY: the output is one fp16 plane
  int af16 = 0;                     // EPI_DY: the saved activation (aux_hi) is one fp16 plane
  float gscale = 1.0f;              // EPI_DY: power-of-two scale carried by A (= dZ) and the output dY; colsum gets alpha (= 1 / gscale)
};

bool rows_path_enabled();
int rows_gemm(rau_ctx* ctx, const RowsGemm& g);
struct SimtGemm;
// nn.Linear-shaped products (forward, dgrad, wgrad) described as a SimtGemm: returns 1 when the rows engine ran it,
// 0 when the description does not fit (the caller falls back to the generic engines), < 0 on error
int rows_contract_try(rau_ctx* ctx, const SimtGemm& g);
// fp32 -> bf16 (hi [, lo]) copy of n contiguous elements in an arena buffer; cache: parameter tensor, packed once per epoch
// fp32 [rows, cols] of pitch ld -> packed bf16 (hi [, lo]) of pitch *ldo (cols rounded up to 8) in an arena buffer
int rows_pack2d(rau_ctx* ctx, const float* src, int64_t ld, int rows, int cols, bool want_lo, bool is_const, const char* slot,
                const bf16** hi, const bf16** lo, int64_t* ldo);
// LSTM weights [4H, K] with rows permuted for the fused cell epilogue (cached per epoch), and the permuted bias sum
int rows_pack_lstm(rau_ctx* ctx, const float* W, int H, int K, int gate_order, bool want_lo, const bf16** hi, const bf16** lo,
                   int64_t* ldo);
int rows_perm_lstm_bias(rau_ctx* ctx, const float* b1, const float* b2, int H, int gate_order, const float** out);
// same, into a caller-owned packed twin (pitch ldo >= cols, zero padded)
// wcols > 0: write only that many columns per row (the destination is a column block of a wider packed matrix)
int rows_pack_into(rau_ctx* ctx, const float* src, int64_t ld, int rows, int cols, bf16* hi, bf16* lo, int64_t ldo, int wcols = 0);
int rows_pack(rau_ctx* ctx, const float* W, int64_t n, bool want_lo, bool cache, const char* slot, const bf16** hi, const bf16** lo,
              bool f16 = false);   // f16: one fp16 plane instead of bf16 hi [, lo]
// gen != 0: draw the keep bits inline from Philox stream `stream_id` with drop rate p_drop (bits is then ignored)
// the same for nHop hops in one launch (training step, drawn masks): hop h writes hi/lo + h*hop_stride, stream id ^ h
int k_xprep_rows_hops(rau_ctx* ctx, const float* X, const void* X16, int B, int C, int S, int nHop, float scale, bf16* hi, bf16* lo,
                      int64_t hop_stride, float p_drop, uint64_t stream_id, int f16 = 0);
int k_unpack_hilo(rau_ctx* ctx, const bf16* hi, const bf16* lo, int64_t n, float* out, int f16 = 0);   // out = hi + lo (tests)
int k_xprep_rows(rau_ctx* ctx, const float* X, int B, int C, int S, const uint32_t* bits, float scale, bf16* hi, bf16* lo,
                 int gen = 0, float p_drop = 0.0f, uint64_t stream_id = 0, int f16 = 0, int hop = -1, int nHop = 0);
// (gen: hop >= 0 and nHop tell the kernel which hop of how many this pack belongs to -- stream_id is then base ^ hop -- so that
// it draws the same keep bits as the all-hops launch, see rau_xmask_shared)
int k_unprep_rows(rau_ctx* ctx, const float* dXr, int B, int C, int S, const uint32_t* bits, float scale, float* dX);
int k_attn_rows_fwd(rau_ctx* ctx, int B, int M, int S, const float* logit, const float* mem, const bf16* I_hi, const bf16* I_lo,
                    float* p, float* a, bf16* p_hi = nullptr, bf16* p_lo = nullptr, int ldp = 0, int f16 = 0);   // p_hi/p_lo: packed twin of p
// one encoder LSTM layer over all T steps in a single persistent launch (k_rows_tc.cu lstm_seq_kernel)
struct LstmSeq {
  int B = 0, H = 0, T = 0;
  const bf16* Wh_hi = nullptr; const bf16* Wh_lo = nullptr; int64_t ldwh = 0;   // rows_pack_lstm layout [4H, H]
  const float* Gx = nullptr; int64_t gx_t = 0; int ldg = 0;                    // hoisted input projection [T][B][4H]
  float* c_out = nullptr; float* h_out = nullptr; int64_t s_t = 0; int lds = 0;   // state rows of step 1, step stride, pitch
  float* lsaved = nullptr; int64_t ls_t = 0, plane = 0;                         // saved gates of step 1
  bf16* hpk_hi = nullptr; bf16* hpk_lo = nullptr;                               // packed h stack [(T+1)][B][H], step 0 = zeros
  // optional: the state of step t == lengths[b] also goes to sel_c / sel_h [B, H] (row pitch sel_ld): the encoder's
  // length selection (F:472-478) without a launch of its own.  Rows whose length is outside 1..T are left untouched.
  const float* lengths = nullptr; float* sel_c = nullptr; float* sel_h = nullptr; int sel_ld = 0;
  unsigned int* counter = nullptr;   // optional: 64 zeroed words for the row tiles' step counters (else cleared here)
};
int rows_lstm_seq(rau_ctx* ctx, const LstmSeq& d, int* done);

// logit[r] = ws . tanh(Z[r,:] + qadd[b(r),:])  (Z = I Wa^T precomputed; qadd = Wqa qf + bqa + ba)
int k_attn_rows_score(rau_ctx* ctx, int B, int A, int S, const float* Z, const float* qadd, const float* ws, int fast_tanh,
                      float* logit);
int k_attn_rows_bwd(rau_ctx* ctx, int B, int M, int A, int S, const float* E, const bf16* I_hi, const bf16* I_lo, const float* ws,
                    const float* p, const float* dp_in, const float* da, float* ds, bf16* dZ_hi, bf16* dZ_lo, float* dqa,
                    float* gws_part, bf16* ds_hi = nullptr, bf16* ds_lo = nullptr, int ldds = 0,
                    const float* qadd = nullptr, int fast_tanh = 0, int acc_zeroed = 0, int f16 = 0, float gscale = 1.0f);
