// rau_model.cuh -- internal interfaces between the orchestration files.
#pragma once
#include "rau_kernels.cuh"
#include "rau_layout.cuh"

// contraction front-end: routes a SimtGemm description to the tcgen05 engine (k_gemm_tc.cu) when the
// ctx precision mode and the operand shapes allow it, else to the fp32 CUDA-core engine (k_gemm_simt.cu).
int rau_contract(rau_ctx* ctx, const SimtGemm& g);
// k_gemm_tc.cu: returns 1 when it ran the product on tcgen05, 0 when the shape is not eligible, <0 on error
int tc_gemm_try(rau_ctx* ctx, const SimtGemm& g);

// nn.Linear helpers over row-major activations [rows, *] and weights [N, K]
static inline SimtGemm lin_fwd(int rows, int N, int K, const float* X, int ldx, const float* W, float* Y, int ldy) {
  SimtGemm g;
  g.M = rows; g.N = N; g.K = K;
  g.A = X; g.sam = ldx; g.sak = 1;
  g.B = W; g.sbk = 1; g.sbn = K;
  g.C = Y; g.scm = ldy; g.scn = 1;
  g.b_const = 1;
  return g;
}
static inline void lin_seg2(SimtGemm& g, int K2, const float* X2, int ldx2, const float* W2) {
  g.A2 = X2; g.sam2 = ldx2; g.sak2 = 1;
  g.B2 = W2; g.sbk2 = 1; g.sbn2 = K2; g.K2 = K2;
}
// dX[rows,K] = dY[rows,N] W[N,K]
static inline SimtGemm lin_dgrad(int rows, int N, int K, const float* dY, int lddy, const float* W, float* dX, int lddx) {
  SimtGemm g;
  g.M = rows; g.N = K; g.K = N;
  g.A = dY; g.sam = lddy; g.sak = 1;
  g.B = W; g.sbk = K; g.sbn = 1;
  g.C = dX; g.scm = lddx; g.scn = 1;
  g.b_const = 1;
  return g;
}
// gW[N,K] += scale * dY[rows,N]^T X[rows,K]
static inline SimtGemm lin_wgrad(int rows, int N, int K, const float* dY, int lddy, const float* X, int ldx, float* gW, float scale) {
  SimtGemm g;
  g.M = N; g.N = K; g.K = rows;
  g.A = dY; g.sam = 1; g.sak = lddy;
  g.B = X; g.sbk = ldx; g.sbn = 1;
  g.C = gW; g.scm = K; g.scn = 1;
  g.accumulate = 1;
  g.alpha = scale;
  return g;
}

// packed bf16 (hi, lo) twin of an fp32 activation [rows, ld]: the tcgen05 operand form, written by the producing kernel
struct PK { bf16* hi = nullptr; bf16* lo = nullptr; int64_t ld = 0; };

// activations one answering unit keeps for its backward pass (carved out of one caller- or arena-owned block)
struct HopSaved {
  uint32_t *qbits, *xbits, *mbits;   // packed keep bits of the three dropouts (F:233, F:239, F:277)
  float *qd, *qf, *I, *E, *p, *j, *lsav, *hout, *m, *dop;
  float* qatt;   // rows path: Wqa qf + bqa + ba [B, A]; E then holds Z = I Wa^T and tanh(Z + qatt[b]) is redone where it is read
  bf16 *Xd_hi, *Xd_lo, *I_hi, *I_lo;   // tcgen05 modes: packed operands kept for the backward pass
  // training step on the rows engine: the feature-dropout bits are drawn inside the transposing pack kernel from this
  // Philox stream instead of being materialised in xbits first (nothing in the step reads them again: dX is not formed)
  int x_philox = 0;
  int x_done = 0;        // the pack already ran (all hops in one launch)
  uint64_t x_stream = 0;
  int x_hop = -1, x_nhop = 0;   // which hop of how many (the shared draw of p = 1/2 needs both; -1: a lone pack)
  // training step: packed twins of the small activations (all NULL through the module-level API, which packs on demand)
  PK qd_pk, qf_pk, p_pk, j_pk, hin_pk, hout_pk, m_pk;
  // training step: Wq drop_h(q) + bq of this hop, computed for all hops in one product before the unroll (q is the same
  // encoder state for every hop); hop_forward then skips the q dropout and the Wq segment
  const float* qpre = nullptr;
};
bool hop_rows_path(const rau_ctx* ctx, const rau_config* cfg);

// Cross-stream scheduling of one hop inside the training step (rows path only).  The i_embed product (feature pack +
// tanh(Wi X + bi)) does not depend on the recurrent state, and the three heavy backward products (dY, gWa, gWi) feed
// nothing the previous hop's backward needs: both run on ctx->side while the chain of small kernels advances.
struct HopAsync {
  cudaEvent_t pre_done = nullptr;    // forward: I of this hop is ready (recorded on the side stream); NULL = compute inline
  int pre_skip = 0;                  // forward: Xd / I / Z in the saved block are already this hop's (eval mode: hop-invariant)
  int bwd_side = 0;                  // backward: put dY / gWa / gWi on the side stream
  int head_side = 0;                 // forward: the answer head (Wo, dropout, score, do_pred) does not feed the next hop: side stream
  int hop = 0;                       // selects the per-hop dZ buffer when bwd_side
};
int hop_forward_pre(rau_ctx* ctx, const rau_config* cfg, int B, const MultT<const float*>& P, const float* X, int train,
                    const HopSaved& sv);
cudaEvent_t rau_side_event(rau_ctx* ctx);
size_t hop_saved_layout(const rau_config* cfg, int B, void* base, HopSaved* sv);

// Backward scratch of one hop that the weight-gradient products read.  In the training step these point into
// tensor-major stacks [nHop][B][dim], hop_backward skips every nn.Linear accGradParameters, and hop_wgrads() issues
// each of them ONCE over all nHop*B rows after the last hop (the clones share one gradWeight, F:344, so the sum over
// hops is what the reference accumulates anyway).
struct HopGrads {
  float *du, *dG, *dj, *ds, *dqa, *dpre, *gwsp;
  PK dscore_pk, du_pk, dG_pk, ds_pk, dpre_pk;   // packed twins written by the producing kernels (may be NULL)
  // training step: the answer head's backward depends on forward results only, so du = drop'(Ws^T dscore) and
  // dh2h = Wo^T du are computed for all hops in two products before the backward unroll; likewise dq is formed once
  // after it from the stacked dpre.  Non-NULL dh2h switches hop_backward to that form.
  const float* dh2h = nullptr;
  int acc_zeroed = 0;   // dqa / gwsp (atomic accumulators of the attention backward) were cleared for all hops at once
};
struct HopStacks {   // every member is [nHop][B][dim]
  const float *dscore, *m, *du, *hout, *dG, *j, *h_in, *dj, *p, *ds, *dqa, *qf, *dpre, *qd, *gwsp;
  PK dscore_pk, m_pk, du_pk, hout_pk, dG_pk, j_pk, hin_pk, p_pk, ds_pk, qf_pk, dpre_pk, qd_pk;   // packed stacks (may be NULL)
};
int hop_wgrads(rau_ctx* ctx, const rau_config* cfg, int rows, const MultT<float*>& G, const HopStacks& st);

int hop_forward(rau_ctx* ctx, const rau_config* cfg, int B, const MultT<const float*>& P,
                const float* q, const float* X, const float* c, const float* h, int train, const HopSaved& sv,
                float* score, float* do_pred, float* p_out, float* c_out, float* h_out, const HopAsync* as = nullptr);
int hop_backward(rau_ctx* ctx, const rau_config* cfg, int B, const MultT<const float*>& P, const MultT<float*>& G,
                 const float* X, const float* c, const float* h, int train, const HopSaved& sv,
                 const float* dscore, const float* ddo_pred, const float* dp_att, const float* dc_out, const float* dh_out,
                 float* dq, int dq_accumulate, float* dX, float* dc, float* dh, const HopGrads* deferred = nullptr,
                 const HopAsync* as = nullptr);
