// rau_step.cu -- the training step of the experiment scripts as device work enqueued on one stream:
//   rau_feval      = feval (F:445-615): encoder unroll, nHop answering units, joint loss, BPTT
//   rau_noise_clip = gradient noise + per-group clip (F:617-648)
//   rau_optim_step = utils/optim_updates.lua (OU:7-87)
//   rau_train_step = feval -> [all-reduce] -> noise/clip -> optimizer x3 (F:786-791)
//   rau_predict    = predict_result (F:652-724)
// Differences from the reference's schedule that do not change results: the layer-1 and layer-2 input
// projections of the encoder are hoisted out of the time loop (one [T*B,in]x[in,4H] product each), the
// per-row "select the state at t == len" host loops (F:472-478, F:604-610) are device kernels, the encoder's
// weight gradients are single contractions over all T*B rows, and dX of the image features is never formed.
#include "rau_model.cuh"
#include "rau_rows.cuh"
#include <math.h>

int rau_check_cfg(const rau_config* cfg);
int rau_check_dev(const void* p, const char* what);
int rau_prepare_mask(rau_ctx* ctx, uint32_t* bits, int64_t n, float p, int train, const uint8_t* bytes, uint64_t stream_id);
int rau_allreduce_internal(rau_ctx* ctx, float* buf, int64_t n);   // rau_comm.cu
int rau_allreduce_group(rau_ctx* ctx, int begin);
bool rau_comm_attached(rau_ctx* ctx);
int rau_comm_rank(rau_ctx* ctx);
int rau_check_async_error(rau_ctx* ctx);   // rau_ctx.cu

#define ARENA(ptr, type, name, count) \
  type* ptr = nullptr;                \
  RAU_TRY(ctx->arena.get(name, sizeof(type) * (size_t)(count), (void**)&ptr))

static inline float drop_scale(float p) { return p > 0.0f ? 1.0f / (1.0f - p) : 1.0f; }

// Philox stream ids: [step:40][kind:8][index:16]; dropout streams are offset per rank, noise streams are not
static inline uint64_t stream_of(int64_t step, int kind, int idx, int rank) {
  return ((uint64_t)step << 24) ^ ((uint64_t)kind << 16) ^ (uint64_t)idx ^ ((uint64_t)rank << 56);
}
enum { SK_EMBED = 1, SK_RNN = 2, SK_Q = 3, SK_X = 4, SK_M = 5, SK_NOISE = 9 };

struct Encoder {
  int B, Tm;          // batch, number of unrolled steps
  uint32_t *ebits, *rbits;
  float *e_all, *G1x, *G2x, *u2, *S_all, *sav1, *sav2, *Gt, *rnn_out;
  int proj0_done = 0;   // a part-1 call already formed layer 1's hoisted input projection
  cudaEvent_t w_ready = nullptr;   // the encoder's weight shadows were packed on the aux stream: wait before the first product
  cudaEvent_t w_ready0 = nullptr;  // ... the part of them layer 1's input projection needs (recorded first)
  int h0_zeroed = 0;               // ... and the packed h_0 rows of both layers were cleared there
  int rbits_done = 0;              // ... and the recurrent dropout mask was prepared there
  unsigned int* seq_cnt = nullptr; // ... and 2 x 64 zeroed step counters for the persistent recurrences, with rnn_out cleared:
                                   //     the recurrence kernels write the length-selected rows themselves (no select launch)
};

// the question encoder runs on the tcgen05 rows engine (hoisted input projections + persistent recurrence)
static bool encoder_fused(const rau_ctx* ctx, const rau_config* cfg, int B) {
  return ctx->precision != RAU_PREC_F32 && rows_path_enabled() && cfg->Hq % 8 == 0 && cfg->embed % 8 == 0 &&
         (int64_t)B * 4 * cfg->Hq * cfg->Hq >= (1 << 18);
}

static int encoder_alloc(rau_ctx* ctx, const rau_config* cfg, int B, Encoder* en) {
  const int T = cfg->T, Hq = cfg->Hq, E = cfg->embed, Q = 4 * Hq;
  ARENA(ebits, uint32_t, "enc.ebits", mask_words((int64_t)T * B * E));
  ARENA(rbits, uint32_t, "enc.rbits", mask_words((int64_t)T * B * Hq));
  ARENA(e_all, float, "enc.e", (size_t)T * B * E);
  ARENA(G1x, float, "enc.G1x", (size_t)T * B * 4 * Hq);
  ARENA(G2x, float, "enc.G2x", (size_t)T * B * 4 * Hq);
  ARENA(u2, float, "enc.u2", (size_t)T * B * Hq);
  ARENA(S_all, float, "enc.S", (size_t)(T + 1) * B * Q);
  ARENA(sav1, float, "enc.sav1", (size_t)T * 5 * B * Hq);
  ARENA(sav2, float, "enc.sav2", (size_t)T * 5 * B * Hq);
  ARENA(Gt, float, "enc.G", (size_t)B * 4 * Hq);
  ARENA(rnn_out, float, "enc.out", (size_t)B * Q);
  en->B = B; en->ebits = ebits; en->rbits = rbits; en->e_all = e_all; en->G1x = G1x; en->G2x = G2x; en->u2 = u2;
  en->S_all = S_all; en->sav1 = sav1; en->sav2 = sav2; en->Gt = Gt; en->rnn_out = rnn_out;
  return RAU_OK;
}

// The bf16 (hi, lo) shadows of the encoder's weights in the permuted gate order and the permuted bias sums: six small
// launches that depend on the parameters only.  The training step issues them on the aux stream at its very start (they
// are cached per public call, so the encoder's own calls below find them done) instead of on the chain in front of each
// layer's products.
// first != 0: only what layer 1's hoisted input projection needs (its Wi shadow and bias sum); first == 0: everything else
static int encoder_pack_weights(rau_ctx* ctx, const rau_config* cfg, const float* Pr, int B, int first) {
  RnnLayerOff L[4];
  rnn_offsets(cfg, L);
  const bool x3 = prec_x3(ctx);
  const int Hq = cfg->Hq;
  for (int layer = 0; layer < 2; ++layer) {
    const int in = layer == 0 ? cfg->embed : Hq;
    const bf16 *h, *l;
    int64_t ld;
    const float* bperm;
    if (first || layer == 1) {   // (cached per public call: the second pass does not repeat layer 1's)
      RAU_TRY(rows_pack_lstm(ctx, Pr + L[layer].Wi, Hq, in, RAU_GATES_IFOG, x3, &h, &l, &ld));
      RAU_TRY(rows_perm_lstm_bias(ctx, Pr + L[layer].bi, Pr + L[layer].bh, Hq, RAU_GATES_IFOG, &bperm));
    }
    if (first) return RAU_OK;
    RAU_TRY(rows_pack_lstm(ctx, Pr + L[layer].Wh, Hq, Hq, RAU_GATES_IFOG, x3, &h, &l, &ld));
  }
  // ... and the packed h_0 = 0 rows the two recurrences start from (Encoder::h0_zeroed)
  const size_t hb = (size_t)B * Hq;
  ARENA(hpk_all, bf16, "enc.hpk", (size_t)4 * (cfg->T + 1) * hb);   // [layer][hi, lo][T+1][B][Hq]
  for (int k = 0; k < 4; ++k)
    if (x3 || (k & 1) == 0) RAU_CHECK_CUDA(cudaMemsetAsync(hpk_all + (size_t)k * (cfg->T + 1) * hb, 0, hb * sizeof(bf16), ctx->stream));
  return RAU_OK;
}

// F:460-479.  State rows are [c1|h1|c2|h2] (D:23-24, D:68); S_all[t] is the state after step t, S_all[0] = 0.
// part: 0 = everything, 1 = masks + word embedding + (tcgen05 path) layer 1's hoisted input projection, 2 = the rest (after a
// part-1 call).  The training step releases the side stream between the two parts: its all-hops feature pack keeps 84 SMs
// and much of HBM busy for ~370 us, and the projection -- the last launch of the chain that is NOT latency-bound -- ran
// three times slower next to it.
static int encoder_forward(rau_ctx* ctx, const rau_config* cfg, const rau_batch* bt, const float* Pe, const float* Pr,
                           int train, const rau_masks* masks, int64_t step_t, Encoder* en, int part = 0) {
  const int B = bt->B, Hq = cfg->Hq, E = cfg->embed, Q = 4 * Hq, G4 = 4 * Hq;
  const int Tm = (bt->max_len > 0 && bt->max_len <= cfg->T) ? bt->max_len : cfg->T;
  en->Tm = Tm;
  RnnLayerOff L[4];
  rnn_offsets(cfg, L);
  const int rank = rau_comm_rank(ctx);
  const bool de = train && cfg->p_embed > 0, dr = train && cfg->p_rnn > 0;
  const bool fused = encoder_fused(ctx, cfg, B);
  // tcgen05 path: the packed (hi, lo) inputs of the two hoisted projections -- e for layer 1, u2 = drop(h1) for layer 2 -- are
  // written by their producers (the embedding lookup, the dropout) into the buffers the backward pass's weight gradients
  // read ("rp.enc.x0" / "rp.enc.x1"), not by a pack launch in front of each projection
  auto x_buf = [&](int layer, bf16** x_h, bf16** x_l, int64_t* ldx) -> int {
    const int in = layer == 0 ? E : Hq;
    *ldx = (in + 7) / 8 * 8;
    const size_t half = ((size_t)Tm * B * (size_t)*ldx * sizeof(bf16) + 1023) / 1024 * 1024;
    char* buf = nullptr;
    RAU_TRY(ctx->arena.get(layer == 0 ? "rp.enc.x0" : "rp.enc.x1", half * (prec_x3(ctx) ? 2 : 1), (void**)&buf));
    *x_h = (bf16*)buf;
    *x_l = prec_x3(ctx) ? (bf16*)(buf + half) : nullptr;
    return RAU_OK;
  };
  if (part != 2) {
    RAU_TRY(rau_prepare_mask(ctx, en->ebits, (int64_t)Tm * B * E, cfg->p_embed, train, masks ? masks->embed : nullptr,
                             stream_of(step_t, SK_EMBED, 0, rank)));
    if (!en->rbits_done)
      RAU_TRY(rau_prepare_mask(ctx, en->rbits, (int64_t)Tm * B * Hq, cfg->p_rnn, train, masks ? masks->rnn : nullptr,
                               stream_of(step_t, SK_RNN, 0, rank)));
    // word_embed for every step at once (F:203-206, F:468)
    bf16 *e_h = nullptr, *e_l = nullptr;
    int64_t lde = 0;
    if (fused) RAU_TRY(x_buf(0, &e_h, &e_l, &lde));
    RAU_TRY(k_embed_fwd(ctx, bt->tokens, Tm * B, E, cfg->V, Pe, de ? en->ebits : nullptr, drop_scale(cfg->p_embed),
                        en->e_all, e_h, (int)lde, e_l));
    if (ctx->phases == 2) rau_phase_mark(ctx, "enc embed done");
  }
  if (fused && en->w_ready0 != nullptr && part != 2)
    RAU_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, ctx->tune.enc_w0 ? en->w_ready0 : en->w_ready, 0));
  if (fused && en->w_ready != nullptr && part != 1) RAU_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, en->w_ready, 0));
  // hoisted input projection of one layer for every step at once, columns in the permuted gate order (tcgen05 path)
  auto input_projection = [&](int layer) -> int {
    const bool x3 = prec_x3(ctx);
    const int in = layer == 0 ? E : Hq;
    const bf16 *Wi_h, *Wi_l;
    bf16 *x_h, *x_l;
    int64_t ldwi, ldx;
    const float* bperm;
    RAU_TRY(rows_pack_lstm(ctx, Pr + L[layer].Wi, Hq, in, RAU_GATES_IFOG, x3, &Wi_h, &Wi_l, &ldwi));
    RAU_TRY(rows_perm_lstm_bias(ctx, Pr + L[layer].bi, Pr + L[layer].bh, Hq, RAU_GATES_IFOG, &bperm));
    RAU_TRY(x_buf(layer, &x_h, &x_l, &ldx));   // (written by k_embed_fwd / k_dropout_pack)
    RowsGemm g;
    g.M = Tm * B; g.N = G4; g.K = in;
    g.A.hi = x_h; g.A.lo = x_l; g.A.ld = ldx;
    g.B.hi = Wi_h; g.B.lo = Wi_l; g.B.ld = ldwi;
    g.epi = ROWS_EPI_LINEAR; g.bias = bperm; g.out_f = layer == 0 ? en->G1x : en->G2x; g.ldo = G4;
    return rows_gemm(ctx, g);
  };
  if (part == 1) {
    en->proj0_done = 0;
    if (fused) {
      const int cap = ctx->main_cta_cap;   // (the side stream is not running yet: the whole device)
      ctx->main_cta_cap = 0;
      const int rc = input_projection(0);
      ctx->main_cta_cap = cap;
      RAU_TRY(rc);
      en->proj0_done = 1;
    }
    return RAU_OK;
  }
  if (part == 0) en->proj0_done = 0;
  // layer 1 input projection hoisted over time: G1x = e Wi1^T + bi1 + bh1
  if (!fused) {
    SimtGemm g = lin_fwd(Tm * B, G4, E, en->e_all, E, Pr + L[0].Wi, en->G1x, G4);
    g.bias_n = Pr + L[0].bi; g.bias_n2 = Pr + L[0].bh;
    RAU_TRY(rau_contract(ctx, g));
  }
  RAU_TRY(k_fill(ctx, en->S_all, (int64_t)B * Q, 0.0f));
  if (fused) {
    // tcgen05 path: each recurrent step is ONE launch -- gate product h_{t-1} Wh^T (+ hoisted input projection) with the
    // cell update fused behind it in the epilogue (EPI_LSTM), which also emits h_t packed for the next step
    const bool x3 = prec_x3(ctx);
    const size_t hb = (size_t)B * Hq;
    ARENA(hpk_all, bf16, "enc.hpk", (size_t)4 * (cfg->T + 1) * hb);   // [layer][hi, lo][T+1][B][Hq], kept for the backward pass
    bool sel_fused = true;
    for (int layer = 0; layer < 2; ++layer) {
      bf16* hpk_hi = hpk_all + (size_t)(2 * layer) * (cfg->T + 1) * hb;
      bf16* hpk_lo = hpk_hi + (size_t)(cfg->T + 1) * hb;
      const bf16 *Wh_h, *Wh_l;
      int64_t ldwh;
      RAU_TRY(rows_pack_lstm(ctx, Pr + L[layer].Wh, Hq, Hq, RAU_GATES_IFOG, x3, &Wh_h, &Wh_l, &ldwh));
      float* Gx = layer == 0 ? en->G1x : en->G2x;
      if (layer == 1) {   // u2 = drop(h1) for every step (D:38-39), straight into the packed operand of the projection
        bf16 *u_h, *u_l;
        int64_t ldu;
        RAU_TRY(x_buf(1, &u_h, &u_l, &ldu));
        RAU_TRY(k_dropout_pack(ctx, en->S_all + (size_t)B * Q + Hq, (int64_t)Tm * B, Hq, dr ? en->rbits : nullptr,
                               drop_scale(cfg->p_rnn), u_h, u_l, (int)ldu, Q));
      }
      if (!(layer == 0 && en->proj0_done)) RAU_TRY(input_projection(layer));
      if (!en->h0_zeroed) {
        RAU_CHECK_CUDA(cudaMemsetAsync(hpk_hi, 0, hb * sizeof(bf16), ctx->stream));
        if (x3) RAU_CHECK_CUDA(cudaMemsetAsync(hpk_lo, 0, hb * sizeof(bf16), ctx->stream));
      }
      if (ctx->phases == 2) rau_phase_mark(ctx, "enc layer input projection done");
      int seq_done = 0;
      {   // the whole recurrence in one persistent launch when the layer fits (weights resident in shared memory)
        LstmSeq d;
        d.B = B; d.H = Hq; d.T = Tm;
        if (en->seq_cnt) {   // fused length selection into rnn_out [B, Q] = [c1 | h1 | c2 | h2] (cleared on the aux stream)
          d.counter = en->seq_cnt + 64 * layer;
          d.lengths = bt->lengths; d.sel_c = en->rnn_out + 2 * layer * Hq; d.sel_h = d.sel_c + Hq; d.sel_ld = Q;
        }
        d.Wh_hi = Wh_h; d.Wh_lo = Wh_l; d.ldwh = ldwh;
        d.Gx = Gx; d.gx_t = (int64_t)B * G4; d.ldg = G4;
        d.c_out = en->S_all + (size_t)B * Q + 2 * layer * Hq; d.h_out = d.c_out + Hq; d.s_t = (int64_t)B * Q; d.lds = Q;
        d.lsaved = layer == 0 ? en->sav1 : en->sav2; d.ls_t = (int64_t)5 * B * Hq; d.plane = (int64_t)B * Hq;
        d.hpk_hi = hpk_hi; d.hpk_lo = x3 ? hpk_lo : nullptr;
        RAU_TRY(rows_lstm_seq(ctx, d, &seq_done));
        if (ctx->phases == 2) rau_phase_mark(ctx, "enc layer recurrence done");
        if (!(seq_done && d.lengths)) sel_fused = false;
      }
      for (int t = 1; t <= Tm && !seq_done; ++t) {
        float* Sp_ = en->S_all + (size_t)(t - 1) * B * Q + 2 * layer * Hq;
        float* Sn = en->S_all + (size_t)t * B * Q + 2 * layer * Hq;
        RowsGemm g;
        g.M = B; g.N = G4; g.K = Hq;
        g.A.hi = hpk_hi + (size_t)(t - 1) * hb; g.A.lo = x3 ? hpk_lo + (size_t)(t - 1) * hb : nullptr; g.A.ld = Hq;
        g.B.hi = Wh_h; g.B.lo = Wh_l; g.B.ld = ldwh;
        g.epi = ROWS_EPI_LSTM;
        g.addend = Gx + (size_t)(t - 1) * B * G4; g.ldadd = G4;
        g.c_prev = Sp_; g.ldcp = Q; g.c_out = Sn; g.ldc = Q; g.h_out = Sn + Hq; g.ldh = Q;
        g.lsaved = (layer == 0 ? en->sav1 : en->sav2) + (size_t)(t - 1) * 5 * B * Hq;
        g.hpk_hi = hpk_hi + (size_t)t * hb; g.hpk_lo = x3 ? hpk_lo + (size_t)t * hb : nullptr; g.ldhp = Hq;
        RAU_TRY(rows_gemm(ctx, g));
      }
    }
    if (!sel_fused) RAU_TRY(k_select_state(ctx, en->S_all, Tm, B, Q, bt->lengths, en->rnn_out));
    return RAU_OK;
  }
  for (int t = 1; t <= Tm; ++t) {   // layer 1 recurrence
    float* Sp_ = en->S_all + (size_t)(t - 1) * B * Q;
    float* Sn = en->S_all + (size_t)t * B * Q;
    SimtGemm g = lin_fwd(B, G4, Hq, Sp_ + Hq, Q, Pr + L[0].Wh, en->Gt, G4);
    g.addend = en->G1x + (size_t)(t - 1) * B * G4; g.sdm = G4; g.sdn = 1;
    RAU_TRY(rau_contract(ctx, g));
    RAU_TRY(k_lstm_fwd(ctx, B, Hq, RAU_GATES_IFOG, en->Gt, G4, Sp_, Q, Sn, Q, Sn + Hq, Q, nullptr, 0,
                       en->sav1 + (size_t)(t - 1) * 5 * B * Hq));
  }
  // u2 = drop(h1) for every step (D:38-39), then the layer-2 input projection hoisted over time
  RAU_TRY(k_dropout(ctx, en->S_all + (size_t)B * Q + Hq, (int64_t)Tm * B, Hq, Q, dr ? en->rbits : nullptr,
                    drop_scale(cfg->p_rnn), en->u2, Hq, nullptr, 0, Hq));
  {
    SimtGemm g = lin_fwd(Tm * B, G4, Hq, en->u2, Hq, Pr + L[1].Wi, en->G2x, G4);
    g.bias_n = Pr + L[1].bi; g.bias_n2 = Pr + L[1].bh;
    RAU_TRY(rau_contract(ctx, g));
  }
  for (int t = 1; t <= Tm; ++t) {   // layer 2 recurrence
    float* Sp_ = en->S_all + (size_t)(t - 1) * B * Q;
    float* Sn = en->S_all + (size_t)t * B * Q;
    SimtGemm g = lin_fwd(B, G4, Hq, Sp_ + 3 * Hq, Q, Pr + L[1].Wh, en->Gt, G4);
    g.addend = en->G2x + (size_t)(t - 1) * B * G4; g.sdm = G4; g.sdn = 1;
    RAU_TRY(rau_contract(ctx, g));
    RAU_TRY(k_lstm_fwd(ctx, B, Hq, RAU_GATES_IFOG, en->Gt, G4, Sp_ + 2 * Hq, Q, Sn + 2 * Hq, Q, Sn + 3 * Hq, Q, nullptr, 0,
                       en->sav2 + (size_t)(t - 1) * 5 * B * Hq));
  }
  // rnn_out[k] = state at t == x_len[k] (F:472-478)
  RAU_TRY(k_select_state(ctx, en->S_all, Tm, B, Q, bt->lengths, en->rnn_out));
  return RAU_OK;
}

// F:600-615 with dq = sum over hops of the answering units' gradient w.r.t. rnn_out (branch:backward, F:598)
static int encoder_backward(rau_ctx* ctx, const rau_config* cfg, const rau_batch* bt, const float* Pr, float* gE,
                            float* gR, int train, const Encoder* en, const float* dq, bool side_ok = false) {
  const int B = bt->B, Hq = cfg->Hq, E = cfg->embed, Q = 4 * Hq, G4 = 4 * Hq, Tm = en->Tm;
  RnnLayerOff L[4];
  rnn_offsets(cfg, L);
  const bool de = train && cfg->p_embed > 0, dr = train && cfg->p_rnn > 0;
  ARENA(dG1, float, "encb.dG1", (size_t)cfg->T * B * G4);
  ARENA(dG2, float, "encb.dG2", (size_t)cfg->T * B * G4);
  ARENA(dC, float, "encb.dC", (size_t)B * Hq);
  ARENA(dH, float, "encb.dH", (size_t)B * Hq);
  ARENA(du2, float, "encb.du2", (size_t)cfg->T * B * Hq);
  ARENA(de_all, float, "encb.de", (size_t)cfg->T * B * E);
  const bool fused = encoder_fused(ctx, cfg, B);
  if (fused) {
    // tcgen05 path: the pointwise cell backward writes dG packed (hi, lo) next to the fp32 copy, the recurrent dgrad is a
    // split-K product on the rows engine straight from it, and the four weight gradients reuse the packed operands the
    // forward pass left behind (x, h_{t-1}) -- no pack kernels in the time loop
    const bool x3 = prec_x3(ctx);
    const size_t gb = (size_t)B * G4, hb = (size_t)B * Hq;
    const int R = Tm * B;
    ARENA(dGp, bf16, "encb.dGp", (size_t)4 * cfg->T * gb);   // [layer][hi, lo][T][B][4H]
    ARENA(dH2, float, "encb.dH2", (size_t)2 * hb);            // ping-pong: the cell backward of step t clears the buffer
    float* dHb[2] = {dH2, dH2 + hb};                          // the dgrad of step t reduces into, while reading the other
    bf16* hpk_all = nullptr;
    RAU_TRY(ctx->arena.get("enc.hpk", sizeof(bf16) * (size_t)4 * (cfg->T + 1) * hb, (void**)&hpk_all));
    // Wavefront (default; RAU_ENC_BWD_WAVE=0 keeps the layer-after-layer form below): step t of layer 1 needs step t of
    // layer 2 (through the dropped-out h1 that feeds it, D:38-39) and its own step t+1, nothing else -- so the two layers'
    // backward recurrences run on two streams one step apart: 2 T serial steps become T + 1.  Per step, layer 2 forms
    // [dH2_{t-1} | du2_t] = dG2_t [Wh2 | Wi2] in ONE split-K product (the input gradient is not hoisted over time any more)
    // and layer 1 applies the dropout mask of u2 inside its cell backward.
    const bool wave = ctx->tune.enc_bwd_wave != 0 && ctx->aux2 != nullptr && Tm >= 2;
    if (wave) {
      bf16* dG1_hi = dGp;
      bf16* dG1_lo = x3 ? dG1_hi + (size_t)cfg->T * gb : nullptr;
      bf16* dG2_hi = dGp + (size_t)2 * cfg->T * gb;
      bf16* dG2_lo = x3 ? dG2_hi + (size_t)cfg->T * gb : nullptr;
      const bf16 *Wh1_h, *Wh1_l, *Wi1_h, *Wi1_l;
      int64_t ldwh1, ldwi1;
      RAU_TRY(rows_pack2d(ctx, Pr + L[0].Wh, Hq, G4, Hq, x3, true, nullptr, &Wh1_h, &Wh1_l, &ldwh1));
      RAU_TRY(rows_pack2d(ctx, Pr + L[0].Wi, E, G4, E, x3, true, nullptr, &Wi1_h, &Wi1_l, &ldwi1));
      // [Wh2 | Wi2] side by side: [4H rows, 2 Hq columns] (hi, lo)
      ARENA(Wcat, bf16, "encb.Wcat", (size_t)2 * G4 * 2 * Hq);
      bf16* Wcat_hi = Wcat;
      bf16* Wcat_lo = x3 ? Wcat + (size_t)G4 * 2 * Hq : nullptr;
      if (ctx->tc_epoch["encb.Wcat"] != ctx->epoch) {   // (packed once per public call, like every weight shadow)
        RAU_TRY(rows_pack_into(ctx, Pr + L[1].Wh, Hq, G4, Hq, Wcat_hi, Wcat_lo, 2 * Hq, Hq));
        RAU_TRY(rows_pack_into(ctx, Pr + L[1].Wi, Hq, G4, Hq, Wcat_hi + Hq, Wcat_lo ? Wcat_lo + Hq : nullptr, 2 * Hq, Hq));
        ctx->tc_epoch["encb.Wcat"] = ctx->epoch;
      }
      // out2[t] = [dH2 into step t | du2 of step t+1 ...]: slab s (1-based step that PRODUCED it) holds dG2_s [Wh2 | Wi2]:
      // columns 0..Hq-1 = gradient into h2_{s-1} (read by layer 2's step s-1), Hq.. = gradient into u2_s (layer 1's step s)
      ARENA(out2, float, "encb.out2", (size_t)(cfg->T + 1) * B * 2 * Hq);
      ARENA(dC1, float, "encb.dC1", (size_t)B * Hq);
      ARENA(dH1, float, "encb.dH1", (size_t)2 * hb);
      RAU_CHECK_CUDA(cudaMemsetAsync(out2, 0, sizeof(float) * (size_t)(Tm + 1) * B * 2 * Hq, ctx->stream));
      cudaStream_t laneA = ctx->stream, laneB = ctx->aux2;
      cudaEvent_t fork = rau_side_event(ctx);
      RAU_REQUIRE(fork != nullptr, "cudaEventCreate failed");
      RAU_CHECK_CUDA(cudaEventRecord(fork, laneA));
      RAU_CHECK_CUDA(cudaStreamWaitEvent(laneB, fork, 0));
      struct Back2 { rau_ctx* c; cudaStream_t s; ~Back2() { c->stream = s; } } back2{ctx, laneA};
      float* dH1b[2] = {dH1, dH1 + hb};
      // Weight gradients gWi += dG^T x, gWh += dG^T h_{t-1}, gb += sum dG of both layers over the rows of steps s0+1 .. s0+ns
      // (packed operands of the forward pass), on the side stream when allowed.  They go out in two halves: the late steps'
      // rows as soon as both lanes are past them, the early steps' rows at the end -- only the second half is exposed.
      const bool wg_side = side_ok && ctx->side != nullptr;
      const int th = Tm >= 8 ? Tm / 2 + 1 : 1;   // the loop hands over steps th .. Tm when it has finished step th
      auto wgrads = [&](int s0, int ns) -> int {
        if (ns <= 0) return RAU_OK;
        cudaStream_t cur = ctx->stream;
        if (wg_side) {   // after everything both lanes have enqueued so far
          cudaEvent_t ea = rau_side_event(ctx), eb = rau_side_event(ctx);
          RAU_REQUIRE(ea != nullptr && eb != nullptr, "cudaEventCreate failed");
          RAU_CHECK_CUDA(cudaEventRecord(ea, laneA));
          RAU_CHECK_CUDA(cudaEventRecord(eb, laneB));
          RAU_CHECK_CUDA(cudaStreamWaitEvent(ctx->side, ea, 0));
          RAU_CHECK_CUDA(cudaStreamWaitEvent(ctx->side, eb, 0));
          ctx->stream = ctx->side;
          ctx->rows_cta_cap = ctx->side_ctas;
        } else {         // one stream: lane B's rows must be complete before the chain reads them
          cudaEvent_t eb = rau_side_event(ctx);
          RAU_REQUIRE(eb != nullptr, "cudaEventCreate failed");
          RAU_CHECK_CUDA(cudaEventRecord(eb, laneB));
          RAU_CHECK_CUDA(cudaStreamWaitEvent(laneA, eb, 0));
          ctx->stream = laneA;
        }
        struct Back { rau_ctx* c; cudaStream_t s; ~Back() { c->stream = s; c->rows_cta_cap = 0; } } back{ctx, cur};
        const int64_t r0 = (int64_t)s0 * B;
        const int Rn = ns * B;
        for (int layer = 1; layer >= 0; --layer) {
          const int in = layer == 0 ? E : Hq;
          const float* dG = (layer == 1 ? dG2 : dG1) + r0 * G4;
          const bf16* dGh = (layer == 1 ? dG2_hi : dG1_hi) + r0 * G4;
          const bf16* dGl = x3 ? (layer == 1 ? dG2_lo : dG1_lo) + r0 * G4 : nullptr;
          const bf16* x_h = nullptr;
          const int64_t ldx = (in + 7) / 8 * 8;
          const size_t xhalf = ((size_t)R * ldx * sizeof(bf16) + 1023) / 1024 * 1024;
          RAU_TRY(ctx->arena.get(layer == 0 ? "rp.enc.x0" : "rp.enc.x1", xhalf * (x3 ? 2 : 1), (void**)&x_h));
          const bf16* x_l = x3 ? (const bf16*)((const char*)x_h + xhalf) : nullptr;
          const bf16* hp_h = hpk_all + (size_t)(2 * layer) * (cfg->T + 1) * hb;
          const bf16* hp_l = x3 ? hp_h + (size_t)(cfg->T + 1) * hb : nullptr;
          for (int which = 0; which < 2; ++which) {
            RowsGemm g;
            g.M = G4; g.N = which == 0 ? in : Hq; g.K = Rn;
            g.A.hi = dGh; g.A.lo = dGl; g.A.mn = 1; g.A.ld = G4;
            g.B.hi = (which == 0 ? x_h + r0 * ldx : hp_h + r0 * Hq);
            g.B.lo = x3 ? (which == 0 ? x_l + r0 * ldx : hp_l + r0 * Hq) : nullptr;
            g.B.mn = 1; g.B.ld = which == 0 ? ldx : Hq;
            g.epi = ROWS_EPI_RED; g.out_f = gR + (which == 0 ? L[layer].Wi : L[layer].Wh); g.ldo = which == 0 ? in : Hq;
            RAU_TRY(rows_gemm(ctx, g));
          }
          RAU_TRY(k_colsum(ctx, dG, Rn, G4, G4, gR + L[layer].bi, 1, gR + L[layer].bh));
        }
        if (wg_side && ctx->phases == 2) rau_phase_mark(ctx, "enc weight gradients (half) done");
        return RAU_OK;
      };
      for (int t = Tm; t >= 1; --t) {
        const bool last = t == Tm;
        // ---- lane A: layer 2, step t
        ctx->stream = laneA;
        {
          const float* Sp_ = en->S_all + (size_t)(t - 1) * B * Q + 2 * Hq;
          const float* dh_in = last ? nullptr : out2 + (size_t)(t + 1) * B * 2 * Hq;   // (slab t+1, columns 0..Hq-1)
          RAU_TRY(k_lstm_bwd(ctx, B, Hq, RAU_GATES_IFOG, last ? nullptr : dC, Hq, dh_in, 2 * Hq, nullptr, 0, bt->lengths, t,
                             dq + 2 * Hq, dq + 3 * Hq, Q, Sp_, Q, en->sav2 + (size_t)(t - 1) * 5 * hb, dG2 + (size_t)(t - 1) * gb,
                             dG2_hi + (size_t)(t - 1) * gb, dC, Hq, dG2_lo ? dG2_lo + (size_t)(t - 1) * gb : nullptr, nullptr));
          RowsGemm g;
          g.M = B; g.N = 2 * Hq; g.K = G4;
          g.A.hi = dG2_hi + (size_t)(t - 1) * gb; g.A.lo = dG2_lo ? dG2_lo + (size_t)(t - 1) * gb : nullptr; g.A.ld = G4;
          g.B.hi = Wcat_hi; g.B.lo = Wcat_lo; g.B.mn = 1; g.B.ld = 2 * Hq;
          g.epi = ROWS_EPI_RED; g.out_f = out2 + (size_t)t * B * 2 * Hq; g.ldo = 2 * Hq;
          // (both lanes' split-K products size themselves for the chain's whole SM share: halving it per lane measured
          // slower, 4.49 vs 4.38 ms -- the launch that comes first finishes sooner and the other fills in behind it)
          RAU_TRY(rows_gemm(ctx, g));
        }
        cudaEvent_t ev = rau_side_event(ctx);
        RAU_REQUIRE(ev != nullptr, "cudaEventCreate failed");
        RAU_CHECK_CUDA(cudaEventRecord(ev, laneA));
        // ---- lane B: layer 1, step t (one step behind layer 2)
        RAU_CHECK_CUDA(cudaStreamWaitEvent(laneB, ev, 0));
        ctx->stream = laneB;
        {
          const float* Sp_ = en->S_all + (size_t)(t - 1) * B * Q;
          float* dH = dH1b[t & 1];
          float* dH_next = dH1b[(t + 1) & 1];
          RAU_TRY(k_lstm_bwd(ctx, B, Hq, RAU_GATES_IFOG, last ? nullptr : dC1, Hq, last ? nullptr : dH, Hq,
                             out2 + (size_t)t * B * 2 * Hq + Hq, 2 * Hq, bt->lengths, t, dq, dq + Hq, Q, Sp_, Q,
                             en->sav1 + (size_t)(t - 1) * 5 * hb, dG1 + (size_t)(t - 1) * gb, dG1_hi + (size_t)(t - 1) * gb, dC1, Hq,
                             dG1_lo ? dG1_lo + (size_t)(t - 1) * gb : nullptr, t > 1 ? dH_next : nullptr,
                             dr ? en->rbits : nullptr, (int64_t)(t - 1) * B * Hq, drop_scale(cfg->p_rnn)));
          if (t > 1) {
            RowsGemm g;
            g.M = B; g.N = Hq; g.K = G4;
            g.A.hi = dG1_hi + (size_t)(t - 1) * gb; g.A.lo = dG1_lo ? dG1_lo + (size_t)(t - 1) * gb : nullptr; g.A.ld = G4;
            g.B.hi = Wh1_h; g.B.lo = Wh1_l; g.B.mn = 1; g.B.ld = ldwh1;
            g.epi = ROWS_EPI_RED; g.out_f = dH_next; g.ldo = Hq;
            RAU_TRY(rows_gemm(ctx, g));
          }
        }
        if (t == th && th > 1) RAU_TRY(wgrads(th - 1, Tm - th + 1));   // steps th .. Tm are final on both lanes
      }
      // the early steps' rows of the weight gradients (the late ones went out in the middle of the loop)
      RAU_TRY(wgrads(0, th > 1 ? th - 1 : Tm));
      // lane B rejoins the chain: gradient into the word embedding for every step at once, de = dG1 Wi1
      cudaEvent_t join = rau_side_event(ctx);
      RAU_REQUIRE(join != nullptr, "cudaEventCreate failed");
      RAU_CHECK_CUDA(cudaEventRecord(join, laneB));
      ctx->stream = laneA;
      RAU_CHECK_CUDA(cudaStreamWaitEvent(laneA, join, 0));
      {
        RowsGemm g;
        g.M = R; g.N = E; g.K = G4;
        g.A.hi = dG1_hi; g.A.lo = dG1_lo; g.A.ld = G4;
        g.B.hi = Wi1_h; g.B.lo = Wi1_l; g.B.mn = 1; g.B.ld = ldwi1;
        g.epi = ROWS_EPI_LINEAR; g.out_f = de_all; g.ldo = E;
        RAU_TRY(rows_gemm(ctx, g));
      }
      RAU_TRY(k_embed_bwd(ctx, bt->tokens, Tm * B, E, cfg->V, en->e_all, de ? en->ebits : nullptr, drop_scale(cfg->p_embed),
                          de_all, E, gE));
      return RAU_OK;
    }
    for (int layer = 1; layer >= 0; --layer) {
      float* dG = layer == 1 ? dG2 : dG1;
      bf16* dG_hi = dGp + (size_t)(2 * layer) * cfg->T * gb;
      bf16* dG_lo = x3 ? dG_hi + (size_t)cfg->T * gb : nullptr;
      const bf16 *Wh_h, *Wh_l, *Wi_h, *Wi_l;
      int64_t ldwh, ldwi;
      const int in = layer == 0 ? E : Hq;
      RAU_TRY(rows_pack2d(ctx, Pr + L[layer].Wh, Hq, G4, Hq, x3, true, nullptr, &Wh_h, &Wh_l, &ldwh));
      RAU_TRY(rows_pack2d(ctx, Pr + L[layer].Wi, in, G4, in, x3, true, nullptr, &Wi_h, &Wi_l, &ldwi));
      for (int t = Tm; t >= 1; --t) {
        const bool last = t == Tm;
        const float* Sp_ = en->S_all + (size_t)(t - 1) * B * Q + 2 * layer * Hq;
        float* dH = dHb[t & 1];            // written by step t+1's dgrad
        float* dH_next = dHb[(t + 1) & 1];   // cleared here, reduced into by this step's dgrad
        RAU_TRY(k_lstm_bwd(ctx, B, Hq, RAU_GATES_IFOG, last ? nullptr : dC, Hq, last ? nullptr : dH, Hq,
                           layer == 0 ? du2 + (size_t)(t - 1) * hb : nullptr, Hq, bt->lengths, t, dq + 2 * layer * Hq,
                           dq + (2 * layer + 1) * Hq, Q, Sp_, Q,
                           (layer == 1 ? en->sav2 : en->sav1) + (size_t)(t - 1) * 5 * hb, dG + (size_t)(t - 1) * gb,
                           dG_hi + (size_t)(t - 1) * gb, dC, Hq, dG_lo ? dG_lo + (size_t)(t - 1) * gb : nullptr,
                           t > 1 ? dH_next : nullptr));
        if (t > 1) {   // dH = dG_t Wh: K = 4H over a [B, Hq] output -> split over the SMs, partial sums TMA-reduced
          RowsGemm g;
          g.M = B; g.N = Hq; g.K = G4;
          g.A.hi = dG_hi + (size_t)(t - 1) * gb; g.A.lo = dG_lo ? dG_lo + (size_t)(t - 1) * gb : nullptr; g.A.ld = G4;
          g.B.hi = Wh_h; g.B.lo = Wh_l; g.B.mn = 1; g.B.ld = ldwh;
          g.epi = ROWS_EPI_RED; g.out_f = dH_next; g.ldo = Hq;
          RAU_TRY(rows_gemm(ctx, g));
        }
      }
      {   // gradient into the layer input for every step at once: dX = dG Wi
        RowsGemm g;
        g.M = R; g.N = in; g.K = G4;
        g.A.hi = dG_hi; g.A.lo = dG_lo; g.A.ld = G4;
        g.B.hi = Wi_h; g.B.lo = Wi_l; g.B.mn = 1; g.B.ld = ldwi;
        g.epi = ROWS_EPI_LINEAR; g.out_f = layer == 1 ? du2 : de_all; g.ldo = in;
        RAU_TRY(rows_gemm(ctx, g));
      }
      if (layer == 1)
        RAU_TRY(k_dropout_bwd_acc(ctx, du2, (int64_t)R * Hq, dr ? en->rbits : nullptr, drop_scale(cfg->p_rnn), du2, 0));
      // weight gradients over all R = Tm*B rows: gWi += dG^T x, gWh += dG^T h_{t-1} (packed operands of the forward pass).
      // Nothing on the chain reads them: in the training step they go to the side stream (joined by the caller).
      cudaStream_t chain = ctx->stream;
      const bool wg_side = side_ok && ctx->side != nullptr;
      if (wg_side) {
        cudaEvent_t ev = rau_side_event(ctx);
        RAU_REQUIRE(ev != nullptr, "cudaEventCreate failed");
        RAU_CHECK_CUDA(cudaEventRecord(ev, chain));
        RAU_CHECK_CUDA(cudaStreamWaitEvent(ctx->side, ev, 0));
        ctx->stream = ctx->side;
        ctx->rows_cta_cap = ctx->side_ctas;
      }
      struct Back { rau_ctx* c; cudaStream_t s; ~Back() { c->stream = s; c->rows_cta_cap = 0; } } back{ctx, chain};
      const bf16* x_h = nullptr;
      const int64_t ldx = (in + 7) / 8 * 8;
      const size_t xhalf = ((size_t)R * ldx * sizeof(bf16) + 1023) / 1024 * 1024;
      RAU_TRY(ctx->arena.get(layer == 0 ? "rp.enc.x0" : "rp.enc.x1", xhalf * (x3 ? 2 : 1), (void**)&x_h));
      const bf16* x_l = x3 ? (const bf16*)((const char*)x_h + xhalf) : nullptr;
      const bf16* hp_h = hpk_all + (size_t)(2 * layer) * (cfg->T + 1) * hb;
      const bf16* hp_l = x3 ? hp_h + (size_t)(cfg->T + 1) * hb : nullptr;
      for (int which = 0; which < 2; ++which) {
        RowsGemm g;
        g.M = G4; g.N = which == 0 ? in : Hq; g.K = R;
        g.A.hi = dG_hi; g.A.lo = dG_lo; g.A.mn = 1; g.A.ld = G4;
        g.B.hi = which == 0 ? x_h : hp_h; g.B.lo = which == 0 ? x_l : hp_l; g.B.mn = 1; g.B.ld = which == 0 ? ldx : Hq;
        g.epi = ROWS_EPI_RED; g.out_f = gR + (which == 0 ? L[layer].Wi : L[layer].Wh); g.ldo = which == 0 ? in : Hq;
        RAU_TRY(rows_gemm(ctx, g));
      }
      RAU_TRY(k_colsum(ctx, dG, R, G4, G4, gR + L[layer].bi, 1, gR + L[layer].bh));
      if (wg_side && ctx->phases == 2) rau_phase_mark(ctx, "enc layer weight gradients done");
    }
    RAU_TRY(k_embed_bwd(ctx, bt->tokens, Tm * B, E, cfg->V, en->e_all, de ? en->ebits : nullptr, drop_scale(cfg->p_embed),
                        de_all, E, gE));
    return RAU_OK;
  }
  // layer 2, t = Tm..1
  for (int t = Tm; t >= 1; --t) {
    const bool last = t == Tm;
    const float* Sp_ = en->S_all + (size_t)(t - 1) * B * Q;
    float* dGt = dG2 + (size_t)(t - 1) * B * G4;
    RAU_TRY(k_lstm_bwd(ctx, B, Hq, RAU_GATES_IFOG, last ? nullptr : dC, Hq, last ? nullptr : dH, Hq, nullptr, 0,
                       bt->lengths, t, dq + 2 * Hq, dq + 3 * Hq, Q, Sp_ + 2 * Hq, Q,
                       en->sav2 + (size_t)(t - 1) * 5 * B * Hq, dGt, nullptr, dC, Hq));
    if (t > 1) RAU_TRY(rau_contract(ctx, lin_dgrad(B, G4, Hq, dGt, G4, Pr + L[1].Wh, dH, Hq)));
  }
  // gradient into h1 through layer 2's input, all steps at once: du2 = (dG2 Wi2) * mask
  RAU_TRY(rau_contract(ctx, lin_dgrad(Tm * B, G4, Hq, dG2, G4, Pr + L[1].Wi, du2, Hq)));
  RAU_TRY(k_dropout_bwd_acc(ctx, du2, (int64_t)Tm * B * Hq, dr ? en->rbits : nullptr, drop_scale(cfg->p_rnn), du2, 0));
  // layer 1, t = Tm..1
  for (int t = Tm; t >= 1; --t) {
    const bool last = t == Tm;
    const float* Sp_ = en->S_all + (size_t)(t - 1) * B * Q;
    float* dGt = dG1 + (size_t)(t - 1) * B * G4;
    RAU_TRY(k_lstm_bwd(ctx, B, Hq, RAU_GATES_IFOG, last ? nullptr : dC, Hq, last ? nullptr : dH, Hq,
                       du2 + (size_t)(t - 1) * B * Hq, Hq, bt->lengths, t, dq, dq + Hq, Q, Sp_, Q,
                       en->sav1 + (size_t)(t - 1) * 5 * B * Hq, dGt, nullptr, dC, Hq));
    if (t > 1) RAU_TRY(rau_contract(ctx, lin_dgrad(B, G4, Hq, dGt, G4, Pr + L[0].Wh, dH, Hq)));
  }
  // word embedding: de = dG1 Wi1 for all steps, scatter-add into gE (LookupTable accGradParameters)
  RAU_TRY(rau_contract(ctx, lin_dgrad(Tm * B, G4, E, dG1, G4, Pr + L[0].Wi, de_all, E)));
  RAU_TRY(k_embed_bwd(ctx, bt->tokens, Tm * B, E, cfg->V, en->e_all, de ? en->ebits : nullptr, drop_scale(cfg->p_embed),
                      de_all, E, gE));
  // weight gradients as single contractions over all Tm*B rows
  const int R = Tm * B;
  const int ks = R >= 2048 ? 8 : (R >= 512 ? 4 : 1);
  auto wg = [&](const float* dG, const float* Xin, int ldx, int K, float* gW) {
    SimtGemm g = lin_wgrad(R, G4, K, dG, G4, Xin, ldx, gW, 1.0f);
    g.ksplit = ks;
    return rau_contract(ctx, g);
  };
  RAU_TRY(wg(dG1, en->e_all, E, E, gR + L[0].Wi));
  RAU_TRY(wg(dG1, en->S_all + Hq, Q, Hq, gR + L[0].Wh));          // h1 of the previous step
  RAU_TRY(wg(dG2, en->u2, Hq, Hq, gR + L[1].Wi));
  RAU_TRY(wg(dG2, en->S_all + 3 * Hq, Q, Hq, gR + L[1].Wh));      // h2 of the previous step
  RAU_TRY(k_colsum(ctx, dG1, R, G4, G4, gR + L[0].bi, 1));
  RAU_TRY(k_colsum(ctx, dG1, R, G4, G4, gR + L[0].bh, 1));
  RAU_TRY(k_colsum(ctx, dG2, R, G4, G4, gR + L[1].bi, 1));
  RAU_TRY(k_colsum(ctx, dG2, R, G4, G4, gR + L[1].bh, 1));
  return RAU_OK;
}

// f16_ok: the entry point can run from the fp16 features alone (rau_batch.feats_f16, `feats` == NULL)
static int check_batch(const rau_config* cfg, const rau_batch* bt, bool f16_ok) {
  RAU_REQUIRE(bt != nullptr, "batch == NULL");
  RAU_REQUIRE(bt->B > 0, "batch size %d", bt->B);
  RAU_REQUIRE(bt->max_len >= 0 && bt->max_len <= cfg->T, "max_len %d outside [0, T=%d]", bt->max_len, cfg->T);
  if (bt->feats == nullptr && bt->feats_f16 != nullptr) {
    RAU_REQUIRE(f16_ok, "this entry point needs float32 features: batch.feats is NULL (only batch.feats_f16 is set)");
    RAU_TRY(rau_check_dev(bt->feats_f16, "batch.feats_f16"));
    RAU_REQUIRE(((uintptr_t)bt->feats_f16 & 7) == 0, "batch.feats_f16 must be 8-byte aligned");
  } else {
    RAU_TRY(rau_check_dev(bt->feats, "batch.feats"));
  }
  RAU_TRY(rau_check_dev(bt->tokens, "batch.tokens"));
  RAU_TRY(rau_check_dev(bt->lengths, "batch.lengths"));
  return RAU_OK;
}

extern "C" {

static int feval_validate(rau_ctx* ctx, const rau_config* cfg, const rau_batch* bt, float* const params[3],
                          float* const grads[3]) {
  RAU_REQUIRE(ctx, "ctx == NULL");
  RAU_TRY(rau_check_cfg(cfg));
  RAU_REQUIRE(cfg->nlayer == 2, "the fused encoder supports nlayer == 2 (F:209), got %d", cfg->nlayer);
  RAU_TRY(check_batch(cfg, bt, true));
  RAU_TRY(rau_check_dev(bt->labels, "batch.labels"));
  RAU_REQUIRE(params && grads, "params/grads == NULL");
  for (int g = 0; g < 3; ++g) {
    RAU_TRY(rau_check_dev(params[g], "params[g]"));
    RAU_TRY(rau_check_dev(grads[g], "grads[g]"));
  }
  RAU_CHECK_CUDA(cudaSetDevice(ctx->device));
  return RAU_OK;
}

// stage the per-step scalars on the device (outside any captured graph)
static int upload_step_state(rau_ctx* ctx, const StepState& st) {
  const int slot = ctx->ss_slot;
  ctx->ss_slot = (ctx->ss_slot + 1) % 64;
  // a caller may enqueue more than 64 steps without synchronising (graph replay costs the host microseconds): the slot
  // is rewritten only after the copy that read it has run
  if (ctx->ss_ev[slot] == nullptr) RAU_CHECK_CUDA(cudaEventCreateWithFlags(&ctx->ss_ev[slot], cudaEventDisableTiming));
  else RAU_CHECK_CUDA(cudaEventSynchronize(ctx->ss_ev[slot]));
  StepState* h = ctx->h_ss + slot;
  *h = st;
  RAU_CHECK_CUDA(cudaMemcpyAsync(ctx->d_ss, h, sizeof(StepState), cudaMemcpyHostToDevice, ctx->stream));
  RAU_CHECK_CUDA(cudaEventRecord(ctx->ss_ev[slot], ctx->stream));
  return RAU_OK;
}

// enqueue feval; step-dependent values come from ctx->ss_active (device) so the sequence can be graph-captured
static int feval_enqueue(rau_ctx* ctx, const rau_config* cfg, const rau_batch* bt, float* const params[3],
                         float* const grads[3], const float* hop_mask, const rau_masks* masks, const rau_step_out* out) {
  ctx->epoch++;
  const int64_t step_t = 0;   // the step part of every Philox stream id is added on the device
  const int B = bt->B, nHop = cfg->nHop, N = cfg->N, S = cfg->S, H = cfg->H, Q = 4 * cfg->Hq;
  const int Bg = bt->B_global > 0 ? bt->B_global : B;
  const int rank = rau_comm_rank(ctx);
  const int train = 1;

  rau_phase_mark(ctx, "begin");
  // cross-stream overlap (rows path): RAU_OVERLAP bit 0 = heavy backward products on the side stream, bit 1 = the
  // state-independent i_embed products of all hops on the side stream, next to the encoder and the chain
  const int overlap_mode = ctx->tune.overlap;
  const bool rows_hops = hop_rows_path(ctx, cfg) && ctx->side != nullptr;
  const bool ov_bwd = rows_hops && (overlap_mode & 1);
  const bool ov_fwd = rows_hops && (overlap_mode & 2);
  const bool ov_head = rows_hops && (overlap_mode & 4);   // bit 2: the answer heads + criteria of the forward unroll
  // fp16-only features (rau_batch.feats_f16 without feats): the all-hops feature pack is the one kernel that reads them
  RAU_REQUIRE(bt->feats != nullptr || (ov_fwd && masks == nullptr && cfg->p_x > 0 && cfg->nHop > 1 && cfg->S <= 200 &&
                                       ctx->tune.xprep_hops != 0),
              "batch.feats is NULL: fp16-only features need the all-hops feature pack (tcgen05 rows path with the side stream, "
              "drawn masks, p_x > 0, nHop > 1, S <= 200)");
  if (ctx->side_ctas == 0) {
    ctx->side_ctas = ctx->tune.side_ctas > 0 ? ctx->tune.side_ctas : (ctx->sm_count * 4) / 7;   // 84 of 148 SMs (profiles/README.md)
    if (ctx->side_ctas < 8 || ctx->side_ctas > ctx->sm_count) ctx->side_ctas = ctx->sm_count;
    ctx->side_ctas_bwd = ctx->tune.side_ctas_bwd;
    if (ctx->side_ctas_bwd < 0 || ctx->side_ctas_bwd > ctx->sm_count) ctx->side_ctas_bwd = 0;
    ctx->side_ctas_fwd = ctx->tune.side_ctas_fwd > 0 ? ctx->tune.side_ctas_fwd : ctx->side_ctas;
    if (ctx->side_ctas_fwd < 8 || ctx->side_ctas_fwd > ctx->sm_count) ctx->side_ctas_fwd = ctx->sm_count;
  }
  ctx->side_ev_next = 0;
  {
    // the chain's split-K products size themselves for the SMs the side stream leaves free (one wave, not two)
    const int mc = ctx->tune.main_ctas;
    const int auto_cap = ctx->sm_count - ctx->side_ctas >= 32 ? ctx->sm_count - ctx->side_ctas : 0;
    ctx->main_cta_cap = (ov_bwd || ov_fwd) ? (mc > 0 ? mc : auto_cap) : 0;
  }
  struct CapGuard { rau_ctx* c; ~CapGuard() { c->main_cta_cap = 0; } } cap_guard{ctx};
  // the side stream's forward work (feature packs, i_embed, Z) needs nothing the chain enqueues below: it forks here
  cudaEvent_t fork0 = nullptr;
  if (ov_fwd) {
    fork0 = rau_side_event(ctx);
    RAU_REQUIRE(fork0 != nullptr, "cudaEventCreate failed");
    RAU_CHECK_CUDA(cudaEventRecord(fork0, ctx->stream));
  }
  // Off the chain's first microseconds, onto the aux stream: the encoder's weight shadows (six small launches the chain
  // otherwise runs in front of its layers' products) and the zero fill of the three gradient vectors (F:446-448; nothing
  // accumulates into them before the backward pass, which is ordered behind the aux stream's join after the encoder).
  cudaEvent_t enc_w_ready = nullptr;
  bool grads_fill_on_aux = false;
  Encoder en;
  RAU_TRY(encoder_alloc(ctx, cfg, B, &en));
  if (fork0 != nullptr && ctx->aux != nullptr) {
    RAU_CHECK_CUDA(cudaStreamWaitEvent(ctx->aux, fork0, 0));
    cudaStream_t chain = ctx->stream;
    ctx->stream = ctx->aux;
    int rc = RAU_OK;
    if (encoder_fused(ctx, cfg, B)) {
      rc = encoder_pack_weights(ctx, cfg, params[1], B, 1);
      en.w_ready0 = rau_side_event(ctx);
      if (rc == RAU_OK && (en.w_ready0 == nullptr || cudaEventRecord(en.w_ready0, ctx->aux) != cudaSuccess)) rc = RAU_ECUDA;
      if (rc == RAU_OK) rc = encoder_pack_weights(ctx, cfg, params[1], B, 0);
      // the recurrent dropout mask (first read by layer 2's dropout), the cleared encoder output and the persistent
      // recurrences' step counters: three more small launches the chain does not have to carry
      const int Tm0 = (bt->max_len > 0 && bt->max_len <= cfg->T) ? bt->max_len : cfg->T;
      if (rc == RAU_OK)
        rc = rau_prepare_mask(ctx, en.rbits, (int64_t)Tm0 * B * cfg->Hq, cfg->p_rnn, train, masks ? masks->rnn : nullptr,
                              stream_of(step_t, SK_RNN, 0, rank));
      unsigned int* cnt = nullptr;
      if (rc == RAU_OK) rc = ctx->arena.get("enc.seqcnt", sizeof(unsigned int) * 128, (void**)&cnt);
      if (rc == RAU_OK && (cudaMemsetAsync(cnt, 0, sizeof(unsigned int) * 128, ctx->aux) != cudaSuccess ||
                           cudaMemsetAsync(en.rnn_out, 0, sizeof(float) * (size_t)B * 4 * cfg->Hq, ctx->aux) != cudaSuccess))
        rc = RAU_ECUDA;
      if (rc == RAU_OK) { en.rbits_done = 1; en.seq_cnt = cnt; }
      enc_w_ready = rau_side_event(ctx);
      if (rc == RAU_OK && (enc_w_ready == nullptr || cudaEventRecord(enc_w_ready, ctx->aux) != cudaSuccess)) rc = RAU_ECUDA;
    }
    ctx->stream = chain;
    RAU_TRY(rc);
    grads_fill_on_aux = true;   // (at the END of the aux stream's preparation below: the fills' thousands of blocks, on the
                                //  higher-priority stream, kept layer 1's input projection off the SMs for ~20 us)
  } else {
    for (int g = 0; g < 3; ++g) RAU_TRY(k_fill(ctx, grads[g], rau_group_size(cfg, g), 0.0f));
  }
  bool side_used = false;
  // Work the chain needs only LATER runs on an auxiliary stream next to it (forked here, joined by the returned event):
  // sized for `cap` SMs so that it does not crowd the chain's own launches.
  auto off_chain = [&](cudaStream_t aux, int cap, const std::function<int()>& fn, cudaEvent_t* done) -> int {
    cudaEvent_t fork = rau_side_event(ctx);
    *done = rau_side_event(ctx);
    RAU_REQUIRE(fork != nullptr && *done != nullptr, "cudaEventCreate failed");
    RAU_CHECK_CUDA(cudaEventRecord(fork, ctx->stream));
    RAU_CHECK_CUDA(cudaStreamWaitEvent(aux, fork, 0));
    cudaStream_t chain = ctx->stream;
    const int cap_saved = ctx->main_cta_cap;
    ctx->stream = aux;
    if (cap > 0) ctx->main_cta_cap = cap;
    const int rc = fn();
    ctx->stream = chain;
    ctx->main_cta_cap = cap_saved;
    RAU_TRY(rc);
    RAU_CHECK_CUDA(cudaEventRecord(*done, aux));
    return RAU_OK;
  };
  // (only when every operand of the moved products has a producer-written packed twin and a pre-packed weight shadow: a
  // product that packs an operand on demand does so into a scratch slot shared by all streams)
  bool split_ok = (ov_fwd || ov_bwd) && ctx->aux != nullptr && ctx->aux2 != nullptr && nHop >= 2;
  en.w_ready = enc_w_ready;
  en.h0_zeroed = enc_w_ready != nullptr ? 1 : 0;
  // The chain's first launches (masks + word embedding) go out before the side stream is released: the all-hops feature
  // pack saturates HBM for its first ~250 us and would stretch these latency-bound launches threefold.
  cudaEvent_t fork_side = fork0;
  if (ov_fwd) {
    RAU_TRY(encoder_forward(ctx, cfg, bt, params[0], params[1], train, masks, step_t, &en, 1));
    fork_side = rau_side_event(ctx);
    RAU_REQUIRE(fork_side != nullptr, "cudaEventCreate failed");
    RAU_CHECK_CUDA(cudaEventRecord(fork_side, ctx->stream));
  }

  // answering units (F:495-537)
  const size_t sv_bytes = hop_saved_layout(cfg, B, nullptr, nullptr);
  ARENA(sv_base, char, "step.saved", sv_bytes * nHop);
  ARENA(c_all, float, "step.c", (size_t)(nHop + 1) * B * H);
  ARENA(h_all, float, "step.h", (size_t)(nHop + 1) * B * H);
  ARENA(scores_own, float, "step.scores", (size_t)nHop * B * N);
  ARENA(dscore, float, "step.dscore", (size_t)nHop * B * N);
  ARENA(dop_own, float, "step.dop", (size_t)nHop * B);
  ARENA(att_own, float, "step.att", (size_t)nHop * B * S);
  ARENA(ans_own, float, "step.ans", (size_t)(nHop + 2) * B);
  ARENA(loss_own, float, "step.loss", 2 * nHop + 2);
  float* scores = (out && out->scores) ? out->scores : scores_own;
  float* dop = (out && out->do_pred) ? out->do_pred : dop_own;
  float* att = (out && out->attprob) ? out->attprob : att_own;
  float* ans = (out && out->answers) ? out->answers : ans_own;
  float* loss = (out && out->loss) ? out->loss : loss_own;
  float* loss_dp = (out && out->loss_do_pred) ? out->loss_do_pred : loss_own + nHop + 2;
  // zero fills and the hops' dropout masks feed nothing before the first hop: they run on the aux stream next to the
  // encoder (joined after it)
  cudaStream_t prep_chain = ctx->stream;
  const bool prep_aux = fork0 != nullptr && ctx->aux != nullptr;
  if (prep_aux) {
    RAU_CHECK_CUDA(cudaStreamWaitEvent(ctx->aux, fork0, 0));
    ctx->stream = ctx->aux;
  }
  struct StreamGuard { rau_ctx* c; cudaStream_t s; ~StreamGuard() { if (c->stream == c->aux) c->stream = s; } } prep_guard{ctx, prep_chain};
  RAU_TRY(k_fill(ctx, c_all, (int64_t)B * H, 0.0f));   // F:362-364
  RAU_TRY(k_fill(ctx, h_all, (int64_t)B * H, 0.0f));
  RAU_TRY(k_fill(ctx, loss, nHop + 2, 0.0f));
  RAU_TRY(k_fill(ctx, loss_dp, nHop, 0.0f));
  MultT<const float*> P = mult_views<const float*, const float>(cfg, params[2]);
  MultT<float*> G = mult_views<float*, float>(cfg, grads[2]);
  // small saved activations and backward scratch as tensor-major stacks [nHop][B][dim]: the nn.Linear weight gradients
  // of all hops are then single contractions over nHop*B rows (hop_wgrads)
  const int M_ = cfg->M, A_ = cfg->A;
  ARENA(st_qd, float, "stack.qd", (size_t)nHop * B * Q);
  ARENA(st_qf, float, "stack.qf", (size_t)nHop * B * M_);
  ARENA(st_j, float, "stack.j", (size_t)nHop * B * M_);
  ARENA(st_m, float, "stack.m", (size_t)nHop * B * M_);
  ARENA(st_du, float, "stack.du", (size_t)nHop * B * M_);
  ARENA(st_dG, float, "stack.dG", (size_t)nHop * B * 4 * H);
  ARENA(st_dj, float, "stack.dj", (size_t)nHop * B * M_);
  ARENA(st_ds, float, "stack.ds", (size_t)nHop * B * S);
  ARENA(st_dqa, float, "stack.dqa", (size_t)nHop * B * A_);
  ARENA(st_dpre, float, "stack.dpre", (size_t)nHop * B * M_);
  ARENA(st_gwsp, float, "stack.gwsp", (size_t)nHop * B * A_);
  // packed bf16 (hi, lo) twins of the same stacks, written by the producing kernels: the tcgen05 products of the chain
  // and the deferred weight gradients read them directly (no pack launches)
  const bool twins = ctx->precision != RAU_PREC_F32 && rows_path_enabled();
  const int Sp8 = (S + 7) / 8 * 8;
  PK pk_qd, pk_qf, pk_p, pk_j, pk_h, pk_m, pk_dscore, pk_du, pk_dG, pk_ds, pk_dpre;
  auto mkpk = [&](const char* name, size_t rows, int64_t ld, PK* out) -> int {
    if (!twins || ld % 8 != 0) return RAU_OK;
    bf16* base = nullptr;
    const size_t n = rows * (size_t)ld;
    RAU_TRY(ctx->arena.get(name, sizeof(bf16) * 2 * n, (void**)&base));
    out->hi = base; out->lo = base + n; out->ld = ld;
    return RAU_OK;
  };
  auto slice = [](const PK& pk, size_t row) {
    PK r;
    if (pk.hi) { r.hi = pk.hi + row * pk.ld; r.lo = pk.lo + row * pk.ld; r.ld = pk.ld; }
    return r;
  };
  const size_t nb = (size_t)nHop * B;
  RAU_TRY(mkpk("pk.qd", nb, Q, &pk_qd));      RAU_TRY(mkpk("pk.qf", nb, M_, &pk_qf));   RAU_TRY(mkpk("pk.p", nb, Sp8, &pk_p));
  RAU_TRY(mkpk("pk.j", nb, M_, &pk_j));       RAU_TRY(mkpk("pk.h", nb + B, H, &pk_h));  RAU_TRY(mkpk("pk.m", nb, M_, &pk_m));
  RAU_TRY(mkpk("pk.dscore", nb, N, &pk_dscore)); RAU_TRY(mkpk("pk.du", nb, M_, &pk_du)); RAU_TRY(mkpk("pk.dG", nb, 4 * H, &pk_dG));
  RAU_TRY(mkpk("pk.ds", nb, Sp8, &pk_ds));    RAU_TRY(mkpk("pk.dpre", nb, M_, &pk_dpre));
  if (pk_h.hi) {   // h_0 = 0 (F:362-364)
    RAU_CHECK_CUDA(cudaMemsetAsync(pk_h.hi, 0, sizeof(bf16) * (size_t)B * H, ctx->stream));
    RAU_CHECK_CUDA(cudaMemsetAsync(pk_h.lo, 0, sizeof(bf16) * (size_t)B * H, ctx->stream));
  }
  std::vector<HopSaved> sv(nHop);
  std::vector<HopAsync> as(nHop);
  // drawn masks (no host-provided ones): the q and m keep bits of all hops come from one launch each
  const bool batched_masks = train && masks == nullptr && nHop < 65536;
  if (batched_masks) {
    HopSaved s0;
    hop_saved_layout(cfg, B, sv_base, &s0);
    if (cfg->p_q > 0)
      RAU_TRY(k_mask_gen(ctx, s0.qbits, (int64_t)B * Q, cfg->p_q, ctx->seed, stream_of(step_t, SK_Q, 0, rank), nHop, (int64_t)(sv_bytes / 4)));
    if (cfg->p_m > 0)
      RAU_TRY(k_mask_gen(ctx, s0.mbits, (int64_t)B * cfg->M, cfg->p_m, ctx->seed, stream_of(step_t, SK_M, 0, rank), nHop, (int64_t)(sv_bytes / 4)));
  }
  for (int hp = 0; hp < nHop; ++hp) {
    hop_saved_layout(cfg, B, sv_base + sv_bytes * hp, &sv[hp]);
    as[hp].hop = hp;
    as[hp].bwd_side = ov_bwd ? 1 : 0;
    as[hp].head_side = ov_head ? 1 : 0;
    const size_t r0 = (size_t)hp * B;
    sv[hp].qd_pk = slice(pk_qd, r0); sv[hp].qf_pk = slice(pk_qf, r0); sv[hp].p_pk = slice(pk_p, r0); sv[hp].j_pk = slice(pk_j, r0);
    sv[hp].hin_pk = slice(pk_h, r0); sv[hp].hout_pk = slice(pk_h, r0 + B); sv[hp].m_pk = slice(pk_m, r0);
    sv[hp].qd = st_qd + (size_t)hp * B * Q;
    sv[hp].qf = st_qf + (size_t)hp * B * M_;
    sv[hp].p = att + (size_t)hp * B * S;                 // the module outputs double as the saved copies
    sv[hp].j = st_j + (size_t)hp * B * M_;
    sv[hp].hout = h_all + (size_t)(hp + 1) * B * H;
    sv[hp].dop = dop + (size_t)hp * B;
    sv[hp].m = st_m + (size_t)hp * B * M_;
    if (!batched_masks) {
      RAU_TRY(rau_prepare_mask(ctx, sv[hp].qbits, (int64_t)B * Q, cfg->p_q, train,
                               masks && masks->q ? masks->q + (size_t)hp * B * Q : nullptr, stream_of(step_t, SK_Q, hp, rank)));
      RAU_TRY(rau_prepare_mask(ctx, sv[hp].mbits, (int64_t)B * cfg->M, cfg->p_m, train,
                               masks && masks->m ? masks->m + (size_t)hp * B * cfg->M : nullptr,
                               stream_of(step_t, SK_M, hp, rank)));
    }
    if (hop_rows_path(ctx, cfg) && !(masks && masks->x)) {   // drawn inline by the rows pack kernel
      sv[hp].x_philox = 1;
      sv[hp].x_stream = stream_of(step_t, SK_X, hp, rank);
      sv[hp].x_hop = hp; sv[hp].x_nhop = nHop;
    } else {
      RAU_TRY(rau_prepare_mask(ctx, sv[hp].xbits, (int64_t)B * cfg->C * S, cfg->p_x, train,
                               masks && masks->x ? masks->x + (size_t)hp * B * cfg->C * S : nullptr,
                               stream_of(step_t, SK_X, hp, rank)));
    }
  }
  const bool prepacked = prep_aux && ctx->precision != RAU_PREC_F32 && hop_rows_path(ctx, cfg) && rows_path_enabled();
  split_ok = split_ok && prepacked && pk_qd.hi && pk_dscore.hi && pk_du.hi && pk_dpre.hi && Q % 8 == 0 && M_ % 8 == 0 && H % 8 == 0;
  if (prepacked) {
    // the bf16 (hi, lo) shadows of the weights the chain's products read: packed here, next to the encoder, instead of
    // inline at their first use on the chain (the per-epoch cache makes the later calls no-ops)
    const bool x3 = prec_x3(ctx);
    const bf16 *ph, *pl;
    int64_t pld;
    const float* pb;
    auto prepack = [&](const float* W, int Nout, int Kin) {
      return Kin % 8 == 0 ? rows_pack2d(ctx, W, Kin, Nout, Kin, x3, true, nullptr, &ph, &pl, &pld) : RAU_OK;
    };
    RAU_TRY(prepack(P.Wq, M_, Q)); RAU_TRY(prepack(P.Wh, M_, H)); RAU_TRY(prepack(P.Wqa, A_, M_)); RAU_TRY(prepack(P.Wm, S, H));
    RAU_TRY(prepack(P.Wp, M_, S)); RAU_TRY(prepack(P.Wo, M_, H)); RAU_TRY(prepack(P.Ws, N, M_));
    RAU_TRY(prepack(P.Wx, 4 * H, M_)); RAU_TRY(prepack(P.Whh, 4 * H, H));
    if (H % 8 == 0 && M_ % 8 == 0) {
      RAU_TRY(rows_pack_lstm(ctx, P.Wx, H, M_, RAU_GATES_IGFO, x3, &ph, &pl, &pld));
      RAU_TRY(rows_pack_lstm(ctx, P.Whh, H, H, RAU_GATES_IGFO, x3, &ph, &pl, &pld));
      RAU_TRY(rows_perm_lstm_bias(ctx, P.bx, P.bhh, H, RAU_GATES_IGFO, &pb));
    }
    RnnLayerOff L[4];
    rnn_offsets(cfg, L);
    const int Hq = cfg->Hq, E = cfg->embed;
    if (Hq % 8 == 0 && E % 8 == 0)
      for (int layer = 0; layer < 2; ++layer) {   // the encoder backward's dgrad operands
        RAU_TRY(prepack(params[1] + L[layer].Wh, 4 * Hq, Hq));
        RAU_TRY(prepack(params[1] + L[layer].Wi, 4 * Hq, layer == 0 ? E : Hq));
      }
  }
  cudaEvent_t prep_done = nullptr;
  if (grads_fill_on_aux) {   // (prep_aux holds whenever this does: ctx->stream is the aux stream here)
    for (int g = 0; g < 3; ++g) RAU_TRY(k_fill(ctx, grads[g], rau_group_size(cfg, g), 0.0f));
  }
  if (prep_aux) {
    prep_done = rau_side_event(ctx);
    RAU_REQUIRE(prep_done != nullptr, "cudaEventCreate failed");
    RAU_CHECK_CUDA(cudaEventRecord(prep_done, ctx->aux));
    ctx->stream = prep_chain;
  }
  if (ov_fwd) {
    // the i_embed product of every hop depends on the features only: all of them go to the side stream now and run
    // next to the encoder unroll and the hops' chains; each hop waits for its own event before it reads I
    bool early = fork0 != nullptr;   // (materialised feature masks are written on the chain: fork after them)
    for (int hp = 0; hp < nHop; ++hp) early = early && (sv[hp].x_philox || !(train && cfg->p_x > 0));
    cudaEvent_t fork = fork_side;
    if (!early) {
      if (prep_done) RAU_CHECK_CUDA(cudaStreamWaitEvent(ctx->side, prep_done, 0));   // the materialised masks
      fork = rau_side_event(ctx);
      RAU_REQUIRE(fork != nullptr, "cudaEventCreate failed");
      RAU_CHECK_CUDA(cudaEventRecord(fork, ctx->stream));
    }
    RAU_CHECK_CUDA(cudaStreamWaitEvent(ctx->side, fork, 0));
    cudaStream_t chain = ctx->stream;
    ctx->stream = ctx->side;
    ctx->rows_cta_cap = ctx->side_ctas_fwd;
    int rc = RAU_OK;
    // drawn feature masks: the feature pack of all hops in one launch (the fp32 features are read once, not nHop times)
    // (RAU_XPREP_HOPS=0: one pack launch per hop -- same bits; a test compares the two)
    bool all_philox = train && cfg->p_x > 0 && nHop > 1 && nHop < 65536 && S <= 200 && ctx->tune.xprep_hops != 0;   // (S: its smem slabs)
    for (int hp = 0; hp < nHop; ++hp) all_philox = all_philox && sv[hp].x_philox;
    if (!all_philox && bt->feats == nullptr) {
      rau_set_error("batch.feats is NULL: fp16-only features are read by the all-hops feature pack alone (training, drawn masks, "
                    "p_x > 0, nHop > 1, S <= 200, RAU_XPREP_HOPS != 0)");
      rc = RAU_EINVAL;
    } else if (all_philox) {
      rc = k_xprep_rows_hops(ctx, bt->feats, bt->feats == nullptr ? bt->feats_f16 : nullptr, B, cfg->C, S, nHop, drop_scale(cfg->p_x), sv[0].Xd_hi,
                             (prec_x3(ctx) && !prec_x_f16(ctx)) ? sv[0].Xd_lo : nullptr, (int64_t)(sv_bytes / sizeof(bf16)),
                             cfg->p_x, sv[0].x_stream, prec_x_f16(ctx) ? 1 : 0);
      for (int hp = 0; hp < nHop; ++hp) sv[hp].x_done = 1;
    }
    for (int hp = 0; hp < nHop && rc == RAU_OK; ++hp) {
      rc = hop_forward_pre(ctx, cfg, B, P, bt->feats, train, sv[hp]);
      if (rc == RAU_OK) {
        as[hp].pre_done = rau_side_event(ctx);
        if (as[hp].pre_done == nullptr || cudaEventRecord(as[hp].pre_done, ctx->side) != cudaSuccess) rc = RAU_ECUDA;
        if (ctx->phases == 2) rau_phase_mark(ctx, "hop pre done");
      }
    }
    ctx->stream = chain;
    ctx->rows_cta_cap = 0;
    RAU_TRY(rc);
    side_used = true;
  }
  RAU_TRY(encoder_forward(ctx, cfg, bt, params[0], params[1], train, masks, step_t, &en, ov_fwd ? 2 : 0));
  if (prep_done) RAU_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, prep_done, 0));
  rau_phase_mark(ctx, "encoder forward");
  // Everything of the unroll that depends on the encoder state only is hoisted out of the per-hop chains: the q dropout
  // of every hop (one launch) and Wq drop_h(q) + bq of every hop (one [nHop*B, Q] x [Q, M] product).
  ARENA(st_qpre, float, "stack.qpre", (size_t)nHop * B * M_);
  cudaEvent_t qpre_rest_done = nullptr;
  const int64_t bits_stride = (int64_t)(sv_bytes / 4);   // keep bits of consecutive hops are sv_bytes apart
  {
    const bool dq_ = train && cfg->p_q > 0;
    RAU_TRY(k_dropout_hops(ctx, en.rnn_out, (int64_t)B * Q, nHop, dq_ ? sv[0].qbits : nullptr, bits_stride, drop_scale(cfg->p_q),
                           st_qd, pk_qd.hi, prec_x3(ctx) ? pk_qd.lo : nullptr));
    auto qpre_rows = [&](int h0, int nh) -> int {   // Wq drop_h(q) + bq for hops h0 .. h0+nh-1
      const size_t r0 = (size_t)h0 * B;
      const PK qp = slice(pk_qd, r0);
      SimtGemm g = lin_fwd(nh * B, M_, Q, st_qd + r0 * Q, Q, P.Wq, st_qpre + r0 * M_, M_);
      g.bias_n = P.bq;
      g.Ar_hi = qp.hi; g.Ar_lo = qp.lo; g.Ar_ld = qp.ld;
      return rau_contract(ctx, g);
    };
    if (split_ok) {   // the first hop needs only its own slice: the other hops' product runs next to the first hop's chain
      RAU_TRY(off_chain(ctx->aux2, 32, [&]() { return qpre_rows(1, nHop - 1); }, &qpre_rest_done));
      RAU_TRY(qpre_rows(0, 1));
    } else {
      RAU_TRY(qpre_rows(0, nHop));
    }
    for (int hp = 0; hp < nHop; ++hp) sv[hp].qpre = st_qpre + (size_t)hp * B * M_;
  }
  // The answer head's backward needs forward results only: du = drop'(dscore Ws) and Wo^T du for a range of hops in two
  // products over their stacked rows (dscore already carries the hop mask and 1/B_global)
  ARENA(st_dh2h, float, "stack.dh2h", (size_t)nHop * B * H);
  auto head_backward = [&](int h0, int nh) -> int {
    const size_t r0 = (size_t)h0 * B;
    const PK dsc = slice(pk_dscore, r0), dup = slice(pk_du, r0);
    SimtGemm g = lin_dgrad(nh * B, N, M_, dscore + r0 * N, N, P.Ws, st_du + r0 * M_, M_);
    g.Ar_hi = dsc.hi; g.Ar_lo = dsc.lo; g.Ar_ld = dsc.ld;
    RAU_TRY(rau_contract(ctx, g));
    const bool dm_ = train && cfg->p_m > 0;
    RAU_TRY(k_dropout_bwd_hops(ctx, st_du + r0 * M_, (int64_t)B * M_, nh, dm_ ? sv[h0].mbits : nullptr, bits_stride,
                               drop_scale(cfg->p_m), dup.hi, prec_x3(ctx) ? dup.lo : nullptr));
    SimtGemm g2 = lin_dgrad(nh * B, M_, H, st_du + r0 * M_, M_, P.Wo, st_dh2h + r0 * H, H);
    g2.Ar_hi = dup.hi; g2.Ar_lo = dup.lo; g2.Ar_ld = dup.ld;
    return rau_contract(ctx, g2);
  };
  // (running the head backward of hops 0 .. nHop-2 during the last hop's forward measured 1 % SLOWER -- 4.87 vs 4.81 ms: the
  // extra launches contend with the last hop's chain -- so the whole head backward stays between the two unrolls)
  for (int hp = 0; hp < nHop; ++hp) {
    if (hp == 1 && qpre_rest_done) RAU_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, qpre_rest_done, 0));
    RAU_TRY(hop_forward(ctx, cfg, B, P, en.rnn_out, bt->feats, c_all + (size_t)hp * B * H, h_all + (size_t)hp * B * H, train,
                        sv[hp], scores + (size_t)hp * B * N, dop + (size_t)hp * B, att + (size_t)hp * B * S,
                        c_all + (size_t)(hp + 1) * B * H, h_all + (size_t)(hp + 1) * B * H, &as[hp]));
    if (ctx->phases == 2) rau_phase_mark(ctx, "hop forward chain done");
    // criterion forward + backward + argmax in one pass (F:505, F:535, F:585-589)
    const float hm = hop_mask ? hop_mask[hp] : 1.0f;
    const PK dsc = slice(pk_dscore, (size_t)hp * B);
    cudaStream_t chain = ctx->stream;
    if (ov_head) { ctx->stream = ctx->side; side_used = true; }   // behind this hop's head on the side stream
    const int rc_ce = k_softmax_ce(ctx, B, N, scores + (size_t)hp * B * N, bt->labels, 1.0f / Bg, hm / Bg, loss + hp,
                                   dscore + (size_t)hp * B * N, dsc.hi, N, ans + (size_t)hp * B,
                                   prec_x3(ctx) ? dsc.lo : nullptr);
    if (ctx->phases == 2) rau_phase_mark(ctx, "hop head + criterion done");
    ctx->stream = chain;
    RAU_TRY(rc_ce);
  }
  if (ov_head) {   // scores, do_pred, losses and dscore of every hop are complete before the merge and the backward unroll
    cudaEvent_t ev = rau_side_event(ctx);
    RAU_REQUIRE(ev != nullptr, "cudaEventCreate failed");
    RAU_CHECK_CUDA(cudaEventRecord(ev, ctx->side));
    RAU_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, ev, 0));
  }
  rau_phase_mark(ctx, "answering units forward");
  // logging-only losses on the averaged / selected predictions and the do_pred BCE (F:539-574): nothing in the backward
  // pass reads them, so with the heads on the side stream they stay there (behind the event the chain just waited for)
  {
    cudaStream_t chain = ctx->stream;
    if (ov_head) { ctx->stream = ctx->side; side_used = true; }
    const int rc_m = k_merge_preds(ctx, nHop, B, N, S, scores, dop, nullptr, bt->labels, ans, 0, 1.0f / Bg, loss + nHop, loss_dp,
                                   ans + (size_t)nHop * B, nullptr, nullptr, nullptr, nullptr);
    ctx->stream = chain;
    RAU_TRY(rc_m);
  }

  rau_phase_mark(ctx, "merged losses");
  // BPTT through the hops (F:578-597); do_pred and attprob receive zero gradient (F:582-583, F:592)
  ARENA(dcs, float, "step.dc", (size_t)2 * B * H);
  ARENA(dhs, float, "step.dh", (size_t)2 * B * H);
  ARENA(dq, float, "step.dq", (size_t)B * Q);
  // The answer head's backward needs forward results only: du = drop'(dscore Ws) and Wo^T du of every hop in two products
  // over nHop*B rows before the unroll (dscore already carries the hop mask and 1/B_global)
  ARENA(st_dqt, float, "stack.dqt", (size_t)nHop * B * Q);
  cudaEvent_t head_rest_done = nullptr;
  if (split_ok) {   // the unroll starts at the last hop: the other hops' head backward runs next to its chain
    RAU_TRY(off_chain(ctx->aux2, 32, [&]() { return head_backward(0, nHop - 1); }, &head_rest_done));   // (aux carries the hops' dh lane)
    RAU_TRY(head_backward(nHop - 1, 1));
  } else {
    RAU_TRY(head_backward(0, nHop));
  }
  const int main_cap_saved = ctx->main_cta_cap;
  if (ov_bwd && ctx->side_ctas_bwd > 0 && ctx->sm_count - ctx->side_ctas_bwd >= 16) ctx->main_cta_cap = ctx->sm_count - ctx->side_ctas_bwd;
  // the attention backward's atomic accumulators of every hop, cleared at once (not two memsets inside every hop's chain)
  RAU_CHECK_CUDA(cudaMemsetAsync(st_dqa, 0, sizeof(float) * (size_t)nHop * B * A_, ctx->stream));
  RAU_CHECK_CUDA(cudaMemsetAsync(st_gwsp, 0, sizeof(float) * (size_t)nHop * B * A_, ctx->stream));
  cudaEvent_t dqt_done = nullptr;
  auto dqt_rows = [&](int h0, int nh) -> int {   // (dpre_h Wq) for hops h0 .. h0+nh-1, before the dropout mask of q
    const size_t r0 = (size_t)h0 * B;
    const PK dp_ = slice(pk_dpre, r0);
    SimtGemm g = lin_dgrad(nh * B, M_, Q, st_dpre + r0 * M_, M_, P.Wq, st_dqt + r0 * Q, Q);
    g.Ar_hi = dp_.hi; g.Ar_lo = dp_.lo; g.Ar_ld = dp_.ld;
    return rau_contract(ctx, g);
  };
  for (int hp = nHop - 1; hp >= 0; --hp) {
    const bool last = hp == nHop - 1;
    if (hp == nHop - 2 && head_rest_done) RAU_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, head_rest_done, 0));
    const float* dc_in = last ? nullptr : dcs + (size_t)((hp + 1) & 1) * B * H;
    const float* dh_in = last ? nullptr : dhs + (size_t)((hp + 1) & 1) * B * H;
    HopGrads hg;
    hg.du = st_du + (size_t)hp * B * M_; hg.dG = st_dG + (size_t)hp * B * 4 * H; hg.dj = st_dj + (size_t)hp * B * M_;
    hg.ds = st_ds + (size_t)hp * B * S; hg.dqa = st_dqa + (size_t)hp * B * A_; hg.dpre = st_dpre + (size_t)hp * B * M_;
    hg.gwsp = st_gwsp + (size_t)hp * B * A_;
    hg.dscore_pk = slice(pk_dscore, (size_t)hp * B); hg.du_pk = slice(pk_du, (size_t)hp * B); hg.dG_pk = slice(pk_dG, (size_t)hp * B);
    hg.ds_pk = slice(pk_ds, (size_t)hp * B); hg.dpre_pk = slice(pk_dpre, (size_t)hp * B);
    hg.dh2h = st_dh2h + (size_t)hp * B * H;
    hg.acc_zeroed = 1;
    RAU_TRY(hop_backward(ctx, cfg, B, P, G, bt->feats, c_all + (size_t)hp * B * H, h_all + (size_t)hp * B * H, train, sv[hp],
                         dscore + (size_t)hp * B * N, nullptr, nullptr, dc_in, dh_in, dq, last ? 0 : 1, nullptr,
                         dcs + (size_t)(hp & 1) * B * H, dhs + (size_t)(hp & 1) * B * H, &hg, &as[hp]));
    if (ov_bwd) side_used = true;
    // this hop's share of dq = sum_h drop_h'(dpre_h Wq) is not needed before the unroll ends: next to the chain
    if (split_ok && hp >= 1) RAU_TRY(off_chain(ctx->aux2, 32, [&]() { return dqt_rows(hp, 1); }, &dqt_done));
    if (ctx->phases == 2) {
      rau_phase_mark(ctx, "hop backward chain done");
      if (ov_bwd) { cudaStream_t c = ctx->stream; ctx->stream = ctx->side; rau_phase_mark(ctx, "hop backward products done"); ctx->stream = c; }
    }
  }
  {   // dq = sum_h drop_h'(dpre_h Wq): the products over the stacked dpre, then one masked sum over the hops
    if (split_ok) {
      RAU_TRY(dqt_rows(0, 1));
      if (dqt_done) RAU_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, dqt_done, 0));   // (aux2 is in order: the last one covers all)
    } else {
      RAU_TRY(dqt_rows(0, nHop));
    }
    const bool dq_ = train && cfg->p_q > 0;
    RAU_TRY(k_dropout_bwd_sum_hops(ctx, st_dqt, (int64_t)B * Q, nHop, dq_ ? sv[0].qbits : nullptr, bits_stride, drop_scale(cfg->p_q), dq));
  }
  ctx->main_cta_cap = main_cap_saved;
  rau_phase_mark(ctx, "answering units backward");
  {
    HopStacks st;
    st.dscore = dscore; st.m = st_m; st.du = st_du; st.hout = h_all + (size_t)B * H; st.dG = st_dG; st.j = st_j; st.h_in = h_all;
    st.dj = st_dj; st.p = att; st.ds = st_ds; st.dqa = st_dqa; st.qf = st_qf; st.dpre = st_dpre; st.qd = st_qd;
    st.gwsp = st_gwsp;
    st.dscore_pk = pk_dscore; st.m_pk = pk_m; st.du_pk = pk_du; st.hout_pk = slice(pk_h, B); st.dG_pk = pk_dG; st.j_pk = pk_j;
    st.hin_pk = pk_h; st.p_pk = pk_p; st.ds_pk = pk_ds; st.qf_pk = pk_qf; st.dpre_pk = pk_dpre; st.qd_pk = pk_qd;
    if (ov_bwd) {
      // the deferred nn.Linear weight gradients feed nothing in the encoder backward: side stream, behind the last hop's
      // heavy products, while the chain walks the encoder
      cudaEvent_t ev = rau_side_event(ctx);
      RAU_REQUIRE(ev != nullptr, "cudaEventCreate failed");
      RAU_CHECK_CUDA(cudaEventRecord(ev, ctx->stream));
      RAU_CHECK_CUDA(cudaStreamWaitEvent(ctx->side, ev, 0));
      cudaStream_t chain = ctx->stream;
      ctx->stream = ctx->side;
      ctx->rows_cta_cap = ctx->side_ctas;
      int rc = hop_wgrads(ctx, cfg, nHop * B, G, st);
      if (ctx->phases == 2) rau_phase_mark(ctx, "deferred weight gradients done");
      if (rc == RAU_OK && ctx->early_tail && ctx->aux != nullptr) {
        // every gradient of the answering units is final here (the chain's share was written before the event the side
        // stream waited for above): the rest of this group's step runs now, on the aux stream, under the encoder backward
        cudaEvent_t ev_m = rau_side_event(ctx);
        cudaEvent_t done = rau_side_event(ctx);
        if (ev_m == nullptr || done == nullptr || cudaEventRecord(ev_m, ctx->side) != cudaSuccess ||
            cudaStreamWaitEvent(ctx->aux, ev_m, 0) != cudaSuccess) {
          rc = RAU_ECUDA;
        } else {
          ctx->stream = ctx->aux;
          ctx->rows_cta_cap = 0;
          rc = ctx->early_tail();
          if (rc == RAU_OK && cudaEventRecord(done, ctx->aux) != cudaSuccess) rc = RAU_ECUDA;
          if (rc == RAU_OK) ctx->early_tail_done = done;
        }
      }
      ctx->stream = chain;
      ctx->rows_cta_cap = 0;
      RAU_TRY(rc);
      side_used = true;
    } else {
      RAU_TRY(hop_wgrads(ctx, cfg, nHop * B, G, st));
    }
  }
  rau_phase_mark(ctx, "unit weight gradients");
  RAU_TRY(encoder_backward(ctx, cfg, bt, params[1], grads[0], grads[1], train, &en, dq, ov_bwd));
  if (ov_bwd) side_used = true;
  if (ctx->phases == 2) rau_phase_mark(ctx, "encoder backward chain done");
  if (ctx->early_tail0 && side_used) {   // the embedding gradient is final (k_embed_bwd was the chain's last launch)
    RAU_TRY(ctx->early_tail0());
    ctx->early_tail0_ran = true;
    if (ctx->phases == 2) rau_phase_mark(ctx, "embed group done");
  }
  if (side_used) {   // join: the side stream's gradients (gWi, gWa, gbi) are complete before anything downstream
    cudaEvent_t join = rau_side_event(ctx);
    RAU_REQUIRE(join != nullptr, "cudaEventCreate failed");
    RAU_CHECK_CUDA(cudaEventRecord(join, ctx->side));
    RAU_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, join, 0));
  }
  rau_phase_mark(ctx, "encoder backward");
  return RAU_OK;
}

int rau_feval(rau_ctx* ctx, const rau_config* cfg, const rau_batch* bt, float* const params[3], float* const grads[3],
              const float* hop_mask, const rau_masks* masks, int64_t step_t, const rau_step_out* out) {
  RAU_TRY(feval_validate(ctx, cfg, bt, params, grads));
  RAU_TRY(rau_check_async_error(ctx));
  StepState st;
  memset(&st, 0, sizeof(st));
  st.step = (unsigned long long)step_t;
  RAU_TRY(upload_step_state(ctx, st));
  ctx->ss_active = ctx->d_ss;
  const int r = feval_enqueue(ctx, cfg, bt, params, grads, hop_mask, masks, out);
  ctx->ss_active = nullptr;
  return r;
}

// The keep masks rau_feval / rau_train_step draw from the context's Philox streams for iteration step_t (same seed, same
// rank), written as 0/1 bytes in the layouts of rau_masks.  Parity tests run the step with drawn masks (the benchmarked
// schedule: all-hops feature pack, batched mask launches, graph replay) and hand these bytes to the oracle.
int rau_draw_masks(rau_ctx* ctx, const rau_config* cfg, int B, int64_t step_t, const rau_masks_out* out) {
  RAU_REQUIRE(ctx && out, "ctx/out == NULL");
  RAU_TRY(rau_check_cfg(cfg));
  RAU_REQUIRE(B > 0, "B = %d", B);
  RAU_CHECK_CUDA(cudaSetDevice(ctx->device));
  StepState st;
  memset(&st, 0, sizeof(st));
  st.step = (unsigned long long)step_t;
  RAU_TRY(upload_step_state(ctx, st));
  ctx->ss_active = ctx->d_ss;
  struct Off { rau_ctx* c; ~Off() { c->ss_active = nullptr; } } off{ctx};
  const int T = cfg->T, E = cfg->embed, Hq = cfg->Hq, Q = 4 * Hq, nHop = cfg->nHop, rank = rau_comm_rank(ctx);
  auto one = [&](uint8_t* dst, int64_t n, float p, int kind, int hops) -> int {
    if (dst == nullptr) return RAU_OK;
    RAU_TRY(rau_check_dev(dst, "mask output"));
    const int64_t stride = mask_words(n);
    ARENA(bits, uint32_t, "draw.bits", (size_t)stride * hops);
    RAU_TRY(k_mask_gen(ctx, bits, n, p, ctx->seed, stream_of(0, kind, 0, rank), hops, stride));
    for (int h = 0; h < hops; ++h) RAU_TRY(k_mask_unpack(ctx, bits + (size_t)h * stride, n, dst + (size_t)h * n));
    return RAU_OK;
  };
  RAU_TRY(one(out->embed, (int64_t)T * B * E, cfg->p_embed, SK_EMBED, 1));
  RAU_TRY(one(out->rnn, (int64_t)T * B * Hq, cfg->p_rnn, SK_RNN, 1));
  RAU_TRY(one(out->q, (int64_t)B * Q, cfg->p_q, SK_Q, nHop));
  RAU_TRY(one(out->m, (int64_t)B * cfg->M, cfg->p_m, SK_M, nHop));
  if (out->x) {
    if (hop_rows_path(ctx, cfg)) {   // drawn inline by the rows pack kernels (16-bit draws shared by channel pairs)
      RAU_TRY(rau_check_dev(out->x, "mask output"));
      RAU_TRY(k_xmask16_bytes(ctx, out->x, B, cfg->C, cfg->S, nHop, cfg->p_x, stream_of(0, SK_X, 0, rank)));
    } else {
      RAU_TRY(one(out->x, (int64_t)B * cfg->C * cfg->S, cfg->p_x, SK_X, nHop));
    }
  }
  return RAU_OK;
}

int rau_noise_clip(rau_ctx* ctx, const rau_config* cfg, float* const grads[3], int64_t step_t, float eta, float gamma,
                   float clip, const float* const noise_override[3], float* norms) {
  RAU_REQUIRE(ctx, "ctx == NULL");
  RAU_TRY(rau_check_cfg(cfg));
  RAU_CHECK_CUDA(cudaSetDevice(ctx->device));
  ctx->epoch++;
  ARENA(norm2, double, "opt.norm2", 4);
  RAU_CHECK_CUDA(cudaMemsetAsync(norm2, 0, sizeof(double) * 4, ctx->stream));
  const float std_ = (eta > 0.0f && gamma > 0.0f) ? sqrtf(eta / ((float)(step_t + 1) * gamma)) : 0.0f;   // F:617-618
  for (int g = 0; g < 3; ++g) {
    RAU_TRY(rau_check_dev(grads[g], "grads[g]"));
    const int64_t n = rau_group_size(cfg, g);
    RAU_TRY(k_noise_norm(ctx, grads[g], n, std_, noise_override ? noise_override[g] : nullptr, ctx->seed,
                         stream_of(step_t, SK_NOISE, g, 0), norm2 + g));
  }
  for (int g = 0; g < 3; ++g)
    RAU_TRY(k_clip_optim(ctx, -1, rau_group_size(cfg, g), nullptr, grads[g], norm2 + g, clip, 0, 0, 0, 0, nullptr, nullptr, 0,
                         norms ? norms + g : nullptr));
  return RAU_OK;
}

int rau_optim_step(rau_ctx* ctx, int optim, int64_t n, float* x, const float* dx, float lr, float h0, float h1, float h2,
                   float* state0, float* state1, int64_t t) {
  RAU_REQUIRE(ctx, "ctx == NULL");
  RAU_REQUIRE(optim >= RAU_OPT_SGD && optim <= RAU_OPT_ADAM, "unknown optimizer %d", optim);
  RAU_REQUIRE(n > 0, "n = %lld", (long long)n);
  RAU_TRY(rau_check_dev(x, "x")); RAU_TRY(rau_check_dev(dx, "dx"));
  if (optim != RAU_OPT_SGD) RAU_TRY(rau_check_dev(state0, "state0"));
  if (optim == RAU_OPT_ADAM) RAU_TRY(rau_check_dev(state1, "state1"));
  RAU_CHECK_CUDA(cudaSetDevice(ctx->device));
  ctx->epoch++;
  return k_clip_optim(ctx, optim, n, x, (float*)dx, nullptr, 0.0f, lr, h0, h1, h2, state0, state1, t, nullptr);
}

static int train_enqueue(rau_ctx* ctx, const rau_config* cfg, const rau_batch* bt, float* const params[3],
                         float* const grads[3], float* const opt_state[3][2], const float* hop_mask, const rau_masks* masks,
                         const rau_train_hparams* hp, const rau_step_out* out) {
  const bool dp = rau_comm_attached(ctx);
  ARENA(norm2, double, "opt.norm2", 4);
  RAU_CHECK_CUDA(cudaMemsetAsync(norm2, 0, sizeof(double) * 4, ctx->stream));
  const float std_flag = (hp->eta > 0.0f && hp->gamma > 0.0f) ? 1.0f : 0.0f;   // the value itself is StepState.noise_std
  auto finish_group = [&](int g) -> int {   // gradient noise + norm, then per-group clip + the optimizer rule (F:617-648, F:788-790)
    RAU_TRY(k_noise_norm(ctx, grads[g], rau_group_size(cfg, g), std_flag, hp->noise_override ? hp->noise_override[g] : nullptr,
                         ctx->seed, stream_of(0, SK_NOISE, g, 0), norm2 + g));
    return k_clip_optim(ctx, hp->optim, rau_group_size(cfg, g), params[g], grads[g], norm2 + g, hp->clip, hp->lr[g], hp->h0,
                        hp->h1, hp->h2, opt_state ? opt_state[g][0] : nullptr, opt_state ? opt_state[g][1] : nullptr, 1,
                        (out && out->norms) ? out->norms + g : nullptr, g);
  };
  ctx->early_tail_done = nullptr;
  ctx->early_tail = [&]() -> int {   // (called by feval_enqueue with ctx->stream = aux, once group 2's gradients are final)
      if (dp) RAU_TRY(rau_allreduce_internal(ctx, grads[2], rau_group_size(cfg, 2)));
      return finish_group(2);
    };
  ctx->early_tail0_ran = false;
  ctx->early_tail0 = [&]() -> int {   // (called by feval_enqueue on the chain between the encoder backward and the join)
      if (dp) RAU_TRY(rau_allreduce_internal(ctx, grads[0], rau_group_size(cfg, 0)));
      return finish_group(0);
    };
  const int rc_f = feval_enqueue(ctx, cfg, bt, params, grads, hop_mask, masks, out);
  ctx->early_tail = nullptr;
  ctx->early_tail0 = nullptr;
  RAU_TRY(rc_f);
  cudaEvent_t early = ctx->early_tail_done;
  ctx->early_tail_done = nullptr;
  const bool todo[3] = {!ctx->early_tail0_ran, true, early == nullptr};   // groups whose tail has not gone out yet
  ctx->early_tail0_ran = false;
  if (dp) {   // data parallel: one sum over ranks of each flat gradient, the remaining ones as one launch
    RAU_TRY(rau_allreduce_group(ctx, 1));
    for (int g = 0; g < 3; ++g)
      if (todo[g]) RAU_TRY(rau_allreduce_internal(ctx, grads[g], rau_group_size(cfg, g)));
    if (out && out->loss) RAU_TRY(rau_allreduce_internal(ctx, out->loss, cfg->nHop + 2));
    if (out && out->loss_do_pred) RAU_TRY(rau_allreduce_internal(ctx, out->loss_do_pred, cfg->nHop));
    RAU_TRY(rau_allreduce_group(ctx, 0));
  }
  for (int g = 0; g < 3; ++g)
    if (todo[g]) RAU_TRY(finish_group(g));
  if (early) RAU_CHECK_CUDA(cudaStreamWaitEvent(ctx->stream, early, 0));
  rau_phase_mark(ctx, "noise + clip + optimizer");
  return RAU_OK;
}

int rau_train_step(rau_ctx* ctx, const rau_config* cfg, const rau_batch* bt, float* const params[3], float* const grads[3],
                   float* const opt_state[3][2], const float* hop_mask, const rau_masks* masks, int64_t step_t,
                   const rau_train_hparams* hp, const rau_step_out* out) {
  RAU_REQUIRE(hp != nullptr, "hparams == NULL");
  RAU_REQUIRE(hp->optim >= RAU_OPT_SGD && hp->optim <= RAU_OPT_ADAM, "unknown optimizer %d", hp->optim);
  RAU_TRY(feval_validate(ctx, cfg, bt, params, grads));
  RAU_TRY(rau_check_async_error(ctx));   // a failed recurrence of an EARLIER step surfaces here without a device sync
  if (hp->optim != RAU_OPT_SGD) RAU_REQUIRE(opt_state && opt_state[0][0], "optimizer state is required");
  // per-step scalars (evaluated in double on the host like Lua numbers, OU:80-83)
  StepState st;
  st.step = (unsigned long long)step_t;
  st.noise_std = (hp->eta > 0.0f && hp->gamma > 0.0f) ? (float)sqrt((double)hp->eta / ((double)(step_t + 1) * hp->gamma)) : 0.0f;
  for (int g = 0; g < 3; ++g) {
    double stp = hp->lr[g];
    if (hp->optim == RAU_OPT_ADAM) {
      const double t = hp->opt_t > 0 ? (double)hp->opt_t : (double)(step_t + 1);   // state.t after its increment (OU:79)
      stp = (double)hp->lr[g] * sqrt(1.0 - pow((double)hp->h1, t)) / (1.0 - pow((double)hp->h0, t));
    }
    st.opt_step[g] = (float)stp;
  }
  RAU_TRY(upload_step_state(ctx, st));

  // whole-step CUDA graph: after two eager runs with identical arguments (every arena buffer is then sized), the
  // enqueue sequence is captured once and replayed; only the StepState upload above changes between replays
  RauGraph& gr = ctx->graph;
  if (ctx->phases < 0) rau_phase_mark(ctx, "init");
  const bool graphable = !gr.disabled && masks == nullptr && hp->noise_override == nullptr && ctx->phases != 1;
  std::vector<uint64_t> key;
  RauGraphEntry* ent = nullptr;
  if (graphable) {
    auto put = [&](const void* p, size_t n) {
      const unsigned char* c = (const unsigned char*)p;
      for (size_t i = 0; i < n; i += 8) {
        uint64_t v = 0;
        memcpy(&v, c + i, n - i < 8 ? n - i : 8);
        key.push_back(v);
      }
    };
    rau_train_hparams hk = *hp;
    hk.opt_t = 0;   // per-step scalar: travels in StepState, not in the captured launch arguments
    put(cfg, sizeof(*cfg)); put(bt, sizeof(*bt)); put(&hk, sizeof(hk));
    for (int g = 0; g < 3; ++g) {
      key.push_back((uint64_t)(uintptr_t)params[g]); key.push_back((uint64_t)(uintptr_t)grads[g]);
      key.push_back((uint64_t)(uintptr_t)(opt_state ? opt_state[g][0] : nullptr));
      key.push_back((uint64_t)(uintptr_t)(opt_state ? opt_state[g][1] : nullptr));
    }
    if (out) put(out, sizeof(*out));
    for (int h = 0; h < cfg->nHop; ++h) { float v = hop_mask ? hop_mask[h] : 1.0f; uint32_t u; memcpy(&u, &v, 4); key.push_back(u); }
    key.push_back((uint64_t)ctx->precision); key.push_back((uint64_t)ctx->seed); key.push_back((uint64_t)(uintptr_t)ctx->comm);
    // a re-allocated arena buffer (a larger validation batch, another configuration) invalidates every captured step:
    // their kernels carry the freed addresses
    if (gr.arena_generation != ctx->arena.generation) { gr.clear(); gr.arena_generation = ctx->arena.generation; }
    if (gr.entries.size() > 16 && gr.entries.find(key) == gr.entries.end()) gr.clear();
    ent = &gr.entries[key];
    if (ent->exec) {
      RAU_CHECK_CUDA(cudaGraphLaunch(ent->exec, ctx->stream));
      ctx->launches += ent->kernels;
      return RAU_OK;
    }
    ent->seen++;
  }
  if (graphable && ent->seen > 2) {
    cudaStream_t user = ctx->stream;
    // the capture stream must see the StepState upload and everything queued before it
    RAU_CHECK_CUDA(cudaEventRecord(ctx->ev0, user));
    RAU_CHECK_CUDA(cudaStreamWaitEvent(ctx->gstream, ctx->ev0, 0));
    ctx->stream = ctx->gstream;
    const int64_t l0 = ctx->launches;
    cudaGraph_t graph = nullptr;
    int r = RAU_OK;
    if (cudaStreamBeginCapture(ctx->gstream, cudaStreamCaptureModeRelaxed) != cudaSuccess) {
      cudaGetLastError();
      ctx->stream = user;
      gr.disabled = true;
    } else {
      ctx->ss_active = ctx->d_ss;
      r = train_enqueue(ctx, cfg, bt, params, grads, opt_state, hop_mask, masks, hp, out);
      ctx->ss_active = nullptr;
      cudaError_t e = cudaStreamEndCapture(ctx->gstream, &graph);
      ctx->stream = user;
      if (r != RAU_OK || e != cudaSuccess || graph == nullptr) {
        cudaGetLastError();
        if (graph) cudaGraphDestroy(graph);
        gr.disabled = true;            // fall back to eager enqueue for good
        ctx->launches = l0;
        if (r != RAU_OK) return r;
      } else {
        e = cudaGraphInstantiate(&ent->exec, graph, 0);
        cudaGraphDestroy(graph);
        if (e != cudaSuccess) {
          cudaGetLastError();
          ent->exec = nullptr;
          gr.disabled = true;
          ctx->launches = l0;
        } else {
          ent->kernels = ctx->launches - l0;
          RAU_CHECK_CUDA(cudaGraphLaunch(ent->exec, ctx->stream));
          return RAU_OK;
        }
      }
    }
  }
  ctx->ss_active = ctx->d_ss;
  const int r = train_enqueue(ctx, cfg, bt, params, grads, opt_state, hop_mask, masks, hp, out);
  ctx->ss_active = nullptr;
  return r;
}

static int predict_enqueue(rau_ctx* ctx, const rau_config* cfg, const rau_batch* bt, float* const params[3], float* pred, float* att) {
  RAU_REQUIRE(ctx, "ctx == NULL");
  RAU_TRY(rau_check_async_error(ctx));
  RAU_TRY(rau_check_cfg(cfg));
  RAU_REQUIRE(cfg->nlayer == 2, "the fused encoder supports nlayer == 2 (F:209), got %d", cfg->nlayer);
  RAU_TRY(check_batch(cfg, bt, false));
  RAU_TRY(rau_check_dev(pred, "pred"));
  for (int g = 0; g < 3; ++g) RAU_TRY(rau_check_dev(params[g], "params[g]"));
  RAU_CHECK_CUDA(cudaSetDevice(ctx->device));
  ctx->epoch++;
  const int B = bt->B, nHop = cfg->nHop, N = cfg->N, S = cfg->S, H = cfg->H;
  Encoder en;
  RAU_TRY(encoder_alloc(ctx, cfg, B, &en));
  RAU_TRY(encoder_forward(ctx, cfg, bt, params[0], params[1], 0, nullptr, 0, &en));
  const size_t sv_bytes = hop_saved_layout(cfg, B, nullptr, nullptr);
  ARENA(sv_base, char, "step.saved", sv_bytes);
  ARENA(c_all, float, "step.c", (size_t)2 * B * H);
  ARENA(h_all, float, "step.h", (size_t)2 * B * H);
  ARENA(dop, float, "step.dop", (size_t)nHop * B);
  ARENA(att_own, float, "pred.att", (size_t)(nHop + 2) * B * S);
  float* attp = att ? att : att_own;
  RAU_TRY(k_fill(ctx, c_all, (int64_t)B * H, 0.0f));
  RAU_TRY(k_fill(ctx, h_all, (int64_t)B * H, 0.0f));
  MultT<const float*> P = mult_views<const float*, const float>(cfg, params[2]);
  HopSaved sv;
  hop_saved_layout(cfg, B, sv_base, &sv);
  // eval mode has no dropout, so the image side of a hop that does not see the recurrent state -- the feature transpose,
  // I = tanh(Wi X + bi) and Z = I Wa^T -- is the same for every hop: formed once (nHop x less heavy work)
  HopAsync shared;
  const bool hoist = hop_rows_path(ctx, cfg);
  if (hoist) {
    RAU_TRY(hop_forward_pre(ctx, cfg, B, P, bt->feats, 0, sv));
    shared.pre_skip = 1;
  }
  for (int hp = 0; hp < nHop; ++hp)
    RAU_TRY(hop_forward(ctx, cfg, B, P, en.rnn_out, bt->feats, c_all + (size_t)(hp & 1) * B * H, h_all + (size_t)(hp & 1) * B * H,
                        0, sv, pred + (size_t)hp * B * N, dop + (size_t)hp * B, attp + (size_t)hp * B * S,
                        c_all + (size_t)((hp + 1) & 1) * B * H, h_all + (size_t)((hp + 1) & 1) * B * H, hoist ? &shared : nullptr));
  // uni = mean over hops, select = first hop with do_pred > 0.5, forced at the last hop (F:699-721)
  RAU_TRY(k_merge_preds(ctx, nHop, B, N, S, pred, dop, attp, nullptr, nullptr, 1, 0.0f, nullptr, nullptr, nullptr,
                        pred + (size_t)nHop * B * N, pred + (size_t)(nHop + 1) * B * N, attp + (size_t)nHop * B * S,
                        attp + (size_t)(nHop + 1) * B * S));
  return RAU_OK;
}

int rau_predict(rau_ctx* ctx, const rau_config* cfg, const rau_batch* bt, float* const params[3], float* pred, float* att) {
  return predict_enqueue(ctx, cfg, bt, params, pred, att);
}

// predict_result followed by the answer extraction of the test loop (F:903-918) on the device: open-ended answers
// (argmax over the N answers) and, when the batch carries multiple-choice candidates, the masked argmax
int rau_predict_answers(rau_ctx* ctx, const rau_config* cfg, const rau_batch* bt, float* const params[3], const float* mc_choices,
                        int nmc, float* oe_answers, float* mc_answers, float* pred, float* att) {
  RAU_REQUIRE(ctx && cfg && bt, "ctx/cfg/batch == NULL");
  RAU_TRY(rau_check_dev(oe_answers, "oe_answers"));
  if (mc_choices) {
    RAU_REQUIRE(nmc > 0 && nmc <= 32, "nmc = %d (1..32)", nmc);
    RAU_TRY(rau_check_dev(mc_choices, "mc_choices"));
    RAU_TRY(rau_check_dev(mc_answers, "mc_answers"));
  }
  float* pr = pred;
  if (pr == nullptr) {
    RAU_TRY(rau_check_cfg(cfg));
    RAU_REQUIRE(bt->B > 0, "batch size %d", bt->B);
    RAU_TRY(ctx->arena.get("pred.scores", sizeof(float) * (size_t)(cfg->nHop + 2) * bt->B * cfg->N, (void**)&pr));
  }
  RAU_TRY(predict_enqueue(ctx, cfg, bt, params, pr, att));
  return k_answers(ctx, (cfg->nHop + 2) * bt->B, bt->B, cfg->N, pr, mc_choices, nmc, oe_answers, mc_answers);
}

int rau_time_iembed(rau_ctx* ctx, const rau_config* cfg, int B, const float* mult_params, const float* X, int iters,
                    float* ms_per_launch) {
  RAU_REQUIRE(ctx && ms_per_launch, "ctx/ms == NULL");
  RAU_TRY(rau_check_cfg(cfg));
  RAU_REQUIRE(B > 0 && iters > 0, "B=%d iters=%d", B, iters);
  RAU_TRY(rau_check_dev(mult_params, "mult_params")); RAU_TRY(rau_check_dev(X, "X"));
  RAU_CHECK_CUDA(cudaSetDevice(ctx->device));
  ctx->epoch++;
  const int Sp = rau_sp(cfg->S), M = cfg->M, C = cfg->C, S = cfg->S;
  MultT<const float*> P = mult_views<const float*, const float>(cfg, mult_params);
  const bool tc = ctx->precision != RAU_PREC_F32 && S % 4 == 0;
  const int f16 = (prec_x_f16(ctx) && hop_rows_path(ctx, cfg)) ? 1 : 0;   // Xd / Wi fp16
  const int of16 = (prec_img_f16(ctx) && hop_rows_path(ctx, cfg)) ? 1 : 0;  // I fp16
  const bool x3 = prec_x3(ctx) && !f16;
  if (tc && rows_path_enabled() && C % 64 == 0 && M % 64 == 0) {
    // the training launch of the rows engine: packed dropped-out features in, tanh epilogue, packed (hi, lo) I out
    const int R = B * S;
    ARENA(Xh, bf16, "time.Xh", (size_t)R * C);
    ARENA(Xl, bf16, "time.Xl", (size_t)R * C);
    ARENA(Ih, bf16, "time.Ih", (size_t)R * M);
    ARENA(Il, bf16, "time.Il", (size_t)R * M);
    const bf16 *Wi_h, *Wi_l;
    RAU_TRY(rows_pack(ctx, P.Wi, (int64_t)M * C, x3, true, nullptr, &Wi_h, &Wi_l, f16));
    RAU_TRY(k_xprep_rows(ctx, X, B, C, S, nullptr, 1.0f, Xh, x3 ? Xl : nullptr, 0, 0.0f, 0, f16));
    RowsGemm rg;
    rg.M = R; rg.N = M; rg.K = C;
    rg.A.hi = Xh; rg.A.lo = x3 ? Xl : nullptr; rg.A.ld = C;
    rg.B.hi = Wi_h; rg.B.lo = Wi_l; rg.B.ld = C;
    rg.epi = ROWS_EPI_TANH; rg.bias = P.bi; rg.f16 = f16; rg.of16 = of16;
    rg.out_hi = Ih; rg.out_lo = (prec_x3(ctx) && !of16) ? Il : nullptr; rg.ldo = M;
    RAU_TRY(rows_gemm(ctx, rg));   // warm-up
    RAU_CHECK_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
    for (int i = 0; i < iters; ++i) RAU_TRY(rows_gemm(ctx, rg));
    RAU_CHECK_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
    RAU_CHECK_CUDA(cudaEventSynchronize(ctx->ev1));
    float ms = 0.0f;
    RAU_CHECK_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
    *ms_per_launch = ms / iters;
    return RAU_OK;
  }
  ARENA(I, float, "time.I", (size_t)B * M * Sp);
  SimtGemm g;
  g.M = M; g.N = Sp; g.K = C;
  g.A = P.Wi; g.sam = C; g.sak = 1; g.a_const = 1;
  g.sbk = Sp; g.sbn = 1; g.bB = (int64_t)C * Sp;
  g.C = I; g.scm = Sp; g.scn = 1; g.bC = (int64_t)M * Sp;
  g.batch = B; g.bias_m = P.bi; g.act = 1; g.n_valid = S;
  if (tc) {   // exactly the training launch: packed operand in, fp32 + packed (hi, lo) result out
    ARENA(Xh, bf16, "time.Xh", (size_t)B * C * Sp);
    ARENA(Xl, bf16, "time.Xl", (size_t)B * C * Sp);
    ARENA(Ih, bf16, "time.Ih", (size_t)B * M * Sp);
    ARENA(Il, bf16, "time.Il", (size_t)B * M * Sp);
    RAU_TRY(k_dropout_pack(ctx, X, (int64_t)B * C, S, nullptr, 1.0f, Xh, x3 ? Xl : nullptr, Sp));
    g.B_hi = Xh; g.B_lo = Xl; g.C_hi = Ih; g.C_lo = Il;
  } else {
    ARENA(Xd, float, "hop.Xd", (size_t)B * C * Sp);
    RAU_TRY(k_dropout(ctx, X, (int64_t)B * C, S, S, nullptr, 1.0f, Xd, Sp, nullptr, 0, Sp));
    g.B = Xd;
  }
  RAU_TRY(rau_contract(ctx, g));   // warm-up
  RAU_CHECK_CUDA(cudaEventRecord(ctx->ev0, ctx->stream));
  for (int i = 0; i < iters; ++i) RAU_TRY(rau_contract(ctx, g));
  RAU_CHECK_CUDA(cudaEventRecord(ctx->ev1, ctx->stream));
  RAU_CHECK_CUDA(cudaEventSynchronize(ctx->ev1));
  float ms = 0.0f;
  RAU_CHECK_CUDA(cudaEventElapsedTime(&ms, ctx->ev0, ctx->ev1));
  *ms_per_launch = ms / iters;
  return RAU_OK;
}

}  // extern "C"

int rau_contract(rau_ctx* ctx, const SimtGemm& g) {
  if (ctx->precision != RAU_PREC_F32) {
    const int rr = rows_contract_try(ctx, g);   // nn.Linear-shaped products: the persistent rows engine
    if (rr < 0) return rr;
    if (rr == 1) return RAU_OK;
    const int r = tc_gemm_try(ctx, g);
    if (r < 0) return r;
    if (r == 1) return RAU_OK;
    if (g.A_hi || g.B_hi) {
      rau_set_error("internal: a packed-operand product was refused by the tcgen05 engine");
      return RAU_EINVAL;
    }
  }
  RAU_TRY(simt_gemm(ctx, g));
  if (ctx->precision != RAU_PREC_F32 && g.C_hi && g.scn == 1 && g.batch == 1)   // keep the result's packed twin in step
    RAU_TRY(rows_pack_into(ctx, g.C, g.scm, g.M, g.N, g.C_hi, prec_x3(ctx) ? g.C_lo : nullptr, g.scm));
  return RAU_OK;
}
