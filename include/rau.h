/* rau.h -- C ABI of librau.so: the B200 (sm_100a) implementation of the RAU_VQA
 * recurrent-answering-unit training/inference hot path.
 *
 * This header is the whole boundary.  It is plain C (no C++ types, no callbacks, no torch
 * types) so that the same text can be handed to LuaJIT `ffi.cdef` (the reference's host
 * language, see lua/ and INTEGRATION.md) and to Python cffi (tests/, bench.py).
 *
 * Reference interfaces replaced (paths relative to the reference repo root):
 *   F:  experiments/Ours_Full/LstmAttCtrlGradNoiseDontSelect.lua
 *   A:  model/ATTLSTM.lua        D:  model/DeepLSTM.lua
 *   OU: utils/optim_updates.lua  MU: utils/model_utils.lua
 *
 * Conventions
 *   - every tensor is a caller-owned DEVICE pointer, float32, row-major, batch-first,
 *     contiguous unless a leading dimension `ld*` is given; the library never frees or
 *     retains caller pointers past the call (nn.Module clones re-point storages, F:322-347);
 *   - token ids / labels are 1-based and delivered as float, like the reference (F:454-455);
 *   - all calls enqueue on the ctx stream and return immediately; scalars are written to
 *     device memory and are valid after rau_sync();
 *   - return 0 on success, a negative rau_status otherwise; text via rau_last_error();
 *   - there is NO CPU fallback: host pointers or a non-sm_100 device are errors.
 */
#ifndef RAU_H
#define RAU_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct rau_ctx rau_ctx;

typedef enum {
  RAU_OK = 0,
  RAU_EINVAL = -1,   /* bad shape / null / misaligned / host pointer */
  RAU_ECUDA = -2,    /* CUDA runtime or driver error */
  RAU_ENCCL = -3,    /* NCCL error or libnccl not loadable */
  RAU_EARCH = -4,    /* device is not sm_100 (B200) */
  RAU_ENOMEM = -5,
  RAU_ESTATE = -6    /* call sequence error (e.g. bwd without fwd) */
} rau_status;

/* gate chunk order of the 4H pre-activation vector */
typedef enum {
  RAU_GATES_IFOG = 0,  /* model/DeepLSTM.lua:47-54  [in|forget|out|transform] */
  RAU_GATES_IGFO = 1   /* model/ATTLSTM.lua:12-19   [in|transform|forget|out] */
} rau_gate_order;

/* arithmetic mode of the dense contractions (pointwise math, cell state, softmax, losses and all
 * weight-gradient accumulators are float32 in every mode) */
typedef enum {
  RAU_PREC_F32 = 0,     /* fp32 operands, CUDA-core FMA: the exact mode */
  RAU_PREC_BF16 = 1,    /* bf16 operands, tcgen05 MMA, fp32 accumulate in TMEM: the fast mode (~3e-3 relative) */
  RAU_PREC_BF16X3 = 2,  /* bf16 hi/lo split operands (hi*hi + hi*lo + lo*hi, 3 MMA passes into one TMEM accumulator):
                         * ~2e-5 relative on tcgen05 */
  RAU_PREC_MIXED = 3,   /* the DEFAULT.  The two feature-width products of an answering unit -- I = tanh(Wi drop(X) + bi) and
                         * gWi += dY^T drop(X): 57 % (C = 512) to 87 % (C = 2048) of the step's flops -- run as ONE fp16 pass
                         * (11-bit significands: operand rounding 2^-12, eight times finer than bf16; dY is carried times a
                         * power of two so that it sits in fp16's normal range and gWi's fp32 result is scaled back).
                         * Everything else -- Z = I Wa^T, dY, gWa, the recurrent chain, the encoder -- stays bf16x3. */
  RAU_PREC_F16IMG = 4   /* the whole image side of an answering unit as single fp16 planes (I, Z, dZ, dY, gWa as well): a third
                         * of bf16x3's tensor work and half its activation bytes, at 3e-4 .. 1.1e-3 per-tensor error against
                         * the float64 oracle -- NOT inside the 1e-3 bar with margin, hence opt-in */
} rau_precision;

typedef enum {
  RAU_OPT_SGD = 0,      /* OU:7-9   */
  RAU_OPT_SGDM = 1,     /* OU:11-19 */
  RAU_OPT_SGDMOM = 2,   /* OU:21-31 */
  RAU_OPT_ADAGRAD = 3,  /* OU:33-43 */
  RAU_OPT_RMSPROP = 4,  /* OU:46-57 */
  RAU_OPT_ADAM = 5      /* OU:59-87 */
} rau_optim;

/* The constants the experiment scripts hard-code (F:202-229) plus run-time sizes. */
typedef struct {
  int V;        /* vocab_size                 F:204 */
  int embed;    /* embed_dim = 200            F:202 */
  int Hq;       /* rnn_size = 512             F:208 */
  int nlayer;   /* nrnn_layer = 2             F:209 (only 2 is supported by the fused encoder) */
  int C;        /* cnnout_dim 512 / 2048      F:216, RN:217 */
  int S;        /* cnnout_w*cnnout_h = 196    F:219 */
  int M;        /* multfeat_dim = 512         F:220 */
  int A;        /* attfeat_dim = 256          F:221 */
  int H;        /* att_rnn_size = 512         F:225 */
  int N;        /* netout_dim = answer_size   F:222 */
  int nHop;     /* -nhop                      F:53  */
  int T;        /* seq_len (<= 26)            LD:1418 */
  float p_embed, p_rnn, p_q, p_x, p_m;   /* dropout rates F:205, F:210, F:233, F:239, F:277 */
} rau_config;

/* ------------------------------------------------------------------ context ------------- */
int rau_version(void);
const char* rau_last_error(void);                      /* thread-local message of the last failure */
int rau_ctx_create(rau_ctx** out, int device, void* cuda_stream /* cudaStream_t or NULL */);
int rau_ctx_destroy(rau_ctx* ctx);
int rau_set_stream(rau_ctx* ctx, void* cuda_stream);
int rau_set_seed(rau_ctx* ctx, uint64_t seed);         /* Philox key for dropout masks and gradient noise */
int rau_set_precision(rau_ctx* ctx, int precision);    /* rau_precision */
int rau_get_precision(rau_ctx* ctx);
int rau_sync(rau_ctx* ctx);
/* debugging aid: with RAU_PHASES=1 in the environment every step runs eagerly with an event at each phase boundary
 * (encoder, answering units, loss, BPTT, weight gradients, update); this prints the split to stderr and clears it */
int rau_phase_report(rau_ctx* ctx);
/* number of kernels this library launched on ctx since creation (bench.py's gpu_launches) */
int64_t rau_launch_count(rau_ctx* ctx);

/* flat parameter layout (our own; nngraph's order is not recoverable):
 *   group 0 "embed": E[V,embed]
 *   group 1 "rnn"  : per layer L: Wi[4Hq,in] bi[4Hq] Wh[4Hq,Hq] bh[4Hq]
 *   group 2 "mult" : Wq bq Wh bh Wi bi Wqa bqa Wa ba ws Wm bm Wp bp Wx bx Whh bhh Wo bo Ws bso wd bs bd
 *                    (the two scalars last, so that every tensor starts 16-byte aligned when N % 4 == 0) */
int64_t rau_group_size(const rau_config* cfg, int group);
/* offset (in floats) of a named tensor inside its group, -1 if unknown; e.g. ("mult","Wi") */
int64_t rau_param_offset(const rau_config* cfg, int group, const char* name);

/* ------------------------------------------------------------------ a1/a2: LSTM cell ----- */
/* One LSTM layer step: G = x Wi^T + bi + h_prev Wh^T + bh; gates per `gate_order`;
 * c = f*c_prev + i*g; h = o*tanh(c).  Replaces the nngraph built by lstm() A:4-28 and by the
 * loop body D:33-65.  ld* are row pitches in floats (DeepLSTM keeps c/h as Narrow views of a
 * [B, 2*Hq*nlayer] state, D:23-24).  `saved` receives 5*B*H floats (i,f,o,g,tanh(c)). */
typedef struct {
  int B, in_size, H, gate_order;
  int ldx, ldc_prev, ldh_prev, ldc, ldh;
} rau_lstm_desc;
size_t rau_lstm_saved_bytes(const rau_lstm_desc* d);
int rau_lstm_cell_fwd(rau_ctx* ctx, const rau_lstm_desc* d,
                      const float* x, const float* c_prev, const float* h_prev,
                      const float* Wi, const float* bi, const float* Wh, const float* bh,
                      float* c, float* h, float* saved);
/* updateGradInput + accGradParameters(scale) in one call.  dc/dh are the gradOutput pair (either may
 * be NULL = zeros); dh_extra (NULL or contiguous [B,H]) is added to dh: in a stack, next_h feeds both the
 * module output and the layer above (D:39, A:52).  dx/dc_prev/dh_prev are the gradInput triple
 * (overwritten); gW* accumulate (+= scale * ...), any of them may be NULL to skip. */
int rau_lstm_cell_bwd(rau_ctx* ctx, const rau_lstm_desc* d,
                      const float* x, const float* c_prev, const float* h_prev,
                      const float* Wi, const float* Wh, const float* saved,
                      const float* dc, const float* dh, int lddc, int lddh, const float* dh_extra,
                      float* dx, float* dc_prev, float* dh_prev, int lddx, int lddc_prev, int lddh_prev,
                      float* gWi, float* gbi, float* gWh, float* gbh, float scale);

/* ------------------------------------------------------------------ a3: word embedding --- */
/* tanh(dropout(E[x])) for n = rows ids (F:203-206). mask: optional 0/1 keep bytes [n,embed]
 * (parity tests); NULL + train!=0 draws Philox masks with `stream_id`; train==0 is evaluate(). */
int rau_embed_fwd(rau_ctx* ctx, const rau_config* cfg, int n, const float* ids, const float* E,
                  int train, const uint8_t* mask, uint64_t stream_id, float* out);
int rau_embed_bwd(rau_ctx* ctx, const rau_config* cfg, int n, const float* ids, const float* out,
                  int train, const uint8_t* mask, uint64_t stream_id, const float* dout, float* gE);

/* nn.Dropout (v2) on n floats: y = x * keep / (1-p) in training, y = x in evaluate() (A:52, D:39, F:205).
 * The backward pass of Dropout is the same call on the gradient with the same mask / stream_id. */
int rau_dropout(rau_ctx* ctx, int64_t n, const float* x, float p, int train, const uint8_t* mask,
                uint64_t stream_id, float* y);

/* ------------------------------------------------------------------ fused step ----------- */
/* Optional explicit dropout masks (0/1 keep bytes) for parity runs; any NULL member = Philox. */
typedef struct {
  const uint8_t* embed;   /* [T,B,embed] */
  const uint8_t* rnn;     /* [T,B,Hq]    (input of encoder layer 2, D:39) */
  const uint8_t* q;       /* [nHop,B,Q]  Q = 2*Hq*nlayer */
  const uint8_t* x;       /* [nHop,B,C,S] */
  const uint8_t* m;       /* [nHop,B,M]  */
} rau_masks;

/* The keep masks the step with iteration number step_t draws (Philox, this context's seed and rank) as 0/1 bytes in the
 * rau_masks layouts above; NULL members are skipped.  Test hook: the parity tests run the step with drawn masks -- the
 * schedule bench.py measures -- and give these bytes to the CPU oracle. */
typedef struct {
  uint8_t* embed;   /* [T,B,embed] */
  uint8_t* rnn;     /* [T,B,Hq]    */
  uint8_t* q;       /* [nHop,B,Q]  */
  uint8_t* x;       /* [nHop,B,C,S] */
  uint8_t* m;       /* [nHop,B,M]  */
} rau_masks_out;
int rau_draw_masks(rau_ctx* ctx, const rau_config* cfg, int B, int64_t step_t, const rau_masks_out* out);

typedef struct {
  int B;                    /* local batch */
  int B_global;             /* loss/gradient normaliser (= B on one GPU; world*B under data parallel) */
  const float* feats;       /* [B,C,14,14]   F:451-453 */
  const float* tokens;      /* [T,B] 1-based, pad = 1 */
  const float* lengths;     /* [B]   question lengths */
  int max_len;              /* host-known x_len:max() (F:460); 0 = run all T steps (same result) */
  const float* labels;      /* [B]   1-based answers */
  const void* feats_f16;    /* optional: the same features as IEEE fp16 [B,C,14,14] on the device.  rau_train_step / rau_feval
                             * in the default precision mode pack them straight into the fp16 operand of the i_embed
                             * product (the bits the tensor pipe sees are those of a float32 `feats` rounded to fp16); `feats`
                             * may then be NULL.  Entry points that need float32 features (rau_predict*, explicit masks,
                             * schedules without the all-hops feature pack) reject a batch without `feats`. */
} rau_batch;

typedef struct {
  float* loss;          /* [nHop+2]  tab_loss (F:535, F:548, F:557); local-batch sums / B_global */
  float* loss_do_pred;  /* [nHop]    tab_loss_do_pred (F:572) */
  float* answers;       /* [nHop+2,B] argmax per hop, uni, select (1-based), may be NULL */
  float* scores;        /* [nHop,B,N] may be NULL */
  float* attprob;       /* [nHop,B,S] may be NULL */
  float* do_pred;       /* [nHop,B]   may be NULL */
  float* norms;         /* [3] pre-clip L2 norms of embed/rnn/mult grads (F:627-648), may be NULL */
} rau_step_out;

/* feval forward+backward (F:445-615): zeroes the three flat grads, runs the encoder unroll, the nHop
 * answering units, the joint loss and BPTT.  hop_mask[h] (HOST array of nHop floats 0/1, may be NULL =
 * all 1) is tab_multhop_compute_loss (F:587-589).  Gradients are left in grads[3] BEFORE noise/clip so that a
 * data-parallel caller can all-reduce them. */
int rau_feval(rau_ctx* ctx, const rau_config* cfg, const rau_batch* batch,
              float* const params[3], float* const grads[3],
              const float* hop_mask, const rau_masks* masks, int64_t step_t, const rau_step_out* out);

/* gradient noise + per-group clip (F:617-648): g += N(0, sqrt(eta/((step_t+1)*gamma))) (noise_override
 * [3] optional explicit noise tensors; eta<=0 disables), then g *= clip/||g|| when ||g|| > clip.
 * norms (device [3]) may be NULL. */
int rau_noise_clip(rau_ctx* ctx, const rau_config* cfg, float* const grads[3], int64_t step_t,
                   float eta, float gamma, float clip, const float* const noise_override[3], float* norms);

/* utils/optim_updates.lua on one flat vector, in place.  state0/state1 are caller-owned device
 * vectors of n floats (adam: m,v; rmsprop/adagrad/sgdmom: m; sgdm: v); t is the 1-based step count
 * AFTER increment (adam bias correction, OU:80-83).  h0..h2: sgd -; sgdm alpha; sgdmom alpha;
 * adagrad eps; rmsprop alpha,eps; adam beta1,beta2,eps. */
int rau_optim_step(rau_ctx* ctx, int optim, int64_t n, float* x, const float* dx, float lr,
                   float h0, float h1, float h2, float* state0, float* state1, int64_t t);

/* One whole training iteration = rau_feval -> [all-reduce when a communicator is attached] ->
 * rau_noise_clip -> rau_optim_step x3 (F:786-791).  opt_state[g][0..1] as in rau_optim_step.
 * step_t is the `it` the reference passes to feval (F:445, F:787; its main loop counts from 1, F:783): it keys the Philox
 * dropout / noise streams and sets the noise variance eta / ((step_t + 1) * gamma) exactly as F:617-618 does.  Adam's
 * bias correction does NOT use it: the reference keeps its own counter in the optimizer state (OU:79), here hp->opt_t. */
typedef struct {
  int optim;                 /* rau_optim */
  float lr[3];               /* learningrate, learningrate, multlearningrate (F:788-790) */
  float h0, h1, h2;          /* optimizer hyper-parameters (see rau_optim_step) */
  float eta, gamma, clip;    /* F:54-55, F:49 */
  const float* const* noise_override; /* NULL or [3] */
  int64_t opt_t;             /* adam: state.t AFTER its increment (OU:79-83), kept by the caller next to opt_state like the
                              * reference's per-group state table (F:754-756); 0 = step_t + 1 (a run that starts at it = 0) */
} rau_train_hparams;
int rau_train_step(rau_ctx* ctx, const rau_config* cfg, const rau_batch* batch,
                   float* const params[3], float* const grads[3], float* const opt_state[3][2],
                   const float* hop_mask, const rau_masks* masks, int64_t step_t,
                   const rau_train_hparams* hp, const rau_step_out* out);

/* predict_result (F:652-724): evaluate()-mode forward; pred[nHop+2,B,N] and att[nHop+2,B,S]
 * (hops, uni = mean over hops, select = first hop with do_pred>0.5, forced at the last hop). */
int rau_predict(rau_ctx* ctx, const rau_config* cfg, const rau_batch* batch,
                float* const params[3], float* pred, float* att);
/* The test loop's answer extraction on the device (F:903-918): rau_predict, then for each of the nHop+2 prediction tables
 * the open-ended answer argmax_n pred[n] and -- when mc_choices [B, nmc] (1-based candidate ids as next_batch_feat's ans_mc
 * delivers them, 0 = empty slot; nmc <= 32) is given -- the multiple-choice answer argmax_n pred[n] * mask[n] with the
 * reference's multiplicative 0/1 mask.  Answers are 1-based floats [nHop+2, B], ties resolve to the lowest index.
 * pred / att may be NULL when the dense tables are not wanted. */
int rau_predict_answers(rau_ctx* ctx, const rau_config* cfg, const rau_batch* batch, float* const params[3],
                        const float* mc_choices, int nmc, float* oe_answers, float* mc_answers, float* pred, float* att);

/* ------------------------------------------------------------------ module-level hop ----- */
/* protos.multimodal forward (F:292-307) for one hop: {q,X,c,h} -> {score,do_pred,p,c',h'}.
 * `saved` is a caller-owned buffer of rau_hop_saved_bytes() (each Lua clone owns its own). */
size_t rau_hop_saved_bytes(const rau_config* cfg, int B);
int rau_hop_fwd(rau_ctx* ctx, const rau_config* cfg, int B, const float* mult_params,
                const float* q, const float* X, const float* c, const float* h,
                int train, const uint8_t* mask_q, const uint8_t* mask_x, const uint8_t* mask_m,
                uint64_t stream_id,
                float* score, float* do_pred, float* p, float* c_out, float* h_out, void* saved);
/* multimodals[h]:backward (F:590-593): gradOutput {dscore, ddo_pred, dp, dc', dh'} (any NULL = 0),
 * gradInput {dq, dX (NULL to skip: the caller discards it, F:598), dc, dh}; mult grads accumulate. */
int rau_hop_bwd(rau_ctx* ctx, const rau_config* cfg, int B, const float* mult_params, float* mult_grads,
                const float* q, const float* X, const float* c, const float* h, int train, const void* saved,
                const float* dscore, const float* ddo_pred, const float* dp, const float* dc_out, const float* dh_out,
                float* dq, float* dX, float* dc, float* dh);

/* ------------------------------------------------------------------ batch feed */
/* Replaces the synchronous per-step upload of the reference: next_batch_feat hands feval HOST tensors feats[B,C,14,14],
 * x[T,B], x_len[B], y[B] (utils/vqa_prepro_loader.lua:1009) which it casts and copies on the compute stream, B*C*196*4
 * bytes per step (F:452-456).  rau_feed keeps `depth` batches in flight: the loader fills a slot's pinned staging, submit
 * enqueues its upload on a copy stream under the running step, acquire makes the context's stream wait for it and describes
 * the device copy as a rau_batch, release lets a later submit overwrite it.  RAU_FEED_F16 stages the features as fp16 --
 * half the PCIe bytes; in the default precision mode they enter the tensor pipe as fp16 anyway, so the operand the
 * tensor pipe sees does not change by a bit (tests/test_gpu_feed.py).  RAU_FEED_F16 widens the upload to a float32
 * `feats` on the copy stream (every entry point accepts the batch); RAU_FEED_F16_DIRECT skips that pass -- acquire
 * returns `feats` = NULL and `feats_f16` = the uploaded buffer, which the training step's feature pack reads as it is
 * (154 MB less HBM traffic per 256-sample batch beside the step, 51 MB less inside it). */
typedef struct rau_feed rau_feed;
typedef enum { RAU_FEED_F32 = 0, RAU_FEED_F16 = 1, RAU_FEED_F16_DIRECT = 2 } rau_feed_format;
int rau_feed_create(rau_ctx* ctx, const rau_config* cfg, int B, int format, int depth /* 2..8 */, rau_feed** out);
int rau_feed_destroy(rau_feed* feed);
size_t rau_feed_host_bytes(const rau_feed* feed);   /* bytes one submit moves host -> device */
int rau_feed_host_slot(rau_feed* feed, int slot, void** feats, float** tokens, float** lengths, float** labels);
/* the host-side cast of F:452-456 (feats:float()): n values, float64 (src_is_f64 != 0) or float32 -> the staging format */
int rau_feed_convert(const rau_feed* feed, const void* src, int src_is_f64, int64_t n, void* dst);
int rau_feed_submit(rau_feed* feed, int slot);
int rau_feed_acquire(rau_feed* feed, int slot, rau_batch* batch);
int rau_feed_release(rau_feed* feed, int slot);
/* The features of a whole split resident in HBM as fp16 (train2014 at C = 512: 16 GB of the 180 GB): a batch's feature
 * tensor is a gather by image index, no per-step feature upload.  put: HOST float32 [n,C,S] -> images first .. first+n-1
 * (0-based), load time, synchronous; gather: image_index = DEVICE array of B floats, 1-based -> feats[B,C,S] float32. */
typedef struct rau_feat_cache rau_feat_cache;
int rau_feat_cache_create(rau_ctx* ctx, int64_t n_images, int C, int S, rau_feat_cache** out);
int rau_feat_cache_destroy(rau_feat_cache* cache);
int rau_feat_cache_put(rau_feat_cache* cache, int64_t first, int64_t n, const float* host_feats);
int rau_feat_cache_gather(rau_feat_cache* cache, const float* image_index, int B, float* feats);
/* the same without widening: feats_f16[B,C,S] fp16 for rau_batch.feats_f16 (C*S % 8 == 0, 16-byte aligned destination) */
int rau_feat_cache_gather_f16(rau_feat_cache* cache, const float* image_index, int B, void* feats_f16);

/* ------------------------------------------------------------------ multi-GPU - */
/* One process per GPU.  rank 0 calls rau_comm_unique_id, ships the 128 bytes to the other ranks,
 * every rank calls rau_comm_init.  rau_allreduce_grads sums the three flat grads across ranks. */
int rau_comm_unique_id(uint8_t id_out[128]);
int rau_comm_init(rau_ctx* ctx, const uint8_t id[128], int rank, int world);
int rau_comm_destroy(rau_ctx* ctx);
int rau_allreduce_grads(rau_ctx* ctx, float* const grads[3], const int64_t sizes[3]);
int rau_allreduce(rau_ctx* ctx, float* buf, int64_t n);

/* ------------------------------------------------------------------ building blocks ------ */
/* Exposed for unit parity tests and the kernel sweep in bench.py. */
/* C[M,N] (+)= A[M,K] B[N,K]^T on the ctx precision mode's engine; ta/tb: 0 = operand stored
 * [rows,K] (K-major), 1 = stored [K,rows].  accumulate != 0 adds into C. */
int rau_gemm(rau_ctx* ctx, int M, int N, int K, const float* A, int lda, int ta,
             const float* B, int ldb, int tb, float* C, int ldc, int accumulate);
/* D[M,N] (+)= A B^T on the persistent rows-layout tcgen05 engine that carries the answering unit's image-side products
 * (bf16 / bf16x3 modes only).  a_mn/b_mn: 0 = operand stored [rows, ld >= K], 1 = stored [K, ld >= rows]; pitches are
 * multiples of 8 floats.  reduce != 0: split-K partial sums are ADDED into D (TMA reduce-add), else D is overwritten. */
int rau_rows_gemm(rau_ctx* ctx, int M, int N, int K, const float* A, int lda, int a_mn,
                  const float* B, int ldb, int b_mn, float* D, int ldd, int reduce);
/* GPU microseconds per launch of one rows-engine product on synthetic operands (iters launches replayed from a CUDA
 * graph, so host launch cost is excluded); reduce != 0 selects the split-K TMA-reduce form */
int rau_rows_gemm_time(rau_ctx* ctx, int M, int N, int K, int a_mn, int b_mn, int reduce, int iters, float* us_per_launch);
/* The feature dropout + rows transpose + bf16 split of the hop forward (replaces nn.Dropout on the image features,
 * train_rau_vqa.lua:239) with keep bits drawn inline from the context's Philox streams (stream_id ^ hop), for nHop hops:
 * out[h][b*S+s][c] = the packed value (hi + lo) as fp32.  all_hops != 0 runs the one-launch form the training step uses,
 * 0 one launch per hop; both draw identical bits.  Test / inspection hook. */
int rau_feature_pack(rau_ctx* ctx, const float* X, int B, int C, int S, int nHop, float p, uint64_t stream_id, int all_hops,
                     float* out);
/* debugging aid (RAU_ROWS_TRACE=1): copies the per-CTA clock stamps [148][16] of the last rows-engine launch to HOST
 * memory `out` (n 64-bit words): 0 start, 1 prologue done, 2 first TMA issue, 3/4 first/second stage landed, 5 MMAs
 * of the first item issued, 6 first accumulator ready, 7 epilogue issued, 8 stores drained, 9 end */
int rau_rows_trace(rau_ctx* ctx, uint64_t* out, int n);
/* softmax cross-entropy of score[B,N] against 1-based labels: loss_sum += scale*sum_b nll_b,
 * dscore = scale*(softmax - onehot), answers = argmax (1-based, ties -> lowest index). */
int rau_softmax_ce(rau_ctx* ctx, int B, int N, const float* score, const float* labels, float scale,
                   float* loss_sum, float* dscore, float* answers);
/* kernel-only timing hook used by bench.py for the roofline line: runs the attention-hop projection
 * kernel (I = tanh(Wi drop(X) + bi)) `iters` times on the ctx stream between two events and returns the
 * average milliseconds per launch. */
int rau_time_iembed(rau_ctx* ctx, const rau_config* cfg, int B, const float* mult_params,
                    const float* X, int iters, float* ms_per_launch);

/* Attention-kernel sweep (BASELINE.json configs[4]): each kernel of one answering unit that touches the 196 x C feature block
 * or the [B*196, M] activation derived from it, launched alone on synthetic operands at batch B in the context's precision
 * mode; us_out[RAU_SWEEP_COUNT] receives microseconds per launch (CUDA events around every launch; flush_l2 != 0 evicts L2
 * with a 256 MB memset before each one).  bench.py turns these into HBM and tensor-pipe fractions. */
typedef enum {
  RAU_SWEEP_PACK = 0,        /* dropout + transpose to rows + split of X        (F:239)        HBM    */
  RAU_SWEEP_IEMBED = 1,      /* I = tanh(Wi Xd + bi)                            (F:240-241)    tensor */
  RAU_SWEEP_Z = 2,           /* Z = I Wa^T                                      (F:247)        tensor */
  RAU_SWEEP_SCORE = 3,       /* logit = ws . tanh(Z + qatt)                     (F:246-251)    HBM    */
  RAU_SWEEP_SOFTMAX_SUM = 4, /* p = softmax(logit + mem), a = sum_s p_s I[:,s]  (F:285-290, F:254-263) HBM */
  RAU_SWEEP_BWD_DP_DZ = 5,   /* dp = da . I ; softmax / tanh backward -> dZ     (two launches) HBM    */
  RAU_SWEEP_DY = 6,          /* dY = (dZ Wa + da p^T)(1 - I^2), gbi             tensor                */
  RAU_SWEEP_GWA = 7,         /* gWa += dZ^T I                                   tensor                */
  RAU_SWEEP_GWI = 8,         /* gWi += dY^T Xd                                  tensor                */
  RAU_SWEEP_COUNT = 9
} rau_sweep_kernel;
int rau_sweep_attention(rau_ctx* ctx, const rau_config* cfg, int B, const float* mult_params, const float* X, int iters,
                        int flush_l2, float* us_out);

#ifdef __cplusplus
}
#endif
#endif /* RAU_H */
