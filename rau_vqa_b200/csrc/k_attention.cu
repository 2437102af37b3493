// k_attention.cu -- the attention-proper kernels (score conv 256->1, memory logits, softmax over the
// 14x14 grid, weighted sum) on saved projections I=[B,M,Sp], E=[B,A,Sp] (Sp = S padded to 208), plus the
// joint-loss kernels.  These are the HBM-bound part of the hop: ~0.6 MB read per sample, < 1 flop/byte.
// Reference nodes replaced: F:251 (conv 256->1), F:287-289 (Linear + CAddTable + SoftMax),
// F:254-263 (Replicate + CMulTable + Sum(3), which materialises a [B,512,196] temporary), F:310 criterion.
#include "rau_kernels.cuh"

namespace {
template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<bf16>(const bf16* p) { return __bfloat162float(*p); }
template <typename T> __device__ __forceinline__ void stf(T* p, float v);
template <> __device__ __forceinline__ void stf<float>(float* p, float v) { *p = v; }
template <> __device__ __forceinline__ void stf<bf16>(bf16* p, float v) { *p = __float2bfloat16(v); }

constexpr int ATT_T = 256;   // threads per image CTA (>= S)

// one CTA per image: s = ws.E + mem ; p = softmax(s) ; a = I p
template <typename T>
__global__ void __launch_bounds__(ATT_T) attn_fwd_kernel(int M, int A, int S, int Sp, const T* __restrict__ E,
                                                         const T* __restrict__ I, const float* __restrict__ ws,
                                                         const float* __restrict__ mem, float* __restrict__ p_out,
                                                         bf16* __restrict__ p_b, int ldpb, float* __restrict__ a_out,
                                                         bf16* __restrict__ a_b) {
  RAU_PDL_ENTRY();
  extern __shared__ float sm[];
  float* p = sm;            // [Sp]
  float* wsm = sm + Sp;     // [A]
  __shared__ float red[32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const T* Eb = E + (int64_t)b * A * Sp;
  const T* Ib = I + (int64_t)b * M * Sp;
  for (int a = tid; a < A; a += ATT_T) wsm[a] = ws[a];
  __syncthreads();
  float logit = -INFINITY;
  if (tid < S) {
    float acc = 0.0f;
#pragma unroll 4
    for (int a = 0; a < A; ++a) acc = fmaf(wsm[a], ldf<T>(Eb + (int64_t)a * Sp + tid), acc);
    logit = acc + mem[(int64_t)b * S + tid];
  }
  const float mx = block_max(logit, red);
  const float e = tid < S ? __expf(logit - mx) : 0.0f;
  const float den = block_sum(e, red);
  const float pv = e / den;
  if (tid < Sp) p[tid] = tid < S ? pv : 0.0f;
  if (tid < S) p_out[(int64_t)b * S + tid] = pv;
  if (p_b && tid < ldpb) p_b[(int64_t)b * ldpb + tid] = __float2bfloat16(tid < S ? pv : 0.0f);
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nw = ATT_T / 32;
  for (int m = warp; m < M; m += nw) {
    const T* row = Ib + (int64_t)m * Sp;
    float acc = 0.0f;
    for (int s = lane; s < S; s += 32) acc = fmaf(ldf<T>(row + s), p[s], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      a_out[(int64_t)b * M + m] = acc;
      if (a_b) a_b[(int64_t)b * M + m] = __float2bfloat16(acc);
    }
  }
}

// backward of the above for one image:
//   dp = dp_in + I^T da ; ds = p (dp - <p,dp>) ; dZ[a,s] = ws[a] ds[s] (1 - E[a,s]^2)
//   dqa[a] = sum_s dZ[a,s] ; gws_part[b,a] = sum_s ds[s] E[a,s]
template <typename T>
__global__ void __launch_bounds__(ATT_T) attn_bwd_kernel(int M, int A, int S, int Sp, const T* __restrict__ E,
                                                         const T* __restrict__ I, const float* __restrict__ ws,
                                                         const float* __restrict__ p_in, const float* __restrict__ dp_in,
                                                         const float* __restrict__ da, float* __restrict__ ds_out,
                                                         bf16* __restrict__ ds_b, int lddsb, T* __restrict__ dZ,
                                                         float* __restrict__ dqa, bf16* __restrict__ dqa_b,
                                                         float* __restrict__ gws_part, bf16* __restrict__ dZ_hi,
                                                         bf16* __restrict__ dZ_lo) {
  RAU_PDL_ENTRY();
  extern __shared__ float sm[];
  float* ds = sm;          // [Sp]
  float* das = sm + Sp;    // [M]
  __shared__ float red[32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const T* Eb = E + (int64_t)b * A * Sp;
  const T* Ib = I + (int64_t)b * M * Sp;
  T* dZb = dZ ? dZ + (int64_t)b * A * Sp : nullptr;
  bf16* hib = dZ_hi ? dZ_hi + (int64_t)b * A * Sp : nullptr;
  bf16* lob = dZ_lo ? dZ_lo + (int64_t)b * A * Sp : nullptr;
  for (int m = tid; m < M; m += ATT_T) das[m] = da[(int64_t)b * M + m];
  __syncthreads();
  float pv = 0.0f, dpv = 0.0f;
  if (tid < S) {
    float acc = dp_in ? dp_in[(int64_t)b * S + tid] : 0.0f;
#pragma unroll 4
    for (int m = 0; m < M; ++m) acc = fmaf(das[m], ldf<T>(Ib + (int64_t)m * Sp + tid), acc);
    dpv = acc;
    pv = p_in[(int64_t)b * S + tid];
  }
  const float dot = block_sum(pv * dpv, red);
  const float dsv = tid < S ? pv * (dpv - dot) : 0.0f;
  if (tid < Sp) ds[tid] = dsv;
  if (tid < S) ds_out[(int64_t)b * S + tid] = dsv;
  if (ds_b && tid < lddsb) ds_b[(int64_t)b * lddsb + tid] = __float2bfloat16(dsv);
  __syncthreads();
  const int warp = tid >> 5, lane = tid & 31, nw = ATT_T / 32;
  for (int a = warp; a < A; a += nw) {
    const T* row = Eb + (int64_t)a * Sp;
    T* drow = dZb ? dZb + (int64_t)a * Sp : nullptr;
    const float w = ws[a];
    float sz = 0.0f, sg = 0.0f;
    for (int s = lane; s < Sp; s += 32) {
      float dz = 0.0f;
      if (s < S) {
        const float e = ldf<T>(row + s);
        dz = w * ds[s] * (1.0f - e * e);
        sg = fmaf(ds[s], e, sg);
        sz += dz;
      }
      if (drow) stf<T>(drow + s, dz);
      if (hib) {
        const bf16 h = __float2bfloat16(dz);
        hib[(int64_t)a * Sp + s] = h;
        if (lob) lob[(int64_t)a * Sp + s] = __float2bfloat16(dz - __bfloat162float(h));
      }
    }
    sz = warp_sum(sz);
    sg = warp_sum(sg);
    if (lane == 0) {
      dqa[(int64_t)b * A + a] = sz;
      if (dqa_b) dqa_b[(int64_t)b * A + a] = __float2bfloat16(sz);
      gws_part[(int64_t)b * A + a] = sg;
    }
  }
}

template <typename T>
__global__ void iembed_bwd_pw_kernel(int64_t total, int M, int S, int Sp, const float* __restrict__ dI,
                                     const T* __restrict__ I, const float* __restrict__ da,
                                     const float* __restrict__ p, T* __restrict__ dY) {
  RAU_PDL_ENTRY();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int s = (int)(i % Sp);
    const int64_t bm = i / Sp;
    const int b = (int)(bm / M);
    float v = 0.0f;
    if (s < S) {
      const float y = ldf<T>(I + i);
      v = (dI[i] + da[bm] * p[(int64_t)b * S + s]) * (1.0f - y * y);
    }
    stf<T>(dY + i, v);
  }
}

// one warp per (image, channel) row of 196 cells: dY = (dI + da p) (1 - I^2)
__global__ void __launch_bounds__(256) iembed_bwd_rows_kernel(int64_t rows, int M, int S, int Sp, const float* __restrict__ dI,
                                                              const float* __restrict__ I, const float* __restrict__ da,
                                                              const float* __restrict__ p, float* __restrict__ dY,
                                                              bf16* __restrict__ hi, bf16* __restrict__ lo,
                                                              float* __restrict__ gbi) {
  RAU_PDL_ENTRY();
  const int lane = threadIdx.x & 31;
  const int64_t w0 = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5, nw = ((int64_t)gridDim.x * blockDim.x) >> 5;
  for (int64_t row = w0; row < rows; row += nw) {
    const int64_t b = row / M;
    const int m = (int)(row % M);
    const float d = da[row];
    const float* pr = p + b * S;
    float sum = 0.0f;
    for (int s = lane; s < Sp; s += 32) {
      float v = 0.0f;
      if (s < S) {
        const float y = I[row * Sp + s];
        v = (dI[row * Sp + s] + d * pr[s]) * (1.0f - y * y);
        sum += v;
      }
      if (dY) dY[row * Sp + s] = v;
      if (hi) {
        const bf16 h = __float2bfloat16(v);
        hi[row * Sp + s] = h;
        if (lo) lo[row * Sp + s] = __float2bfloat16(v - __bfloat162float(h));
      }
    }
    sum = warp_sum(sum);
    if (lane == 0 && gbi) atomicAdd(&gbi[m], sum);
  }
}

// ---------------------------------------------------------------- criteria
// one CTA per row: CrossEntropyCriterion forward + backward + argmax (F:505, F:535, F:585)
__global__ void __launch_bounds__(256) softmax_ce_kernel(int N, const float* __restrict__ score, const float* __restrict__ labels,
                                                         float loss_scale, float grad_scale, float* __restrict__ loss_sum,
                                                         float* __restrict__ dscore_f, bf16* __restrict__ dscore_b, int lddb,
                                                         float* __restrict__ answers, bf16* __restrict__ dscore_lo) {
  RAU_PDL_ENTRY();
  __shared__ float red[32];
  __shared__ int redi[32];
  const int b = blockIdx.x, tid = threadIdx.x;
  const float* row = score + (int64_t)b * N;
  float mx = -INFINITY;
  int arg = 0x7fffffff;
  for (int n = tid; n < N; n += blockDim.x) {
    const float v = row[n];
    if (v > mx) { mx = v; arg = n; }
  }
  // (max, lowest index) reduction
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
  }
  const int lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
  if (lane == 0) { red[w] = mx; redi[w] = arg; }
  __syncthreads();
  if (w == 0) {
    mx = lane < nw ? red[lane] : -INFINITY;
    arg = lane < nw ? redi[lane] : 0x7fffffff;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, mx, o);
      const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
      if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
    }
    if (lane == 0) { red[0] = mx; redi[0] = arg; }
  }
  __syncthreads();
  mx = red[0];
  arg = redi[0];
  __syncthreads();
  float se = 0.0f;
  for (int n = tid; n < N; n += blockDim.x) se += __expf(row[n] - mx);
  se = block_sum(se, red);
  const int y = (int)labels[b] - 1;
  if (tid == 0) {
    if (answers) answers[b] = (float)(arg + 1);
    if (loss_sum && y >= 0 && y < N) atomicAdd(loss_sum, loss_scale * (logf(se) - (row[y] - mx)));
  }
  if (dscore_f || dscore_b) {
    const float inv = 1.0f / se;
    for (int n = tid; n < (dscore_b ? lddb : N); n += blockDim.x) {
      float g = 0.0f;
      if (n < N) g = grad_scale * (__expf(row[n] - mx) * inv - (n == y ? 1.0f : 0.0f));
      if (dscore_f && n < N) dscore_f[(int64_t)b * N + n] = g;
      if (dscore_b) {
        const bf16 h = __float2bfloat16(g);
        dscore_b[(int64_t)b * lddb + n] = h;
        if (dscore_lo) dscore_lo[(int64_t)b * lddb + n] = __float2bfloat16(g - __bfloat162float(h));
      }
    }
  }
}

// Open-ended and multiple-choice answers of predict_result's nHop + 2 prediction tables (F:903-918): a warp per
// (table, sample) row.  oe = argmax_n pred[n]; mc = argmax_n pred[n] * mask[n], where mask[n] = 1 for the sample's candidate
// answers (ans_mc, 1-based ids, 0 = empty slot) and 0 elsewhere -- the reference MULTIPLIES by the mask (mc_pred:cmul), so a
// non-candidate scores 0, not -inf, and wins over all-negative candidates; kept as is.  Ties -> lowest index.  1-based.
__global__ void __launch_bounds__(256) answers_kernel(int rows, int B, int N, const float* __restrict__ pred,
                                                      const float* __restrict__ mc, int nmc, float* __restrict__ oe_out,
                                                      float* __restrict__ mc_out) {
  const int row = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (row >= rows) return;
  const int b = row % B;
  const float* p = pred + (int64_t)row * N;
  float cand[32];   // (nmc <= 32: the VQA multiple-choice task has 18 candidates)
  for (int k = 0; k < 32; ++k) cand[k] = (mc != nullptr && k < nmc) ? mc[(int64_t)b * nmc + k] : 0.0f;
  float mo = -INFINITY, mm = -INFINITY;
  int ao = 0x7fffffff, am = 0x7fffffff;
  for (int n = lane; n < N; n += 32) {
    const float v = p[n];
    if (v > mo) { mo = v; ao = n; }
    if (mc_out != nullptr) {
      bool in = false;
      for (int k = 0; k < 32; ++k) in = in || (cand[k] == (float)(n + 1));
      const float w = in ? v : v * 0.0f;   // (cmul by the 0/1 mask: NaN / inf propagate like the reference's product)
      if (w > mm) { mm = w; am = n; }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float v1 = __shfl_xor_sync(0xffffffffu, mo, o);
    const int a1 = __shfl_xor_sync(0xffffffffu, ao, o);
    if (v1 > mo || (v1 == mo && a1 < ao)) { mo = v1; ao = a1; }
    const float v2 = __shfl_xor_sync(0xffffffffu, mm, o);
    const int a2 = __shfl_xor_sync(0xffffffffu, am, o);
    if (v2 > mm || (v2 == mm && a2 < am)) { mm = v2; am = a2; }
  }
  if (lane == 0) {
    oe_out[row] = (float)(ao + 1);
    if (mc_out != nullptr) mc_out[row] = (float)(am + 1);
  }
}

// logging-only merged predictions (F:539-574) and predict_result's merge (F:699-721); one CTA per row.
__global__ void __launch_bounds__(256) merge_preds_kernel(int nHop, int B, int N, int S, const float* __restrict__ scores,
                                                          const float* __restrict__ do_pred, const float* __restrict__ attprob,
                                                          const float* __restrict__ labels, const float* __restrict__ answers_hop,
                                                          int force_last, float inv_bglobal, float* __restrict__ loss_uni_sel,
                                                          float* __restrict__ loss_do_pred, float* __restrict__ answers_uni_sel,
                                                          float* __restrict__ pred_uni, float* __restrict__ pred_sel,
                                                          float* __restrict__ att_uni, float* __restrict__ att_sel) {
  RAU_PDL_ENTRY();
  __shared__ float red[32];
  __shared__ int redi[32];
  __shared__ float cur[64];
  const int b = blockIdx.x, tid = threadIdx.x;
  if (tid == 0) {
    float did = 0.0f;
    for (int h = 0; h < nHop; ++h) {
      float dp = do_pred[(int64_t)h * B + b] > 0.5f ? 1.0f : 0.0f;
      if (force_last && h == nHop - 1) dp = 1.0f;                       // F:704
      cur[h] = fminf(fmaxf(dp - did, 0.0f), 1.0f);                      // F:522 / F:705
      did = fminf(did + dp, 1.0f);                                      // F:532 / F:716
      if (loss_do_pred && labels && answers_hop) {                      // BCE against is_correct (F:572)
        const float x = do_pred[(int64_t)h * B + b];
        const float t = answers_hop[(int64_t)h * B + b] == labels[b] ? 1.0f : 0.0f;
        const float l = -(t * logf(x + 1e-12f) + (1.0f - t) * logf(1.0f - x + 1e-12f));
        atomicAdd(&loss_do_pred[h], l * inv_bglobal);
      }
    }
  }
  __syncthreads();
  const float invh = 1.0f / (float)nHop;
  for (int which = 0; which < 2; ++which) {   // 0 = uni (mean over hops), 1 = select
    float mx = -INFINITY;
    int arg = 0x7fffffff;
    for (int n = tid; n < N; n += blockDim.x) {
      float v = 0.0f;
      for (int h = 0; h < nHop; ++h) {
        const float sc = scores[((int64_t)h * B + b) * N + n];
        v += which == 0 ? sc : sc * cur[h];
      }
      if (which == 0) v *= invh;
      float* dst = which == 0 ? pred_uni : pred_sel;
      if (dst) dst[(int64_t)b * N + n] = v;
      if (v > mx) { mx = v; arg = n; }
    }
    if (labels) {
      const int lane = tid & 31, w = tid >> 5, nw = blockDim.x >> 5;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        const float om = __shfl_xor_sync(0xffffffffu, mx, o);
        const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
        if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
      }
      __syncthreads();
      if (lane == 0) { red[w] = mx; redi[w] = arg; }
      __syncthreads();
      if (tid == 0) {
        for (int k = 1; k < nw; ++k)
          if (red[k] > mx || (red[k] == mx && redi[k] < arg)) { mx = red[k]; arg = redi[k]; }
        red[0] = mx; redi[0] = arg;
      }
      __syncthreads();
      mx = red[0];
      arg = redi[0];
      __syncthreads();
      float se = 0.0f, vy = 0.0f;
      const int y = (int)labels[b] - 1;
      for (int n = tid; n < N; n += blockDim.x) {
        float v = 0.0f;
        for (int h = 0; h < nHop; ++h) {
          const float sc = scores[((int64_t)h * B + b) * N + n];
          v += which == 0 ? sc : sc * cur[h];
        }
        if (which == 0) v *= invh;
        se += __expf(v - mx);
        if (n == y) vy = v;
      }
      se = block_sum(se, red);
      vy = block_sum(vy, red);
      if (tid == 0) {
        if (answers_uni_sel) answers_uni_sel[(int64_t)which * B + b] = (float)(arg + 1);
        if (loss_uni_sel && y >= 0 && y < N) atomicAdd(&loss_uni_sel[which], (logf(se) - (vy - mx)) * inv_bglobal);
      }
    }
    float* adst = which == 0 ? att_uni : att_sel;
    if (adst && attprob) {
      for (int s = tid; s < S; s += blockDim.x) {
        float v = 0.0f;
        for (int h = 0; h < nHop; ++h) {
          const float pv = attprob[((int64_t)h * B + b) * S + s];
          v += which == 0 ? pv : pv * cur[h];
        }
        adst[(int64_t)b * S + s] = which == 0 ? v * invh : v;
      }
    }
    __syncthreads();
  }
}
}  // namespace

template <typename T>
int k_attn_fwd(rau_ctx* ctx, int B, int M, int A, int S, int Sp, const T* E, const T* I, const float* ws, const float* mem,
               float* p, bf16* p_b, int ldpb, float* a, bf16* a_b) {
  if (S > ATT_T || Sp > ATT_T || ldpb > ATT_T) { rau_set_error("attention grid S=%d too large", S); return RAU_EINVAL; }
  RAU_LAUNCH_PDL(ctx->stream, (attn_fwd_kernel<T>), B, ATT_T, (Sp + A) * sizeof(float), M, A, S, Sp, E, I, ws, mem, p, p_b, ldpb, a, a_b);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
template int k_attn_fwd<float>(rau_ctx*, int, int, int, int, int, const float*, const float*, const float*, const float*, float*, bf16*, int, float*, bf16*);
template int k_attn_fwd<bf16>(rau_ctx*, int, int, int, int, int, const bf16*, const bf16*, const float*, const float*, float*, bf16*, int, float*, bf16*);

template <typename T>
int k_attn_bwd(rau_ctx* ctx, int B, int M, int A, int S, int Sp, const T* E, const T* I, const float* ws, const float* p,
               const float* dp_in, const float* da, float* ds, bf16* ds_b, int lddsb, T* dZ, float* dqa, bf16* dqa_b,
               float* gws_part, bf16* dZ_hi, bf16* dZ_lo) {
  if (S > ATT_T || Sp > ATT_T || lddsb > ATT_T) { rau_set_error("attention grid S=%d too large", S); return RAU_EINVAL; }
  RAU_LAUNCH_PDL(ctx->stream, (attn_bwd_kernel<T>), B, ATT_T, (Sp + M) * sizeof(float), M, A, S, Sp, E, I, ws, p, dp_in, da, ds, ds_b, lddsb,
                                                                          dZ, dqa, dqa_b, gws_part, dZ_hi, dZ_lo);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
template int k_attn_bwd<float>(rau_ctx*, int, int, int, int, int, const float*, const float*, const float*, const float*, const float*, const float*, float*, bf16*, int, float*, float*, bf16*, float*, bf16*, bf16*);
template int k_attn_bwd<bf16>(rau_ctx*, int, int, int, int, int, const bf16*, const bf16*, const float*, const float*, const float*, const float*, float*, bf16*, int, bf16*, float*, bf16*, float*, bf16*, bf16*);

template <typename T>
int k_iembed_bwd_pw(rau_ctx* ctx, int B, int M, int S, int Sp, const float* dI, const T* I, const float* da, const float* p, T* dY) {
  const int64_t total = (int64_t)B * M * Sp;
  int64_t blocks = (total + 1023) / 1024;
  if (blocks > 148 * 16) blocks = 148 * 16;
  RAU_LAUNCH_PDL(ctx->stream, (iembed_bwd_pw_kernel<T>), (int)blocks, 256, 0, total, M, S, Sp, dI, I, da, p, dY);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
template int k_iembed_bwd_pw<float>(rau_ctx*, int, int, int, int, const float*, const float*, const float*, const float*, float*);
template int k_iembed_bwd_pw<bf16>(rau_ctx*, int, int, int, int, const float*, const bf16*, const float*, const float*, bf16*);

int k_iembed_bwd_rows(rau_ctx* ctx, int B, int M, int S, int Sp, const float* dI, const float* I, const float* da,
                      const float* p, float* dY, bf16* dY_hi, bf16* dY_lo, float* gbi) {
  const int64_t rows = (int64_t)B * M;
  int64_t blocks = (rows + 7) / 8;
  if (blocks > 148 * 8) blocks = 148 * 8;
  RAU_LAUNCH_PDL(ctx->stream, (iembed_bwd_rows_kernel), (int)blocks, 256, 0, rows, M, S, Sp, dI, I, da, p, dY, dY_hi, dY_lo, gbi);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}

int k_softmax_ce(rau_ctx* ctx, int B, int N, const float* score, const float* labels, float loss_scale, float grad_scale,
                 float* loss_sum, float* dscore_f, bf16* dscore_b, int lddb, float* answers, bf16* dscore_lo) {
  RAU_LAUNCH_PDL(ctx->stream, (softmax_ce_kernel), B, 256, 0, N, score, labels, loss_scale, grad_scale, loss_sum, dscore_f, dscore_b, lddb, answers,
                 dscore_lo);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}

int k_answers(rau_ctx* ctx, int rows, int B, int N, const float* pred, const float* mc, int nmc, float* oe_out, float* mc_out) {
  if (nmc > 32) { rau_set_error("k_answers: %d multiple-choice candidates > 32", nmc); return RAU_EINVAL; }
  answers_kernel<<<(rows + 7) / 8, 256, 0, ctx->stream>>>(rows, B, N, pred, mc, nmc, oe_out, mc != nullptr ? mc_out : nullptr);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}

int k_merge_preds(rau_ctx* ctx, int nHop, int B, int N, int S, const float* scores, const float* do_pred, const float* attprob,
                  const float* labels, const float* answers_hop, int force_last, float inv_bglobal, float* loss_uni_sel,
                  float* loss_do_pred, float* answers_uni_sel, float* pred_uni, float* pred_sel, float* att_uni, float* att_sel) {
  if (nHop > 64) { rau_set_error("nHop=%d > 64", nHop); return RAU_EINVAL; }
  RAU_LAUNCH_PDL(ctx->stream, (merge_preds_kernel), B, 256, 0, nHop, B, N, S, scores, do_pred, attprob, labels, answers_hop, force_last,
                                                 inv_bglobal, loss_uni_sel, loss_do_pred, answers_uni_sel, pred_uni, pred_sel,
                                                 att_uni, att_sel);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
