// k_pointwise.cu -- HBM-bound elementwise / gather / reduction kernels of the RAU path.  Each replaces a
// chain of tiny cunn kernels of the reference: Dropout+Tanh+LookupTable, the ~12 pointwise
// nodes of an LSTM cell, Narrow/CAddTable glue, and the per-row host copy loops of F:472-478 / F:604-610.
#include "rau_kernels.cuh"

namespace {
constexpr int TPB = 256;
inline int grid_for(int64_t n, int per_thread = 1) {
  int64_t b = (n + (int64_t)TPB * per_thread - 1) / ((int64_t)TPB * per_thread);
  if (b < 1) b = 1;
  if (b > 148 * 32) b = 148 * 32;
  return (int)b;
}

// ---------------------------------------------------------------- masks
__global__ void mask_gen_kernel(uint32_t* __restrict__ bits, int64_t nwords, int64_t n, uint32_t thresh,
                                int keep_all, uint2 key, uint32_t stream_lo, uint32_t stream_hi,
                                const StepState* __restrict__ ss, int64_t hop_stride) {
  RAU_PDL_ENTRY();
  // blockIdx.y = hop: the masks of all hops in one launch (streams differ in their index field, words hop_stride apart)
  bits += (int64_t)blockIdx.y * hop_stride;
  stream_lo ^= blockIdx.y;
  if (ss) {   // graph replay: the step part of the stream id lives on the device
    const unsigned long long sid = (((unsigned long long)stream_hi << 32) | stream_lo) ^ (ss->step << 24);
    stream_lo = (uint32_t)sid;
    stream_hi = (uint32_t)(sid >> 32);
  }
  for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < nwords; w += (int64_t)gridDim.x * blockDim.x) {
    uint32_t word = 0;
#pragma unroll
    for (int q = 0; q < 8; ++q) {  // 8 Philox calls x 4 lanes = 32 keep bits
      uint4 r = philox4x32(make_uint4((uint32_t)(w * 8 + q), (uint32_t)((w * 8 + q) >> 32), stream_lo, stream_hi), key);
      word |= (uint32_t)(r.x < thresh) << (q * 4 + 0);
      word |= (uint32_t)(r.y < thresh) << (q * 4 + 1);
      word |= (uint32_t)(r.z < thresh) << (q * 4 + 2);
      word |= (uint32_t)(r.w < thresh) << (q * 4 + 3);
    }
    if (keep_all) word = 0xffffffffu;
    bits[w] = word;
  }
}

__global__ void mask_pack_kernel(uint32_t* __restrict__ bits, const uint8_t* __restrict__ bytes, int64_t nwords, int64_t n) {
  RAU_PDL_ENTRY();
  for (int64_t w = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; w < nwords; w += (int64_t)gridDim.x * blockDim.x) {
    uint32_t word = 0;
    for (int k = 0; k < 32; ++k) {
      int64_t i = w * 32 + k;
      if (i < n && bytes[i]) word |= 1u << k;
    }
    bits[w] = word;
  }
}

// packed keep bits -> 0/1 bytes (rau_draw_masks: the masks a step draws, exported for the parity tests)
__global__ void mask_unpack_kernel(const uint32_t* __restrict__ bits, int64_t n, uint8_t* __restrict__ bytes) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    bytes[i] = (uint8_t)((bits[i >> 5] >> (i & 31)) & 1u);
}

// The feature-dropout keep decisions of the rows pack kernels (k_rows_tc.cu xprep_rows_kernel / xprep_rows_hops_kernel)
// as 0/1 bytes in the [nHop, B, C, S] layout of rau_masks.x: element (b, c, s) of hop h takes 16 bits of one Philox draw
// shared by the channel pair (c & ~1, c | 1) x 4 consecutive grid cells -- word s & 3, low half for the even channel.
__global__ void xmask16_bytes_kernel(uint8_t* __restrict__ out, int B, int C, int S, int nHop, uint32_t thresh, uint2 key,
                                     uint32_t stream_lo, uint32_t stream_hi, const StepState* __restrict__ ss) {
  if (ss) {
    const unsigned long long sid = (((unsigned long long)stream_hi << 32) | stream_lo) ^ (ss->step << 24);
    stream_lo = (uint32_t)sid;
    stream_hi = (uint32_t)(sid >> 32);
  }
  const int64_t per_hop = (int64_t)B * C * S, total = per_hop * nHop;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int h = (int)(i / per_hop);
    const int64_t e = i - (int64_t)h * per_hop;          // (b*C + c)*S + s
    const int s = (int)(e % S);
    const int64_t bc = e / S;
    const int c = (int)(bc % C);
    const uint64_t ctr = (uint64_t)(((bc - (c & 1)) * S) + (s & ~3)) >> 2;
    const bool shared = rau_xmask_shared(thresh, nHop);
    const uint4 r = philox4x32(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32),
                                          shared ? stream_lo ^ RAU_XMASK_SHARED_TAG : stream_lo ^ (uint32_t)h, stream_hi), key);
    const uint32_t w = (s & 3) == 0 ? r.x : (s & 3) == 1 ? r.y : (s & 3) == 2 ? r.z : r.w;
    if (shared) out[i] = (uint8_t)(((w >> ((c & 1) * 16 + h)) & 1u) ? 0 : 1);
    else out[i] = (uint8_t)((((w >> ((c & 1) * 16)) & 0xffffu) < thresh) ? 1 : 0);
  }
}

// Evicts L2 by READING a buffer larger than it (a memset would leave L2 full of dirty lines whose write-back the next
// kernel pays for); the sum is stored only under a condition that never holds, so the loads are not optimised away.
__global__ void l2_evict_kernel(const float4* __restrict__ buf, int64_t n4, float* __restrict__ sink) {
  float acc = 0.0f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 v = __ldcg(buf + i);
    acc += v.x + v.y + v.z + v.w;
  }
  if (acc == 1.2345e38f) *sink = acc;
}

// ---------------------------------------------------------------- embedding
__global__ void embed_fwd_kernel(const float* __restrict__ ids, int n, int D, int V, const float* __restrict__ E,
                                 const uint32_t* __restrict__ bits, float scale, float* __restrict__ out_f,
                                 bf16* __restrict__ out_b, int ldb, bf16* __restrict__ out_lo) {
  RAU_PDL_ENTRY();
  const int64_t total = (int64_t)n * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / D), d = (int)(i % D);
    int id = (int)ids[r] - 1;
    id = min(max(id, 0), V - 1);
    const float v = tanhf(E[(int64_t)id * D + d] * keep_scale(bits, i, scale));
    if (out_f) out_f[i] = v;
    if (out_b) {   // packed twin for the tcgen05 products: hi [, lo] (the layer-1 input projection reads it as stored)
      const bf16 h = __float2bfloat16(v);
      out_b[(int64_t)r * ldb + d] = h;
      if (out_lo) out_lo[(int64_t)r * ldb + d] = __float2bfloat16(v - __bfloat162float(h));
    }
  }
}

__global__ void embed_bwd_kernel(const float* __restrict__ ids, int n, int D, int V, const float* __restrict__ out,
                                 const uint32_t* __restrict__ bits, float scale, const float* __restrict__ dout, int lddout,
                                 float* __restrict__ gE) {
  RAU_PDL_ENTRY();
  const int64_t total = (int64_t)n * D;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = (int)(i / D), d = (int)(i % D);
    int id = (int)ids[r] - 1;
    id = min(max(id, 0), V - 1);
    const float e = out[i];
    const float g = dout[(int64_t)r * lddout + d] * (1.0f - e * e) * keep_scale(bits, i, scale);
    if (g != 0.0f) atomicAdd(&gE[(int64_t)id * D + d], g);
  }
}

// ---------------------------------------------------------------- LSTM pointwise
// chunk index of gates (i,f,o,g) inside the 4H pre-activation vector
__device__ __forceinline__ void gate_chunks(int order, int& ci, int& cf, int& co, int& cg) {
  if (order == RAU_GATES_IFOG) { ci = 0; cf = 1; co = 2; cg = 3; }
  else { ci = 0; cg = 1; cf = 2; co = 3; }
}

__global__ void lstm_fwd_kernel(int B, int H, int order, const float* __restrict__ G, int ldg,
                                const float* __restrict__ c_prev, int ldcp, float* __restrict__ c, int ldc,
                                float* __restrict__ h, int ldh, bf16* __restrict__ h_b, int ldhb, float* __restrict__ saved) {
  RAU_PDL_ENTRY();
  int ci, cf, co, cg;
  gate_chunks(order, ci, cf, co, cg);
  const int64_t total = (int64_t)B * H, plane = total;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / H), j = (int)(idx % H);
    const float* g = G + (int64_t)b * ldg;
    const float i_ = sigmoidf_(g[ci * H + j]);
    const float f_ = sigmoidf_(g[cf * H + j]);
    const float o_ = sigmoidf_(g[co * H + j]);
    const float g_ = tanhf(g[cg * H + j]);
    const float cp = c_prev ? c_prev[(int64_t)b * ldcp + j] : 0.0f;
    const float cn = f_ * cp + i_ * g_;
    const float tc = tanhf(cn);
    const float hn = o_ * tc;
    c[(int64_t)b * ldc + j] = cn;
    h[(int64_t)b * ldh + j] = hn;
    if (h_b) h_b[(int64_t)b * ldhb + j] = __float2bfloat16(hn);
    if (saved) {
      saved[idx] = i_; saved[plane + idx] = f_; saved[2 * plane + idx] = o_;
      saved[3 * plane + idx] = g_; saved[4 * plane + idx] = tc;
    }
  }
}

__global__ void lstm_bwd_kernel(int B, int H, int order, const float* __restrict__ dc_out, int lddc,
                                const float* __restrict__ dh_out, int lddh, const float* __restrict__ dh_extra, int ldhe,
                                const float* __restrict__ lengths, int t, const float* __restrict__ dq_c,
                                const float* __restrict__ dq_h, int lddq,
                                const float* __restrict__ c_prev, int ldcp, const float* __restrict__ saved,
                                float* __restrict__ dG, bf16* __restrict__ dG_b, float* __restrict__ dc_prev, int lddcp,
                                bf16* __restrict__ dG_lo, float* __restrict__ zero_out, const uint32_t* __restrict__ extra_bits,
                                int64_t extra_bit0, float extra_scale) {
  RAU_PDL_ENTRY();
  // zero_out ([B, H] contiguous): the buffer the split-K dgrad that follows reduces into (saves a memset node per step)
  if (zero_out)
    for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < (int64_t)B * H; idx += (int64_t)gridDim.x * blockDim.x)
      zero_out[idx] = 0.0f;
  int ci, cf, co, cg;
  gate_chunks(order, ci, cf, co, cg);
  const int64_t total = (int64_t)B * H, plane = total;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(idx / H), j = (int)(idx % H);
    float dc_o = dc_out ? dc_out[(int64_t)b * lddc + j] : 0.0f;
    float dh_o = dh_out ? dh_out[(int64_t)b * lddh + j] : 0.0f;
    if (lengths != nullptr && (int)lengths[b] == t) {   // drnn_out[k] = d_feats[1][k]  (F:604-610)
      dc_o = dq_c[(int64_t)b * lddq + j];
      dh_o = dq_h[(int64_t)b * lddq + j];
    }
    // (extra_bits: dh_extra is the gradient w.r.t. a dropped-out copy of h -- D:39 -- whose keep mask is applied here)
    if (dh_extra) dh_o += dh_extra[(int64_t)b * ldhe + j] * keep_scale(extra_bits, extra_bit0 + idx, extra_scale);
    const float i_ = saved[idx], f_ = saved[plane + idx], o_ = saved[2 * plane + idx];
    const float g_ = saved[3 * plane + idx], tc = saved[4 * plane + idx];
    const float cp = c_prev ? c_prev[(int64_t)b * ldcp + j] : 0.0f;
    const float d_o = dh_o * tc;
    const float dc = dc_o + dh_o * o_ * (1.0f - tc * tc);
    const float d_f = dc * cp, d_i = dc * g_, d_g = dc * i_;
    if (dc_prev) dc_prev[(int64_t)b * lddcp + j] = dc * f_;
    const float gi = d_i * i_ * (1.0f - i_), gf = d_f * f_ * (1.0f - f_);
    const float go = d_o * o_ * (1.0f - o_), gg = d_g * (1.0f - g_ * g_);
    const int64_t row = (int64_t)b * 4 * H;
    if (dG) {
      dG[row + ci * H + j] = gi; dG[row + cf * H + j] = gf; dG[row + co * H + j] = go; dG[row + cg * H + j] = gg;
    }
    if (dG_b) {
      dG_b[row + ci * H + j] = __float2bfloat16(gi); dG_b[row + cf * H + j] = __float2bfloat16(gf);
      dG_b[row + co * H + j] = __float2bfloat16(go); dG_b[row + cg * H + j] = __float2bfloat16(gg);
      if (dG_lo) {   // bf16x3 operand: lo = bf16(x - hi)
        dG_lo[row + ci * H + j] = __float2bfloat16(gi - __bfloat162float(__float2bfloat16(gi)));
        dG_lo[row + cf * H + j] = __float2bfloat16(gf - __bfloat162float(__float2bfloat16(gf)));
        dG_lo[row + co * H + j] = __float2bfloat16(go - __bfloat162float(__float2bfloat16(go)));
        dG_lo[row + cg * H + j] = __float2bfloat16(gg - __bfloat162float(__float2bfloat16(gg)));
      }
    }
  }
}

// The same cell backward, four hidden units per thread: 16-byte loads and stores, 8-byte packed (hi / lo) stores, 32-bit
// index arithmetic.  The scalar form issues ~25 memory instructions and two 64-bit divisions per element; this launch sits
// 2 T times on the encoder backward's serial lanes (and once per hop), so its latency is what matters, not its bytes.
__device__ __forceinline__ uint2 pack4_bf16(float a, float b, float c, float d) {
  const __nv_bfloat162 p0 = __floats2bfloat162_rn(a, b), p1 = __floats2bfloat162_rn(c, d);
  return make_uint2(*reinterpret_cast<const uint32_t*>(&p0), *reinterpret_cast<const uint32_t*>(&p1));
}
__global__ void lstm_bwd_vec4_kernel(int B, int H, int order, const float* __restrict__ dc_out, int lddc,
                                     const float* __restrict__ dh_out, int lddh, const float* __restrict__ dh_extra, int ldhe,
                                     const float* __restrict__ lengths, int t, const float* __restrict__ dq_c,
                                     const float* __restrict__ dq_h, int lddq,
                                     const float* __restrict__ c_prev, int ldcp, const float* __restrict__ saved,
                                     float* __restrict__ dG, bf16* __restrict__ dG_b, float* __restrict__ dc_prev, int lddcp,
                                     bf16* __restrict__ dG_lo, float* __restrict__ zero_out, const uint32_t* __restrict__ extra_bits,
                                     int64_t extra_bit0, float extra_scale) {
  RAU_PDL_ENTRY();
  const int total4 = (B * H) >> 2, h4 = H >> 2;
  const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
  int ci, cf, co, cg;
  gate_chunks(order, ci, cf, co, cg);
  const size_t plane = (size_t)B * H;
  for (int i4 = blockIdx.x * blockDim.x + threadIdx.x; i4 < total4; i4 += gridDim.x * blockDim.x) {
    if (zero_out) reinterpret_cast<float4*>(zero_out)[i4] = z4;
    const int b = i4 / h4, j = (i4 - b * h4) << 2;
    const int idx = i4 << 2;
    float4 dc_o = dc_out ? *reinterpret_cast<const float4*>(dc_out + (size_t)b * lddc + j) : z4;
    float4 dh_o = dh_out ? *reinterpret_cast<const float4*>(dh_out + (size_t)b * lddh + j) : z4;
    if (lengths != nullptr && (int)lengths[b] == t) {   // drnn_out[k] = d_feats[1][k]  (F:604-610)
      dc_o = *reinterpret_cast<const float4*>(dq_c + (size_t)b * lddq + j);
      dh_o = *reinterpret_cast<const float4*>(dq_h + (size_t)b * lddq + j);
    }
    if (dh_extra) {   // gradient w.r.t. a dropped-out copy of h (D:39): its keep mask is applied here
      const float4 e = *reinterpret_cast<const float4*>(dh_extra + (size_t)b * ldhe + j);
      float k0 = 1.0f, k1 = 1.0f, k2 = 1.0f, k3 = 1.0f;
      if (extra_bits) {   // (extra_bit0 and idx are multiples of 4: the four keep bits sit in one word)
        const int64_t bit = extra_bit0 + idx;
        const uint32_t w = extra_bits[bit >> 5] >> (bit & 31);
        k0 = (w & 1u) ? extra_scale : 0.0f; k1 = (w & 2u) ? extra_scale : 0.0f;
        k2 = (w & 4u) ? extra_scale : 0.0f; k3 = (w & 8u) ? extra_scale : 0.0f;
      }
      dh_o.x = fmaf(e.x, k0, dh_o.x); dh_o.y = fmaf(e.y, k1, dh_o.y); dh_o.z = fmaf(e.z, k2, dh_o.z); dh_o.w = fmaf(e.w, k3, dh_o.w);
    }
    const float4 i_ = reinterpret_cast<const float4*>(saved)[i4], f_ = reinterpret_cast<const float4*>(saved + plane)[i4];
    const float4 o_ = reinterpret_cast<const float4*>(saved + 2 * plane)[i4], g_ = reinterpret_cast<const float4*>(saved + 3 * plane)[i4];
    const float4 tc = reinterpret_cast<const float4*>(saved + 4 * plane)[i4];
    const float4 cp = c_prev ? *reinterpret_cast<const float4*>(c_prev + (size_t)b * ldcp + j) : z4;
    float gi[4], gf[4], go[4], gg[4], dcp[4];
#define RAU_CELL_BWD(k, C)                                                        \
    {                                                                             \
      const float d_o = dh_o.C * tc.C;                                            \
      const float dc = dc_o.C + dh_o.C * o_.C * (1.0f - tc.C * tc.C);             \
      const float d_f = dc * cp.C, d_i = dc * g_.C, d_g = dc * i_.C;              \
      dcp[k] = dc * f_.C;                                                         \
      gi[k] = d_i * i_.C * (1.0f - i_.C); gf[k] = d_f * f_.C * (1.0f - f_.C);     \
      go[k] = d_o * o_.C * (1.0f - o_.C); gg[k] = d_g * (1.0f - g_.C * g_.C);     \
    }
    RAU_CELL_BWD(0, x) RAU_CELL_BWD(1, y) RAU_CELL_BWD(2, z) RAU_CELL_BWD(3, w)
#undef RAU_CELL_BWD
    if (dc_prev) *reinterpret_cast<float4*>(dc_prev + (size_t)b * lddcp + j) = make_float4(dcp[0], dcp[1], dcp[2], dcp[3]);
    const size_t row = (size_t)b * 4 * H + j;
    auto put = [&](int chunk, float a, float b2, float c2, float d) {
      const size_t o = row + (size_t)chunk * H;
      if (dG) *reinterpret_cast<float4*>(dG + o) = make_float4(a, b2, c2, d);
      if (dG_b) {
        const uint2 hi = pack4_bf16(a, b2, c2, d);
        *reinterpret_cast<uint2*>(dG_b + o) = hi;
        if (dG_lo)   // bf16x3 operand: lo = bf16(x - hi)
          *reinterpret_cast<uint2*>(dG_lo + o) = pack4_bf16(a - __uint_as_float(hi.x << 16), b2 - __uint_as_float(hi.x & 0xffff0000u),
                                                           c2 - __uint_as_float(hi.y << 16), d - __uint_as_float(hi.y & 0xffff0000u));
      }
    };
    put(ci, gi[0], gi[1], gi[2], gi[3]);
    put(cf, gf[0], gf[1], gf[2], gf[3]);
    put(co, go[0], go[1], go[2], go[3]);
    put(cg, gg[0], gg[1], gg[2], gg[3]);
  }
}

// ---------------------------------------------------------------- elementwise
__global__ void dropout_kernel(const float* __restrict__ x, int64_t rows, int cols, int ldx,
                               const uint32_t* __restrict__ bits, float scale,
                               float* __restrict__ y_f, int ldyf, bf16* __restrict__ y_b, int ldyb, int cols_pad,
                               bf16* __restrict__ y_lo) {
  RAU_PDL_ENTRY();
  const int64_t total = rows * cols_pad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols_pad;
    const int c = (int)(i % cols_pad);
    float v = 0.0f;
    if (c < cols) v = x[r * ldx + c] * keep_scale(bits, r * cols + c, scale);
    if (y_f) y_f[r * ldyf + c] = v;
    if (y_b) {
      const bf16 h = __float2bfloat16(v);
      y_b[r * ldyb + c] = h;
      if (y_lo) y_lo[r * ldyb + c] = __float2bfloat16(v - __bfloat162float(h));   // bf16x3 operand: (hi, lo)
    }
  }
}

// nn.Dropout fused with the tcgen05 operand packing: y = x * keep * scale written only as bf16 (hi, lo) rows of
// pitch cols_pad (zero padded).  Four columns per thread: 128-bit loads, 64-bit stores.
__global__ void dropout_pack_kernel(const float* __restrict__ x, int64_t rows, int cols, const uint32_t* __restrict__ bits,
                                    float scale, bf16* __restrict__ hi, bf16* __restrict__ lo, int cols_pad, int64_t ldx) {
  RAU_PDL_ENTRY();
  const int q = cols_pad >> 2;
  const int64_t total = rows * q;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / q;
    const int c = (int)(i % q) * 4;
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (c + 3 < cols) {
      const float4 t = *reinterpret_cast<const float4*>(x + r * ldx + c);
      v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w;
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (c + k < cols) v[k] = x[r * ldx + c + k];
    }
    __align__(8) bf16 h[4], l[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float y = (c + k < cols) ? v[k] * keep_scale(bits, r * cols + c + k, scale) : 0.0f;
      h[k] = __float2bfloat16(y);
      l[k] = __float2bfloat16(y - __bfloat162float(h[k]));
    }
    *reinterpret_cast<uint2*>(hi + i * 4) = *reinterpret_cast<const uint2*>(h);
    if (lo) *reinterpret_cast<uint2*>(lo + i * 4) = *reinterpret_cast<const uint2*>(l);
  }
}

// The same input dropped out once per hop with that hop's keep bits (hop h's bits start bits_stride words after hop
// h-1's): y[h][i] = x[i] * keep_h(i) * scale, fp32 + packed (hi, lo).  One launch for the q_embed inputs of all hops.
__global__ void dropout_hops_kernel(const float* __restrict__ x, int64_t n, int nHop, const uint32_t* __restrict__ bits,
                                    int64_t bits_stride, float scale, float* __restrict__ y, bf16* __restrict__ y_hi,
                                    bf16* __restrict__ y_lo) {
  RAU_PDL_ENTRY();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float xv = x[i];
    for (int h = 0; h < nHop; ++h) {
      const float v = xv * keep_scale(bits ? bits + h * bits_stride : nullptr, i, scale);
      y[h * n + i] = v;
      if (y_hi) {
        const bf16 hb = __float2bfloat16(v);
        y_hi[h * n + i] = hb;
        if (y_lo) y_lo[h * n + i] = __float2bfloat16(v - __bfloat162float(hb));
      }
    }
  }
}
// In-place dropout backward over a [nHop][n] stack with per-hop keep bits, fp32 + packed twin
__global__ void dropout_bwd_hops_kernel(float* __restrict__ y, int64_t n, int nHop, const uint32_t* __restrict__ bits,
                                        int64_t bits_stride, float scale, bf16* __restrict__ y_hi, bf16* __restrict__ y_lo) {
  RAU_PDL_ENTRY();
  const int64_t total = n * nHop;
  for (int64_t idx = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int h = (int)(idx / n);
    const int64_t i = idx - h * n;
    const float v = y[idx] * keep_scale(bits ? bits + h * bits_stride : nullptr, i, scale);
    y[idx] = v;
    if (y_hi) {
      const bf16 hb = __float2bfloat16(v);
      y_hi[idx] = hb;
      if (y_lo) y_lo[idx] = __float2bfloat16(v - __bfloat162float(hb));
    }
  }
}
// out[i] = sum_h dx[h][i] * keep_h(i) * scale: the gradient of the shared input of the per-hop dropouts
__global__ void dropout_bwd_sum_hops_kernel(const float* __restrict__ dx, int64_t n, int nHop, const uint32_t* __restrict__ bits,
                                            int64_t bits_stride, float scale, float* __restrict__ out) {
  RAU_PDL_ENTRY();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.0f;
    for (int h = 0; h < nHop; ++h) acc += dx[h * n + i] * keep_scale(bits ? bits + h * bits_stride : nullptr, i, scale);
    out[i] = acc;
  }
}

__global__ void dropout_bwd_acc_kernel(const float* __restrict__ dx, int64_t n, const uint32_t* __restrict__ bits,
                                       float scale, float* __restrict__ y, int accumulate, bf16* __restrict__ y_hi,
                                       bf16* __restrict__ y_lo) {
  RAU_PDL_ENTRY();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float v = dx[i] * keep_scale(bits, i, scale);
    if (accumulate) v += y[i];
    y[i] = v;
    if (y_hi) {
      const bf16 h = __float2bfloat16(v);
      y_hi[i] = h;
      if (y_lo) y_lo[i] = __float2bfloat16(v - __bfloat162float(h));
    }
  }
}

__global__ void tanh_bwd_kernel(const float* __restrict__ dy, const float* __restrict__ y, int64_t n,
                                float* __restrict__ dx_f, bf16* __restrict__ dx_b, bf16* __restrict__ dx_lo) {
  RAU_PDL_ENTRY();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float v = dy[i] * (1.0f - y[i] * y[i]);
    if (dx_f) dx_f[i] = v;
    if (dx_b) {
      const bf16 h = __float2bfloat16(v);
      dx_b[i] = h;
      if (dx_lo) dx_lo[i] = __float2bfloat16(v - __bfloat162float(h));
    }
  }
}

__global__ void add_kernel(const float* __restrict__ a, const float* __restrict__ b, int64_t n, float* __restrict__ y) {
  RAU_PDL_ENTRY();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = a[i] + b[i];
}
__global__ void axpy_kernel(float alpha, const float* __restrict__ x, int64_t n, float* __restrict__ y) {
  RAU_PDL_ENTRY();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] += alpha * x[i];
}
__global__ void fill_kernel(float* __restrict__ x, int64_t n, float v) {
  RAU_PDL_ENTRY();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) x[i] = v;
}
__global__ void to_bf16_kernel(const float* __restrict__ x, int64_t rows, int cols, int ldx, bf16* __restrict__ y, int ldy, int cols_pad) {
  RAU_PDL_ENTRY();
  const int64_t total = rows * cols_pad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int64_t r = i / cols_pad;
    const int c = (int)(i % cols_pad);
    y[r * ldy + c] = __float2bfloat16(c < cols ? x[r * ldx + c] : 0.0f);
  }
}

// y[b] = sigmoid(x[b,:] . w + bias)   -- the do_pred head, F:281 (Linear(M,1) -> Sigmoid -> Sum(2))
__global__ void rowdot_sigmoid_kernel(const float* __restrict__ x, int B, int K, const float* __restrict__ w,
                                      const float* __restrict__ bias, float* __restrict__ y) {
  RAU_PDL_ENTRY();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= B) return;
  float s = 0.0f;
  for (int k = lane; k < K; k += 32) s += x[(int64_t)warp * K + k] * w[k];
  s = warp_sum(s);
  if (lane == 0) y[warp] = 1.0f / (1.0f + expf(-(s + bias[0])));
}

__global__ void dopred_bwd_kernel(const float* __restrict__ ddo, const float* __restrict__ dop, const float* __restrict__ m,
                                  const float* __restrict__ wd, int B, int K, float* __restrict__ dm_acc,
                                  float* __restrict__ gwd, float* __restrict__ gbd) {
  RAU_PDL_ENTRY();
  // one block per feature chunk; loops over the batch (tiny: only used when a caller passes a do_pred gradient)
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  float gw = 0.0f, gb = 0.0f;
  for (int b = 0; b < B; ++b) {
    const float dl = ddo[b] * dop[b] * (1.0f - dop[b]);
    if (k < K) {
      dm_acc[(int64_t)b * K + k] += dl * wd[k];
      gw += dl * m[(int64_t)b * K + k];
    }
    gb += dl;
  }
  if (k < K) gwd[k] += gw;
  if (k == 0) gbd[0] += gb;
}

// ---------------------------------------------------------------- reductions
// out[c] (+)= sum_r x[r, c]; block = 32 columns x 8 row-lanes; gridDim.y row slices add atomically (accumulate mode only);
// out2 (optional) receives the same sums: the two biases of an LSTM layer (i2h, h2h) share one gradient
__global__ void colsum_kernel(const float* __restrict__ x, int64_t rows, int cols, int ld, float* __restrict__ out, int accumulate,
                              float* __restrict__ out2) {
  RAU_PDL_ENTRY();
  __shared__ float red[8][33];
  const int c = blockIdx.x * 32 + threadIdx.x;
  const int64_t per = (rows + gridDim.y - 1) / gridDim.y;
  const int64_t r_lo = blockIdx.y * per, r_hi = r_lo + per < rows ? r_lo + per : rows;
  float s = 0.0f;
  if (c < cols)
    for (int64_t r = r_lo + threadIdx.y; r < r_hi; r += 8) s += x[r * ld + c];
  red[threadIdx.y][threadIdx.x] = s;
  __syncthreads();
  if (threadIdx.y == 0 && c < cols) {
    float t = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    if (gridDim.y > 1) {
      atomicAdd(out + c, t);
      if (out2) atomicAdd(out2 + c, t);
    } else {
      out[c] = accumulate ? out[c] + t : t;
      if (out2) out2[c] = accumulate ? out2[c] + t : t;
    }
  }
}

template <typename T> __device__ __forceinline__ float ldf(const T* p);
template <> __device__ __forceinline__ float ldf<float>(const float* p) { return *p; }
template <> __device__ __forceinline__ float ldf<bf16>(const bf16* p) { return __bfloat162float(*p); }

template <typename T>
__global__ void rowsum_bms_kernel(const T* __restrict__ x, int B, int M, int S, int Sp, float* __restrict__ out) {
  RAU_PDL_ENTRY();
  __shared__ float red[32];
  const int m = blockIdx.x;
  float s = 0.0f;
  for (int b = blockIdx.y; b < B; b += gridDim.y) {
    const T* row = x + ((int64_t)b * M + m) * Sp;
    for (int k = threadIdx.x; k < S; k += blockDim.x) s += ldf<T>(row + k);
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(&out[m], s);
}

__global__ void sum_all_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out, int accumulate) {
  RAU_PDL_ENTRY();
  __shared__ float red[32];
  float s = 0.0f;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) s += x[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[0] = accumulate ? out[0] + s : s;
}

// rnn_out[b] = state_{t = len_b}[b]   (F:472-478 host loop, fused)
__global__ void select_state_kernel(const float* __restrict__ S_all, int T, int B, int Q,
                                    const float* __restrict__ lengths, float* __restrict__ out) {
  RAU_PDL_ENTRY();
  const int64_t total = (int64_t)B * Q;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / Q);
    int t = (int)lengths[b];
    float v = 0.0f;                       // rnn_out:zero() when no step matches (F:464)
    if (t >= 1 && t <= T) v = S_all[((int64_t)t * B + b) * Q + (i % Q)];
    out[i] = v;
  }
}
}  // namespace

// ================================================================== host wrappers
int k_mask_gen(rau_ctx* ctx, uint32_t* bits, int64_t n, float p, uint64_t seed, uint64_t stream_id, int nHop,
               int64_t hop_stride) {
  const int64_t nw = (n + 31) / 32;
  double keep = 1.0 - (double)p;
  uint32_t thresh = keep >= 1.0 ? 0xffffffffu : (uint32_t)(keep * 4294967296.0);
  RAU_LAUNCH_PDL(ctx->stream, (mask_gen_kernel), dim3(grid_for(nw), nHop), TPB, 0, bits, nw, n, thresh, p <= 0.0f ? 1 : 0,
      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)), (uint32_t)stream_id, (uint32_t)(stream_id >> 32), ctx->ss_active,
      hop_stride);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
int k_mask_unpack(rau_ctx* ctx, const uint32_t* bits, int64_t n, uint8_t* bytes) {
  mask_unpack_kernel<<<grid_for(n), TPB, 0, ctx->stream>>>(bits, n, bytes);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
int k_xmask16_bytes(rau_ctx* ctx, uint8_t* out, int B, int C, int S, int nHop, float p_drop, uint64_t stream_id) {
  const double keep = 1.0 - (double)p_drop;
  const uint32_t thresh = keep >= 1.0 ? 65536u : (uint32_t)(keep * 65536.0 + 0.5);   // as k_xprep_rows_hops
  xmask16_bytes_kernel<<<grid_for((int64_t)B * C * S * nHop), TPB, 0, ctx->stream>>>(
      out, B, C, S, nHop, thresh, make_uint2((uint32_t)ctx->seed, (uint32_t)(ctx->seed >> 32)), (uint32_t)stream_id,
      (uint32_t)(stream_id >> 32), ctx->ss_active);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
int k_l2_evict(rau_ctx* ctx, const void* buf, size_t bytes, float* sink) {
  l2_evict_kernel<<<148 * 8, 256, 0, ctx->stream>>>((const float4*)buf, (int64_t)(bytes / 16), sink);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
int k_mask_pack(rau_ctx* ctx, uint32_t* bits, const uint8_t* bytes, int64_t n) {
  const int64_t nw = (n + 31) / 32;
  RAU_LAUNCH_PDL(ctx->stream, (mask_pack_kernel), grid_for(nw), TPB, 0, bits, bytes, nw, n);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
int k_embed_fwd(rau_ctx* ctx, const float* ids, int n, int D, int V, const float* E, const uint32_t* bits, float scale,
                float* out_f, bf16* out_b, int ldb, bf16* out_lo) {
  RAU_LAUNCH_PDL(ctx->stream, (embed_fwd_kernel), grid_for((int64_t)n * D), TPB, 0, ids, n, D, V, E, bits, scale, out_f, out_b, ldb, out_lo);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
int k_embed_bwd(rau_ctx* ctx, const float* ids, int n, int D, int V, const float* out, const uint32_t* bits, float scale,
                const float* dout, int lddout, float* gE) {
  RAU_LAUNCH_PDL(ctx->stream, (embed_bwd_kernel), grid_for((int64_t)n * D), TPB, 0, ids, n, D, V, out, bits, scale, dout, lddout, gE);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
int k_lstm_fwd(rau_ctx* ctx, int B, int H, int order, const float* G, int ldg, const float* c_prev, int ldcp,
               float* c, int ldc, float* h, int ldh, bf16* h_b, int ldhb, float* saved) {
  RAU_LAUNCH_PDL(ctx->stream, (lstm_fwd_kernel), grid_for((int64_t)B * H), TPB, 0, B, H, order, G, ldg, c_prev, ldcp, c, ldc, h, ldh, h_b, ldhb, saved);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
int k_lstm_bwd(rau_ctx* ctx, int B, int H, int order, const float* dc_out, int lddc, const float* dh_out, int lddh,
               const float* dh_extra, int ldhe, const float* lengths, int t, const float* dq_c, const float* dq_h, int lddq,
               const float* c_prev, int ldcp, const float* saved, float* dG, bf16* dG_b, float* dc_prev, int lddcp, bf16* dG_lo,
               float* zero_out, const uint32_t* extra_bits, int64_t extra_bit0, float extra_scale) {
  auto al16 = [](const void* q) { return ((uintptr_t)q & 15) == 0; };
  const bool vec4 = H % 4 == 0 && (int64_t)B * H < (1ll << 30) && lddc % 4 == 0 && lddh % 4 == 0 && ldhe % 4 == 0 && lddq % 4 == 0 &&
                    ldcp % 4 == 0 && lddcp % 4 == 0 && extra_bit0 % 4 == 0 && al16(dc_out) && al16(dh_out) && al16(dh_extra) &&
                    al16(dq_c) && al16(dq_h) && al16(c_prev) && al16(saved) && al16(dG) && al16(dG_b) && al16(dc_prev) && al16(dG_lo) &&
                    al16(zero_out);
  if (vec4) {
    RAU_LAUNCH_PDL(ctx->stream, (lstm_bwd_vec4_kernel), grid_for((int64_t)B * H / 4), TPB, 0, B, H, order, dc_out, lddc, dh_out, lddh,
        dh_extra, ldhe, lengths, t, dq_c, dq_h, lddq, c_prev, ldcp, saved, dG, dG_b, dc_prev, lddcp, dG_lo, zero_out, extra_bits,
        extra_bit0, extra_scale);
    RAU_LAUNCH_CHECK(ctx);
    return RAU_OK;
  }
  RAU_LAUNCH_PDL(ctx->stream, (lstm_bwd_kernel), grid_for((int64_t)B * H), TPB, 0, B, H, order, dc_out, lddc, dh_out, lddh, dh_extra, ldhe,
      lengths, t, dq_c, dq_h, lddq, c_prev, ldcp, saved, dG, dG_b, dc_prev, lddcp, dG_lo, zero_out, extra_bits, extra_bit0,
      extra_scale);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
int k_dropout(rau_ctx* ctx, const float* x, int64_t rows, int cols, int ldx, const uint32_t* bits, float scale,
              float* y_f, int ldyf, bf16* y_b, int ldyb, int cols_pad, bf16* y_lo) {
  RAU_LAUNCH_PDL(ctx->stream, (dropout_kernel), grid_for(rows * cols_pad, 4), TPB, 0, x, rows, cols, ldx, bits, scale, y_f, ldyf, y_b, ldyb, cols_pad, y_lo);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
int k_dropout_pack(rau_ctx* ctx, const float* x, int64_t rows, int cols, const uint32_t* bits, float scale, bf16* hi, bf16* lo,
                   int cols_pad, int64_t ldx) {
  if (ldx <= 0) ldx = cols;
  if (cols % 4 != 0 || cols_pad % 4 != 0 || ldx % 4 != 0 || ((uintptr_t)x & 15) != 0) {
    rau_set_error("k_dropout_pack: cols=%d cols_pad=%d ldx=%lld must be multiples of 4 and x 16-byte aligned", cols, cols_pad, (long long)ldx);
    return RAU_EINVAL;
  }
  RAU_LAUNCH_PDL(ctx->stream, (dropout_pack_kernel), grid_for(rows * (cols_pad / 4), 2), TPB, 0, x, rows, cols, bits, scale, hi, lo, cols_pad, ldx);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
int k_dropout_bwd_acc(rau_ctx* ctx, const float* dx, int64_t n, const uint32_t* bits, float scale, float* y, int accumulate,
                      bf16* y_hi, bf16* y_lo) {
  RAU_LAUNCH_PDL(ctx->stream, (dropout_bwd_acc_kernel), grid_for(n, 4), TPB, 0, dx, n, bits, scale, y, accumulate, y_hi, y_lo);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
int k_dropout_hops(rau_ctx* ctx, const float* x, int64_t n, int nHop, const uint32_t* bits, int64_t bits_stride, float scale,
                   float* y, bf16* y_hi, bf16* y_lo) {
  RAU_LAUNCH_PDL(ctx->stream, (dropout_hops_kernel), grid_for(n, 2), TPB, 0, x, n, nHop, bits, bits_stride, scale, y, y_hi, y_lo);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
int k_dropout_bwd_hops(rau_ctx* ctx, float* y, int64_t n, int nHop, const uint32_t* bits, int64_t bits_stride, float scale,
                       bf16* y_hi, bf16* y_lo) {
  RAU_LAUNCH_PDL(ctx->stream, (dropout_bwd_hops_kernel), grid_for(n * nHop, 4), TPB, 0, y, n, nHop, bits, bits_stride, scale, y_hi, y_lo);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
int k_dropout_bwd_sum_hops(rau_ctx* ctx, const float* dx, int64_t n, int nHop, const uint32_t* bits, int64_t bits_stride,
                           float scale, float* out) {
  RAU_LAUNCH_PDL(ctx->stream, (dropout_bwd_sum_hops_kernel), grid_for(n, 2), TPB, 0, dx, n, nHop, bits, bits_stride, scale, out);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
int k_tanh_bwd(rau_ctx* ctx, const float* dy, const float* y, int64_t n, float* dx_f, bf16* dx_b, bf16* dx_lo) {
  RAU_LAUNCH_PDL(ctx->stream, (tanh_bwd_kernel), grid_for(n, 4), TPB, 0, dy, y, n, dx_f, dx_b, dx_lo);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
int k_add(rau_ctx* ctx, const float* a, const float* b, int64_t n, float* y) {
  RAU_LAUNCH_PDL(ctx->stream, (add_kernel), grid_for(n, 4), TPB, 0, a, b, n, y);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
int k_axpy(rau_ctx* ctx, float alpha, const float* x, int64_t n, float* y) {
  RAU_LAUNCH_PDL(ctx->stream, (axpy_kernel), grid_for(n, 4), TPB, 0, alpha, x, n, y);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
int k_fill(rau_ctx* ctx, float* x, int64_t n, float v) {
  if (n <= 0) return RAU_OK;
  RAU_LAUNCH_PDL(ctx->stream, (fill_kernel), grid_for(n, 4), TPB, 0, x, n, v);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
int k_to_bf16(rau_ctx* ctx, const float* x, int64_t rows, int cols, int ldx, bf16* y, int ldy, int cols_pad) {
  RAU_LAUNCH_PDL(ctx->stream, (to_bf16_kernel), grid_for(rows * cols_pad, 4), TPB, 0, x, rows, cols, ldx, y, ldy, cols_pad);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
int k_rowdot_sigmoid(rau_ctx* ctx, const float* x, int B, int K, const float* w, const float* b, float* y) {
  RAU_LAUNCH_PDL(ctx->stream, (rowdot_sigmoid_kernel), cdiv((int64_t)B * 32, TPB), TPB, 0, x, B, K, w, b, y);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
int k_dopred_bwd(rau_ctx* ctx, const float* ddo, const float* dop, const float* m, const float* wd, int B, int K,
                 float* dm_acc, float* gwd, float* gbd) {
  RAU_LAUNCH_PDL(ctx->stream, (dopred_bwd_kernel), cdiv(K, TPB), TPB, 0, ddo, dop, m, wd, B, K, dm_acc, gwd, gbd);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
int k_colsum(rau_ctx* ctx, const float* x, int64_t rows, int cols, int ld, float* out, int accumulate, float* out2) {
  int slices = 1;
  if (accumulate && rows >= 512) {   // enough rows to spread over the SMs
    slices = (int)(rows / 128);
    const int want = 148 * 4 / cdiv(cols, 32);
    if (slices > want) slices = want;
    if (slices < 1) slices = 1;
  }
  RAU_LAUNCH_PDL(ctx->stream, (colsum_kernel), dim3(cdiv(cols, 32), slices), dim3(32, 8), 0, x, rows, cols, ld, out, accumulate, out2);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
template <typename T>
int k_rowsum_bms(rau_ctx* ctx, const T* x, int B, int M, int S, int Sp, float* out) {
  RAU_LAUNCH_PDL(ctx->stream, (rowsum_bms_kernel<T>), dim3(M, B >= 16 ? 16 : 1), 64, 0, x, B, M, S, Sp, out);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
template int k_rowsum_bms<float>(rau_ctx*, const float*, int, int, int, int, float*);
template int k_rowsum_bms<bf16>(rau_ctx*, const bf16*, int, int, int, int, float*);
// out[0] += sum x: many CTAs, one atomic each (the single-CTA form below took 60 us for the 0.4 M-element gbs reduction)
__global__ void sum_all_acc_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ out) {
  RAU_PDL_ENTRY();
  __shared__ float red[32];
  float s = 0.0f;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) s += x[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) atomicAdd(out, s);
}

int k_sum_all(rau_ctx* ctx, const float* x, int64_t n, float* out, int accumulate) {
  if (accumulate && n >= 1024) {
    RAU_LAUNCH_PDL(ctx->stream, (sum_all_acc_kernel), (int)((n + 4095) / 4096 < 148 ? (n + 4095) / 4096 : 148), 256, 0, x, n, out);
    RAU_LAUNCH_CHECK(ctx);
    return RAU_OK;
  }
  RAU_LAUNCH_PDL(ctx->stream, (sum_all_kernel), 1, 1024, 0, x, n, out, accumulate);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
int k_select_state(rau_ctx* ctx, const float* S_all, int T, int B, int Q, const float* lengths, float* out) {
  RAU_LAUNCH_PDL(ctx->stream, (select_state_kernel), grid_for((int64_t)B * Q), TPB, 0, S_all, T, B, Q, lengths, out);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
