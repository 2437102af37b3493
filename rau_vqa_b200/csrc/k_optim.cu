// k_optim.cu -- gradient noise + per-group L2 clip (F:617-648) and the utils/optim_updates.lua update rules
// (OU:7-87) as two fused passes over a flat vector: pass 1 adds noise and reduces ||g||^2 on the device
// (no host sync, unlike the three :norm() calls of the reference), pass 2 applies the clip scale and the
// optimizer in one sweep.  Traffic: adam 4 reads + 4 writes per parameter, the HBM floor for this update.
#include "rau_kernels.cuh"
#include <math.h>

namespace {
__global__ void __launch_bounds__(256) noise_norm_kernel(float* __restrict__ g, int64_t n, float std,
                                                         const float* __restrict__ noise, uint2 key,
                                                         uint32_t stream_lo, uint32_t stream_hi, double* __restrict__ norm2,
                                                         const StepState* __restrict__ ss) {
  RAU_PDL_ENTRY();
  __shared__ float red[32];
  if (ss) {
    const unsigned long long sid = (((unsigned long long)stream_hi << 32) | stream_lo) ^ (ss->step << 24);
    stream_lo = (uint32_t)sid;
    stream_hi = (uint32_t)(sid >> 32);
    if (std > 0.0f) std = ss->noise_std;
  }
  float acc = 0.0f;
  const int64_t nq = (n + 3) / 4;
  for (int64_t q = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; q < nq; q += (int64_t)gridDim.x * blockDim.x) {
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    if (noise == nullptr && std > 0.0f) {
      const uint4 r = philox4x32(make_uint4((uint32_t)q, (uint32_t)(q >> 32), stream_lo, stream_hi), key);
      // Box-Muller on (0,1] uniforms
      const float u0 = ((float)r.x + 1.0f) * 2.3283064365386963e-10f, u1 = (float)r.y * 2.3283064365386963e-10f;
      const float u2 = ((float)r.z + 1.0f) * 2.3283064365386963e-10f, u3 = (float)r.w * 2.3283064365386963e-10f;
      const float ra = sqrtf(-2.0f * logf(u0)), rb = sqrtf(-2.0f * logf(u2));
      float s0, c0, s1, c1;
      sincospif(2.0f * u1, &s0, &c0);
      sincospif(2.0f * u3, &s1, &c1);
      z[0] = ra * c0 * std; z[1] = ra * s0 * std; z[2] = rb * c1 * std; z[3] = rb * s1 * std;
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int64_t i = q * 4 + k;
      if (i < n) {
        float v = g[i] + (noise ? noise[i] : z[k]);
        g[i] = v;
        acc = fmaf(v, v, acc);
      }
    }
  }
  acc = block_sum(acc, red);
  if (threadIdx.x == 0) atomicAdd(norm2, (double)acc);
}

struct OptArgs {
  int optim; int64_t n; float clip, lr, h0, h1, h2, step;
};

__global__ void __launch_bounds__(256) clip_optim_kernel(OptArgs a, float* __restrict__ x, float* __restrict__ g,
                                                         const double* __restrict__ norm2, float* __restrict__ s0,
                                                         float* __restrict__ s1, float* __restrict__ norm_out,
                                                         const StepState* __restrict__ ss, int group,
                                                         const unsigned int* __restrict__ failed) {
  RAU_PDL_ENTRY();
  // a persistent recurrence kernel of this step gave up on a peer CTA: the gradients are garbage, leave x and the
  // optimizer state alone (the host reports the failure at its next call into the library)
  if (failed != nullptr && a.optim >= 0 && *failed != 0u) return;
  float scale = 1.0f;
  if (ss && group >= 0) a.step = ss->opt_step[group];
  if (norm2 != nullptr) {
    const float nrm = (float)sqrt(*norm2);
    if (a.clip > 0.0f && nrm > a.clip) scale = a.clip / nrm;          // F:628-630
    if (norm_out && blockIdx.x == 0 && threadIdx.x == 0) *norm_out = nrm;
  }
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < a.n; i += (int64_t)gridDim.x * blockDim.x) {
    float gi = g[i];
    if (scale != 1.0f) { gi *= scale; g[i] = gi; }
    if (a.optim < 0) continue;   // clip only (rau_noise_clip)
    float xi = x[i];
    switch (a.optim) {
      case RAU_OPT_SGD: xi -= a.lr * gi; break;                                         // OU:8
      case RAU_OPT_SGDM: { float v = s0[i] * a.h0 + a.lr * gi; s0[i] = v; xi -= v; } break;  // OU:16-18
      case RAU_OPT_SGDMOM: {                                                             // OU:26-30
        const float tmp = s0[i];
        const float m = tmp * a.h0 - a.lr * gi;
        s0[i] = m;
        xi = xi - a.h0 * tmp + (1.0f + a.h0) * m;
      } break;
      case RAU_OPT_ADAGRAD: { float m = s0[i] + gi * gi; s0[i] = m; xi -= a.lr * gi / (sqrtf(m) + a.h0); } break;  // OU:39-42
      case RAU_OPT_RMSPROP: {                                                            // OU:52-56
        const float m = s0[i] * a.h0 + (1.0f - a.h0) * gi * gi;
        s0[i] = m;
        xi -= a.lr * gi / (sqrtf(m) + a.h1);
      } break;
      case RAU_OPT_ADAM: {                                                               // OU:76-86
        const float m = s0[i] * a.h0 + (1.0f - a.h0) * gi;
        const float v = s1[i] * a.h1 + (1.0f - a.h1) * gi * gi;
        s0[i] = m; s1[i] = v;
        xi -= a.step * m / (sqrtf(v) + a.h2);
      } break;
    }
    x[i] = xi;
  }
}
}  // namespace

int k_noise_norm(rau_ctx* ctx, float* g, int64_t n, float std, const float* noise_override, uint64_t seed,
                 uint64_t stream_id, double* norm2_out) {
  int64_t blocks = (n / 4 + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  RAU_LAUNCH_PDL(ctx->stream, (noise_norm_kernel), (int)blocks, 256, 0, g, n, std, noise_override,
      make_uint2((uint32_t)seed, (uint32_t)(seed >> 32)), (uint32_t)stream_id, (uint32_t)(stream_id >> 32), norm2_out,
      ctx->ss_active);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}

int k_clip_optim(rau_ctx* ctx, int optim, int64_t n, float* x, float* g, const double* norm2, float clip, float lr,
                 float h0, float h1, float h2, float* s0, float* s1, int64_t t, float* norm_out, int group) {
  OptArgs a;
  a.optim = optim; a.n = n; a.clip = clip; a.lr = lr; a.h0 = h0; a.h1 = h1; a.h2 = h2; a.step = lr;
  if (optim == RAU_OPT_ADAM) {   // OU:80-83, evaluated in double on the host like Lua numbers
    const double bc1 = 1.0 - pow((double)h0, (double)t), bc2 = 1.0 - pow((double)h1, (double)t);
    a.step = (float)((double)lr * sqrt(bc2) / bc1);
  }
  int64_t blocks = (n + 255) / 256;
  if (blocks < 1) blocks = 1;
  if (blocks > 148 * 8) blocks = 148 * 8;
  RAU_LAUNCH_PDL(ctx->stream, (clip_optim_kernel), (int)blocks, 256, 0, a, x, g, norm2, s0, s1, norm_out, ctx->ss_active, group,
      (const unsigned int*)ctx->d_err);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
