// k_gemm_simt.cu -- fp32 CUDA-core GEMM with generic strides: the contraction engine of the exact
// mode (RAU_PREC_F32) and of the small odd-shaped products in every mode.  It stands in for the
// cuBLAS sgemm calls behind nn.Linear / 1x1 SpatialConvolution in the reference.
#include "rau_common.cuh"

namespace {
constexpr int BM = 64, BN = 64, BK = 16, NT = 256;

__device__ __forceinline__ void load_tiles(const float* __restrict__ A, int64_t sam, int64_t sak,
                                           const float* __restrict__ B, int64_t sbk, int64_t sbn,
                                           int M, int N, int K, int m0, int n0, int k0,
                                           float (*As)[BM + 4], float (*Bs)[BN + 4]) {
  const int tid = threadIdx.x;
#pragma unroll
  for (int i = 0; i < (BM * BK) / NT; ++i) {
    int idx = tid + i * NT;
    int m, k;
    if (sak == 1) { k = idx % BK; m = idx / BK; } else { m = idx % BM; k = idx / BM; }
    float v = 0.0f;
    if (m0 + m < M && k0 + k < K) v = A[(int64_t)(m0 + m) * sam + (int64_t)(k0 + k) * sak];
    As[k][m] = v;
  }
#pragma unroll
  for (int i = 0; i < (BN * BK) / NT; ++i) {
    int idx = tid + i * NT;
    int n, k;
    if (sbk == 1) { k = idx % BK; n = idx / BK; } else { n = idx % BN; k = idx / BN; }
    float v = 0.0f;
    if (n0 + n < N && k0 + k < K) v = B[(int64_t)(k0 + k) * sbk + (int64_t)(n0 + n) * sbn];
    Bs[k][n] = v;
  }
}

__global__ void __launch_bounds__(NT) simt_gemm_kernel(SimtGemm g) {
  __shared__ float As[BK][BM + 4];
  __shared__ float Bs[BK][BN + 4];
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;
  const int bz = blockIdx.z / g.ksplit, ks = blockIdx.z % g.ksplit;
  const int tx = threadIdx.x % 16, ty = threadIdx.x / 16;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;

  for (int seg = 0; seg < 2; ++seg) {
    const float* Aseg = seg == 0 ? g.A : g.A2;
    const float* Bseg = seg == 0 ? g.B : g.B2;
    if (Aseg == nullptr) continue;
    const int64_t sam = seg == 0 ? g.sam : g.sam2, sak = seg == 0 ? g.sak : g.sak2;
    const int64_t sbk = seg == 0 ? g.sbk : g.sbk2, sbn = seg == 0 ? g.sbn : g.sbn2;
    const int K = seg == 0 ? g.K : g.K2;
    const int kbatch = seg == 0 ? g.kbatch : 1;
    // split-K: slice ks of ksplit owns a contiguous range of the reduced batches (kbatch > 1) or of K
    int kb_lo = 0, kb_hi = kbatch, k_lo = 0, k_hi = K;
    if (g.ksplit > 1) {
      if (kbatch > 1) {
        const int per = (kbatch + g.ksplit - 1) / g.ksplit;
        kb_lo = ks * per; kb_hi = min(kbatch, kb_lo + per);
      } else {
        const int per = ((K + g.ksplit - 1) / g.ksplit + BK - 1) / BK * BK;
        k_lo = ks * per; k_hi = min(K, k_lo + per);
      }
    }
    for (int kb = kb_lo; kb < kb_hi; ++kb) {
      const float* Ab = Aseg + (seg == 0 ? (int64_t)bz * g.bA + (int64_t)kb * g.kA : 0);
      const float* Bb = Bseg + (seg == 0 ? (int64_t)bz * g.bB + (int64_t)kb * g.kB : 0);
      for (int k0 = k_lo; k0 < k_hi; k0 += BK) {
        load_tiles(Ab, sam, sak, Bb, sbk, sbn, g.M, g.N, k_hi, m0, n0, k0, As, Bs);
        __syncthreads();
#pragma unroll
        for (int k = 0; k < BK; ++k) {
          float a[4], b[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) a[i] = As[k][ty * 4 + i];
#pragma unroll
          for (int j = 0; j < 4; ++j) b[j] = Bs[k][tx * 4 + j];
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
        }
        __syncthreads();
      }
    }
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int m = m0 + ty * 4 + i;
    if (m >= g.M) continue;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n >= g.N) continue;
      float v = acc[i][j] * g.alpha;
      if (g.bias_m) v += g.bias_m[m];
      if (g.bias_n) v += g.bias_n[n];
      if (g.bias_n2) v += g.bias_n2[n];
      if (g.bias_bm) v += g.bias_bm[(int64_t)bz * g.M + m];
      if (g.addend) {
        const int64_t di = (int64_t)bz * g.bD + (int64_t)m * g.sdm + (int64_t)n * g.sdn;
        v += g.addend[di];
        if (g.addend2) v += g.addend2[di];
      }
      if (g.act == 1) v = tanhf(v);
      else if (g.act == 2) v = 1.0f / (1.0f + expf(-v));
      if (g.n_valid >= 0 && n >= g.n_valid) v = 0.0f;
      float* c = g.C + (int64_t)bz * g.bC + (int64_t)m * g.scm + (int64_t)n * g.scn;
      if (g.ksplit > 1) { atomicAdd(c, v); continue; }
      if (g.accumulate) v += *c;
      *c = v;
    }
  }
}
}  // namespace

int simt_gemm(rau_ctx* ctx, const SimtGemm& g) {
  if (g.M <= 0 || g.N <= 0) return RAU_OK;
  if (g.ksplit < 1) { rau_set_error("simt_gemm: ksplit < 1"); return RAU_EINVAL; }
  if (g.ksplit > 1 && (!g.accumulate || g.act || g.bias_m || g.bias_n || g.bias_n2 || g.bias_bm || g.addend || g.A2 ||
                       g.alpha != 1.0f || g.n_valid >= 0)) {
    rau_set_error("simt_gemm: split-K requires a plain accumulating product");
    return RAU_EINVAL;
  }
  dim3 grid(cdiv(g.N, BN), cdiv(g.M, BM), g.batch * g.ksplit);
  simt_gemm_kernel<<<grid, NT, 0, ctx->stream>>>(g);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
