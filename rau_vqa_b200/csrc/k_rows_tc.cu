// k_rows_tc.cu -- the persistent tcgen05 engine of the answering unit's heavy products (SURVEY.md 8a rows a6/a7 and
// their backward).  All image-side activations of a hop live in a "rows" layout: one row per grid cell of one image,
// r = b*S + s, features contiguous ([R, C] dropped-out features, [R, M] i_embed output I, [R, A] attention hidden E).
// In that layout the reference's per-image 1x1 convolutions (F:240, F:247) are plain [R, K] x [K, N] products whose
// 128-row tiles are always full (no 196 -> 256 padding), and the backward weight gradients are [K = R] reductions.
//
//   D[M, N] = sum_k A[m, k] * B[n, k]         A, B: bf16 (hi [, lo]) arrays, K-major ([rows, K]) or MN-major ([K, rows])
//
// One CTA per SM walks a static list of work items (tile_m, tile_n, k_slice).  Warp roles:
//   warp 0      TMA producer: 48 KB stages (A 128 x BK, B 256 x BK; hi and lo tiles in bf16x3 mode, BK = 32; BK = 64 else)
//   warp 1      MMA issuer: tcgen05.mma kind::f16, M = 128, N = 256, fp32 accumulators in TMEM; bf16x3 issues
//               hi*hi + hi*lo + lo*hi into the same accumulator.  Two 256-column accumulators alternate between items,
//               so the epilogue of item i overlaps the MMAs of item i+1.
//   warps 2..9  epilogue: tcgen05.ld (one TMEM lane = one row per thread, 32 columns at a time; the two warps of a lane
//               quarter split the 256 columns) -> fused math -> swizzled staging in shared memory -> TMA store (bf16
//               hi/lo or fp32) or TMA reduce-add (split-K weight gradients).
// Epilogues (template parameter):
//   EPI_PLAIN  D (tests, dX)                                   EPI_RED   D added into fp32 global (split-K wgrad)
//   EPI_TANH   tanh(D + bias[n]) -> bf16 hi/lo                 (i_embed F:240-241)
//   EPI_ATT    E = tanh(D + bias[n] + rowvec[b(r), n]) -> fp32, logit[r] = sum_n colw[n] E[r, n]   (F:246-251)
//   EPI_DY     dY = (D + rowvec[b(r), n] * rowscale[r]) * (1 - I[r, n]^2) -> bf16 hi/lo, colsum[n] += sum_r dY
#include "rau_model.cuh"
#include <cuda.h>
#include <cuda_fp16.h>
#include <stdlib.h>

namespace {

constexpr int RT_BM = 128, RT_BN = 256;
constexpr int RT_STAGE = 48 * 1024;
constexpr int RT_THREADS = 352;          // A producer, MMA issuer, 8 epilogue warps, B producer
constexpr int RT_STG_WARP = 4096;          // per epilogue warp: one 32 x 128 B (fp32) or two 32 x 64 B (bf16 hi, lo) buffers
constexpr int RT_MAXSTAGES = 8;
constexpr int RT_SMEM_BUDGET = 4 * RT_STAGE + 8 * RT_STG_WARP;   // pipeline + staging bytes a CTA may use (224 KB)

enum { EPI_PLAIN = 0, EPI_RED = 1, EPI_TANH = 2, EPI_ATT = 3, EPI_DY = 4, EPI_LINEAR = 5, EPI_LSTM = 6 };

struct RtParams {
  CUtensorMap mapA[2], mapB[2];   // [hi, lo] operand tiles
  CUtensorMap mapA2[2], mapB2[2]; // optional second K segment (x Wi^T + h Wh^T of an LSTM layer)
  CUtensorMap mapO[3];            // outputs: bf16 (hi, lo) or fp32 ([0]); EPI_LINEAR: fp32 [0] + optional bf16 hi [1], lo [2]
  int M, N, K;
  int a_mn, b_mn, x3, BK, nkb;
  int nkb1;                        // k-blocks of the first segment (== nkb without a second segment)
  int cg2;                         // launched as CTA pairs (cta_group::2, M = 256 per item)
  int a_swap, b_swap;              // bf16x3: plane 0 of the operand's 3-D box is the lo array (lo lies below hi in memory)
  int BN;                          // accumulator columns per tile: 64, 128 or 256
  int stg_warp;                    // staging bytes per epilogue warp
  int out_f, out_hi;               // EPI_LINEAR: which outputs exist
  const float* bias2;              // [N]
  const float* addend; const float* addend2; long long ldadd;   // [M, ldadd]
  int act;                         // EPI_LINEAR: 0 none, 1 tanh, 2 sigmoid
  int vec;                         // EPI_LINEAR: bias / addend pointers are 16-byte aligned and N % 32 == 0
  unsigned long long* dbg;         // RAU_ROWS_TRACE: per-CTA clock stamps [16] (tools/rows_trace.py), else NULL
  // EPI_LSTM: the cell update fused behind the gate product (A:12-25, D:47-61).  Accumulator columns are permuted gate
  // pre-activations: every 32-column chunk holds (i, f, o, g) of 8 consecutive hidden units.
  const float* c_prev; long long ldcp;         // [M, H] (NULL = zeros)
  float* c_out; long long ldc;                 // [M, H]
  float* h_out; long long ldh;                 // [M, H]
  float* lsaved; long long plane;              // 5 planes (i, f, o, g, tanh c) of M*H floats
  bf16* hpk_hi; bf16* hpk_lo; long long ldhp;  // packed h for the next product (optional)
  int ksplit, kb_per;
  int tiles_m, tiles_n, stages, stage_bytes;
  unsigned int ksplit_magic, tiles_n_magic;   // ceil(2^32 / d): item -> (tile, k slice) -> (tm, tn) without integer divisions
                                              // (a runtime division is ~150 dependent cycles; three of them sat between the
                                              // prologue and every CTA's first TMA)
  int out_lo;                      // bf16 outputs: also write the lo array
  const float* bias;
  const float* rowvec;             // [B, N]
  const float* colw;               // [N]
  float* rowout;                   // [M]
  const float* rowscale;           // [M]
  const bf16* aux_hi; const bf16* aux_lo; long long ldaux;
  float* colsum;                   // [N]
  int S;
  float alpha;
  int fast_tanh;
  int ew;                          // epilogue warps of the launch (8, or 16 for the 1-pass i_embed product)
  int f16;                         // the operands are fp16, not bf16 (single plane)
  int of16;                        // EPI_TANH / EPI_DY: the output is ONE fp16 plane (not bf16 hi [, lo])
  int af16;                        // EPI_DY: the saved activation (aux) is one fp16 plane
  float gscale;                    // EPI_DY: power-of-two gradient scale carried by the A operand dZ and by the output dY
};

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
// floor(x / d) for 0 <= x, d < 2^16 with magic = ceil(2^32 / d) (exact in that range); d == 1 has no 32-bit magic
__device__ __forceinline__ int fast_div(int x, int d, unsigned int magic) {
  return d == 1 ? x : (int)__umulhi((unsigned int)x, magic);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(smem_u32(bar)), "r"(parity)
        : "memory");
  } while (!ok);
}
// one lane of a converged warp (the CUTLASS elect_one_sync idiom): lets the compiler keep the surrounding loop
// warp-uniform, so the single-thread TMA / MMA instructions are issued without a per-instruction lane loop
__device__ __forceinline__ uint32_t elect_one() {
  uint32_t pred = 0;
  asm volatile(
      "{\n\t.reg .b32 rx;\n\t.reg .pred px;\n\t"
      "elect.sync rx|px, %1;\n\t"
      "@px mov.s32 %0, 1;\n\t}"
      : "+r"(pred)
      : "r"(0xffffffffu));
  return pred;
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
// ---- cta_group::2 (a CTA pair on one TPC runs M = 256 MMAs; each CTA stages its own 128 rows of A and HALF of B):
// the peer's TMA loads signal the LEADER's barrier (peer bit of the shared::cluster address cleared), tcgen05.commit
// multicasts its arrival to both CTAs, and accumulator-drained arrivals go to the leader.  PTX forms as in CUTLASS
// (cute/arch/copy_sm100_tma.hpp SM100_TMA_2SM_LOAD_*, cutlass/arch/barrier.h umma_arrive_multicast_2x1SM).
constexpr uint32_t RT_PEER_MASK = 0xFEFFFFFFu;
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ void tma_load_2d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar) & RT_PEER_MASK), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tma_load_3d_2sm(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1, int c2) {
  asm volatile(
      "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(map), "r"(smem_u32(bar) & RT_PEER_MASK), "r"(c0), "r"(c1), "r"(c2)
      : "memory");
}
__device__ __forceinline__ void umma_f16_2sm(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit_2sm(uint64_t* bar) {   // arrives on this barrier in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
                   smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {   // arrive on the even CTA's copy of this barrier
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & RT_PEER_MASK) : "memory");
}
__device__ __forceinline__ void tma_store_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void tma_reduce_add_2d(const CUtensorMap* map, const void* src, int c0, int c1) {
  asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];" ::"l"(map),
               "r"(smem_u32(src)), "r"(c0), "r"(c1)
               : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

__device__ __forceinline__ void umma_f16(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float* v) {
  uint32_t r[32];
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
  for (int k = 0; k < 32; ++k) v[k] = __uint_as_float(r[k]);
}

__device__ __forceinline__ float exp2f_ftz(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float rcp_ftz(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// tanh to ~1e-7 ABSOLUTE (bf16x3 mode): 1 - 2/(e^2|x| + 1).  Near zero the relative error grows, which is harmless here:
// the result feeds contractions and (1 - y^2), both of which see the absolute error only.
// (no |x| / copysign: e^{2x} -> 0 gives -1, -> inf gives rcp(inf) = 0 and +1, and the absolute error is the same either way)
__device__ __forceinline__ float tanh_acc(float x) {
  const float e = exp2f_ftz(2.8853900817779268f * x);
  return fmaf(-2.0f, rcp_ftz(e + 1.0f), 1.0f);
}
__device__ __forceinline__ float tanh_hw(float x) {
  float y;
  asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// shared-memory matrix descriptor (cute::UMMA::SmemDescriptor): start>>4 [0,14) | LBO>>4 [16,30) | SBO>>4 [32,46) |
// version = 1 [46,48) | layout type [61,64): 2 = SWIZZLE_128B, 4 = SWIZZLE_64B
__device__ __forceinline__ uint64_t make_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes, uint32_t layout) {
  return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)(lbo_bytes >> 4) << 16) | ((uint64_t)(sbo_bytes >> 4) << 32) |
         (1ull << 46) | ((uint64_t)layout << 61);
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
// hi = bf16(a), lo = bf16(a - hi) for a pair
// (ONE packed conversion per pair for hi -- F2FP on the ALU pipe -- instead of two scalar F2F, which share the MUFU pipe with
// the epilogues' ex2 / rcp: ncu showed the one-pass i_embed epilogue queueing on that pipe)
__device__ __forceinline__ void split_pair(float a, float b, uint32_t& hi, uint32_t& lo) {
  hi = pack_bf16x2(a, b);   // low half = bf16(a), high half = bf16(b), round to nearest
  lo = pack_bf16x2(a - __uint_as_float(hi << 16), b - __uint_as_float(hi & 0xffff0000u));
}
// fp16 pair, round-to-nearest, saturating (a scaled gradient that overflows becomes 65504, never inf)
__device__ __forceinline__ uint32_t pack_f16x2(float a, float b) {
  uint32_t r;
  asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(b), "f"(a));   // upper half <- first source operand
  return r;
}
__device__ __forceinline__ float hf_lo(uint32_t u) { return __half2float(__ushort_as_half((unsigned short)(u & 0xffffu))); }
__device__ __forceinline__ float hf_hi(uint32_t u) { return __half2float(__ushort_as_half((unsigned short)(u >> 16))); }
__device__ __forceinline__ float bf_lo(uint32_t u) { return __uint_as_float(u << 16); }
__device__ __forceinline__ float bf_hi(uint32_t u) { return __uint_as_float(u & 0xffff0000u); }

// write 32 consecutive 32-bit words of this lane's row into a 32-row x 128-byte SWIZZLE_128B staging buffer
__device__ __forceinline__ void stage_row128(uint32_t buf, int lane, const uint32_t* w) {
  const uint32_t base = buf + (uint32_t)lane * 128u;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const uint32_t a = base + (uint32_t)((j ^ (lane & 7)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w[4 * j]), "r"(w[4 * j + 1]), "r"(w[4 * j + 2]),
                 "r"(w[4 * j + 3])
                 : "memory");
  }
}
// write 16 consecutive 32-bit words (32 bf16) of this lane's row into a 32-row x 64-byte SWIZZLE_64B staging buffer
__device__ __forceinline__ void stage_row64(uint32_t buf, int lane, const uint32_t* w) {
  const uint32_t base = buf + (uint32_t)lane * 64u;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    const uint32_t a = base + (uint32_t)((j ^ ((lane >> 1) & 3)) << 4);
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(w[4 * j]), "r"(w[4 * j + 1]), "r"(w[4 * j + 2]),
                 "r"(w[4 * j + 3])
                 : "memory");
  }
}

// lane L ends with sum over lanes of x[L] (31 shuffles)
__device__ __forceinline__ float warp_transpose_sum32(float* x, int lane) {
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < off; ++i) {
      const float send = up ? x[i] : x[i + off];
      const float keep = up ? x[i + off] : x[i];
      x[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
  }
  return x[0];
}

// EW = epilogue warps (8 or 16; warps 2 .. EW+1, the B producer is warp EW+2).  16 warps put four instead of two warps on
// every scheduler: the 1-pass fp16 i_embed product is epilogue-latency bound (tmem load -> tanh -> split -> staging -> TMA
// store per 32-column chunk; ncu: issue slots 35 % busy with 2 warps per scheduler), not tensor bound.
template <int EPI, int X3, int NSTEPS, int CG2, int EW>
__global__ void __launch_bounds__(32 * (EW + 3), 1) rows_gemm_kernel(const __grid_constant__ RtParams p) {
  asm volatile("griddepcontrol.launch_dependents;");   // PDL: the next kernel may begin its prologue
  // CG2: launched as clusters of two CTAs (one TPC).  The pair owns 256 rows x BN columns per item: each CTA stages its own
  // 128 rows of A and its half of the B tile, the leader (cluster rank 0) issues cta_group::2 MMAs with M = 256, each CTA
  // drains its own 128 accumulator lanes.  Halves the B-operand shared-memory traffic per SM (the 1-CTA bound).
  const uint32_t cta_rank = CG2 ? cluster_ctarank() : 0u;
  const bool leader = cta_rank == 0;
  const int worker = CG2 ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int nworkers = CG2 ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[RT_MAXSTAGES], empty_bar[RT_MAXSTAGES], tfull_bar[2], tempty_bar[2];
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  uint8_t* staging = smem + (size_t)p.stages * p.stage_bytes;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int items = (CG2 ? (p.tiles_m + 1) / 2 : p.tiles_m) * p.tiles_n * p.ksplit;
  unsigned long long* dbg = p.dbg ? p.dbg + (size_t)blockIdx.x * 16 : nullptr;
#define RT_STAMP(slot) do { if (dbg) dbg[slot] = clock64(); } while (0)
  if (threadIdx.x == 0) RT_STAMP(0);
  unsigned long long gt0 = 0;   // (trace only) wall-clock ns of this CTA's lifetime go to slot 15: cycles / ns = the SM clock under load
  if (dbg && threadIdx.x == 0) asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt0));

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 2); mbar_init(&empty_bar[s], 1); }
    for (int s = 0; s < 2; ++s) { mbar_init(&tfull_bar[s], 1); mbar_init(&tempty_bar[s], CG2 ? 2 * EW : EW); }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.mapA[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.mapB[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.mapO[0]) : "memory");
  }
  if (warp == 1) {
    if (CG2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (CG2) cluster_sync_all();   // both CTAs' barriers are initialised before any remote arrive / peer TMA completion
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;
  if (threadIdx.x == 0) RT_STAMP(1);   // prologue done
  asm volatile("griddepcontrol.wait;" ::: "memory");   // PDL: everything earlier kernels wrote is visible from here on

  const uint32_t a_bytes = (uint32_t)RT_BM * p.BK * 2, b_bytes = (uint32_t)(CG2 ? p.BN / 2 : p.BN) * p.BK * 2;
  constexpr int nt = X3 ? 2 : 1;

  if (warp == 0 || warp == EW + 2) {
    // ===================== TMA producers: warp 0 feeds operand A, warp 10 operand B (one thread issues roughly one
    // cp.async.bulk.tensor per 250 cycles, so the two operands are issued from different warps, and the hi and lo
    // tiles of the bf16x3 split travel as the two planes of ONE 3-D box)
    {
      const bool isB = warp == EW + 2;
      const int mn = isB ? p.b_mn : p.a_mn;
      const uint32_t my_bytes = nt * (isB ? b_bytes : a_bytes);
      const int nchunk = (isB ? (CG2 ? p.BN / 2 : p.BN) : RT_BM) / 64;
      uint32_t it = 0;
      int st = 0;          // stage ring position and phase, carried instead of it % stages, it / stages
      uint32_t ph = 0;
      for (int item = worker; item < items; item += nworkers) {
        const int tile = fast_div(item, p.ksplit, p.ksplit_magic);
        const int ks = item - tile * p.ksplit;
        const int tm = fast_div(tile, p.tiles_n, p.tiles_n_magic);
        const int tn = tile - tm * p.tiles_n;
        const int r0 = CG2 ? (isB ? tn * p.BN + (int)cta_rank * (p.BN / 2) : tm * 2 * RT_BM + (int)cta_rank * RT_BM)
                           : (isB ? tn * p.BN : tm * RT_BM);
        const int kb0 = ks * p.kb_per, kb1 = min(p.nkb, kb0 + p.kb_per);
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          const bool seg2 = kb >= p.nkb1;
          const CUtensorMap* map = isB ? (seg2 ? &p.mapB2[0] : &p.mapB[0]) : (seg2 ? &p.mapA2[0] : &p.mapA[0]);
          mbar_wait(&empty_bar[st], ph ^ 1u);
          if (elect_one()) {
            if (it == 0 && !isB) RT_STAMP(2);      // first TMA about to issue
            uint8_t* dst = smem + (size_t)st * p.stage_bytes + (isB ? nt * a_bytes : 0);
            if (!CG2) mbar_expect_tx(&full_bar[st], my_bytes);
            else if (leader) mbar_expect_tx(&full_bar[st], 2 * my_bytes);   // own tile + the peer's, both land on this barrier
            const int k0 = (seg2 ? kb - p.nkb1 : kb) * p.BK;
            if (CG2 && mn) {   // boxes of 64 rows (contiguous) x BK k [x 2 planes], completing on the leader's barrier
              for (int u = 0; u < nchunk; ++u) {
                if (X3) tma_load_3d_2sm(dst + u * (nt * p.BK * 128), map, &full_bar[st], r0 + 64 * u, k0, 0);
                else tma_load_2d_2sm(dst + u * (p.BK * 128), map, &full_bar[st], r0 + 64 * u, k0);
              }
            } else if (CG2) {
              if (X3) tma_load_3d_2sm(dst, map, &full_bar[st], k0, r0, 0);
              else tma_load_2d_2sm(dst, map, &full_bar[st], k0, r0);
            } else if (mn) {   // boxes of 64 rows (contiguous) x BK k [x 2 planes]
              for (int u = 0; u < nchunk; ++u) {
                if (X3) tma_load_3d(dst + u * (nt * p.BK * 128), map, &full_bar[st], r0 + 64 * u, k0, 0);
                else tma_load_2d(dst + u * (p.BK * 128), map, &full_bar[st], r0 + 64 * u, k0);
              }
            } else {
              if (X3) tma_load_3d(dst, map, &full_bar[st], k0, r0, 0);
              else tma_load_2d(dst, map, &full_bar[st], k0, r0);
            }
            if (it == 0 && !isB) RT_STAMP(10);     // stage 0 fully issued
            if (it == 1 && !isB) RT_STAMP(11);     // stage 1 fully issued
            if (it == 7 && !isB) RT_STAMP(13);     // stage 7 fully issued
          }
          __syncwarp();
          if (++st == p.stages) { st = 0; ph ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: the warp walks the loops converged, one elected lane issues (pair leader only)
    if (!CG2 || leader) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): D=f32 [4,6) | A=bf16 [7,10) | B=bf16 [10,13) |
      // a_major [15] | b_major [16] | N>>3 [17,23) | M>>4 [24,29)
      const uint32_t fmt = p.f16 ? 0u : 1u;   // A / B element format: 0 = f16, 1 = bf16
      const uint32_t idesc = (1u << 4) | (fmt << 7) | (fmt << 10) | ((uint32_t)p.a_mn << 15) | ((uint32_t)p.b_mn << 16) |
                             ((uint32_t)(p.BN >> 3) << 17) | ((uint32_t)((CG2 ? 2 * RT_BM : RT_BM) >> 4) << 24);
      // K-major tile: rows of BK*2 bytes (64 B -> SWIZZLE_64B, 128 B -> SWIZZLE_128B), 8-row groups SBO apart; a 16-k
      //   step is +32 bytes inside the swizzled row.
      // MN-major tile: 64-wide chunks (LBO apart), 8-k-row groups 1024 bytes apart (SBO), SWIZZLE_128B; a 16-k step is
      //   16 rows = 2048 bytes.
      constexpr uint32_t BK_ = NSTEPS * 16;
      constexpr uint32_t k_layout = BK_ == 32 ? 4u : 2u, k_sbo = BK_ == 32 ? 512u : 1024u;
      constexpr uint32_t mn_chunk = BK_ * 128u;
      constexpr uint32_t NT = X3 ? 2u : 1u;
      const uint64_t a_step = p.a_mn ? (2048u >> 4) : (32u >> 4), b_step = p.b_mn ? (2048u >> 4) : (32u >> 4);
      // plane 0 / plane 1 of an operand's box: K-major tiles a_bytes (b_bytes) apart; MN-major chunks interleave the
      // planes, so a chunk pitch (LBO) spans both and plane 1 starts BK*128 bytes in.  p.a_swap: plane 0 is `lo`.
      const uint64_t a_p1 = (uint64_t)((p.a_mn ? mn_chunk : a_bytes) >> 4), b_p1 = (uint64_t)((p.b_mn ? mn_chunk : b_bytes) >> 4);
      const uint64_t a_hi_off = p.a_swap ? a_p1 : 0, a_lo_off = p.a_swap ? 0 : a_p1;
      const uint64_t b_hi_off = p.b_swap ? b_p1 : 0, b_lo_off = p.b_swap ? 0 : b_p1;
      uint32_t it = 0, li = 0;
      int st = 0;
      uint32_t ph = 0;
      for (int item = worker; item < items; item += nworkers, ++li) {
        const int ks = item - fast_div(item, p.ksplit, p.ksplit_magic) * p.ksplit;
        const int kb0 = ks * p.kb_per, kb1 = min(p.nkb, kb0 + p.kb_per);
        const uint32_t ab = li & 1u, aph = (li >> 1) & 1u;
        mbar_wait(&tempty_bar[ab], aph ^ 1u);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const uint32_t dcol = tmem_base + ab * (uint32_t)RT_BN;
        for (int kb = kb0; kb < kb1; ++kb, ++it) {
          mbar_wait(&full_bar[st], ph);
          asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
          if (elect_one()) {
            if (it == 0) RT_STAMP(3);      // first stage landed
            if (it == 1) RT_STAMP(4);      // second stage landed
            const uint32_t sa = smem_u32(smem + (size_t)st * p.stage_bytes);
            const uint32_t sb = sa + NT * a_bytes;
            const uint64_t da0 = p.a_mn ? make_desc(sa, NT * mn_chunk, 1024u, 2u) : make_desc(sa, 16u, k_sbo, k_layout);
            const uint64_t db0 = p.b_mn ? make_desc(sb, NT * mn_chunk, 1024u, 2u) : make_desc(sb, 16u, k_sbo, k_layout);
            const uint64_t da_hi = da0 + a_hi_off, da_lo = da0 + a_lo_off, db_hi = db0 + b_hi_off, db_lo = db0 + b_lo_off;
            const uint32_t acc0 = kb > kb0 ? 1u : 0u;
#pragma unroll
            for (int k = 0; k < NSTEPS; ++k) {
              const uint64_t oa = (uint64_t)k * a_step, ob = (uint64_t)k * b_step;
              if (CG2) {
                umma_f16_2sm(dcol, da_hi + oa, db_hi + ob, idesc, k == 0 ? acc0 : 1u);
                if (X3) {
                  umma_f16_2sm(dcol, da_hi + oa, db_lo + ob, idesc, 1u);
                  umma_f16_2sm(dcol, da_lo + oa, db_hi + ob, idesc, 1u);
                }
              } else {
                umma_f16(dcol, da_hi + oa, db_hi + ob, idesc, k == 0 ? acc0 : 1u);
                if (X3) {
                  umma_f16(dcol, da_hi + oa, db_lo + ob, idesc, 1u);
                  umma_f16(dcol, da_lo + oa, db_hi + ob, idesc, 1u);
                }
              }
            }
            if (CG2) umma_commit_2sm(&empty_bar[st]);   // frees the stage in both CTAs
            else umma_commit(&empty_bar[st]);           // frees the stage once these MMAs have read it
            if (it == 0) RT_STAMP(12);       // MMAs of k-block 0 issued
            if (it == 7) RT_STAMP(14);       // MMAs of k-block 7 issued
          }
          __syncwarp();
          if (++st == p.stages) { st = 0; ph ^= 1u; }
        }
        if (elect_one()) {
          if (CG2) umma_commit_2sm(&tfull_bar[ab]);   // both CTAs' epilogues
          else umma_commit(&tfull_bar[ab]);           // accumulator of this item complete
          if (li == 0) RT_STAMP(5);          // all MMAs of the first item issued
        }
        __syncwarp();
      }
    }
  } else {
    // ===================== epilogue warps 2..9: TMEM lanes 32*(warp%4) .. +31; the two warps of a lane quarter split
    // the 256 accumulator columns in halves and walk them in 32-column chunks
    // (a warp reads the TMEM lanes 32 * (warp % 4) ..; the EW / 4 warps of a lane quarter split the accumulator's columns)
    const int q = warp & 3, half = (warp - 2) >> 2;
    constexpr int NPART = EW / 4;
    uint8_t* stg_p = staging + (size_t)(warp - 2) * p.stg_warp;
    const uint32_t stg0 = smem_u32(stg_p), stg1 = stg0 + 2048u;
    uint32_t li = 0;
    // EPI_DY operand prefetch (registers, filled one chunk / one item ahead): ah/al = the saved activation of the next
    // chunk, fetched COALESCED (a lane reads 16 bytes of row lane/4 + 8i: one instruction covers 8 rows x 64 B instead of
    // 32 rows x 16 B, 8x fewer L1 wavefronts than a lane-per-row fetch; the warp's scratch then hands every lane its own
    // row); rvp = this warp's slice of the per-image row vector for the (at most two) images its 32 rows belong to
    uint4 ah[4], al[4];
    float4 rvp[2];
    float rs_next = 0.0f;   // the next item's row scale (p[r]) for this thread's row
    bool dy_ready = false;
    const int arow = lane >> 2, aseg = lane & 3;
    const int ncols_w = p.BN / 2;   // accumulator columns of this warp (EPI_DY runs with EW = 8)
    auto aux_fetch = [&](int rowbase_x, int ncx) {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int row = min(rowbase_x + arow + 8 * i, p.M - 1);
        ah[i] = __ldg(reinterpret_cast<const uint4*>(p.aux_hi + (long long)row * p.ldaux + ncx) + aseg);
        al[i] = p.aux_lo ? __ldg(reinterpret_cast<const uint4*>(p.aux_lo + (long long)row * p.ldaux + ncx) + aseg)
                         : make_uint4(0u, 0u, 0u, 0u);
      }
    };
    auto rv_fetch = [&](int rowbase_x, int nbase) {
      const int b0 = min(rowbase_x, p.M - 1) / p.S, b1 = min(rowbase_x + 31, p.M - 1) / p.S;
      const int per = ncols_w >> 2;   // float4 per image
#pragma unroll
      for (int j = 0; j < 2; ++j) {
        const int idx = lane + 32 * j;
        rvp[j] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (idx < 2 * per) {
          const int col = min(nbase + (idx % per) * 4, p.N - 4);
          rvp[j] = __ldg(reinterpret_cast<const float4*>(p.rowvec + (long long)(idx < per ? b0 : b1) * p.N + col));
        }
      }
    };
    for (int item = worker; item < items; item += nworkers, ++li) {
      const int tile = fast_div(item, p.ksplit, p.ksplit_magic);
      const int tm = fast_div(tile, p.tiles_n, p.tiles_n_magic);
      const int tn = tile - tm * p.tiles_n;
      const int m0 = CG2 ? tm * 2 * RT_BM + (int)cta_rank * RT_BM : tm * RT_BM, n0 = tn * p.BN;
      const uint32_t ab = li & 1u, aph = (li >> 1) & 1u;
      const int r = m0 + q * 32 + lane;          // global row of this thread
      const bool r_ok = r < p.M;
      const int rr = r_ok ? r : p.M - 1;
      const int rowbase = m0 + q * 32;           // first row of this warp's 32-row slab
      mbar_wait(&tfull_bar[ab], aph);
      if (li == 0 && warp == 2 && lane == 0) RT_STAMP(6);   // first accumulator ready
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      const uint32_t taddr = tmem_base + ab * (uint32_t)RT_BN + ((uint32_t)(q * 32) << 16);
      float rowacc = 0.0f;
      const float* rv = nullptr;
      float rs = 0.0f;
      if (EPI == EPI_ATT || EPI == EPI_DY) rv = p.rowvec + (long long)(rr / p.S) * p.N;
      if (EPI == EPI_DY) rs = (dy_ready ? rs_next : p.rowscale[rr]) * p.gscale;   // (dY is written times gscale)
      const int cper = (p.BN / 32 + NPART - 1) / NPART;   // 32-column chunks per warp
      const int c_lo = half * cper, c_hi = min(p.BN / 32, c_lo + cper);
      bool released = false;
      uint32_t rv_s = 0;
      if (EPI == EPI_DY && n0 + c_lo * 32 >= p.N) dy_ready = false;   // (no columns for this warp: nothing consumes the prefetch)
      else if (EPI == EPI_DY) {
        if (!dy_ready) {   // first item of this CTA: nothing was prefetched
          aux_fetch(rowbase, min(n0 + c_lo * 32, p.N - 32));
          rv_fetch(rowbase, n0 + c_lo * 32);
        }
        // row vector slice -> scratch [2 images][ncols_w] (after the previous item's readers are done with it)
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 2; ++j) {
          const int idx = lane + 32 * j;
          if (idx < 2 * (ncols_w >> 2))
            asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(stg0 + 8192u + (uint32_t)idx * 16u), "f"(rvp[j].x),
                         "f"(rvp[j].y), "f"(rvp[j].z), "f"(rvp[j].w) : "memory");
        }
        __syncwarp();
        const int b0 = min(rowbase, p.M - 1) / p.S;
        rv_s = stg0 + 8192u + (uint32_t)((rr / p.S == b0) ? 0 : ncols_w * 4);
      }
#pragma unroll 1
      for (int c = c_lo; c < c_hi; ++c) {
        const int nc = n0 + c * 32;
        if (nc >= p.N) break;
        float v[32];
        tmem_ld32(taddr + (uint32_t)(c * 32), v);
        if (c == c_hi - 1 || nc + 32 >= p.N) {   // last read of this accumulator: hand it back to the MMA warp
          asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
          __syncwarp();
          if (lane == 0) { if (CG2) mbar_arrive_leader(&tempty_bar[ab]); else mbar_arrive(&tempty_bar[ab]); }
          released = true;
        }
        uint32_t w0[32];     // fp32 outputs: 32 words; bf16 outputs: w0[0..15] = hi pairs, w0[16..31] = lo pairs
        if (EPI == EPI_PLAIN || EPI == EPI_RED) {
#pragma unroll
          for (int k = 0; k < 32; ++k) w0[k] = __float_as_uint(v[k] * p.alpha);
        } else if (EPI == EPI_TANH) {
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + nc) + k4);
            v[4 * k4] += b4.x; v[4 * k4 + 1] += b4.y; v[4 * k4 + 2] += b4.z; v[4 * k4 + 3] += b4.w;
          }
          if (p.fast_tanh) {
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = tanh_hw(v[k]);
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = tanh_acc(v[k]);
          }
          if (p.of16) {
#pragma unroll
            for (int k = 0; k < 16; ++k) w0[k] = pack_f16x2(v[2 * k], v[2 * k + 1]);
          } else if (p.out_lo) {
#pragma unroll
            for (int k = 0; k < 16; ++k) split_pair(v[2 * k], v[2 * k + 1], w0[k], w0[16 + k]);
          } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) w0[k] = pack_bf16x2(v[2 * k], v[2 * k + 1]);
          }
        } else if (EPI == EPI_ATT) {
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + nc) + k4);
            const float4 q4 = __ldg(reinterpret_cast<const float4*>(rv + nc) + k4);
            v[4 * k4] += b4.x + q4.x; v[4 * k4 + 1] += b4.y + q4.y; v[4 * k4 + 2] += b4.z + q4.z; v[4 * k4 + 3] += b4.w + q4.w;
          }
          if (p.fast_tanh) {
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = tanh_hw(v[k]);
          } else {
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = tanh_acc(v[k]);
          }
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            const float4 c4 = __ldg(reinterpret_cast<const float4*>(p.colw + nc) + k4);
            rowacc = fmaf(c4.x, v[4 * k4], rowacc); rowacc = fmaf(c4.y, v[4 * k4 + 1], rowacc);
            rowacc = fmaf(c4.z, v[4 * k4 + 2], rowacc); rowacc = fmaf(c4.w, v[4 * k4 + 3], rowacc);
          }
#pragma unroll
          for (int k = 0; k < 32; ++k) w0[k] = __float_as_uint(v[k]);
        } else if (EPI == EPI_LSTM) {
          const int u0 = (nc >> 5) << 3;   // first of the 8 hidden units of this chunk
#pragma unroll
          for (int k4 = 0; k4 < 8; ++k4) {
            float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
            if (p.bias) { const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + nc) + k4); t.x += b4.x; t.y += b4.y; t.z += b4.z; t.w += b4.w; }
            if (p.addend) { const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.addend + (long long)rr * p.ldadd + nc) + k4); t.x += b4.x; t.y += b4.y; t.z += b4.z; t.w += b4.w; }
            v[4 * k4] += t.x; v[4 * k4 + 1] += t.y; v[4 * k4 + 2] += t.z; v[4 * k4 + 3] += t.w;
          }
          float cp[8];
          if (p.c_prev) {
            const float4 c0 = __ldg(reinterpret_cast<const float4*>(p.c_prev + (long long)rr * p.ldcp + u0));
            const float4 c1 = __ldg(reinterpret_cast<const float4*>(p.c_prev + (long long)rr * p.ldcp + u0) + 1);
            cp[0] = c0.x; cp[1] = c0.y; cp[2] = c0.z; cp[3] = c0.w; cp[4] = c1.x; cp[5] = c1.y; cp[6] = c1.z; cp[7] = c1.w;
          } else {
#pragma unroll
            for (int u = 0; u < 8; ++u) cp[u] = 0.0f;
          }
          float gi[8], gf[8], go[8], gg[8], tc[8], cn[8], hn[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) {
            gi[u] = rcp_ftz(1.0f + exp2f_ftz(-1.4426950408889634f * v[u]));
            gf[u] = rcp_ftz(1.0f + exp2f_ftz(-1.4426950408889634f * v[8 + u]));
            go[u] = rcp_ftz(1.0f + exp2f_ftz(-1.4426950408889634f * v[16 + u]));
            gg[u] = tanh_acc(v[24 + u]);
            cn[u] = fmaf(gf[u], cp[u], gi[u] * gg[u]);
            tc[u] = tanh_acc(cn[u]);
            hn[u] = go[u] * tc[u];
          }
          if (r_ok) {
            float4* d;
            d = reinterpret_cast<float4*>(p.c_out + (long long)r * p.ldc + u0);
            d[0] = make_float4(cn[0], cn[1], cn[2], cn[3]); d[1] = make_float4(cn[4], cn[5], cn[6], cn[7]);
            d = reinterpret_cast<float4*>(p.h_out + (long long)r * p.ldh + u0);
            d[0] = make_float4(hn[0], hn[1], hn[2], hn[3]); d[1] = make_float4(hn[4], hn[5], hn[6], hn[7]);
            if (p.lsaved) {
              const long long H_ = p.N >> 2;
              float* sbase = p.lsaved + (long long)r * H_ + u0;
              d = reinterpret_cast<float4*>(sbase);
              d[0] = make_float4(gi[0], gi[1], gi[2], gi[3]); d[1] = make_float4(gi[4], gi[5], gi[6], gi[7]);
              d = reinterpret_cast<float4*>(sbase + p.plane);
              d[0] = make_float4(gf[0], gf[1], gf[2], gf[3]); d[1] = make_float4(gf[4], gf[5], gf[6], gf[7]);
              d = reinterpret_cast<float4*>(sbase + 2 * p.plane);
              d[0] = make_float4(go[0], go[1], go[2], go[3]); d[1] = make_float4(go[4], go[5], go[6], go[7]);
              d = reinterpret_cast<float4*>(sbase + 3 * p.plane);
              d[0] = make_float4(gg[0], gg[1], gg[2], gg[3]); d[1] = make_float4(gg[4], gg[5], gg[6], gg[7]);
              d = reinterpret_cast<float4*>(sbase + 4 * p.plane);
              d[0] = make_float4(tc[0], tc[1], tc[2], tc[3]); d[1] = make_float4(tc[4], tc[5], tc[6], tc[7]);
            }
            if (p.hpk_hi) {
              uint32_t hh[4], hl[4];
#pragma unroll
              for (int u = 0; u < 4; ++u) split_pair(hn[2 * u], hn[2 * u + 1], hh[u], hl[u]);
              *reinterpret_cast<uint4*>(p.hpk_hi + (long long)r * p.ldhp + u0) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
              if (p.hpk_lo) *reinterpret_cast<uint4*>(p.hpk_lo + (long long)r * p.ldhp + u0) = make_uint4(hl[0], hl[1], hl[2], hl[3]);
            }
          }
          continue;
        } else if (EPI == EPI_LINEAR) {
          // nn.Linear epilogue: act(alpha*acc + bias + bias2 + addend + addend2), fp32 and/or packed bf16 (hi, lo) out
          if (!p.vec) {   // ragged or misaligned: guarded scalar loads
#pragma unroll
            for (int k = 0; k < 32; ++k) {
              const int n = nc + k;
              float t = 0.0f;
              if (n < p.N) {
                if (p.bias) t += __ldg(p.bias + n);
                if (p.bias2) t += __ldg(p.bias2 + n);
                if (p.addend) t += __ldg(p.addend + (long long)rr * p.ldadd + n);
                if (p.addend2 && p.act != 3) t += __ldg(p.addend2 + (long long)rr * p.ldadd + n);
              }
              v[k] = fmaf(v[k], p.alpha, t);
              if (p.act == 3 && n < p.N) {   // tanh backward: times (1 - y^2), y = addend2
                const float y = __ldg(p.addend2 + (long long)rr * p.ldadd + n);
                v[k] *= fmaf(-y, y, 1.0f);
              }
            }
          } else {
#pragma unroll
            for (int k4 = 0; k4 < 8; ++k4) {
              float4 t = make_float4(0.f, 0.f, 0.f, 0.f);
              if (p.bias) { const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias + nc) + k4); t.x += b4.x; t.y += b4.y; t.z += b4.z; t.w += b4.w; }
              if (p.bias2) { const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.bias2 + nc) + k4); t.x += b4.x; t.y += b4.y; t.z += b4.z; t.w += b4.w; }
              if (p.addend) { const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.addend + (long long)rr * p.ldadd + nc) + k4); t.x += b4.x; t.y += b4.y; t.z += b4.z; t.w += b4.w; }
              if (p.addend2 && p.act != 3) { const float4 b4 = __ldg(reinterpret_cast<const float4*>(p.addend2 + (long long)rr * p.ldadd + nc) + k4); t.x += b4.x; t.y += b4.y; t.z += b4.z; t.w += b4.w; }
              v[4 * k4] = fmaf(v[4 * k4], p.alpha, t.x); v[4 * k4 + 1] = fmaf(v[4 * k4 + 1], p.alpha, t.y);
              v[4 * k4 + 2] = fmaf(v[4 * k4 + 2], p.alpha, t.z); v[4 * k4 + 3] = fmaf(v[4 * k4 + 3], p.alpha, t.w);
              if (p.act == 3) {   // tanh backward fused behind the product: times (1 - y^2), y = addend2 (the tanh's output)
                const float4 y4 = __ldg(reinterpret_cast<const float4*>(p.addend2 + (long long)rr * p.ldadd + nc) + k4);
                v[4 * k4] *= fmaf(-y4.x, y4.x, 1.0f); v[4 * k4 + 1] *= fmaf(-y4.y, y4.y, 1.0f);
                v[4 * k4 + 2] *= fmaf(-y4.z, y4.z, 1.0f); v[4 * k4 + 3] *= fmaf(-y4.w, y4.w, 1.0f);
              }
            }
          }
          if (p.act == 1) {
            if (p.fast_tanh) {
#pragma unroll
              for (int k = 0; k < 32; ++k) v[k] = tanh_hw(v[k]);
            } else {
#pragma unroll
              for (int k = 0; k < 32; ++k) v[k] = tanh_acc(v[k]);
            }
          } else if (p.act == 2) {
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = rcp_ftz(1.0f + exp2f_ftz(-1.4426950408889634f * v[k]));
          }
#pragma unroll
          for (int k = 0; k < 32; ++k) w0[k] = __float_as_uint(v[k]);
          if (lane == 0) bulk_wait_read0();
          __syncwarp();
          if (p.out_f) stage_row128(stg0, lane, w0);
          if (p.out_hi) {
            uint32_t wh[32];
            if (p.out_lo) {
#pragma unroll
              for (int k = 0; k < 16; ++k) split_pair(v[2 * k], v[2 * k + 1], wh[k], wh[16 + k]);
            } else {
#pragma unroll
              for (int k = 0; k < 16; ++k) wh[k] = pack_bf16x2(v[2 * k], v[2 * k + 1]);
            }
            stage_row64(stg0 + 4096u, lane, wh);
            if (p.out_lo) stage_row64(stg0 + 6144u, lane, wh + 16);
          }
          fence_async_smem();
          __syncwarp();
          if (lane == 0 && rowbase < p.M) {
            if (p.out_f) tma_store_2d(&p.mapO[0], stg_p, nc, rowbase);
            if (p.out_hi) tma_store_2d(&p.mapO[1], stg_p + 4096, nc, rowbase);
            if (p.out_hi && p.out_lo) tma_store_2d(&p.mapO[2], stg_p + 6144, nc, rowbase);
            bulk_commit();
          }
          continue;
        } else {   // EPI_DY
          const uint32_t ax = stg0 + 4096u;
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            const int r_ = arow + 8 * i;
            const uint32_t a = ax + (uint32_t)r_ * 64u + (uint32_t)((aseg ^ ((r_ >> 1) & 3)) << 4);
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a), "r"(ah[i].x), "r"(ah[i].y), "r"(ah[i].z), "r"(ah[i].w) : "memory");
            asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(a + 2048u), "r"(al[i].x), "r"(al[i].y), "r"(al[i].z), "r"(al[i].w) : "memory");
          }
          __syncwarp();
          // the registers are free again: next chunk's activation (or the next item's first chunk and row-vector slice)
          // is in flight during the math below
          if (c + 1 < c_hi && nc + 32 < p.N) {
            aux_fetch(rowbase, nc + 32);
          } else {
            const int item2 = item + nworkers;
            dy_ready = item2 < items;
            if (dy_ready) {
              const int tile2 = fast_div(item2, p.ksplit, p.ksplit_magic);
              const int tm2 = fast_div(tile2, p.tiles_n, p.tiles_n_magic);
              const int tn2 = tile2 - tm2 * p.tiles_n;
              const int rb2 = (CG2 ? tm2 * 2 * RT_BM + (int)cta_rank * RT_BM : tm2 * RT_BM) + q * 32;
              aux_fetch(rb2, min(tn2 * p.BN + c_lo * 32, p.N - 32));
              rv_fetch(rb2, tn2 * p.BN + c_lo * 32);
              rs_next = p.rowscale[min(rb2 + lane, p.M - 1)];
            }
          }
#pragma unroll
          for (int k8 = 0; k8 < 4; ++k8) {   // 8 columns per pass: own row's activation and the row-vector slice from smem
            uint4 h, l;
            float4 d0, d1;
            const uint32_t ya = ax + (uint32_t)lane * 64u + (uint32_t)((k8 ^ ((lane >> 1) & 3)) << 4);
            const uint32_t da = rv_s + (uint32_t)((c - c_lo) * 128 + k8 * 32);
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(h.x), "=r"(h.y), "=r"(h.z), "=r"(h.w) : "r"(ya) : "memory");
            asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(l.x), "=r"(l.y), "=r"(l.z), "=r"(l.w) : "r"(ya + 2048u) : "memory");
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(d0.x), "=f"(d0.y), "=f"(d0.z), "=f"(d0.w) : "r"(da) : "memory");
            asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(d1.x), "=f"(d1.y), "=f"(d1.z), "=f"(d1.w) : "r"(da + 16u) : "memory");
            const uint32_t hh[4] = {h.x, h.y, h.z, h.w}, ll[4] = {l.x, l.y, l.z, l.w};
            const float dd[8] = {d0.x, d0.y, d0.z, d0.w, d1.x, d1.y, d1.z, d1.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              const float y0 = p.af16 ? hf_lo(hh[j]) : bf_lo(hh[j]) + bf_lo(ll[j]);
              const float y1 = p.af16 ? hf_hi(hh[j]) : bf_hi(hh[j]) + bf_hi(ll[j]);
              const int k = 8 * k8 + 2 * j;
              // (the accumulator carries the output's power-of-two scale through its A operand dZ; rs = p[r] * gscale)
              v[k] = fmaf(dd[2 * j], rs, v[k]) * fmaf(-y0, y0, 1.0f);
              v[k + 1] = fmaf(dd[2 * j + 1], rs, v[k + 1]) * fmaf(-y1, y1, 1.0f);
            }
          }
          __syncwarp();   // every lane has read its row before the next chunk overwrites the scratch
          if (!r_ok) {    // rows past M (a ragged last tile only) contribute nothing to the column sums
#pragma unroll
            for (int k = 0; k < 32; ++k) v[k] = 0.0f;
          }
          if (p.of16) {
#pragma unroll
            for (int k = 0; k < 16; ++k) w0[k] = pack_f16x2(v[2 * k], v[2 * k + 1]);
          } else if (p.out_lo) {
#pragma unroll
            for (int k = 0; k < 16; ++k) split_pair(v[2 * k], v[2 * k + 1], w0[k], w0[16 + k]);
          } else {
#pragma unroll
            for (int k = 0; k < 16; ++k) w0[k] = pack_bf16x2(v[2 * k], v[2 * k + 1]);
          }
          if (p.colsum) {
            const float s0 = warp_transpose_sum32(v, lane);
            if (nc + lane < p.N) atomicAdd(p.colsum + nc + lane, s0 * p.alpha);   // (alpha = 1 / gscale)
          }
        }
        // the staging buffer is free once the previous chunk's bulk stores have read it
        if (lane == 0) bulk_wait_read0();
        __syncwarp();
        const bool f32_out = (EPI == EPI_PLAIN || EPI == EPI_RED || EPI == EPI_ATT);
        if (f32_out) {
          stage_row128(stg0, lane, w0);
        } else {
          stage_row64(stg0, lane, w0);
          if (p.out_lo) stage_row64(stg1, lane, w0 + 16);
        }
        fence_async_smem();
        __syncwarp();
        if (lane == 0 && rowbase < p.M) {
          if (EPI == EPI_RED) {
            tma_reduce_add_2d(&p.mapO[0], stg_p, nc, rowbase);
          } else if (f32_out) {
            tma_store_2d(&p.mapO[0], stg_p, nc, rowbase);
          } else {
            tma_store_2d(&p.mapO[0], stg_p, nc, rowbase);
            if (p.out_lo) tma_store_2d(&p.mapO[1], stg_p + 2048, nc, rowbase);
          }
          bulk_commit();
        }
      }
      if (!released) {   // this half had no columns inside N: still hand the accumulator back
        asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
        __syncwarp();
        if (lane == 0) { if (CG2) mbar_arrive_leader(&tempty_bar[ab]); else mbar_arrive(&tempty_bar[ab]); }
      }
      if (EPI == EPI_ATT && r_ok && p.rowout && n0 + c_lo * 32 < p.N) atomicAdd(p.rowout + r, rowacc);
    }
    if (warp == 2 && lane == 0) RT_STAMP(7);   // epilogue math + stores issued
    if (lane == 0) bulk_wait0();
    if (warp == 2 && lane == 0) RT_STAMP(8);   // bulk stores drained
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (threadIdx.x == 0) RT_STAMP(9);
  if (dbg && threadIdx.x == 0) {
    unsigned long long gt1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt1));
    dbg[15] = gt1 - gt0;
  }
  if (CG2) cluster_sync_all();   // no CTA of the pair leaves while the other may still signal its barriers / read its smem
  if (warp == 1) {
    if (CG2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

}  // namespace

// ================================================================== host side
#include "rau_rows.cuh"

namespace {

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
EncodeFn g_encode = nullptr;

int get_encode() {
  if (g_encode) return RAU_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qr;
  cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qr);
  if (e != cudaSuccess || fn == nullptr) {
    rau_set_error("cuTensorMapEncodeTiled is not available: %s", cudaGetErrorString(e));
    return RAU_ECUDA;
  }
  g_encode = (EncodeFn)fn;
  return RAU_OK;
}

// 2-D map over a row-major array: dim0 (contiguous) x dim1 rows of pitch ld elements
int encode_2d(CUtensorMap* m, CUtensorMapDataType dt, int esize, const void* base, uint64_t dim0, uint64_t dim1, uint64_t ld,
              uint32_t box0, uint32_t box1, CUtensorMapSwizzle sw) {
  cuuint64_t dims[2] = {dim0, dim1};
  cuuint64_t strides[1] = {ld * (uint64_t)esize};
  cuuint32_t box[2] = {box0, box1};
  cuuint32_t es[2] = {1, 1};
  CUresult r = g_encode(m, dt, 2, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, sw,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    rau_set_error("cuTensorMapEncodeTiled failed (%d): dims=%llu,%llu ld=%llu box=%u,%u esize=%d", (int)r,
                  (unsigned long long)dim0, (unsigned long long)dim1, (unsigned long long)ld, box0, box1, esize);
    return RAU_ECUDA;
  }
  return RAU_OK;
}

// operand tiles: 2-D over one bf16 array, or (bf16x3) 3-D over the hi and lo arrays as two planes `plane` bytes apart
int encode_operand(CUtensorMap* m, const bf16* hi, const bf16* lo, int* swap, int mn, int rows, int K, int64_t ld, int BK,
                   int box_rows) {
  const uint64_t d0 = mn ? (uint64_t)rows : (uint64_t)K, d1 = mn ? (uint64_t)K : (uint64_t)rows;
  const uint32_t b0 = mn ? 64u : (uint32_t)BK, b1 = mn ? (uint32_t)BK : (uint32_t)box_rows;
  const CUtensorMapSwizzle sw = (mn || BK == 64) ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B;
  *swap = 0;
  if (lo == nullptr) return encode_2d(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, hi, d0, d1, (uint64_t)ld, b0, b1, sw);
  const bf16* base = hi;
  if (lo < hi) { base = lo; *swap = 1; }
  const uint64_t plane = (uint64_t)((const char*)(*swap ? hi : lo) - (const char*)base);
  if (plane % 16 != 0 || plane >= (1ull << 40)) {
    rau_set_error("rows_gemm: the hi and lo arrays of an operand are %llu bytes apart (need a multiple of 16 below 2^40)",
                  (unsigned long long)plane);
    return RAU_EINVAL;
  }
  cuuint64_t dims[3] = {d0, d1, 2};
  cuuint64_t strides[2] = {(cuuint64_t)ld * 2, plane};
  cuuint32_t box[3] = {b0, b1, 2};
  cuuint32_t es[3] = {1, 1, 1};
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void*)base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        sw, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    rau_set_error("cuTensorMapEncodeTiled (3-D operand) failed (%d): dims=%llu,%llu ld=%lld plane=%llu box=%u,%u", (int)r,
                  (unsigned long long)d0, (unsigned long long)d1, (long long)ld, (unsigned long long)plane, b0, b1);
    return RAU_ECUDA;
  }
  return RAU_OK;
}

bool g_attr_done[8] = {false, false, false, false, false, false, false, false};

template <int EPI, int X3, int NSTEPS, int CG2, int EW = 8>
int launch_rows_v(rau_ctx* ctx, const RtParams& p, int grid, int smem_bytes) {
  static bool attr_done = false;
  if (!attr_done) {
    RAU_CHECK_CUDA(cudaFuncSetAttribute(rows_gemm_kernel<EPI, X3, NSTEPS, CG2, EW>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                        RT_SMEM_BUDGET + 1024));
    attr_done = true;
  }
  if (CG2) {   // clusters of two CTAs (one TPC): the cta_group::2 pair
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(32 * (EW + 3));
    cfg.dynamicSmemBytes = smem_bytes;
    cfg.stream = ctx->stream;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = rau_pdl_enabled() ? 1 : 0;
    cfg.attrs = at;
    cfg.numAttrs = 2;
    (void)cudaLaunchKernelEx(&cfg, rows_gemm_kernel<EPI, X3, NSTEPS, CG2, EW>, p);
  } else {
    RAU_LAUNCH_PDL(ctx->stream, (rows_gemm_kernel<EPI, X3, NSTEPS, CG2, EW>), grid, 32 * (EW + 3), smem_bytes, p);
  }
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}
// variants: bf16x3 with 32-wide k-blocks (256-column tiles) or 64-wide (narrow tiles); single-pass bf16 with 64-wide;
// the CTA-pair form exists for the big K-major products with fused epilogues
template <int EPI>
int launch_rows(rau_ctx* ctx, const RtParams& p, int grid, int smem_bytes) {
  constexpr bool pairable = EPI == EPI_PLAIN || EPI == EPI_TANH || EPI == EPI_ATT || EPI == EPI_DY || EPI == EPI_RED || EPI == EPI_LINEAR;
  if (EPI == EPI_TANH && p.ew == 16 && p.cg2 && !p.x3) return launch_rows_v<EPI_TANH, 0, 4, 1, 16>(ctx, p, grid, smem_bytes);
  if (pairable && p.cg2) {
    if (p.x3) return launch_rows_v<EPI, 1, 2, pairable ? 1 : 0>(ctx, p, grid, smem_bytes);
    return launch_rows_v<EPI, 0, 4, pairable ? 1 : 0>(ctx, p, grid, smem_bytes);
  }
  if (p.x3) return p.BK == 32 ? launch_rows_v<EPI, 1, 2, 0>(ctx, p, grid, smem_bytes) : launch_rows_v<EPI, 1, 4, 0>(ctx, p, grid, smem_bytes);
  return launch_rows_v<EPI, 0, 4, 0>(ctx, p, grid, smem_bytes);
}

__global__ void unpack_hilo_kernel(const bf16* __restrict__ hi, const bf16* __restrict__ lo, int64_t n, float* __restrict__ out,
                                   int f16) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = f16 ? __half2float(reinterpret_cast<const __half*>(hi)[i])
                 : __bfloat162float(hi[i]) + (lo ? __bfloat162float(lo[i]) : 0.0f);
}

__global__ void pack_hilo_kernel(const float* __restrict__ in, int64_t n4, bf16* __restrict__ hi, bf16* __restrict__ lo, int f16) {
  RAU_PDL_ENTRY();
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n4; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 x = reinterpret_cast<const float4*>(in)[i];
    uint32_t h0, l0, h1, l1;
    if (f16) {
      reinterpret_cast<uint2*>(hi)[i] = make_uint2(pack_f16x2(x.x, x.y), pack_f16x2(x.z, x.w));
      continue;
    }
    split_pair(x.x, x.y, h0, l0);
    split_pair(x.z, x.w, h1, l1);
    reinterpret_cast<uint2*>(hi)[i] = make_uint2(h0, h1);
    if (lo) reinterpret_cast<uint2*>(lo)[i] = make_uint2(l0, l1);
  }
}

// X [B, C, S] fp32 -> dropped-out features in rows layout [B*S, C] as bf16 (hi, lo); keep bit index = (b*C + c)*S + s
// gen != 0: the keep bits are drawn here (Philox4x32-10, the same counter -> bit mapping as mask_gen_kernel: element e
// takes lane e & 3 of the draw with counter e >> 2), so the training step needs no separate mask pass over [B, C, S]
__global__ void __launch_bounds__(256) xprep_rows_kernel(const float* __restrict__ X, const uint32_t* __restrict__ bits, float scale,
                                                         int C, int S, bf16* __restrict__ hi, bf16* __restrict__ lo, int gen,
                                                         uint32_t thresh, uint2 key, uint32_t stream_lo, uint32_t stream_hi,
                                                         const StepState* __restrict__ ss, int f16, int hop, int nHop) {
  RAU_PDL_ENTRY();
  extern __shared__ float sT[];   // [S][66]
  const int b = blockIdx.y, c0 = blockIdx.x * 64;
  const int S4 = S >> 2;
  // (hop >= 0: stream_* is the hop's own stream id = base ^ hop; the shared draw of p = 1/2 uses the base)
  const bool shared = gen && hop >= 0 && rau_xmask_shared(thresh, nHop);
  if (shared) stream_lo = stream_lo ^ (uint32_t)hop ^ RAU_XMASK_SHARED_TAG;
  if (gen && ss) {   // graph replay: the step part of the stream id lives on the device
    const unsigned long long sid = (((unsigned long long)stream_hi << 32) | stream_lo) ^ (ss->step << 24);
    stream_lo = (uint32_t)sid;
    stream_hi = (uint32_t)(sid >> 32);
  }
  for (int i = threadIdx.x; i < 64 * S4; i += 256) {
    const int c = i / S4, s4 = i - c * S4;
    const int64_t e = ((int64_t)b * C + c0 + c) * S + 4 * s4;
    float4 t = *reinterpret_cast<const float4*>(X + e);
    if (gen) {
      // inline-drawn keep bits: one Philox call serves a channel PAIR (even, odd) x 4 grid cells -- word k of the call
      // decides cell 4*s4+k, its low half for the even channel, its high half for the odd one (16-bit threshold)
      const int cc = c0 + c;
      const uint64_t ctr = (uint64_t)(((int64_t)b * C + (cc & ~1)) * S + 4 * s4) >> 2;
      const uint4 r = philox4x32(make_uint4((uint32_t)ctr, (uint32_t)(ctr >> 32), stream_lo, stream_hi), key);
      const int sh = (cc & 1) * 16;
      if (shared) {   // bit `hop` of the element's 16-bit lane, clear = keep
        const int bit = sh + hop;
        t.x = ((r.x >> bit) & 1u) ? 0.0f : t.x * scale;
        t.y = ((r.y >> bit) & 1u) ? 0.0f : t.y * scale;
        t.z = ((r.z >> bit) & 1u) ? 0.0f : t.z * scale;
        t.w = ((r.w >> bit) & 1u) ? 0.0f : t.w * scale;
      } else {
        t.x = ((r.x >> sh) & 0xffffu) < thresh ? t.x * scale : 0.0f;
        t.y = ((r.y >> sh) & 0xffffu) < thresh ? t.y * scale : 0.0f;
        t.z = ((r.z >> sh) & 0xffffu) < thresh ? t.z * scale : 0.0f;
        t.w = ((r.w >> sh) & 0xffffu) < thresh ? t.w * scale : 0.0f;
      }
    } else if (bits) {
      const uint32_t w = bits[e >> 5] >> (e & 31);
      t.x = (w & 1u) ? t.x * scale : 0.0f;
      t.y = (w & 2u) ? t.y * scale : 0.0f;
      t.z = (w & 4u) ? t.z * scale : 0.0f;
      t.w = (w & 8u) ? t.w * scale : 0.0f;
    }
    float* d = sT + (4 * s4) * 66 + c;
    d[0] = t.x; d[66] = t.y; d[132] = t.z; d[198] = t.w;
  }
  __syncthreads();
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int s = warp; s < S; s += 8) {
    const float2 v = *reinterpret_cast<const float2*>(sT + s * 66 + 2 * lane);
    uint32_t h, l;
    const int64_t o = (((int64_t)b * S + s) * C + c0) >> 1;
    if (f16) { reinterpret_cast<uint32_t*>(hi)[o + lane] = pack_f16x2(v.x, v.y); continue; }
    split_pair(v.x, v.y, h, l);
    reinterpret_cast<uint32_t*>(hi)[o + lane] = h;
    if (lo) reinterpret_cast<uint32_t*>(lo)[o + lane] = l;
  }
}

// dXr [B*S, C] fp32 -> dX [B, C, S] = dXr^T * keep * scale  (backward of the feature dropout, F:239)
// The same pack for ALL hops of the training step in one launch: every hop drops out the SAME features with its own
// Philox stream (stream id ^ hop), so the fp32 tile is read and transposed once and written nHop times (hop h's arrays are
// hop_stride elements after hop 0's).  Identical bits to nHop launches of xprep_rows_kernel.
// Persistent: one 1024-thread CTA per SM it may use (four 256-thread groups, each with its own [S][66] slab and named
// barrier, walk the (image, 64-channel) tiles).  The big CTAs keep the launch on `gridDim.x` SMs, so the chain's tcgen05
// launches still find free SMs next to it (a grid of small CTAs spread over every SM and filled their shared memory).
// X16 != NULL: the features arrive as fp16 [B, C, S] (rau_batch.feats_f16, the fp16 feed) and are widened on load.
__global__ void __launch_bounds__(1024, 1) xprep_rows_hops_kernel(const float* __restrict__ X, const __half* __restrict__ X16,
                                                                  float scale, int C, int S, int B, int nHop,
                                                                  bf16* __restrict__ hi, bf16* __restrict__ lo, long long hop_stride,
                                                                  uint32_t thresh, uint2 key, uint32_t stream_lo, uint32_t stream_hi,
                                                                  const StepState* __restrict__ ss, int f16) {
  RAU_PDL_ENTRY();
  extern __shared__ float sT_all[];   // 4 x [S][66]
  const int sub = threadIdx.x >> 8, tid = threadIdx.x & 255;
  float* sT = sT_all + (size_t)sub * S * 66;
  const int S4 = S >> 2, ctiles = C / 64, ntiles = ctiles * B;
  if (ss) {
    const unsigned long long sid = (((unsigned long long)stream_hi << 32) | stream_lo) ^ (ss->step << 24);
    stream_lo = (uint32_t)sid;
    stream_hi = (uint32_t)(sid >> 32);
  }
  const int warp = tid >> 5, lane = tid & 31;
  for (int tile = blockIdx.x * 4 + sub; tile < ntiles; tile += gridDim.x * 4) {
    const int b = tile / ctiles, c0 = (tile - b * ctiles) * 64;
    // the tile's 64 x S floats are contiguous in X: batches of 7 independent 16-byte loads per thread, then the
    // transposing stores (a load -> store loop exposes one memory latency per iteration)
    const float4* src = reinterpret_cast<const float4*>(X + ((int64_t)b * C + c0) * S);
    const uint2* src16 = reinterpret_cast<const uint2*>(X16 + ((int64_t)b * C + c0) * S);
    const int n4 = 64 * S4;
    for (int i0 = tid; i0 < n4; i0 += 256 * 7) {
      float4 t[7];
      if (X16) {
        uint2 u[7];
#pragma unroll
        for (int k = 0; k < 7; ++k) {
          const int i = i0 + 256 * k;
          if (i < n4) u[k] = __ldg(src16 + i);
        }
#pragma unroll
        for (int k = 0; k < 7; ++k) t[k] = make_float4(hf_lo(u[k].x), hf_hi(u[k].x), hf_lo(u[k].y), hf_hi(u[k].y));
      } else {
#pragma unroll
        for (int k = 0; k < 7; ++k) {
          const int i = i0 + 256 * k;
          if (i < n4) t[k] = __ldg(src + i);
        }
      }
#pragma unroll
      for (int k = 0; k < 7; ++k) {
        const int i = i0 + 256 * k;
        if (i < n4) {
          const int c = i / S4, s4 = i - c * S4;
          float* d = sT + (4 * s4) * 66 + c;
          d[0] = t[k].x; d[66] = t[k].y; d[132] = t[k].z; d[198] = t[k].w;
        }
      }
    }
    asm volatile("bar.sync %0, 256;" ::"r"(1 + sub) : "memory");
    // a lane owns the channel pair 2*lane, 2*lane+1: one Philox call covers the pair x 4 consecutive grid cells
    const uint64_t e0 = ((uint64_t)b * C + c0 + 2 * lane) * (uint64_t)S;
    for (int s4 = warp; s4 < S4; s4 += 8) {
      float2 v[4];
#pragma unroll
      for (int k = 0; k < 4; ++k) v[k] = *reinterpret_cast<const float2*>(sT + (4 * s4 + k) * 66 + 2 * lane);
      const uint64_t ct0 = (e0 + 4 * s4) >> 2;
      if (rau_xmask_shared(thresh, nHop)) {
        // p = 1/2: ONE draw for all hops (bit h of a 16-bit lane decides hop h).  The scaled values are packed once; a hop's
        // stores are the packed words ANDed with a 0 / 0xffff mask per half: 4-5 instructions per channel pair and hop.
        const uint4 r0 = philox4x32(make_uint4((uint32_t)ct0, (uint32_t)(ct0 >> 32), stream_lo ^ RAU_XMASK_SHARED_TAG, stream_hi), key);
        const uint32_t keepw[4] = {~r0.x, ~r0.y, ~r0.z, ~r0.w};
        uint32_t wh[4], wl[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          if (f16) { wh[k] = pack_f16x2(v[k].x * scale, v[k].y * scale); wl[k] = 0u; }
          else split_pair(v[k].x * scale, v[k].y * scale, wh[k], wl[k]);
        }
        for (int h = 0; h < nHop; ++h) {
          uint32_t* ph = reinterpret_cast<uint32_t*>(hi + (long long)h * hop_stride);
          uint32_t* pl = (lo && !f16) ? reinterpret_cast<uint32_t*>(lo + (long long)h * hop_stride) : nullptr;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint32_t m = ((keepw[k] >> h) & 0x00010001u) * 0xffffu;
            const int64_t o = (((int64_t)b * S + 4 * s4 + k) * C + c0) >> 1;
            ph[o + lane] = wh[k] & m;
            if (pl) pl[o + lane] = wl[k] & m;
          }
        }
        continue;
      }
      for (int h = 0; h < nHop; ++h) {
        const uint4 r0 = philox4x32(make_uint4((uint32_t)ct0, (uint32_t)(ct0 >> 32), stream_lo ^ (uint32_t)h, stream_hi), key);
        const uint32_t q0[4] = {r0.x & 0xffffu, r0.y & 0xffffu, r0.z & 0xffffu, r0.w & 0xffffu};
        const uint32_t q1[4] = {r0.x >> 16, r0.y >> 16, r0.z >> 16, r0.w >> 16};
        uint32_t* ph = reinterpret_cast<uint32_t*>(hi + (long long)h * hop_stride);
        uint32_t* pl = lo ? reinterpret_cast<uint32_t*>(lo + (long long)h * hop_stride) : nullptr;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const float a = q0[k] < thresh ? v[k].x * scale : 0.0f, c = q1[k] < thresh ? v[k].y * scale : 0.0f;
          uint32_t hh, ll;
          const int64_t o = (((int64_t)b * S + 4 * s4 + k) * C + c0) >> 1;
          if (f16) { ph[o + lane] = pack_f16x2(a, c); continue; }
          split_pair(a, c, hh, ll);
          ph[o + lane] = hh;
          if (pl) pl[o + lane] = ll;
        }
      }
    }
    asm volatile("bar.sync %0, 256;" ::"r"(1 + sub) : "memory");   // the slab is free for the group's next tile
  }
}

__global__ void __launch_bounds__(256) unprep_rows_kernel(const float* __restrict__ dXr, const uint32_t* __restrict__ bits, float scale,
                                                          int C, int S, float* __restrict__ dX) {
  RAU_PDL_ENTRY();
  extern __shared__ float sT[];   // [S][66]
  const int b = blockIdx.y, c0 = blockIdx.x * 64;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int s = warp; s < S; s += 8) {
    const float2 v = *reinterpret_cast<const float2*>(dXr + ((int64_t)b * S + s) * C + c0 + 2 * lane);
    *reinterpret_cast<float2*>(sT + s * 66 + 2 * lane) = v;
  }
  __syncthreads();
  const int S4 = S >> 2;
  for (int i = threadIdx.x; i < 64 * S4; i += 256) {
    const int c = i / S4, s4 = i - c * S4;
    const int64_t e = ((int64_t)b * C + c0 + c) * S + 4 * s4;
    const float* d = sT + (4 * s4) * 66 + c;
    float4 t = make_float4(d[0], d[66], d[132], d[198]);
    if (bits) {
      const uint32_t w = bits[e >> 5] >> (e & 31);
      t.x = (w & 1u) ? t.x * scale : 0.0f;
      t.y = (w & 2u) ? t.y * scale : 0.0f;
      t.z = (w & 4u) ? t.z * scale : 0.0f;
      t.w = (w & 8u) ? t.w * scale : 0.0f;
    }
    *reinterpret_cast<float4*>(dX + e) = t;
  }
}

// attbycontent, the state-dependent half (F:244-252): logit[r] = sum_n ws[n] tanh(Z[r,n] + qadd[b(r),n]) where
// Z = I Wa^T was formed ahead of the recurrence (it does not depend on the state) and qadd = Wqa qf + bqa + ba.
// A warp per 4 rows, 16-byte loads.  Z is read once and only the R logits are written, but the pass is instruction- and
// latency-bound, not HBM-bound (ncu: 47 % issue slots at 37 % active warps, DRAM at 30 %), so the inner loop is cut to five
// instructions per element: tanh(x) = 1 - 2 / (2^(c x) + 1) with c = 2 log2(e) gives
//   logit = sum_n ws_n - 2 sum_n ws_n r_n,  r_n = 1 / (2^(c Z_n + c q_n) + 1)
// (c q and -2 ws are formed once per 4 rows; the four row sums are reduced together, 6 shuffles instead of 20).
template <int FAST>
__global__ void __launch_bounds__(256) attn_rows_score_kernel(int R, int S, int A, const float* __restrict__ Z,
                                                              const float* __restrict__ qadd, const float* __restrict__ ws,
                                                              float* __restrict__ logit) {
  RAU_PDL_ENTRY();
  const int lane = threadIdx.x & 31;
  const int a4 = A >> 2;
  constexpr float C2 = 2.8853900817779268f;   // 2 log2(e)
  // a warp walks batches of 4 rows, warps_total batches apart (the grid is sized to one resident wave)
  const int nb = (R + 3) >> 2, wstride = gridDim.x * 8;
  for (int batch = blockIdx.x * 8 + (threadIdx.x >> 5); batch < nb; batch += wstride) {
    const int rbase = batch * 4;
    const int b0 = rbase / S;
    const bool one_image = (min(rbase + 3, R - 1)) / S == b0;   // (warp-uniform)
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
    float wsum = 0.0f;
    for (int w = lane; w < a4; w += 32) {
      const float4 w4 = __ldg(reinterpret_cast<const float4*>(ws) + w);
      float4 z[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) z[i] = __ldg(reinterpret_cast<const float4*>(Z + (int64_t)min(rbase + i, R - 1) * A) + w);
      if (FAST) {
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 qv = __ldg(reinterpret_cast<const float4*>(qadd + (int64_t)(min(rbase + i, R - 1) / S) * A) + w);
          acc[i] = fmaf(w4.x, tanh_hw(z[i].x + qv.x), acc[i]);
          acc[i] = fmaf(w4.y, tanh_hw(z[i].y + qv.y), acc[i]);
          acc[i] = fmaf(w4.z, tanh_hw(z[i].z + qv.z), acc[i]);
          acc[i] = fmaf(w4.w, tanh_hw(z[i].w + qv.w), acc[i]);
        }
        continue;
      }
      wsum += (w4.x + w4.y) + (w4.z + w4.w);
      const float m0 = -2.0f * w4.x, m1 = -2.0f * w4.y, m2 = -2.0f * w4.z, m3 = -2.0f * w4.w;
      float4 qs = __ldg(reinterpret_cast<const float4*>(qadd + (int64_t)b0 * A) + w);
      qs.x *= C2; qs.y *= C2; qs.z *= C2; qs.w *= C2;
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        if (!one_image && i > 0) {   // the batch straddles two images: this row's own q
          qs = __ldg(reinterpret_cast<const float4*>(qadd + (int64_t)(min(rbase + i, R - 1) / S) * A) + w);
          qs.x *= C2; qs.y *= C2; qs.z *= C2; qs.w *= C2;
        }
        acc[i] = fmaf(m0, rcp_ftz(exp2f_ftz(fmaf(z[i].x, C2, qs.x)) + 1.0f), acc[i]);
        acc[i] = fmaf(m1, rcp_ftz(exp2f_ftz(fmaf(z[i].y, C2, qs.y)) + 1.0f), acc[i]);
        acc[i] = fmaf(m2, rcp_ftz(exp2f_ftz(fmaf(z[i].z, C2, qs.z)) + 1.0f), acc[i]);
        acc[i] = fmaf(m3, rcp_ftz(exp2f_ftz(fmaf(z[i].w, C2, qs.w)) + 1.0f), acc[i]);
      }
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] += wsum;
    // the four row sums reduced together: after the xor-16 and xor-8 exchanges lane l carries row 2*(l>>4&1) + (l>>3&1)
    const bool hi16 = (lane & 16) != 0, hi8 = (lane & 8) != 0;
    float v0 = hi16 ? acc[2] : acc[0], s0 = hi16 ? acc[0] : acc[2];
    float v1 = hi16 ? acc[3] : acc[1], s1 = hi16 ? acc[1] : acc[3];
    v0 += __shfl_xor_sync(0xffffffffu, s0, 16);
    v1 += __shfl_xor_sync(0xffffffffu, s1, 16);
    float u = hi8 ? v1 : v0;
    const float su = hi8 ? v0 : v1;
    u += __shfl_xor_sync(0xffffffffu, su, 8);
    u += __shfl_xor_sync(0xffffffffu, u, 4);
    u += __shfl_xor_sync(0xffffffffu, u, 2);
    u += __shfl_xor_sync(0xffffffffu, u, 1);
    const int row = rbase + (hi16 ? 2 : 0) + (hi8 ? 1 : 0);
    if ((lane & 7) == 0 && row < R) logit[row] = u;
  }
}

// attbymemory + attselect on the rows layout (F:285-290, F:254-263): p = softmax(logit + mem), a = sum_s p_s I[b*S+s, :]
// grid (B, M/256), 256 threads: a CTA owns 256 channels of one image; 32 threads cover a row slice with 16-byte loads,
// 8 row groups walk the image in parallel (HBM-bound: I hi+lo is read exactly once)
// SCORE != 0: `logit` is not read; the CTA first forms its image's content logits itself, ws . tanh(Z[r,:] + qadd[b,:])
// (the attn_rows_score_kernel pass fused in: one launch less on the recurrent chain; both channel halves of an image redo
// the 196 logits, the second read of Z comes out of L2).  1 = accurate tanh, 2 = MUFU.TANH.
template <int SCORE>
__global__ void __launch_bounds__(256) attn_rows_fwd_kernel(int S, int M, const float* __restrict__ logit,
                                                            const float* __restrict__ mem, const bf16* __restrict__ I_hi,
                                                            const bf16* __restrict__ I_lo, float* __restrict__ p_out,
                                                            float* __restrict__ a_out, bf16* __restrict__ p_hi,
                                                            bf16* __restrict__ p_lo, int ldp, int A, const float* __restrict__ Z,
                                                            const float* __restrict__ qadd, const float* __restrict__ ws, int f16) {
  RAU_PDL_ENTRY();
  __shared__ float p[256];
  __shared__ float part[8][256];
  __shared__ float red[32];
  const int b = blockIdx.x, tid = threadIdx.x, c0 = blockIdx.y * 256;
  const int64_t r0 = (int64_t)b * S;
  if (SCORE) {
    const int wq = tid >> 5, ln = tid & 31, a4 = A >> 2;
    for (int s0 = wq * 2; s0 < S; s0 += 16) {   // a warp takes two rows at a time (independent loads in flight)
      float acc0 = 0.0f, acc1 = 0.0f;
      const int s1 = min(s0 + 1, S - 1);
      for (int w = ln; w < a4; w += 32) {
        const float4 w4 = __ldg(reinterpret_cast<const float4*>(ws) + w);
        const float4 q4 = __ldg(reinterpret_cast<const float4*>(qadd + (int64_t)b * A) + w);
        const float4 z0 = __ldg(reinterpret_cast<const float4*>(Z + (r0 + s0) * A) + w);
        const float4 z1 = __ldg(reinterpret_cast<const float4*>(Z + (r0 + s1) * A) + w);
#define RAU_T(x) (SCORE == 2 ? tanh_hw(x) : tanh_acc(x))
        acc0 = fmaf(w4.x, RAU_T(z0.x + q4.x), acc0); acc0 = fmaf(w4.y, RAU_T(z0.y + q4.y), acc0);
        acc0 = fmaf(w4.z, RAU_T(z0.z + q4.z), acc0); acc0 = fmaf(w4.w, RAU_T(z0.w + q4.w), acc0);
        acc1 = fmaf(w4.x, RAU_T(z1.x + q4.x), acc1); acc1 = fmaf(w4.y, RAU_T(z1.y + q4.y), acc1);
        acc1 = fmaf(w4.z, RAU_T(z1.z + q4.z), acc1); acc1 = fmaf(w4.w, RAU_T(z1.w + q4.w), acc1);
#undef RAU_T
      }
      acc0 = warp_sum(acc0); acc1 = warp_sum(acc1);
      if (ln == 0) { p[s0] = acc0; if (s0 + 1 < S) p[s0 + 1] = acc1; }
    }
    __syncthreads();
  }
  const float lc = SCORE ? (tid < S ? p[tid] : 0.0f) : (tid < S ? logit[r0 + tid] : 0.0f);
  if (SCORE) __syncthreads();   // p[] is rewritten with the probabilities below
  const float l0 = tid < S ? lc + mem[r0 + tid] : -INFINITY;
  const float mx = block_max(l0, red);
  const float e0 = tid < S ? __expf(l0 - mx) : 0.0f;
  const float den = block_sum(e0, red);
  p[tid] = e0 / den;
  if (blockIdx.y == 0 && tid < S) p_out[r0 + tid] = e0 / den;
  if (blockIdx.y == 0 && p_hi && tid < ldp) {   // packed twin of p (pitch ldp >= S, zero padded) for the Wp product
    const float v = tid < S ? e0 / den : 0.0f;
    const bf16 h = __float2bfloat16_rn(v);
    p_hi[(int64_t)b * ldp + tid] = h;
    if (p_lo) p_lo[(int64_t)b * ldp + tid] = __float2bfloat16_rn(v - __bfloat162float(h));
  }
  __syncthreads();
  const int cg = tid & 31, rg = tid >> 5;
  float a[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) a[k] = 0.0f;
#pragma unroll 5
  for (int s = rg; s < S; s += 8) {
    const uint4 h = __ldg(reinterpret_cast<const uint4*>(I_hi + (r0 + s) * M + c0) + cg);
    uint4 l = make_uint4(0u, 0u, 0u, 0u);
    if (I_lo) l = __ldg(reinterpret_cast<const uint4*>(I_lo + (r0 + s) * M + c0) + cg);
    const float w = p[s];
    if (f16) {
      a[0] = fmaf(w, hf_lo(h.x), a[0]); a[1] = fmaf(w, hf_hi(h.x), a[1]); a[2] = fmaf(w, hf_lo(h.y), a[2]); a[3] = fmaf(w, hf_hi(h.y), a[3]);
      a[4] = fmaf(w, hf_lo(h.z), a[4]); a[5] = fmaf(w, hf_hi(h.z), a[5]); a[6] = fmaf(w, hf_lo(h.w), a[6]); a[7] = fmaf(w, hf_hi(h.w), a[7]);
      continue;
    }
    a[0] = fmaf(w, bf_lo(h.x) + bf_lo(l.x), a[0]); a[1] = fmaf(w, bf_hi(h.x) + bf_hi(l.x), a[1]);
    a[2] = fmaf(w, bf_lo(h.y) + bf_lo(l.y), a[2]); a[3] = fmaf(w, bf_hi(h.y) + bf_hi(l.y), a[3]);
    a[4] = fmaf(w, bf_lo(h.z) + bf_lo(l.z), a[4]); a[5] = fmaf(w, bf_hi(h.z) + bf_hi(l.z), a[5]);
    a[6] = fmaf(w, bf_lo(h.w) + bf_lo(l.w), a[6]); a[7] = fmaf(w, bf_hi(h.w) + bf_hi(l.w), a[7]);
  }
  float4* dst = reinterpret_cast<float4*>(&part[rg][8 * cg]);
  dst[0] = make_float4(a[0], a[1], a[2], a[3]);
  dst[1] = make_float4(a[4], a[5], a[6], a[7]);
  __syncthreads();
  float v = 0.0f;
#pragma unroll
  for (int k = 0; k < 8; ++k) v += part[k][tid];
  a_out[(int64_t)b * M + c0 + tid] = v;
}

// backward, part 1 (a warp per row of I, fully parallel over the B*S rows): dp[r] = dp_in[r] + sum_m da[b(r), m] I[r, m]
__global__ void __launch_bounds__(256) attn_rows_dp_kernel(int R, int S, int M, const bf16* __restrict__ I_hi,
                                                           const bf16* __restrict__ I_lo, const float* __restrict__ da,
                                                           const float* __restrict__ dp_in, float* __restrict__ dp, int f16) {
  RAU_PDL_ENTRY();
  const int lane = threadIdx.x & 31;
  const int r = blockIdx.x * 8 + (threadIdx.x >> 5);
  if (r >= R) return;
  const int b = r / S;
  const uint4* ih = reinterpret_cast<const uint4*>(I_hi + (int64_t)r * M);
  const uint4* il = I_lo ? reinterpret_cast<const uint4*>(I_lo + (int64_t)r * M) : nullptr;
  const float* d = da + (int64_t)b * M;
  const int w8 = M >> 3;
  float acc = 0.0f;
#pragma unroll 2
  for (int w = lane; w < w8; w += 32) {
    const uint4 h = __ldg(ih + w);
    uint4 l = make_uint4(0u, 0u, 0u, 0u);
    if (il) l = __ldg(il + w);
    const float4 d0 = __ldg(reinterpret_cast<const float4*>(d + 8 * w));
    const float4 d1 = __ldg(reinterpret_cast<const float4*>(d + 8 * w) + 1);
    if (f16) {
      acc = fmaf(d0.x, hf_lo(h.x), acc); acc = fmaf(d0.y, hf_hi(h.x), acc); acc = fmaf(d0.z, hf_lo(h.y), acc); acc = fmaf(d0.w, hf_hi(h.y), acc);
      acc = fmaf(d1.x, hf_lo(h.z), acc); acc = fmaf(d1.y, hf_hi(h.z), acc); acc = fmaf(d1.z, hf_lo(h.w), acc); acc = fmaf(d1.w, hf_hi(h.w), acc);
      continue;
    }
    acc = fmaf(d0.x, bf_lo(h.x) + bf_lo(l.x), acc); acc = fmaf(d0.y, bf_hi(h.x) + bf_hi(l.x), acc);
    acc = fmaf(d0.z, bf_lo(h.y) + bf_lo(l.y), acc); acc = fmaf(d0.w, bf_hi(h.y) + bf_hi(l.y), acc);
    acc = fmaf(d1.x, bf_lo(h.z) + bf_lo(l.z), acc); acc = fmaf(d1.y, bf_hi(h.z) + bf_hi(l.z), acc);
    acc = fmaf(d1.z, bf_lo(h.w) + bf_lo(l.w), acc); acc = fmaf(d1.w, bf_hi(h.w) + bf_hi(l.w), acc);
  }
  acc = warp_sum(acc);
  if (lane == 0) dp[r] = acc + (dp_in ? dp_in[r] : 0.0f);
}

// backward, part 2: grid (B, NSL) -- every CTA redoes the image's softmax backward (196 scalars), then emits dZ for its
// slice of the image's rows:  ds = p (dp - <p,dp>) ; dZ[r,n] = ws[n] ds[s] (1 - E[r,n]^2) -> bf16 (hi, lo) ;
// dqa[b,n] += sum_s dZ[r,n] ; gws_part[b,n] += sum_s ds[s] E[r,n]   (both zeroed by the caller, slices add atomically)
template <int RECOMP>   // 0: E holds tanh(.) ; 1: E holds Z = I Wa^T, e = tanh(Z + qadd[b]) ; 2: same with MUFU.TANH
__global__ void __launch_bounds__(256) attn_rows_dz_kernel(int S, int A, const float* __restrict__ E,
                                                           const float* __restrict__ qadd,
                                                           const float* __restrict__ ws, const float* __restrict__ p_in,
                                                           const float* __restrict__ dp, float* __restrict__ ds_out,
                                                           bf16* __restrict__ dZ_hi, bf16* __restrict__ dZ_lo,
                                                           float* __restrict__ dqa, float* __restrict__ gws_part,
                                                           bf16* __restrict__ ds_hi, bf16* __restrict__ ds_lo, int ldds,
                                                           int f16, float gscale) {
  RAU_PDL_ENTRY();
  __shared__ float ds[256];
  __shared__ float red[32];
  __shared__ float red2[2][1024];
  const int b = blockIdx.x, tid = threadIdx.x;
  const int64_t r0 = (int64_t)b * S;
  const float pv = tid < S ? p_in[r0 + tid] : 0.0f;
  const float dpv = tid < S ? dp[r0 + tid] : 0.0f;
  const float dot = block_sum(pv * dpv, red);
  const float dsv = tid < S ? pv * (dpv - dot) : 0.0f;
  ds[tid] = dsv;
  if (blockIdx.y == 0 && tid < S) ds_out[r0 + tid] = dsv;
  if (blockIdx.y == 0 && ds_hi && tid < ldds) {   // packed twin of ds for the Wm products
    const bf16 h = __float2bfloat16_rn(dsv);
    ds_hi[(int64_t)b * ldds + tid] = h;
    if (ds_lo) ds_lo[(int64_t)b * ldds + tid] = __float2bfloat16_rn(dsv - __bfloat162float(h));
  }
  __syncthreads();
  const int per = (S + gridDim.y - 1) / gridDim.y;
  const int s_lo = blockIdx.y * per, s_hi = min(S, s_lo + per);
  const int tpr = A >> 2, ngroups = 256 / tpr;
  const int ng = tid % tpr, rg = tid / tpr;
  const float4 w4 = *reinterpret_cast<const float4*>(ws + 4 * ng);
  float4 qa = make_float4(0.f, 0.f, 0.f, 0.f);
  if (RECOMP) qa = __ldg(reinterpret_cast<const float4*>(qadd + (int64_t)b * A) + ng);
  float sz0 = 0.f, sz1 = 0.f, sz2 = 0.f, sz3 = 0.f, sg0 = 0.f, sg1 = 0.f, sg2 = 0.f, sg3 = 0.f;
#pragma unroll 4
  for (int s = s_lo + rg; s < s_hi; s += ngroups) {
    float4 e = __ldg(reinterpret_cast<const float4*>(E + (r0 + s) * A) + ng);
    if (RECOMP == 1) { e.x = tanh_acc(e.x + qa.x); e.y = tanh_acc(e.y + qa.y); e.z = tanh_acc(e.z + qa.z); e.w = tanh_acc(e.w + qa.w); }
    if (RECOMP == 2) { e.x = tanh_hw(e.x + qa.x); e.y = tanh_hw(e.y + qa.y); e.z = tanh_hw(e.z + qa.z); e.w = tanh_hw(e.w + qa.w); }
    const float d = ds[s];
    const float z0 = w4.x * d * (1.0f - e.x * e.x), z1 = w4.y * d * (1.0f - e.y * e.y);
    const float z2 = w4.z * d * (1.0f - e.z * e.z), z3 = w4.w * d * (1.0f - e.w * e.w);
    sg0 = fmaf(d, e.x, sg0); sg1 = fmaf(d, e.y, sg1); sg2 = fmaf(d, e.z, sg2); sg3 = fmaf(d, e.w, sg3);
    sz0 += z0; sz1 += z1; sz2 += z2; sz3 += z3;
    uint32_t h0, l0, h1, l1;
    // dZ is carried times gscale, a power of two (the products that read it scale their fp32 results back)
    if (f16) {   // one fp16 plane
      reinterpret_cast<uint2*>(dZ_hi + (r0 + s) * A)[ng] =
          make_uint2(pack_f16x2(z0 * gscale, z1 * gscale), pack_f16x2(z2 * gscale, z3 * gscale));
      continue;
    }
    split_pair(z0 * gscale, z1 * gscale, h0, l0);
    split_pair(z2 * gscale, z3 * gscale, h1, l1);
    reinterpret_cast<uint2*>(dZ_hi + (r0 + s) * A)[ng] = make_uint2(h0, h1);
    if (dZ_lo) reinterpret_cast<uint2*>(dZ_lo + (r0 + s) * A)[ng] = make_uint2(l0, l1);
  }
  *reinterpret_cast<float4*>(&red2[0][rg * A + 4 * ng]) = make_float4(sz0, sz1, sz2, sz3);
  *reinterpret_cast<float4*>(&red2[1][rg * A + 4 * ng]) = make_float4(sg0, sg1, sg2, sg3);
  __syncthreads();
  for (int n = tid; n < A; n += 256) {
    float z = 0.0f, g = 0.0f;
    for (int k = 0; k < ngroups; ++k) { z += red2[0][k * A + n]; g += red2[1][k * A + n]; }
    atomicAdd(dqa + (int64_t)b * A + n, z);
    atomicAdd(gws_part + (int64_t)b * A + n, g);
  }
}


// ================================================================== persistent LSTM recurrence (question encoder)
// One launch walks ALL time steps of one LSTM layer: G_t = Gx_t + h_{t-1} Wh^T, cell update, h_t (D:22-45, the encoder's
// recurrent half; the input half Gx was hoisted over time).  A CTA owns 128 batch rows x 64 permuted gate columns (16
// hidden units) for the whole sequence: its slice of Wh stays in shared memory (hi and lo planes, 128 KB at Hq = 512) and
// its cell state in registers, so a step moves only h_{t-1} (L2 -> smem by TMA) and the step's outputs.  The CTAs of one
// row tile exchange h_t through global memory: writers release a per-tile counter, the TMA producer acquires it.
struct LsParams {
  CUtensorMap mapA, mapB;          // h stack [(T+1)*B, H] (hi, lo planes) ; Wh permuted [4H, H] (hi, lo planes)
  int B, H, T, tiles_n, nkb, stages, a_swap, b_swap;
  const float* Gx; long long gx_t; int ldg;     // [T][B][4H] hoisted input projection (+ biases), permuted columns
  float* c_out; float* h_out; long long s_t; int lds;   // state rows of step 1 (fp32), step stride, row pitch
  float* lsaved; long long ls_t, plane;         // saved gates of step 1: 5 planes (i, f, o, g, tanh c') of [B, H]
  bf16* hpk_hi; bf16* hpk_lo; long long hp_t;   // packed h of step 0 (zeros); step t at + t*hp_t
  unsigned int* counter;                        // [tiles_m] zeroed by the caller
  unsigned int* err;                            // sticky failure word (device: the optimizer kernel reads it)
  unsigned int* err_host;                       // the same, host-mapped: checked at every public entry point
  int fence_all;                                // every epilogue thread fences before the publish barrier (A/B)
  const float* lengths; float* sel_c; float* sel_h; int sel_ld;   // optional length selection (LstmSeq)
};
constexpr int LS_THREADS = 320;   // TMA producer, MMA issuer, 8 epilogue warps

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// 256-bit global accesses (sm_100): the per-row pieces of the recurrence's inputs and outputs are 32 contiguous bytes per
// thread -- one request each instead of two.  (Measured per step of lstm_seq_kernel with %globaltimer stamps: the SM's
// memory port is what bounds it -- 256 KB of h tiles + 32 KB of Gx in, 65 KB out per CTA -- and generic stores still in
// flight stretch the next handshake's gpu-scope fences by 1-2 us.)
__device__ __forceinline__ void st_global_v8(float* p, const float* v) {
  asm volatile("st.global.v8.f32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]),
               "f"(v[4]), "f"(v[5]), "f"(v[6]), "f"(v[7])
               : "memory");
}
__device__ __forceinline__ void ld_global_nc_v8(const float* p, float* v) {
  asm volatile("ld.global.nc.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void red_release_add_u32(unsigned int* p, unsigned int v) {
  asm volatile("red.release.gpu.global.add.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}

template <int X3>
__global__ void __launch_bounds__(LS_THREADS, 1) lstm_seq_kernel(const __grid_constant__ LsParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  __shared__ uint64_t full_bar[RT_MAXSTAGES], empty_bar[RT_MAXSTAGES], w_bar, tfull_bar, tempty_bar;
  __shared__ uint32_t tmem_base_s;
  uint8_t* smem = (uint8_t*)(((uintptr_t)smem_raw + 1023) & ~(uintptr_t)1023);
  constexpr uint32_t NT = X3 ? 2u : 1u;
  constexpr uint32_t w_plane = 64u * 64u * 2u, w_tile = NT * w_plane;        // per k-block of the weight slice
  constexpr uint32_t a_plane = 128u * 64u * 2u, a_stage = NT * a_plane;
  uint8_t* wsm = smem;
  uint8_t* ast = smem + (size_t)p.nkb * w_tile;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tn = blockIdx.x % p.tiles_n, tm = blockIdx.x / p.tiles_n;
  const int m0 = tm * 128, n0 = tn * 64;

  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
    mbar_init(&w_bar, 1); mbar_init(&tfull_bar, 1); mbar_init(&tempty_bar, 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.mapA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.mapB) : "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_base_s)), "r"(64u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = tmem_base_s;

  if (warp == 0) {
    // ===================== TMA producer: the weight slice once, then h_{t-1} of every step through the stage ring
    if (elect_one()) {
      mbar_expect_tx(&w_bar, (uint32_t)p.nkb * w_tile);
      for (int kb = 0; kb < p.nkb; ++kb) {
        if (X3) tma_load_3d(wsm + (size_t)kb * w_tile, &p.mapB, &w_bar, kb * 64, n0, 0);
        else tma_load_2d(wsm + (size_t)kb * w_tile, &p.mapB, &w_bar, kb * 64, n0);
      }
    }
    __syncwarp();
    int st = 0;          // stage ring position and phase (carried: no it % stages, it / stages in the loop)
    uint32_t ph = 0;
    for (int t = 1; t <= p.T; ++t) {
      if (t > 1) {   // every CTA of this row tile has published its columns of h_{t-1}
        const unsigned int need = (unsigned int)p.tiles_n * (unsigned int)(t - 1);
        if (lane == 0) {
          const long long t0 = clock64();
          while (ld_acquire_u32(p.counter + tm) < need) {
            if (clock64() - t0 > 4000000000ll) { *p.err = 1u; *p.err_host = 1u; break; }   // never hang the device on a lost peer
          }
        }
        __syncwarp();
        asm volatile("fence.proxy.async;" ::: "memory");   // generic-proxy writes of the peers -> this thread's TMA reads
      }
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(&empty_bar[st], ph ^ 1u);
        if (elect_one()) {
          mbar_expect_tx(&full_bar[st], a_stage);
          if (X3) tma_load_3d(ast + (size_t)st * a_stage, &p.mapA, &full_bar[st], kb * 64, (t - 1) * p.B + m0, 0);
          else tma_load_2d(ast + (size_t)st * a_stage, &p.mapA, &full_bar[st], kb * 64, (t - 1) * p.B + m0);
        }
        __syncwarp();
        if (++st == p.stages) { st = 0; ph ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer: M = 128, N = 64, K-major SWIZZLE_128B operands (see rows_gemm_kernel)
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((uint32_t)(128 >> 4) << 24);
    const uint64_t a_p1 = (uint64_t)(a_plane >> 4), b_p1 = (uint64_t)(w_plane >> 4);
    const uint64_t a_hi_off = p.a_swap ? a_p1 : 0, a_lo_off = p.a_swap ? 0 : a_p1;
    const uint64_t b_hi_off = p.b_swap ? b_p1 : 0, b_lo_off = p.b_swap ? 0 : b_p1;
    mbar_wait(&w_bar, 0u);
    int st = 0;
    uint32_t ph = 0;
    for (int t = 1; t <= p.T; ++t) {
      mbar_wait(&tempty_bar, ((uint32_t)(t - 1) & 1u) ^ 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      for (int kb = 0; kb < p.nkb; ++kb) {
        mbar_wait(&full_bar[st], ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        if (elect_one()) {
          const uint64_t da0 = make_desc(smem_u32(ast + (size_t)st * a_stage), 16u, 1024u, 2u);
          const uint64_t db0 = make_desc(smem_u32(wsm + (size_t)kb * w_tile), 16u, 1024u, 2u);
          const uint64_t da_hi = da0 + a_hi_off, da_lo = da0 + a_lo_off, db_hi = db0 + b_hi_off, db_lo = db0 + b_lo_off;
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t o = (uint64_t)k * (32u >> 4);
            umma_f16(tmem_base, da_hi + o, db_hi + o, idesc, (kb > 0 || k > 0) ? 1u : 0u);
            if (X3) {
              umma_f16(tmem_base, da_hi + o, db_lo + o, idesc, 1u);
              umma_f16(tmem_base, da_lo + o, db_hi + o, idesc, 1u);
            }
          }
          umma_commit(&empty_bar[st]);
        }
        __syncwarp();
        if (++st == p.stages) { st = 0; ph ^= 1u; }
      }
      if (elect_one()) umma_commit(&tfull_bar);
      __syncwarp();
    }
  } else {
    // ===================== epilogue warps 2..9: TMEM lanes 32*(warp%4).., one 32-column chunk (8 hidden units x
    // i, f, o, g) per warp; the cell state of those units lives in registers for the whole sequence
    const int q = warp & 3, c = (warp - 2) >> 2;
    const int r = m0 + q * 32 + lane;
    const bool r_ok = r < p.B;
    const int rr = r_ok ? r : p.B - 1;
    const int nc = n0 + c * 32;            // first permuted gate column
    const int u0 = (nc >> 5) << 3;         // first hidden unit
    const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32);
    float cs[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) cs[u] = 0.0f;
    const int len_r = p.lengths ? (int)p.lengths[rr] : 0;   // (0: never selected)
    for (int t = 1; t <= p.T; ++t) {
      float gx[32];   // the input half of this step's gates does not depend on the recurrence: fetch before waiting
      const float* gsrc = p.Gx + (long long)(t - 1) * p.gx_t + (long long)rr * p.ldg + nc;
#pragma unroll
      for (int k8 = 0; k8 < 4; ++k8) ld_global_nc_v8(gsrc + 8 * k8, gx + 8 * k8);
      mbar_wait(&tfull_bar, (uint32_t)(t - 1) & 1u);
      asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
      float v[32];
      tmem_ld32(taddr, v);
      asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar);
#pragma unroll
      for (int k = 0; k < 32; ++k) v[k] += gx[k];
      float gi[8], gf[8], go[8], gg[8], tc[8], hn[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        gi[u] = rcp_ftz(1.0f + exp2f_ftz(-1.4426950408889634f * v[u]));
        gf[u] = rcp_ftz(1.0f + exp2f_ftz(-1.4426950408889634f * v[8 + u]));
        go[u] = rcp_ftz(1.0f + exp2f_ftz(-1.4426950408889634f * v[16 + u]));
        gg[u] = tanh_acc(v[24 + u]);
        cs[u] = fmaf(gf[u], cs[u], gi[u] * gg[u]);
        tc[u] = tanh_acc(cs[u]);
        hn[u] = go[u] * tc[u];
      }
      if (r_ok) {
        // h_t packed first: it is what the peers are waiting for
        uint32_t hh[4], hl[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) split_pair(hn[2 * u], hn[2 * u + 1], hh[u], hl[u]);
        const long long ho = (long long)t * p.hp_t + (long long)r * p.H + u0;
        *reinterpret_cast<uint4*>(p.hpk_hi + ho) = make_uint4(hh[0], hh[1], hh[2], hh[3]);
        if (X3) *reinterpret_cast<uint4*>(p.hpk_lo + ho) = make_uint4(hl[0], hl[1], hl[2], hl[3]);
      }
      // publish: every writer fences, the epilogue warps meet, one thread's gpu-scope release bumps the tile counter.
      // (RAU_SEQ_FENCE=0 drops the per-thread fence and relies on the release's cumulativity over the barrier, the
      // grid-sync idiom: measured no faster -- 4.78 vs 4.78 ms -- so the belt-and-braces form stays the default.)
      if (p.fence_all) __threadfence();
      asm volatile("fence.proxy.async;" ::: "memory");
      asm volatile("bar.sync 1, 256;" ::: "memory");
      if (warp == 2 && lane == 0) red_release_add_u32(p.counter + tm, 1u);
      if (r_ok) {
        const long long so = (long long)(t - 1) * p.s_t + (long long)r * p.lds + u0;
        st_global_v8(p.c_out + so, cs);
        st_global_v8(p.h_out + so, hn);
        float* sbase = p.lsaved + (long long)(t - 1) * p.ls_t + (long long)r * p.H + u0;
        st_global_v8(sbase, gi);
        st_global_v8(sbase + p.plane, gf);
        st_global_v8(sbase + 2 * p.plane, go);
        st_global_v8(sbase + 3 * p.plane, gg);
        st_global_v8(sbase + 4 * p.plane, tc);
        if (t == len_r) {   // the question ends here: this row's encoder output (F:472-478)
          st_global_v8(p.sel_c + (long long)r * p.sel_ld + u0, cs);
          st_global_v8(p.sel_h + (long long)r * p.sel_ld + u0, hn);
        }
      }
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(64u) : "memory");
  }
}


bool g_prep_attr = false;

}  // namespace

bool rows_path_enabled() { return rau_process_tuning().rows != 0; }

int rows_gemm(rau_ctx* ctx, const RowsGemm& g) {
  RAU_REQUIRE(g.M > 0 && g.N > 0 && g.K > 0, "rows_gemm: bad shape %dx%dx%d", g.M, g.N, g.K);
  RAU_REQUIRE(g.A.hi && g.B.hi, "rows_gemm: operand missing");
  RAU_REQUIRE((g.A.lo != nullptr) == (g.B.lo != nullptr), "rows_gemm: hi/lo split must be on both operands or neither");
  RAU_REQUIRE(g.A.ld % 8 == 0 && g.B.ld % 8 == 0, "rows_gemm: operand pitches must be multiples of 8 elements");
  RAU_REQUIRE((((uintptr_t)g.A.hi | (uintptr_t)g.B.hi | (uintptr_t)g.A.lo | (uintptr_t)g.B.lo) & 15) == 0,
              "rows_gemm: operands must be 16-byte aligned");
  RAU_TRY(get_encode());
  RtParams p;
  memset(&p, 0, sizeof(p));
  // work enqueued on the side stream shares the GPU with the critical chain: it gets rows_cta_cap SMs
  const int sm_avail = (ctx->rows_cta_cap > 0 && ctx->rows_cta_cap < ctx->sm_count) ? ctx->rows_cta_cap
                       : (ctx->main_cta_cap > 0 && ctx->main_cta_cap < ctx->sm_count) ? ctx->main_cta_cap : ctx->sm_count;
  p.M = g.M; p.N = g.N; p.K = g.K;
  p.a_mn = g.A.mn; p.b_mn = g.B.mn;
  p.x3 = g.A.lo ? 1 : 0;
  p.tiles_m = (g.M + RT_BM - 1) / RT_BM;
  // accumulator width: whole 256-column tiles when there is enough work to fill the SMs, narrower tiles for the skinny
  // nn.Linear products (M = batch rows) so that more CTAs share the latency-bound work
  int BN = g.BN;
  if (BN == 0) {
    BN = 256;
    // (split-K reductions keep 256-wide tiles unless the output is a thin [<= 256, N] slab: their SM fill comes from
    // the K split, and wide tiles re-read less of the A operand)
    if (g.epi == EPI_LINEAR || g.epi == EPI_PLAIN || g.epi == EPI_LSTM || (g.epi == EPI_RED && p.tiles_m <= 2 && g.K < 8192))
      while (BN > 64 && (long long)p.tiles_m * ((g.N + BN - 1) / BN) < ctx->sm_count) BN >>= 1;
    while (BN > 64 && g.N <= BN / 2) BN >>= 1;
  }
  RAU_REQUIRE(BN == 64 || BN == 128 || BN == 256, "rows_gemm: BN = %d", BN);
  p.BN = BN;
  p.tiles_n = (g.N + BN - 1) / BN;
  // k elements per stage: one TMA instruction costs its issuing thread ~250 cycles whatever the box size, so the narrow
  // tiles (little MMA time per k-block) take 64-wide k-blocks even in bf16x3; 256-wide bf16x3 tiles keep 48 KB stages
  p.BK = (p.x3 && BN == 256) ? 32 : 64;
  p.nkb = (g.K + p.BK - 1) / p.BK;
  const bool seg2 = g.K2 > 0;
  {   // CTA pairs (cta_group::2) for the big K-major products: RAU_CG2=0 keeps everything on single CTAs
    const int cg2_on = ctx->tune.cg2;
    // (EPI_LINEAR: the encoder's hoisted input projections and input gradients, [T*B, in] x [in, 4H]: RAU_LIN_CG2=0 keeps single CTAs)
    const bool pair_epi = g.epi == EPI_PLAIN || g.epi == EPI_TANH || g.epi == EPI_ATT || g.epi == EPI_DY || g.epi == EPI_RED ||
                          (g.epi == EPI_LINEAR && ctx->tune.lin_cg2 != 0);
    // big row counts (the image-side products), or split-K reductions with at least one pair of row tiles
    const bool pair_shape = g.epi == EPI_RED ? (p.tiles_m >= 2 && p.tiles_m % 2 == 0 && p.nkb >= 64) : p.tiles_m >= 8;
    p.cg2 = (cg2_on && pair_epi && pair_shape && BN == 256 && !seg2 && sm_avail >= 2) ? 1 : 0;
  }
  if (seg2) {
    RAU_REQUIRE(g.A2.hi && g.B2.hi && g.A2.mn == g.A.mn && g.B2.mn == g.B.mn && (g.A2.lo != nullptr) == (g.A.lo != nullptr) &&
                    (g.B2.lo != nullptr) == (g.B.lo != nullptr) && g.A2.ld % 8 == 0 && g.B2.ld % 8 == 0,
                "rows_gemm: bad second K segment");
    RAU_REQUIRE((((uintptr_t)g.A2.hi | (uintptr_t)g.B2.hi | (uintptr_t)g.A2.lo | (uintptr_t)g.B2.lo) & 15) == 0,
                "rows_gemm: operands must be 16-byte aligned");
  }
  p.nkb1 = p.nkb;
  if (seg2) p.nkb += (g.K2 + p.BK - 1) / p.BK;
  // shared memory: 8 staging buffers for the epilogue warps + as many operand stages as fit (latency-bound skinny
  // products want many small stages in flight, the big ones four 48 KB stages)
  p.stg_warp = (g.epi == EPI_LINEAR && g.out_hi) ? 8192 : RT_STG_WARP;
  p.stage_bytes = (p.x3 ? 2 : 1) * (RT_BM + (p.cg2 ? BN / 2 : BN)) * p.BK * 2;
  if (g.epi == EPI_DY) p.stg_warp = 9216;   // + a 32 x 64 B (hi, lo) scratch per warp (the saved activation is fetched coalesced) + 1 KB row vector
  {   // the 1-pass (fp16 / bf16) CTA-pair i_embed product is epilogue bound: 16 epilogue warps (RAU_TANH_EW=8: A/B switch)
    const int ew_tanh = ctx->tune.tanh_ew;
    // (K <= 1024: at K = 2048 the product is tensor bound and the stage the extra staging buffers cost matters more)
    p.ew = (g.epi == EPI_TANH && p.cg2 && !p.x3 && g.K <= 1024) ? ew_tanh : 8;
  }
  p.stages = (RT_SMEM_BUDGET - p.ew * p.stg_warp) / p.stage_bytes;
  if (p.stages > RT_MAXSTAGES) p.stages = RT_MAXSTAGES;
  const int tiles = (p.cg2 ? (p.tiles_m + 1) / 2 : p.tiles_m) * p.tiles_n;   // work items before any K split
  p.ksplit = 1;
  const int workers = p.cg2 ? sm_avail / 2 : sm_avail;   // CTAs, or CTA pairs
  if (g.epi == EPI_RED && tiles < workers) {
    int want = workers / tiles;
    if (want > p.nkb) want = p.nkb;
    if (want < 1) want = 1;
    p.ksplit = want;
  }
  p.kb_per = (p.nkb + p.ksplit - 1) / p.ksplit;
  p.ksplit = (p.nkb + p.kb_per - 1) / p.kb_per;
  RAU_REQUIRE((long long)tiles * p.ksplit < 65536 && p.ksplit < 65536 && p.tiles_n < 65536, "rows_gemm: %d work items", tiles * p.ksplit);
  p.ksplit_magic = (unsigned int)((0x100000000ull + (unsigned long long)p.ksplit - 1) / (unsigned long long)p.ksplit);     // (unused when 1)
  p.tiles_n_magic = (unsigned int)((0x100000000ull + (unsigned long long)p.tiles_n - 1) / (unsigned long long)p.tiles_n);
  RAU_TRY(encode_operand(&p.mapA[0], g.A.hi, g.A.lo, &p.a_swap, g.A.mn, g.M, g.K, g.A.ld, p.BK, RT_BM));
  RAU_TRY(encode_operand(&p.mapB[0], g.B.hi, g.B.lo, &p.b_swap, g.B.mn, g.N, g.K, g.B.ld, p.BK, p.cg2 ? BN / 2 : BN));
  if (seg2) {
    int sa2 = 0, sb2 = 0;
    RAU_TRY(encode_operand(&p.mapA2[0], g.A2.hi, g.A2.lo, &sa2, g.A2.mn, g.M, g.K2, g.A2.ld, p.BK, RT_BM));
    RAU_TRY(encode_operand(&p.mapB2[0], g.B2.hi, g.B2.lo, &sb2, g.B2.mn, g.N, g.K2, g.B2.ld, p.BK, BN));
    RAU_REQUIRE(sa2 == p.a_swap && sb2 == p.b_swap, "rows_gemm: the two K segments order their hi/lo arrays differently");
  }
  const bool f32_out = g.epi == EPI_PLAIN || g.epi == EPI_RED || g.epi == EPI_ATT;
  if (g.epi == EPI_LINEAR) {
    RAU_REQUIRE(g.out_f || g.out_hi, "rows_gemm: EPI_LINEAR without an output");
    if (g.out_f) {
      RAU_REQUIRE(g.ldo % 4 == 0 && ((uintptr_t)g.out_f & 15) == 0, "rows_gemm: fp32 output must be 16-byte aligned");
      RAU_TRY(encode_2d(&p.mapO[0], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g.out_f, (uint64_t)g.N, (uint64_t)g.M, (uint64_t)g.ldo, 32, 32,
                        CU_TENSOR_MAP_SWIZZLE_128B));
      p.out_f = 1;
    }
    if (g.out_hi) {
      RAU_REQUIRE(g.ldo_b % 8 == 0 && (((uintptr_t)g.out_hi | (uintptr_t)g.out_lo) & 15) == 0, "rows_gemm: bf16 output alignment");
      RAU_TRY(encode_2d(&p.mapO[1], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.out_hi, (uint64_t)g.N, (uint64_t)g.M, (uint64_t)g.ldo_b, 32,
                        32, CU_TENSOR_MAP_SWIZZLE_64B));
      if (g.out_lo)
        RAU_TRY(encode_2d(&p.mapO[2], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.out_lo, (uint64_t)g.N, (uint64_t)g.M, (uint64_t)g.ldo_b,
                          32, 32, CU_TENSOR_MAP_SWIZZLE_64B));
      p.out_hi = 1;
      p.out_lo = g.out_lo ? 1 : 0;
    }
    p.bias2 = g.bias2; p.addend = g.addend; p.addend2 = g.addend2; p.ldadd = g.ldadd; p.act = g.act;
    p.vec = (g.N % 32 == 0) && ((((uintptr_t)g.bias | (uintptr_t)g.bias2 | (uintptr_t)g.addend | (uintptr_t)g.addend2) & 15) == 0) &&
            (g.ldadd % 4 == 0);
  } else if (g.epi == EPI_LSTM) {
    const int H_ = g.N / 4;
    RAU_REQUIRE(g.N % 32 == 0 && g.c_out && g.h_out, "rows_gemm: bad EPI_LSTM arguments (4H = %d)", g.N);
    RAU_REQUIRE(g.ldc % 4 == 0 && g.ldh % 4 == 0 && g.ldcp % 4 == 0 && g.ldadd % 4 == 0 && g.ldhp % 8 == 0 &&
                    ((((uintptr_t)g.c_out | (uintptr_t)g.h_out | (uintptr_t)g.c_prev | (uintptr_t)g.lsaved | (uintptr_t)g.bias |
                       (uintptr_t)g.addend | (uintptr_t)g.hpk_hi | (uintptr_t)g.hpk_lo) & 15) == 0),
                "rows_gemm: EPI_LSTM tensors must be 16-byte aligned");
    p.addend = g.addend; p.ldadd = g.ldadd;
    p.c_prev = g.c_prev; p.ldcp = g.ldcp; p.c_out = g.c_out; p.ldc = g.ldc; p.h_out = g.h_out; p.ldh = g.ldh;
    p.lsaved = g.lsaved; p.plane = (long long)g.M * H_;
    p.hpk_hi = g.hpk_hi; p.hpk_lo = g.hpk_lo; p.ldhp = g.ldhp;
  } else if (f32_out) {
    RAU_REQUIRE(g.out_f && g.ldo % 4 == 0 && ((uintptr_t)g.out_f & 15) == 0, "rows_gemm: fp32 output must be 16-byte aligned");
    RAU_TRY(encode_2d(&p.mapO[0], CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g.out_f, (uint64_t)g.N, (uint64_t)g.M, (uint64_t)g.ldo, 32, 32,
                      CU_TENSOR_MAP_SWIZZLE_128B));
  } else {
    RAU_REQUIRE(g.out_hi && g.ldo % 8 == 0 && (((uintptr_t)g.out_hi | (uintptr_t)g.out_lo) & 15) == 0,
                "rows_gemm: bf16 output must be 16-byte aligned");
    RAU_TRY(encode_2d(&p.mapO[0], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.out_hi, (uint64_t)g.N, (uint64_t)g.M, (uint64_t)g.ldo, 32, 32,
                      CU_TENSOR_MAP_SWIZZLE_64B));
    if (g.out_lo)
      RAU_TRY(encode_2d(&p.mapO[1], CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, g.out_lo, (uint64_t)g.N, (uint64_t)g.M, (uint64_t)g.ldo, 32, 32,
                        CU_TENSOR_MAP_SWIZZLE_64B));
    p.out_lo = g.out_lo ? 1 : 0;
  }
  if (g.epi == EPI_TANH || g.epi == EPI_ATT || g.epi == EPI_DY)
    RAU_REQUIRE(g.N % 32 == 0, "rows_gemm: fused epilogues need N %% 32 == 0 (N = %d)", g.N);
  if (g.epi == EPI_ATT) RAU_REQUIRE(g.N <= BN && g.S > 0 && g.rowvec && g.colw && g.bias, "rows_gemm: bad EPI_ATT arguments");
  if (g.epi == EPI_DY) RAU_REQUIRE(g.S > 0 && g.rowvec && g.rowscale && g.aux_hi && g.ldaux % 8 == 0, "rows_gemm: bad EPI_DY arguments");
  if (g.epi == EPI_TANH) RAU_REQUIRE(g.bias != nullptr, "rows_gemm: EPI_TANH needs a bias");
  p.bias = g.bias; p.rowvec = g.rowvec; p.colw = g.colw; p.rowout = g.rowout; p.rowscale = g.rowscale;
  p.aux_hi = g.aux_hi; p.aux_lo = g.aux_lo; p.ldaux = g.ldaux; p.colsum = g.colsum; p.S = g.S > 0 ? g.S : 1;
  p.alpha = g.alpha;
  p.f16 = g.f16 ? 1 : 0;
  p.of16 = g.of16 ? 1 : 0;
  p.af16 = g.af16 ? 1 : 0;
  p.gscale = g.gscale;
  RAU_REQUIRE(!(p.f16 && p.x3), "rows_gemm: fp16 operands are single-plane");
  RAU_REQUIRE(!(p.of16 && g.out_lo), "rows_gemm: an fp16 output is single-plane");
  // single-pass bf16: the result is rounded to bf16 anyway, MUFU.TANH (2^-11) is below that; fp16 keeps the accurate form
  p.fast_tanh = (p.x3 || p.f16) ? 0 : 1;
  const int items = tiles * p.ksplit;
  int grid = items < sm_avail ? items : sm_avail;
  if (p.cg2) grid = 2 * (items < sm_avail / 2 ? items : sm_avail / 2);   // whole CTA pairs
  {
    if (ctx->tune.rows_trace) {   // debugging aid: the stamps of the LAST launch are left in the arena buffer "rows.trace"
      void* buf = nullptr;
      RAU_TRY(ctx->arena.get("rows.trace", sizeof(unsigned long long) * 16 * 148, &buf));
      RAU_CHECK_CUDA(cudaMemsetAsync(buf, 0, sizeof(unsigned long long) * 16 * 148, ctx->stream));
      p.dbg = (unsigned long long*)buf;
    }
  }
  const int smem_bytes = p.stages * p.stage_bytes + p.ew * p.stg_warp + 1024;
  switch (g.epi) {
    case EPI_PLAIN: return launch_rows<EPI_PLAIN>(ctx, p, grid, smem_bytes);
    case EPI_RED: return launch_rows<EPI_RED>(ctx, p, grid, smem_bytes);
    case EPI_TANH: return launch_rows<EPI_TANH>(ctx, p, grid, smem_bytes);
    case EPI_ATT: return launch_rows<EPI_ATT>(ctx, p, grid, smem_bytes);
    case EPI_DY: return launch_rows<EPI_DY>(ctx, p, grid, smem_bytes);
    case EPI_LINEAR: return launch_rows<EPI_LINEAR>(ctx, p, grid, smem_bytes);
    case EPI_LSTM: return launch_rows<EPI_LSTM>(ctx, p, grid, smem_bytes);
    default: rau_set_error("rows_gemm: unknown epilogue %d", g.epi); return RAU_EINVAL;
  }
}

int rows_pack(rau_ctx* ctx, const float* W, int64_t n, bool want_lo, bool cache, const char* slot, const bf16** hi, const bf16** lo,
              bool f16) {
  RAU_REQUIRE(n % 8 == 0 && ((uintptr_t)W & 15) == 0, "rows_pack: %lld elements / alignment", (long long)n);
  if (f16) want_lo = false;
  char name[128];
  bool cached = false;
  if (cache) {
    snprintf(name, sizeof(name), "rw.%p.%lld.%d", (const void*)W, (long long)n, f16 ? 2 : (want_lo ? 1 : 0));
    auto it = ctx->tc_epoch.find(name);
    cached = it != ctx->tc_epoch.end() && it->second == ctx->epoch;
  } else {
    snprintf(name, sizeof(name), "rw.%s", slot);
  }
  void* buf = nullptr;
  const size_t half = ((size_t)n * sizeof(bf16) + 1023) / 1024 * 1024;
  RAU_TRY(ctx->arena.get(name, half * (want_lo ? 2 : 1), &buf));
  bf16* h = (bf16*)buf;
  bf16* l = want_lo ? (bf16*)((char*)buf + half) : nullptr;
  if (!cached) {
    int64_t blocks = (n / 4 + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    RAU_LAUNCH_PDL(ctx->stream, (pack_hilo_kernel), (int)blocks, 256, 0, W, n / 4, h, l, f16 ? 1 : 0);
    RAU_LAUNCH_CHECK(ctx);
    if (cache) ctx->tc_epoch[name] = ctx->epoch;
  }
  *hi = h;
  *lo = l;
  return RAU_OK;
}

static int prep_attr() {
  if (!g_prep_attr) {
    RAU_CHECK_CUDA(cudaFuncSetAttribute(xprep_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 66 * 4));
    RAU_CHECK_CUDA(cudaFuncSetAttribute(unprep_rows_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 256 * 66 * 4));
    g_prep_attr = true;
  }
  return RAU_OK;
}

int k_unpack_hilo(rau_ctx* ctx, const bf16* hi, const bf16* lo, int64_t n, float* out, int f16) {
  unpack_hilo_kernel<<<(int)((n + 255) / 256 < 4096 ? (n + 255) / 256 : 4096), 256, 0, ctx->stream>>>(hi, lo, n, out, f16);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}

int k_xprep_rows_hops(rau_ctx* ctx, const float* X, const void* X16, int B, int C, int S, int nHop, float scale, bf16* hi, bf16* lo,
                      int64_t hop_stride, float p_drop, uint64_t stream_id, int f16) {
  RAU_REQUIRE(C % 64 == 0 && S % 4 == 0 && S <= 200 && hop_stride % 2 == 0, "k_xprep_rows_hops: C=%d S=%d", C, S);
  RAU_REQUIRE(X16 ? ((uintptr_t)X16 & 7) == 0 : (X != nullptr && ((uintptr_t)X & 15) == 0), "k_xprep_rows_hops: feature alignment");
  const int smem = 4 * S * 66 * 4;
  static bool attr = false;
  if (!attr) {
    RAU_CHECK_CUDA(cudaFuncSetAttribute(xprep_rows_hops_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 4 * 200 * 66 * 4));
    attr = true;
  }
  const double keep = 1.0 - (double)p_drop;
  const uint32_t thresh = keep >= 1.0 ? 65536u : (uint32_t)(keep * 65536.0 + 0.5);   // 16-bit keep threshold
  const int cap = (ctx->rows_cta_cap > 0 && ctx->rows_cta_cap < ctx->sm_count) ? ctx->rows_cta_cap : ctx->sm_count;
  const int ntiles = (C / 64) * B;
  const int grid = (ntiles + 3) / 4 < cap ? (ntiles + 3) / 4 : cap;
  RAU_LAUNCH_PDL(ctx->stream, (xprep_rows_hops_kernel), grid, 1024, smem,
      X, (const __half*)X16, scale, C, S, B, nHop, hi, lo, (long long)hop_stride, thresh, make_uint2((uint32_t)ctx->seed, (uint32_t)(ctx->seed >> 32)),
      (uint32_t)stream_id, (uint32_t)(stream_id >> 32), ctx->ss_active, f16);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}

int k_xprep_rows(rau_ctx* ctx, const float* X, int B, int C, int S, const uint32_t* bits, float scale, bf16* hi, bf16* lo,
                 int gen, float p_drop, uint64_t stream_id, int f16, int hop, int nHop) {
  RAU_REQUIRE(C % 64 == 0 && S % 4 == 0 && S <= 256 && ((uintptr_t)X & 15) == 0, "k_xprep_rows: C=%d S=%d", C, S);
  RAU_TRY(prep_attr());
  const double keep = 1.0 - (double)p_drop;
  const uint32_t thresh = keep >= 1.0 ? 65536u : (uint32_t)(keep * 65536.0 + 0.5);   // 16-bit keep threshold (gen path)
  RAU_LAUNCH_PDL(ctx->stream, (xprep_rows_kernel), dim3(C / 64, B), 256, S * 66 * 4, 
      X, bits, scale, C, S, hi, lo, gen, thresh, make_uint2((uint32_t)ctx->seed, (uint32_t)(ctx->seed >> 32)), (uint32_t)stream_id,
      (uint32_t)(stream_id >> 32), ctx->ss_active, f16, hop, nHop);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}

int k_unprep_rows(rau_ctx* ctx, const float* dXr, int B, int C, int S, const uint32_t* bits, float scale, float* dX) {
  RAU_REQUIRE(C % 64 == 0 && S % 4 == 0 && S <= 256 && ((uintptr_t)dX & 15) == 0, "k_unprep_rows: C=%d S=%d", C, S);
  RAU_TRY(prep_attr());
  RAU_LAUNCH_PDL(ctx->stream, (unprep_rows_kernel), dim3(C / 64, B), 256, S * 66 * 4, dXr, bits, scale, C, S, dX);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}

int k_attn_rows_fwd(rau_ctx* ctx, int B, int M, int S, const float* logit, const float* mem, const bf16* I_hi, const bf16* I_lo,
                    float* p, float* a, bf16* p_hi, bf16* p_lo, int ldp, int f16) {
  RAU_REQUIRE(M % 256 == 0 && S <= 256 && ldp <= 256, "k_attn_rows_fwd: M=%d S=%d", M, S);
  RAU_LAUNCH_PDL(ctx->stream, (attn_rows_fwd_kernel<0>), dim3(B, M / 256), 256, 0, S, M, logit, mem, I_hi, I_lo, p, a, p_hi, p_lo, ldp,
                 0, (const float*)nullptr, (const float*)nullptr, (const float*)nullptr, f16);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}

int k_attn_rows_score(rau_ctx* ctx, int B, int A, int S, const float* Z, const float* qadd, const float* ws, int fast_tanh,
                      float* logit) {
  RAU_REQUIRE(A % 4 == 0, "k_attn_rows_score: A=%d", A);
  const int R = B * S;
  int grid = (R + 31) / 32;
  {   // exactly one resident wave (the occupancy the compiler's register count allows: 56 registers -> 4 CTAs per SM, not
      // the 6 an earlier cut assumed, which left a 0.32-wave tail): a warp then walks several batches of 4 rows
    static int per_sm[2] = {0, 0};
    const int v = fast_tanh ? 1 : 0;
    if (per_sm[v] == 0) {
      int n = 0;
      cudaError_t e = v ? cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, attn_rows_score_kernel<1>, 256, 0)
                        : cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, attn_rows_score_kernel<0>, 256, 0);
      per_sm[v] = (e == cudaSuccess && n > 0) ? n : 4;
    }
    const int one_wave = ctx->sm_count * per_sm[v];
    if (grid > one_wave) grid = one_wave;
  }
  if (fast_tanh) RAU_LAUNCH_PDL(ctx->stream, (attn_rows_score_kernel<1>), grid, 256, 0, R, S, A, Z, qadd, ws, logit);
  else RAU_LAUNCH_PDL(ctx->stream, (attn_rows_score_kernel<0>), grid, 256, 0, R, S, A, Z, qadd, ws, logit);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}

int k_attn_rows_bwd(rau_ctx* ctx, int B, int M, int A, int S, const float* E, const bf16* I_hi, const bf16* I_lo, const float* ws,
                    const float* p, const float* dp_in, const float* da, float* ds, bf16* dZ_hi, bf16* dZ_lo, float* dqa,
                    float* gws_part, bf16* ds_hi, bf16* ds_lo, int ldds, const float* qadd, int fast_tanh, int acc_zeroed, int f16,
                    float gscale) {
  RAU_REQUIRE(M % 8 == 0 && S <= 256 && ldds <= 256 && (A == 64 || A == 128 || A == 256), "k_attn_rows_bwd: M=%d A=%d S=%d", M, A, S);
  const int R = B * S;
  float* dp = nullptr;
  RAU_TRY(ctx->arena.get("attn.dp", sizeof(float) * (size_t)R, (void**)&dp));
  RAU_LAUNCH_PDL(ctx->stream, (attn_rows_dp_kernel), (R + 7) / 8, 256, 0, R, S, M, I_hi, I_lo, da, dp_in, dp, f16);
  RAU_LAUNCH_CHECK(ctx);
  if (!acc_zeroed) {
    RAU_CHECK_CUDA(cudaMemsetAsync(dqa, 0, sizeof(float) * (size_t)B * A, ctx->stream));
    RAU_CHECK_CUDA(cudaMemsetAsync(gws_part, 0, sizeof(float) * (size_t)B * A, ctx->stream));
  }
  const int nsl = B >= 128 ? 4 : (B >= 32 ? 8 : 16);   // row slices per image: ~1000 CTAs
  if (!qadd)
    RAU_LAUNCH_PDL(ctx->stream, (attn_rows_dz_kernel<0>), dim3(B, nsl), 256, 0, S, A, E, qadd, ws, p, dp, ds, dZ_hi, dZ_lo, dqa,
                   gws_part, ds_hi, ds_lo, ldds, f16, gscale);
  else if (!fast_tanh)
    RAU_LAUNCH_PDL(ctx->stream, (attn_rows_dz_kernel<1>), dim3(B, nsl), 256, 0, S, A, E, qadd, ws, p, dp, ds, dZ_hi, dZ_lo, dqa,
                   gws_part, ds_hi, ds_lo, ldds, f16, gscale);
  else
    RAU_LAUNCH_PDL(ctx->stream, (attn_rows_dz_kernel<2>), dim3(B, nsl), 256, 0, S, A, E, qadd, ws, p, dp, ds, dZ_hi, dZ_lo, dqa,
                   gws_part, ds_hi, ds_lo, ldds, f16, gscale);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}



template <typename P>
static cudaError_t launch_cooperative(cudaStream_t st, void (*kern)(P), int grid, int block, int smem, const P& p) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeCooperative;
  at[0].val.cooperative = 1;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, p);
}

// The whole recurrence of one encoder LSTM layer in one persistent launch (lstm_seq_kernel).  Returns RAU_OK with *done = 0
// when the shape does not fit (the caller then unrolls the per-step EPI_LSTM launches).
int rows_lstm_seq(rau_ctx* ctx, const LstmSeq& d, int* done) {
  *done = 0;
  const int on = ctx->tune.lstm_seq;   // RAU_LSTM_SEQ=0: one launch per recurrent step
  const int B = d.B, H = d.H, T = d.T;
  const bool x3 = d.hpk_lo != nullptr;
  if (!on || H % 64 != 0 || T < 1 || !d.hpk_hi || !d.Wh_hi || (x3 != (d.Wh_lo != nullptr))) return RAU_OK;
  const int nkb = H / 64, tiles_m = (B + 127) / 128, tiles_n = 4 * H / 64;
  const int w_bytes = nkb * (x3 ? 2 : 1) * 64 * 64 * 2, a_stage = (x3 ? 2 : 1) * 128 * 64 * 2;
  int stages = (RT_SMEM_BUDGET - w_bytes) / a_stage;
  if (stages > RT_MAXSTAGES) stages = RT_MAXSTAGES;
  // every CTA of the grid must be resident at once (they wait for each other): one CTA per SM
  if (stages < 2 || tiles_m * tiles_n > ctx->sm_count) return RAU_OK;
  RAU_REQUIRE(d.ldwh % 8 == 0, "rows_lstm_seq: pitches");
  // (the epilogue moves 32-byte pieces: every pitch and base it touches must be a multiple of 32 bytes)
  if (d.ldg % 8 != 0 || d.lds % 8 != 0 || d.gx_t % 8 != 0 || d.s_t % 8 != 0 || d.ls_t % 8 != 0 || d.plane % 8 != 0 ||
      (((uintptr_t)d.Gx | (uintptr_t)d.c_out | (uintptr_t)d.h_out | (uintptr_t)d.lsaved) & 31) != 0)
    return RAU_OK;
  RAU_TRY(get_encode());
  LsParams p;
  memset(&p, 0, sizeof(p));
  RAU_TRY(encode_operand(&p.mapA, d.hpk_hi, d.hpk_lo, &p.a_swap, 0, (T + 1) * B, H, H, 64, 128));
  RAU_TRY(encode_operand(&p.mapB, d.Wh_hi, d.Wh_lo, &p.b_swap, 0, 4 * H, H, d.ldwh, 64, 64));
  p.B = B; p.H = H; p.T = T; p.tiles_n = tiles_n; p.nkb = nkb; p.stages = stages;
  p.Gx = d.Gx; p.gx_t = d.gx_t; p.ldg = d.ldg;
  p.c_out = d.c_out; p.h_out = d.h_out; p.s_t = d.s_t; p.lds = d.lds;
  p.lsaved = d.lsaved; p.ls_t = d.ls_t; p.plane = d.plane;
  p.hpk_hi = d.hpk_hi; p.hpk_lo = d.hpk_lo; p.hp_t = (long long)B * H;
  unsigned int* cnt = d.counter;
  RAU_REQUIRE(tiles_m <= 64, "rows_lstm_seq: %d row tiles", tiles_m);
  if (cnt == nullptr) {
    RAU_TRY(ctx->arena.get("lstmseq.cnt", sizeof(unsigned int) * 64, (void**)&cnt));
    RAU_CHECK_CUDA(cudaMemsetAsync(cnt, 0, sizeof(unsigned int) * 64, ctx->stream));
  }
  p.counter = cnt; p.err = ctx->d_err; p.err_host = ctx->h_err_dev;
  if (d.lengths && d.sel_c && d.sel_h && d.sel_ld % 8 == 0 && ((((uintptr_t)d.sel_c | (uintptr_t)d.sel_h) & 31) == 0)) {
    p.lengths = d.lengths; p.sel_c = d.sel_c; p.sel_h = d.sel_h; p.sel_ld = d.sel_ld;
  } else if (d.lengths) {
    return RAU_OK;   // (the caller asked for the fused selection and the layout does not allow it: per-step launches + select)
  }
  p.fence_all = 1;
  const int smem_bytes = w_bytes + stages * a_stage + 1024;
  static bool attr_done[2] = {false, false};
  if (!attr_done[x3 ? 1 : 0]) {
    if (x3) RAU_CHECK_CUDA(cudaFuncSetAttribute(lstm_seq_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, RT_SMEM_BUDGET + 1024));
    else RAU_CHECK_CUDA(cudaFuncSetAttribute(lstm_seq_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, RT_SMEM_BUDGET + 1024));
    attr_done[x3 ? 1 : 0] = true;
  }
  // the CTAs wait for each other: a cooperative launch makes the driver guarantee that the whole grid is co-resident
  // (next to whatever the side and aux streams keep on the device) or refuse the launch
  if (x3) RAU_CHECK_CUDA(launch_cooperative(ctx->stream, lstm_seq_kernel<1>, tiles_m * tiles_n, LS_THREADS, smem_bytes, p));
  else RAU_CHECK_CUDA(launch_cooperative(ctx->stream, lstm_seq_kernel<0>, tiles_m * tiles_n, LS_THREADS, smem_bytes, p));
  RAU_LAUNCH_CHECK(ctx);
  *done = 1;
  return RAU_OK;
}


// ================================================================== nn.Linear adapter
namespace {
// fp32 [rows, cols] (pitch ld) -> packed bf16 (hi [, lo]) [rows, ldo] with ldo = cols rounded up to 8, zero padded
// wcols (even, <= ldo): columns written per packed row -- ldo for a whole row (zero padded past cols), less when the
// destination is a column block of a wider packed matrix
__global__ void pack2d_kernel(const float* __restrict__ in, int64_t ld, int rows, int cols, int ldo, bf16* __restrict__ hi,
                              bf16* __restrict__ lo, int wcols) {
  RAU_PDL_ENTRY();
  const int q = wcols >> 1;   // pairs written per packed row
  const int64_t total = (int64_t)rows * q, pitch = ldo >> 1;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % q) * 2;
    const int64_t r = i / q;
    const float a = c < cols ? in[r * ld + c] : 0.0f;
    const float b = c + 1 < cols ? in[r * ld + c + 1] : 0.0f;
    uint32_t h, l;
    split_pair(a, b, h, l);
    const int64_t o = r * pitch + (c >> 1);
    reinterpret_cast<uint32_t*>(hi)[o] = h;
    if (lo) reinterpret_cast<uint32_t*>(lo)[o] = l;
  }
}

struct Packed2D { const bf16* hi = nullptr; const bf16* lo = nullptr; int64_t ld = 0; };

// pack a strided fp32 matrix; parameter tensors (is_const) are packed once per epoch
int pack2d(rau_ctx* ctx, const float* src, int64_t ld, int rows, int cols, bool want_lo, bool is_const, const char* slot,
           Packed2D* out) {
  const int ldo = (cols + 7) / 8 * 8;
  char name[128];
  bool cached = false;
  if (is_const) {
    snprintf(name, sizeof(name), "rp.%p.%d.%d.%lld.%d", (const void*)src, rows, cols, (long long)ld, want_lo ? 1 : 0);
    auto it = ctx->tc_epoch.find(name);
    cached = it != ctx->tc_epoch.end() && it->second == ctx->epoch;
  } else {
    // scratch slots are per stream: work enqueued on the side stream must not share them with the chain
    snprintf(name, sizeof(name), ctx->rows_cta_cap > 0 ? "rp.side.%s" : "rp.%s", slot);
  }
  const size_t half = ((size_t)rows * ldo * sizeof(bf16) + 1023) / 1024 * 1024;
  void* buf = nullptr;
  RAU_TRY(ctx->arena.get(name, half * (want_lo ? 2 : 1), &buf));
  out->hi = (bf16*)buf;
  out->lo = want_lo ? (bf16*)((char*)buf + half) : nullptr;
  out->ld = ldo;
  if (!cached) {
    const int64_t work = (int64_t)rows * (ldo / 2);
    int64_t blocks = (work + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (blocks < 1) blocks = 1;
    RAU_LAUNCH_PDL(ctx->stream, (pack2d_kernel), (int)blocks, 256, 0, src, ld, rows, cols, ldo, (bf16*)out->hi, (bf16*)out->lo, ldo);
    RAU_LAUNCH_CHECK(ctx);
    if (is_const) ctx->tc_epoch[name] = ctx->epoch;
  }
  return RAU_OK;
}

}  // namespace
int rows_pack_into(rau_ctx* ctx, const float* src, int64_t ld, int rows, int cols, bf16* hi, bf16* lo, int64_t ldo, int wcols) {
  RAU_REQUIRE(hi != nullptr && ldo % 2 == 0 && ldo >= cols, "rows_pack_into: bad destination (ldo = %lld)", (long long)ldo);
  if (wcols <= 0) wcols = (int)ldo;
  RAU_REQUIRE(wcols % 2 == 0 && wcols <= ldo && wcols >= cols, "rows_pack_into: wcols = %d", wcols);
  const int64_t work = (int64_t)rows * (wcols / 2);
  int64_t blocks = (work + 255) / 256;
  if (blocks > 148 * 8) blocks = 148 * 8;
  if (blocks < 1) blocks = 1;
  RAU_LAUNCH_PDL(ctx->stream, (pack2d_kernel), (int)blocks, 256, 0, src, ld, rows, cols, (int)ldo, hi, lo, wcols);
  RAU_LAUNCH_CHECK(ctx);
  return RAU_OK;
}

int rows_pack2d(rau_ctx* ctx, const float* src, int64_t ld, int rows, int cols, bool want_lo, bool is_const, const char* slot,
                const bf16** hi, const bf16** lo, int64_t* ldo) {
  Packed2D pk;
  RAU_TRY(pack2d(ctx, src, ld, rows, cols, want_lo, is_const, slot, &pk));
  *hi = pk.hi; *lo = pk.lo; *ldo = pk.ld;
  return RAU_OK;
}
namespace {
// LSTM layer weights for the fused cell epilogue: rows permuted so that packed row n' = (u/8)*32 + k*8 + u%8 is gate k
// (k = 0..3 = i, f, o, g) of hidden unit u; bsum[n'] = bi + bh in the same order
__global__ void pack_lstm_kernel(const float* __restrict__ W, int H, int K, int ldo, int c_i, int c_f, int c_o, int c_g,
                                 bf16* __restrict__ hi, bf16* __restrict__ lo) {
  RAU_PDL_ENTRY();
  const int q = ldo >> 1;
  const int64_t total = (int64_t)4 * H * q;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % q) * 2;
    const int np = (int)(i / q);
    const int grp = np >> 5, k = (np & 31) >> 3, uu = np & 7;
    const int chunk = k == 0 ? c_i : (k == 1 ? c_f : (k == 2 ? c_o : c_g));
    const int64_t src = (int64_t)(chunk * H + grp * 8 + uu) * K;
    const float a = c < K ? W[src + c] : 0.0f;
    const float b = c + 1 < K ? W[src + c + 1] : 0.0f;
    uint32_t h, l;
    split_pair(a, b, h, l);
    reinterpret_cast<uint32_t*>(hi)[i] = h;
    if (lo) reinterpret_cast<uint32_t*>(lo)[i] = l;
  }
}
__global__ void perm_bias_kernel(const float* __restrict__ b1, const float* __restrict__ b2, int H, int c_i, int c_f, int c_o,
                                 int c_g, float* __restrict__ out) {
  RAU_PDL_ENTRY();
  const int np = blockIdx.x * blockDim.x + threadIdx.x;
  if (np >= 4 * H) return;
  const int grp = np >> 5, k = (np & 31) >> 3, uu = np & 7;
  const int chunk = k == 0 ? c_i : (k == 1 ? c_f : (k == 2 ? c_o : c_g));
  const int src = chunk * H + grp * 8 + uu;
  out[np] = (b1 ? b1[src] : 0.0f) + (b2 ? b2[src] : 0.0f);
}
}  // namespace

int rows_pack_lstm(rau_ctx* ctx, const float* W, int H, int K, int gate_order, bool want_lo, const bf16** hi, const bf16** lo,
                   int64_t* ldo) {
  RAU_REQUIRE(H % 8 == 0, "rows_pack_lstm: H = %d must be a multiple of 8", H);
  int ci = 0, cf = 1, co = 2, cg = 3;                         // RAU_GATES_IFOG (D:47-54)
  if (gate_order == RAU_GATES_IGFO) { ci = 0; cg = 1; cf = 2; co = 3; }   // A:12-19
  const int ld = (K + 7) / 8 * 8;
  char name[128];
  snprintf(name, sizeof(name), "rl.%p.%d.%d.%d.%d", (const void*)W, H, K, gate_order, want_lo ? 1 : 0);
  auto it = ctx->tc_epoch.find(name);
  const bool cached = it != ctx->tc_epoch.end() && it->second == ctx->epoch;
  const size_t half = ((size_t)4 * H * ld * sizeof(bf16) + 1023) / 1024 * 1024;
  void* buf = nullptr;
  RAU_TRY(ctx->arena.get(name, half * (want_lo ? 2 : 1), &buf));
  *hi = (bf16*)buf;
  *lo = want_lo ? (bf16*)((char*)buf + half) : nullptr;
  *ldo = ld;
  if (!cached) {
    const int64_t work = (int64_t)4 * H * (ld / 2);
    int64_t blocks = (work + 255) / 256;
    if (blocks > 148 * 8) blocks = 148 * 8;
    RAU_LAUNCH_PDL(ctx->stream, (pack_lstm_kernel), (int)blocks, 256, 0, W, H, K, ld, ci, cf, co, cg, (bf16*)*hi, (bf16*)*lo);
    RAU_LAUNCH_CHECK(ctx);
    ctx->tc_epoch[name] = ctx->epoch;
  }
  return RAU_OK;
}

int rows_perm_lstm_bias(rau_ctx* ctx, const float* b1, const float* b2, int H, int gate_order, const float** out) {
  int ci = 0, cf = 1, co = 2, cg = 3;
  if (gate_order == RAU_GATES_IGFO) { ci = 0; cg = 1; cf = 2; co = 3; }
  char name[128];
  snprintf(name, sizeof(name), "rlb.%p.%p.%d.%d", (const void*)b1, (const void*)b2, H, gate_order);
  auto it = ctx->tc_epoch.find(name);
  const bool cached = it != ctx->tc_epoch.end() && it->second == ctx->epoch;
  void* buf = nullptr;
  RAU_TRY(ctx->arena.get(name, sizeof(float) * 4 * (size_t)H, &buf));
  *out = (const float*)buf;
  if (!cached) {
    RAU_LAUNCH_PDL(ctx->stream, (perm_bias_kernel), (4 * H + 255) / 256, 256, 0, b1, b2, H, ci, cf, co, cg, (float*)buf);
    RAU_LAUNCH_CHECK(ctx);
    ctx->tc_epoch[name] = ctx->epoch;
  }
  return RAU_OK;
}

namespace {
// C[m, n] = bias[n] + bias2[n] + addend[m, n] + addend2[m, n]: what a split-K product then adds its partial sums into
__global__ void linear_init_kernel(float* __restrict__ C, long long ldc, int M, int N, const float* __restrict__ bias,
                                   const float* __restrict__ bias2, const float* __restrict__ addend,
                                   const float* __restrict__ addend2, long long ldadd) {
  RAU_PDL_ENTRY();
  const long long total = (long long)M * N;
  for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int n = (int)(i % N);
    const long long m = i / N;
    float t = 0.0f;
    if (bias) t += bias[n];
    if (bias2) t += bias2[n];
    if (addend) t += addend[m * ldadd + n];
    if (addend2) t += addend2[m * ldadd + n];
    C[m * ldc + n] = t;
  }
}

long long rows_min_work() { return rau_process_tuning().tc_min_work; }
}  // namespace

int rows_contract_try(rau_ctx* ctx, const SimtGemm& g) {
  if (ctx->precision == RAU_PREC_F32 || !rows_path_enabled()) return 0;
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return 0;
  if (g.batch != 1 || g.kbatch != 1) return 0;
  if (g.bias_m || g.bias_bm || g.n_valid >= 0) return 0;
  if (g.A_hi || g.B_hi) return 0;                     // producers of the generic engine's packed forms
  if ((long long)g.M * g.N * (long long)(g.K + g.K2) < rows_min_work()) return 0;
  // output: fp32 row-major, TMA-addressable
  if (g.scn != 1 || g.scm % 4 != 0 || ((uintptr_t)g.C & 15) != 0) return 0;
  if (g.C_hi && (g.scm % 8 != 0 || (((uintptr_t)g.C_hi | (uintptr_t)g.C_lo) & 15) != 0)) return 0;
  const bool plain_acc = g.accumulate != 0;
  if (plain_acc && (g.bias_n || g.bias_n2 || g.addend || g.addend2 || g.act != 0 || g.C_hi)) return 0;
  if (!plain_acc && g.ksplit > 1) return 0;
  if (g.addend && g.sdn != 1) return 0;
  // operands: one contiguous dimension each
  const int a_mn = (g.sak == 1) ? 0 : (g.sam == 1 ? 1 : -1);
  const int b_mn = (g.sbk == 1) ? 0 : (g.sbn == 1 ? 1 : -1);
  if (a_mn < 0 || b_mn < 0) return 0;
  const bool x3 = prec_x3(ctx);
  RowsGemm r;
  r.M = g.M; r.N = g.N; r.K = g.K;
  Packed2D pa, pb, pa2, pb2;
  // K-major: stored [rows, K]; MN-major: stored [K, rows].  A producer may have left a packed twin of the operand.
  auto twin_ok = [&](const bf16* hi, const bf16* lo, int64_t ld) {
    return hi != nullptr && (!x3 || lo != nullptr) && ld % 8 == 0 && (((uintptr_t)hi | (uintptr_t)lo) & 15) == 0;
  };
  if (twin_ok(g.Ar_hi, g.Ar_lo, g.Ar_ld)) { pa.hi = g.Ar_hi; pa.lo = x3 ? g.Ar_lo : nullptr; pa.ld = g.Ar_ld; }
  else RAU_TRY(pack2d(ctx, g.A, a_mn ? g.sak : g.sam, a_mn ? g.K : g.M, a_mn ? g.M : g.K, x3, g.a_const != 0, "A", &pa));
  if (twin_ok(g.Br_hi, g.Br_lo, g.Br_ld)) { pb.hi = g.Br_hi; pb.lo = x3 ? g.Br_lo : nullptr; pb.ld = g.Br_ld; }
  else RAU_TRY(pack2d(ctx, g.B, b_mn ? g.sbk : g.sbn, b_mn ? g.K : g.N, b_mn ? g.N : g.K, x3, g.b_const != 0, "B", &pb));
  r.A.hi = pa.hi; r.A.lo = pa.lo; r.A.mn = a_mn; r.A.ld = pa.ld;
  r.B.hi = pb.hi; r.B.lo = pb.lo; r.B.mn = b_mn; r.B.ld = pb.ld;
  if (g.A2) {
    const int a2_mn = (g.sak2 == 1) ? 0 : (g.sam2 == 1 ? 1 : -1);
    const int b2_mn = (g.sbk2 == 1) ? 0 : (g.sbn2 == 1 ? 1 : -1);
    if (a2_mn != a_mn || b2_mn != b_mn) return 0;
    if (twin_ok(g.A2r_hi, g.A2r_lo, g.A2r_ld)) { pa2.hi = g.A2r_hi; pa2.lo = x3 ? g.A2r_lo : nullptr; pa2.ld = g.A2r_ld; }
    else RAU_TRY(pack2d(ctx, g.A2, a_mn ? g.sak2 : g.sam2, a_mn ? g.K2 : g.M, a_mn ? g.M : g.K2, x3, g.a_const != 0, "A2", &pa2));
    RAU_TRY(pack2d(ctx, g.B2, b_mn ? g.sbk2 : g.sbn2, b_mn ? g.K2 : g.N, b_mn ? g.N : g.K2, x3, g.b_const != 0, "B2", &pb2));
    r.A2.hi = pa2.hi; r.A2.lo = pa2.lo; r.A2.mn = a_mn; r.A2.ld = pa2.ld;
    r.B2.hi = pb2.hi; r.B2.lo = pb2.lo; r.B2.mn = b_mn; r.B2.ld = pb2.ld;
    r.K2 = g.K2;
  }
  r.alpha = g.alpha;
  r.out_f = g.C; r.ldo = g.scm;
  // long reductions into a few output tiles (dgrads with K = 4H, the 2000-way head): one CTA per tile would stream its
  // whole K range through a single SM's ~64 B/clk L2 port, so the K range is split over the SMs and the partial sums are
  // added (TMA reduce) into an output that a tiny kernel pre-loads with bias + addend
  const int skinny_tiles = ((g.M + RT_BM - 1) / RT_BM) * ((g.N + 63) / 64);
  const bool split = !plain_acc && g.act == 0 && g.C_hi == nullptr && (g.K + g.K2) >= 1024 && skinny_tiles * 4 <= ctx->sm_count;
  if (split) {
    if (g.bias_n || g.bias_n2 || g.addend || g.addend2) {
      const long long total = (long long)g.M * g.N;
      int blocks = (int)((total + 1023) / 1024);
      if (blocks > 148 * 4) blocks = 148 * 4;
      RAU_LAUNCH_PDL(ctx->stream, (linear_init_kernel), blocks, 256, 0, g.C, g.scm, g.M, g.N, g.bias_n, g.bias_n2, g.addend, g.addend2, g.sdm);
      RAU_LAUNCH_CHECK(ctx);
    } else {
      RAU_CHECK_CUDA(cudaMemset2DAsync(g.C, (size_t)g.scm * 4, 0, (size_t)g.N * 4, (size_t)g.M, ctx->stream));
    }
    r.epi = ROWS_EPI_RED;
  } else if (plain_acc) {
    r.epi = ROWS_EPI_RED;
  } else {
    r.epi = ROWS_EPI_LINEAR;
    r.bias = g.bias_n; r.bias2 = g.bias_n2;
    r.addend = g.addend; r.addend2 = g.addend2; r.ldadd = g.sdm;
    r.act = g.act;
    r.out_hi = g.C_hi; r.out_lo = x3 ? g.C_lo : nullptr; r.ldo_b = g.scm;
  }
  RAU_TRY(rows_gemm(ctx, r));
  return 1;
}
