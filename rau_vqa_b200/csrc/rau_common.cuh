// rau_common.cuh -- context, error plumbing and small device helpers shared by every kernel file.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include <map>
#include <string>
#include <functional>
#include "../../include/rau.h"

typedef __nv_bfloat16 bf16;

void rau_set_error(const char* fmt, ...);

#define RAU_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      rau_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return RAU_ECUDA;                                                                   \
    }                                                                                     \
  } while (0)

#define RAU_REQUIRE(cond, ...)                                                            \
  do {                                                                                    \
    if (!(cond)) {                                                                        \
      rau_set_error(__VA_ARGS__);                                                         \
      return RAU_EINVAL;                                                                  \
    }                                                                                     \
  } while (0)

#define RAU_TRY(expr)                                                                     \
  do {                                                                                    \
    int _s = (expr);                                                                      \
    if (_s != RAU_OK) return _s;                                                          \
  } while (0)

// A growable device arena: buffers are keyed by name, re-used across calls, grown on demand.
struct RauArena {
  struct Buf { void* p = nullptr; size_t bytes = 0; };
  std::map<std::string, Buf> bufs;
  uint64_t generation = 0;   // bumped whenever a buffer is re-allocated: device addresses a captured graph baked in are stale
  int get(const char* name, size_t bytes, void** out);
  void release();
};

// Every environment switch of the library, read in ONE place (rau_tuning_from_env, rau_ctx.cu) when a context is created.
// The defaults are the measured best; the switches exist for A/B runs, tests of the alternative schedules and debugging.
struct RauTuning {
  int graph = 1;          // RAU_GRAPH=0: never capture the step into a CUDA graph
  int overlap = 7;        // RAU_OVERLAP: bit 0 backward products, bit 1 forward pre-work, bit 2 answer heads on the side stream
  int side_ctas = 0;      // RAU_SIDE_CTAS: SMs the side stream's launches size themselves for (0 = 4/7 of the device)
  int side_ctas_fwd = 0;  // RAU_SIDE_CTAS_FWD / _BWD: the same for the forward pre-work / the hops' backward products
  int side_ctas_bwd = 0;
  int main_ctas = 0;      // RAU_MAIN_CTAS: SMs the chain's split-K products size for while the side stream runs (0 = the rest)
  int rows = 1;           // RAU_ROWS=0: route products to the first-cut tcgen05 engine
  int cg2 = 1;            // RAU_CG2=0: no CTA pairs
  int lin_cg2 = 1;        // RAU_LIN_CG2=0: the big-M nn.Linear products (encoder projections) stay on single CTAs
  int enc_w0 = 1;         // RAU_ENC_W0=0: layer 1's input projection waits for ALL of the aux stream's encoder preparation
  int tanh_ew = 16;       // RAU_TANH_EW=8: eight epilogue warps for the 1-pass i_embed product
  int rows_trace = 0;     // RAU_ROWS_TRACE=1: per-CTA clock stamps of rows-engine launches (tools/rows_trace.py)
  int lstm_seq = 1;       // RAU_LSTM_SEQ=0: one launch per recurrent step of the encoder instead of the persistent kernel
  int enc_bwd_wave = 1;   // RAU_ENC_BWD_WAVE=0: encoder backward layer after layer instead of the two-stream wavefront
  int xprep_hops = 1;     // RAU_XPREP_HOPS=0: one feature-pack launch per hop instead of one for all hops
  int pdl = 0;            // RAU_PDL=1: programmatic dependent launch attribute on every launch (measured slower)
  int phases = 0;         // RAU_PHASES: 1 = per-phase events of an eager step, 2 = %globaltimer stamps inside the graph
  long long tc_min_work = 1ll << 18;   // RAU_TC_MIN_WORK: M*N*K below which a product stays on the CUDA-core engine
  int time_cap = 0;       // RAU_TIME_CAP: SM cap for rau_rows_gemm_time
};
RauTuning rau_tuning_from_env();
const RauTuning& rau_process_tuning();   // read once per process: for the few call sites that have no context at hand

struct RauComm;  // rau_comm.cu

// Per-step scalars kept on the device so that a captured CUDA graph of the training step can be replayed
// unchanged: the host uploads this struct before every step instead of baking the values into kernel arguments.
struct StepState {
  unsigned long long step;   // iteration number `it` (0-based): keys the Philox dropout / noise streams
  float noise_std;           // sqrt(eta / ((it+1) * gamma))                                   F:617-618
  float opt_step[3];         // adam: lr * sqrt(1 - b2^t) / (1 - b1^t) per group (OU:80-83); other rules: lr
};

struct RauGraphEntry { cudaGraphExec_t exec = nullptr; int seen = 0; int64_t kernels = 0; };
struct RauGraph {
  std::map<std::vector<uint64_t>, RauGraphEntry> entries;   // one captured step per distinct argument set
  bool disabled = false;
  uint64_t arena_generation = 0;   // RauArena::generation the entries were captured under
  void clear() {
    for (auto& kv : entries)
      if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    entries.clear();
  }
};

struct rau_ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  uint64_t seed = 0x5eed5eedULL;
  int precision = RAU_PREC_MIXED;
  RauTuning tune;
  int sm_count = 148;
  int64_t launches = 0;
  RauArena arena;
  uint64_t epoch = 1;                              // bumped at every public entry: bf16 weight shadows are per epoch
  std::map<std::string, uint64_t> tc_epoch;        // shadow name -> epoch it was packed in
  RauComm* comm = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  StepState* d_ss = nullptr;                       // device copy read by kernels
  StepState* h_ss = nullptr;                       // pinned ring the uploads are staged in
  int ss_slot = 0;
  cudaEvent_t ss_ev[64] = {};                      // recorded behind each slot's upload: a slot is reused only after its copy ran
  // sticky failure word of the persistent recurrence kernels (a CTA gave up waiting for its row tile's peers): d_err is
  // read by the optimizer kernel, which then leaves the parameters untouched; the kernels also set the host-mapped copy,
  // which every public entry point checks without a device synchronisation
  unsigned int* d_err = nullptr;
  volatile unsigned int* h_err = nullptr;          // cudaHostAllocMapped
  unsigned int* h_err_dev = nullptr;               // its device alias
  const StepState* ss_active = nullptr;            // non-null while a whole-step call is being enqueued
  cudaStream_t gstream = nullptr;                  // capture stream (the caller's stream may be the legacy one)
  RauGraph graph;
  // side stream for the heavy image-side products that are off the recurrence's critical path (rau_step.cu): they run
  // on at most side_ctas SMs next to the chain of small dependent kernels.  RAU_OVERLAP=0 disables it.
  cudaStream_t side = nullptr;
  cudaStream_t aux = nullptr;                      // short state-independent preparation (fills, masks) next to the encoder
  cudaStream_t aux2 = nullptr;                     // the encoder backward's second wavefront lane (layer 1 behind layer 2)
  std::vector<cudaEvent_t> side_ev;
  int side_ev_next = 0;
  int side_ctas = 0;
  int side_ctas_bwd = 0;                           // the cap for the hops' backward products (dY, gWa, gWi)
  int side_ctas_fwd = 0;                           // the cap while the forward's state-independent products run
  // the answering units' gradient group is complete once the side stream has run the deferred weight gradients, long
  // before the encoder backward ends: the rest of that group's step (all-reduce when data parallel, gradient noise + norm,
  // clip + optimizer) is issued there on the aux stream (train step only)
  std::function<int()> early_tail;          // set by the train step: [all-reduce,] noise + norm, clip + optimizer of group 2
  cudaEvent_t early_tail_done = nullptr;    // recorded on the aux stream behind it (NULL: it did not run)
  // the word-embedding group is final when the chain has run the encoder backward, before it waits for the side stream's
  // weight gradients: its [all-reduce,] noise + norm, clip + optimizer go out on the chain in that gap
  std::function<int()> early_tail0;
  bool early_tail0_ran = false;
  int main_cta_cap = 0;                            // > 0 while the side stream is in use: SMs the chain's split-K products size for
  int rows_cta_cap = 0;                            // > 0 while work is being enqueued on the side stream
  // RAU_PHASES=1: eager steps with an event at every phase boundary; rau_phase_report() prints the split
  int phases = -1;
  std::vector<std::pair<std::string, cudaEvent_t>> phase_ev;
  // RAU_PHASES=2: a one-thread kernel per mark writes %globaltimer, so the marks survive graph capture and show the
  // replayed step's timeline on both streams
  unsigned long long* stamp_buf = nullptr;
  std::vector<std::string> stamp_names;
};
void rau_phase_mark(rau_ctx* ctx, const char* name);

// precision mode -> operand formats.
//   prec_x3       products on the recurrent chain / encoder / small nn.Linear layers carry bf16 (hi, lo) operands (3 passes)
//   prec_x_f16    the dropped-out features Xd, the Wi shadow and dY are single fp16 planes: I = tanh(Wi Xd + bi) and
//                 gWi += dY^T Xd are one fp16 pass (RAU_PREC_MIXED and RAU_PREC_F16IMG)
//   prec_img_f16  I, dZ and the Wa shadow are single fp16 planes too: Z, dY, gWa are one fp16 pass (RAU_PREC_F16IMG only)
static inline bool prec_x3(const rau_ctx* c) {
  return c->precision == RAU_PREC_BF16X3 || c->precision == RAU_PREC_MIXED || c->precision == RAU_PREC_F16IMG;
}
static inline bool prec_x_f16(const rau_ctx* c) { return c->precision == RAU_PREC_MIXED || c->precision == RAU_PREC_F16IMG; }
static inline bool prec_img_f16(const rau_ctx* c) { return c->precision == RAU_PREC_F16IMG; }

#define RAU_LAUNCH_CHECK(ctx)                                                             \
  do {                                                                                    \
    (ctx)->launches++;                                                                    \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) {                                                              \
      rau_set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
      return RAU_ECUDA;                                                                   \
    }                                                                                     \
  } while (0)

static inline int cdiv(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------ programmatic dependent launch
// The step is a chain of ~700 short dependent kernels.  Every kernel of the library starts with RAU_PDL_ENTRY(): it lets
// the NEXT kernel of the stream be scheduled early (griddepcontrol.launch_dependents) and then waits until everything
// the PREVIOUS kernels wrote is visible (griddepcontrol.wait) before touching memory.  Launches carry the
// programmatic-stream-serialization attribute, so launch latency, CTA scheduling, instruction fetch and the tcgen05
// prologue (barrier init, TMEM allocation, tensor-map prefetch) of kernel i+1 overlap the tail of kernel i; data
// dependencies stay exactly those of a serial stream.  The attribute is OFF by default (RAU_PDL=1 turns it on): inside
// the replayed CUDA graph of the step it measured slightly slower than plain serial edges (profiles/README.md).
#define RAU_PDL_ENTRY()                                              \
  do {                                                               \
    asm volatile("griddepcontrol.launch_dependents;");               \
    asm volatile("griddepcontrol.wait;" ::: "memory");               \
  } while (0)

bool rau_pdl_enabled();

template <typename... Exp, typename... Act>
static inline cudaError_t rau_launch_pdl(cudaStream_t st, void (*kern)(Exp...), dim3 grid, dim3 block, size_t smem, Act&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = rau_pdl_enabled() ? 1 : 0;
  cfg.attrs = at;
  cfg.numAttrs = 1;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<Act&&>(args)...);
}
#define RAU_LAUNCH_PDL(stream, kern, grid, block, smem, ...) \
  (void)rau_launch_pdl(stream, kern, dim3(grid), dim3(block), (size_t)(smem), __VA_ARGS__)

// ------------------------------------------------------------------ device helpers
__device__ __forceinline__ float sigmoidf_(float x) { return 1.0f / (1.0f + __expf(-x)); }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// block-wide sum for blockDim.x <= 1024 (multiple of 32); `red` is >= 32 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
  v = warp_sum(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : 0.0f;
  if (w == 0) r = warp_sum(r);
  if (threadIdx.x == 0) red[0] = r;
  __syncthreads();
  r = red[0];
  return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  v = warp_max(v);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (threadIdx.x < nw) ? red[threadIdx.x] : -INFINITY;
  if (w == 0) r = warp_max(r);
  if (threadIdx.x == 0) red[0] = r;
  __syncthreads();
  r = red[0];
  return r;
}

// Philox4x32-10 (Salmon et al. 2011), counter-based: no state to carry between forward and backward.
__device__ __forceinline__ uint4 philox4x32(uint4 ctr, uint2 key) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0 = __umulhi(M0, ctr.x), lo0 = M0 * ctr.x;
    uint32_t hi1 = __umulhi(M1, ctr.z), lo1 = M1 * ctr.z;
    ctr = make_uint4(hi1 ^ ctr.y ^ key.x, lo1, hi0 ^ ctr.w ^ key.y, lo0);
    key.x += W0; key.y += W1;
  }
  return ctr;
}

// keep-bit lookup in a packed mask (bit i of word i>>5); nullptr = keep everything (evaluate()).
// Feature-dropout masks drawn inline (F:239, one mask per hop over the same features).  General p: element (b, c, s) of hop h
// compares 16 bits of the draw (counter = the element's channel-pair x 4-cell group, stream id ^ h) with a threshold.  For
// p = 1/2 exactly and nHop <= 16 ONE draw serves every hop: hop h keeps the element iff bit h of its 16-bit lane is clear
// (stream id ^ RAU_XMASK_SHARED_TAG, independent of h) -- 1/nHop of the Philox work, which is what bounds the pack kernel.
constexpr uint32_t RAU_XMASK_SHARED_TAG = 0xA5A50000u;
__host__ __device__ __forceinline__ bool rau_xmask_shared(uint32_t thresh16, int nHop) { return thresh16 == 32768u && nHop >= 1 && nHop <= 16; }

__device__ __forceinline__ float keep_scale(const uint32_t* __restrict__ bits, int64_t i, float scale) {
  if (bits == nullptr) return 1.0f;
  return ((bits[i >> 5] >> (i & 31)) & 1u) ? scale : 0.0f;
}

// ------------------------------------------------------------------ internal entry points
// k_gemm_simt.cu
struct SimtGemm {
  int M = 0, N = 0, K = 0;
  const float* A = nullptr; int64_t sam = 0, sak = 0;   // A(m,k) = A[m*sam + k*sak]
  const float* B = nullptr; int64_t sbk = 0, sbn = 0;   // B(k,n) = B[k*sbk + n*sbn]
  float* C = nullptr; int64_t scm = 0, scn = 1;
  int batch = 1; int64_t bA = 0, bB = 0, bC = 0;        // independent problems (blockIdx.z)
  int kbatch = 1; int64_t kA = 0, kB = 0;               // batches reduced into one C
  int ksplit = 1;                                       // >1: grid.z slices of the reduction, atomicAdd into C
  const float* A2 = nullptr; int64_t sam2 = 0, sak2 = 0; // optional second K segment (same M,N)
  const float* B2 = nullptr; int64_t sbk2 = 0, sbn2 = 0; int K2 = 0;
  const float* bias_m = nullptr;   // [M]
  const float* bias_n = nullptr;   // [N]
  const float* bias_n2 = nullptr;  // [N] second bias (i2h + h2h)
  const float* bias_bm = nullptr;  // [batch, M]
  const float* addend = nullptr; int64_t sdm = 0, sdn = 1, bD = 0;  // same indexing as C
  const float* addend2 = nullptr;  // same strides as addend
  int a_const = 0, b_const = 0;    // operand is a parameter tensor (constant within one public call)
  // tcgen05 engine only: operands a producer already wrote as packed bf16 (hi, lo) with the fp32 operand's own
  // indexing, and optional bf16 (hi, lo) copies of the result indexed like C
  const bf16 *A_hi = nullptr, *A_lo = nullptr, *B_hi = nullptr, *B_lo = nullptr;
  bf16 *C_hi = nullptr, *C_lo = nullptr;
  // rows engine only: packed bf16 (hi, lo) twins of the operands AS STORED ([M or K rows, pitch ld], ld % 8 == 0), written
  // by the producing kernel; with them the product needs no pack launch
  const bf16 *Ar_hi = nullptr, *Ar_lo = nullptr; int64_t Ar_ld = 0;
  const bf16 *A2r_hi = nullptr, *A2r_lo = nullptr; int64_t A2r_ld = 0;
  const bf16 *Br_hi = nullptr, *Br_lo = nullptr; int64_t Br_ld = 0;
  int accumulate = 0;
  float alpha = 1.0f;
  int act = 0;                     // 0 none, 1 tanh, 2 sigmoid; 3 (rows engine only) tanh backward: * (1 - addend2^2)
  int n_valid = -1;                // columns >= n_valid are written as 0 (spatial pad)
};
int simt_gemm(rau_ctx* ctx, const SimtGemm& g);
