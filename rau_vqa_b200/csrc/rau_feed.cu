// rau_feed.cu -- the batch feed.  The reference moves every batch synchronously: next_batch_feat
// returns host tensors feats[B,C,14,14] (float64), x[T,B], x_len[B], y[B] (LD:1009) and feval casts and uploads them on
// the compute stream, B*C*196*4 bytes per step (F:452-456: 103 MB at B = 256, C = 512).  Two replacements:
//   rau_feed        pinned, `depth`-deep staging; the upload of batch i+1 runs on a copy stream under step i.  The staging
//                   format is float32 (as the reference uploads) or float16 (half the PCIe bytes: with eight ranks pulling
//                   from one host the float32 feed is upload-bound).  In the default precision mode the image features enter
//                   the tensor pipe as fp16 anyway (x -> fp16(x / (1-p)), and the power-of-two dropout scale commutes with
//                   the rounding), so the fp16 feed changes no bit of what the products read.
//   rau_feat_cache  the features of a whole split resident in HBM as fp16 (train2014 at C = 512: 82 783 x 196 KB = 16 GB of the
//                   180 GB), a batch is a gather by image index: no per-step feature upload at all.
#include "rau_kernels.cuh"
#include <cuda_fp16.h>
#include <thread>

namespace {
__global__ void half_to_float_kernel(const __half2* __restrict__ in, int64_t n2, float2* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __half22float2(in[i]);
}
__global__ void float_to_half_kernel(const float2* __restrict__ in, int64_t n2, __half2* __restrict__ out) {
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n2; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = __float22half2_rn(in[i]);
}
// feats[b] = float(cache[index[b] - 1]) ; one image = per_image halves (even); grid.y = B
__global__ void gather_kernel(const __half2* __restrict__ cache, int64_t per2, int64_t n_images, const float* __restrict__ index,
                              float2* __restrict__ out) {
  const int b = blockIdx.y;
  int64_t img = (int64_t)index[b] - 1;
  if (img < 0) img = 0;
  if (img >= n_images) img = n_images - 1;
  const __half2* src = cache + img * per2;
  float2* dst = out + (int64_t)b * per2;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < per2; i += (int64_t)gridDim.x * blockDim.x)
    dst[i] = __half22float2(src[i]);
}
// the same gather without widening: out[b] = cache[index[b] - 1] as fp16 (16 bytes per thread and step)
__global__ void gather_f16_kernel(const uint4* __restrict__ cache, int64_t per8, int64_t n_images, const float* __restrict__ index,
                                  uint4* __restrict__ out) {
  const int b = blockIdx.y;
  int64_t img = (int64_t)index[b] - 1;
  if (img < 0) img = 0;
  if (img >= n_images) img = n_images - 1;
  const uint4* src = cache + img * per8;
  uint4* dst = out + (int64_t)b * per8;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < per8; i += (int64_t)gridDim.x * blockDim.x) dst[i] = __ldg(src + i);
}
inline int blocks_for(int64_t n) { int64_t b = (n + 255) / 256; return (int)(b < 1 ? 1 : (b > 148 * 16 ? 148 * 16 : b)); }
}  // namespace

struct rau_feed {
  rau_ctx* ctx = nullptr;
  int B = 0, T = 0, C = 0, S = 0, format = 0, depth = 0;
  cudaStream_t copy = nullptr;
  struct Slot {
    void* h_feats = nullptr; float *h_tok = nullptr, *h_len = nullptr, *h_lab = nullptr;     // pinned host staging
    void* d_stage = nullptr;                                                                // device fp16 staging (F16 only)
    float *d_feats = nullptr, *d_tok = nullptr, *d_len = nullptr, *d_lab = nullptr;          // what rau_batch points at
    cudaEvent_t copied = nullptr, ready = nullptr, freed = nullptr;
    int max_len = 0;
    bool in_flight = false;
  };
  std::vector<Slot> slots;
};

struct rau_feat_cache {
  rau_ctx* ctx = nullptr;
  int64_t n_images = 0, per_image = 0;
  __half* data = nullptr;
  float* stage = nullptr;        // device fp32 staging of one upload chunk
  int64_t stage_elems = 0;
};

int rau_check_cfg(const rau_config* cfg);
int rau_check_dev(const void* p, const char* what);

extern "C" {

size_t rau_feed_host_bytes(const rau_feed* f) {
  if (f == nullptr) return 0;
  const size_t feat = (size_t)f->B * f->C * f->S * (f->format == RAU_FEED_F32 ? 4 : 2);
  return feat + sizeof(float) * ((size_t)f->T * f->B + 2 * (size_t)f->B);
}

int rau_feed_destroy(rau_feed* f) {
  if (f == nullptr) return RAU_OK;
  cudaSetDevice(f->ctx->device);
  if (f->copy) cudaStreamSynchronize(f->copy);
  cudaStreamSynchronize(f->ctx->stream);
  for (auto& s : f->slots) {
    if (s.h_feats) cudaFreeHost(s.h_feats);
    if (s.h_tok) cudaFreeHost(s.h_tok);
    if (s.d_stage) cudaFree(s.d_stage);
    if (s.d_feats) cudaFree(s.d_feats);
    if (s.d_tok) cudaFree(s.d_tok);
    if (s.copied) cudaEventDestroy(s.copied);
    if (s.ready) cudaEventDestroy(s.ready);
    if (s.freed) cudaEventDestroy(s.freed);
  }
  if (f->copy) cudaStreamDestroy(f->copy);
  delete f;
  return RAU_OK;
}

int rau_feed_create(rau_ctx* ctx, const rau_config* cfg, int B, int format, int depth, rau_feed** out) {
  RAU_REQUIRE(ctx && out, "ctx/out == NULL");
  *out = nullptr;
  RAU_TRY(rau_check_cfg(cfg));
  RAU_REQUIRE(B > 0 && depth >= 2 && depth <= 8, "rau_feed_create: B = %d, depth = %d (2..8)", B, depth);
  RAU_REQUIRE(format == RAU_FEED_F32 || format == RAU_FEED_F16 || format == RAU_FEED_F16_DIRECT, "rau_feed_create: unknown format %d", format);
  RAU_REQUIRE(((int64_t)cfg->C * cfg->S) % 2 == 0, "rau_feed_create: C*S must be even");
  RAU_CHECK_CUDA(cudaSetDevice(ctx->device));
  rau_feed* f = new rau_feed();
  f->ctx = ctx; f->B = B; f->T = cfg->T; f->C = cfg->C; f->S = cfg->S; f->format = format; f->depth = depth;
  f->slots.resize(depth);
  const size_t nfeat = (size_t)B * cfg->C * cfg->S, nsmall = (size_t)cfg->T * B + 2 * (size_t)B;
  bool ok = cudaStreamCreateWithFlags(&f->copy, cudaStreamNonBlocking) == cudaSuccess;
  for (auto& s : f->slots) {
    if (!ok) break;
    ok = cudaHostAlloc(&s.h_feats, nfeat * (format == RAU_FEED_F32 ? 4 : 2), cudaHostAllocDefault) == cudaSuccess &&
         cudaHostAlloc((void**)&s.h_tok, nsmall * sizeof(float), cudaHostAllocDefault) == cudaSuccess &&
         (format == RAU_FEED_F16_DIRECT || cudaMalloc((void**)&s.d_feats, nfeat * sizeof(float)) == cudaSuccess) &&
         cudaMalloc((void**)&s.d_tok, nsmall * sizeof(float)) == cudaSuccess &&
         (format == RAU_FEED_F32 || cudaMalloc(&s.d_stage, nfeat * 2) == cudaSuccess) &&
         cudaEventCreateWithFlags(&s.copied, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&s.ready, cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&s.freed, cudaEventDisableTiming) == cudaSuccess;
    if (ok) {
      s.h_len = s.h_tok + (size_t)cfg->T * B; s.h_lab = s.h_len + B;
      s.d_len = s.d_tok + (size_t)cfg->T * B; s.d_lab = s.d_len + B;
    }
  }
  if (!ok) {
    rau_set_error("rau_feed_create: allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    rau_feed_destroy(f);
    return RAU_ENOMEM;
  }
  *out = f;
  return RAU_OK;
}

// The pinned staging of one slot, for the loader to fill: feats in the feed's format ([B,C,14,14] float32 or float16),
// tokens [T,B], lengths [B], labels [B] as float, 1-based (F:454-455).  Blocks until the slot's previous upload has left
// the staging memory.
int rau_feed_host_slot(rau_feed* f, int slot, void** feats, float** tokens, float** lengths, float** labels) {
  RAU_REQUIRE(f && slot >= 0 && slot < f->depth, "rau_feed_host_slot: bad feed / slot");
  rau_feed::Slot& s = f->slots[slot];
  if (s.in_flight) RAU_CHECK_CUDA(cudaEventSynchronize(s.copied));
  if (feats) *feats = s.h_feats;
  if (tokens) *tokens = s.h_tok;
  if (lengths) *lengths = s.h_len;
  if (labels) *labels = s.h_lab;
  return RAU_OK;
}

// The host-side cast of F:452-456 (`feats:float()`): n loader values (float64 when src_is_f64, else float32) into the
// feed's staging format, split over a few host threads (the loader's prefetch thread calls this, LD:931-958).
int rau_feed_convert(const rau_feed* f, const void* src, int src_is_f64, int64_t n, void* dst) {
  RAU_REQUIRE(f && src && dst && n >= 0, "rau_feed_convert: bad arguments");
  const int fmt = f->format;
  auto work = [=](int64_t a, int64_t b) {
    for (int64_t i = a; i < b; ++i) {
      const float v = src_is_f64 ? (float)((const double*)src)[i] : ((const float*)src)[i];
      if (fmt != RAU_FEED_F32) ((__half*)dst)[i] = __float2half_rn(v);
      else ((float*)dst)[i] = v;
    }
  };
  unsigned hw = std::thread::hardware_concurrency();
  int nt = (int)(hw == 0 ? 1 : (hw > 8 ? 8 : hw));
  if (n < (1 << 16)) nt = 1;
  std::vector<std::thread> th;
  for (int t = 1; t < nt; ++t) th.emplace_back(work, n * t / nt, n * (t + 1) / nt);
  work(0, n / nt);
  for (auto& t : th) t.join();
  return RAU_OK;
}

// Enqueue the upload of a filled slot on the feed's copy stream and return at once.  The device buffers of the slot are
// rewritten only after the step that last read them has run (rau_feed_release).
int rau_feed_submit(rau_feed* f, int slot) {
  RAU_REQUIRE(f && slot >= 0 && slot < f->depth, "rau_feed_submit: bad feed / slot");
  rau_feed::Slot& s = f->slots[slot];
  RAU_CHECK_CUDA(cudaSetDevice(f->ctx->device));
  const size_t nfeat = (size_t)f->B * f->C * f->S, nsmall = (size_t)f->T * f->B + 2 * (size_t)f->B;
  int ml = 0;   // x_len:max() (F:460), known on the host: the step unrolls that many encoder steps
  for (int b = 0; b < f->B; ++b) ml = s.h_len[b] > ml ? (int)s.h_len[b] : ml;
  s.max_len = ml < 1 ? 1 : (ml > f->T ? f->T : ml);
  RAU_CHECK_CUDA(cudaStreamWaitEvent(f->copy, s.freed, 0));
  RAU_CHECK_CUDA(cudaMemcpyAsync(s.d_tok, s.h_tok, nsmall * sizeof(float), cudaMemcpyHostToDevice, f->copy));
  if (f->format == RAU_FEED_F16_DIRECT) {   // the training step's feature pack reads the fp16 buffer itself
    RAU_CHECK_CUDA(cudaMemcpyAsync(s.d_stage, s.h_feats, nfeat * 2, cudaMemcpyHostToDevice, f->copy));
    RAU_CHECK_CUDA(cudaEventRecord(s.copied, f->copy));
  } else if (f->format == RAU_FEED_F16) {
    RAU_CHECK_CUDA(cudaMemcpyAsync(s.d_stage, s.h_feats, nfeat * 2, cudaMemcpyHostToDevice, f->copy));
    RAU_CHECK_CUDA(cudaEventRecord(s.copied, f->copy));
    // a SMALL grid: the conversion has a whole training step to finish and must not take SMs from the step's persistent
    // kernels (32 CTAs move the 154 MB of a 256-sample batch in well under a millisecond)
    half_to_float_kernel<<<32, 256, 0, f->copy>>>((const __half2*)s.d_stage, (int64_t)nfeat / 2, (float2*)s.d_feats);
    f->ctx->launches++;
    RAU_CHECK_CUDA(cudaGetLastError());
  } else {
    RAU_CHECK_CUDA(cudaMemcpyAsync(s.d_feats, s.h_feats, nfeat * sizeof(float), cudaMemcpyHostToDevice, f->copy));
    RAU_CHECK_CUDA(cudaEventRecord(s.copied, f->copy));
  }
  RAU_CHECK_CUDA(cudaEventRecord(s.ready, f->copy));
  s.in_flight = true;
  return RAU_OK;
}

// Make the context's stream wait for the slot's upload and describe it as a rau_batch (B_global = B; a data-parallel caller
// overwrites it).  The pointers stay valid until rau_feed_release(slot).
int rau_feed_acquire(rau_feed* f, int slot, rau_batch* batch) {
  RAU_REQUIRE(f && batch && slot >= 0 && slot < f->depth, "rau_feed_acquire: bad feed / slot");
  rau_feed::Slot& s = f->slots[slot];
  RAU_REQUIRE(s.in_flight, "rau_feed_acquire: slot %d was not submitted", slot);
  RAU_CHECK_CUDA(cudaStreamWaitEvent(f->ctx->stream, s.ready, 0));
  batch->B = f->B; batch->B_global = f->B;
  batch->feats = s.d_feats; batch->feats_f16 = f->format == RAU_FEED_F16_DIRECT ? s.d_stage : nullptr; batch->tokens = s.d_tok; batch->lengths = s.d_len; batch->labels = s.d_lab;
  batch->max_len = s.max_len;
  return RAU_OK;
}

// Everything enqueued on the context's stream so far may still read the slot; uploads submitted later wait for it.
int rau_feed_release(rau_feed* f, int slot) {
  RAU_REQUIRE(f && slot >= 0 && slot < f->depth, "rau_feed_release: bad feed / slot");
  RAU_CHECK_CUDA(cudaEventRecord(f->slots[slot].freed, f->ctx->stream));
  return RAU_OK;
}

// ------------------------------------------------------------------ device-resident feature cache
int rau_feat_cache_destroy(rau_feat_cache* c) {
  if (c == nullptr) return RAU_OK;
  cudaSetDevice(c->ctx->device);
  cudaStreamSynchronize(c->ctx->stream);
  if (c->data) cudaFree(c->data);
  if (c->stage) cudaFree(c->stage);
  delete c;
  return RAU_OK;
}

int rau_feat_cache_create(rau_ctx* ctx, int64_t n_images, int C, int S, rau_feat_cache** out) {
  RAU_REQUIRE(ctx && out && n_images > 0 && C > 0 && S > 0 && ((int64_t)C * S) % 2 == 0, "rau_feat_cache_create: bad arguments");
  *out = nullptr;
  RAU_CHECK_CUDA(cudaSetDevice(ctx->device));
  rau_feat_cache* c = new rau_feat_cache();
  c->ctx = ctx; c->n_images = n_images; c->per_image = (int64_t)C * S;
  c->stage_elems = c->per_image * 64;   // uploads go through a 64-image fp32 staging block
  if (cudaMalloc((void**)&c->data, sizeof(__half) * (size_t)n_images * c->per_image) != cudaSuccess ||
      cudaMalloc((void**)&c->stage, sizeof(float) * (size_t)c->stage_elems) != cudaSuccess) {
    rau_set_error("rau_feat_cache_create: %lld images x %lld halves do not fit: %s", (long long)n_images,
                  (long long)c->per_image, cudaGetErrorString(cudaGetLastError()));
    rau_feat_cache_destroy(c);
    return RAU_ENOMEM;
  }
  *out = c;
  return RAU_OK;
}

// Store n images starting at `first` (0-based) from HOST float32 features [n, C, S]; converted to fp16 on the device.
// Load-time call (once per split), synchronous.
int rau_feat_cache_put(rau_feat_cache* c, int64_t first, int64_t n, const float* host_feats) {
  RAU_REQUIRE(c && host_feats && first >= 0 && n > 0 && first + n <= c->n_images, "rau_feat_cache_put: range %lld + %lld of %lld",
              (long long)first, (long long)n, c ? (long long)c->n_images : 0);
  RAU_CHECK_CUDA(cudaSetDevice(c->ctx->device));
  cudaStream_t st = c->ctx->stream;
  for (int64_t i = 0; i < n; i += 64) {
    const int64_t k = (n - i < 64 ? n - i : 64) * c->per_image;
    RAU_CHECK_CUDA(cudaMemcpyAsync(c->stage, host_feats + i * c->per_image, sizeof(float) * (size_t)k, cudaMemcpyHostToDevice, st));
    float_to_half_kernel<<<blocks_for(k / 2), 256, 0, st>>>((const float2*)c->stage, k / 2,
                                                            (__half2*)(c->data + (first + i) * c->per_image));
    c->ctx->launches++;
    RAU_CHECK_CUDA(cudaGetLastError());
    RAU_CHECK_CUDA(cudaStreamSynchronize(st));   // (the host block may be pageable and is reused by the caller)
  }
  return RAU_OK;
}

// feats[b] = the cached features of image image_index[b] (DEVICE array of B floats, 1-based like every index the scripts
// handle) as float32 [B, C, S], on the context's stream: the batch's feature tensor without any host traffic.
int rau_feat_cache_gather(rau_feat_cache* c, const float* image_index, int B, float* feats) {
  RAU_REQUIRE(c && B > 0, "rau_feat_cache_gather: bad arguments");
  RAU_TRY(rau_check_dev(image_index, "image_index"));
  RAU_TRY(rau_check_dev(feats, "feats"));
  RAU_CHECK_CUDA(cudaSetDevice(c->ctx->device));
  const int64_t per2 = c->per_image / 2;
  int gx = (int)((per2 + 255) / 256);
  if (gx > 64) gx = 64;
  gather_kernel<<<dim3(gx, B), 256, 0, c->ctx->stream>>>((const __half2*)c->data, per2, c->n_images, image_index, (float2*)feats);
  RAU_LAUNCH_CHECK(c->ctx);
  return RAU_OK;
}

// The same as fp16 [B, C, S] (rau_batch.feats_f16: the training step's feature pack reads it directly): a third of the
// gather's HBM traffic.  Needs C*S % 8 == 0 and a 16-byte aligned destination.
int rau_feat_cache_gather_f16(rau_feat_cache* c, const float* image_index, int B, void* feats_f16) {
  RAU_REQUIRE(c && B > 0, "rau_feat_cache_gather_f16: bad arguments");
  RAU_TRY(rau_check_dev(image_index, "image_index"));
  RAU_TRY(rau_check_dev(feats_f16, "feats_f16"));
  RAU_REQUIRE(c->per_image % 8 == 0 && ((uintptr_t)feats_f16 & 15) == 0, "rau_feat_cache_gather_f16: C*S %% 8 != 0 or misaligned output");
  RAU_CHECK_CUDA(cudaSetDevice(c->ctx->device));
  const int64_t per8 = c->per_image / 8;
  int gx = (int)((per8 + 255) / 256);
  if (gx > 64) gx = 64;
  gather_f16_kernel<<<dim3(gx, B), 256, 0, c->ctx->stream>>>((const uint4*)c->data, per8, c->n_images, image_index, (uint4*)feats_f16);
  RAU_LAUNCH_CHECK(c->ctx);
  return RAU_OK;
}

}  // extern "C"
