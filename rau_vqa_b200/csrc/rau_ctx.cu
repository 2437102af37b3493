// rau_ctx.cu -- context lifetime, error text, device arena and the flat parameter layout of include/rau.h.
#include <algorithm>
#include "rau_layout.cuh"
#include <stdarg.h>
#include <stdlib.h>

static thread_local char g_err[512] = "";

void rau_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

int RauArena::get(const char* name, size_t bytes, void** out) {
  Buf& b = bufs[name];
  if (b.bytes < bytes) {
    if (b.p) {
      // the stream may still be using the old block: drain before freeing
      cudaDeviceSynchronize();
      cudaFree(b.p);
      b.p = nullptr;
      b.bytes = 0;
      generation++;   // captured graphs that baked the old address in must be dropped (rau_train_step checks)
    }
    size_t want = (bytes + 255) & ~(size_t)255;
    cudaError_t e = cudaMalloc(&b.p, want);
    if (e != cudaSuccess) {
      rau_set_error("cudaMalloc(%zu) for '%s' failed: %s", want, name, cudaGetErrorString(e));
      return RAU_ENOMEM;
    }
    b.bytes = want;
  }
  *out = b.p;
  return RAU_OK;
}

void RauArena::release() {
  for (auto& kv : bufs)
    if (kv.second.p) cudaFree(kv.second.p);
  bufs.clear();
}

int rau_comm_destroy_internal(rau_ctx* ctx);  // rau_comm.cu

// the ONLY place the library reads the environment
RauTuning rau_tuning_from_env() {
  RauTuning t;
  auto geti = [](const char* name, int dflt) { const char* e = getenv(name); return e ? atoi(e) : dflt; };
  t.graph = geti("RAU_GRAPH", t.graph);
  t.overlap = geti("RAU_OVERLAP", t.overlap);
  t.side_ctas = geti("RAU_SIDE_CTAS", 0);
  t.side_ctas_fwd = geti("RAU_SIDE_CTAS_FWD", 0);
  t.side_ctas_bwd = geti("RAU_SIDE_CTAS_BWD", 0);
  t.main_ctas = geti("RAU_MAIN_CTAS", 0);
  t.rows = geti("RAU_ROWS", t.rows);
  t.cg2 = geti("RAU_CG2", t.cg2);
  t.tanh_ew = geti("RAU_TANH_EW", t.tanh_ew) == 8 ? 8 : 16;
  t.lin_cg2 = geti("RAU_LIN_CG2", t.lin_cg2);
  t.enc_w0 = geti("RAU_ENC_W0", t.enc_w0);
  t.rows_trace = geti("RAU_ROWS_TRACE", 0);
  t.lstm_seq = geti("RAU_LSTM_SEQ", t.lstm_seq);
  t.enc_bwd_wave = geti("RAU_ENC_BWD_WAVE", t.enc_bwd_wave);
  t.xprep_hops = geti("RAU_XPREP_HOPS", t.xprep_hops);
  t.pdl = geti("RAU_PDL", 0);   // opt-in: measured SLOWER on the graph-replayed step (6.28 vs 4.82 ms)
  t.phases = geti("RAU_PHASES", 0);
  { const char* e = getenv("RAU_TC_MIN_WORK"); if (e) t.tc_min_work = atoll(e); }
  t.time_cap = geti("RAU_TIME_CAP", 0);
  return t;
}
const RauTuning& rau_process_tuning() {
  static const RauTuning t = rau_tuning_from_env();
  return t;
}

bool rau_pdl_enabled() { return rau_process_tuning().pdl != 0; }

namespace {
__global__ void stamp_kernel(unsigned long long* out) {
  unsigned long long t;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
  *out = t;
}
}  // namespace

void rau_phase_mark(rau_ctx* ctx, const char* name) {
  if (ctx->phases < 0) ctx->phases = ctx->tune.phases;
  if (!ctx->phases) return;
  if (ctx->phases == 2) {
    if (!ctx->stamp_buf && cudaMalloc(&ctx->stamp_buf, sizeof(unsigned long long) * 1024) != cudaSuccess) return;
    if (strcmp(name, "begin") == 0) ctx->stamp_names.clear();
    if (ctx->stamp_names.size() >= 1024) return;
    stamp_kernel<<<1, 1, 0, ctx->stream>>>(ctx->stamp_buf + ctx->stamp_names.size());
    ctx->stamp_names.push_back(std::string(ctx->stream == ctx->side ? "S " : "M ") + name);
    return;
  }
  cudaEvent_t ev;
  if (cudaEventCreate(&ev) != cudaSuccess) return;
  cudaEventRecord(ev, ctx->stream);
  ctx->phase_ev.push_back({name, ev});
}

int rau_check_async_error(rau_ctx* ctx);

extern "C" {

int rau_version(void) { return 100; }

/* debugging aid (RAU_PHASES=1): prints the milliseconds between the phase marks recorded since the last report */
int rau_phase_report(rau_ctx* ctx) {
  if (ctx == nullptr) return RAU_EINVAL;
  if (ctx->phases == 2 && ctx->stamp_buf && !ctx->stamp_names.empty()) {
    cudaDeviceSynchronize();
    std::vector<unsigned long long> t(ctx->stamp_names.size());
    cudaMemcpy(t.data(), ctx->stamp_buf, sizeof(unsigned long long) * t.size(), cudaMemcpyDeviceToHost);
    std::vector<size_t> order(t.size());
    for (size_t i = 0; i < order.size(); ++i) order[i] = i;
    std::stable_sort(order.begin(), order.end(), [&](size_t a, size_t b) { return t[a] < t[b]; });
    for (size_t i : order)
      fprintf(stderr, "[rau stamp] %9.1f us  %s\n", (double)(t[i] - t[0]) * 1e-3, ctx->stamp_names[i].c_str());
    return RAU_OK;
  }
  if (ctx->phase_ev.empty()) return RAU_OK;
  cudaStreamSynchronize(ctx->stream);
  for (size_t i = 1; i < ctx->phase_ev.size(); ++i) {
    float ms = 0.0f;
    cudaEventElapsedTime(&ms, ctx->phase_ev[i - 1].second, ctx->phase_ev[i].second);
    fprintf(stderr, "[rau phase] %-28s %8.3f ms\n", ctx->phase_ev[i].first.c_str(), ms);
  }
  for (auto& kv : ctx->phase_ev) cudaEventDestroy(kv.second);
  ctx->phase_ev.clear();
  return RAU_OK;
}

const char* rau_last_error(void) { return g_err; }

int rau_ctx_create(rau_ctx** out, int device, void* cuda_stream) {
  if (out == nullptr) { rau_set_error("rau_ctx_create: out == NULL"); return RAU_EINVAL; }
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    rau_set_error("no CUDA device (%s); librau has no CPU fallback", cudaGetErrorString(e));
    return RAU_ECUDA;
  }
  RAU_REQUIRE(device >= 0 && device < ndev, "device %d out of range (%d devices)", device, ndev);
  cudaDeviceProp prop;
  RAU_CHECK_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10) {
    rau_set_error("device %d is sm_%d%d; librau is built for sm_100a (B200) only", device, prop.major, prop.minor);
    return RAU_EARCH;
  }
  RAU_CHECK_CUDA(cudaSetDevice(device));
  rau_ctx* ctx = new rau_ctx();
  ctx->device = device;
  ctx->stream = (cudaStream_t)cuda_stream;
  ctx->sm_count = prop.multiProcessorCount;
  if (cudaEventCreate(&ctx->ev0) != cudaSuccess || cudaEventCreate(&ctx->ev1) != cudaSuccess) {
    rau_set_error("cudaEventCreate failed");
    delete ctx;
    return RAU_ECUDA;
  }
  if (cudaMalloc(&ctx->d_err, sizeof(unsigned int)) != cudaSuccess || cudaMemset(ctx->d_err, 0, sizeof(unsigned int)) != cudaSuccess ||
      cudaHostAlloc((void**)&ctx->h_err, sizeof(unsigned int), cudaHostAllocMapped) != cudaSuccess ||
      cudaHostGetDevicePointer((void**)&ctx->h_err_dev, (void*)ctx->h_err, 0) != cudaSuccess) {
    rau_set_error("context allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete ctx;
    return RAU_ECUDA;
  }
  *ctx->h_err = 0;
  if (cudaMalloc(&ctx->d_ss, sizeof(StepState)) != cudaSuccess ||
      cudaMallocHost(&ctx->h_ss, sizeof(StepState) * 64) != cudaSuccess ||
      cudaStreamCreateWithPriority(&ctx->gstream, cudaStreamNonBlocking, -1) != cudaSuccess ||
      cudaStreamCreateWithPriority(&ctx->side, cudaStreamNonBlocking, 0) != cudaSuccess ||
      cudaStreamCreateWithPriority(&ctx->aux, cudaStreamNonBlocking, -1) != cudaSuccess ||
      cudaStreamCreateWithPriority(&ctx->aux2, cudaStreamNonBlocking, -1) != cudaSuccess) {
    rau_set_error("context allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete ctx;
    return RAU_ECUDA;
  }
  ctx->tune = rau_tuning_from_env();
  if (ctx->tune.graph == 0) ctx->graph.disabled = true;
  *out = ctx;
  return RAU_OK;
}

int rau_ctx_destroy(rau_ctx* ctx) {
  if (ctx == nullptr) return RAU_OK;
  cudaSetDevice(ctx->device);
  cudaStreamSynchronize(ctx->stream);
  ctx->graph.clear();            // captured steps hold NCCL kernels: they must be gone before the communicator is
  cudaDeviceSynchronize();
  rau_comm_destroy_internal(ctx);
  ctx->arena.release();
  if (ctx->stamp_buf) cudaFree(ctx->stamp_buf);
  if (ctx->d_ss) cudaFree(ctx->d_ss);
  if (ctx->h_ss) cudaFreeHost(ctx->h_ss);
  if (ctx->d_err) cudaFree(ctx->d_err);
  if (ctx->h_err) cudaFreeHost((void*)ctx->h_err);
  for (cudaEvent_t e : ctx->ss_ev) if (e) cudaEventDestroy(e);
  if (ctx->gstream) cudaStreamDestroy(ctx->gstream);
  if (ctx->side) cudaStreamDestroy(ctx->side);
  if (ctx->aux) cudaStreamDestroy(ctx->aux);
  if (ctx->aux2) cudaStreamDestroy(ctx->aux2);
  for (cudaEvent_t e : ctx->side_ev) cudaEventDestroy(e);
  if (ctx->ev0) cudaEventDestroy(ctx->ev0);
  if (ctx->ev1) cudaEventDestroy(ctx->ev1);
  delete ctx;
  return RAU_OK;
}

int rau_set_stream(rau_ctx* ctx, void* cuda_stream) {
  RAU_REQUIRE(ctx, "ctx == NULL");
  ctx->stream = (cudaStream_t)cuda_stream;
  return RAU_OK;
}

int rau_set_seed(rau_ctx* ctx, uint64_t seed) {
  RAU_REQUIRE(ctx, "ctx == NULL");
  ctx->seed = seed;
  return RAU_OK;
}

int rau_set_precision(rau_ctx* ctx, int precision) {
  RAU_REQUIRE(ctx, "ctx == NULL");
  RAU_REQUIRE(precision >= RAU_PREC_F32 && precision <= RAU_PREC_F16IMG, "unknown precision %d", precision);
  ctx->precision = precision;
  return RAU_OK;
}

int rau_get_precision(rau_ctx* ctx) { return ctx ? ctx->precision : RAU_EINVAL; }

int rau_sync(rau_ctx* ctx) {
  RAU_REQUIRE(ctx, "ctx == NULL");
  RAU_CHECK_CUDA(cudaStreamSynchronize(ctx->stream));
  return rau_check_async_error(ctx);
}

int64_t rau_launch_count(rau_ctx* ctx) { return ctx ? ctx->launches : 0; }

int64_t rau_group_size(const rau_config* cfg, int group) {
  if (cfg == nullptr) return -1;
  return rau_layout(cfg, group, nullptr);
}

int64_t rau_param_offset(const rau_config* cfg, int group, const char* name) {
  if (cfg == nullptr || name == nullptr) return -1;
  std::vector<ParamEntry> e;
  if (rau_layout(cfg, group, &e) < 0) return -1;
  for (const auto& p : e)
    if (strcmp(p.name, name) == 0) return p.off;
  return -1;
}

}  // extern "C"

// A persistent recurrence kernel gave up on a peer CTA (instead of hanging the device): its results and everything
// computed from them are invalid.  The optimizer kernel saw the same word and left the parameters untouched.
int rau_check_async_error(rau_ctx* ctx) {
  if (ctx->h_err == nullptr || *ctx->h_err == 0) return RAU_OK;
  cudaStreamSynchronize(ctx->stream);
  *ctx->h_err = 0;
  cudaMemset(ctx->d_err, 0, sizeof(unsigned int));
  rau_set_error("persistent LSTM recurrence: a CTA timed out waiting for its row tile's peers; the step's results are invalid "
                "and its parameter update was skipped");
  return RAU_ECUDA;
}

// ------------------------------------------------------------------ layout
static const char* kRnnNames[4][4] = {{"l1.Wi", "l1.bi", "l1.Wh", "l1.bh"},
                                      {"l2.Wi", "l2.bi", "l2.Wh", "l2.bh"},
                                      {"l3.Wi", "l3.bi", "l3.Wh", "l3.bh"},
                                      {"l4.Wi", "l4.bi", "l4.Wh", "l4.bh"}};

int64_t rau_layout(const rau_config* cfg, int group, std::vector<ParamEntry>* out) {
  int64_t off = 0;
  auto add = [&](const char* name, int64_t rows, int64_t cols) {
    if (out) out->push_back(ParamEntry{name, off, rows, cols});
    off += rows * cols;
  };
  const int Q = 2 * cfg->Hq * cfg->nlayer;
  switch (group) {
    case 0:
      add("E", cfg->V, cfg->embed);
      break;
    case 1:
      if (cfg->nlayer < 1 || cfg->nlayer > 4) return -1;
      for (int L = 0; L < cfg->nlayer; ++L) {
        const int in = L == 0 ? cfg->embed : cfg->Hq;
        add(kRnnNames[L][0], 4 * cfg->Hq, in);
        add(kRnnNames[L][1], 4 * cfg->Hq, 1);
        add(kRnnNames[L][2], 4 * cfg->Hq, cfg->Hq);
        add(kRnnNames[L][3], 4 * cfg->Hq, 1);
      }
      break;
    case 2:
      add("Wq", cfg->M, Q);            add("bq", cfg->M, 1);
      add("Wh", cfg->M, cfg->H);       add("bh", cfg->M, 1);
      add("Wi", cfg->M, cfg->C);       add("bi", cfg->M, 1);
      add("Wqa", cfg->A, cfg->M);      add("bqa", cfg->A, 1);
      add("Wa", cfg->A, cfg->M);       add("ba", cfg->A, 1);
      add("ws", 1, cfg->A);
      add("Wm", cfg->S, cfg->H);       add("bm", cfg->S, 1);
      add("Wp", cfg->M, cfg->S);       add("bp", cfg->M, 1);
      add("Wx", 4 * cfg->H, cfg->M);   add("bx", 4 * cfg->H, 1);
      add("Whh", 4 * cfg->H, cfg->H);  add("bhh", 4 * cfg->H, 1);
      add("Wo", cfg->M, cfg->H);       add("bo", cfg->M, 1);
      add("Ws", cfg->N, cfg->M);       add("bso", cfg->N, 1);
      add("wd", 1, cfg->M);
      // the two scalars last: every matrix and vector above starts on a 16-byte boundary (TMA / float4 access)
      add("bs", 1, 1);                 add("bd", 1, 1);
      break;
    default:
      return -1;
  }
  return off;
}
