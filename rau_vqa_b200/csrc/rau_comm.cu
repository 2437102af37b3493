// rau_comm.cu -- batch-sharded data parallelism: one process per GPU, one NCCL sum of each flat
// gradient per step over NVLink 5 / NVSwitch.  The reference has no multi-GPU path (F:128-146), so the only
// contract is "same result as one big batch".  libnccl is bound at run time (dlopen) so that librau.so loads on
// hosts without NCCL and shares the copy a host process (e.g. torch) may already have mapped.
#include "rau_common.cuh"
#include <dlfcn.h>

typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclFloat32_ = 7, ncclSum_ = 0 };

struct NcclApi {
  void* lib = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;   // optional: several all-reduces issued as one launch
  ncclResult_t (*GroupEnd)() = nullptr;
};

static NcclApi g_nccl;

static int nccl_load() {
  if (g_nccl.lib) return RAU_OK;
  const char* names[] = {"libnccl.so.2", "libnccl.so"};
  void* lib = nullptr;
  for (const char* n : names) {
    lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL);
    if (lib) break;
  }
  if (!lib) { rau_set_error("libnccl.so.2 is not loadable: %s", dlerror()); return RAU_ENCCL; }
  NcclApi a;
  a.lib = lib;
  a.GetUniqueId = (decltype(a.GetUniqueId))dlsym(lib, "ncclGetUniqueId");
  a.CommInitRank = (decltype(a.CommInitRank))dlsym(lib, "ncclCommInitRank");
  a.CommDestroy = (decltype(a.CommDestroy))dlsym(lib, "ncclCommDestroy");
  a.AllReduce = (decltype(a.AllReduce))dlsym(lib, "ncclAllReduce");
  a.GetErrorString = (decltype(a.GetErrorString))dlsym(lib, "ncclGetErrorString");
  a.GroupStart = (decltype(a.GroupStart))dlsym(lib, "ncclGroupStart");
  a.GroupEnd = (decltype(a.GroupEnd))dlsym(lib, "ncclGroupEnd");
  if (!a.GetUniqueId || !a.CommInitRank || !a.CommDestroy || !a.AllReduce || !a.GetErrorString) {
    rau_set_error("libnccl is missing a required symbol");
    return RAU_ENCCL;
  }
  g_nccl = a;
  return RAU_OK;
}

struct RauComm {
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
};

#define RAU_CHECK_NCCL(expr)                                                                  \
  do {                                                                                        \
    ncclResult_t _r = (expr);                                                                 \
    if (_r != 0) {                                                                            \
      rau_set_error("%s failed: %s", #expr, g_nccl.GetErrorString(_r));                       \
      return RAU_ENCCL;                                                                       \
    }                                                                                         \
  } while (0)

bool rau_comm_attached(rau_ctx* ctx) { return ctx->comm != nullptr && ctx->comm->world > 1; }
int rau_comm_rank(rau_ctx* ctx) { return ctx->comm ? ctx->comm->rank : 0; }

int rau_allreduce_internal(rau_ctx* ctx, float* buf, int64_t n) {
  if (!rau_comm_attached(ctx)) return RAU_OK;
  RAU_CHECK_NCCL(g_nccl.AllReduce(buf, buf, (size_t)n, ncclFloat32_, ncclSum_, ctx->comm->comm, ctx->stream));
  return RAU_OK;
}

// brackets for a batch of rau_allreduce_internal calls: NCCL then launches them as one kernel
int rau_allreduce_group(rau_ctx* ctx, int begin) {
  if (!rau_comm_attached(ctx) || !g_nccl.GroupStart || !g_nccl.GroupEnd) return RAU_OK;
  if (begin) RAU_CHECK_NCCL(g_nccl.GroupStart());
  else RAU_CHECK_NCCL(g_nccl.GroupEnd());
  return RAU_OK;
}

int rau_comm_destroy_internal(rau_ctx* ctx) {
  if (ctx->comm) {
    // CUDA graphs that captured an all-reduce reference the communicator: destroy them first (ncclCommDestroy
    // otherwise waits forever), and let every queued collective finish
    ctx->graph.clear();
    cudaStreamSynchronize(ctx->stream);
    if (ctx->comm->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(ctx->comm->comm);
    delete ctx->comm;
    ctx->comm = nullptr;
  }
  return RAU_OK;
}

extern "C" {

int rau_comm_unique_id(uint8_t id_out[128]) {
  if (id_out == nullptr) { rau_set_error("id_out == NULL"); return RAU_EINVAL; }
  RAU_TRY(nccl_load());
  ncclUniqueId id;
  RAU_CHECK_NCCL(g_nccl.GetUniqueId(&id));
  memcpy(id_out, id.internal, 128);
  return RAU_OK;
}

int rau_comm_init(rau_ctx* ctx, const uint8_t id[128], int rank, int world) {
  RAU_REQUIRE(ctx && id, "ctx/id == NULL");
  RAU_REQUIRE(world >= 1 && rank >= 0 && rank < world, "bad rank %d / world %d", rank, world);
  RAU_CHECK_CUDA(cudaSetDevice(ctx->device));
  rau_comm_destroy_internal(ctx);
  RauComm* c = new RauComm();
  c->rank = rank; c->world = world;
  if (world > 1) {
    int s = nccl_load();
    if (s != RAU_OK) { delete c; return s; }
    ncclUniqueId uid;
    memcpy(uid.internal, id, 128);
    ncclResult_t r = g_nccl.CommInitRank(&c->comm, world, uid, rank);
    if (r != 0) {
      rau_set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r));
      delete c;
      return RAU_ENCCL;
    }
  }
  ctx->comm = c;
  return RAU_OK;
}

int rau_comm_destroy(rau_ctx* ctx) {
  RAU_REQUIRE(ctx, "ctx == NULL");
  return rau_comm_destroy_internal(ctx);
}

int rau_allreduce(rau_ctx* ctx, float* buf, int64_t n) {
  RAU_REQUIRE(ctx && buf && n > 0, "bad allreduce arguments");
  return rau_allreduce_internal(ctx, buf, n);
}

int rau_allreduce_grads(rau_ctx* ctx, float* const grads[3], const int64_t sizes[3]) {
  RAU_REQUIRE(ctx && grads && sizes, "bad allreduce arguments");
  for (int g = 0; g < 3; ++g) RAU_TRY(rau_allreduce_internal(ctx, grads[g], sizes[g]));
  return RAU_OK;
}

}  // extern "C"
