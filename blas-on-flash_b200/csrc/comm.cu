// Multi-GPU plumbing behind the C ABI: one communicator rank per context.
//
// north_star (5): output row blocks are sharded over the GPUs with no collective on the compute path; what does
// cross NVLink is (i) the replicated dense operand of gemm / csrmm, which every rank uploads 1/world of and
// broadcasts panel by panel so that it crosses PCIe once per node instead of once per GPU, and
// (ii) the k-means allreduce of centroid sums and counts, the only reduction on the path.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2: the copy the process already holds -- torch's, under
// torchrun -- or the system one), so libbof_b200.so has no link-time dependency on it and single-GPU users never
// load it.  Two ways to get ranks:
//   * one process per GPU (torchrun): rank 0 calls bof_comm_unique_id, the 128 bytes travel by whatever the
//     launcher offers (bench.py / dist.py: a torch.distributed broadcast), every rank calls bof_comm_init;
//   * one process, several GPUs (C++ applications, drivers/): bof_mgpu_create (mgpu.cu) makes one context and one
//     worker thread per device and initialises the ranks itself.
#include "host_internal.cuh"

#include <dlfcn.h>
#include <nccl.h>   // types and enums only; every entry point is looked up with dlsym

#include <mutex>

namespace bof {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommInitRankConfig)(ncclComm_t*, int, ncclUniqueId, int, ncclConfig_t*) = nullptr;   // optional
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*Broadcast)(const void*, void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};

static NcclApi g_nccl;
static std::once_flag g_nccl_once;
static std::string g_nccl_err;

static const NcclApi* nccl_api() {
  std::call_once(g_nccl_once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD | RTLD_GLOBAL);   // the copy already in the process
    if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { g_nccl_err = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
    g_nccl.handle = h;
    bool ok = true;
    auto sym = [&](const char* name) { void* p = dlsym(h, name); if (!p) { ok = false; g_nccl_err = std::string("libnccl lacks ") + name; } return p; };
    g_nccl.GetUniqueId = reinterpret_cast<decltype(g_nccl.GetUniqueId)>(sym("ncclGetUniqueId"));
    g_nccl.CommInitRank = reinterpret_cast<decltype(g_nccl.CommInitRank)>(sym("ncclCommInitRank"));
    g_nccl.CommDestroy = reinterpret_cast<decltype(g_nccl.CommDestroy)>(sym("ncclCommDestroy"));
    g_nccl.Broadcast = reinterpret_cast<decltype(g_nccl.Broadcast)>(sym("ncclBroadcast"));
    g_nccl.AllReduce = reinterpret_cast<decltype(g_nccl.AllReduce)>(sym("ncclAllReduce"));
    g_nccl.AllGather = reinterpret_cast<decltype(g_nccl.AllGather)>(sym("ncclAllGather"));
    g_nccl.GroupStart = reinterpret_cast<decltype(g_nccl.GroupStart)>(sym("ncclGroupStart"));
    g_nccl.GroupEnd = reinterpret_cast<decltype(g_nccl.GroupEnd)>(sym("ncclGroupEnd"));
    g_nccl.GetErrorString = reinterpret_cast<decltype(g_nccl.GetErrorString)>(sym("ncclGetErrorString"));
    g_nccl.CommInitRankConfig = reinterpret_cast<decltype(g_nccl.CommInitRankConfig)>(dlsym(h, "ncclCommInitRankConfig"));
    if (!ok) g_nccl.handle = nullptr;
  });
  return g_nccl.handle ? &g_nccl : nullptr;
}

struct CommState {
  ncclComm_t comm = nullptr;
  int world = 1, rank = 0;
  int max_ctas = 0;   // CTAs an NCCL kernel of this communicator may occupy (0: unknown / unlimited)
};

// CTAs per collective kernel: few, because the panel broadcasts run next to persistent tensor-core kernels and only
// have to keep up with PCIe-rate uploads of the other ranks (BOF_NCCL_MAX_CTAS overrides)
static int nccl_max_ctas() {
  const char* e = getenv("BOF_NCCL_MAX_CTAS");
  const int v = e ? atoi(e) : 8;
  return std::min(32, std::max(1, v));
}

#define BOF_NCCL(ctx, expr)                                                                              \
  do {                                                                                                   \
    ncclResult_t r__ = (expr);                                                                           \
    if (r__ != ncclSuccess)                                                                              \
      return bof::fail((ctx), BOF_ECUDA, "%s failed: %s (%s:%d)", #expr, bof::g_nccl.GetErrorString(r__), __FILE__, __LINE__); \
  } while (0)

int comm_world(const bof_ctx* ctx) { return ctx && ctx->comm ? ctx->comm->world : 1; }
// SMs to keep free for collective kernels that may be waiting on the GPU (pairs of SMs: the GEMM launches 2-CTA clusters)
int comm_sm_reserve(const bof_ctx* ctx) {
  if (!ctx || !ctx->comm || ctx->comm->world <= 1) return 0;
  return 2 * (ctx->comm->max_ctas > 0 ? ctx->comm->max_ctas : 16);
}
int comm_rank(const bof_ctx* ctx) { return ctx && ctx->comm ? ctx->comm->rank : 0; }

// Broadcast `count` floats at `buf` (same address role on every rank) from `root`, on the context's collective
// stream.  The caller orders it against the other streams with events.
int comm_broadcast_f32(bof_ctx* ctx, float* buf, size_t count, int root) {
  if (!ctx->comm) return fail(ctx, BOF_EINVAL, "no communicator on this context (bof_comm_init)");
  BOF_NCCL(ctx, g_nccl.Broadcast(buf, buf, count, ncclFloat32, root, ctx->comm->comm, ctx->coll));
  return BOF_OK;
}

// In-place sum of `count` floats over the ranks on stream `s` (k-means: [K*d sums | counts]).
int comm_allreduce_sum_f32(bof_ctx* ctx, float* buf, size_t count, cudaStream_t s) {
  if (!ctx->comm) return fail(ctx, BOF_EINVAL, "no communicator on this context (bof_comm_init)");
  BOF_NCCL(ctx, g_nccl.AllReduce(buf, buf, count, ncclFloat32, ncclSum, ctx->comm->comm, s));
  return BOF_OK;
}

void comm_destroy(bof_ctx* ctx) {
  if (!ctx || !ctx->comm) return;
  if (ctx->comm->comm && g_nccl.handle) g_nccl.CommDestroy(ctx->comm->comm);
  delete ctx->comm;
  ctx->comm = nullptr;
  if (ctx->coll) { cudaStreamDestroy(ctx->coll); ctx->coll = nullptr; }
}

}  // namespace bof

using namespace bof;

extern "C" {

int bof_comm_unique_id(void* id_out) {
  if (!id_out) return BOF_EINVAL;
  const NcclApi* api = nccl_api();
  if (!api) return BOF_ENODEV;
  static_assert(sizeof(ncclUniqueId) == 128, "BOF_COMM_ID_BYTES");
  ncclUniqueId id;
  if (api->GetUniqueId(&id) != ncclSuccess) return BOF_ECUDA;
  memcpy(id_out, &id, sizeof(id));
  return BOF_OK;
}

int bof_comm_init(bof_ctx* ctx, int world, int rank, const void* id) {
  if (!ctx) return BOF_EINVAL;
  BOF_REQUIRE(ctx, world >= 1 && rank >= 0 && rank < world && id != nullptr, "comm_init: bad world / rank / id");
  BOF_REQUIRE(ctx, ctx->comm == nullptr, "comm_init: this context already has a communicator");
  const NcclApi* api = nccl_api();
  if (!api) return fail(ctx, BOF_ENODEV, "%s", g_nccl_err.c_str());
  BOF_CUDA(ctx, cudaSetDevice(ctx->device));
  ncclUniqueId uid;
  memcpy(&uid, id, sizeof(uid));
  CommState* cs = new CommState();
  cs->world = world; cs->rank = rank;
  ncclResult_t r;
  if (api->CommInitRankConfig != nullptr) {
    ncclConfig_t cfg = NCCL_CONFIG_INITIALIZER;
    cfg.maxCTAs = nccl_max_ctas();
    cfg.minCTAs = 1;
    cfg.cgaClusterSize = 1;   // no CTA clusters: the CTAs must fit into whatever SMs are free
    r = api->CommInitRankConfig(&cs->comm, world, uid, rank, &cfg);
    cs->max_ctas = cfg.maxCTAs;
  } else {
    r = api->CommInitRank(&cs->comm, world, uid, rank);
  }
  if (r != ncclSuccess) {
    delete cs;
    return fail(ctx, BOF_ECUDA, "ncclCommInitRank failed: %s", api->GetErrorString(r));
  }
  ctx->comm = cs;
  BOF_CUDA(ctx, cudaStreamCreateWithFlags(&ctx->coll, cudaStreamNonBlocking));
  return BOF_OK;
}

int bof_comm_finalize(bof_ctx* ctx) {
  if (!ctx) return BOF_OK;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  comm_destroy(ctx);
  return BOF_OK;
}

int bof_comm_world(const bof_ctx* ctx) { return comm_world(ctx); }
int bof_comm_rank(const bof_ctx* ctx) { return comm_rank(ctx); }

}  // extern "C"
