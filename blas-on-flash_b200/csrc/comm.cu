// Multi-GPU plumbing behind the C ABI: one communicator rank per context.
//
// north_star (5): output row blocks are sharded over the GPUs with no collective on the compute path; what does
// cross NVLink is (i) the replicated dense operand of gemm / csrmm, which every rank uploads 1/world of and
// broadcasts panel by panel so that it crosses PCIe once per node instead of once per GPU, and
// (ii) the k-means allreduce of centroid sums and counts, the only reduction on the path.
//
// The replicated operand does NOT travel through NCCL kernels: an NCCL kernel that waits for its root holds SMs, and
// next to a persistent tensor-core grid that stalls the MMAs (measured, profiles/r02).  Instead every rank owns an
// EXCHANGE BUFFER in HBM that its peers map (CUDA IPC between processes, plain peer pointers inside one process), and
// the owner of a panel pushes it into every peer's buffer with copy-engine peer copies over NVLink / NVSwitch,
// followed by a 32-bit flag write (cuStreamWriteValue32); consumers gate their streams with cuStreamWaitValue32.
// No SM is involved, no host thread waits, and nothing crosses a process boundary except device memory.
// NCCL keeps two jobs: the bootstrap all-gather of the IPC handles and the k-means allreduce.
//
// NCCL is resolved at run time (dlopen of libnccl.so.2: the copy the process already holds -- torch's, under
// torchrun -- or the system one), so libbof_b200.so has no link-time dependency on it and single-GPU users never
// load it.  Two ways to get ranks:
//   * one process per GPU (torchrun): rank 0 calls bof_comm_unique_id, the 128 bytes travel by whatever the
//     launcher offers (bench.py / dist.py: a torch.distributed broadcast), every rank calls bof_comm_init;
//   * one process, several GPUs (C++ applications, drivers/): bof_mgpu_create (mgpu.cu) makes one context and one
//     worker thread per device and initialises the ranks itself.
#include "host_internal.cuh"

#include <cuda.h>
#include <dlfcn.h>
#include <nccl.h>
#include <unistd.h>   // types and enums only; every entry point is looked up with dlsym

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

constexpr int kPushStreams = 4;
constexpr int kFlagReady = 0;      // flags[kFlagReady + r]: rank r has finished its previous collective call
constexpr int kFlagData = 64;      // flags[kFlagData + i]: item i of the current call has been pushed into my buffer
constexpr int kFlagCount = 64 + 256;

// what every rank publishes about itself (all-gathered through NCCL)
struct PeerInfo {
  uint64_t pid;
  uint64_t epoch;            // bumps when the exchange buffer is reallocated
  uint64_t xbuf_ptr;         // same-process peers use the pointers directly
  uint64_t flags_ptr;
  uint64_t xbuf_bytes;
  int32_t device;
  int32_t pad;
  cudaIpcMemHandle_t xbuf_handle;
  cudaIpcMemHandle_t flags_handle;
};

struct Peer {
  int device = -1;
  uint64_t epoch = ~0ull;     // epoch of the mapping we hold
  float* xbuf = nullptr;
  uint32_t* flags = nullptr;
  bool xbuf_ipc = false, flags_ipc = false;
};

struct CommState {
  ncclComm_t comm = nullptr;
  int world = 1, rank = 0;
  int max_ctas = 0;   // CTAs an NCCL kernel of this communicator may occupy (0: unknown / unlimited)
  // peer exchange
  float* xbuf = nullptr;       // my exchange buffer (cudaMalloc: IPC-exportable)
  size_t xbuf_bytes = 0;
  uint64_t epoch = 0;
  uint32_t* flags = nullptr;   // kFlagCount words, zero-initialised
  uint32_t seq = 0;            // collective call counter (identical on every rank)
  std::vector<Peer> peers;
  cudaStream_t push[kPushStreams] = {};
  PeerInfo* gather_dev = nullptr;   // world entries, for the bootstrap all-gather
  uint32_t* seq_dev = nullptr;      // device copy of `seq`: source of the 4-byte flag writes into the peers
  CUresult (*WaitValue32)(CUstream, CUdeviceptr, cuuint32_t, unsigned int) = nullptr;
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

// ---- peer exchange ------------------------------------------------------------------------------------------------

static int publish_and_map(bof_ctx* ctx) {
  CommState* cs = ctx->comm;
  PeerInfo mine{};
  mine.pid = (uint64_t)getpid();
  mine.epoch = cs->epoch;
  mine.xbuf_ptr = (uint64_t)(uintptr_t)cs->xbuf;
  mine.flags_ptr = (uint64_t)(uintptr_t)cs->flags;
  mine.xbuf_bytes = cs->xbuf_bytes;
  mine.device = ctx->device;
  if (cs->xbuf) BOF_CUDA(ctx, cudaIpcGetMemHandle(&mine.xbuf_handle, cs->xbuf));
  BOF_CUDA(ctx, cudaIpcGetMemHandle(&mine.flags_handle, cs->flags));
  BOF_CUDA(ctx, cudaMemcpyAsync(cs->gather_dev + cs->rank, &mine, sizeof(mine), cudaMemcpyHostToDevice, ctx->coll));
  BOF_NCCL(ctx, g_nccl.AllGather(cs->gather_dev + cs->rank, cs->gather_dev, sizeof(PeerInfo), ncclInt8, cs->comm, ctx->coll));
  std::vector<PeerInfo> all((size_t)cs->world);
  BOF_CUDA(ctx, cudaMemcpyAsync(all.data(), cs->gather_dev, sizeof(PeerInfo) * cs->world, cudaMemcpyDeviceToHost, ctx->coll));
  BOF_CUDA(ctx, cudaStreamSynchronize(ctx->coll));
  for (int r = 0; r < cs->world; ++r) {
    if (r == cs->rank) continue;
    Peer& pr = cs->peers[r];
    const PeerInfo& pi = all[r];
    const bool same_process = pi.pid == mine.pid;
    if (pr.device != pi.device) {
      pr.device = pi.device;
      if (pi.device != ctx->device) {
        cudaError_t e = cudaDeviceEnablePeerAccess(pi.device, 0);
        if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
          return fail(ctx, BOF_ECUDA, "no peer access from device %d to device %d: %s", ctx->device, pi.device, cudaGetErrorString(e));
        cudaGetLastError();
      }
    }
    if (pr.flags == nullptr) {
      if (same_process) pr.flags = reinterpret_cast<uint32_t*>((uintptr_t)pi.flags_ptr);
      else {
        void* q = nullptr;
        BOF_CUDA(ctx, cudaIpcOpenMemHandle(&q, pi.flags_handle, cudaIpcMemLazyEnablePeerAccess));
        pr.flags = static_cast<uint32_t*>(q);
        pr.flags_ipc = true;
      }
    }
    if (pr.epoch != pi.epoch) {
      if (pr.xbuf && pr.xbuf_ipc) BOF_CUDA(ctx, cudaIpcCloseMemHandle(pr.xbuf));
      pr.xbuf = nullptr; pr.xbuf_ipc = false;
      if (pi.xbuf_ptr != 0) {
        if (same_process) pr.xbuf = reinterpret_cast<float*>((uintptr_t)pi.xbuf_ptr);
        else {
          void* q = nullptr;
          BOF_CUDA(ctx, cudaIpcOpenMemHandle(&q, pi.xbuf_handle, cudaIpcMemLazyEnablePeerAccess));
          pr.xbuf = static_cast<float*>(q);
          pr.xbuf_ipc = true;
        }
      }
      pr.epoch = pi.epoch;
    }
  }
  return BOF_OK;
}

// Start of a collective call that exchanges up to `bytes` through the exchange buffers: grows mine if needed,
// refreshes the peer mappings, and tells every peer that my buffer is free again (my previous call has returned).
// Returns my exchange buffer.
int comm_exchange_begin(bof_ctx* ctx, size_t bytes, float** xbuf_out) {
  CommState* cs = ctx->comm;
  if (!cs || cs->world <= 1) return fail(ctx, BOF_EINVAL, "peer exchange needs a communicator with more than one rank");
  if (!cs->WaitValue32) return fail(ctx, BOF_ENODEV, "stream memory operations (cuStreamWaitValue32) are not available");
  if (cs->xbuf_bytes < bytes) {
    // every rank sees the same sizes (collective call), so every rank regrows in the same call
    BOF_CUDA(ctx, cudaDeviceSynchronize());
    if (cs->xbuf) BOF_CUDA(ctx, cudaFree(cs->xbuf));
    cs->xbuf = nullptr; cs->xbuf_bytes = 0;
    const size_t want = (bytes + ((size_t)64 << 20)) & ~(((size_t)2 << 20) - 1);
    void* q = nullptr;
    if (cudaMalloc(&q, want) != cudaSuccess) { cudaGetLastError(); return fail(ctx, BOF_ENOMEM, "exchange buffer: cudaMalloc of %zu bytes failed", want); }
    cs->xbuf = static_cast<float*>(q);
    cs->xbuf_bytes = want;
    cs->epoch++;
  }
  // Peers may have regrown too; the all-gather also orders "every rank has entered call seq+1" after "every rank has
  // left call seq" for the host side.  ~60 us per call.
  BOF_TRY(publish_and_map(ctx));
  cs->seq++;
  // flags are raised with 4-byte peer copies of this word (in stream order behind the data they announce)
  BOF_CUDA(ctx, cudaMemcpyAsync(cs->seq_dev, &cs->seq, sizeof(uint32_t), cudaMemcpyHostToDevice, cs->push[0]));
  for (int r = 0; r < cs->world; ++r) {
    if (r == cs->rank) continue;
    BOF_CUDA(ctx, cudaMemcpyPeerAsync(cs->peers[r].flags + kFlagReady + cs->rank, cs->peers[r].device, cs->seq_dev, ctx->device,
                                      sizeof(uint32_t), cs->push[0]));
  }
  cudaEvent_t ev = get_event(ctx, 120);
  BOF_CUDA(ctx, cudaEventRecord(ev, cs->push[0]));
  for (int i = 1; i < kPushStreams; ++i) BOF_CUDA(ctx, cudaStreamWaitEvent(cs->push[i], ev, 0));
  *xbuf_out = cs->xbuf;
  return BOF_OK;
}

// Push [offset, offset + count) floats of my exchange buffer into the same place of every peer's, after `ready` (an
// event of this device, e.g. "my upload landed"), then raise data flag `item` there.
int comm_push(bof_ctx* ctx, size_t offset, size_t count, int item, cudaEvent_t ready) {
  CommState* cs = ctx->comm;
  for (int d = 1; d < cs->world; ++d) {
    const int r = (cs->rank + d) % cs->world;      // every owner starts with a different peer
    cudaStream_t st = cs->push[r % kPushStreams];
    Peer& pr = cs->peers[r];
    if (ready) BOF_CUDA(ctx, cudaStreamWaitEvent(st, ready, 0));
    // the peer's buffer is free once the peer has entered this call
    if (cs->WaitValue32(st, (CUdeviceptr)(uintptr_t)(cs->flags + kFlagReady + r), cs->seq, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
      return fail(ctx, BOF_ECUDA, "cuStreamWaitValue32 failed");
    BOF_CUDA(ctx, cudaMemcpyPeerAsync(pr.xbuf + offset, pr.device, cs->xbuf + offset, ctx->device, count * sizeof(float), st));
    BOF_CUDA(ctx, cudaMemcpyPeerAsync(pr.flags + kFlagData + item, pr.device, cs->seq_dev, ctx->device, sizeof(uint32_t), st));
  }
  return BOF_OK;
}

// Make stream `s` wait until item `item` of the current call has arrived in my exchange buffer.
int comm_wait_item(bof_ctx* ctx, cudaStream_t s, int item) {
  CommState* cs = ctx->comm;
  if (cs->WaitValue32(s, (CUdeviceptr)(uintptr_t)(cs->flags + kFlagData + item), cs->seq, CU_STREAM_WAIT_VALUE_GEQ) != CUDA_SUCCESS)
    return fail(ctx, BOF_ECUDA, "cuStreamWaitValue32 failed");
  return BOF_OK;
}

// All my pushes have left (my exchange buffer may be reused); part of sync_all.
void comm_sync_pushes(bof_ctx* ctx) {
  if (!ctx->comm) return;
  for (int i = 0; i < kPushStreams; ++i)
    if (ctx->comm->push[i]) cudaStreamSynchronize(ctx->comm->push[i]);
}

void comm_destroy(bof_ctx* ctx) {
  if (!ctx || !ctx->comm) return;
  for (auto& pr : ctx->comm->peers) {
    if (pr.xbuf && pr.xbuf_ipc) cudaIpcCloseMemHandle(pr.xbuf);
    if (pr.flags && pr.flags_ipc) cudaIpcCloseMemHandle(pr.flags);
  }
  for (int i = 0; i < kPushStreams; ++i) if (ctx->comm->push[i]) cudaStreamDestroy(ctx->comm->push[i]);
  if (ctx->comm->xbuf) cudaFree(ctx->comm->xbuf);
  if (ctx->comm->flags) cudaFree(ctx->comm->flags);
  if (ctx->comm->gather_dev) cudaFree(ctx->comm->gather_dev);
  if (ctx->comm->seq_dev) cudaFree(ctx->comm->seq_dev);
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
  // peer exchange state
  cs->peers.resize((size_t)world);
  for (int i = 0; i < kPushStreams; ++i) BOF_CUDA(ctx, cudaStreamCreateWithFlags(&cs->push[i], cudaStreamNonBlocking));
  BOF_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&cs->flags), kFlagCount * sizeof(uint32_t)));
  BOF_CUDA(ctx, cudaMemset(cs->flags, 0, kFlagCount * sizeof(uint32_t)));
  BOF_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&cs->gather_dev), sizeof(PeerInfo) * world));
  BOF_CUDA(ctx, cudaMalloc(reinterpret_cast<void**>(&cs->seq_dev), 256));
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  if (cudaGetDriverEntryPoint("cuStreamWaitValue32", &fn, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
    cs->WaitValue32 = reinterpret_cast<decltype(cs->WaitValue32)>(fn);
  if (!cs->WaitValue32) cudaGetLastError();
  if (world > 1) BOF_TRY(publish_and_map(ctx));   // flags of every peer (no exchange buffer yet)
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
