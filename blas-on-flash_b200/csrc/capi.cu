// The C ABI of include/bof_b200.h: context, argument validation, the device-tile entry points and
// the host entry points (per-GPU CUDA-stream tile pipelines: host buffer -> H2D -> kernel -> D2H
// -> host buffer).  The pipelines replace, for this path, the reference's scheduler / cache /
// io_executor / file_handle stack (src/scheduler/*.cpp, src/file_handles/*.cpp): residency is
// structural (every output row block is produced by one pass on one GPU), ordering is expressed
// with CUDA events instead of task parents and overlap checks.
#include "common.cuh"

#include <unistd.h>

#include <algorithm>
#include <chrono>
#include <cstdlib>
#include <condition_variable>
#include <deque>
#include <functional>
#include <map>
#include <thread>

using namespace bof;

namespace bof {
// Minimal fork-join pool for the host side of the staging copies (the reference's N_IO_THR threads).
class CopyPool {
 public:
  explicit CopyPool(int n) {
    n = std::max(1, n);
    for (int i = 0; i < n; ++i) workers_.emplace_back([this, i] { loop(i); });
  }
  ~CopyPool() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    for (auto& t : workers_) t.join();
  }
  int size() const { return (int)workers_.size(); }
  // fn(part) for part in [0, parts); returns when all parts are done.  parts <= size().
  void run(int parts, const std::function<void(int)>& fn) {
    if (parts <= 1) { fn(0); return; }
    std::unique_lock<std::mutex> lk(mu_);
    fn_ = &fn; parts_ = parts; pending_ = parts; ++epoch_;
    cv_.notify_all();
    done_cv_.wait(lk, [this] { return pending_ == 0; });
    fn_ = nullptr;
  }

 private:
  void loop(int id) {
    uint64_t seen = 0;
    for (;;) {
      const std::function<void(int)>* fn = nullptr;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || epoch_ != seen; });
        if (stop_) return;
        seen = epoch_;
        if (id >= parts_) continue;
        fn = fn_;
      }
      (*fn)(id);
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (--pending_ == 0) done_cv_.notify_all();
      }
    }
  }
  std::vector<std::thread> workers_;
  std::mutex mu_;
  std::condition_variable cv_, done_cv_;
  const std::function<void(int)>* fn_ = nullptr;
  int parts_ = 0, pending_ = 0;
  uint64_t epoch_ = 0;
  bool stop_ = false;
};

void drain_copy_out(bof_ctx* ctx, StageSlot* sl);  // defined below (needs host_rows_copy)
void trace_host(bof_ctx* ctx, const char* what, long idx);  // BOF_TRACE wall-clock mark (any thread)

// One background thread per context runs the device->host side of the pageable path -- the writer half of the
// reference's IoExecutor threads.  The calling thread only describes a transfer (D2HJob) and goes on staging
// uploads; this thread enqueues the chunked copies into the pinned ring on the job's stream (after `wait_ev`),
// records `record_ev` behind them, and copies every chunk out to the caller's buffer once its DMA has landed
// (the copy-out of chunk c overlaps the DMA of the chunks after it).  A consumer of `record_ev` first calls
// wait_issued(ticket): CUDA ignores waits on events that have not been recorded yet.
struct D2HJob {
  char* host = nullptr; size_t hpitch = 0;
  const char* dev = nullptr; size_t dpitch = 0;
  size_t width = 0, rows = 0;
  bool flat = false;
  cudaStream_t s = nullptr;
  cudaEvent_t wait_ev = nullptr, record_ev = nullptr;
  uint64_t id = 0;
};

class Drainer {
 public:
  explicit Drainer(bof_ctx* ctx) : ctx_(ctx), th_([this] { loop(); }) {}
  ~Drainer() {
    {
      std::lock_guard<std::mutex> lk(mu_);
      stop_ = true;
    }
    cv_.notify_all();
    th_.join();
  }
  uint64_t push(D2HJob job) {
    uint64_t id;
    {
      std::lock_guard<std::mutex> lk(mu_);
      id = job.id = ++pushed_;
      q_.push_back(job);
    }
    cv_.notify_one();
    return id;
  }
  void wait_issued(uint64_t id) {
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] { return issued_ >= id; });
  }
  void wait_all_issued() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] { return issued_ >= pushed_; });
  }
  // every pushed transfer has reached the caller's memory; returns false if one of them failed since the last call
  bool wait_idle() {
    std::unique_lock<std::mutex> lk(mu_);
    cv_done_.wait(lk, [&] { return q_.empty() && !busy_; });
    const bool ok = ok_;
    ok_ = true;
    return ok;
  }

 private:
  void drain(StageSlot& sl) {
    if (!sl.in_flight) return;
    if (cudaEventSynchronize(sl.ev) == cudaSuccess) drain_copy_out(ctx_, &sl);
    else { cudaGetLastError(); failed_ = true; }
    sl.in_flight = false;
    sl.out_dst = nullptr;
  }
  void run(const D2HJob& j) {
    std::vector<StageSlot>& ring = ctx_->stage_out;  // touched by this thread only
    const size_t cap = ctx_->cfg.stage_bytes, total = j.width * j.rows;
    // a flat transfer is cut into cap-sized pseudo rows
    const size_t w = j.flat ? std::min(cap, total) : j.width;
    const size_t total_rows = j.flat ? (total + w - 1) / w : j.rows;
    const size_t rows_per_chunk = std::max<size_t>(1, cap / w);
    bool ok = true;
    trace_host(ctx_, "drainer: job start", (long)j.id);
    if (j.wait_ev) ok = cudaStreamWaitEvent(j.s, j.wait_ev, 0) == cudaSuccess && ok;
    for (size_t r0 = 0; r0 < total_rows; r0 += rows_per_chunk) {
      const size_t rows = std::min(rows_per_chunk, total_rows - r0);
      StageSlot& sl = ring[next_slot_++ % ring.size()];
      drain(sl);
      if (j.flat) {
        const size_t bytes = std::min(total - r0 * w, rows * w);  // the last pseudo row may be short
        ok = cudaMemcpyAsync(sl.ptr, j.dev + r0 * w, bytes, cudaMemcpyDeviceToHost, j.s) == cudaSuccess && ok;
        sl.out_dst = j.host + r0 * w; sl.out_pitch = bytes; sl.out_width = bytes; sl.out_rows = 1;
      } else {
        ok = cudaMemcpy2DAsync(sl.ptr, w, j.dev + r0 * j.dpitch, j.dpitch, w, rows, cudaMemcpyDeviceToHost, j.s) == cudaSuccess && ok;
        sl.out_dst = j.host + r0 * j.hpitch; sl.out_pitch = j.hpitch; sl.out_width = w; sl.out_rows = rows;
      }
      ok = cudaEventRecord(sl.ev, j.s) == cudaSuccess && ok;
      sl.in_flight = true;
    }
    if (j.record_ev) ok = cudaEventRecord(j.record_ev, j.s) == cudaSuccess && ok;
    trace_host(ctx_, "drainer: job issued", (long)j.id);
    if (!ok) { cudaGetLastError(); failed_ = true; }
  }
  void loop() {
    cudaSetDevice(ctx_->device);
    for (;;) {
      D2HJob job;
      {
        std::unique_lock<std::mutex> lk(mu_);
        cv_.wait(lk, [&] { return stop_ || !q_.empty(); });
        if (q_.empty()) return;  // stop requested and nothing left
        job = q_.front();
        q_.pop_front();
        busy_ = true;
      }
      run(job);
      bool more;
      {
        std::lock_guard<std::mutex> lk(mu_);
        issued_ = job.id;
        more = !q_.empty();
      }
      cv_done_.notify_all();
      if (!more)  // nothing queued behind it: finish the chunks still in flight (a later job would recycle them)
        for (auto& sl : ctx_->stage_out) drain(sl);
      {
        std::lock_guard<std::mutex> lk(mu_);
        if (failed_) { ok_ = false; failed_ = false; }
        busy_ = false;
      }
      cv_done_.notify_all();
    }
  }
  bof_ctx* ctx_;
  std::mutex mu_;
  std::condition_variable cv_, cv_done_;
  std::deque<D2HJob> q_;
  uint64_t pushed_ = 0, issued_ = 0;
  size_t next_slot_ = 0;
  bool stop_ = false, busy_ = false, ok_ = true, failed_ = false;
  std::thread th_;
};
}  // namespace bof

namespace {

constexpr int kGemmRing = 5;  // P/C block generations in flight in bof_host_gemm

thread_local std::string g_create_err;

#define BOF_TRY(expr)          \
  do {                         \
    int rc__ = (expr);         \
    if (rc__ != BOF_OK) return rc__; \
  } while (0)

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}

bool is_nt(char c) { return c == 'N' || c == 'T'; }
bool is_rc(char c) { return c == 'R' || c == 'C'; }

cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// slots of the context arena
enum Slot {
  S_DENSE = 0,     // resident dense operand (B of csrmm, x of csrgemv, Q source of gemm)
  S_DENSE_T,       // its transposed / split form
  S_BLK0 = 2,      // per-block buffers, two generations each (b = 0/1 added to the slot id)
  S_OFFS = 2, S_IDX64 = 4, S_IDX32 = 6, S_VALS = 8, S_CBLK = 10, S_CBLK_T = 12,
  S_WS = 18,       // kernel workspaces
  S_OUT0 = 19, S_OUT1, S_OUT2, S_MISC,
  // host gemm: ring of kGemmRing generations (g added to the slot id)
  S_PRAW = 24, S_PPLANES = 30, S_GCBLK = 32,
};

cudaEvent_t get_event(bof_ctx* ctx, size_t i) {
  while (ctx->events.size() <= i) {
    cudaEvent_t e = nullptr;
    cudaEventCreateWithFlags(&e, cudaEventDisableTiming);
    ctx->events.push_back(e);
  }
  return ctx->events[i];
}

// ---- optional timeline of a host pipeline (BOF_TRACE=1) ----
bool trace_on() {
  static const bool on = getenv("BOF_TRACE") != nullptr;
  return on;
}
void trace_mark(bof_ctx* ctx, cudaStream_t s, const char* what, int idx) {
  if (!trace_on()) return;
  if (ctx->trace.empty()) ctx->trace_t0 = now_ms();
  cudaEvent_t e = nullptr;
  if (!ctx->trace_pool.empty()) { e = ctx->trace_pool.back(); ctx->trace_pool.pop_back(); }
  else if (cudaEventCreate(&e) != cudaSuccess) return;
  cudaEventRecord(e, s);
  ctx->trace.push_back({e, what, idx});
}
}  // namespace
namespace bof {
void trace_host(bof_ctx* ctx, const char* what, long idx) {
  if (!trace_on()) return;
  std::lock_guard<std::mutex> lk(ctx->host_trace_mu);
  ctx->host_trace.push_back({now_ms() - ctx->trace_t0, what, idx});
}
}  // namespace bof
namespace {
void trace_dump(bof_ctx* ctx, const char* title) {
  if (trace_on()) {
    std::lock_guard<std::mutex> lk(ctx->host_trace_mu);
    for (auto& h : ctx->host_trace) std::fprintf(stderr, "[bof host ] %9.3f ms  %s %ld\n", h.ms, h.what, h.idx);
    ctx->host_trace.clear();
  }
  if (!trace_on() || ctx->trace.empty()) return;
  std::vector<std::pair<float, size_t>> order;
  for (size_t i = 0; i < ctx->trace.size(); ++i) {
    float ms = 0.f;
    cudaEventElapsedTime(&ms, ctx->trace[0].ev, ctx->trace[i].ev);
    order.push_back({ms, i});
  }
  std::sort(order.begin(), order.end());
  std::fprintf(stderr, "[bof trace] %s\n", title);
  for (auto& o : order) std::fprintf(stderr, "[bof trace] %9.3f ms  %s %d\n", o.first, ctx->trace[o.second].what, ctx->trace[o.second].idx);
  for (auto& m : ctx->trace) ctx->trace_pool.push_back(m.ev);
  ctx->trace.clear();
}

bool host_is_pinned(const void* p) {
  cudaPointerAttributes attr{};
  if (cudaPointerGetAttributes(&attr, p) != cudaSuccess) {
    cudaGetLastError();
    return false;
  }
  return attr.type == cudaMemoryTypeHost || attr.type == cudaMemoryTypeManaged;
}

int ensure_ring(bof_ctx* ctx, std::vector<StageSlot>& ring);
// Both rings are created at the first staged transfer of a context: cudaMallocHost in the middle of a pipeline waits
// for the device (147 ms behind the prologue kernels at 32768^3, BOF_TRACE).
int ensure_rings(bof_ctx* ctx) {
  BOF_TRY(ensure_ring(ctx, ctx->stage_in));
  return ensure_ring(ctx, ctx->stage_out);
}
int ensure_ring(bof_ctx* ctx, std::vector<StageSlot>& ring) {
  if (!ring.empty()) return BOF_OK;
  ring.resize((size_t)ctx->cfg.n_stage_bufs);
  for (auto& sl : ring) {
    if (cudaMallocHost(&sl.ptr, ctx->cfg.stage_bytes) != cudaSuccess ||
        cudaEventCreateWithFlags(&sl.ev, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      for (auto& u : ring) {  // a half-built ring must not be mistaken for a usable one by the next call
        if (u.ptr) cudaFreeHost(u.ptr);
        if (u.ev) cudaEventDestroy(u.ev);
      }
      ring.clear();
      return fail(ctx, BOF_ENOMEM, "cudaMallocHost of a %llu-byte staging buffer failed",
                  (unsigned long long)ctx->cfg.stage_bytes);
    }
  }
  if (!ctx->pool) ctx->pool = new CopyPool(ctx->cfg.n_copy_threads);
  if (!ctx->pool_out) ctx->pool_out = new CopyPool(ctx->cfg.n_copy_threads);
  if (!ctx->drainer) ctx->drainer = new Drainer(ctx);
  return BOF_OK;
}

// File-backed host ranges registered by the flash:: layer (map_file): base address -> (length, fd, file offset
// of the base).  With BOF_STAGE_FD=1 staging copies for such ranges use pread/pwrite on the descriptor instead
// of touching the mapping (the reference's FlashFileHandle::read/write into cache buffers); the page cache keeps
// both views coherent.
struct FileRange { size_t len; int fd; uint64_t file_off; };
std::mutex g_map_mu;
std::map<uintptr_t, FileRange> g_mappings;

bool lookup_mapping(const void* p, size_t bytes, int* fd, uint64_t* file_off) {
  std::lock_guard<std::mutex> lk(g_map_mu);
  const uintptr_t a = reinterpret_cast<uintptr_t>(p);
  auto it = g_mappings.upper_bound(a);
  if (it == g_mappings.begin()) return false;
  --it;
  if (a < it->first || a + bytes > it->first + it->second.len) return false;
  *fd = it->second.fd;
  *file_off = it->second.file_off + (a - it->first);
  return true;
}

bool file_xfer(bool write, int fd, char* buf, size_t len, uint64_t off) {
  while (len > 0) {
    const ssize_t n = write ? ::pwrite(fd, buf, len, (off_t)off) : ::pread(fd, buf, len, (off_t)off);
    if (n < 0 && errno == EINTR) continue;
    if (n <= 0) return false;
    buf += n; off += (uint64_t)n; len -= (size_t)n;
  }
  return true;
}

// rows x width bytes between a pitched host matrix and a tightly packed staging slot, split over the pool
void host_rows_copy(bof_ctx* ctx, char* packed, char* host, size_t hpitch, size_t width, size_t rows, bool to_packed) {
  const double t0 = now_ms();
  const size_t total = width * rows;
  CopyPool* pool = to_packed ? ctx->pool : ctx->pool_out;  // the two directions run on different threads
  const int parts = (int)std::min<size_t>((size_t)pool->size(), std::max<size_t>(1, total >> 20));
  int fd = -1;
  uint64_t foff = 0;
  const size_t span = rows == 0 ? 0 : (rows - 1) * hpitch + width;
  // Measured on the B200 boxes with page-cache-resident files (profiles/r01/trip12_driver_ab.txt): 8 workers
  // copying through the mapping move 44 GB/s in and 19 GB/s out, pread/pwrite 29 / 5.7 GB/s (pwrite to tmpfs
  // is the slow one).  The descriptor path therefore stays opt-in (BOF_STAGE_FD=1) for cold files on real disks,
  // where explicit large reads beat 4 KiB fault-driven readahead.
  static const bool use_fd = getenv("BOF_STAGE_FD") != nullptr;
  const bool via_fd = use_fd && lookup_mapping(host, span, &fd, &foff);
  pool->run(parts, [&](int part) {
    if (hpitch == width) {  // flat: split by bytes
      const size_t b0 = total * part / parts, b1 = total * (part + 1) / parts;
      if (via_fd && file_xfer(!to_packed, fd, packed + b0, b1 - b0, foff + b0)) return;
      if (to_packed) std::memcpy(packed + b0, host + b0, b1 - b0);
      else std::memcpy(host + b0, packed + b0, b1 - b0);
    } else {
      const size_t r0 = rows * part / parts, r1 = rows * (part + 1) / parts;
      for (size_t r = r0; r < r1; ++r) {
        if (via_fd && file_xfer(!to_packed, fd, packed + r * width, width, foff + r * hpitch)) continue;
        if (to_packed) std::memcpy(packed + r * width, host + r * hpitch, width);
        else std::memcpy(host + r * hpitch, packed + r * width, width);
      }
    }
  });
  (to_packed ? ctx->stats.stage_in_ms : ctx->stats.stage_out_ms) += now_ms() - t0;
}

}  // namespace
namespace bof {
void drain_copy_out(bof_ctx* ctx, StageSlot* sl) {
  if (sl->out_dst)
    host_rows_copy(ctx, static_cast<char*>(sl->ptr), sl->out_dst, sl->out_pitch, sl->out_width, sl->out_rows, false);
}
}  // namespace bof
namespace {
// host -> device slots are recycled by the calling thread once their DMA has finished
int drain_slot(bof_ctx* ctx, StageSlot& sl) {
  if (!sl.in_flight) return BOF_OK;
  BOF_CUDA(ctx, cudaEventSynchronize(sl.ev));
  sl.in_flight = false;
  return BOF_OK;
}

// Pageable host memory (e.g. the mmap behind a flash_ptr) -> device through the pinned ring: the host memcpy of
// chunk i+1 overlaps the DMA of chunk i.  Blocks the calling thread, like the reference's synchronous
// FlashFileHandle::read into a cache buffer, but keeps the copy engine at work.
int staged_upload(bof_ctx* ctx, void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                  cudaStream_t s) {
  std::vector<StageSlot>& ring = ctx->stage_in;
  BOF_TRY(ensure_rings(ctx));
  const size_t cap = ctx->cfg.stage_bytes;
  // view the transfer as rows of `w` bytes; a flat transfer is cut into cap-sized pseudo rows
  const bool flat = (dpitch == width && spitch == width) || height == 1;
  const size_t w = flat ? std::min(cap, width * height) : width;
  const size_t total_rows = flat ? ceil_div<size_t>(width * height, w) : height;
  const size_t rows_per_chunk = std::max<size_t>(1, cap / w);
  const size_t flat_bytes = width * height;
  char* host = const_cast<char*>(static_cast<const char*>(src));
  char* dev = static_cast<char*>(dst);
  size_t slot_i = 0;
  for (size_t r0 = 0; r0 < total_rows; r0 += rows_per_chunk, ++slot_i) {
    const size_t rows = std::min(rows_per_chunk, total_rows - r0);
    StageSlot& sl = ring[slot_i % ring.size()];
    BOF_TRY(drain_slot(ctx, sl));
    if (flat) {
      const size_t bytes = std::min(flat_bytes - r0 * w, rows * w);  // the last pseudo row may be short
      host_rows_copy(ctx, static_cast<char*>(sl.ptr), host + r0 * w, bytes, bytes, 1, true);
      BOF_CUDA(ctx, cudaMemcpyAsync(dev + r0 * w, sl.ptr, bytes, cudaMemcpyHostToDevice, s));
    } else {
      host_rows_copy(ctx, static_cast<char*>(sl.ptr), host + r0 * spitch, spitch, w, rows, true);
      BOF_CUDA(ctx, cudaMemcpy2DAsync(dev + r0 * dpitch, dpitch, sl.ptr, w, w, rows, cudaMemcpyHostToDevice, s));
    }
    BOF_CUDA(ctx, cudaEventRecord(sl.ev, s));
    sl.in_flight = true;
  }
  return BOF_OK;
}

bool wants_staging(const bof_ctx* ctx, const void* host, size_t dpitch, size_t spitch, size_t width, size_t height) {
  const bool flat = (dpitch == width && spitch == width) || height == 1;
  return width * height >= (256u << 10) && (flat || width <= ctx->cfg.stage_bytes) && !host_is_pinned(host);
}

// Device -> host transfer on stream `s`, ordered after `wait_ev` (may be null); `record_ev` (may be null) is
// recorded on `s` behind it.  A pinned destination is enqueued right here.  A pageable one is handed to the
// drainer thread, which enqueues it chunk by chunk through the pinned ring while the calling thread goes on;
// *ticket then identifies the transfer and d2h_fence(ticket) must precede any use of `record_ev` (and any
// re-recording of `wait_ev`).  sync_all() completes every transfer.
int d2h_transfer(bof_ctx* ctx, void* dst, size_t dpitch, const void* src, size_t spitch, size_t width, size_t height,
                 cudaStream_t s, cudaEvent_t wait_ev, cudaEvent_t record_ev, uint64_t* ticket) {
  if (ticket) *ticket = 0;
  const bool empty = width == 0 || height == 0;
  ctx->stats.d2h_bytes += (double)width * height;
  if (empty || !wants_staging(ctx, dst, dpitch, spitch, width, height)) {
    if (wait_ev) BOF_CUDA(ctx, cudaStreamWaitEvent(s, wait_ev, 0));
    if (!empty) {
      if (dpitch == width && spitch == width) BOF_CUDA(ctx, cudaMemcpyAsync(dst, src, width * height, cudaMemcpyDeviceToHost, s));
      else BOF_CUDA(ctx, cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, cudaMemcpyDeviceToHost, s));
    }
    if (record_ev) BOF_CUDA(ctx, cudaEventRecord(record_ev, s));
    return BOF_OK;
  }
  BOF_TRY(ensure_rings(ctx));
  D2HJob j;
  j.host = static_cast<char*>(dst); j.hpitch = dpitch;
  j.dev = static_cast<const char*>(src); j.dpitch = spitch;
  j.width = width; j.rows = height;
  j.flat = (dpitch == width && spitch == width) || height == 1;
  j.s = s; j.wait_ev = wait_ev; j.record_ev = record_ev;
  const uint64_t id = ctx->drainer->push(j);
  if (ticket) *ticket = id;
  return BOF_OK;
}

void d2h_fence(bof_ctx* ctx, uint64_t ticket) {
  if (ticket == 0 || !ctx->drainer) return;
  trace_host(ctx, "caller: fence enter", (long)ticket);
  ctx->drainer->wait_issued(ticket);
  trace_host(ctx, "caller: fence leave", (long)ticket);
}

// pitched host<->device copy; collapses to a flat copy when both sides are tight.  Pinned host memory
// is copied asynchronously in place; pageable memory goes through the staging rings.
int copy2d(bof_ctx* ctx, void* dst, size_t dpitch, const void* src, size_t spitch, size_t width,
           size_t height, cudaMemcpyKind kind, cudaStream_t s) {
  if (width == 0 || height == 0) return BOF_OK;
  if (kind == cudaMemcpyDeviceToHost) return d2h_transfer(ctx, dst, dpitch, src, spitch, width, height, s, nullptr, nullptr, nullptr);
  if (kind == cudaMemcpyHostToDevice) {
    ctx->stats.h2d_bytes += (double)width * height;
    if (wants_staging(ctx, src, dpitch, spitch, width, height)) return staged_upload(ctx, dst, dpitch, src, spitch, width, height, s);
  }
  if (dpitch == width && spitch == width) {
    BOF_CUDA(ctx, cudaMemcpyAsync(dst, src, width * height, kind, s));
  } else {
    BOF_CUDA(ctx, cudaMemcpy2DAsync(dst, dpitch, src, spitch, width, height, kind, s));
  }
  return BOF_OK;
}
int copy1d(bof_ctx* ctx, void* dst, const void* src, size_t bytes, cudaMemcpyKind kind, cudaStream_t s) {
  return copy2d(ctx, dst, bytes, src, bytes, bytes, 1, kind, s);
}

void stats_begin(bof_ctx* ctx) {
  ctx->stats = bof_stats{};
  ctx->stats.total_ms = -now_ms();
  ctx->stats.kernel_launches = -ctx->launches.load();
}
void stats_end(bof_ctx* ctx) {
  ctx->stats.total_ms += now_ms();
  ctx->stats.kernel_launches += ctx->launches.load();
}

// all staged device->host chunks have reached the caller's buffer
int drain_wait(bof_ctx* ctx) {
  if (ctx->drainer && !ctx->drainer->wait_idle()) return fail(ctx, BOF_ECUDA, "device->host staging failed");
  return BOF_OK;
}

int sync_all(bof_ctx* ctx) {
  if (ctx->drainer) ctx->drainer->wait_all_issued();  // the drainer may still be enqueueing copies on the streams
  BOF_CUDA(ctx, cudaStreamSynchronize(ctx->h2d));
  BOF_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
  BOF_CUDA(ctx, cudaStreamSynchronize(ctx->d2h));
  return drain_wait(ctx);
}

// After a failed call nothing of it may still be running: queued copies and the drainer thread reference the
// caller's host buffers and the context's slots.  Keeps the recorded error message.
void quiesce(bof_ctx* ctx) {
  if (ctx->drainer) ctx->drainer->wait_all_issued();
  cudaStreamSynchronize(ctx->h2d);
  cudaStreamSynchronize(ctx->compute);
  cudaStreamSynchronize(ctx->d2h);
  if (ctx->drainer) ctx->drainer->wait_idle();
  for (auto& sl : ctx->stage_in) sl.in_flight = false;
  cudaGetLastError();
}

// Every host entry point holds one: any return that did not set `ok` leaves the context quiescent.
struct CallGuard {
  bof_ctx* ctx;
  bool ok = false;
  explicit CallGuard(bof_ctx* c) : ctx(c) {}
  ~CallGuard() { if (!ok && ctx) quiesce(ctx); }
  int done() { ok = true; return BOF_OK; }
};


// Canonical form of a GEMM: Cout[Mo x No] (row-major, ldc) = P[Mo x K] * Q[No x K]^T where
// element (r, kk) of P is psrc[r*p_sr + kk*p_sk] (one of the strides is 1), same for Q.
struct Canon {
  int64_t Mo, No, K;
  const float* psrc; int64_t p_sr, p_sk;
  const float* qsrc; int64_t q_sr, q_sk;
  int64_t ldc;
};

// Leading-dimension defaults and the row/col role table of src/blas/gemm.cpp:52-67.
int canon_gemm(bof_ctx* ctx, char ord, char ta, char tb, int64_t m, int64_t n, int64_t k,
               const float* A, int64_t lda, const float* B, int64_t ldb, int64_t ldc, Canon* out) {
  BOF_REQUIRE(ctx, is_rc(ord), "gemm: mat_ord must be 'R' or 'C' (got '%c')", ord);
  BOF_REQUIRE(ctx, is_nt(ta), "gemm: trans_a must be 'N' or 'T' (got '%c')", ta);
  BOF_REQUIRE(ctx, is_nt(tb), "gemm: trans_b must be 'N' or 'T' (got '%c')", tb);
  BOF_REQUIRE(ctx, m >= 0 && n >= 0 && k >= 0, "gemm: negative dimension");
  const bool col = ord == 'C', tA = ta == 'T', tB = tb == 'T';
  const int64_t a_cols = (tA != col) ? m : k;  // contiguous extent of A as stored
  const int64_t b_cols = (tB != col) ? k : n;
  const int64_t c_cols = col ? m : n;
  if (lda == 0) lda = a_cols;
  if (ldb == 0) ldb = b_cols;
  if (ldc == 0) ldc = c_cols;
  BOF_REQUIRE(ctx, lda >= a_cols && ldb >= b_cols && ldc >= c_cols, "gemm: leading dimension too small");
  // op(A)(i, kk) = A[i*a_si + kk*a_sk]: contiguous in kk iff A is stored with k as its inner extent
  const int64_t a_si = (tA == col) ? lda : 1, a_sk = (tA == col) ? 1 : lda;
  // op(B)(kk, j) = B[j*b_sj + kk*b_sk]: contiguous in kk iff B is stored with k as its inner extent
  const int64_t b_sj = (tB != col) ? ldb : 1, b_sk = (tB != col) ? 1 : ldb;
  Canon c{};
  c.K = k;
  c.ldc = ldc;
  if (!col) {  // C[i*ldc + j]
    c.Mo = m; c.No = n;
    c.psrc = A; c.p_sr = a_si; c.p_sk = a_sk;
    c.qsrc = B; c.q_sr = b_sj; c.q_sk = b_sk;
  } else {     // column-major C is the row-major transpose: C^T = op(B)^T op(A)^T
    c.Mo = n; c.No = m;
    c.psrc = B; c.p_sr = b_sj; c.p_sk = b_sk;
    c.qsrc = A; c.q_sr = a_si; c.q_sk = a_sk;
  }
  *out = c;
  return BOF_OK;
}

int64_t padded_k(int64_t k) { return std::max<int64_t>(32, round_up<int64_t>(k, 32)); }
size_t plane_bytes(int64_t rows, int64_t kp) { return round_up<size_t>((size_t)rows * kp * 4, 256); }

int pick_gemm_path(const bof_ctx* ctx, int64_t Mo, int64_t No, int64_t K) {
  if (ctx->cfg.gemm_force_path) return ctx->cfg.gemm_force_path;
  if ((double)Mo * No * K < 2e6) return 3;  // launch-latency territory: CUDA cores, no planes
  return 2;
}

int64_t k_chunk_of(const bof_ctx* ctx) {
  if (ctx->cfg.gemm_k_chunk < 0) return 0;
  return ctx->cfg.gemm_k_chunk == 0 ? 256 : ctx->cfg.gemm_k_chunk;
}

// GEMM on device-resident canonical operands with a caller-provided plane workspace.
int gemm_canon_device(bof_ctx* ctx, cudaStream_t s, const Canon& c, float alpha, float beta, float* C,
                      void* ws, size_t ws_bytes) {
  if (c.Mo == 0 || c.No == 0) return BOF_OK;
  const int path = pick_gemm_path(ctx, c.Mo, c.No, c.K);
  if (c.K == 0 || path == 3) {
    // K == 0 degenerates to C = beta*C, which the CUDA-core kernel handles as well
    return launch_gemm_ffma(ctx, s, c.Mo, c.No, c.K, alpha, c.psrc, c.p_sr, c.p_sk, c.qsrc, c.q_sk, c.q_sr,
                            beta, C, c.ldc);
  }
  const int64_t kp = padded_k(c.K);
  const size_t pb = plane_bytes(c.Mo, kp), qb = plane_bytes(c.No, kp);
  BOF_REQUIRE(ctx, ws != nullptr && ws_bytes >= 2 * pb + 2 * qb + 256, "gemm: workspace too small");
  uint8_t* base = reinterpret_cast<uint8_t*>(round_up<uintptr_t>(reinterpret_cast<uintptr_t>(ws), 256));
  float* p_hi = reinterpret_cast<float*>(base);
  float* p_lo = reinterpret_cast<float*>(base + pb);
  float* q_hi = reinterpret_cast<float*>(base + 2 * pb);
  float* q_lo = reinterpret_cast<float*>(base + 2 * pb + qb);
  BOF_TRY(launch_split_planes(ctx, s, c.Mo, c.K, c.psrc, c.p_sr, c.p_sk, p_hi, p_lo, kp));
  BOF_TRY(launch_split_planes(ctx, s, c.No, c.K, c.qsrc, c.q_sr, c.q_sk, q_hi, q_lo, kp));
  GemmEpilogue ep;
  ep.alpha = alpha; ep.beta = beta; ep.C = C; ep.ldc = c.ldc;
  return launch_gemm_tc(ctx, s, path == 1 ? 1 : 2, c.Mo, c.No, c.K, kp, p_hi, p_lo, q_hi, q_lo, ep, k_chunk_of(ctx));
}

// Row-block partition by nnz budget (the idea of get_next_blk_size, include/blas_utils.h:72-82):
// blocks[i] .. blocks[i+1] are the rows of block i.
std::vector<int64_t> partition_rows(const int64_t* ia, int64_t m, int64_t max_nnz) {
  std::vector<int64_t> cuts{0};
  int64_t r = 0;
  while (r < m) {
    const int64_t limit = ia[r] + max_nnz;
    int64_t e = std::upper_bound(ia + r + 1, ia + m + 1, limit) - ia - 1;  // last e with ia[e] <= limit
    if (e <= r) e = r + 1;  // a single row above the budget still forms a block
    cuts.push_back(e);
    r = e;
  }
  return cuts;
}

}  // namespace

namespace bof {
int slot_reserve(bof_ctx* ctx, int slot, size_t bytes, void** out) {
  if (bytes == 0) bytes = 256;
  if (ctx->slot_bytes[slot] < bytes) {
    if (ctx->slot_ptr[slot]) {
      // the buffer may still be in use by queued work of a previous call
      BOF_CUDA(ctx, cudaDeviceSynchronize());
      BOF_CUDA(ctx, cudaFree(ctx->slot_ptr[slot]));
      ctx->slot_ptr[slot] = nullptr;
      ctx->slot_bytes[slot] = 0;
    }
    cudaError_t e = cudaMalloc(&ctx->slot_ptr[slot], bytes);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return fail(ctx, BOF_ENOMEM, "cudaMalloc of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    ctx->slot_bytes[slot] = bytes;
  }
  *out = ctx->slot_ptr[slot];
  return BOF_OK;
}
}  // namespace bof

extern "C" {

int bof_abi_version(void) { return 1; }

int bof_ctx_create(const bof_config* cfg, bof_ctx** out) {
  if (!out) return BOF_EINVAL;
  *out = nullptr;
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0) {
    cudaGetLastError();
    g_create_err = std::string("no CUDA device: ") + (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0");
    return BOF_ENODEV;
  }
  bof_ctx* ctx = new bof_ctx();
  if (cfg) ctx->cfg = *cfg;
  ctx->device = ctx->cfg.device;
  auto bail = [&](int code, const std::string& msg) {
    g_create_err = msg;
    delete ctx;
    return code;
  };
  if (ctx->device < 0 || ctx->device >= ndev) return bail(BOF_EINVAL, "device ordinal out of range");
  if ((e = cudaSetDevice(ctx->device)) != cudaSuccess) return bail(BOF_ECUDA, cudaGetErrorString(e));
  cudaDeviceProp prop{};
  if ((e = cudaGetDeviceProperties(&prop, ctx->device)) != cudaSuccess) return bail(BOF_ECUDA, cudaGetErrorString(e));
  if (prop.major != 10) {
    char buf[160];
    snprintf(buf, sizeof(buf), "device %d is sm_%d%d; this library only carries sm_100a code (B200)", ctx->device,
             prop.major, prop.minor);
    return bail(BOF_ENODEV, buf);
  }
  ctx->num_sms = prop.multiProcessorCount;
  ctx->l2_bytes = (size_t)prop.l2CacheSize;
  if (ctx->cfg.n_copy_threads <= 0) ctx->cfg.n_copy_threads = 8;
  if (ctx->cfg.stage_bytes == 0) ctx->cfg.stage_bytes = 16ull << 20;
  if (ctx->cfg.n_stage_bufs <= 0) ctx->cfg.n_stage_bufs = 4;
  if (ctx->cfg.csrmm_max_nnz == 0) ctx->cfg.csrmm_max_nnz = 64ull << 20;
  if (ctx->cfg.gemm_row_block == 0) ctx->cfg.gemm_row_block = 4096;
  cudaStreamCreateWithFlags(&ctx->compute, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&ctx->h2d, cudaStreamNonBlocking);
  cudaStreamCreateWithFlags(&ctx->d2h, cudaStreamNonBlocking);
  cudaDriverEntryPointQueryResult qres;
  void* fn = nullptr;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres) == cudaSuccess &&
      qres == cudaDriverEntryPointSuccess)
    ctx->tmap_encode = fn;
  else
    cudaGetLastError();
  *out = ctx;
  return BOF_OK;
}

int bof_ctx_destroy(bof_ctx* ctx) {
  if (!ctx) return BOF_OK;
  cudaSetDevice(ctx->device);
  cudaDeviceSynchronize();
  for (int i = 0; i < bof_ctx::kSlots; ++i)
    if (ctx->slot_ptr[i]) cudaFree(ctx->slot_ptr[i]);
  for (auto ev : ctx->events) cudaEventDestroy(ev);
  for (auto& m : ctx->trace) cudaEventDestroy(m.ev);
  for (auto ev : ctx->trace_pool) cudaEventDestroy(ev);
  delete ctx->drainer;  // joins its thread before the rings go away
  delete ctx->pool;
  delete ctx->pool_out;
  for (auto* ring : {&ctx->stage_in, &ctx->stage_out})
    for (auto& sl : *ring) {
      if (sl.ptr) cudaFreeHost(sl.ptr);
      if (sl.ev) cudaEventDestroy(sl.ev);
    }
  if (ctx->sync_ctr) cudaFree(ctx->sync_ctr);
  if (ctx->tk0) cudaEventDestroy(ctx->tk0);
  if (ctx->tk1) cudaEventDestroy(ctx->tk1);
  if (ctx->compute) cudaStreamDestroy(ctx->compute);
  if (ctx->h2d) cudaStreamDestroy(ctx->h2d);
  if (ctx->d2h) cudaStreamDestroy(ctx->d2h);
  delete ctx;
  return BOF_OK;
}

const char* bof_last_error(const bof_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_err.c_str(); }

int bof_get_stats(const bof_ctx* ctx, bof_stats* out) {
  if (!ctx || !out) return BOF_EINVAL;
  *out = ctx->stats;
  // kernel_ms: device time of the most recent tensor-core GEMM kernel, once it has finished
  if (ctx->tk_valid && cudaEventQuery(ctx->tk1) == cudaSuccess) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, ctx->tk0, ctx->tk1) == cudaSuccess) out->kernel_ms = ms;
  }
  cudaGetLastError();
  return BOF_OK;
}

int64_t bof_launch_count(const bof_ctx* ctx) { return ctx ? ctx->launches.load() : 0; }

int bof_register_mapping(const void* base, size_t len, int fd, uint64_t file_offset) {
  if (base == nullptr || len == 0 || fd < 0) return BOF_EINVAL;
  std::lock_guard<std::mutex> lk(g_map_mu);
  g_mappings[reinterpret_cast<uintptr_t>(base)] = FileRange{len, fd, file_offset};
  return BOF_OK;
}

int bof_unregister_mapping(const void* base) {
  std::lock_guard<std::mutex> lk(g_map_mu);
  return g_mappings.erase(reinterpret_cast<uintptr_t>(base)) ? BOF_OK : BOF_EINVAL;
}

// ---- device-tile entry points ---------------------------------------------------------------

size_t bof_spmm_workspace_bytes(char ord, int64_t m, int64_t n, int64_t k) {
  if (ord != 'C') return 0;
  return round_up<size_t>((size_t)n * k * 4, 256) + round_up<size_t>((size_t)m * k * 4, 256) + 256;
}

int bof_spmm_csr_f32(bof_ctx* ctx, void* stream, char ord, int64_t m, int64_t n, int64_t k, float alpha,
                     const float* vals, const int32_t* idx, const int64_t* offs, const float* B, int64_t ldb,
                     float beta, float* C, int64_t ldc, void* workspace, size_t workspace_bytes) {
  if (!ctx) return BOF_EINVAL;
  BOF_REQUIRE(ctx, is_rc(ord), "csrmm: ord_b must be 'R' or 'C' (got '%c')", ord);
  BOF_REQUIRE(ctx, m >= 0 && n >= 0 && k >= 0, "csrmm: negative dimension");
  BOF_REQUIRE(ctx, n < (1ll << 31), "csrmm: column count must fit int32 on the device");
  cudaStream_t s = as_stream(stream);
  if (m == 0 || k == 0) return BOF_OK;
  if (ord == 'R') {
    BOF_REQUIRE(ctx, ldb >= k && ldc >= k, "csrmm: leading dimension too small");
    return launch_spmm_rm(ctx, s, m, k, alpha, vals, idx, offs, B, ldb, beta, C, ldc);
  }
  // Column-major B (n x k, ldb >= n) and C (m x k, ldc >= m): transpose-on-stage around the
  // row-major kernel; the gathers need B rows contiguous.
  BOF_REQUIRE(ctx, ldb >= n && ldc >= m, "csrmm: leading dimension too small");
  BOF_REQUIRE(ctx, workspace && workspace_bytes >= bof_spmm_workspace_bytes(ord, m, n, k),
              "csrmm: column-major needs bof_spmm_workspace_bytes() of workspace");
  uint8_t* base = reinterpret_cast<uint8_t*>(round_up<uintptr_t>(reinterpret_cast<uintptr_t>(workspace), 256));
  float* Bt = reinterpret_cast<float*>(base);
  float* Ct = reinterpret_cast<float*>(base + round_up<size_t>((size_t)n * k * 4, 256));
  BOF_TRY(launch_transpose(ctx, s, k, n, B, ldb, Bt, k));
  BOF_TRY(launch_spmm_rm(ctx, s, m, k, 1.f, vals, idx, offs, Bt, k, 0.f, Ct, k));
  return launch_transpose_axpby(ctx, s, m, k, alpha, Ct, k, beta, C, ldc);
}

int bof_spmv_csr_f32(bof_ctx* ctx, void* stream, char trans, int64_t m, int64_t n, const float* vals,
                     const int32_t* idx, const int64_t* offs, const float* x, float* y) {
  if (!ctx) return BOF_EINVAL;
  BOF_REQUIRE(ctx, is_nt(trans), "csrgemv: trans_a must be 'N' or 'T' (got '%c')", trans);
  BOF_REQUIRE(ctx, m >= 0 && n >= 0 && n < (1ll << 31), "csrgemv: bad dimension");
  return launch_spmv(ctx, as_stream(stream), trans, m, n, vals, idx, offs, x, y);
}

int bof_idx_narrow(bof_ctx* ctx, void* stream, const int64_t* in, int32_t* out, int64_t count) {
  if (!ctx) return BOF_EINVAL;
  return launch_idx_narrow(ctx, as_stream(stream), in, out, count);
}
int bof_idx_widen(bof_ctx* ctx, void* stream, const int32_t* in, int64_t* out, int64_t count) {
  if (!ctx) return BOF_EINVAL;
  return launch_idx_widen(ctx, as_stream(stream), in, out, count);
}

size_t bof_sgemm_workspace_bytes(int64_t m, int64_t n, int64_t k) {
  const int64_t kp = padded_k(k);
  return 2 * plane_bytes(m, kp) + 2 * plane_bytes(n, kp) + 512;
}

int bof_sgemm_f32(bof_ctx* ctx, void* stream, char ord, char ta, char tb, int64_t m, int64_t n, int64_t k,
                  float alpha, const float* A, int64_t lda, const float* B, int64_t ldb, float beta, float* C,
                  int64_t ldc, void* workspace, size_t workspace_bytes) {
  if (!ctx) return BOF_EINVAL;
  Canon c;
  BOF_TRY(canon_gemm(ctx, ord, ta, tb, m, n, k, A, lda, B, ldb, ldc, &c));
  return gemm_canon_device(ctx, as_stream(stream), c, alpha, beta, C, workspace, workspace_bytes);
}

size_t bof_csr2csc_workspace_bytes(int64_t m, int64_t n, int64_t nnz) { return csr2csc_workspace_bytes(m, n, nnz); }

int bof_csr2csc(bof_ctx* ctx, void* stream, int64_t m, int64_t n, int64_t nnz, const int64_t* offs,
                const int32_t* idx, const float* vals, int64_t* offs_t, int32_t* idx_t, float* vals_t,
                void* workspace, size_t workspace_bytes) {
  if (!ctx) return BOF_EINVAL;
  BOF_REQUIRE(ctx, m >= 0 && n >= 0 && nnz >= 0, "csrcsc: negative dimension");
  return launch_csr2csc(ctx, as_stream(stream), m, n, nnz, offs, idx, vals, offs_t, idx_t, vals_t, workspace,
                        workspace_bytes);
}

int bof_row_sqnorm_f32(bof_ctx* ctx, void* stream, int64_t rows, int64_t dim, const float* X, int64_t ldx,
                       float* out) {
  if (!ctx) return BOF_EINVAL;
  BOF_REQUIRE(ctx, rows >= 0 && dim >= 0 && ldx >= dim, "row_sqnorm: bad dimension");
  return launch_row_sqnorm(ctx, as_stream(stream), rows, dim, X, ldx, out);
}

size_t bof_kmeans_point_planes_bytes(int64_t npoints, int64_t dim) { return 2 * plane_bytes(npoints, padded_k(dim)) + 256; }

size_t bof_kmeans_workspace_bytes(int64_t npoints, int64_t ncenters, int64_t dim, int with_point_planes) {
  size_t b = 2 * plane_bytes(ncenters, padded_k(dim)) + 512;
  if (with_point_planes) b += bof_kmeans_point_planes_bytes(npoints, dim);
  return b;
}

int bof_kmeans_prepare_points(bof_ctx* ctx, void* stream, int64_t npoints, int64_t dim, const float* points,
                              void* points_planes) {
  if (!ctx) return BOF_EINVAL;
  BOF_REQUIRE(ctx, npoints >= 0 && dim > 0 && points_planes, "kmeans: bad argument");
  const int64_t kp = padded_k(dim);
  uint8_t* base = reinterpret_cast<uint8_t*>(round_up<uintptr_t>(reinterpret_cast<uintptr_t>(points_planes), 256));
  return launch_split_planes(ctx, as_stream(stream), npoints, dim, points, dim, 1, reinterpret_cast<float*>(base),
                             reinterpret_cast<float*>(base + plane_bytes(npoints, kp)), kp);
}

int bof_kmeans_assign(bof_ctx* ctx, void* stream, int64_t npoints, int64_t ncenters, int64_t dim,
                      const float* points, const float* centers, const float* c_l2sq, const float* p_l2sq,
                      int32_t* assign, const void* points_planes, void* workspace, size_t workspace_bytes) {
  if (!ctx) return BOF_EINVAL;
  BOF_REQUIRE(ctx, npoints >= 0 && ncenters > 0 && dim > 0, "kmeans: bad dimension");
  BOF_REQUIRE(ctx, workspace && workspace_bytes >= bof_kmeans_workspace_bytes(npoints, ncenters, dim, points_planes == nullptr),
              "kmeans: workspace too small");
  if (npoints == 0) return BOF_OK;
  cudaStream_t s = as_stream(stream);
  const int64_t kp = padded_k(dim);
  uint8_t* base = reinterpret_cast<uint8_t*>(round_up<uintptr_t>(reinterpret_cast<uintptr_t>(workspace), 256));
  float* c_hi = reinterpret_cast<float*>(base);
  float* c_lo = reinterpret_cast<float*>(base + plane_bytes(ncenters, kp));
  BOF_TRY(launch_split_planes(ctx, s, ncenters, dim, centers, dim, 1, c_hi, c_lo, kp));
  const uint8_t* pp;
  if (points_planes) {
    pp = reinterpret_cast<const uint8_t*>(round_up<uintptr_t>(reinterpret_cast<uintptr_t>(points_planes), 256));
  } else {
    uint8_t* mine = base + 2 * plane_bytes(ncenters, kp);
    BOF_TRY(bof_kmeans_prepare_points(ctx, stream, npoints, dim, points, mine));
    pp = reinterpret_cast<const uint8_t*>(round_up<uintptr_t>(reinterpret_cast<uintptr_t>(mine), 256));
  }
  GemmEpilogue ep;
  ep.C = nullptr;
  ep.row_add = p_l2sq;
  ep.col_add = c_l2sq;
  ep.argmin_out = assign;
  const int cg = ctx->cfg.gemm_force_path == 1 ? 1 : 2;
  return launch_gemm_tc(ctx, s, cg, npoints, ncenters, dim, kp, reinterpret_cast<const float*>(pp),
                        reinterpret_cast<const float*>(pp + plane_bytes(npoints, kp)), c_hi, c_lo, ep, 0);
}

size_t bof_kmeans_reduce_workspace_bytes(int64_t npoints, int64_t ncenters, int64_t dim) {
  return kmeans_reduce_workspace_bytes(npoints, ncenters, dim);
}

int bof_kmeans_reduce(bof_ctx* ctx, void* stream, int64_t npoints, int64_t ncenters, int64_t dim,
                      const float* points, const int32_t* assign, float* sums, float* counts, void* workspace,
                      size_t workspace_bytes) {
  if (!ctx) return BOF_EINVAL;
  return launch_kmeans_reduce_ws(ctx, as_stream(stream), npoints, ncenters, dim, points, assign, sums, counts,
                                 workspace, workspace_bytes);
}

int bof_kmeans_finalize(bof_ctx* ctx, void* stream, int64_t ncenters, int64_t dim, const float* sums,
                        const float* counts, float* centers, float* c_l2sq) {
  if (!ctx) return BOF_EINVAL;
  return launch_kmeans_finalize(ctx, as_stream(stream), ncenters, dim, sums, counts, centers, c_l2sq);
}

// ---- host entry points -------------------------------------------------------------------------

// flash::csrmm.  'N': B is uploaded once and stays resident; A streams in nnz-balanced row blocks
// (offsets, indices, values), double-buffered: while block i runs, block i+1 uploads and block
// i-1's C rows download.  'T': whole-matrix csr2csc on the device, then the same kernel.
static int host_csrmm_impl(bof_ctx* ctx, char trans_a, int64_t m, int64_t n, int64_t k, float alpha, float beta,
                           const float* a, const int64_t* ia, const int64_t* ja, char ord_b, const float* b, float* c,
                           const float* b_dev) {
  if (!ctx) return BOF_EINVAL;
  BOF_REQUIRE(ctx, is_nt(trans_a), "csrmm: unrecognized value for param trans_a = '%c'", trans_a);
  BOF_REQUIRE(ctx, b_dev == nullptr || (trans_a == 'N' && ord_b == 'R'),
              "csrmm: a device-resident B is supported for trans_a='N', ord_b='R' only");
  BOF_REQUIRE(ctx, is_rc(ord_b), "csrmm: unrecognized value for param ord_b = '%c'", ord_b);
  BOF_REQUIRE(ctx, m >= 0 && n >= 0 && k >= 0 && m < (1ll << 31) && n < (1ll << 31), "csrmm: bad dimension");
  BOF_CUDA(ctx, cudaSetDevice(ctx->device));
  stats_begin(ctx);
  CallGuard call_guard(ctx);
  trace_mark(ctx, ctx->h2d, "start", 0);
  const int64_t out_rows = trans_a == 'N' ? m : n;   // rows of C
  const int64_t in_rows = trans_a == 'N' ? n : m;    // rows of B
  if (out_rows == 0 || k == 0) { stats_end(ctx); return call_guard.done(); }
  const bool colmaj = ord_b == 'C';
  const cudaMemcpyKind H2D = cudaMemcpyHostToDevice, D2H = cudaMemcpyDeviceToHost;

  // resident dense operand, always row-major [in_rows x k] on the device
  float* Bd = nullptr;
  if (b_dev != nullptr) {
    Bd = const_cast<float*>(b_dev);  // already in HBM (e.g. all-gathered over NVLink); read-only here
  } else {
    BOF_TRY(slot_reserve(ctx, S_DENSE, (size_t)in_rows * k, &Bd));
  }
  if (b_dev != nullptr) {
    // nothing to upload
  } else if (colmaj) {
    float* Braw = nullptr;
    BOF_TRY(slot_reserve(ctx, S_DENSE_T, (size_t)in_rows * k, &Braw));
    BOF_TRY(copy1d(ctx, Braw, b, (size_t)in_rows * k * 4, H2D, ctx->h2d));
    cudaEvent_t evB = get_event(ctx, 0);
    BOF_CUDA(ctx, cudaEventRecord(evB, ctx->h2d));
    BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, evB, 0));
    BOF_TRY(launch_transpose(ctx, ctx->compute, k, in_rows, Braw, in_rows, Bd, k));
  } else {
    BOF_TRY(copy1d(ctx, Bd, b, (size_t)in_rows * k * 4, H2D, ctx->h2d));
    cudaEvent_t evB = get_event(ctx, 0);
    BOF_CUDA(ctx, cudaEventRecord(evB, ctx->h2d));
    BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, evB, 0));
  }

  trace_mark(ctx, ctx->h2d, "h2d: dense operand landed", 0);
  const int64_t* offs_host = ia;
  const int64_t nnz = ia[m] - ia[0];
  std::vector<int64_t> tr_offs_host;  // only for 'T'
  const int32_t* idx_dev_all = nullptr;  // 'T': transposed matrix resident on the device
  const float* vals_dev_all = nullptr;
  const int64_t* offs_dev_all = nullptr;

  if (trans_a == 'T') {
    // A^T in CSR on the device (K6), then the no-transpose kernel on n rows.
    BOF_REQUIRE(ctx, nnz < (1ll << 31), "csrmm('T'): nnz must be below 2^31");
    int64_t *offs_d, *offs_t, *idx64;
    int32_t *idx32, *idx_t;
    float *vals_d, *vals_t;
    void* ws;
    const size_t wsb = csr2csc_workspace_bytes(m, n, nnz);
    BOF_TRY(slot_reserve(ctx, S_OFFS, (size_t)m + 1, &offs_d));
    BOF_TRY(slot_reserve(ctx, S_IDX64, (size_t)std::max<int64_t>(nnz, 1), &idx64));
    BOF_TRY(slot_reserve(ctx, S_IDX32, (size_t)std::max<int64_t>(nnz, 1), &idx32));
    BOF_TRY(slot_reserve(ctx, S_VALS, (size_t)std::max<int64_t>(nnz, 1), &vals_d));
    BOF_TRY(slot_reserve(ctx, S_OUT0, (size_t)n + 1, &offs_t));
    BOF_TRY(slot_reserve(ctx, S_OUT1, (size_t)std::max<int64_t>(nnz, 1), &idx_t));
    BOF_TRY(slot_reserve(ctx, S_OUT2, (size_t)std::max<int64_t>(nnz, 1), &vals_t));
    BOF_TRY(slot_reserve(ctx, S_WS, wsb, &ws));
    BOF_TRY(copy1d(ctx, offs_d, ia, (size_t)(m + 1) * 8, H2D, ctx->h2d));
    BOF_TRY(copy1d(ctx, idx64, ja, (size_t)nnz * 8, H2D, ctx->h2d));
    BOF_TRY(copy1d(ctx, vals_d, a, (size_t)nnz * 4, H2D, ctx->h2d));
    cudaEvent_t ev = get_event(ctx, 1);
    BOF_CUDA(ctx, cudaEventRecord(ev, ctx->h2d));
    BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, ev, 0));
    BOF_TRY(launch_idx_narrow(ctx, ctx->compute, idx64, idx32, nnz));
    BOF_TRY(launch_csr2csc(ctx, ctx->compute, m, n, nnz, offs_d, idx32, vals_d, offs_t, idx_t, vals_t, ws, wsb));
    offs_dev_all = offs_t;
    idx_dev_all = idx_t;
    vals_dev_all = vals_t;
  }

  if (trans_a == 'T') {
    // whole C on the device (n x k); beta needs the old C
    float* Cd = nullptr;
    BOF_TRY(slot_reserve(ctx, S_CBLK, (size_t)out_rows * k, &Cd));
    float* Cio = Cd;  // what is copied from/to the host
    float* Ccm = nullptr;
    if (colmaj) { BOF_TRY(slot_reserve(ctx, S_CBLK_T, (size_t)out_rows * k, &Ccm)); Cio = Ccm; }
    if (beta != 0.f) {
      BOF_TRY(copy1d(ctx, Cio, c, (size_t)out_rows * k * 4, H2D, ctx->h2d));
      cudaEvent_t ev = get_event(ctx, 2);
      BOF_CUDA(ctx, cudaEventRecord(ev, ctx->h2d));
      BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, ev, 0));
    }
    if (colmaj) {
      BOF_TRY(launch_spmm_rm(ctx, ctx->compute, out_rows, k, 1.f, vals_dev_all, idx_dev_all, offs_dev_all, Bd, k, 0.f, Cd, k));
      BOF_TRY(launch_transpose_axpby(ctx, ctx->compute, out_rows, k, alpha, Cd, k, beta, Ccm, out_rows));
    } else {
      BOF_TRY(launch_spmm_rm(ctx, ctx->compute, out_rows, k, alpha, vals_dev_all, idx_dev_all, offs_dev_all, Bd, k, beta, Cd, k));
    }
    BOF_TRY(copy1d(ctx, c, Cio, (size_t)out_rows * k * 4, D2H, ctx->compute));
    BOF_TRY(sync_all(ctx));
    stats_end(ctx);
    return call_guard.done();
  }

  // ---- 'N': streamed row blocks ----
  int64_t budget = (int64_t)ctx->cfg.csrmm_max_nnz;
  budget = std::min(budget, std::max<int64_t>(nnz / 8, 1 << 20));  // >= 8 blocks when the matrix is big enough
  const std::vector<int64_t> cuts = partition_rows(offs_host, m, budget);
  const int nblk = (int)cuts.size() - 1;
  int64_t max_rows = 1, max_nnz = 1;
  for (int i = 0; i < nblk; ++i) {
    max_rows = std::max(max_rows, cuts[i + 1] - cuts[i]);
    max_nnz = std::max(max_nnz, offs_host[cuts[i + 1]] - offs_host[cuts[i]]);
  }
  int64_t* offs_d[2]; int64_t* idx64_d[2]; int32_t* idx32_d[2]; float* vals_d[2]; float* cblk[2]; float* cblk_t[2] = {nullptr, nullptr};
  for (int g = 0; g < 2; ++g) {
    BOF_TRY(slot_reserve(ctx, S_OFFS + g, (size_t)max_rows + 1, &offs_d[g]));
    BOF_TRY(slot_reserve(ctx, S_IDX64 + g, (size_t)max_nnz, &idx64_d[g]));
    BOF_TRY(slot_reserve(ctx, S_IDX32 + g, (size_t)max_nnz, &idx32_d[g]));
    BOF_TRY(slot_reserve(ctx, S_VALS + g, (size_t)max_nnz, &vals_d[g]));
    BOF_TRY(slot_reserve(ctx, S_CBLK + g, (size_t)max_rows * k, &cblk[g]));
    if (colmaj) BOF_TRY(slot_reserve(ctx, S_CBLK_T + g, (size_t)max_rows * k, &cblk_t[g]));
  }
  // events: 4+g uploaded, 6+g computed, 8+g downloaded.  Software-pipelined issue order: block i+1 is
  // uploaded and launched before block i is downloaded, so a (host-blocking) staged download of
  // block i overlaps the kernel of block i+1 and the copy engines never wait on the host.
  bool used[2] = {false, false};
  uint64_t down_ticket[2] = {0, 0};  // pageable C: the drainer enqueues the download and records ev_down
  auto stage_block = [&](int i) -> int {
    const int g = i & 1;
    const int64_t r0 = cuts[i], r1 = cuts[i + 1], rows = r1 - r0;
    const int64_t z0 = offs_host[r0] - offs_host[0], z1 = offs_host[r1] - offs_host[0], bnnz = z1 - z0;
    cudaEvent_t ev_up = get_event(ctx, 4 + g), ev_done = get_event(ctx, 6 + g), ev_down = get_event(ctx, 8 + g);
    if (used[g]) {
      d2h_fence(ctx, down_ticket[g]);  // ev_down of block i-2 has been recorded (and ev_done may be re-recorded)
      BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->h2d, ev_done, 0));   // inputs of block i-2 consumed
      // Only an upload of old C rows (beta != 0) touches the C buffer from this stream; waiting for the download
      // unconditionally idled the H2D engine ~9 ms every other block (BOF_TRACE timeline, cfg-3: 419 -> 37x ms).
      if (beta != 0.f) BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->h2d, ev_down, 0));   // its C rows left the device
    }
    BOF_TRY(copy1d(ctx, offs_d[g], offs_host + r0, (size_t)(rows + 1) * 8, H2D, ctx->h2d));
    BOF_TRY(copy1d(ctx, idx64_d[g], ja + z0, (size_t)bnnz * 8, H2D, ctx->h2d));
    BOF_TRY(copy1d(ctx, vals_d[g], a + z0, (size_t)bnnz * 4, H2D, ctx->h2d));
    float* c_io = colmaj ? cblk_t[g] : cblk[g];
    if (beta != 0.f) {
      if (colmaj) BOF_TRY(copy2d(ctx, c_io, (size_t)rows * 4, c + r0, (size_t)m * 4, (size_t)rows * 4, (size_t)k, H2D, ctx->h2d));
      else BOF_TRY(copy1d(ctx, c_io, c + r0 * k, (size_t)rows * k * 4, H2D, ctx->h2d));
    }
    BOF_CUDA(ctx, cudaEventRecord(ev_up, ctx->h2d));
    trace_mark(ctx, ctx->h2d, "h2d: A block landed", i);
    BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, ev_up, 0));
    if (used[g]) BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, ev_down, 0));
    trace_mark(ctx, ctx->compute, "compute: block start", i);
    BOF_TRY(launch_idx_narrow(ctx, ctx->compute, idx64_d[g], idx32_d[g], bnnz));
    if (colmaj) {
      BOF_TRY(launch_spmm_rm(ctx, ctx->compute, rows, k, 1.f, vals_d[g], idx32_d[g], offs_d[g], Bd, k, 0.f, cblk[g], k));
      BOF_TRY(launch_transpose_axpby(ctx, ctx->compute, rows, k, alpha, cblk[g], k, beta, cblk_t[g], rows));
    } else {
      BOF_TRY(launch_spmm_rm(ctx, ctx->compute, rows, k, alpha, vals_d[g], idx32_d[g], offs_d[g], Bd, k, beta, cblk[g], k));
    }
    BOF_CUDA(ctx, cudaEventRecord(ev_done, ctx->compute));
    trace_mark(ctx, ctx->compute, "compute: block end", i);
    used[g] = true;
    return BOF_OK;
  };
  auto fetch_block = [&](int i) -> int {
    const int g = i & 1;
    const int64_t r0 = cuts[i], rows = cuts[i + 1] - r0;
    float* c_io = colmaj ? cblk_t[g] : cblk[g];
    cudaEvent_t ev_done = get_event(ctx, 6 + g), ev_down = get_event(ctx, 8 + g);
    if (colmaj) BOF_TRY(d2h_transfer(ctx, c + r0, (size_t)m * 4, c_io, (size_t)rows * 4, (size_t)rows * 4, (size_t)k, ctx->d2h, ev_done, ev_down, &down_ticket[g]));
    else BOF_TRY(d2h_transfer(ctx, c + r0 * k, (size_t)rows * k * 4, c_io, (size_t)rows * k * 4, (size_t)rows * k * 4, 1, ctx->d2h, ev_done, ev_down, &down_ticket[g]));
    trace_mark(ctx, ctx->d2h, "d2h: C block downloaded", i);
    return BOF_OK;
  };
  if (nblk > 0) BOF_TRY(stage_block(0));
  for (int i = 0; i < nblk; ++i) {
    if (i + 1 < nblk) BOF_TRY(stage_block(i + 1));
    BOF_TRY(fetch_block(i));
  }
  BOF_TRY(sync_all(ctx));
  trace_dump(ctx, "bof_host_csrmm");
  stats_end(ctx);
  return call_guard.done();
}

int bof_host_csrmm(bof_ctx* ctx, char trans_a, int64_t m, int64_t n, int64_t k, float alpha, float beta,
                   const float* a, const int64_t* ia, const int64_t* ja, char ord_b, const float* b, float* c) {
  return host_csrmm_impl(ctx, trans_a, m, n, k, alpha, beta, a, ia, ja, ord_b, b, c, nullptr);
}

int bof_host_csrmm_devb(bof_ctx* ctx, int64_t m, int64_t n, int64_t k, float alpha, float beta, const float* a,
                        const int64_t* ia, const int64_t* ja, const float* b_dev, float* c) {
  if (ctx && b_dev == nullptr) return fail(ctx, BOF_EINVAL, "csrmm_devb: b_dev is null");
  return host_csrmm_impl(ctx, 'N', m, n, k, alpha, beta, a, ia, ja, 'R', nullptr, c, b_dev);
}

// flash::gemm.  The canonical Q operand (op(B)^T for row-major problems) is uploaded and split
// into TF32 planes once; the canonical P operand and the output stream in row blocks,
// double-buffered (upload / split + MMA / download overlap).  The reference's k-dimension
// accumulate chain (src/blas/gemm.cpp:114-126) is an I/O artefact: the whole k extent is reduced
// on the device, so each C block crosses PCIe once.
static int host_gemm_impl(bof_ctx* ctx, char ord, char ta, char tb, int64_t m, int64_t n, int64_t k, float alpha,
                          float beta, const float* a, const float* b, float* c, int64_t lda, int64_t ldb, int64_t ldc,
                          bool q_on_device, const float* term_m = nullptr, const float* term_n = nullptr) {
  if (!ctx) return BOF_EINVAL;
  Canon cn;
  BOF_TRY(canon_gemm(ctx, ord, ta, tb, m, n, k, a, lda, b, ldb, ldc, &cn));
  BOF_REQUIRE(ctx, !q_on_device || ord == 'R', "gemm: a device-resident B is supported for mat_ord='R' only");
  BOF_CUDA(ctx, cudaSetDevice(ctx->device));
  stats_begin(ctx);
  CallGuard call_guard(ctx);
  if (cn.Mo == 0 || cn.No == 0) { stats_end(ctx); return call_guard.done(); }
  const cudaMemcpyKind H2D = cudaMemcpyHostToDevice, D2H = cudaMemcpyDeviceToHost;
  const int path = pick_gemm_path(ctx, cn.Mo, cn.No, cn.K);
  const int64_t K = cn.K, kp = padded_k(K);
  const bool tensor = (K > 0 && path != 3);
  const int cg = path == 1 ? 1 : 2;

  // upload rows [r0, r1) of a canonical operand as a tight device matrix; returns its strides
  auto upload_rows = [&](const float* src, int64_t s_r, int64_t s_k, int64_t r0, int64_t r1, float* dst,
                         int64_t* d_sr, int64_t* d_sk, cudaStream_t s) -> int {
    const int64_t rows = r1 - r0;
    if (K == 0) { *d_sr = 1; *d_sk = 1; return BOF_OK; }
    if (s_k == 1) {  // rows contiguous in k
      *d_sr = K; *d_sk = 1;
      return copy2d(ctx, dst, (size_t)K * 4, src + r0 * s_r, (size_t)s_r * 4, (size_t)K * 4, (size_t)rows, H2D, s);
    }
    *d_sr = 1; *d_sk = rows;  // stored k-major: a column range of a [K x ld] matrix
    return copy2d(ctx, dst, (size_t)rows * 4, src + r0, (size_t)s_k * 4, (size_t)rows * 4, (size_t)K, H2D, s);
  };

  // ---- buffers ----
  float* qraw = nullptr;
  float* q_hi = nullptr;
  float* q_lo = nullptr;
  if (q_on_device) qraw = const_cast<float*>(cn.qsrc);  // B already in HBM in its source layout; read-only here
  else BOF_TRY(slot_reserve(ctx, S_DENSE, (size_t)cn.No * std::max<int64_t>(K, 1), &qraw));
  const size_t qb = plane_bytes(cn.No, kp);
  if (tensor) {
    void* p;
    BOF_TRY(slot_reserve(ctx, S_DENSE_T, 2 * qb, &p));
    q_hi = static_cast<float*>(p);
    q_lo = reinterpret_cast<float*>(static_cast<uint8_t*>(p) + qb);
  }
  int64_t rb = (int64_t)ctx->cfg.gemm_row_block;
  rb = std::max<int64_t>(256, round_up<int64_t>(rb, 256));
  rb = std::min(rb, round_up<int64_t>(cn.Mo, 256));
  const int nblk = (int)ceil_div<int64_t>(cn.Mo, rb);
  const size_t pb = plane_bytes(rb, kp);
  constexpr int NB = kGemmRing;
  const int ngen = std::min(NB, nblk);
  // Ring of block generations.  Planes and C blocks of the ring are each ONE allocation (generation g at
  // row g*rb), so that consecutive generations can be multiplied in a single launch during the prologue.
  float* praw[NB];
  float* hi_all = nullptr; float* lo_all = nullptr; float* c_all = nullptr;
  for (int g = 0; g < ngen; ++g)
    BOF_TRY(slot_reserve(ctx, S_PRAW + g, (size_t)rb * std::max<int64_t>(K, 1), &praw[g]));
  if (tensor) {
    void* p;
    BOF_TRY(slot_reserve(ctx, S_PPLANES, 2 * (size_t)ngen * pb, &p));
    hi_all = static_cast<float*>(p);
    lo_all = reinterpret_cast<float*>(static_cast<uint8_t*>(p) + (size_t)ngen * pb);
  }
  BOF_TRY(slot_reserve(ctx, S_GCBLK, (size_t)ngen * rb * cn.No, &c_all));
  auto p_hi_of = [&](int g) { return hi_all + (size_t)g * rb * kp; };
  auto p_lo_of = [&](int g) { return lo_all + (size_t)g * rb * kp; };
  auto cblk_of = [&](int g) { return c_all + (size_t)g * rb * cn.No; };

  // flash::kmeans: C(i, j) += term_m[i], then += term_n[j] (i over m, j over n), applied to each block on the
  // device before it is downloaded.  In canonical (row-major output) form the rows are m for 'R', n for 'C'.
  float* term_rows_d = nullptr;
  float* term_cols_d = nullptr;
  const bool with_terms = term_m != nullptr && term_n != nullptr;
  const bool canon_rows_are_m = ord == 'R';
  if (with_terms) {
    float* t;
    BOF_TRY(slot_reserve(ctx, S_MISC, (size_t)(cn.Mo + cn.No), &t));
    term_rows_d = t;
    term_cols_d = t + cn.Mo;
    BOF_TRY(copy1d(ctx, term_rows_d, canon_rows_are_m ? term_m : term_n, (size_t)cn.Mo * 4, H2D, ctx->h2d));
    BOF_TRY(copy1d(ctx, term_cols_d, canon_rows_are_m ? term_n : term_m, (size_t)cn.No * 4, H2D, ctx->h2d));
    BOF_CUDA(ctx, cudaEventRecord(get_event(ctx, 3), ctx->h2d));
    BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, get_event(ctx, 3), 0));
  }

  // events: 8+g P block uploaded, 16+g P block split, 24+g block computed, 32+g block downloaded, 40+j Q panel
  // (reused for the column slabs of the last block)
  constexpr int EV_UP = 8, EV_SPLIT = 16, EV_DONE = 24, EV_DOWN = 32, EV_QPAN = 40;
  static_assert(kGemmRing <= 8, "event ids are spaced for at most 8 generations");
  bool used[NB] = {};
  uint64_t down_ticket[NB] = {};  // pageable C: the drainer enqueues the download and records EV_DOWN
  int64_t q_sr = 1, q_sk = 1;  // strides of the raw Q copy on the device

  auto upload_block = [&](int i) -> int {  // P rows (+ old C rows when beta != 0) of block i
    const int g = i % NB;
    const int64_t r0 = (int64_t)i * rb, r1 = std::min(cn.Mo, r0 + rb), rows = r1 - r0;
    if (used[g]) {
      d2h_fence(ctx, down_ticket[g]);  // EV_DOWN of block i-NB has been recorded; its EV_DONE may be re-recorded
      BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->h2d, get_event(ctx, (tensor ? EV_SPLIT : EV_DONE) + g), 0));  // raw P of block i-NB consumed
      if (beta != 0.f) BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->h2d, get_event(ctx, EV_DOWN + g), 0));       // C buffer free
    }
    int64_t p_sr, p_sk;
    trace_host(ctx, "caller: upload P block begin", i);
    BOF_TRY(upload_rows(cn.psrc, cn.p_sr, cn.p_sk, r0, r1, praw[g], &p_sr, &p_sk, ctx->h2d));
    trace_host(ctx, "caller: upload P block end", i);
    if (beta != 0.f)
      BOF_TRY(copy2d(ctx, cblk_of(g), (size_t)cn.No * 4, c + r0 * cn.ldc, (size_t)cn.ldc * 4, (size_t)cn.No * 4, (size_t)rows, H2D, ctx->h2d));
    BOF_CUDA(ctx, cudaEventRecord(get_event(ctx, EV_UP + g), ctx->h2d));
    trace_mark(ctx, ctx->h2d, "h2d: P block landed", i);
    return BOF_OK;
  };
  // compute of block i against Q rows [n0, n1) (the whole Q when not panelled); `first`/`last` bracket the block
  auto rows_of = [&](int i) { return std::min(cn.Mo, (int64_t)(i + 1) * rb) - (int64_t)i * rb; };
  // wait for block i's upload (and for its buffers), split its rows into planes
  auto prepare_block = [&](int i) -> int {
    const int g = i % NB;
    const int64_t rows = rows_of(i);
    const int64_t p_sr = cn.p_sk == 1 ? K : 1, p_sk = cn.p_sk == 1 ? 1 : rows;
    BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, get_event(ctx, EV_UP + g), 0));
    if (used[g]) {
      d2h_fence(ctx, down_ticket[g]);
      BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, get_event(ctx, EV_DOWN + g), 0));
    }
    if (tensor) {
      BOF_TRY(launch_split_planes(ctx, ctx->compute, rows, K, praw[g], p_sr, p_sk, p_hi_of(g), p_lo_of(g), kp));
      BOF_CUDA(ctx, cudaEventRecord(get_event(ctx, EV_SPLIT + g), ctx->compute));
    }
    return BOF_OK;
  };
  // blocks [i0, i0 + cnt) (consecutive generations, no ring wrap) against Q rows [n0, n1), one launch
  auto gemm_blocks = [&](int i0, int cnt, int64_t n0, int64_t n1) -> int {
    const int g = i0 % NB;
    int64_t rows = 0;
    for (int i = i0; i < i0 + cnt; ++i) rows += rows_of(i);
    if (tensor) {
      GemmEpilogue ep;
      ep.alpha = alpha; ep.beta = beta; ep.C = cblk_of(g) + n0; ep.ldc = cn.No;
      trace_mark(ctx, ctx->compute, "compute: gemm start, first block", i0);
      const int rc = launch_gemm_tc(ctx, ctx->compute, cg, rows, n1 - n0, K, kp, p_hi_of(g), p_lo_of(g), q_hi + n0 * kp,
                                    q_lo + n0 * kp, ep, k_chunk_of(ctx));
      trace_mark(ctx, ctx->compute, "compute: gemm end, rows", (int)rows);
      return rc;
    }
    const int64_t p_sr = cn.p_sk == 1 ? K : 1, p_sk = cn.p_sk == 1 ? 1 : rows;  // cnt == 1 on this path
    return launch_gemm_ffma(ctx, ctx->compute, rows, n1 - n0, K, alpha, praw[g], p_sr, p_sk, qraw + n0 * q_sr, q_sk, q_sr,
                            beta, cblk_of(g) + n0, cn.No);
  };
  auto finish_block = [&](int i) -> int {
    const int g = i % NB;
    if (with_terms)  // term_m is added first (kmeans_task.h:74-80): it is the row term iff the canonical rows are m
      BOF_TRY(launch_add_outer_terms(ctx, ctx->compute, cblk_of(g), rows_of(i), cn.No, cn.No, term_rows_d + (int64_t)i * rb,
                                     term_cols_d, canon_rows_are_m ? 1 : 0));
    BOF_CUDA(ctx, cudaEventRecord(get_event(ctx, EV_DONE + g), ctx->compute));
    used[g] = true;
    return BOF_OK;
  };
  auto fetch_block = [&](int i) -> int {
    const int g = i % NB;
    const int64_t r0 = (int64_t)i * rb, rows = std::min(cn.Mo, r0 + rb) - r0;
    BOF_TRY(d2h_transfer(ctx, c + r0 * cn.ldc, (size_t)cn.ldc * 4, cblk_of(g), (size_t)cn.No * 4, (size_t)cn.No * 4, (size_t)rows,
                         ctx->d2h, get_event(ctx, EV_DONE + g), get_event(ctx, EV_DOWN + g), &down_ticket[g]));
    if (down_ticket[g] == 0) trace_mark(ctx, ctx->d2h, "d2h: C block downloaded", i);
    return BOF_OK;
  };

  // ---- Q (resident) and block 0 ----
  // When Q is K-major in the source its rows upload as contiguous panels: panel 0, then P block 0, then the
  // remaining panels; block 0 is computed panel by panel as they land, so only one panel and one P block of
  // PCIe time are exposed before the tensor cores start.
  const bool q_panels = tensor && !q_on_device && (size_t)cn.No * K * 4 >= (256u << 20);
  static const int64_t n_q_panels = getenv("BOF_GEMM_QPANELS") ? std::min(8, std::max(1, atoi(getenv("BOF_GEMM_QPANELS")))) : 8;
  const int64_t qpan_rows = q_panels ? std::max<int64_t>(256, round_up<int64_t>(ceil_div<int64_t>(cn.No, n_q_panels), 256)) : cn.No;
  const int n_qpan = (int)ceil_div<int64_t>(cn.No, qpan_rows);
  // Every panel is kept as its own tight block at qraw + n0*K: [rows x K] when Q is K-major in the source,
  // [K x rows] (a column range of the stored matrix, pitched copy) when it is not.
  std::vector<int64_t> pan_sr((size_t)n_qpan, 1), pan_sk((size_t)n_qpan, 1);
  auto upload_q_panel = [&](int j) -> int {
    const int64_t n0 = (int64_t)j * qpan_rows, n1 = std::min(cn.No, n0 + qpan_rows);
    if (q_on_device) { pan_sr[j] = cn.q_sr; pan_sk[j] = cn.q_sk; }  // single panel, source strides
    else BOF_TRY(upload_rows(cn.qsrc, cn.q_sr, cn.q_sk, n0, n1, qraw + n0 * K, &pan_sr[j], &pan_sk[j], ctx->h2d));
    if (n_qpan == 1) { q_sr = pan_sr[0]; q_sk = pan_sk[0]; }
    BOF_CUDA(ctx, cudaEventRecord(get_event(ctx, EV_QPAN + j), ctx->h2d));
    trace_mark(ctx, ctx->h2d, "h2d: Q panel landed", j);
    return BOF_OK;
  };
  auto split_q_panel = [&](int j) -> int {
    const int64_t n0 = (int64_t)j * qpan_rows, n1 = std::min(cn.No, n0 + qpan_rows);
    BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, get_event(ctx, EV_QPAN + j), 0));
    if (!tensor) return BOF_OK;
    return launch_split_planes(ctx, ctx->compute, n1 - n0, K, qraw + n0 * K, pan_sr[j], pan_sk[j], q_hi + n0 * kp,
                               q_lo + n0 * kp, kp);
  };
  // Prologue: the first blocks ride between the Q panels (h2d order Q0 P0 Q1 P1 Q2 P2 Q3 P3 Q4 .. Qn) and are
  // computed against each panel as it lands (compute order = arrival order), so the tensor cores start after
  // one panel + one block of PCIe time and stay fed while the rest of Q uploads: every new panel unlocks one
  // tile per prologue block.
  trace_mark(ctx, ctx->h2d, "start", 0);
  // One generation stays out of the prologue: the prologue blocks all finish together (with the last panel), so
  // the first steady-state block would otherwise wait for a whole C block to be downloaded (11 ms at 32768^3,
  // seen with BOF_TRACE=1) before it could reuse generation 0.
  const int npro = n_qpan > 1 ? std::min({nblk, NB - 1, n_qpan}) : 1;  // blocks handled by the prologue
  auto pan = [&](int j, int64_t* n0, int64_t* n1) { *n0 = (int64_t)j * qpan_rows; *n1 = std::min(cn.No, *n0 + qpan_rows); };
  const bool merge = tensor;  // the CUDA-core path multiplies one block per launch
  // Uploads and launches are issued panel by panel: with pinned memory the order of issue is immaterial (everything
  // is asynchronous), but a pageable upload blocks this thread while it is staged, and launching only after the
  // whole prologue had been uploaded left the GPU idle for the first 170 ms at 32768^3 (BOF_TRACE).
  for (int t = 0; t < n_qpan; ++t) {
    int64_t n0, n1;
    pan(t, &n0, &n1);
    BOF_TRY(upload_q_panel(t));
    if (t < npro) BOF_TRY(upload_block(t));
    BOF_TRY(split_q_panel(t));
    // panel t against the blocks that landed before it: one launch over those consecutive generations
    const int older = std::min(t, npro);
    if (older > 0) {
      if (merge) BOF_TRY(gemm_blocks(0, older, n0, n1));
      else for (int i = 0; i < older; ++i) BOF_TRY(gemm_blocks(i, 1, n0, n1));
    }
    // block t landed right after panel t: all panels so far at once
    if (t < npro) {
      BOF_TRY(prepare_block(t));
      BOF_TRY(gemm_blocks(t, 1, 0, n1));
    }
    if (t == n_qpan - 1)
      for (int i = 0; i < npro; ++i) BOF_TRY(finish_block(i));
  }
  // ---- steady state: Q complete; keep NB blocks in flight, fetch the oldest before reusing its buffers ----
  int next_fetch = 0;
  int fetch_end = nblk;  // blocks [next_fetch, fetch_end) still have to be downloaded whole
  for (int i = npro; i < nblk; ++i) {
    if (next_fetch < i) BOF_TRY(fetch_block(next_fetch++));            // keeps the downloads flowing
    while (next_fetch <= i - NB) BOF_TRY(fetch_block(next_fetch++));   // block i-NB owned these buffers
    BOF_TRY(upload_block(i));
    BOF_TRY(prepare_block(i));
    if (i == nblk - 1 && tensor && cn.No >= 2048) {
      // Drain: the last block is multiplied and downloaded in four column slabs, so only the last slab's
      // download (a quarter of a block) is exposed after the tensor cores stop.
      const int g = i % NB;
      const int64_t r0 = (int64_t)i * rb, rows = rows_of(i);
      const int64_t slab = round_up<int64_t>(ceil_div<int64_t>(cn.No, 4), 256);
      while (next_fetch < i) BOF_TRY(fetch_block(next_fetch++));  // d2h is FIFO: earlier blocks first
      int j = 0;
      for (int64_t n0 = 0; n0 < cn.No; n0 += slab, ++j) {
        const int64_t n1 = std::min(cn.No, n0 + slab);
        BOF_TRY(gemm_blocks(i, 1, n0, n1));
        if (with_terms)
          BOF_TRY(launch_add_outer_terms(ctx, ctx->compute, cblk_of(g) + n0, rows, n1 - n0, cn.No, term_rows_d + r0,
                                         term_cols_d + n0, canon_rows_are_m ? 1 : 0));
        BOF_CUDA(ctx, cudaEventRecord(get_event(ctx, EV_QPAN + j), ctx->compute));
        BOF_TRY(d2h_transfer(ctx, c + r0 * cn.ldc + n0, (size_t)cn.ldc * 4, cblk_of(g) + n0, (size_t)cn.No * 4, (size_t)(n1 - n0) * 4,
                             (size_t)rows, ctx->d2h, get_event(ctx, EV_QPAN + j), n1 == cn.No ? get_event(ctx, EV_DOWN + g) : nullptr,
                             &down_ticket[g]));
        if (down_ticket[g] == 0) trace_mark(ctx, ctx->d2h, "d2h: last block, slab downloaded", j);
      }
      BOF_CUDA(ctx, cudaEventRecord(get_event(ctx, EV_DONE + g), ctx->compute));
      used[g] = true;
      fetch_end = i;
      break;
    }
    BOF_TRY(gemm_blocks(i, 1, 0, cn.No));
    BOF_TRY(finish_block(i));
  }
  while (next_fetch < fetch_end) BOF_TRY(fetch_block(next_fetch++));
  BOF_TRY(sync_all(ctx));
  trace_dump(ctx, "bof_host_gemm");
  stats_end(ctx);
  return call_guard.done();
}

int bof_host_gemm(bof_ctx* ctx, char ord, char ta, char tb, int64_t m, int64_t n, int64_t k, float alpha,
                  float beta, const float* a, const float* b, float* c, int64_t lda, int64_t ldb, int64_t ldc) {
  return host_gemm_impl(ctx, ord, ta, tb, m, n, k, alpha, beta, a, b, c, lda, ldb, ldc, false);
}

int bof_host_kmeans_dist(bof_ctx* ctx, char ord, char ta, char tb, int64_t m, int64_t n, int64_t k, float alpha,
                         float beta, const float* a, const float* b, float* c, int64_t lda, int64_t ldb, int64_t ldc,
                         const float* c_l2sq, const float* p_l2sq) {
  if (ctx && (c_l2sq == nullptr || p_l2sq == nullptr)) return fail(ctx, BOF_EINVAL, "kmeans: c_l2sq / p_l2sq is null");
  return host_gemm_impl(ctx, ord, ta, tb, m, n, k, alpha, beta, a, b, c, lda, ldb, ldc, false, c_l2sq, p_l2sq);
}

int bof_host_gemm_devb(bof_ctx* ctx, char ta, char tb, int64_t m, int64_t n, int64_t k, float alpha, float beta,
                       const float* a, const float* b_dev, float* c, int64_t lda, int64_t ldb, int64_t ldc) {
  return host_gemm_impl(ctx, 'R', ta, tb, m, n, k, alpha, beta, a, b_dev, c, lda, ldb, ldc, true);
}

// flash::csrgemv: x resident, A streams in row blocks; 'N' writes disjoint y rows, 'T' accumulates
// every block into the full y on the device (zeroed once, as src/blas/csrgemv.cpp:64 does).
int bof_host_csrgemv(bof_ctx* ctx, char trans_a, int64_t m, int64_t n, const float* a, const int64_t* ia,
                     const int64_t* ja, const float* b, float* c) {
  if (!ctx) return BOF_EINVAL;
  BOF_REQUIRE(ctx, is_nt(trans_a), "csrgemv trans_a error : expected=N or T, found=%c", trans_a);
  BOF_REQUIRE(ctx, m >= 0 && n >= 0 && m < (1ll << 31) && n < (1ll << 31), "csrgemv: bad dimension");
  BOF_CUDA(ctx, cudaSetDevice(ctx->device));
  stats_begin(ctx);
  CallGuard call_guard(ctx);
  const cudaMemcpyKind H2D = cudaMemcpyHostToDevice, D2H = cudaMemcpyDeviceToHost;
  const bool tr = trans_a == 'T';
  const int64_t xlen = tr ? m : n, ylen = tr ? n : m;
  if (ylen == 0) { stats_end(ctx); return call_guard.done(); }
  float *xd, *yd;
  BOF_TRY(slot_reserve(ctx, S_DENSE, (size_t)std::max<int64_t>(xlen, 1), &xd));
  BOF_TRY(slot_reserve(ctx, S_CBLK, (size_t)ylen, &yd));
  BOF_TRY(copy1d(ctx, xd, b, (size_t)xlen * 4, H2D, ctx->h2d));
  cudaEvent_t evx = get_event(ctx, 0);
  BOF_CUDA(ctx, cudaEventRecord(evx, ctx->h2d));
  BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, evx, 0));
  BOF_CUDA(ctx, cudaMemsetAsync(yd, 0, (size_t)ylen * 4, ctx->compute));

  const int64_t nnz = ia[m] - ia[0];
  int64_t budget = std::min<int64_t>((int64_t)ctx->cfg.csrmm_max_nnz, std::max<int64_t>(nnz / 8, 1 << 20));
  const std::vector<int64_t> cuts = partition_rows(ia, m, budget);
  const int nblk = (int)cuts.size() - 1;
  int64_t max_rows = 1, max_nnz = 1;
  for (int i = 0; i < nblk; ++i) {
    max_rows = std::max(max_rows, cuts[i + 1] - cuts[i]);
    max_nnz = std::max(max_nnz, ia[cuts[i + 1]] - ia[cuts[i]]);
  }
  int64_t* offs_d[2]; int64_t* idx64_d[2]; int32_t* idx32_d[2]; float* vals_d[2];
  for (int g = 0; g < 2; ++g) {
    BOF_TRY(slot_reserve(ctx, S_OFFS + g, (size_t)max_rows + 1, &offs_d[g]));
    BOF_TRY(slot_reserve(ctx, S_IDX64 + g, (size_t)max_nnz, &idx64_d[g]));
    BOF_TRY(slot_reserve(ctx, S_IDX32 + g, (size_t)max_nnz, &idx32_d[g]));
    BOF_TRY(slot_reserve(ctx, S_VALS + g, (size_t)max_nnz, &vals_d[g]));
  }
  bool used[2] = {false, false};
  for (int i = 0; i < nblk; ++i) {
    const int g = i & 1;
    const int64_t r0 = cuts[i], r1 = cuts[i + 1], rows = r1 - r0;
    const int64_t z0 = ia[r0] - ia[0], bnnz = ia[r1] - ia[r0];
    cudaEvent_t ev_up = get_event(ctx, 4 + g), ev_done = get_event(ctx, 6 + g);
    if (used[g]) BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->h2d, ev_done, 0));
    BOF_TRY(copy1d(ctx, offs_d[g], ia + r0, (size_t)(rows + 1) * 8, H2D, ctx->h2d));
    BOF_TRY(copy1d(ctx, idx64_d[g], ja + z0, (size_t)bnnz * 8, H2D, ctx->h2d));
    BOF_TRY(copy1d(ctx, vals_d[g], a + z0, (size_t)bnnz * 4, H2D, ctx->h2d));
    BOF_CUDA(ctx, cudaEventRecord(ev_up, ctx->h2d));
    BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, ev_up, 0));
    BOF_TRY(launch_idx_narrow(ctx, ctx->compute, idx64_d[g], idx32_d[g], bnnz));
    if (!tr) {
      BOF_TRY(launch_spmv(ctx, ctx->compute, 'N', rows, n, vals_d[g], idx32_d[g], offs_d[g], xd, yd + r0));
    } else {
      // y += A_blk^T x_blk : launch the accumulate kernel directly (y was zeroed once above)
      BOF_TRY(launch_spmv(ctx, ctx->compute, 't', rows, n, vals_d[g], idx32_d[g], offs_d[g], xd + r0, yd));
    }
    BOF_CUDA(ctx, cudaEventRecord(ev_done, ctx->compute));
    used[g] = true;
  }
  BOF_TRY(copy1d(ctx, c, yd, (size_t)ylen * 4, D2H, ctx->compute));
  BOF_TRY(sync_all(ctx));
  stats_end(ctx);
  return call_guard.done();
}

// flash::csrcsc: the whole matrix is transposed in HBM in one shot (the reference's two-phase
// row-block transpose + column-block merge exists only because a block had to fit in DRAM).
int bof_host_csrcsc(bof_ctx* ctx, int64_t m, int64_t n, const int64_t* ia, const int64_t* ja, const float* a,
                    int64_t* ia_tr, int64_t* ja_tr, float* a_tr) {
  if (!ctx) return BOF_EINVAL;
  BOF_REQUIRE(ctx, m >= 0 && n >= 0 && m < (1ll << 31) && n < (1ll << 31), "csrcsc: bad dimension");
  BOF_CUDA(ctx, cudaSetDevice(ctx->device));
  stats_begin(ctx);
  CallGuard call_guard(ctx);
  const cudaMemcpyKind H2D = cudaMemcpyHostToDevice, D2H = cudaMemcpyDeviceToHost;
  const int64_t nnz = ia[m] - ia[0];
  BOF_REQUIRE(ctx, nnz >= 0 && nnz < (1ll << 31), "csrcsc: nnz must be in [0, 2^31)");
  const size_t z = (size_t)std::max<int64_t>(nnz, 1);
  int64_t *offs_d, *offs_t, *idx64;
  int32_t *idx32, *idx_t;
  float *vals_d, *vals_t;
  void* ws;
  const size_t wsb = csr2csc_workspace_bytes(m, n, nnz);
  BOF_TRY(slot_reserve(ctx, S_OFFS, (size_t)m + 1, &offs_d));
  BOF_TRY(slot_reserve(ctx, S_IDX64, z, &idx64));
  BOF_TRY(slot_reserve(ctx, S_IDX32, z, &idx32));
  BOF_TRY(slot_reserve(ctx, S_VALS, z, &vals_d));
  BOF_TRY(slot_reserve(ctx, S_OUT0, (size_t)n + 1, &offs_t));
  BOF_TRY(slot_reserve(ctx, S_OUT1, z, &idx_t));
  BOF_TRY(slot_reserve(ctx, S_OUT2, z, &vals_t));
  BOF_TRY(slot_reserve(ctx, S_WS, wsb, &ws));
  // values ride a second stream so that both copy engines' queues stay busy
  BOF_TRY(copy1d(ctx, offs_d, ia, (size_t)(m + 1) * 8, H2D, ctx->h2d));
  BOF_TRY(copy1d(ctx, idx64, ja, (size_t)nnz * 8, H2D, ctx->h2d));
  BOF_TRY(copy1d(ctx, vals_d, a, (size_t)nnz * 4, H2D, ctx->h2d));
  cudaEvent_t ev = get_event(ctx, 0);
  BOF_CUDA(ctx, cudaEventRecord(ev, ctx->h2d));
  BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, ev, 0));
  BOF_TRY(launch_idx_narrow(ctx, ctx->compute, idx64, idx32, nnz));
  BOF_TRY(launch_csr2csc(ctx, ctx->compute, m, n, nnz, offs_d, idx32, vals_d, offs_t, idx_t, vals_t, ws, wsb));
  // the int64 staging buffer of the input indices is free again: reuse it for the widened output
  BOF_TRY(launch_idx_widen(ctx, ctx->compute, idx_t, idx64, nnz));
  cudaEvent_t evk = get_event(ctx, 1);
  BOF_CUDA(ctx, cudaEventRecord(evk, ctx->compute));
  BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->d2h, evk, 0));
  BOF_TRY(copy1d(ctx, a_tr, vals_t, (size_t)nnz * 4, D2H, ctx->d2h));
  BOF_TRY(copy1d(ctx, ja_tr, idx64, (size_t)nnz * 8, D2H, ctx->compute));
  BOF_TRY(copy1d(ctx, ia_tr, offs_t, (size_t)(n + 1) * 8, D2H, ctx->compute));  // offsets last, as csrcsc.cpp:150
  BOF_TRY(sync_all(ctx));
  stats_end(ctx);
  return call_guard.done();
}

// ---- resident CSR: A stays in HBM across calls -----------------------------------
// The reference re-reads A from flash on every flash::csrmm / flash::csrgemv call (its Cache is flushed when a
// kernel returns, src/blas/csrmm.cpp:259); the Krylov / eigensolver loops that call them re-multiply the same A.
// A bof_csr uploads A once (indices narrowed to int32), optionally keeps A^T next to it, and each product then
// moves only the dense operands over PCIe.

struct bof_csr {
  bof_ctx* ctx = nullptr;
  int64_t m = 0, n = 0, nnz = 0;
  // [0] = A (m rows), [1] = A^T in CSR (n rows), built on first use
  float* vals[2] = {nullptr, nullptr};
  int32_t* idx[2] = {nullptr, nullptr};
  int64_t* offs[2] = {nullptr, nullptr};
  std::vector<int64_t> offs_host[2];
  bool have_t = false;
};

static void csr_free(bof_csr* h) {
  if (!h) return;
  for (int t = 0; t < 2; ++t) {
    if (h->vals[t]) cudaFree(h->vals[t]);
    if (h->idx[t]) cudaFree(h->idx[t]);
    if (h->offs[t]) cudaFree(h->offs[t]);
  }
  delete h;
}

static int csr_alloc(bof_ctx* ctx, bof_csr* h, int t, int64_t rows) {
  const size_t z = (size_t)std::max<int64_t>(h->nnz, 1);
  struct Req { void** p; size_t bytes; } reqs[] = {
      {(void**)&h->vals[t], z * 4}, {(void**)&h->idx[t], z * 4}, {(void**)&h->offs[t], (size_t)(rows + 1) * 8}};
  for (auto& r : reqs) {
    if (cudaMalloc(r.p, r.bytes) != cudaSuccess) {
      cudaGetLastError();
      return fail(ctx, BOF_ENOMEM, "csr: cudaMalloc of %zu bytes failed", r.bytes);
    }
  }
  return BOF_OK;
}

int bof_csr_open(bof_ctx* ctx, int64_t m, int64_t n, const float* a, const int64_t* ia, const int64_t* ja,
                 bof_csr** out) {
  if (!ctx || !out) return BOF_EINVAL;
  *out = nullptr;
  BOF_REQUIRE(ctx, m >= 0 && n >= 0 && m < (1ll << 31) && n < (1ll << 31), "csr_open: bad dimension");
  BOF_REQUIRE(ctx, ia != nullptr, "csr_open: ia is null");
  const int64_t nnz = ia[m] - ia[0];
  BOF_REQUIRE(ctx, nnz >= 0 && nnz < (1ll << 31), "csr_open: nnz must be in [0, 2^31)");
  BOF_CUDA(ctx, cudaSetDevice(ctx->device));
  stats_begin(ctx);
  CallGuard call_guard(ctx);
  bof_csr* h = new bof_csr();
  h->ctx = ctx; h->m = m; h->n = n; h->nnz = nnz;
  auto guard = [&](int rc) { if (rc != BOF_OK) { quiesce(ctx); csr_free(h); } return rc; };
  if (int rc = guard(csr_alloc(ctx, h, 0, m))) return rc;
  h->offs_host[0].assign(ia, ia + m + 1);
  const cudaMemcpyKind H2D = cudaMemcpyHostToDevice;
  if (int rc = guard(copy1d(ctx, h->offs[0], ia, (size_t)(m + 1) * 8, H2D, ctx->h2d))) return rc;
  if (int rc = guard(copy1d(ctx, h->vals[0], a, (size_t)nnz * 4, H2D, ctx->h2d))) return rc;
  // int64 column indices cross PCIe as they are on disk and are narrowed chunk by chunk (two staging generations)
  const int64_t chunk = std::max<int64_t>(std::min<int64_t>((int64_t)ctx->cfg.csrmm_max_nnz, nnz), 1);
  int64_t* st[2];
  for (int g = 0; g < 2; ++g)
    if (int rc = guard(slot_reserve(ctx, S_IDX64 + g, (size_t)chunk, &st[g]))) return rc;
  bool used[2] = {false, false};
  int ci = 0;
  for (int64_t z = 0; z < nnz; z += chunk, ++ci) {
    const int g = ci & 1;
    const int64_t cnt = std::min(chunk, nnz - z);
    cudaEvent_t ev_up = get_event(ctx, 4 + g), ev_done = get_event(ctx, 6 + g);
    if (used[g] && cudaStreamWaitEvent(ctx->h2d, ev_done, 0) != cudaSuccess) return guard(fail(ctx, BOF_ECUDA, "csr_open: wait failed"));
    if (int rc = guard(copy1d(ctx, st[g], ja + z, (size_t)cnt * 8, H2D, ctx->h2d))) return rc;
    cudaEventRecord(ev_up, ctx->h2d);
    cudaStreamWaitEvent(ctx->compute, ev_up, 0);
    if (int rc = guard(launch_idx_narrow(ctx, ctx->compute, st[g], h->idx[0] + z, cnt))) return rc;
    cudaEventRecord(ev_done, ctx->compute);
    used[g] = true;
  }
  if (int rc = guard(sync_all(ctx))) return rc;
  stats_end(ctx);
  *out = h;
  return call_guard.done();
}

int bof_csr_build_transpose(bof_csr* h) {
  if (!h) return BOF_EINVAL;
  if (h->have_t) return BOF_OK;
  bof_ctx* ctx = h->ctx;
  BOF_CUDA(ctx, cudaSetDevice(ctx->device));
  CallGuard call_guard(ctx);
  BOF_TRY(csr_alloc(ctx, h, 1, h->n));
  void* ws;
  const size_t wsb = csr2csc_workspace_bytes(h->m, h->n, h->nnz);
  BOF_TRY(slot_reserve(ctx, S_WS, wsb, &ws));
  BOF_TRY(launch_csr2csc(ctx, ctx->compute, h->m, h->n, h->nnz, h->offs[0], h->idx[0], h->vals[0], h->offs[1],
                         h->idx[1], h->vals[1], ws, wsb));
  h->offs_host[1].resize((size_t)h->n + 1);
  BOF_TRY(copy1d(ctx, h->offs_host[1].data(), h->offs[1], (size_t)(h->n + 1) * 8, cudaMemcpyDeviceToHost, ctx->compute));
  BOF_TRY(sync_all(ctx));
  h->have_t = true;
  return call_guard.done();
}

int bof_csr_arrays(bof_csr* h, char trans_a, const float** vals, const int32_t** idx, const int64_t** offs,
                   int64_t* nnz) {
  if (!h) return BOF_EINVAL;
  BOF_REQUIRE(h->ctx, is_nt(trans_a), "csr_arrays: unrecognized value for param trans_a = '%c'", trans_a);
  const int t = trans_a == 'T';
  if (t) BOF_TRY(bof_csr_build_transpose(h));
  if (vals) *vals = h->vals[t];
  if (idx) *idx = h->idx[t];
  if (offs) *offs = h->offs[t];
  if (nnz) *nnz = h->nnz;
  return BOF_OK;
}

// C = alpha op(A) B + beta C with host B, C.  The dense operands move in column panels of 64 (a panel is an
// independent product), so the upload of panel p+1, the kernels of panel p and the download of its row blocks
// overlap on the three streams; A is only read from HBM.
int bof_csr_mm(bof_csr* h, char trans_a, int64_t k, float alpha, float beta, char ord_b, const float* b, float* c) {
  if (!h) return BOF_EINVAL;
  bof_ctx* ctx = h->ctx;
  BOF_REQUIRE(ctx, is_nt(trans_a), "csrmm: unrecognized value for param trans_a = '%c'", trans_a);
  BOF_REQUIRE(ctx, is_rc(ord_b), "csrmm: unrecognized value for param ord_b = '%c'", ord_b);
  BOF_REQUIRE(ctx, k >= 0, "csrmm: bad dimension");
  BOF_CUDA(ctx, cudaSetDevice(ctx->device));
  const int t = trans_a == 'T';
  if (t) BOF_TRY(bof_csr_build_transpose(h));
  stats_begin(ctx);
  CallGuard call_guard(ctx);
  const int64_t out_rows = t ? h->n : h->m, in_rows = t ? h->m : h->n;
  if (out_rows == 0 || k == 0) { stats_end(ctx); return call_guard.done(); }
  const bool colmaj = ord_b == 'C';
  const cudaMemcpyKind H2D = cudaMemcpyHostToDevice, D2H = cudaMemcpyDeviceToHost;
  const float* vals = h->vals[t];
  const int32_t* idx = h->idx[t];
  const int64_t* offs = h->offs[t];
  const int64_t* oh = h->offs_host[t].data();

  // panel width: measured on the cfg-3 matrix at k = 256 (profiles/r01/resident_panel_sweep.txt): 32 -> 304 ms
  // (kernel-bound), 64 -> 217 ms, 128 -> 260 ms, 256 (no panels) -> 309 ms
  static const int64_t panel_cols = getenv("BOF_CSR_PANEL") ? std::max(4, atoi(getenv("BOF_CSR_PANEL"))) : 64;
  const int64_t kb = std::min<int64_t>(k, panel_cols);
  const int npan = (int)ceil_div<int64_t>(k, kb);
  const int64_t rows_blk = std::max<int64_t>(4096, (64ll << 20) / (kb * 4));
  const int nblk = (int)ceil_div<int64_t>(out_rows, rows_blk);
  const size_t pan_elems = (size_t)std::max<int64_t>(in_rows, 1) * kb;
  float *bpan_all, *braw_all = nullptr, *cblk[2], *cblk_t[2] = {nullptr, nullptr};
  BOF_TRY(slot_reserve(ctx, S_DENSE, 2 * pan_elems, &bpan_all));
  if (colmaj) BOF_TRY(slot_reserve(ctx, S_DENSE_T, 2 * pan_elems, &braw_all));
  for (int g = 0; g < 2; ++g) {
    BOF_TRY(slot_reserve(ctx, S_CBLK + g, (size_t)std::min(rows_blk, out_rows) * kb, &cblk[g]));
    if (colmaj) BOF_TRY(slot_reserve(ctx, S_CBLK_T + g, (size_t)std::min(rows_blk, out_rows) * kb, &cblk_t[g]));
  }
  // events: 0+gp panel uploaded, 2+gp panel consumed, 4+g old C block uploaded, 6+g block computed, 8+g block downloaded
  bool pan_used[2] = {false, false}, blk_used[2] = {false, false};
  uint64_t down_ticket[2] = {0, 0};  // pageable C: the drainer enqueues the download and records event 8+g
  auto upload_panel = [&](int p) -> int {
    const int gp = p & 1;
    const int64_t j0 = (int64_t)p * kb, kbp = std::min(kb, k - j0);
    float* bp = bpan_all + (size_t)gp * pan_elems;
    if (pan_used[gp]) BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->h2d, get_event(ctx, 2 + gp), 0));
    if (in_rows > 0) {
      if (colmaj) {
        float* raw = braw_all + (size_t)gp * pan_elems;  // columns j0.. of a column-major B are contiguous
        BOF_TRY(copy1d(ctx, raw, b + j0 * in_rows, (size_t)in_rows * kbp * 4, H2D, ctx->h2d));
      } else if (npan == 1) {
        BOF_TRY(copy1d(ctx, bp, b, (size_t)in_rows * k * 4, H2D, ctx->h2d));
      } else {
        BOF_TRY(copy2d(ctx, bp, (size_t)kbp * 4, b + j0, (size_t)k * 4, (size_t)kbp * 4, (size_t)in_rows, H2D, ctx->h2d));
      }
    }
    BOF_CUDA(ctx, cudaEventRecord(get_event(ctx, 0 + gp), ctx->h2d));
    pan_used[gp] = true;
    return BOF_OK;
  };
  int bc = 0;  // running block counter -> C buffer generation
  auto run_block = [&](int p, int i, int g) -> int {
    const int gp = p & 1;
    const int64_t j0 = (int64_t)p * kb, kbp = std::min(kb, k - j0);
    const int64_t r0 = (int64_t)i * rows_blk, rows = std::min(rows_blk, out_rows - r0);
    const int64_t z0 = oh[r0] - oh[0];
    float* bp = bpan_all + (size_t)gp * pan_elems;
    float* c_io = colmaj ? cblk_t[g] : cblk[g];
    if (blk_used[g]) d2h_fence(ctx, down_ticket[g]);  // event 8+g recorded, 6+g may be re-recorded
    if (beta != 0.f) {
      if (blk_used[g]) {
        BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->h2d, get_event(ctx, 6 + g), 0));
        BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->h2d, get_event(ctx, 8 + g), 0));
      }
      if (colmaj) BOF_TRY(copy2d(ctx, c_io, (size_t)rows * 4, c + j0 * out_rows + r0, (size_t)out_rows * 4, (size_t)rows * 4, (size_t)kbp, H2D, ctx->h2d));
      else BOF_TRY(copy2d(ctx, c_io, (size_t)kbp * 4, c + r0 * k + j0, (size_t)k * 4, (size_t)kbp * 4, (size_t)rows, H2D, ctx->h2d));
      BOF_CUDA(ctx, cudaEventRecord(get_event(ctx, 4 + g), ctx->h2d));
      BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, get_event(ctx, 4 + g), 0));
    }
    if (i == 0) {
      BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, get_event(ctx, 0 + gp), 0));
      if (colmaj && in_rows > 0)
        BOF_TRY(launch_transpose(ctx, ctx->compute, kbp, in_rows, braw_all + (size_t)gp * pan_elems, in_rows, bp, kbp));
    }
    if (blk_used[g]) BOF_CUDA(ctx, cudaStreamWaitEvent(ctx->compute, get_event(ctx, 8 + g), 0));
    if (colmaj) {
      BOF_TRY(launch_spmm_rm(ctx, ctx->compute, rows, kbp, 1.f, vals + z0, idx + z0, offs + r0, bp, kbp, 0.f, cblk[g], kbp));
      BOF_TRY(launch_transpose_axpby(ctx, ctx->compute, rows, kbp, alpha, cblk[g], kbp, beta, cblk_t[g], rows));
    } else {
      BOF_TRY(launch_spmm_rm(ctx, ctx->compute, rows, kbp, alpha, vals + z0, idx + z0, offs + r0, bp, kbp, beta, cblk[g], kbp));
    }
    BOF_CUDA(ctx, cudaEventRecord(get_event(ctx, 6 + g), ctx->compute));
    if (i == nblk - 1) BOF_CUDA(ctx, cudaEventRecord(get_event(ctx, 2 + gp), ctx->compute));
    blk_used[g] = true;
    return BOF_OK;
  };
  auto fetch_block = [&](int p, int i, int g) -> int {
    const int64_t j0 = (int64_t)p * kb, kbp = std::min(kb, k - j0);
    const int64_t r0 = (int64_t)i * rows_blk, rows = std::min(rows_blk, out_rows - r0);
    float* c_io = colmaj ? cblk_t[g] : cblk[g];
    cudaEvent_t ev_done = get_event(ctx, 6 + g), ev_down = get_event(ctx, 8 + g);
    uint64_t* tk = &down_ticket[g];
    if (colmaj) BOF_TRY(d2h_transfer(ctx, c + j0 * out_rows + r0, (size_t)out_rows * 4, c_io, (size_t)rows * 4, (size_t)rows * 4, (size_t)kbp, ctx->d2h, ev_done, ev_down, tk));
    else if (npan == 1) BOF_TRY(d2h_transfer(ctx, c + r0 * k, (size_t)rows * k * 4, c_io, (size_t)rows * k * 4, (size_t)rows * k * 4, 1, ctx->d2h, ev_done, ev_down, tk));
    else BOF_TRY(d2h_transfer(ctx, c + r0 * k + j0, (size_t)k * 4, c_io, (size_t)kbp * 4, (size_t)kbp * 4, (size_t)rows, ctx->d2h, ev_done, ev_down, tk));
    return BOF_OK;
  };
  // software pipeline over (panel, block): launch step s, then download step s-1
  BOF_TRY(upload_panel(0));
  int prev_p = -1, prev_i = -1, prev_g = -1;
  for (int p = 0; p < npan; ++p) {
    if (beta == 0.f && p + 1 < npan) BOF_TRY(upload_panel(p + 1));
    for (int i = 0; i < nblk; ++i, ++bc) {
      const int g = bc & 1;
      BOF_TRY(run_block(p, i, g));
      if (prev_p >= 0) BOF_TRY(fetch_block(prev_p, prev_i, prev_g));
      prev_p = p; prev_i = i; prev_g = g;
    }
    // with beta != 0 the old-C uploads share the h2d stream, so the next panel is queued behind them
    if (beta != 0.f && p + 1 < npan) BOF_TRY(upload_panel(p + 1));
  }
  if (prev_p >= 0) BOF_TRY(fetch_block(prev_p, prev_i, prev_g));
  BOF_TRY(sync_all(ctx));
  stats_end(ctx);
  return call_guard.done();
}

// y = op(A) x with host x, y.  'T' uses the resident A^T when it has been built (a deterministic gather SpMV),
// else the scatter kernel on A.
int bof_csr_mv(bof_csr* h, char trans_a, const float* x, float* y) {
  if (!h) return BOF_EINVAL;
  bof_ctx* ctx = h->ctx;
  BOF_REQUIRE(ctx, is_nt(trans_a), "csrgemv trans_a error : expected=N or T, found=%c", trans_a);
  BOF_CUDA(ctx, cudaSetDevice(ctx->device));
  stats_begin(ctx);
  CallGuard call_guard(ctx);
  const bool tr = trans_a == 'T';
  const int64_t xlen = tr ? h->m : h->n, ylen = tr ? h->n : h->m;
  if (ylen == 0) { stats_end(ctx); return call_guard.done(); }
  float *xd, *yd;
  BOF_TRY(slot_reserve(ctx, S_MISC, (size_t)std::max<int64_t>(xlen, 1), &xd));
  BOF_TRY(slot_reserve(ctx, S_OUT0, (size_t)ylen, &yd));
  cudaStream_t s = ctx->compute;
  BOF_TRY(copy1d(ctx, xd, x, (size_t)xlen * 4, cudaMemcpyHostToDevice, s));
  if (!tr) BOF_TRY(launch_spmv(ctx, s, 'N', h->m, h->n, h->vals[0], h->idx[0], h->offs[0], xd, yd));
  else if (h->have_t) BOF_TRY(launch_spmv(ctx, s, 'N', h->n, h->m, h->vals[1], h->idx[1], h->offs[1], xd, yd));
  else BOF_TRY(launch_spmv(ctx, s, 'T', h->m, h->n, h->vals[0], h->idx[0], h->offs[0], xd, yd));
  BOF_TRY(copy1d(ctx, y, yd, (size_t)ylen * 4, cudaMemcpyDeviceToHost, s));
  BOF_TRY(sync_all(ctx));
  stats_end(ctx);
  return call_guard.done();
}

int bof_csr_close(bof_csr* h) {
  if (!h) return BOF_OK;
  cudaSetDevice(h->ctx->device);
  cudaDeviceSynchronize();
  csr_free(h);
  return BOF_OK;
}

// ---- k-means: points shard resident across iterations -------------------------------------------

struct bof_kmeans {
  bof_ctx* ctx;
  int64_t npoints, ncenters, dim;
  float* points;        // P x dim
  void* point_planes;   // TF32 hi/lo planes of the points (filled once)
  float* p_l2sq;        // P
  float* centers;       // K x dim
  float* c_l2sq;        // K
  float* partial;       // [K*dim sums | K counts]
  int32_t* assign;      // P
  void* ws_assign; size_t ws_assign_bytes;
  void* ws_reduce; size_t ws_reduce_bytes;
  int64_t* assign64;    // P, for bof_kmeans_get
};

static void kmeans_free(bof_kmeans* km) {
  if (!km) return;
  void* ptrs[] = {km->points, km->point_planes, km->p_l2sq, km->centers, km->c_l2sq, km->partial,
                  km->assign, km->ws_assign, km->ws_reduce, km->assign64};
  for (void* p : ptrs) if (p) cudaFree(p);
  delete km;
}

int bof_kmeans_open(bof_ctx* ctx, int64_t npoints, int64_t ncenters, int64_t dim, const float* points_host,
                    const float* centers_host, bof_kmeans** out) {
  if (!ctx || !out) return BOF_EINVAL;
  *out = nullptr;
  BOF_REQUIRE(ctx, npoints >= 0 && ncenters > 0 && dim > 0 && npoints < (1ll << 31), "kmeans: bad dimension");
  BOF_CUDA(ctx, cudaSetDevice(ctx->device));
  bof_kmeans* km = new bof_kmeans();
  km->ctx = ctx; km->npoints = npoints; km->ncenters = ncenters; km->dim = dim;
  const size_t P = (size_t)std::max<int64_t>(npoints, 1);
  km->ws_assign_bytes = bof_kmeans_workspace_bytes(npoints, ncenters, dim, 0);
  km->ws_reduce_bytes = kmeans_reduce_workspace_bytes(npoints, ncenters, dim);
  struct Req { void** p; size_t bytes; } reqs[] = {
      {(void**)&km->points, P * dim * 4}, {&km->point_planes, bof_kmeans_point_planes_bytes(npoints, dim)},
      {(void**)&km->p_l2sq, P * 4}, {(void**)&km->centers, (size_t)ncenters * dim * 4},
      {(void**)&km->c_l2sq, (size_t)ncenters * 4}, {(void**)&km->partial, ((size_t)ncenters * dim + ncenters) * 4},
      {(void**)&km->assign, P * 4}, {&km->ws_assign, km->ws_assign_bytes}, {&km->ws_reduce, km->ws_reduce_bytes},
      {(void**)&km->assign64, P * 8}};
  for (auto& r : reqs) {
    if (cudaMalloc(r.p, r.bytes) != cudaSuccess) {
      cudaGetLastError();
      kmeans_free(km);
      return fail(ctx, BOF_ENOMEM, "kmeans: cudaMalloc of %zu bytes failed", r.bytes);
    }
  }
  cudaStream_t s = ctx->compute;
  auto guard = [&](int rc) { if (rc != BOF_OK) { quiesce(ctx); kmeans_free(km); } return rc; };
  if (int rc = guard(copy1d(ctx, km->points, points_host, (size_t)npoints * dim * 4, cudaMemcpyHostToDevice, s))) return rc;
  if (int rc = guard(copy1d(ctx, km->centers, centers_host, (size_t)ncenters * dim * 4, cudaMemcpyHostToDevice, s))) return rc;
  if (int rc = guard(launch_row_sqnorm(ctx, s, npoints, dim, km->points, dim, km->p_l2sq))) return rc;
  if (int rc = guard(launch_row_sqnorm(ctx, s, ncenters, dim, km->centers, dim, km->c_l2sq))) return rc;
  if (int rc = guard(bof_kmeans_prepare_points(ctx, s, npoints, dim, km->points, km->point_planes))) return rc;
  if (cudaStreamSynchronize(s) != cudaSuccess) { kmeans_free(km); return fail(ctx, BOF_ECUDA, "kmeans: upload failed"); }
  *out = km;
  return BOF_OK;
}

int bof_kmeans_local_step(bof_kmeans* km, void** dev_partial, size_t* partial_floats) {
  if (!km) return BOF_EINVAL;
  bof_ctx* ctx = km->ctx;
  cudaStream_t s = ctx->compute;
  BOF_TRY(bof_kmeans_assign(ctx, s, km->npoints, km->ncenters, km->dim, km->points, km->centers, km->c_l2sq,
                            km->p_l2sq, km->assign, km->point_planes, km->ws_assign, km->ws_assign_bytes));
  BOF_TRY(launch_kmeans_reduce_ws(ctx, s, km->npoints, km->ncenters, km->dim, km->points, km->assign, km->partial,
                                  km->partial + km->ncenters * km->dim, km->ws_reduce, km->ws_reduce_bytes));
  if (dev_partial) *dev_partial = km->partial;
  if (partial_floats) *partial_floats = (size_t)km->ncenters * km->dim + km->ncenters;
  return BOF_OK;
}

int bof_kmeans_update(bof_kmeans* km) {
  if (!km) return BOF_EINVAL;
  return launch_kmeans_finalize(km->ctx, km->ctx->compute, km->ncenters, km->dim, km->partial,
                                km->partial + km->ncenters * km->dim, km->centers, km->c_l2sq);
}

int bof_kmeans_get(bof_kmeans* km, float* centers_host, int64_t* assign_host) {
  if (!km) return BOF_EINVAL;
  bof_ctx* ctx = km->ctx;
  cudaStream_t s = ctx->compute;
  CallGuard call_guard(ctx);
  if (centers_host) BOF_TRY(copy1d(ctx, centers_host, km->centers, (size_t)km->ncenters * km->dim * 4, cudaMemcpyDeviceToHost, s));
  if (assign_host && km->npoints > 0) {
    BOF_TRY(launch_idx_widen(ctx, s, km->assign, km->assign64, km->npoints));
    BOF_TRY(copy1d(ctx, assign_host, km->assign64, (size_t)km->npoints * 8, cudaMemcpyDeviceToHost, s));
  }
  BOF_TRY(sync_all(ctx));
  return call_guard.done();
}

void* bof_kmeans_stream(bof_kmeans* km) { return km ? (void*)km->ctx->compute : nullptr; }

int bof_kmeans_close(bof_kmeans* km) {
  if (!km) return BOF_OK;
  cudaStreamSynchronize(km->ctx->compute);
  kmeans_free(km);
  return BOF_OK;
}

}  // extern "C"
