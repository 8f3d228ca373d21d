// Host <-> device staging for the tile pipelines: pinned-memory fast path, the pinned ring with copy workers for
// pageable memory (the mmap behind a flash_ptr), the background device->host job queue, the context's slot arena,
// stream/event helpers and the BOF_TRACE timeline.  Together with the pipelines this replaces, for the hot path,
// the reference's scheduler / cache / io_executor / file_handle stack (src/scheduler/*.cpp, src/file_handles/*.cpp).
#include "host_internal.cuh"

#include <sys/mman.h>
#include <unistd.h>

#include <chrono>
#include <condition_variable>
#include <cstring>
#include <deque>
#include <functional>
#include <map>
#include <thread>

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

namespace bof {

double now_ms() {
  using namespace std::chrono;
  return duration<double, std::milli>(steady_clock::now().time_since_epoch()).count();
}


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
void trace_host(bof_ctx* ctx, const char* what, long idx) {
  if (!trace_on()) return;
  std::lock_guard<std::mutex> lk(ctx->host_trace_mu);
  ctx->host_trace.push_back({now_ms() - ctx->trace_t0, what, idx});
}
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
struct FileRange { size_t len; int fd; uint64_t file_off; bool pinned = false; };
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
#ifndef MADV_POPULATE_READ
#define MADV_POPULATE_READ 22
#define MADV_POPULATE_WRITE 23
#endif
// Map the pages of a host range in one kernel call before touching them: a freshly mmap'd file (the flash_ptr
// case) otherwise takes one minor fault per 4 KiB page inside the memcpy.  Opt-in (BOF_POPULATE=1): the A/B on
// /dev/shm files was mixed (profiles/r01/populate_ab.txt: gemm 0.83 -> 0.60 s, csrmm 1.08 -> 1.35 s, single runs).
void populate_range(char* p, size_t len, bool for_write) {
  static const bool on = getenv("BOF_POPULATE") && atoi(getenv("BOF_POPULATE")) != 0;
  static const uintptr_t page = (uintptr_t)sysconf(_SC_PAGESIZE);
  if (!on || len < (64u << 10)) return;
  const uintptr_t a = reinterpret_cast<uintptr_t>(p) & ~(page - 1);
  const uintptr_t b = (reinterpret_cast<uintptr_t>(p) + len + page - 1) & ~(page - 1);
  (void)madvise(reinterpret_cast<void*>(a), b - a, for_write ? MADV_POPULATE_WRITE : MADV_POPULATE_READ);  // best effort
}

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
      populate_range(host + b0, b1 - b0, !to_packed);
      if (to_packed) std::memcpy(packed + b0, host + b0, b1 - b0);
      else std::memcpy(host + b0, packed + b0, b1 - b0);
    } else {
      const size_t r0 = rows * part / parts, r1 = rows * (part + 1) / parts;
      if (!via_fd && r1 > r0 && width * 2 >= hpitch)  // dense enough that most pages of the span are touched
        populate_range(host + r0 * hpitch, (r1 - r0 - 1) * hpitch + width, !to_packed);
      for (size_t r = r0; r < r1; ++r) {
        if (via_fd && file_xfer(!to_packed, fd, packed + r * width, width, foff + r * hpitch)) continue;
        if (to_packed) std::memcpy(packed + r * width, host + r * hpitch, width);
        else std::memcpy(host + r * hpitch, packed + r * width, width);
      }
    }
  });
  (to_packed ? ctx->stats.stage_in_ms : ctx->stats.stage_out_ms) += now_ms() - t0;
}

void drain_copy_out(bof_ctx* ctx, StageSlot* sl) {
  if (sl->out_dst)
    host_rows_copy(ctx, static_cast<char*>(sl->ptr), sl->out_dst, sl->out_pitch, sl->out_width, sl->out_rows, false);
}
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
// Ordering contract of the pageable path: the copies are enqueued on `s` LATER, from the drainer thread, so they are
// NOT ordered against work the caller enqueues on `s` after this call returns -- the source buffer must stay
// untouched (and `wait_ev` unre-recorded) until d2h_fence(ticket) or sync_all().  Every call site keeps to that:
// block buffers are reused only behind their ticket's fence, one-shot downloads are followed by sync_all().
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
  if (ctx->coll) BOF_CUDA(ctx, cudaStreamSynchronize(ctx->coll));
  comm_sync_pushes(ctx);
  BOF_CUDA(ctx, cudaStreamSynchronize(ctx->compute));
  BOF_CUDA(ctx, cudaStreamSynchronize(ctx->d2h));
  return drain_wait(ctx);
}

// After a failed call nothing of it may still be running: queued copies and the drainer thread reference the
// caller's host buffers and the context's slots.  Keeps the recorded error message.
void quiesce(bof_ctx* ctx) {
  if (ctx->drainer) ctx->drainer->wait_all_issued();
  cudaStreamSynchronize(ctx->h2d);
  if (ctx->coll) cudaStreamSynchronize(ctx->coll);
  comm_sync_pushes(ctx);
  cudaStreamSynchronize(ctx->compute);
  cudaStreamSynchronize(ctx->d2h);
  if (ctx->drainer) ctx->drainer->wait_idle();
  for (auto& sl : ctx->stage_in) sl.in_flight = false;
  cudaGetLastError();
}


void staging_destroy(bof_ctx* ctx) {
  delete ctx->drainer;  // joins its thread before the rings go away
  delete ctx->pool;
  delete ctx->pool_out;
  ctx->drainer = nullptr; ctx->pool = nullptr; ctx->pool_out = nullptr;
  for (auto* ring : {&ctx->stage_in, &ctx->stage_out}) {
    for (auto& sl : *ring) {
      if (sl.ptr) cudaFreeHost(sl.ptr);
      if (sl.ev) cudaEventDestroy(sl.ev);
    }
    ring->clear();
  }
}

}  // namespace bof

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

using namespace bof;

extern "C" {

// BOF_PIN_MAPPINGS=0 switches the page-locking of file mappings off (staging through the pinned ring is used then)
static bool pin_mappings() {
  static const bool on = [] { const char* e = getenv("BOF_PIN_MAPPINGS"); return e == nullptr || atoi(e) != 0; }();
  return on;
}

int bof_register_mapping(const void* base, size_t len, int fd, uint64_t file_offset) {
  if (base == nullptr || len == 0 || fd < 0) return BOF_EINVAL;
  FileRange fr{len, fd, file_offset};
  // Page-lock the mapping and make it DMA-able: the page-cache pages behind a MAP_SHARED mapping are then what the
  // copy engines read and write -- file -> HBM and HBM -> file without the pinned bounce buffer, the host memcpy and
  // the page faults of a fresh mapping inside the timed call (the faults happen here, once, at map time).  The
  // counterpart of the reference's O_DIRECT reads into its own aligned buffers
  // (src/file_handles/flash_file_handle.cpp:247-407), minus the buffer.  Needs a current CUDA context (flash_setup
  // creates it before the files are mapped); on failure -- no context, locked-memory limit, a filesystem whose pages
  // cannot be pinned -- the mapping simply stays pageable and is staged as before.
  int ndev = 0;
  if (pin_mappings() && cudaGetDeviceCount(&ndev) == cudaSuccess && ndev > 0) {
    const double t0 = now_ms();
    cudaError_t e = cudaHostRegister(const_cast<void*>(base), len, cudaHostRegisterPortable);
    if (e == cudaSuccess) {
      fr.pinned = true;
      if (trace_on()) std::fprintf(stderr, "[bof host ] mapping of %zu bytes page-locked in %.1f ms\n", len, now_ms() - t0);
    } else {
      cudaGetLastError();
      if (trace_on()) std::fprintf(stderr, "[bof host ] cudaHostRegister of a %zu-byte mapping failed (%s): staged copies\n", len, cudaGetErrorString(e));
    }
  } else {
    cudaGetLastError();
  }
  std::lock_guard<std::mutex> lk(g_map_mu);
  g_mappings[reinterpret_cast<uintptr_t>(base)] = fr;
  return BOF_OK;
}

int bof_unregister_mapping(const void* base) {
  std::lock_guard<std::mutex> lk(g_map_mu);
  auto it = g_mappings.find(reinterpret_cast<uintptr_t>(base));
  if (it == g_mappings.end()) return BOF_EINVAL;
  if (it->second.pinned) {
    cudaDeviceSynchronize();   // nothing may still be reading or writing the pages
    cudaHostUnregister(const_cast<void*>(base));
    cudaGetLastError();
  }
  g_mappings.erase(it);
  return BOF_OK;
}

}  // extern "C"
