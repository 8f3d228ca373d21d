#include <fcntl.h>
#include <sys/mman.h>
#include <unistd.h>
#include <chrono>
#include <condition_variable>
#include <cstdio>
#include <cstring>
#include <functional>
#include <mutex>
#include <thread>
#include <vector>
#include <cstdlib>
#include <cerrno>
class CopyPool {
 public:
  explicit CopyPool(int n) { for (int i = 0; i < n; ++i) workers_.emplace_back([this, i] { loop(i); }); }
  ~CopyPool() { { std::lock_guard<std::mutex> lk(mu_); stop_ = true; } cv_.notify_all(); for (auto& t : workers_) t.join(); }
  int size() const { return (int)workers_.size(); }
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
      { std::unique_lock<std::mutex> lk(mu_); cv_.wait(lk, [&] { return stop_ || epoch_ != seen; }); if (stop_) return; seen = epoch_; if (id >= parts_) continue; fn = fn_; }
      (*fn)(id);
      { std::lock_guard<std::mutex> lk(mu_); if (--pending_ == 0) done_cv_.notify_all(); }
    }
  }
  std::vector<std::thread> workers_; std::mutex mu_; std::condition_variable cv_, done_cv_;
  const std::function<void(int)>* fn_ = nullptr; int parts_ = 0, pending_ = 0; uint64_t epoch_ = 0; bool stop_ = false;
};
double now() { return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count(); }
int main(int argc, char** argv) {
  const size_t total = 2ull << 30, chunk = 16u << 20;
  int fd = open("/dev/shm/pooltest.bin", O_RDWR | O_CREAT, 0666); ftruncate(fd, total);
  std::vector<char> init(chunk, 1); for (size_t o = 0; o < total; o += chunk) pwrite(fd, init.data(), chunk, o);
  char* buf = (char*)aligned_alloc(4096, chunk); memset(buf, 0, chunk);
  for (int nthr : {1, 4, 8}) {
    CopyPool pool(nthr);
    double t0 = now();
    for (size_t o = 0; o < total; o += chunk) {
      int parts = nthr;
      pool.run(parts, [&](int p) { size_t b0 = chunk * p / parts, b1 = chunk * (p + 1) / parts; size_t len = b1 - b0; char* d = buf + b0; size_t off = o + b0;
        while (len) { ssize_t n = pread(fd, d, len, off); if (n <= 0) break; d += n; off += n; len -= n; } });
    }
    double t = now() - t0; printf("pread threads=%d: %.2f GB/s\n", nthr, total / t / 1e9);
    char* m = (char*)mmap(nullptr, total, PROT_READ, MAP_SHARED, fd, 0);
    t0 = now();
    for (size_t o = 0; o < total; o += chunk) { int parts = nthr; pool.run(parts, [&](int p) { size_t b0 = chunk * p / parts, b1 = chunk * (p + 1) / parts; memcpy(buf + b0, m + o + b0, b1 - b0); }); }
    t = now() - t0; printf("mmap-memcpy threads=%d: %.2f GB/s\n", nthr, total / t / 1e9);
    munmap(m, total);
  }
  unlink("/dev/shm/pooltest.bin");
}
