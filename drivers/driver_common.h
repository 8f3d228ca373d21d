// Shared helpers of the command-line drivers.  The drivers keep the positional CLIs of the reference's
// drivers/*.cpp (same argument order and meaning, cited per driver) so that scripts such as
// misc/gemm_run.sh keep working; the work itself goes through include/flash_blas.h.
#pragma once

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <string>
#include <vector>

#include "flash_blas.h"
#include "lib_funcs.h"

namespace drv {

inline void usage_exit(const char* text) {
  std::fprintf(stderr, "usage : %s\n", text);
  std::exit(2);
}

inline FBLAS_UINT to_u(const char* s) { return (FBLAS_UINT)std::stoll(s); }
inline FPTYPE to_f(const char* s) { return (FPTYPE)std::stof(s); }

template <typename T>
std::vector<T> read_file(const std::string& name, size_t count) {
  std::vector<T> v(count);
  std::ifstream in(name, std::ios::binary);
  if (!in || !in.read(reinterpret_cast<char*>(v.data()), (std::streamsize)(count * sizeof(T)))) {
    std::fprintf(stderr, "cannot read %zu values from %s\n", count, name.c_str());
    std::exit(1);
  }
  return v;
}

template <typename T>
void write_file(const std::string& name, const std::vector<T>& v) {
  std::ofstream out(name, std::ios::binary);
  out.write(reinterpret_cast<const char*>(v.data()), (std::streamsize)(v.size() * sizeof(T)));
}

struct StopWatch {
  std::chrono::steady_clock::time_point t0 = std::chrono::steady_clock::now();
  double seconds() const { return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count(); }
};

inline void report(const char* what, double secs, FBLAS_INT rc) {
  bof_stats st{};
  if (bof_ctx* ctx = flash::flash_context()) bof_get_stats(ctx, &st);
  std::printf("%s took %.3f s, returned %lld; h2d %.1f MB, d2h %.1f MB, kernels %lld, host staging in %.0f ms / out %.0f ms, "
              "library wall %.0f ms\n", what, secs, (long long)rc, st.h2d_bytes / 1e6, st.d2h_bytes / 1e6,
              (long long)st.kernel_launches, st.stage_in_ms, st.stage_out_ms, st.total_ms);
}

}  // namespace drv
