// K6/K7 (stable CSR -> CSC) and K10/K11 (k-means centroid reduce, row norms) for sm_100a.
//
// Reference bodies replaced:
//   mkl_scsrcsc + index rebase      include/tasks/csrcsc_task.h:68-80
//   row-block-ordered segment merge include/tasks/csrcsc_task.h:143-162
//   bucket + cblas_saxpy loop       drivers/in_mem_kmeans.cpp:105-125
//   cblas_sdot row norms            drivers/in_mem_kmeans.cpp:75-78,179-182
//
// The transpose is the unique *stable* counting sort of the nonzeros by column.  It is an LSD radix sort with
// WIDE digits (up to 12 bits), so that the 23-bit columns of the 8M x 8M matrix take two passes instead of
// three, each pass = histogram kernel -> device-wide exclusive scan -> ranked scatter ("histogram-plus-scan
// scatter"; no atomic anywhere, bit-reproducible: within a column the entries keep ascending source row and
// duplicates keep storage order -- what mkl_scsrcsc on row blocks + the reference's in-order merge produce).
//
//  * A CTA owns a SUPERTILE (up to 16 tiles of 8192 items) and keeps one running output cursor per digit in
//    shared memory, so the histogram has 4096 counters per 131072 items instead of per tile (4 % of the data).
//  * Inside a tile the items are ranked stably by the wide digit in two levels of 6 bits (7 ballots each on
//    warp-private counters), then staged through shared memory in sorted order so that every run of equal
//    digits leaves the SM as one contiguous segment.
//  * Pass 1 derives the source row of every nonzero from the CSR offsets on the fly (no materialised row array).
//  * Pass 2 cuts its supertiles at the boundaries of the pass-1 buckets; the scanned pass-2 histogram then
//    holds, at the first supertile of every bucket, exactly the CSC offset of (high digit, low digit): the
//    n + 1 output offsets are a gather from it -- no sorted-key array is written or re-read.
//  * The last pass writes only the payloads (row ids, values).
// Traffic at 8M x 8M: 4 + 1 + 8 + 12 (pass 1) + 4 + 1 + 12 + 8 (pass 2) = 50 B/nnz against 88 B/nnz of the
// three-pass 8-bit version of round 1 (single-pass ideal of SURVEY.md 8d: 20 B/nnz).
#include "common.cuh"

#include <algorithm>
#include <cstdlib>

namespace bof {
namespace {

constexpr int RS_THREADS = 512;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_IPT = 16;                     // items per thread
constexpr int RS_TILE = RS_THREADS * RS_IPT;   // 8192 items per tile
constexpr int RS_WARP_ITEMS = 32 * RS_IPT;     // a warp owns 512 consecutive positions of the tile
constexpr int RS_MAX_BITS = 12;
constexpr int RS_MAX_BINS = 1 << RS_MAX_BITS;
constexpr int RS_LVL_DIG = 64;                 // digits per ranking level (6 bits)
constexpr int RS_MAX_SUB = 16;                 // tiles per supertile

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

struct SortSmem {
  uint32_t cnt[RS_WARPS][RS_LVL_DIG + 1];  // warp-private digit counters of the current ranking level
  uint32_t dig_off[RS_LVL_DIG];
  uint32_t warp_tot[2];
  union {
    struct {
      uint16_t words[RS_TILE];   // wide digits in level-A order
      uint16_t slotb[RS_TILE];   // level-A position -> final slot
    } a;
    uint32_t stage[RS_TILE];     // one array of the tile in sorted order, on its way out
  } u;
  uint16_t sdig[RS_TILE];         // wide digit of every sorted slot
  uint32_t running[RS_MAX_BINS];  // per digit: next global output position (scatter) / count (histogram)
};

struct RowSmem {
  uint32_t rowid[RS_TILE];
  uint32_t wmax[RS_WARPS];
  long long r_lo, r_hi;
};

// Stable ranks of this thread's items by a 6-bit digit.  Item r of a thread sits at tile position
// warp * 512 + r * 32 + lane; pos[r] = number of valid items with a smaller digit, or with the same digit at a
// lower position.  Seven ballots per item on warp-private counters with one writer per digit
// (__match_any_sync measured ~10x slower than the ballots on sm_100, profiles/r01).
template <class DigitFn>
__device__ __forceinline__ void rank_level(SortSmem& sm, DigitFn digit, uint32_t valid_bits, uint32_t (&pos)[RS_IPT]) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int i = tid; i < RS_WARPS * (RS_LVL_DIG + 1); i += RS_THREADS) (&sm.cnt[0][0])[i] = 0;
  __syncthreads();
  const unsigned lt = lanemask_lt();
#pragma unroll
  for (int r = 0; r < RS_IPT; ++r) {
    const bool valid = (valid_bits >> r) & 1u;
    const uint32_t d = digit(r);
    unsigned peers = __ballot_sync(0xffffffffu, valid);
#pragma unroll
    for (int b = 0; b < 6; ++b) {
      const bool bit = (d >> b) & 1u;
      const unsigned bal = __ballot_sync(0xffffffffu, bit);
      peers &= bit ? bal : ~bal;
    }
    uint32_t old = 0;
    if (valid) old = sm.cnt[warp][d];
    __syncwarp();
    if (valid && (peers & lt) == 0) sm.cnt[warp][d] = old + __popc(peers);
    __syncwarp();
    pos[r] = old + __popc(peers & lt);
  }
  __syncthreads();
  // per digit: exclusive prefix over the warps (in place); digit totals -> exclusive scan over the 64 digits
  uint32_t total = 0;
  if (tid < RS_LVL_DIG) {
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      const uint32_t t = sm.cnt[w][tid];
      sm.cnt[w][tid] = total;
      total += t;
    }
  }
  uint32_t incl = total;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += v;
  }
  if (tid < RS_LVL_DIG && lane == 31) sm.warp_tot[warp] = incl;
  __syncthreads();
  if (tid < RS_LVL_DIG) sm.dig_off[tid] = (warp == 1 ? sm.warp_tot[0] : 0u) + incl - total;
  __syncthreads();
#pragma unroll
  for (int r = 0; r < RS_IPT; ++r) {
    if ((valid_bits >> r) & 1u) {
      const uint32_t d = digit(r);
      pos[r] += sm.dig_off[d] + sm.cnt[warp][d];
    }
  }
}

// Stable sort of one tile by a digit of up to 12 bits (two levels of 6).  On return sm.sdig[s] is the digit of
// sorted slot s (s < count) and, if WANT_FS, fs[r] the slot of the thread's item r.  The caller must
// __syncthreads() before it overwrites sm.u.stage.
template <bool TWO, bool WANT_FS, class DigitFn>
__device__ __forceinline__ void local_sort(SortSmem& sm, DigitFn wide_digit, uint32_t valid_bits, int count,
                                           uint32_t (&fs)[RS_IPT]) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  rank_level(sm, [&](int r) { return wide_digit(r) & 63u; }, valid_bits, fs);
  if constexpr (!TWO) {
#pragma unroll
    for (int r = 0; r < RS_IPT; ++r)
      if ((valid_bits >> r) & 1u) sm.sdig[fs[r]] = (uint16_t)wide_digit(r);
    __syncthreads();
  } else {
#pragma unroll
    for (int r = 0; r < RS_IPT; ++r)
      if ((valid_bits >> r) & 1u) sm.u.a.words[fs[r]] = (uint16_t)wide_digit(r);
    __syncthreads();
    // level B runs over the level-A order: position q holds the q-th item by low digit
    const int q0 = warp * RS_WARP_ITEMS + lane;
    uint32_t vb = 0;
#pragma unroll
    for (int r = 0; r < RS_IPT; ++r) vb |= (uint32_t)(q0 + r * 32 < count) << r;
    uint32_t posb[RS_IPT];
    rank_level(sm, [&](int r) { return (uint32_t)(sm.u.a.words[min(q0 + r * 32, RS_TILE - 1)] >> 6); }, vb, posb);
#pragma unroll
    for (int r = 0; r < RS_IPT; ++r) {
      if ((vb >> r) & 1u) {
        const int q = q0 + r * 32;
        sm.sdig[posb[r]] = sm.u.a.words[q];
        if constexpr (WANT_FS) sm.u.a.slotb[q] = (uint16_t)posb[r];
      }
    }
    __syncthreads();
    if constexpr (WANT_FS) {
#pragma unroll
      for (int r = 0; r < RS_IPT; ++r)
        if ((valid_bits >> r) & 1u) fs[r] = sm.u.a.slotb[fs[r]];
    }
  }
}

struct PassArgs {
  const uint32_t* key_in;
  const uint32_t* p1_in;       // payload 1 (P1_ARRAY)
  const uint32_t* p2_in;       // payload 2 or nullptr
  uint32_t* key_out;           // nullptr: keys are not written (last pass)
  uint32_t* p1_out;
  uint32_t* p2_out;
  const int64_t* row_offs;     // P1_ROWS: CSR offsets (m + 1), payload 1 = source row of the item
  int64_t m;
  int64_t n;                   // items
  int shift, bits;
  uint32_t* counts;            // histogram out / scanned offsets in: [digit * nsuper + supertile]
  uint32_t nsuper;             // stride of `counts` (upper bound of the supertile count)
  const int64_t* super_start;  // nsuper + 1 supertile boundaries, or nullptr: uniform supertiles of super_len
  const uint32_t* nsuper_actual;  // device scalar with the supertile count actually used, or nullptr
  int64_t super_len;
};

__device__ __forceinline__ bool supertile_range(const PassArgs& a, int64_t& sbeg, int64_t& send) {
  const uint32_t b = blockIdx.x;
  if (a.nsuper_actual != nullptr && b >= *a.nsuper_actual) return false;
  if (a.super_start != nullptr) {
    sbeg = a.super_start[b];
    send = a.super_start[b + 1];
  } else {
    sbeg = (int64_t)b * a.super_len;
    send = min(a.n, sbeg + a.super_len);
  }
  return sbeg < send;
}

// counts[digit * nsuper + supertile] = number of items of the supertile with that digit
template <bool TWO>
__global__ void __launch_bounds__(RS_THREADS, 2) radix_hist_kernel(const PassArgs a) {
  extern __shared__ __align__(16) uint8_t rs_smem_raw[];
  SortSmem& sm = *reinterpret_cast<SortSmem*>(rs_smem_raw);
  int64_t sbeg, send;
  if (!supertile_range(a, sbeg, send)) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nbins = 1 << a.bits;
  const uint32_t mask = (uint32_t)nbins - 1u;
  for (int d = tid; d < nbins; d += RS_THREADS) sm.running[d] = 0;
  for (int64_t base = sbeg; base < send; base += RS_TILE) {
    const int count = (int)min((int64_t)RS_TILE, send - base);
    const int p0 = warp * RS_WARP_ITEMS + lane;
    uint32_t key[RS_IPT];
    uint32_t valid_bits = 0;
#pragma unroll
    for (int r = 0; r < RS_IPT; ++r) {
      const int p = p0 + r * 32;
      const bool ok = p < count;
      key[r] = ok ? __ldcs(a.key_in + base + p) : 0u;
      valid_bits |= (uint32_t)ok << r;
    }
    uint32_t fs[RS_IPT];
    local_sort<TWO, false>(sm, [&](int r) { return (key[r] >> a.shift) & mask; }, valid_bits, count, fs);
    // run heads subtract their slot, run tails add theirs + 1: every digit gains its run length
    for (int s = tid; s < count; s += RS_THREADS) {
      const uint16_t d = sm.sdig[s];
      if (s == 0 || sm.sdig[s - 1] != d) sm.running[d] -= (uint32_t)s;
    }
    __syncthreads();
    for (int s = tid; s < count; s += RS_THREADS) {
      const uint16_t d = sm.sdig[s];
      if (s == count - 1 || sm.sdig[s + 1] != d) sm.running[d] += (uint32_t)s + 1u;
    }
    __syncthreads();
  }
  for (int d = tid; d < nbins; d += RS_THREADS) a.counts[(size_t)d * a.nsuper + blockIdx.x] = sm.running[d];
}

enum { P1_ARRAY = 0, P1_ROWS = 1, P1_IOTA = 2 };

// rowid[p] = CSR row of global item base + p, for p < count (offsets may be un-rebased: offs[0] != 0)
__device__ __forceinline__ void tile_row_ids(RowSmem& rs, const int64_t* __restrict__ offs, int64_t m, int64_t base,
                                             int count) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t off0 = offs[0];
  if (tid < 2) {
    // last row r with offs[r] - off0 <= g: the row that holds item g
    const int64_t g = off0 + (tid == 0 ? base : base + count - 1);
    int64_t lo = 0, hi = m;  // invariant: offs[lo] <= g, offs[hi] > g (offs[m] = off0 + nnz > g)
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (offs[mid] <= g) lo = mid; else hi = mid;
    }
    if (tid == 0) rs.r_lo = lo; else rs.r_hi = lo;
  }
  for (int p = tid; p < RS_TILE; p += RS_THREADS) rs.rowid[p] = 0;
  __syncthreads();
  const int64_t r_lo = rs.r_lo, r_hi = rs.r_hi;
  if (tid == 0) rs.rowid[0] = (uint32_t)r_lo;
  for (int64_t r = r_lo + 1 + tid; r <= r_hi; r += RS_THREADS) {
    const int64_t st = offs[r] - off0 - base;
    if (offs[r + 1] > offs[r] && st < count) rs.rowid[st] = (uint32_t)r;   // st > 0 because r > r_lo
  }
  __syncthreads();
  // inclusive max-scan over the tile: thread t owns positions [16 t, 16 t + 16)
  uint32_t v[RS_IPT];
  uint32_t run = 0;
#pragma unroll
  for (int u = 0; u < RS_IPT; ++u) {
    run = max(run, rs.rowid[tid * RS_IPT + u]);
    v[u] = run;
  }
  uint32_t incl = run;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl = max(incl, t);
  }
  if (lane == 31) rs.wmax[warp] = incl;
  uint32_t excl = __shfl_up_sync(0xffffffffu, incl, 1);
  if (lane == 0) excl = 0;
  __syncthreads();
  for (int w = 0; w < warp; ++w) excl = max(excl, rs.wmax[w]);
#pragma unroll
  for (int u = 0; u < RS_IPT; ++u) rs.rowid[tid * RS_IPT + u] = max(v[u], excl);
  __syncthreads();
}

// Stable scatter of one supertile.  a.counts = exclusive scan of the histogram kernel's counts.
template <bool TWO, int P1MODE>
__global__ void __launch_bounds__(RS_THREADS, 2) radix_scatter_kernel(const PassArgs a) {
  extern __shared__ __align__(16) uint8_t rs_smem_raw[];
  SortSmem& sm = *reinterpret_cast<SortSmem*>(rs_smem_raw);
  RowSmem& rs = *reinterpret_cast<RowSmem*>(rs_smem_raw + ((sizeof(SortSmem) + 15) & ~(size_t)15));
  int64_t sbeg, send;
  if (!supertile_range(a, sbeg, send)) return;
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int nbins = 1 << a.bits;
  const uint32_t mask = (uint32_t)nbins - 1u;
  for (int d = tid; d < nbins; d += RS_THREADS) sm.running[d] = a.counts[(size_t)d * a.nsuper + blockIdx.x];
  // (the first __syncthreads of the loop body orders these writes before their use)
  for (int64_t base = sbeg; base < send; base += RS_TILE) {
    const int count = (int)min((int64_t)RS_TILE, send - base);
    const int p0 = warp * RS_WARP_ITEMS + lane;
    uint32_t key[RS_IPT];
    uint32_t valid_bits = 0;
#pragma unroll
    for (int r = 0; r < RS_IPT; ++r) {
      const int p = p0 + r * 32;
      const bool ok = p < count;
      key[r] = ok ? __ldcs(a.key_in + base + p) : 0u;
      valid_bits |= (uint32_t)ok << r;
    }
    if constexpr (P1MODE == P1_ROWS) tile_row_ids(rs, a.row_offs, a.m, base, count);
    uint32_t fs[RS_IPT];
    local_sort<TWO, true>(sm, [&](int r) { return (key[r] >> a.shift) & mask; }, valid_bits, count, fs);
    for (int s = tid; s < count; s += RS_THREADS) {
      const uint16_t d = sm.sdig[s];
      if (s == 0 || sm.sdig[s - 1] != d) sm.running[d] -= (uint32_t)s;   // global position of slot s = running[d] + s
    }
    __syncthreads();   // also: every fs[] has been read out of sm.u.a before sm.u.stage is written

    if (a.key_out != nullptr) {
#pragma unroll
      for (int r = 0; r < RS_IPT; ++r)
        if ((valid_bits >> r) & 1u) sm.u.stage[fs[r]] = key[r];
      __syncthreads();
      for (int s = tid; s < count; s += RS_THREADS) a.key_out[sm.running[sm.sdig[s]] + (uint32_t)s] = sm.u.stage[s];
      __syncthreads();
    }
    {
#pragma unroll
      for (int r = 0; r < RS_IPT; ++r) {
        if ((valid_bits >> r) & 1u) {
          const int p = p0 + r * 32;
          uint32_t v;
          if constexpr (P1MODE == P1_ROWS) v = rs.rowid[p];
          else if constexpr (P1MODE == P1_IOTA) v = (uint32_t)(base + p);
          else v = __ldcs(a.p1_in + base + p);
          sm.u.stage[fs[r]] = v;
        }
      }
      __syncthreads();
      for (int s = tid; s < count; s += RS_THREADS) a.p1_out[sm.running[sm.sdig[s]] + (uint32_t)s] = sm.u.stage[s];
      __syncthreads();
    }
    if (a.p2_in != nullptr) {
#pragma unroll
      for (int r = 0; r < RS_IPT; ++r)
        if ((valid_bits >> r) & 1u) sm.u.stage[fs[r]] = __ldcs(a.p2_in + base + p0 + r * 32);
      __syncthreads();
      for (int s = tid; s < count; s += RS_THREADS) a.p2_out[sm.running[sm.sdig[s]] + (uint32_t)s] = sm.u.stage[s];
      __syncthreads();
    }
    for (int s = tid; s < count; s += RS_THREADS) {
      const uint16_t d = sm.sdig[s];
      if (s == count - 1 || sm.sdig[s + 1] != d) sm.running[d] += (uint32_t)s + 1u;
    }
    __syncthreads();
  }
}

// ---- device-wide exclusive scan of uint32 (in place), three kernels ---------------------------
constexpr int SC_THREADS = 512;
constexpr int SC_IPT = 8;
constexpr int SC_CHUNK = SC_THREADS * SC_IPT;

__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total_out) {
  __shared__ uint32_t wsum[SC_THREADS / 32];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  uint32_t incl = v;
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
    if (lane >= o) incl += t;
  }
  if (lane == 31) wsum[warp] = incl;
  __syncthreads();
  uint32_t add = 0, tot = 0;
  for (int w = 0; w < SC_THREADS / 32; ++w) {
    const uint32_t t = wsum[w];
    if (w < warp) add += t;
    tot += t;
  }
  __syncthreads();
  *total_out = tot;
  return add + incl - v;
}

__global__ void __launch_bounds__(SC_THREADS)
scan_reduce_kernel(const uint32_t* __restrict__ data, int64_t n, uint32_t* __restrict__ block_sums) {
  const int64_t base = (int64_t)blockIdx.x * SC_CHUNK;
  uint32_t s = 0;
#pragma unroll
  for (int u = 0; u < SC_IPT; ++u) {
    const int64_t i = base + u * SC_THREADS + threadIdx.x;
    if (i < n) s += data[i];
  }
  uint32_t tot;
  block_excl_scan(s, &tot);
  if (threadIdx.x == 0) block_sums[blockIdx.x] = tot;
}

__global__ void __launch_bounds__(SC_THREADS)
scan_blocksums_kernel(uint32_t* __restrict__ block_sums, int64_t nb) {
  uint32_t carry = 0;
  for (int64_t base = 0; base < nb; base += SC_THREADS) {
    const int64_t i = base + threadIdx.x;
    const uint32_t v = (i < nb) ? block_sums[i] : 0u;
    uint32_t tot;
    const uint32_t ex = block_excl_scan(v, &tot);
    if (i < nb) block_sums[i] = carry + ex;
    carry += tot;
  }
}

__global__ void __launch_bounds__(SC_THREADS)
scan_apply_kernel(uint32_t* __restrict__ data, int64_t n, const uint32_t* __restrict__ block_sums) {
  // thread t owns SC_IPT consecutive elements so that the scan order is the index order
  const int64_t base = (int64_t)blockIdx.x * SC_CHUNK + (int64_t)threadIdx.x * SC_IPT;
  uint32_t v[SC_IPT];
  uint32_t s = 0;
#pragma unroll
  for (int u = 0; u < SC_IPT; ++u) {
    v[u] = (base + u < n) ? data[base + u] : 0u;
    s += v[u];
  }
  uint32_t tot;
  uint32_t run = block_excl_scan(s, &tot) + block_sums[blockIdx.x];
#pragma unroll
  for (int u = 0; u < SC_IPT; ++u) {
    if (base + u < n) data[base + u] = run;
    run += v[u];
  }
}

// Supertiles of pass 2, cut at the bucket boundaries of pass 1.  scanned1[l * nsuper1] is the start of low-digit
// bucket l in the pass-1 output.  One block.  Writes first_super[0 .. nb1] (first supertile of every bucket; entry
// nb1 = number of supertiles, also stored in *nsuper_actual) and super_start[0 .. count].
constexpr int ST_THREADS = 1024;
__global__ void __launch_bounds__(ST_THREADS)
super_table_kernel(const uint32_t* __restrict__ scanned1, uint32_t nsuper1, int nb1, int64_t n_items,
                   int64_t super_len, uint32_t* __restrict__ first_super, int64_t* __restrict__ super_start,
                   uint32_t* __restrict__ nsuper_actual) {
  __shared__ uint32_t wsum[ST_THREADS / 32];
  __shared__ uint32_t carry_s;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  if (tid == 0) carry_s = 0;
  __syncthreads();
  for (int l0 = 0; l0 < nb1; l0 += ST_THREADS) {
    const int l = l0 + tid;
    int64_t beg = 0, end = 0;
    if (l < nb1) {
      beg = scanned1[(size_t)l * nsuper1];
      end = (l + 1 < nb1) ? (int64_t)scanned1[(size_t)(l + 1) * nsuper1] : n_items;
    }
    const uint32_t cnt = (uint32_t)((end - beg + super_len - 1) / super_len);
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    uint32_t add = carry_s;
    for (int w = 0; w < warp; ++w) add += wsum[w];
    const uint32_t first = add + incl - cnt;
    if (l < nb1) {
      first_super[l] = first;
      for (uint32_t j = 0; j < cnt; ++j) super_start[first + j] = beg + (int64_t)j * super_len;
    }
    __syncthreads();
    if (tid == ST_THREADS - 1) carry_s = add + incl;
    __syncthreads();
  }
  if (tid == 0) {
    first_super[nb1] = carry_s;
    super_start[carry_s] = n_items;
    *nsuper_actual = carry_s;
  }
}

// One pass: seg_offs[c] = first output position of digit c = scanned[c * nsuper] (entry nbins * nsuper holds the
// total).  c runs over [0, nseg], nseg <= nbins.
template <typename OutT>
__global__ void __launch_bounds__(256)
offsets_one_pass_kernel(const uint32_t* __restrict__ scanned, uint32_t nsuper, int64_t nseg, OutT* __restrict__ seg_offs) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c <= nseg; c += stride)
    seg_offs[c] = (OutT)scanned[(size_t)c * nsuper];
}

// Two passes: the start of (high digit h, low digit l) in the final order is the scanned pass-2 count of digit h
// at the first supertile of pass-1 bucket l (supertiles never straddle a bucket; the array ends with the total).
template <typename OutT>
__global__ void __launch_bounds__(256)
offsets_two_pass_kernel(const uint32_t* __restrict__ scanned2, uint32_t nsuper2, const uint32_t* __restrict__ first_super,
                        int bits1, int64_t nseg, OutT* __restrict__ seg_offs) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int64_t mask1 = ((int64_t)1 << bits1) - 1;
  for (int64_t c = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; c <= nseg; c += stride)
    seg_offs[c] = (OutT)scanned2[(size_t)(c >> bits1) * nsuper2 + first_super[c & mask1]];
}

// Three or more passes (keys wider than 24 bits): seg_offs[c] = first position whose sorted key is >= c
template <typename OutT>
__global__ void __launch_bounds__(256)
segment_offsets_kernel(const uint32_t* __restrict__ sorted_keys, int64_t n, int64_t nseg,
                       OutT* __restrict__ seg_offs) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i <= n; i += stride) {
    // keys above nseg (never produced by our kernels; a caller-supplied assignment could hold one) must not
    // run the loop past the (nseg + 1)-entry output
    const int64_t lo = (i == 0) ? 0 : (int64_t)sorted_keys[i - 1] + 1;
    const int64_t hi = (i == n) ? nseg : min((int64_t)sorted_keys[i], nseg);
    for (int64_t c = lo; c <= hi; ++c) seg_offs[c] = (OutT)i;
  }
}

int key_bits(int64_t nkeys) {
  int b = 1;
  while (b < 32 && ((int64_t)1 << b) < nkeys) ++b;
  return b;
}

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

int exclusive_scan_u32(bof_ctx* ctx, cudaStream_t s, uint32_t* data, int64_t n, uint32_t* block_sums) {
  const int64_t nb = ceil_div<int64_t>(n, SC_CHUNK);
  scan_reduce_kernel<<<(unsigned)nb, SC_THREADS, 0, s>>>(data, n, block_sums);
  BOF_LAUNCH_CHECK(ctx, "scan_reduce_kernel");
  scan_blocksums_kernel<<<1, SC_THREADS, 0, s>>>(block_sums, nb);
  BOF_LAUNCH_CHECK(ctx, "scan_blocksums_kernel");
  scan_apply_kernel<<<(unsigned)nb, SC_THREADS, 0, s>>>(data, n, block_sums);
  BOF_LAUNCH_CHECK(ctx, "scan_apply_kernel");
  return BOF_OK;
}

// ---- plan: digit widths, supertiles, workspace layout ------------------------------------------------------------

// bof_config.radix_max_bits (test knob): cap of the digit width, e.g. 8 makes a 23-bit key take three passes
int clamp_bits(int cfg_bits) { return cfg_bits <= 0 ? RS_MAX_BITS : std::min(RS_MAX_BITS, cfg_bits); }

constexpr int RS_MAX_PASSES = 8;

struct SortPlan {
  int passes = 0;
  int bits[RS_MAX_PASSES] = {};
  int shift[RS_MAX_PASSES] = {};
  int64_t super_len = RS_TILE;
  uint32_t nsuper[RS_MAX_PASSES] = {};   // supertile count (pass 2 of a two-pass plan: upper bound)
  size_t counts_elems[RS_MAX_PASSES] = {};  // histogram entries incl. the trailing total slot
  // workspace layout (byte offsets from the 256-aligned base)
  size_t off_buf[2][3] = {};   // ping-pong triples (key, p1, p2)
  size_t off_counts = 0, off_block_sums = 0, off_super_start = 0, off_first_super = 0, off_nsuper = 0;
  size_t bytes = 0;
};

SortPlan make_plan(int64_t n_items, int kbits, bool has_p2, int num_sms, int maxb) {
  SortPlan p;
  p.passes = std::max(1, ceil_div(kbits, maxb));
  int left = kbits, sh = 0;
  for (int i = 0; i < p.passes; ++i) {
    const int b = ceil_div(left, p.passes - i);   // widest digits first: 23 -> 12 + 11
    p.bits[i] = b;
    p.shift[i] = sh;
    sh += b;
    left -= b;
  }
  // supertiles: enough of them to fill the machine several times over, at most RS_MAX_SUB tiles each
  const int64_t n1 = std::max<int64_t>(n_items, 1);
  const int64_t want = (int64_t)num_sms * 2 * 4;
  int64_t sub = std::min<int64_t>(RS_MAX_SUB, std::max<int64_t>(1, n1 / ((int64_t)RS_TILE * want)));
  p.super_len = sub * RS_TILE;
  const uint32_t uniform = (uint32_t)ceil_div<int64_t>(n1, p.super_len);
  size_t max_counts = 0;
  for (int i = 0; i < p.passes; ++i) {
    p.nsuper[i] = uniform;
    if (p.passes == 2 && i == 1) p.nsuper[i] = uniform + (1u << p.bits[0]);   // one partial supertile per bucket
    p.counts_elems[i] = ((size_t)1 << p.bits[i]) * p.nsuper[i] + 1;
    max_counts = std::max(max_counts, p.counts_elems[i]);
  }
  const size_t arr = align_up((size_t)n1 * 4);
  size_t off = 0;
  const int nbuf = p.passes >= 3 ? 2 : (p.passes == 2 ? 1 : 0);
  for (int b = 0; b < nbuf; ++b)
    for (int q = 0; q < 3; ++q) {
      p.off_buf[b][q] = off;
      if (q < 2 || has_p2) off += arr;
    }
  // passes >= 3 park the sorted keys of the last pass in the key array of the buffer that pass does not read
  p.off_counts = off;
  // a two-pass plan keeps both histograms (the offsets gather needs the second, the table kernel the first)
  const size_t counts_total = p.passes == 2 ? align_up(p.counts_elems[0] * 4) + align_up(p.counts_elems[1] * 4)
                                            : align_up(max_counts * 4);
  off += counts_total;
  p.off_block_sums = off;
  off += align_up((size_t)ceil_div<int64_t>((int64_t)max_counts, SC_CHUNK) * 4);
  p.off_super_start = off;
  off += align_up(((size_t)p.nsuper[p.passes == 2 ? 1 : 0] + 2) * 8);
  p.off_first_super = off;
  off += align_up((((size_t)1 << p.bits[0]) + 1) * 4);
  p.off_nsuper = off;
  off += 256;
  p.bytes = off + 256;
  return p;
}

template <typename K>
int set_smem_attr(bof_ctx* ctx, K kern, size_t bytes, PerDeviceOnce& once) {
  if (once.need(ctx->device)) {
    BOF_CUDA(ctx, cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
    once.done(ctx->device);
  }
  return BOF_OK;
}

constexpr size_t kSortSmemBytes = (sizeof(SortSmem) + 15) & ~(size_t)15;
constexpr size_t kRowSmemBytes = kSortSmemBytes + sizeof(RowSmem);

int launch_hist(bof_ctx* ctx, cudaStream_t s, const PassArgs& a) {
  static PerDeviceOnce once[2];
  if (a.bits > 6) {
    if (int rc = set_smem_attr(ctx, radix_hist_kernel<true>, kSortSmemBytes, once[1])) return rc;
    radix_hist_kernel<true><<<a.nsuper, RS_THREADS, kSortSmemBytes, s>>>(a);
  } else {
    if (int rc = set_smem_attr(ctx, radix_hist_kernel<false>, kSortSmemBytes, once[0])) return rc;
    radix_hist_kernel<false><<<a.nsuper, RS_THREADS, kSortSmemBytes, s>>>(a);
  }
  BOF_LAUNCH_CHECK(ctx, "radix_hist_kernel");
  return BOF_OK;
}

template <bool TWO, int P1MODE>
int launch_scatter_t(bof_ctx* ctx, cudaStream_t s, const PassArgs& a) {
  static PerDeviceOnce once;
  const size_t smem = P1MODE == P1_ROWS ? kRowSmemBytes : kSortSmemBytes;
  if (int rc = set_smem_attr(ctx, radix_scatter_kernel<TWO, P1MODE>, smem, once)) return rc;
  radix_scatter_kernel<TWO, P1MODE><<<a.nsuper, RS_THREADS, smem, s>>>(a);
  BOF_LAUNCH_CHECK(ctx, "radix_scatter_kernel");
  return BOF_OK;
}

int launch_scatter(bof_ctx* ctx, cudaStream_t s, const PassArgs& a, int p1mode) {
  const bool two = a.bits > 6;
  switch (p1mode) {
    case P1_ROWS: return two ? launch_scatter_t<true, P1_ROWS>(ctx, s, a) : launch_scatter_t<false, P1_ROWS>(ctx, s, a);
    case P1_IOTA: return two ? launch_scatter_t<true, P1_IOTA>(ctx, s, a) : launch_scatter_t<false, P1_IOTA>(ctx, s, a);
    default: return two ? launch_scatter_t<true, P1_ARRAY>(ctx, s, a) : launch_scatter_t<false, P1_ARRAY>(ctx, s, a);
  }
}

// Stable sort of (key, payload 1[, payload 2]) by a key of `kbits` bits + the nseg + 1 segment offsets of the
// sorted order (seg_offs[c] = number of items with key < c).  Payload 1 of the input is an array (p1_in), the CSR
// row of the item (row_offs, m) or its position (neither given).
int radix_sort_segments(bof_ctx* ctx, cudaStream_t s, int64_t n_items, int kbits, const uint32_t* key_in,
                        const uint32_t* p1_in, const int64_t* row_offs, int64_t m, const uint32_t* p2_in,
                        uint32_t* p1_out, uint32_t* p2_out, int64_t nseg, int64_t* seg_offs, void* ws, size_t ws_bytes) {
  const SortPlan pl = make_plan(n_items, kbits, p2_in != nullptr, ctx->num_sms, clamp_bits(ctx->cfg.radix_max_bits));
  BOF_REQUIRE(ctx, pl.passes <= RS_MAX_PASSES, "radix sort: too many passes");
  BOF_REQUIRE(ctx, ws != nullptr && ws_bytes >= pl.bytes, "radix sort: workspace too small");
  BOF_REQUIRE(ctx, n_items > 0 && n_items < (1ll << 32) - RS_TILE, "radix sort: item count must be below 2^32");
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  auto buf = [&](int b, int q) { return reinterpret_cast<uint32_t*>(base + pl.off_buf[b][q]); };
  uint32_t* counts[2] = {reinterpret_cast<uint32_t*>(base + pl.off_counts),
                         reinterpret_cast<uint32_t*>(base + pl.off_counts + (pl.passes == 2 ? align_up(pl.counts_elems[0] * 4) : 0))};
  uint32_t* block_sums = reinterpret_cast<uint32_t*>(base + pl.off_block_sums);
  int64_t* super_start = reinterpret_cast<int64_t*>(base + pl.off_super_start);
  uint32_t* first_super = reinterpret_cast<uint32_t*>(base + pl.off_first_super);
  uint32_t* nsuper_actual = reinterpret_cast<uint32_t*>(base + pl.off_nsuper);
  const int p1mode0 = p1_in ? P1_ARRAY : (row_offs ? P1_ROWS : P1_IOTA);
  const unsigned off_grid = (unsigned)std::min<int64_t>(ceil_div<int64_t>(nseg + 1, 256), (int64_t)ctx->num_sms * 32);

  const uint32_t* kin = key_in;
  const uint32_t* q1 = p1_in;
  const uint32_t* q2 = p2_in;
  uint32_t* last_keys = nullptr;
  for (int p = 0; p < pl.passes; ++p) {
    const bool last = p == pl.passes - 1;
    const bool aligned = pl.passes == 2 && p == 1;
    uint32_t* cnt = counts[pl.passes == 2 ? p : 0];
    PassArgs a{};
    a.key_in = kin; a.p1_in = q1; a.p2_in = q2;
    a.row_offs = row_offs; a.m = m; a.n = n_items;
    a.shift = pl.shift[p]; a.bits = pl.bits[p];
    a.counts = cnt; a.nsuper = pl.nsuper[p]; a.super_len = pl.super_len;
    if (aligned) {
      super_table_kernel<<<1, ST_THREADS, 0, s>>>(counts[0], pl.nsuper[0], 1 << pl.bits[0], n_items, pl.super_len,
                                                  first_super, super_start, nsuper_actual);
      BOF_LAUNCH_CHECK(ctx, "super_table_kernel");
      a.super_start = super_start;
      a.nsuper_actual = nsuper_actual;
      // supertiles past the actual count never run: their histogram entries must read as zero
      BOF_CUDA(ctx, cudaMemsetAsync(cnt, 0, pl.counts_elems[p] * 4, s));
    } else {
      BOF_CUDA(ctx, cudaMemsetAsync(cnt + pl.counts_elems[p] - 1, 0, 4, s));   // the trailing total slot
    }
    if (last) {
      a.p1_out = p1_out; a.p2_out = p2_out;
      if (pl.passes >= 3) {
        last_keys = buf(p % 2, 0);   // this pass reads the other buffer
        a.key_out = last_keys;
      }
    } else {
      a.key_out = buf(p % 2, 0); a.p1_out = buf(p % 2, 1); a.p2_out = p2_in ? buf(p % 2, 2) : nullptr;
    }
    if (int rc = launch_hist(ctx, s, a)) return rc;
    if (int rc = exclusive_scan_u32(ctx, s, cnt, (int64_t)pl.counts_elems[p], block_sums)) return rc;
    if (int rc = launch_scatter(ctx, s, a, p == 0 ? p1mode0 : P1_ARRAY)) return rc;
    kin = a.key_out; q1 = a.p1_out; q2 = a.p2_out;
  }
  if (pl.passes == 1) {
    offsets_one_pass_kernel<int64_t><<<off_grid, 256, 0, s>>>(counts[0], pl.nsuper[0], nseg, seg_offs);
    BOF_LAUNCH_CHECK(ctx, "offsets_one_pass_kernel");
  } else if (pl.passes == 2) {
    offsets_two_pass_kernel<int64_t><<<off_grid, 256, 0, s>>>(counts[1], pl.nsuper[1], first_super, pl.bits[0], nseg, seg_offs);
    BOF_LAUNCH_CHECK(ctx, "offsets_two_pass_kernel");
  } else {
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div<int64_t>(n_items + 1, 256), (int64_t)ctx->num_sms * 32);
    segment_offsets_kernel<int64_t><<<grid, 256, 0, s>>>(last_keys, n_items, nseg, seg_offs);
    BOF_LAUNCH_CHECK(ctx, "segment_offsets_kernel");
  }
  return BOF_OK;
}

size_t radix_sort_workspace_bytes(int64_t n_items, int kbits, bool has_p2) {
  // The plan depends on the SM count through the supertile length (more SMs -> shorter supertiles -> more
  // counters) and on the digit cap (the workspace query has no context): size it for an SM count no device
  // exceeds and take the largest layout over the digit caps.
  size_t best = 0;
  for (int maxb = 1; maxb <= RS_MAX_BITS; ++maxb) {
    if (ceil_div(kbits, maxb) > RS_MAX_PASSES) continue;
    best = std::max(best, make_plan(n_items, kbits, has_p2, 256, maxb).bytes);
  }
  return best;
}

// ---- k-means ------------------------------------------------------------------------------------

__global__ void __launch_bounds__(256)
row_sqnorm_kernel(int64_t rows, int64_t dim, const float* __restrict__ X, int64_t ldx,
                  float* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= rows) return;
  const float* x = X + row * ldx;
  float acc = 0.f;
  for (int64_t j = lane; j < dim; j += 32) {
    const float v = __ldg(x + j);
    acc = fmaf(v, v, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (lane == 0) out[row] = acc;
}

// Centroid partial sums, load-balanced and deterministic.  Clusters are cut into segments of KM_SEG
// points of the id list sorted by (cluster, p); one block sums one segment (one thread per
// dimension, sequential fp32 adds in ascending p, 8 row loads in flight), a second kernel adds the
// segments of each cluster in ascending order.  A 15x cluster-size imbalance (seen on cfg-5) no
// longer serialises behind one block.
constexpr int KM_SEG = 256;

// seg_base[c] = number of segments of clusters < c (exclusive scan of ceil(count_c / KM_SEG)); one block
__global__ void __launch_bounds__(1024)
kmeans_plan_kernel(int64_t ncenters, const int64_t* __restrict__ seg_offs, int32_t* __restrict__ seg_base) {
  __shared__ int32_t wsum[32];
  __shared__ int32_t carry_s;
  if (threadIdx.x == 0) carry_s = 0;
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int64_t c0 = 0; c0 < ncenters; c0 += blockDim.x) {
    const int64_t c = c0 + threadIdx.x;
    const int32_t v = (c < ncenters) ? (int32_t)((seg_offs[c + 1] - seg_offs[c] + KM_SEG - 1) / KM_SEG) : 0;
    int32_t incl = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int32_t t = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += t;
    }
    if (lane == 31) wsum[warp] = incl;
    __syncthreads();
    int32_t add = carry_s;
    for (int w = 0; w < warp; ++w) add += wsum[w];
    if (c < ncenters) seg_base[c] = add + incl - v;
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) carry_s = add + incl;
    __syncthreads();
  }
  if (threadIdx.x == 0) seg_base[ncenters] = carry_s;
}

__global__ void __launch_bounds__(256)
kmeans_partial_kernel(int64_t ncenters, int64_t dim, const float* __restrict__ points,
                      const uint32_t* __restrict__ sorted_ids, const int64_t* __restrict__ seg_offs,
                      const int32_t* __restrict__ seg_base, float* __restrict__ partial) {
  __shared__ int64_t s_beg, s_end;
  const int32_t b = blockIdx.x;
  if (b >= seg_base[ncenters]) return;
  if (threadIdx.x == 0) {
    int64_t lo = 0, hi = ncenters;  // last c with seg_base[c] <= b
    while (hi - lo > 1) {
      const int64_t mid = (lo + hi) >> 1;
      if (seg_base[mid] <= b) lo = mid; else hi = mid;
    }
    const int64_t beg = seg_offs[lo] + (int64_t)(b - seg_base[lo]) * KM_SEG;
    s_beg = beg;
    s_end = min(seg_offs[lo + 1], beg + KM_SEG);
  }
  __syncthreads();
  const int64_t beg = s_beg, end = s_end;
  for (int64_t j = threadIdx.x; j < dim; j += blockDim.x) {
    float acc = 0.f;
    int64_t i = beg;
    for (; i + 8 <= end; i += 8) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; ++u) v[u] = __ldcs(points + (int64_t)sorted_ids[i + u] * dim + j);
#pragma unroll
      for (int u = 0; u < 8; ++u) acc += v[u];
    }
    for (; i < end; ++i) acc += __ldcs(points + (int64_t)sorted_ids[i] * dim + j);
    partial[(int64_t)b * dim + j] = acc;
  }
}

__global__ void __launch_bounds__(256)
kmeans_combine_kernel(int64_t dim, const float* __restrict__ partial, const int64_t* __restrict__ seg_offs,
                      const int32_t* __restrict__ seg_base, float* __restrict__ sums, float* __restrict__ counts) {
  const int64_t c = blockIdx.x;
  const int32_t s0 = seg_base[c], s1 = seg_base[c + 1];
  for (int64_t j = threadIdx.x; j < dim; j += blockDim.x) {
    float acc = 0.f;
    for (int32_t sgm = s0; sgm < s1; ++sgm) acc += partial[(int64_t)sgm * dim + j];
    sums[c * dim + j] = acc;
  }
  if (threadIdx.x == 0) counts[c] = (float)(seg_offs[c + 1] - seg_offs[c]);
}

__global__ void __launch_bounds__(256)
kmeans_finalize_kernel(int64_t dim, const float* __restrict__ sums, const float* __restrict__ counts,
                       float* __restrict__ centers, float* __restrict__ c_l2sq) {
  __shared__ float red[8];
  const int64_t c = blockIdx.x;
  const float cnt = counts[c];
  const float inv = cnt > 0.f ? 1.f / cnt : 0.f;
  float acc = 0.f;
  for (int64_t j = threadIdx.x; j < dim; j += blockDim.x) {
    const float v = cnt > 0.f ? sums[c * dim + j] * inv : 0.f;  // empty cluster => zero vector
    centers[c * dim + j] = v;
    acc = fmaf(v, v, acc);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    c_l2sq[c] = t;
  }
}

}  // namespace

size_t csr2csc_workspace_bytes(int64_t m, int64_t n, int64_t nnz) {
  (void)m;
  return radix_sort_workspace_bytes(std::max<int64_t>(nnz, 1), key_bits(n), true) + 256;
}

int launch_csr2csc(bof_ctx* ctx, cudaStream_t s, int64_t m, int64_t n, int64_t nnz,
                   const int64_t* offs, const int32_t* idx, const float* vals, int64_t* offs_t,
                   int32_t* idx_t, float* vals_t, void* ws, size_t ws_bytes) {
  BOF_REQUIRE(ctx, nnz >= 0 && nnz < (1ll << 32) - RS_TILE, "csr2csc: nnz must be below 2^32");
  BOF_REQUIRE(ctx, m < (1ll << 31) && n < (1ll << 31), "csr2csc: dimensions must be below 2^31");
  if (nnz == 0) {
    BOF_CUDA(ctx, cudaMemsetAsync(offs_t, 0, (size_t)(n + 1) * sizeof(int64_t), s));
    return BOF_OK;
  }
  BOF_REQUIRE(ctx, ws != nullptr && ws_bytes >= csr2csc_workspace_bytes(m, n, nnz), "csr2csc: workspace too small");
  // keys = column indices, payload 1 = source row (derived from the offsets inside the first pass), payload 2 =
  // the value bits; the last pass lands in the caller's arrays and the offsets come out of the scanned histogram
  return radix_sort_segments(ctx, s, nnz, key_bits(n), reinterpret_cast<const uint32_t*>(idx), nullptr, offs, m,
                             reinterpret_cast<const uint32_t*>(vals), reinterpret_cast<uint32_t*>(idx_t),
                             reinterpret_cast<uint32_t*>(vals_t), n, offs_t, ws, ws_bytes);
}

int launch_row_sqnorm(bof_ctx* ctx, cudaStream_t s, int64_t rows, int64_t dim, const float* X,
                      int64_t ldx, float* out) {
  if (rows == 0) return BOF_OK;
  row_sqnorm_kernel<<<(unsigned)ceil_div<int64_t>(rows, 8), 256, 0, s>>>(rows, dim, X, ldx, out);
  BOF_LAUNCH_CHECK(ctx, "row_sqnorm_kernel");
  return BOF_OK;
}

static int64_t km_max_segments(int64_t npoints, int64_t ncenters) { return npoints / KM_SEG + ncenters + 1; }

size_t kmeans_reduce_workspace_bytes(int64_t npoints, int64_t ncenters, int64_t dim) {
  const size_t arr = align_up((size_t)std::max<int64_t>(npoints, 1) * 4);
  return arr + radix_sort_workspace_bytes(std::max<int64_t>(npoints, 1), key_bits(ncenters), false) +
         align_up((size_t)(ncenters + 1) * 8) + align_up((size_t)(ncenters + 1) * 4) +
         align_up((size_t)km_max_segments(npoints, ncenters) * (size_t)std::max<int64_t>(dim, 1) * 4) + 512;
}

int launch_kmeans_reduce_ws(bof_ctx* ctx, cudaStream_t s, int64_t npoints, int64_t ncenters,
                            int64_t dim, const float* points, const int32_t* assign, float* sums,
                            float* counts_out, void* ws, size_t ws_bytes) {
  BOF_REQUIRE(ctx, npoints < (1ll << 31) && ncenters < (1ll << 31), "kmeans_reduce: extents must be below 2^31");
  BOF_REQUIRE(ctx, ncenters > 0 && dim > 0, "kmeans_reduce: empty problem");
  BOF_REQUIRE(ctx, ws != nullptr && ws_bytes >= kmeans_reduce_workspace_bytes(npoints, ncenters, dim),
              "kmeans_reduce: workspace too small");
  const size_t arr = align_up((size_t)std::max<int64_t>(npoints, 1) * 4);
  const size_t sort_bytes = radix_sort_workspace_bytes(std::max<int64_t>(npoints, 1), key_bits(ncenters), false);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  uint32_t* sorted_ids = reinterpret_cast<uint32_t*>(base);
  uint8_t* sort_ws = base + arr;
  int64_t* seg = reinterpret_cast<int64_t*>(sort_ws + align_up(sort_bytes));
  int32_t* seg_base = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(seg) + align_up((size_t)(ncenters + 1) * 8));
  float* partial = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(seg_base) + align_up((size_t)(ncenters + 1) * 4));

  if (npoints > 0) {
    // point ids grouped by cluster, ascending id inside a cluster (stable): keys = assignment, payload = position
    int rc = radix_sort_segments(ctx, s, npoints, key_bits(ncenters), reinterpret_cast<const uint32_t*>(assign), nullptr,
                                 nullptr, 0, nullptr, sorted_ids, nullptr, ncenters, seg, sort_ws, sort_bytes);
    if (rc) return rc;
  } else {
    BOF_CUDA(ctx, cudaMemsetAsync(seg, 0, (size_t)(ncenters + 1) * 8, s));
  }
  kmeans_plan_kernel<<<1, 1024, 0, s>>>(ncenters, seg, seg_base);
  BOF_LAUNCH_CHECK(ctx, "kmeans_plan_kernel");
  kmeans_partial_kernel<<<(unsigned)km_max_segments(npoints, ncenters), 256, 0, s>>>(ncenters, dim, points, sorted_ids, seg,
                                                                                     seg_base, partial);
  BOF_LAUNCH_CHECK(ctx, "kmeans_partial_kernel");
  kmeans_combine_kernel<<<(unsigned)ncenters, 256, 0, s>>>(dim, partial, seg, seg_base, sums, counts_out);
  BOF_LAUNCH_CHECK(ctx, "kmeans_combine_kernel");
  return BOF_OK;
}

int launch_kmeans_finalize(bof_ctx* ctx, cudaStream_t s, int64_t ncenters, int64_t dim,
                           const float* sums, const float* counts, float* centers, float* c_l2sq) {
  if (ncenters == 0) return BOF_OK;
  kmeans_finalize_kernel<<<(unsigned)ncenters, 256, 0, s>>>(dim, sums, counts, centers, c_l2sq);
  BOF_LAUNCH_CHECK(ctx, "kmeans_finalize_kernel");
  return BOF_OK;
}

}  // namespace bof
