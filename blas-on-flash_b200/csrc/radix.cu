// K6/K7 (stable CSR -> CSC) and K10/K11 (k-means centroid reduce, row norms) for sm_100a.
//
// Reference bodies replaced:
//   mkl_scsrcsc + index rebase      include/tasks/csrcsc_task.h:68-80
//   row-block-ordered segment merge include/tasks/csrcsc_task.h:143-162
//   bucket + cblas_saxpy loop       drivers/in_mem_kmeans.cpp:105-125
//   cblas_sdot row norms            drivers/in_mem_kmeans.cpp:75-78,179-182
//
// The transpose is the unique *stable* counting sort of the nonzeros by column.  It is built from
// least-significant-digit radix passes (8-bit digits), each pass = per-tile digit histogram ->
// device-wide exclusive scan -> ranked scatter.  Ranks come from warp match/ballot on warp-private
// counters, so there is no atomic anywhere and the result is bit-reproducible: within a column the
// entries keep ascending source row, duplicates keep storage order -- exactly what mkl_scsrcsc on
// row blocks followed by the reference's in-order merge produces.
//
// All of it is HBM-bound byte shuffling; tiles are staged through shared memory so that the
// scattered runs leave the SM as contiguous segments.  The source row of every nonzero is derived from the
// CSR offsets inside the first pass (no materialised row array).
//
// Measured and rejected in round 2 (profiles/r02/csrcsc_wide_digits.md): 12-bit digits (two passes for 2^23
// columns) with supertile histograms and a two-level 6+6-bit ballot ranking.  Bit-exact, but the ranking costs
// 220-290 warp instructions per 32 keys (issue-bound) and the 4096 x #CTA open write frontiers of 8-byte runs
// overflow L2, so DRAM saw 3.6x write and 5.8x read amplification: 98 ms against 31 ms for this 8-bit version.
#include "common.cuh"

#include <algorithm>

namespace bof {
namespace {

constexpr int RS_THREADS = 256;
constexpr int RS_WARPS = RS_THREADS / 32;
constexpr int RS_IPT = 16;                       // items per thread
constexpr int RS_TILE = RS_THREADS * RS_IPT;     // 4096 items per tile
constexpr int RS_BINS = 256;
constexpr int RS_CNT_STRIDE = RS_BINS + 1;       // +1: sentinel bin for out-of-range items

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

// Lanes of the warp holding the same 9-bit digit (8 data bits + the out-of-range sentinel bit).
// Nine ballots and a few logic ops: MATCH.ANY measured ~10x slower than this on sm_100
// (profiles/r01: radix_hist_kernel was issue-bound at 96% SM throughput with __match_any_sync).
__device__ __forceinline__ unsigned digit_peers(uint32_t d) {
  unsigned peers = 0xffffffffu;
#pragma unroll
  for (int b = 0; b < 9; ++b) {
    const bool bit = (d >> b) & 1u;
    const unsigned bal = __ballot_sync(0xffffffffu, bit);
    peers &= bit ? bal : ~bal;
  }
  return peers;
}

// counts[digit * num_tiles + tile] = number of keys of the tile whose digit is `digit`.
// Counting needs no ranks, so every thread keeps private one-byte counters in shared memory
// (cnt8[warp][digit][lane], at most RS_TILE / 128 = 32 increments each): three instructions per key
// and no cross-lane traffic, against ~45 for the ballot-based peer search the scatter kernel needs.
// One block of 128 threads per scatter tile.
constexpr int RH_THREADS = 128;
constexpr int RH_WARPS = RH_THREADS / 32;
constexpr int RH_IPT = RS_TILE / RH_THREADS;  // 32 keys per thread and tile
// A block can own a SUPERTILE of RS_SUPER consecutive tiles (histogram: one set of counters for all of them; scatter:
// a running output cursor per digit carried from tile to tile).  Measured at cfg-4 with RS_SUPER = 4
// (profiles/r02/t15_csrcsc_supertile4_per_launch.txt): histogram 1.75 -> 1.13 ms and the scans vanish, but the scatter
// passes go from 8.9 / 7.3 / 6.9 ms to 12.0 / 13.3 / 9.3 ms with 40-60 % more DRAM traffic -- adjacent output regions
// are then written by ONE block minutes of GPU time apart instead of by neighbouring blocks at the same time, and L2
// evicts the partially written sectors in between.  So: one tile per block.
constexpr int RS_SUPER = 1;
static_assert(RH_IPT * RS_SUPER <= 255, "one-byte counters");

__global__ void __launch_bounds__(RH_THREADS)
radix_hist_kernel(const uint32_t* __restrict__ keys, int64_t n, int shift, uint32_t mask,
                  uint32_t* __restrict__ counts, unsigned num_tiles) {
  __shared__ __align__(16) uint8_t cnt8[RH_WARPS][RS_BINS][32];  // 32 KiB
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  uint4* z = reinterpret_cast<uint4*>(&cnt8[0][0][0]);
  for (int i = threadIdx.x; i < (int)(sizeof(cnt8) / 16); i += RH_THREADS) z[i] = make_uint4(0, 0, 0, 0);
  __syncthreads();
  for (int t = 0; t < RS_SUPER; ++t) {
    const int64_t tile_base = ((int64_t)blockIdx.x * RS_SUPER + t) * RS_TILE;
    if (tile_base >= n) break;
    const int64_t warp_base = tile_base + (int64_t)warp * 32 * RH_IPT;
    uint32_t k[8];
    for (int r0 = 0; r0 < RH_IPT; r0 += 8) {
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t i = warp_base + (r0 + u) * 32 + lane;
        k[u] = (i < n) ? __ldcs(keys + i) : 0xffffffffu;
      }
#pragma unroll
      for (int u = 0; u < 8; ++u) {
        const int64_t i = warp_base + (r0 + u) * 32 + lane;
        if (i < n) cnt8[warp][(k[u] >> shift) & mask][lane]++;
      }
    }
  }
  __syncthreads();
  // digit totals: 4 warps x 32 one-byte counters = 32 words per digit, summed with dp4a
  for (int d = threadIdx.x; d < RS_BINS; d += RH_THREADS) {
    uint32_t t = 0;
#pragma unroll
    for (int w = 0; w < RH_WARPS; ++w) {
      const uint32_t* row = reinterpret_cast<const uint32_t*>(&cnt8[w][d][0]);
#pragma unroll
      for (int q = 0; q < 8; ++q) t = __dp4a(row[q], 0x01010101u, t);
    }
    counts[(size_t)d * num_tiles + blockIdx.x] = t;
  }
}

struct RowSmem {
  uint32_t rowid[RS_TILE];
  uint32_t wmax[RS_WARPS];
  long long r_lo, r_hi;
};

// Last row r with offs[r] <= g (the row that holds nonzero g; offs[m] > g).  Called by one whole warp; `guess`
// is where a matrix with uniform row lengths would have it: one coalesced probe of 32 consecutive offsets
// brackets g there, otherwise a 32-ary search (5 dependent loads for 2^23 rows) takes over.
__device__ __forceinline__ int64_t warp_find_row(const int64_t* __restrict__ offs, int64_t m, int64_t g, int64_t guess) {
  const int lane = threadIdx.x & 31;
  int64_t lo, hi;
  {
    const int64_t w0 = max((int64_t)0, min(guess - 8, m - 31));
    const int64_t r = w0 + lane;
    const bool le = r <= m && offs[min(r, m)] <= g;
    const unsigned b = __ballot_sync(0xffffffffu, le);
    if (b != 0u && b != 0xffffffffu) return w0 + __popc(b) - 1;
    if (b == 0u) { lo = 0; hi = w0; } else { lo = w0 + 31; hi = m; }
  }
  while (hi - lo > 1) {
    const int64_t step = (hi - lo + 31) / 32;
    const int64_t p = min(lo + (int64_t)(lane + 1) * step, hi);
    const bool le = p < hi && offs[p] <= g;
    const int cnt = __popc(__ballot_sync(0xffffffffu, le));
    const int64_t nlo = lo + (int64_t)cnt * step;
    if (cnt < 32) hi = min(hi, lo + (int64_t)(cnt + 1) * step);
    lo = nlo;
  }
  return lo;
}

// rowid[p] = CSR row of nonzero tile_base + p for p < count (offsets may be un-rebased: offs[0] != 0)
__device__ __forceinline__ void tile_row_ids(RowSmem& rs, const int64_t* __restrict__ offs, int64_t m, int64_t nnz,
                                             int64_t tile_base, int count) {
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const int64_t off0 = offs[0];
  if (warp < 2) {
    const int64_t rel = warp == 0 ? tile_base : tile_base + count - 1;
    const int64_t guess = (int64_t)((double)rel * (double)m / (double)nnz);
    const int64_t r = warp_find_row(offs, m, off0 + rel, guess);
    if (lane == 0) { if (warp == 0) rs.r_lo = r; else rs.r_hi = r; }
  }
  for (int p = tid; p < RS_TILE; p += RS_THREADS) rs.rowid[p] = 0;
  __syncthreads();
  const int64_t r_lo = rs.r_lo, r_hi = rs.r_hi;
  if (tid == 0) rs.rowid[0] = (uint32_t)r_lo;
  for (int64_t r = r_lo + 1 + tid; r <= r_hi; r += RS_THREADS) {
    const int64_t st = offs[r] - off0 - tile_base;   // > 0 because r > r_lo
    if (offs[r + 1] > offs[r] && st < count) rs.rowid[st] = (uint32_t)r;
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

struct ScatterSmem {
  uint32_t stage[RS_TILE];
  uint16_t sdig[RS_TILE];
  uint32_t cnt[RS_WARPS][RS_CNT_STRIDE];
  uint32_t digit_off[RS_BINS + 1];   // start of each digit inside the block-sorted tile
  uint32_t gbase[RS_BINS];           // global start of (digit, this tile) minus digit_off:
                                     // global position of tile slot s = gbase[digit(s)] + s
  uint32_t run[RS_BINS];             // running global cursor of every digit across the tiles of the supertile
};

// Stable scatter of one tile.  offsets = exclusive scan of the histogram kernel's counts.
// Moves the key (keys_out may be null) and up to two 32-bit payloads.  ROWS: payload 1 of the input is the CSR row
// of the item, derived from `row_offs` on the fly instead of read from p1_in.
template <bool ROWS>
__global__ void __launch_bounds__(RS_THREADS, 3)
radix_scatter_kernel(const uint32_t* __restrict__ keys_in, uint32_t* __restrict__ keys_out,
                     const uint32_t* __restrict__ p1_in, uint32_t* __restrict__ p1_out,
                     const uint32_t* __restrict__ p2_in, uint32_t* __restrict__ p2_out, int64_t n,
                     int shift, uint32_t mask, const uint32_t* __restrict__ offsets, unsigned num_tiles,
                     const int64_t* __restrict__ row_offs, int64_t m) {
  extern __shared__ __align__(16) uint8_t rs_smem_raw[];
  ScatterSmem& sm = *reinterpret_cast<ScatterSmem*>(rs_smem_raw);
  RowSmem& rs = *reinterpret_cast<RowSmem*>(rs_smem_raw + ((sizeof(ScatterSmem) + 15) & ~(size_t)15));
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < RS_BINS) sm.run[threadIdx.x] = offsets[(size_t)threadIdx.x * num_tiles + blockIdx.x];
  for (int tile = 0; tile < RS_SUPER; ++tile) {
  const int64_t tile_base = ((int64_t)blockIdx.x * RS_SUPER + tile) * RS_TILE;
  if (tile_base >= n) break;
  __syncthreads();   // the previous tile's staging buffer, counters and cursors are no longer in use
  for (int i = threadIdx.x; i < RS_WARPS * RS_CNT_STRIDE; i += RS_THREADS) (&sm.cnt[0][0])[i] = 0;
  __syncthreads();

  const int64_t warp_base = tile_base + (int64_t)warp * 32 * RS_IPT;
  const int tile_count = (int)min((int64_t)RS_TILE, n - tile_base);

  // per item: the key and its rank (later: its slot in the block-sorted tile); digits are recomputed
  uint32_t key[RS_IPT];
  uint32_t slot[RS_IPT];
  auto digit_of = [&](int r) -> uint32_t {
    return (warp_base + r * 32 + lane < n) ? ((key[r] >> shift) & mask) : 256u;
  };
#pragma unroll
  for (int r = 0; r < RS_IPT; ++r) {
    const int64_t i = warp_base + r * 32 + lane;
    key[r] = (i < n) ? __ldcs(keys_in + i) : 0u;
  }
  if constexpr (ROWS) tile_row_ids(rs, row_offs, m, n, tile_base, tile_count);
#pragma unroll
  for (int r = 0; r < RS_IPT; ++r) {
    const uint32_t d = digit_of(r);
    const unsigned peers = digit_peers(d);
    const uint32_t old = sm.cnt[warp][d];
    __syncwarp();
    if ((peers & lanemask_lt()) == 0) sm.cnt[warp][d] = old + __popc(peers);
    __syncwarp();
    slot[r] = old + __popc(peers & lanemask_lt());
  }
  __syncthreads();

  // per digit: exclusive prefix over warps (in place), digit totals -> exclusive scan over digits
  uint32_t total = 0;
  if (threadIdx.x < RS_BINS) {
#pragma unroll
    for (int w = 0; w < RS_WARPS; ++w) {
      const uint32_t t = sm.cnt[w][threadIdx.x];
      sm.cnt[w][threadIdx.x] = total;
      total += t;
    }
  }
  // block-wide exclusive scan of `total` (one digit per thread)
  {
    uint32_t incl = total;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o);
      if (lane >= o) incl += v;
    }
    __shared__ uint32_t warp_tot[RS_WARPS];
    if (lane == 31) warp_tot[warp] = incl;
    __syncthreads();
    uint32_t add = 0;
    for (int w = 0; w < warp; ++w) add += warp_tot[w];
    if (threadIdx.x < RS_BINS) sm.digit_off[threadIdx.x] = add + incl - total;
    if (threadIdx.x == RS_BINS - 1) sm.digit_off[RS_BINS] = add + incl;
  }
  __syncthreads();

  constexpr uint32_t NO_SLOT = 0xffffffffu;
#pragma unroll
  for (int r = 0; r < RS_IPT; ++r) {
    const uint32_t d = digit_of(r);
    slot[r] = (d < 256u) ? sm.digit_off[d] + sm.cnt[warp][d] + slot[r] : NO_SLOT;
  }
  if (threadIdx.x < RS_BINS) {
    sm.gbase[threadIdx.x] = sm.run[threadIdx.x] - sm.digit_off[threadIdx.x];  // wraps mod 2^32, undone by + s
    sm.run[threadIdx.x] += sm.digit_off[threadIdx.x + 1] - sm.digit_off[threadIdx.x];   // next tile of the supertile
  }

  // keys: local sort into smem, then contiguous runs to global
#pragma unroll
  for (int r = 0; r < RS_IPT; ++r)
    if (slot[r] != NO_SLOT) {
      sm.stage[slot[r]] = key[r];
      sm.sdig[slot[r]] = (uint16_t)((key[r] >> shift) & mask);
    }
  __syncthreads();
  if (keys_out != nullptr)
    for (int s = threadIdx.x; s < tile_count; s += RS_THREADS)
      keys_out[(uint32_t)(sm.gbase[sm.sdig[s]] + (uint32_t)s)] = sm.stage[s];
  // payloads reuse the staging buffer
#pragma unroll
  for (int q = 0; q < 2; ++q) {
    const uint32_t* pin = q == 0 ? p1_in : p2_in;
    uint32_t* pout = q == 0 ? p1_out : p2_out;
    const bool from_rows = ROWS && q == 0;
    if (pin == nullptr && !from_rows) continue;
    __syncthreads();
#pragma unroll
    for (int r = 0; r < RS_IPT; ++r) {
      const int64_t i = warp_base + r * 32 + lane;
      if (slot[r] != NO_SLOT) sm.stage[slot[r]] = from_rows ? rs.rowid[warp * 32 * RS_IPT + r * 32 + lane] : __ldcs(pin + i);
    }
    __syncthreads();
    for (int s = threadIdx.x; s < tile_count; s += RS_THREADS)
      pout[(uint32_t)(sm.gbase[sm.sdig[s]] + (uint32_t)s)] = sm.stage[s];
  }
  }  // tiles of the supertile
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

// seg_offs[c] = first position whose (sorted) key is >= c, for c in [0, nseg]; int64 or fp32 out
template <typename OutT>
__global__ void __launch_bounds__(256)
segment_offsets_kernel(const uint32_t* __restrict__ sorted_keys, int64_t n, int64_t nseg,
                       OutT* __restrict__ seg_offs) {
  // four consecutive positions per thread (one 16-byte load when aligned); position i closes the segments of the
  // keys in (key[i-1], key[i]]
  const int64_t stride = (int64_t)gridDim.x * blockDim.x * 4;
  for (int64_t i0 = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 4; i0 <= n; i0 += stride) {
    uint32_t k[5];
    k[0] = i0 == 0 ? 0u : sorted_keys[i0 - 1];
    if (i0 + 4 <= n && (reinterpret_cast<uintptr_t>(sorted_keys + i0) & 15) == 0) {
      const uint4 v = __ldcs(reinterpret_cast<const uint4*>(sorted_keys + i0));
      k[1] = v.x; k[2] = v.y; k[3] = v.z; k[4] = v.w;
    } else {
#pragma unroll
      for (int u = 0; u < 4; ++u) k[1 + u] = (i0 + u < n) ? sorted_keys[i0 + u] : 0u;
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int64_t i = i0 + u;
      if (i > n) break;
      // keys above nseg (never produced by our kernels; a caller-supplied assignment could hold one) must not
      // run the loop past the (nseg + 1)-entry output
      const int64_t lo = (i == 0) ? 0 : (int64_t)k[u] + 1;
      const int64_t hi = (i == n) ? nseg : min((int64_t)k[1 + u], nseg);
      for (int64_t c = lo; c <= hi; ++c) seg_offs[c] = (OutT)i;
    }
  }
}

__global__ void iota_kernel(uint32_t* __restrict__ out, int64_t n) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) out[i] = (uint32_t)i;
}


// ---- deterministic y = A^T x (K5) --------------------------------------------------------------------------------
// prod[j] = vals[j] * x[row(j)], a warp per row
__global__ void __launch_bounds__(256)
row_products_kernel(int64_t m, const float* __restrict__ vals, const int64_t* __restrict__ offs,
                    const float* __restrict__ x, uint32_t* __restrict__ prod) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
  if (row >= m) return;
  const int64_t base = offs[0];
  const int64_t beg = offs[row] - base, end = offs[row + 1] - base;
  const float xr = __ldg(x + row);
  for (int64_t j = beg + lane; j < end; j += 32) prod[j] = __float_as_uint(__ldcs(vals + j) * xr);
}

// y[c] (+)= sum of the products of column c in ascending source row: one lane group of 8 per column, the partial
// sums of the 8 lanes combined in lane order -- a fixed summation tree, so the result is reproducible
__global__ void __launch_bounds__(256)
segment_sum_kernel(int64_t n, const int64_t* __restrict__ seg_offs, const uint32_t* __restrict__ prod,
                   float* __restrict__ y, int accumulate) {
  const int lane = threadIdx.x & 31, sl = lane & 7;
  const int64_t c = ((int64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * 4 + (lane >> 3);
  float acc = 0.f;
  if (c < n) {
    const int64_t beg = seg_offs[c], end = seg_offs[c + 1];
    for (int64_t j = beg + sl; j < end; j += 8) acc += __uint_as_float(__ldcs(prod + j));
  }
#pragma unroll
  for (int o = 4; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if (c < n && sl == 0) y[c] = accumulate ? y[c] + acc : acc;
}

int key_bits(int64_t nkeys) {
  int b = 1;
  while (b < 32 && ((int64_t)1 << b) < nkeys) ++b;
  return b;
}

size_t align_up(size_t x) { return (x + 255) & ~(size_t)255; }

size_t counts_bytes(int64_t n) {
  const int64_t tiles = ceil_div<int64_t>(std::max<int64_t>(n, 1), (int64_t)RS_TILE * RS_SUPER);
  const int64_t ncounts = tiles * RS_BINS;
  const int64_t nblocks = ceil_div<int64_t>(ncounts, SC_CHUNK);
  return align_up((size_t)ncounts * 4) + align_up((size_t)nblocks * 4);
}

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

struct Triple {
  uint32_t* key;
  uint32_t* p1;
  uint32_t* p2;
};

// One stable pass on digit `shift`: (in) -> (out).  `counts` holds tiles*256 + scan block sums.
int radix_pass(bof_ctx* ctx, cudaStream_t s, int64_t n, int shift, int bits, const uint32_t* key_in,
               const uint32_t* p1_in, const uint32_t* p2_in, Triple out, uint32_t* counts,
               const int64_t* row_offs = nullptr, int64_t m = 0) {
  const int64_t tiles = ceil_div<int64_t>(n, (int64_t)RS_TILE * RS_SUPER);   // supertiles: one block each
  const int64_t ncounts = tiles * RS_BINS;
  uint32_t* block_sums = reinterpret_cast<uint32_t*>(reinterpret_cast<uint8_t*>(counts) + align_up((size_t)ncounts * 4));
  const uint32_t mask = (1u << bits) - 1u;
  constexpr size_t kScatterSmem = (sizeof(ScatterSmem) + 15) & ~(size_t)15;
  static PerDeviceOnce attr_set[2];
  if (attr_set[0].need(ctx->device)) {
    BOF_CUDA(ctx, cudaFuncSetAttribute(radix_scatter_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kScatterSmem));
    attr_set[0].done(ctx->device);
  }
  if (attr_set[1].need(ctx->device)) {
    BOF_CUDA(ctx, cudaFuncSetAttribute(radix_scatter_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                       (int)(kScatterSmem + sizeof(RowSmem))));
    attr_set[1].done(ctx->device);
  }
  radix_hist_kernel<<<(unsigned)tiles, RH_THREADS, 0, s>>>(key_in, n, shift, mask, counts, (unsigned)tiles);
  BOF_LAUNCH_CHECK(ctx, "radix_hist_kernel");
  int rc = exclusive_scan_u32(ctx, s, counts, ncounts, block_sums);
  if (rc) return rc;
  if (row_offs != nullptr)
    radix_scatter_kernel<true><<<(unsigned)tiles, RS_THREADS, kScatterSmem + sizeof(RowSmem), s>>>(
        key_in, out.key, nullptr, out.p1, p2_in, out.p2, n, shift, mask, counts, (unsigned)tiles, row_offs, m);
  else
    radix_scatter_kernel<false><<<(unsigned)tiles, RS_THREADS, kScatterSmem, s>>>(
        key_in, out.key, p1_in, out.p1, p2_in, out.p2, n, shift, mask, counts, (unsigned)tiles, nullptr, 0);
  BOF_LAUNCH_CHECK(ctx, "radix_scatter_kernel");
  return BOF_OK;
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
                      const int32_t* __restrict__ seg_base, float* __restrict__ sums, float* __restrict__ counts,
                      float* __restrict__ counts_hi) {
  const int64_t c = blockIdx.x;
  const int32_t s0 = seg_base[c], s1 = seg_base[c + 1];
  for (int64_t j = threadIdx.x; j < dim; j += blockDim.x) {
    float acc = 0.f;
    for (int32_t sgm = s0; sgm < s1; ++sgm) acc += partial[(int64_t)sgm * dim + j];
    sums[c * dim + j] = acc;
  }
  if (threadIdx.x == 0) {
    // counts travel through an fp32 sum over the ranks: split into 12 low bits and the rest so that both parts stay
    // exactly representable (a single float is exact only below 2^24 points per cluster)
    const int64_t cnt = seg_offs[c + 1] - seg_offs[c];
    if (counts_hi != nullptr) {
      counts[c] = (float)(cnt & 4095);
      counts_hi[c] = (float)(cnt >> 12);
    } else {
      counts[c] = (float)cnt;
    }
  }
}

__global__ void __launch_bounds__(256)
kmeans_finalize_kernel(int64_t dim, const float* __restrict__ sums, const float* __restrict__ counts,
                       const float* __restrict__ counts_hi, float* __restrict__ centers, float* __restrict__ c_l2sq) {
  __shared__ float red[8];
  const int64_t c = blockIdx.x;
  // with the split form both parts are exact integers (sums over the ranks included); recombine in fp64
  const float cnt = counts_hi ? (float)((double)counts_hi[c] * 4096.0 + (double)counts[c]) : counts[c];
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

// digit width of the radix passes: 8 bits unless bof_config.radix_max_bits asks for less (test knob: more passes)
static int digit_bits(const bof_ctx* ctx) {
  const int b = ctx->cfg.radix_max_bits;
  return (b <= 0 || b > 8) ? 8 : b;
}

size_t csr2csc_workspace_bytes(int64_t m, int64_t n, int64_t nnz) {
  (void)m;
  (void)n;
  const size_t arr = align_up((size_t)std::max<int64_t>(nnz, 1) * 4);
  return 6 * arr + counts_bytes(nnz) + 512;
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
  const size_t arr = align_up((size_t)nnz * 4);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  Triple X{reinterpret_cast<uint32_t*>(base), reinterpret_cast<uint32_t*>(base + arr),
           reinterpret_cast<uint32_t*>(base + 2 * arr)};
  Triple Y{reinterpret_cast<uint32_t*>(base + 3 * arr), reinterpret_cast<uint32_t*>(base + 4 * arr),
           reinterpret_cast<uint32_t*>(base + 5 * arr)};
  uint32_t* counts = reinterpret_cast<uint32_t*>(base + 6 * arr);

  const int db = digit_bits(ctx), kb = key_bits(n);
  const int passes = ceil_div(kb, db);
  const uint32_t* kin = reinterpret_cast<const uint32_t*>(idx);
  const uint32_t* rin = nullptr;   // pass 0 derives the source rows from `offs`
  const uint32_t* vin = reinterpret_cast<const uint32_t*>(vals);
  uint32_t* final_keys = nullptr;
  for (int p = 0; p < passes; ++p) {
    Triple dst = (p % 2 == 0) ? X : Y;
    if (p == passes - 1) {
      // last pass lands in the caller's arrays; the sorted keys go to the idle key buffer
      final_keys = dst.key;
      dst.p1 = reinterpret_cast<uint32_t*>(idx_t);
      dst.p2 = reinterpret_cast<uint32_t*>(vals_t);
    }
    int rc = radix_pass(ctx, s, nnz, db * p, std::min(db, kb - db * p), kin, rin, vin, dst, counts,
                        p == 0 ? offs : nullptr, m);
    if (rc) return rc;
    kin = dst.key;
    rin = dst.p1;
    vin = dst.p2;
  }
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div<int64_t>(nnz + 1, 1024), (int64_t)ctx->num_sms * 32);
  segment_offsets_kernel<int64_t><<<grid, 256, 0, s>>>(final_keys, nnz, n, offs_t);
  BOF_LAUNCH_CHECK(ctx, "segment_offsets_kernel");
  return BOF_OK;
}


// y = A^T x without atomics: products v * x[row] are sorted stably by column (the radix passes of the
// transpose, one 32-bit payload) and every column sums its run in a fixed order.  accumulate: y += instead of y =.
size_t spmv_t_workspace_bytes(int64_t n, int64_t nnz) {
  const size_t arr = align_up((size_t)std::max<int64_t>(nnz, 1) * 4);
  return 5 * arr + counts_bytes(nnz) + align_up((size_t)(n + 1) * 8) + 512;
}

int launch_spmv_t_sorted(bof_ctx* ctx, cudaStream_t s, int accumulate, int64_t m, int64_t n, int64_t nnz, const float* vals,
                         const int32_t* idx, const int64_t* offs, const float* x, float* y, void* ws, size_t ws_bytes) {
  BOF_REQUIRE(ctx, nnz >= 0 && nnz < (1ll << 32) - RS_TILE, "csrgemv 'T': nnz must be below 2^32");
  if (nnz == 0 || m == 0) {
    if (!accumulate) BOF_CUDA(ctx, cudaMemsetAsync(y, 0, (size_t)n * sizeof(float), s));
    return BOF_OK;
  }
  BOF_REQUIRE(ctx, ws != nullptr && ws_bytes >= spmv_t_workspace_bytes(n, nnz), "csrgemv 'T': workspace too small");
  const size_t arr = align_up((size_t)nnz * 4);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  uint32_t* prod0 = reinterpret_cast<uint32_t*>(base);
  Triple X{reinterpret_cast<uint32_t*>(base + arr), reinterpret_cast<uint32_t*>(base + 2 * arr), nullptr};
  Triple Y{reinterpret_cast<uint32_t*>(base + 3 * arr), reinterpret_cast<uint32_t*>(base + 4 * arr), nullptr};
  uint32_t* counts = reinterpret_cast<uint32_t*>(base + 5 * arr);
  int64_t* seg = reinterpret_cast<int64_t*>(base + 5 * arr + counts_bytes(nnz));
  row_products_kernel<<<(unsigned)ceil_div<int64_t>(m, 8), 256, 0, s>>>(m, vals, offs, x, prod0);
  BOF_LAUNCH_CHECK(ctx, "row_products_kernel");
  const int db = digit_bits(ctx), kb = key_bits(n);
  const int passes = ceil_div(kb, db);
  const uint32_t* kin = reinterpret_cast<const uint32_t*>(idx);
  const uint32_t* pin = prod0;
  for (int p = 0; p < passes; ++p) {
    Triple dst = (p % 2 == 0) ? X : Y;
    int rc = radix_pass(ctx, s, nnz, db * p, std::min(db, kb - db * p), kin, pin, nullptr, dst, counts);
    if (rc) return rc;
    kin = dst.key;
    pin = dst.p1;
  }
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div<int64_t>(nnz + 1, 1024), (int64_t)ctx->num_sms * 32);
  segment_offsets_kernel<int64_t><<<grid, 256, 0, s>>>(kin, nnz, n, seg);
  BOF_LAUNCH_CHECK(ctx, "segment_offsets_kernel");
  segment_sum_kernel<<<(unsigned)ceil_div<int64_t>(n, 32), 256, 0, s>>>(n, seg, pin, y, accumulate);
  BOF_LAUNCH_CHECK(ctx, "segment_sum_kernel");
  return BOF_OK;
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
  return 4 * arr + counts_bytes(npoints) + align_up((size_t)(ncenters + 1) * 8) + align_up((size_t)(ncenters + 1) * 4) +
         align_up((size_t)km_max_segments(npoints, ncenters) * (size_t)std::max<int64_t>(dim, 1) * 4) + 256;
}

int launch_kmeans_reduce_ws(bof_ctx* ctx, cudaStream_t s, int64_t npoints, int64_t ncenters,
                            int64_t dim, const float* points, const int32_t* assign, float* sums,
                            float* counts_out, void* ws, size_t ws_bytes, float* counts_hi_out) {
  BOF_REQUIRE(ctx, npoints < (1ll << 31) && ncenters < (1ll << 31), "kmeans_reduce: extents must be below 2^31");
  BOF_REQUIRE(ctx, ncenters > 0 && dim > 0, "kmeans_reduce: empty problem");
  BOF_REQUIRE(ctx, ws != nullptr && ws_bytes >= kmeans_reduce_workspace_bytes(npoints, ncenters, dim),
              "kmeans_reduce: workspace too small");
  const size_t arr = align_up((size_t)std::max<int64_t>(npoints, 1) * 4);
  uint8_t* base = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(ws) + 255) & ~(uintptr_t)255);
  Triple X{reinterpret_cast<uint32_t*>(base), reinterpret_cast<uint32_t*>(base + arr), nullptr};
  Triple Y{reinterpret_cast<uint32_t*>(base + 2 * arr), reinterpret_cast<uint32_t*>(base + 3 * arr), nullptr};
  uint32_t* counts = reinterpret_cast<uint32_t*>(base + 4 * arr);
  int64_t* seg = reinterpret_cast<int64_t*>(base + 4 * arr + counts_bytes(npoints));
  int32_t* seg_base = reinterpret_cast<int32_t*>(reinterpret_cast<uint8_t*>(seg) + align_up((size_t)(ncenters + 1) * 8));
  float* partial = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(seg_base) + align_up((size_t)(ncenters + 1) * 4));

  const uint32_t* kin = reinterpret_cast<const uint32_t*>(assign);
  const uint32_t* pin = nullptr;
  if (npoints > 0) {
    // ids 0..P-1 as the payload of the first pass
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div<int64_t>(npoints, 256), (int64_t)ctx->num_sms * 32);
    iota_kernel<<<grid, 256, 0, s>>>(Y.p1, npoints);
    BOF_LAUNCH_CHECK(ctx, "iota_kernel");
    pin = Y.p1;
    const int db = digit_bits(ctx), kb = key_bits(ncenters);
    const int passes = ceil_div(kb, db);
    for (int p = 0; p < passes; ++p) {
      Triple dst = (p % 2 == 0) ? X : Y;
      int rc = radix_pass(ctx, s, npoints, db * p, std::min(db, kb - db * p), kin, pin, nullptr, dst, counts);
      if (rc) return rc;
      kin = dst.key;
      pin = dst.p1;
    }
  }
  const unsigned grid = (unsigned)std::min<int64_t>(ceil_div<int64_t>(npoints + 1, 256), (int64_t)ctx->num_sms * 32);
  segment_offsets_kernel<int64_t><<<grid, 256, 0, s>>>(kin, npoints, ncenters, seg);
  BOF_LAUNCH_CHECK(ctx, "segment_offsets_kernel");
  kmeans_plan_kernel<<<1, 1024, 0, s>>>(ncenters, seg, seg_base);
  BOF_LAUNCH_CHECK(ctx, "kmeans_plan_kernel");
  kmeans_partial_kernel<<<(unsigned)km_max_segments(npoints, ncenters), 256, 0, s>>>(ncenters, dim, points, pin, seg,
                                                                                     seg_base, partial);
  BOF_LAUNCH_CHECK(ctx, "kmeans_partial_kernel");
  kmeans_combine_kernel<<<(unsigned)ncenters, 256, 0, s>>>(dim, partial, seg, seg_base, sums, counts_out, counts_hi_out);
  BOF_LAUNCH_CHECK(ctx, "kmeans_combine_kernel");
  return BOF_OK;
}

int launch_kmeans_finalize(bof_ctx* ctx, cudaStream_t s, int64_t ncenters, int64_t dim,
                           const float* sums, const float* counts, float* centers, float* c_l2sq,
                           const float* counts_hi) {
  if (ncenters == 0) return BOF_OK;
  kmeans_finalize_kernel<<<(unsigned)ncenters, 256, 0, s>>>(dim, sums, counts, counts_hi, centers, c_l2sq);
  BOF_LAUNCH_CHECK(ctx, "kmeans_finalize_kernel");
  return BOF_OK;
}

}  // namespace bof
