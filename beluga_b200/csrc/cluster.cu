// Cluster-based estimate, device side (see cluster.cuh).  Hand-written sm_100a kernels:
//   cluster_insert / cluster_flag   spatial hash per particle, hash set "cell -> first particle"
//   cluster_cell_of                 dense cell id per particle (cells numbered by first occurrence)
//   radix_histogram / radix_scatter stable LSD radix sort of the particle indices by cell id
//   cluster_cells                   one warp per cell: ordered weight sum, moments, representative
// Nothing here uses floating-point atomics: every sum has a fixed order, so results are reproducible
// run to run and the per-cell weight is the reference's sequential sum bit for bit.
#include "cluster.cuh"

#include <algorithm>

namespace bb200 {

namespace {

constexpr int kWarp = 32;
constexpr unsigned long long kEmpty = ~0ull;

__device__ __forceinline__ Pose2 load_state(const Pose2* p) {
  const double2 a = *reinterpret_cast<const double2*>(p);
  const double2 b = *(reinterpret_cast<const double2*>(p) + 1);
  return Pose2{a.x, a.y, b.x, b.y};
}

__device__ __forceinline__ uint64_t probe_start(unsigned long long key, uint64_t mask) {
  const unsigned long long h = key * 0x9E3779B97F4A7C15ull;
  return (h ^ (h >> 29)) & mask;
}

// ---- cells -------------------------------------------------------------------------------------------

/// spatial_hash<SE2d>{linear, linear, angular} (cluster_based_estimation.hpp:321-325) of every state,
/// inserted into an open-addressing set that keeps the smallest particle index per cell.
__global__ void __launch_bounds__(256) cluster_insert_kernel(const Pose2* __restrict__ states, uint64_t n, double linear, double angular,
                                                             unsigned long long* __restrict__ hashes, unsigned long long* keys,
                                                             unsigned int* first, uint64_t mask) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  const unsigned long long hash = spatial_hash(load_state(states + i), linear, linear, angular);
  hashes[i] = hash;
  const unsigned long long key = hash == kEmpty ? kEmpty - 1 : hash;  // the sentinel itself cannot be stored (2^-64 event)
  uint64_t pos = probe_start(key, mask);
  for (;;) {
    const unsigned long long prev = atomicCAS(keys + pos, kEmpty, key);
    if (prev == kEmpty || prev == key) {
      atomicMin(first + pos, static_cast<unsigned int>(i));
      return;
    }
    pos = (pos + 1) & mask;
  }
}

/// flags[i] = 1 when particle i is the first of its cell (try_emplace inserted, :151-156).
__global__ void __launch_bounds__(256) cluster_flag_kernel(const unsigned long long* __restrict__ hashes, uint64_t n,
                                                           const unsigned long long* __restrict__ keys, const unsigned int* __restrict__ first,
                                                           uint64_t mask, uint32_t* __restrict__ slot_of, uint32_t* __restrict__ flags) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  const unsigned long long hash = hashes[i];
  const unsigned long long key = hash == kEmpty ? kEmpty - 1 : hash;
  uint64_t pos = probe_start(key, mask);
  while (keys[pos] != key) pos = (pos + 1) & mask;
  slot_of[i] = static_cast<uint32_t>(pos);
  flags[i] = first[pos] == static_cast<unsigned int>(i) ? 1u : 0u;
}

/// cell id = number of first occurrences before the cell's first particle; also the cell sizes.
__global__ void __launch_bounds__(256) cluster_cell_of_kernel(const uint32_t* __restrict__ slot_of, const unsigned int* __restrict__ first,
                                                              const uint32_t* __restrict__ exclusive, uint64_t n, uint32_t* __restrict__ cell_of,
                                                              uint32_t* counts) {
  const uint64_t i = static_cast<uint64_t>(blockIdx.x) * 256 + threadIdx.x;
  if (i >= n) return;
  const uint32_t cell = exclusive[first[slot_of[i]]];
  cell_of[i] = cell;
  atomicAdd(counts + cell, 1u);
}

// ---- stable radix sort of (cell id, particle index) ----------------------------------------------
// LSD, 8 bits per pass, tiles of 2048 keys.  Stability (equal keys keep their input order) is what
// leaves every cell's particles in particle order after the last pass.

constexpr int kSortThreads = 256;
constexpr int kSortItems = 8;
constexpr int kSortTile = kSortThreads * kSortItems;
constexpr int kRadix = 256;

__global__ void __launch_bounds__(kSortThreads) radix_histogram_kernel(const uint32_t* __restrict__ keys, uint32_t n, int shift, uint32_t tiles,
                                                                      uint32_t* __restrict__ histogram) {
  __shared__ uint32_t s_hist[kRadix];
  s_hist[threadIdx.x] = 0;
  __syncthreads();
  const uint32_t base = blockIdx.x * kSortTile;
#pragma unroll
  for (int r = 0; r < kSortItems; ++r) {
    const uint32_t e = base + r * kSortThreads + threadIdx.x;
    if (e < n) atomicAdd(&s_hist[(keys[e] >> shift) & (kRadix - 1)], 1u);
  }
  __syncthreads();
  histogram[threadIdx.x * tiles + blockIdx.x] = s_hist[threadIdx.x];  // digit-major: one scan gives every (digit, tile) base
}

__global__ void __launch_bounds__(kSortThreads) radix_scatter_kernel(const uint32_t* __restrict__ keys_in, const uint32_t* __restrict__ idx_in,
                                                                    uint32_t n, int shift, uint32_t tiles, const uint32_t* __restrict__ bases,
                                                                    uint32_t* __restrict__ keys_out, uint32_t* __restrict__ idx_out) {
  constexpr int kWarps = kSortThreads / kWarp;
  __shared__ uint32_t s_base[kRadix];            // next free output position per digit for this tile
  __shared__ uint32_t s_count[kWarps][kRadix];   // keys per (warp, digit) in the current round
  const int warp = threadIdx.x / kWarp, lane = threadIdx.x % kWarp;
  s_base[threadIdx.x] = bases[threadIdx.x * tiles + blockIdx.x];
  const uint32_t base = blockIdx.x * kSortTile;
  for (int r = 0; r < kSortItems; ++r) {
#pragma unroll
    for (int w = 0; w < kWarps; ++w) s_count[w][threadIdx.x] = 0;
    __syncthreads();
    const uint32_t e = base + r * kSortThreads + threadIdx.x;
    const bool valid = e < n;
    const uint32_t key = valid ? keys_in[e] : 0u;
    const uint32_t digit = valid ? ((key >> shift) & (kRadix - 1)) : kRadix;  // invalid lanes form their own group
    const unsigned peers = __match_any_sync(0xffffffffu, digit);
    const uint32_t rank = __popc(peers & ((1u << lane) - 1u));
    if (valid && rank == 0) s_count[warp][digit] = __popc(peers);
    __syncthreads();
    // Thread d turns the per-warp counts of digit d into offsets (warp order = element order).
    {
      uint32_t running = s_base[threadIdx.x];
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        const uint32_t c = s_count[w][threadIdx.x];
        s_count[w][threadIdx.x] = running;
        running += c;
      }
      s_base[threadIdx.x] = running;
    }
    __syncthreads();
    if (valid) {
      const uint32_t pos = s_count[warp][digit] + rank;
      keys_out[pos] = key;
      idx_out[pos] = idx_in != nullptr ? idx_in[e] : e;
    }
    __syncthreads();
  }
}

// ---- one record per cell ---------------------------------------------------------------------------

constexpr int kCellThreads = 128;

__global__ void __launch_bounds__(kCellThreads) cluster_cells_kernel(const Pose2* __restrict__ states, const double* __restrict__ weights,
                                                                    const unsigned long long* __restrict__ hashes,
                                                                    const uint32_t* __restrict__ sorted_idx, const uint32_t* __restrict__ starts,
                                                                    uint32_t cells, double px, double py, CellRecord* __restrict__ records) {
  const uint32_t cell = (blockIdx.x * kCellThreads + threadIdx.x) / kWarp;
  if (cell >= cells) return;  // whole warps leave together
  const int lane = threadIdx.x % kWarp;
  const uint32_t begin = starts[cell], end = starts[cell + 1];

  double m[kMomentCount];
#pragma unroll
  for (int k = 0; k < kMomentCount; ++k) m[k] = 0.0;
  bool unit = true;
  auto add = [&](const Pose2& st, double w) {
    const double dx = st.x - px, dy = st.y - py;
    m[0] += w;
    m[1] += w * w;
    m[2] += w * st.c;
    m[3] += w * st.s;
    m[4] += w * dx;
    m[5] += w * dy;
    m[6] += w * dx * dx;
    m[7] += w * dx * dy;
    m[8] += w * dy * dy;
    unit = unit && w == 1.0;
  };
  uint32_t p = begin + lane;
  for (; p + 3 * kWarp < end; p += 4 * kWarp) {  // four gathers in flight per lane
    const uint32_t i0 = sorted_idx[p], i1 = sorted_idx[p + kWarp], i2 = sorted_idx[p + 2 * kWarp], i3 = sorted_idx[p + 3 * kWarp];
    const double w0 = weights[i0], w1 = weights[i1], w2 = weights[i2], w3 = weights[i3];
    const Pose2 s0 = load_state(states + i0), s1 = load_state(states + i1), s2 = load_state(states + i2), s3 = load_state(states + i3);
    add(s0, w0);
    add(s1, w1);
    add(s2, w2);
    add(s3, w3);
  }
  for (; p < end; p += kWarp) {
    const uint32_t i = sorted_idx[p];
    add(load_state(states + i), weights[i]);
  }
#pragma unroll
  for (int k = 0; k < kMomentCount; ++k) {
#pragma unroll
    for (int off = kWarp / 2; off > 0; off >>= 1) m[k] = m[k] + __shfl_xor_sync(0xffffffffu, m[k], off);
  }
  unit = __all_sync(0xffffffffu, unit);

  // ClusterCell::weight: `entry.weight += weight` in particle order (:153).  With unit weights (the
  // state after every resample) all partial sums are integers, so the count is that sum exactly;
  // otherwise the warp replays the sequential chain, 32 weights per coalesced load.
  double total = static_cast<double>(end - begin);
  if (!unit) {
    total = 0.0;
    double w = begin + lane < end ? weights[sorted_idx[begin + lane]] : 0.0;
    for (uint32_t base = begin; base < end; base += kWarp) {
      const uint32_t q = base + kWarp + lane;  // the next 32 weights travel while this batch is added up
      const double w_next = q < end ? weights[sorted_idx[q]] : 0.0;
      const int valid = static_cast<int>(min(static_cast<uint32_t>(kWarp), end - base));
      for (int k = 0; k < valid; ++k) total = total + __shfl_sync(0xffffffffu, w, k);
      w = w_next;
    }
  }
  if (lane == 0) {
    const uint32_t first = sorted_idx[begin];  // stable sort: the cell's earliest particle
    CellRecord r;
    r.hash = hashes[first];
    r.first_index = first;
    r.count = end - begin;
    r.weight = total;
    r.representative = load_state(states + first);
#pragma unroll
    for (int k = 0; k < kMomentCount; ++k) r.moments[k] = m[k];
    records[cell] = r;
  }
}

int bits_for(uint32_t values) {  // bits needed to represent 0 .. values-1
  int b = 0;
  while (b < 32 && (1ull << b) < values) ++b;
  return b;
}

}  // namespace

uint32_t cluster_sort_tiles(uint64_t n) { return static_cast<uint32_t>((n + kSortTile - 1) / kSortTile); }

void launch_cluster_cells_begin(const Pose2* states, uint64_t n, double linear_resolution, double angular_resolution, const ClusterScratch& s,
                                cudaStream_t stream) {
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  cudaMemsetAsync(s.keys, 0xFF, s.table_size * sizeof(unsigned long long), stream);
  cudaMemsetAsync(s.first, 0xFF, s.table_size * sizeof(unsigned int), stream);
  cluster_insert_kernel<<<blocks, 256, 0, stream>>>(states, n, linear_resolution, angular_resolution, s.hashes, s.keys, s.first, s.table_size - 1);
  cluster_flag_kernel<<<blocks, 256, 0, stream>>>(s.hashes, n, s.keys, s.first, s.table_size - 1, s.slot_of, s.flags);
  launch_scan_u32(s.flags, s.flags, static_cast<uint32_t>(n), s.words + 0, s.tile_state, s.words + 1, stream);
}

const uint32_t* launch_cluster_sort(uint64_t n, uint32_t cells, const ClusterScratch& s, cudaStream_t stream, int* launches) {
  const unsigned blocks = static_cast<unsigned>((n + 255) / 256);
  const uint32_t n32 = static_cast<uint32_t>(n);
  cudaMemsetAsync(s.starts, 0, (static_cast<size_t>(cells) + 1) * sizeof(uint32_t), stream);
  cluster_cell_of_kernel<<<blocks, 256, 0, stream>>>(s.slot_of, s.first, s.flags, n, s.cell_of, s.starts);
  launch_scan_u32(s.starts, s.starts, cells + 1, s.words + 0, s.tile_state, nullptr, stream);

  const uint32_t tiles = cluster_sort_tiles(n);
  const int passes = std::max(1, (bits_for(cells) + 7) / 8);
  const uint32_t* keys_in = s.cell_of;
  const uint32_t* idx_in = nullptr;  // first pass: the identity
  uint32_t* keys_out = s.keys_a;
  uint32_t* idx_out = s.idx_a;
  for (int pass = 0; pass < passes; ++pass) {
    radix_histogram_kernel<<<tiles, kSortThreads, 0, stream>>>(keys_in, n32, 8 * pass, tiles, s.histogram);
    launch_scan_u32(s.histogram, s.histogram, kRadix * tiles, s.words + 0, s.tile_state, nullptr, stream);
    radix_scatter_kernel<<<tiles, kSortThreads, 0, stream>>>(keys_in, idx_in, n32, 8 * pass, tiles, s.histogram, keys_out, idx_out);
    keys_in = keys_out;
    idx_in = idx_out;
    keys_out = keys_out == s.keys_a ? s.keys_b : s.keys_a;
    idx_out = idx_out == s.idx_a ? s.idx_b : s.idx_a;
  }
  if (launches != nullptr) *launches = 2 + 3 * passes;
  return idx_in;
}

void launch_cluster_records(const Pose2* states, const double* weights, const uint32_t* sorted_idx, uint32_t cells, double pivot_x, double pivot_y,
                            const ClusterScratch& s, cudaStream_t stream) {
  const unsigned cell_blocks = static_cast<unsigned>((static_cast<uint64_t>(cells) * kWarp + kCellThreads - 1) / kCellThreads);
  cluster_cells_kernel<<<cell_blocks, kCellThreads, 0, stream>>>(states, weights, s.hashes, sorted_idx, s.starts, cells, pivot_x, pivot_y, s.records);
}

}  // namespace bb200
