// Device side of the cluster-based estimate (reference: beluga/algorithm/cluster_based_estimation.hpp).
//
// The reference groups the particles into spatial-hash cells (make_cluster_map, :141-161), then
// runs a priority-queue flood over the cells (assign_clusters, :205-253) and finally estimates
// mean/covariance per cluster (estimate_clusters, :351-398).  Everything per PARTICLE runs here on
// the device and produces one CellRecord per occupied cell; the per-CELL flood (sequential, order
// dependent, a few thousand cells) runs on the host (cluster_host.hpp).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>

#include "kernels.cuh"

namespace bb200 {

/// What the host pass needs to know about one occupied cell; cells are numbered in the order in
/// which the particle sequence first touches them (the insertion order of the reference's map).
struct CellRecord {
  Pose2 representative;         // state of the first particle that fell into the cell (ClusterCell::representative_state)
  unsigned long long hash;      // spatial hash of the cell
  unsigned int first_index;     // index of that first particle
  unsigned int count;           // ClusterCell::num_particles
  double weight;                // sum of the particle weights in particle order (ClusterCell::weight before normalisation)
  double moments[kMomentCount]; // raw moments of the cell's particles about the pivot (layout of kMomentCount)
};
static_assert(sizeof(CellRecord) == 128, "CellRecord is copied to the host as raw bytes");

/// Scratch buffers of the clusterizer (owned by Filter, sized for `capacity` particles).
struct ClusterScratch {
  uint64_t capacity{0};
  uint64_t table_size{0};              // power of two >= 2 * capacity
  unsigned long long* hashes{nullptr}; // [capacity]
  unsigned long long* keys{nullptr};   // [table_size]
  unsigned int* first{nullptr};        // [table_size] smallest particle index per key
  uint32_t* slot_of{nullptr};          // [capacity] table slot of each particle's cell
  uint32_t* flags{nullptr};            // [capacity] first-occurrence flags, scanned in place
  uint32_t* cell_of{nullptr};          // [capacity] dense cell id per particle
  uint32_t* starts{nullptr};           // [capacity + 1] cell counts, scanned in place into segment starts
  uint32_t* keys_a{nullptr};           // radix sort ping-pong: keys and particle indices
  uint32_t* keys_b{nullptr};
  uint32_t* idx_a{nullptr};
  uint32_t* idx_b{nullptr};
  uint32_t* histogram{nullptr};        // [256 * sort tiles]
  unsigned long long* tile_state{nullptr};
  unsigned long long* words{nullptr};  // [4]: scan ticket, cell count, spare
  CellRecord* records{nullptr};
  uint64_t records_capacity{0};
};

uint32_t cluster_sort_tiles(uint64_t n);

/// hashes -> hash set (key -> smallest particle index) -> first-occurrence flags -> dense cell ids
/// base; leaves the number of occupied cells in scratch.words[1].
void launch_cluster_cells_begin(const Pose2* states, uint64_t n, double linear_resolution, double angular_resolution, const ClusterScratch& s,
                                cudaStream_t stream);

/// Cell id per particle, cell sizes and the stable sort of the particle indices by cell (particle order
/// preserved inside a cell).  `cells` = value read back from words[1].  Returns the buffer holding the
/// sorted particle indices and the number of kernels launched.
const uint32_t* launch_cluster_sort(uint64_t n, uint32_t cells, const ClusterScratch& s, cudaStream_t stream, int* launches);
/// One CellRecord per cell from the sorted indices.
void launch_cluster_records(const Pose2* states, const double* weights, const uint32_t* sorted_idx, uint32_t cells, double pivot_x, double pivot_y,
                            const ClusterScratch& s, cudaStream_t stream);

}  // namespace bb200
