// Host pass of the cluster-based estimate: the per-cell flood of
// beluga/algorithm/cluster_based_estimation.hpp (normalize_and_cap_weights :177-192,
// assign_clusters :205-253) and the per-cluster selection of estimate_clusters /
// cluster_based_estimate (:351-432), fed with one record per occupied cell from the device.
#pragma once

#include <cstddef>
#include <cstdint>
#include <vector>

namespace bb200 {

/// Host view of one device CellRecord (cluster.cuh), field for field.
struct HostCell {
  double representative[4];  // cos, sin, x, y
  unsigned long long hash;
  unsigned int first_index;
  unsigned int count;
  double weight;
  double moments[9];
};
static_assert(sizeof(HostCell) == 128, "must match bb200::CellRecord");

struct ClusterSelection {
  std::vector<uint32_t> cluster_of_cell;  // cluster id of every cell (assign_clusters)
  uint32_t clusters{0};                   // number of cluster ids handed out
  bool found{false};                      // a cluster with more than one particle exists
  uint32_t best{0};                       // its id: the heaviest such cluster (first of equals)
  double moments[9]{};                    // raw moments to estimate from: that cluster's, or the whole set's
};

/// cells in first-occurrence order (the insertion order of the reference's unordered_map);
/// n_particles sizes the map's initial reserve like make_cluster_map (:146).
ClusterSelection select_cluster(const HostCell* cells, std::size_t n_cells, std::uint64_t n_particles, double linear_resolution,
                                double angular_resolution, double weight_cap_percentile);

}  // namespace bb200
