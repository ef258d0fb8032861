// oracle/cluster_oracle.hpp -- CPU restatement of beluga's cluster-based estimate.
//
// TEST INFRASTRUCTURE ONLY: nothing under beluga_b200/ may include, link or execute this file.
//
// Follows beluga/include/beluga/algorithm/cluster_based_estimation.hpp (cited per function).  The
// reference's result depends on two libstdc++ behaviours, which this restatement gets by using the
// very same containers from the same standard library (GCC 13 here):
//   * the iteration order of std::unordered_map<std::size_t, ClusterCell> (:127) -- it fixes the
//     layout of the heap built in make_priority_queue (:73-95), hence which of several equally heavy
//     cells is popped first;
//   * std::priority_queue's push/pop order for equal priorities.
// Parity notes ("unpinned" items): estimate_clusters sorts the particles by cluster with
// ranges::sort (:362), an unstable introsort from range-v3 (absent here), so the summation order
// inside one cluster is unspecified; this file sums in particle order (std::stable_sort).  That only
// moves the per-cluster mean/covariance at rounding level (the reference tests use 1e-3 / 1e-6).
#pragma once

#include <algorithm>
#include <cstddef>
#include <cstdint>
#include <optional>
#include <queue>
#include <unordered_map>
#include <utility>
#include <vector>

#include "beluga_oracle.hpp"

namespace oracle {

/// cluster_based_estimation.hpp:119-124
struct ClusterCell {
  SE2 representative_state;
  double weight{0.0};
  std::size_t num_particles{0};
  std::optional<std::size_t> cluster_id;
};

using ClusterMap = std::unordered_map<std::size_t, ClusterCell>;  // :127

/// :107-112
inline double calculate_percentile_threshold(std::vector<double> values, double percentile) {
  const auto n = static_cast<std::ptrdiff_t>(static_cast<double>(values.size()) * percentile);
  std::nth_element(values.begin(), values.begin() + n, values.end());
  return values[static_cast<std::size_t>(n)];
}

/// :141-161
inline ClusterMap make_cluster_map(const std::vector<SE2>& states, const std::vector<double>& weights, const std::vector<std::size_t>& hashes) {
  ClusterMap map;
  map.reserve(states.size() / 5);
  for (std::size_t i = 0; i < states.size(); ++i) {
    auto [it, inserted] = map.try_emplace(hashes[i], ClusterCell{});
    ClusterCell& entry = it->second;
    entry.weight += weights[i];
    entry.num_particles++;
    if (inserted) entry.representative_state = states[i];
  }
  return map;
}

/// :177-192
inline void normalize_and_cap_weights(ClusterMap& map, double percentile) {
  for (auto& kv : map) kv.second.weight /= static_cast<double>(kv.second.num_particles);
  std::vector<double> values;
  values.reserve(map.size());
  for (const auto& kv : map) values.push_back(kv.second.weight);
  const double max_weight = calculate_percentile_threshold(std::move(values), percentile);
  for (auto& kv : map) kv.second.weight = std::min(kv.second.weight, max_weight);
}

/// :205-253.  `neighbors(state)` returns the hashes of the adjacent cells in the reference's order.
template <class NeighborsFunction>
void assign_clusters(ClusterMap& map, NeighborsFunction&& neighbors) {
  struct KeyWithPriority {  // :75-82
    double priority;
    std::size_t key;
    bool operator<(const KeyWithPriority& other) const { return priority < other.priority; }
  };
  std::vector<KeyWithPriority> initial;
  initial.reserve(map.size());
  for (const auto& kv : map) initial.push_back(KeyWithPriority{kv.second.weight, kv.first});  // map iteration order (:88-94)
  std::priority_queue<KeyWithPriority> queue(initial.begin(), initial.end());
  const double max_priority = queue.top().priority;

  std::size_t next_cluster_id = 0;
  while (!queue.empty()) {
    const std::size_t hash = queue.top().key;
    queue.pop();
    ClusterCell& cell = map[hash];
    if (!cell.cluster_id.has_value()) cell.cluster_id = next_cluster_id++;
    for (const std::size_t neighbor_hash : neighbors(cell.representative_state)) {
      auto it = map.find(neighbor_hash);
      const bool valid = it != map.end() && !it->second.cluster_id.has_value() && it->second.weight <= cell.weight;  // :232-238
      if (!valid) continue;
      it->second.cluster_id = cell.cluster_id;
      queue.push(KeyWithPriority{max_priority + it->second.weight, neighbor_hash});  // :247
    }
  }
}

/// :259-276
struct ParticleClusterizerParam {
  double linear_hash_resolution = 0.20;
  double angular_hash_resolution = 0.524;
  double weight_cap_percentile = 0.90;
};

/// ParticleClusterizer (:279-338)
struct ParticleClusterizer {
  ParticleClusterizerParam p;

  [[nodiscard]] std::size_t hash(const SE2& s) const {
    return spatial_hash(s, p.linear_hash_resolution, p.linear_hash_resolution, p.angular_hash_resolution);
  }

  /// :289-293 with the six adjacent poses of :330-337 (+x, -x, +y, -y, +theta, -theta).
  [[nodiscard]] std::vector<std::size_t> neighbors(const SE2& pose) const {
    const double l = p.linear_hash_resolution, a = p.angular_hash_resolution;
    const SE2 adjacent[6] = {SE2{SO2{0.0}, +l, 0.0}, SE2{SO2{0.0}, -l, 0.0}, SE2{SO2{0.0}, 0.0, +l},
                             SE2{SO2{0.0}, 0.0, -l}, SE2{SO2{+a}, 0.0, 0.0}, SE2{SO2{-a}, 0.0, 0.0}};
    std::vector<std::size_t> out;
    for (const SE2& n : adjacent) out.push_back(hash(pose * n));
    return out;
  }

  /// :304-316
  [[nodiscard]] std::vector<std::size_t> operator()(const std::vector<SE2>& states, const std::vector<double>& weights) const {
    std::vector<std::size_t> hashes(states.size());
    for (std::size_t i = 0; i < states.size(); ++i) hashes[i] = hash(states[i]);
    ClusterMap map = make_cluster_map(states, weights, hashes);
    normalize_and_cap_weights(map, p.weight_cap_percentile);
    assign_clusters(map, [this](const SE2& s) { return neighbors(s); });
    std::vector<std::size_t> out(states.size());
    for (std::size_t i = 0; i < states.size(); ++i) out[i] = map[hashes[i]].cluster_id.value();
    return out;
  }
};

struct ClusterEstimate {
  double weight;
  Estimate estimate;
  std::size_t cluster;
};

/// estimate_clusters (:351-398): one estimate per cluster with more than one particle, ascending cluster id.
inline std::vector<ClusterEstimate> estimate_clusters(const std::vector<SE2>& states, const std::vector<double>& weights,
                                                      const std::vector<std::size_t>& clusters) {
  std::vector<std::size_t> order(states.size());
  for (std::size_t i = 0; i < order.size(); ++i) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](std::size_t a, std::size_t b) { return clusters[a] < clusters[b]; });
  std::vector<ClusterEstimate> out;
  std::size_t begin = 0;
  while (begin < order.size()) {
    std::size_t end = begin;
    while (end < order.size() && clusters[order[end]] == clusters[order[begin]]) ++end;
    if (end - begin > 1) {  // :383 a single sample has no covariance
      std::vector<SE2> s;
      std::vector<double> w;
      for (std::size_t k = begin; k < end; ++k) s.push_back(states[order[k]]), w.push_back(weights[order[k]]);
      double total = 0.0;
      for (const double v : w) total += v;
      out.push_back(ClusterEstimate{total, estimate(s, w), clusters[order[begin]]});
    }
    begin = end;
  }
  return out;
}

/// cluster_based_estimate (:415-432)
inline Estimate cluster_based_estimate(const std::vector<SE2>& states, const std::vector<double>& weights, const ParticleClusterizerParam& p,
                                       std::vector<std::size_t>* clusters_out = nullptr) {
  const std::vector<std::size_t> clusters = ParticleClusterizer{p}(states, weights);
  if (clusters_out != nullptr) *clusters_out = clusters;
  const std::vector<ClusterEstimate> per_cluster = estimate_clusters(states, weights, clusters);
  if (per_cluster.empty()) return estimate(states, weights);
  // ranges::max_element: the first of the largest.
  const ClusterEstimate* best = &per_cluster.front();
  for (const ClusterEstimate& e : per_cluster)
    if (best->weight < e.weight) best = &e;
  return best->estimate;
}

}  // namespace oracle
