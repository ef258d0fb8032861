// Test-only C wrappers around csrc/map_host.cpp (pure host code), so that the map preprocessing the device kernels index --
// the likelihood field, the free-distance map of the beam walk, the free-cell list -- is checked on the CPU against the
// oracle and against brute force.  Built by tests/test_map_host_cpu.py with g++; not part of the product library.
#include <cstring>

#include "../../beluga_b200/csrc/map_host.hpp"

extern "C" {

int probe_likelihood_field(const bb200_likelihood_field_param* p, const int8_t* cells, int32_t width, int32_t height, double resolution,
                           float* out) {
  const bb200_occupancy_grid grid{cells, width, height, resolution, {1.0, 0.0, 0.0, 0.0}};
  const std::vector<float> field = bb200::make_likelihood_field(*p, grid);
  if (field.size() != static_cast<size_t>(width) * static_cast<size_t>(height)) return -1;
  std::memcpy(out, field.data(), field.size() * sizeof(float));
  return 0;
}

int probe_free_distance(const int8_t* cells, int32_t width, int32_t height, uint8_t* out) {
  const bb200_occupancy_grid grid{cells, width, height, 1.0, {1.0, 0.0, 0.0, 0.0}};
  const std::vector<uint8_t> d = bb200::make_free_distance(grid);
  if (d.size() != static_cast<size_t>(width) * static_cast<size_t>(height)) return -1;
  std::memcpy(out, d.data(), d.size());
  return 0;
}

int64_t probe_free_cells(const int8_t* cells, int32_t width, int32_t height, uint32_t* out, int64_t capacity) {
  const bb200_occupancy_grid grid{cells, width, height, 1.0, {1.0, 0.0, 0.0, 0.0}};
  const std::vector<uint32_t> f = bb200::make_free_cells(grid);
  if (static_cast<int64_t>(f.size()) > capacity) return -1;
  std::memcpy(out, f.data(), f.size() * sizeof(uint32_t));
  return static_cast<int64_t>(f.size());
}
}
