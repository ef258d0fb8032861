#include "map_host.hpp"

#include <algorithm>
#include <cmath>
#include <queue>

namespace bb200 {

namespace {

constexpr int8_t kFree = 0;        // beluga_ros/occupancy_grid.hpp:50
constexpr int8_t kUnknown = -1;    // :52
constexpr int8_t kOccupied = 100;  // :54

struct Frontier {
  uint32_t source;  // nearest obstacle cell found so far for `cell`
  uint32_t cell;
};

// Brushfire over the 4-neighbourhood with the expansion order right, down, left, up
// (sensor/data/linear_grid.hpp:113-130).  Keys are float squared distances between cell
// centroids; ties pop in std::priority_queue order like the reference.
std::vector<float> brushfire(const std::vector<uint8_t>& seeds, uint32_t width, uint32_t height, double resolution, float cap) {
  const size_t count = seeds.size();
  std::vector<float> dist(count, cap);
  std::vector<uint8_t> seen(count, 0);
  auto later = [&dist](const Frontier& a, const Frontier& b) { return dist[a.cell] > dist[b.cell]; };
  std::priority_queue<Frontier, std::vector<Frontier>, decltype(later)> open(later);

  for (size_t i = 0; i < count; ++i) {
    if (seeds[i] != 0) {
      seen[i] = 1;
      dist[i] = 0;
      open.push(Frontier{static_cast<uint32_t>(i), static_cast<uint32_t>(i)});
    }
  }
  // (cell + 0.5) * resolution -- regular_grid.hpp:87-89; the difference of two centroids squared,
  // evaluated in double and narrowed to float (likelihood_field_model_base.hpp:131-133).
  auto centroid = [resolution](uint32_t c) { return (static_cast<double>(static_cast<int>(c)) + 0.5) * resolution; };
  auto sqdist = [&](uint32_t a, uint32_t b) {
    const double dx = centroid(a % width) - centroid(b % width);
    const double dy = centroid(a / width) - centroid(b / width);
    return static_cast<float>(dx * dx + dy * dy);
  };
  auto visit = [&](uint32_t source, uint32_t cell) {
    if (seen[cell]) return;
    seen[cell] = 1;
    const float d = sqdist(source, cell);
    if (d < cap) {
      dist[cell] = d;
      open.push(Frontier{source, cell});
    }
  };
  while (!open.empty()) {
    const Frontier f = open.top();
    open.pop();
    const uint32_t xi = f.cell % width, yi = f.cell / width;
    if (xi + 1 < width) visit(f.source, f.cell + 1);
    if (yi + 1 < height) visit(f.source, f.cell + width);
    if (xi > 0) visit(f.source, f.cell - 1);
    if (yi > 0) visit(f.source, f.cell - width);
  }
  return dist;
}

}  // namespace

std::vector<float> make_likelihood_field(const bb200_likelihood_field_param& p, const bb200_occupancy_grid& g) {
  const uint32_t width = static_cast<uint32_t>(g.width), height = static_cast<uint32_t>(g.height);
  const size_t count = static_cast<size_t>(width) * height;

  std::vector<uint8_t> occupied(count), edge(count, 0);
  for (size_t i = 0; i < count; ++i) occupied[i] = g.cells[i] == kOccupied;
  // obstacle_edge_mask (occupancy_grid.hpp:184-201): occupied with at least one FREE 4-neighbour.
  for (size_t i = 0; i < count; ++i) {
    if (!occupied[i]) continue;
    const uint32_t xi = static_cast<uint32_t>(i % width), yi = static_cast<uint32_t>(i / width);
    bool touches_free = false;
    if (xi + 1 < width) touches_free |= g.cells[i + 1] == kFree;
    if (yi + 1 < height) touches_free |= g.cells[i + width] == kFree;
    if (xi > 0) touches_free |= g.cells[i - 1] == kFree;
    if (yi > 0) touches_free |= g.cells[i - width] == kFree;
    edge[i] = touches_free;
  }

  const double two_squared_sigma = 2 * p.sigma_hit * p.sigma_hit;
  const double pi = 3.14159265358979323846;
  const double amplitude = p.z_hit / (p.sigma_hit * std::sqrt(2 * pi));
  const double offset = p.z_random / p.max_laser_distance;
  const float cap = static_cast<float>(p.max_obstacle_distance * p.max_obstacle_distance);

  std::vector<float> sq = brushfire(p.only_obstacle_boundaries ? edge : occupied, width, height, g.resolution, cap);

  if (p.model_unknown_space) {  // likelihood_field_model_base.hpp:158-177
    const double inverse_max_distance = 1 / p.max_laser_distance;
    const double background = -two_squared_sigma * std::log((inverse_max_distance - offset) / amplitude);
    const float fill = std::min(cap, static_cast<float>(background));
    for (size_t i = 0; i < count; ++i) {
      const bool unknown = g.cells[i] == kUnknown;
      const bool masked = p.only_obstacle_boundaries ? (unknown || (occupied[i] && !edge[i])) : unknown;
      if (masked) sq[i] = fill;
    }
  }
  for (size_t i = 0; i < count; ++i) {
    sq[i] = static_cast<float>(amplitude * std::exp(-static_cast<double>(sq[i]) / two_squared_sigma) + offset);
  }
  return sq;
}

std::vector<uint8_t> make_free_distance(const bb200_occupancy_grid& g) {
  const int w = g.width, h = g.height;
  std::vector<int> d(static_cast<size_t>(w) * h);
  const int cap = 255;
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const size_t i = static_cast<size_t>(y) * w + x;
      const int border = std::min(std::min(x + 1, y + 1), std::min(w - x, h - y));  // distance to the outside
      d[i] = g.cells[i] != kFree ? 0 : std::min(cap, border);
    }
  // Two-pass chamfer with unit weights on the 8-neighbourhood: exact for the Chebyshev metric.
  auto relax = [&](size_t i, int x, int y, int dx, int dy) {
    const int nx = x + dx, ny = y + dy;
    if (nx < 0 || ny < 0 || nx >= w || ny >= h) return;
    const int v = d[static_cast<size_t>(ny) * w + nx] + 1;
    if (v < d[i]) d[i] = v;
  };
  for (int y = 0; y < h; ++y)
    for (int x = 0; x < w; ++x) {
      const size_t i = static_cast<size_t>(y) * w + x;
      relax(i, x, y, -1, 0), relax(i, x, y, -1, -1), relax(i, x, y, 0, -1), relax(i, x, y, 1, -1);
    }
  for (int y = h - 1; y >= 0; --y)
    for (int x = w - 1; x >= 0; --x) {
      const size_t i = static_cast<size_t>(y) * w + x;
      relax(i, x, y, 1, 0), relax(i, x, y, 1, 1), relax(i, x, y, 0, 1), relax(i, x, y, -1, 1);
    }
  std::vector<uint8_t> out(d.size());
  for (size_t i = 0; i < d.size(); ++i) out[i] = static_cast<uint8_t>(std::min(d[i], cap));
  return out;
}

std::vector<uint32_t> make_free_cells(const bb200_occupancy_grid& g) {
  std::vector<uint32_t> out;
  const size_t count = static_cast<size_t>(g.width) * static_cast<size_t>(g.height);
  for (size_t i = 0; i < count; ++i)
    if (g.cells[i] == kFree) out.push_back(static_cast<uint32_t>(i));
  return out;
}

}  // namespace bb200
