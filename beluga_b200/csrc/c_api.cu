// extern "C" surface of libbeluga_b200.so (declared in include/beluga_b200.h).
#include <algorithm>
#include <cmath>
#include <limits>
#include <memory>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "../../include/beluga_b200.h"
#include "amcl.hpp"
#include "cluster_host.hpp"
#include "filter.hpp"

using bb200::Amcl;
using bb200::Filter;

struct bb200_filter {
  Filter impl;
  explicit bb200_filter(const bb200_filter_config& c) : impl(c) {}
};

struct bb200_amcl {
  Amcl impl;
  bb200_filter* filter_view;  // non-owning alias handed out by bb200_amcl_filter
  bb200_amcl(const bb200_amcl_param& p, const bb200_motion_param& m) : impl(p, m), filter_view(nullptr) {}
};

/// One filter over several shards driven by ONE host thread (the shape of beluga_ros's single-process node).
struct bb200_sharded_amcl {
  std::vector<bb200_amcl*> ranks;
  mutable std::string error;
  void record_error(const std::string& m) const { error = m; }
  ~bb200_sharded_amcl() {
    for (bb200_amcl* a : ranks) delete a;
  }
};

namespace {
thread_local std::string g_create_error;

// bb200_filter is layout-compatible with its only member, so the Amcl-owned Filter can be viewed
// through the same handle type without a second allocation.
static_assert(sizeof(bb200_filter) == sizeof(Filter), "bb200_filter must wrap Filter exactly");

// Nothing throws across the C boundary: the host side allocates (std::vector, std::string, unordered_map), so every
// entry point runs its body under this guard and turns an exception into a status + last_error message.
template <class Context, class Body>
int guarded(Context& context, Body&& body) noexcept {
  try {
    return body();
  } catch (const std::bad_alloc&) {
    try {
      context.record_error("out of host memory");
    } catch (...) {
    }
    return BB200_ERR_CAPACITY;
  } catch (const std::exception& e) {
    try {
      context.record_error(std::string("internal error: ") + e.what());
    } catch (...) {
    }
    return BB200_ERR_STATE;
  } catch (...) {
    try {
      context.record_error("internal error: unknown exception");
    } catch (...) {
    }
    return BB200_ERR_STATE;
  }
}
struct CreateErrorContext {
  void record_error(const std::string& m) { g_create_error = m; }
};
template <class Body>
int guarded_create(Body&& body) noexcept {
  CreateErrorContext c;
  return guarded(c, body);
}
}  // namespace

extern "C" {

int bb200_abi_version(void) { return BB200_ABI_VERSION; }

int bb200_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    (void)cudaGetLastError();
    return 0;
  }
  return n;
}

const char* bb200_create_error(void) { return g_create_error.c_str(); }

int bb200_filter_create(const bb200_filter_config* config, bb200_filter** out) {
  if (config == nullptr || out == nullptr) {
    g_create_error = "null argument";
    return BB200_ERR_INVALID_ARGUMENT;
  }
  *out = nullptr;
  return guarded_create([&] {
    bb200_filter* f = new bb200_filter(*config);
    if (!f->impl.ok()) {
      g_create_error = f->impl.last_error();
      const int st = f->impl.create_status();
      delete f;
      return st;
    }
    *out = f;
    return static_cast<int>(BB200_OK);
  });
}

void bb200_filter_destroy(bb200_filter* f) { delete f; }
const char* bb200_last_error(const bb200_filter* f) { return f != nullptr ? f->impl.last_error() : "null filter"; }

#define BB_REQUIRE(cond)                              \
  do {                                                \
    if (!(cond)) return BB200_ERR_INVALID_ARGUMENT;   \
  } while (0)

int bb200_filter_set_likelihood_field_map(bb200_filter* f, const bb200_likelihood_field_param* p, const bb200_occupancy_grid* grid, int prob) {
  BB_REQUIRE(f && p && grid);
  return guarded(f->impl, [&] { return f->impl.set_likelihood_field_map(*p, *grid, prob != 0); });
}
int bb200_filter_set_beam_map(bb200_filter* f, const bb200_beam_param* p, const bb200_occupancy_grid* grid) {
  BB_REQUIRE(f && p && grid);
  return guarded(f->impl, [&] { return f->impl.set_beam_map(*p, *grid); });
}
int bb200_filter_get_likelihood_field(const bb200_filter* f, float* out, uint64_t capacity) {
  BB_REQUIRE(f && out);
  return guarded(f->impl, [&] { return f->impl.get_likelihood_field(out, capacity); });
}
int bb200_filter_set_particles(bb200_filter* f, const double* states, const double* weights, uint64_t n) {
  BB_REQUIRE(f);
  return guarded(f->impl, [&] { return f->impl.set_particles(states, weights, n); });
}
int bb200_filter_size(const bb200_filter* f, uint64_t* n) {
  BB_REQUIRE(f && n);
  *n = f->impl.size();
  return BB200_OK;
}
int bb200_filter_get_particles(bb200_filter* f, double* states, double* weights, uint64_t capacity) {
  BB_REQUIRE(f);
  return guarded(f->impl, [&] { return f->impl.get_particles(states, weights, capacity); });
}
int bb200_filter_initialize_normal(bb200_filter* f, const double mean_xytheta[3], const double cov[9], uint64_t n) {
  BB_REQUIRE(f && mean_xytheta && cov);
  return guarded(f->impl, [&] { return f->impl.initialize_normal(mean_xytheta, cov, n); });
}
int bb200_filter_initialize_uniform(bb200_filter* f, uint64_t n) {
  BB_REQUIRE(f);
  return guarded(f->impl, [&] { return f->impl.initialize_uniform(n); });
}
int bb200_filter_propagate(bb200_filter* f, const bb200_motion_sampling* s, uint32_t step) {
  BB_REQUIRE(f && s);
  return guarded(f->impl, [&] { return f->impl.propagate_reweight(s, step, nullptr, 0); });
}
int bb200_filter_reweight(bb200_filter* f, const double* points_xy, uint64_t n_points) {
  BB_REQUIRE(f && (points_xy || n_points == 0));
  static const double kNoPoints[2] = {0.0, 0.0};
  return guarded(f->impl, [&] { return f->impl.propagate_reweight(nullptr, 0, points_xy != nullptr ? points_xy : kNoPoints, n_points); });
}
int bb200_filter_propagate_reweight(bb200_filter* f, const bb200_motion_sampling* s, uint32_t step, const double* points_xy, uint64_t n_points) {
  BB_REQUIRE(f && s && (points_xy || n_points == 0));
  static const double kNoPoints[2] = {0.0, 0.0};
  return guarded(f->impl, [&] { return f->impl.propagate_reweight(s, step, points_xy != nullptr ? points_xy : kNoPoints, n_points); });
}
int bb200_filter_max_weight(bb200_filter* f, double* wmax) {
  BB_REQUIRE(f && wmax);
  return guarded(f->impl, [&] { return f->impl.max_weight(wmax); });
}
int bb200_filter_build_cdf(bb200_filter* f, double global_wmax, uint64_t* local_total, int* exponent) {
  BB_REQUIRE(f);
  return guarded(f->impl, [&] { return f->impl.build_cdf(global_wmax, local_total, exponent); });
}
int bb200_filter_normalize_by(bb200_filter* f, uint64_t global_total, double* local_sum_sq) {
  BB_REQUIRE(f);
  return guarded(f->impl, [&] { return f->impl.normalize_by(global_total, local_sum_sq); });
}
int bb200_filter_normalize(bb200_filter* f, double* factor, double* sum_sq) {
  BB_REQUIRE(f);
  return guarded(f->impl, [&] { return f->impl.normalize(factor, sum_sq); });
}
int bb200_filter_resample(bb200_filter* f, const bb200_resample_opts* o, uint64_t* new_size) {
  BB_REQUIRE(f && o);
  return guarded(f->impl, [&] { return f->impl.resample(*o, new_size); });
}
int bb200_filter_resample_range(bb200_filter* f, const bb200_resample_opts* o, uint64_t global_total, uint64_t cdf_offset, uint64_t slot_begin,
                                uint64_t slot_end) {
  BB_REQUIRE(f && o);
  return guarded(f->impl, [&] { return f->impl.resample_range(*o, global_total, cdf_offset, slot_begin, slot_end); });
}
int bb200_filter_adopt(bb200_filter* f, uint64_t n, int from_staging) {
  BB_REQUIRE(f);
  return guarded(f->impl, [&] { return f->impl.adopt(n, from_staging); });
}
int bb200_systematic_comb(uint64_t seed, uint32_t step, uint64_t global_total, uint64_t total_slots, uint64_t* stride, uint64_t* offset) {
  BB_REQUIRE(stride && offset && total_slots > 0);
  *stride = global_total / total_slots;
  *offset = bb200::mulhi64(bb200::counter_draw(seed, 0, step, bb200::kStreamSystematic).a, *stride);
  return BB200_OK;
}
int bb200_estimate_from_moments(const double moments[9], const double pivot_xy[2], bb200_estimate* out) {
  BB_REQUIRE(moments && pivot_xy && out);
  // Pure host arithmetic (estimation.hpp:436-475 from raw moments); needs no device.
  bb200::Filter::estimate_from_moments_static(moments, pivot_xy, out);
  return BB200_OK;
}
int bb200_filter_set_stream(bb200_filter* f, void* cuda_stream) {
  BB_REQUIRE(f);
  return guarded(f->impl, [&] { return f->impl.set_stream(cuda_stream); });
}
int bb200_filter_enqueue_propagate_reweight(bb200_filter* f, const bb200_motion_sampling* s, uint32_t step, const double* points_xy, uint64_t n_points) {
  BB_REQUIRE(f && s && points_xy);
  return guarded(f->impl, [&] { return f->impl.enqueue_propagate_reweight(s, step, points_xy, n_points); });
}
int bb200_filter_enqueue_build_cdf(bb200_filter* f) {
  BB_REQUIRE(f);
  return guarded(f->impl, [&] { return f->impl.enqueue_build_cdf(); });
}
int bb200_filter_enqueue_resample_range(bb200_filter* f, const bb200_resample_opts* o, uint64_t global_total, uint64_t cdf_offset, uint64_t slot_begin,
                                        uint64_t slot_end) {
  BB_REQUIRE(f && o);
  return guarded(f->impl, [&] { return f->impl.enqueue_resample_range(*o, global_total, cdf_offset, slot_begin, slot_end); });
}
int bb200_filter_enqueue_adopt(bb200_filter* f, uint64_t n) {
  BB_REQUIRE(f);
  return guarded(f->impl, [&] { return f->impl.enqueue_adopt(n); });
}
int bb200_filter_enqueue_moments(bb200_filter* f, const double pivot_xy[2]) {
  BB_REQUIRE(f && pivot_xy);
  return guarded(f->impl, [&] { return f->impl.enqueue_moments(pivot_xy); });
}
int bb200_filter_ipc_handles(bb200_filter* f, void* out128) {
  BB_REQUIRE(f && out128);
  return guarded(f->impl, [&] { return f->impl.ipc_handles(out128); });
}
int bb200_filter_open_peers(bb200_filter* f, int world, int rank, const void* handles) {
  BB_REQUIRE(f && handles);
  return guarded(f->impl, [&] { return f->impl.open_peers(world, rank, handles); });
}
int bb200_filter_enqueue_resample_push(bb200_filter* f, const bb200_resample_opts* o, uint64_t global_total, uint64_t cdf_offset, uint64_t slot_begin,
                                       uint64_t slot_end, uint64_t shard, const double pivot_xy[2]) {
  BB_REQUIRE(f && o && pivot_xy && slot_end >= slot_begin && shard > 0);
  return guarded(f->impl, [&] { return f->impl.enqueue_resample_push(*o, global_total, cdf_offset, slot_begin, slot_end, shard, pivot_xy); });
}
int bb200_filter_enqueue_resample_push_device(bb200_filter* f, const bb200_resample_opts* o, const uint64_t* rank_totals_device, int rank, int world,
                                              uint64_t shard, const double pivot_xy[2]) {
  BB_REQUIRE(f && o && rank_totals_device && pivot_xy && shard > 0);
  return guarded(f->impl, [&] { return f->impl.enqueue_resample_push_device(*o, rank_totals_device, rank, world, shard, pivot_xy); });
}
int bb200_filter_enqueue_reduce_moments(bb200_filter* f) {
  BB_REQUIRE(f);
  return guarded(f->impl, [&] { return f->impl.enqueue_reduce_moments(); });
}
int bb200_filter_enqueue_flip_adopt(bb200_filter* f, uint64_t n) {
  BB_REQUIRE(f);
  return guarded(f->impl, [&] { return f->impl.enqueue_flip_adopt(n); });
}
int bb200_filter_ancestors(bb200_filter* f, int64_t* out, uint64_t capacity) {
  BB_REQUIRE(f && out);
  return guarded(f->impl, [&] { return f->impl.ancestors(out, capacity); });
}
int bb200_filter_cdf(bb200_filter* f, uint64_t* out, uint64_t capacity) {
  BB_REQUIRE(f && out);
  return guarded(f->impl, [&] { return f->impl.cdf(out, capacity); });
}
int bb200_filter_estimate(bb200_filter* f, bb200_estimate* out) {
  BB_REQUIRE(f && out);
  return guarded(f->impl, [&] { return f->impl.estimate(out); });
}
void bb200_cluster_param_default(bb200_cluster_param* p) {
  if (p == nullptr) return;
  p->linear_hash_resolution = 0.20;
  p->angular_hash_resolution = 0.524;
  p->weight_cap_percentile = 0.90;
}
int bb200_filter_cluster_estimate(bb200_filter* f, const bb200_cluster_param* p, bb200_estimate* out, uint32_t* cluster_ids, uint64_t ids_capacity,
                                  uint32_t* n_cells, uint32_t* n_clusters) {
  BB_REQUIRE(f && p);
  return guarded(f->impl, [&] { return f->impl.cluster_estimate(*p, out, cluster_ids, ids_capacity, n_cells, n_clusters); });
}
int bb200_cluster_select_host(const bb200_cluster_cell* cells, uint64_t n_cells, uint64_t n_particles, const bb200_cluster_param* p,
                              uint32_t* cluster_of_cell, uint32_t* n_clusters, int* found, uint32_t* best, double moments_out[9]) {
  if ((cells == nullptr && n_cells > 0) || p == nullptr || !(p->linear_hash_resolution > 0.0) || !(p->angular_hash_resolution > 0.0) ||
      !(p->weight_cap_percentile >= 0.0) || !(p->weight_cap_percentile < 1.0))
    return BB200_ERR_INVALID_ARGUMENT;
  static_assert(sizeof(bb200_cluster_cell) == sizeof(bb200::HostCell), "public and internal cell records must agree");
  return guarded_create([&] {
    const bb200::ClusterSelection sel = bb200::select_cluster(reinterpret_cast<const bb200::HostCell*>(cells), n_cells, n_particles,
                                                              p->linear_hash_resolution, p->angular_hash_resolution, p->weight_cap_percentile);
    if (cluster_of_cell != nullptr)
      for (uint64_t k = 0; k < n_cells; ++k) cluster_of_cell[k] = sel.cluster_of_cell[k];
    if (n_clusters != nullptr) *n_clusters = sel.clusters;
    if (found != nullptr) *found = sel.found ? 1 : 0;
    if (best != nullptr) *best = sel.best;
    if (moments_out != nullptr)
      for (int j = 0; j < 9; ++j) moments_out[j] = sel.moments[j];
    return static_cast<int>(BB200_OK);
  });
}
int bb200_filter_particle_histogram(bb200_filter* f, double linear_resolution, double angular_resolution, bb200_cluster_cell* bins, uint64_t capacity,
                                    uint64_t* n_bins, double* max_bin_weight) {
  BB_REQUIRE(f);
  return guarded(f->impl, [&] { return f->impl.particle_histogram(linear_resolution, angular_resolution, bins, capacity, n_bins, max_bin_weight); });
}
int bb200_filter_sample_states(bb200_filter* f, uint64_t count, uint32_t step, double* states_out) {
  BB_REQUIRE(f);
  return guarded(f->impl, [&] { return f->impl.sample_states(count, step, states_out); });
}

namespace {
// detail::alphaHueToRGBA (beluga_ros/particle_cloud.hpp:56-70): single-precision HSV -> RGB for V = S = 1.
void alpha_hue_to_rgba(float hue, float alpha, float rgba[4]) {
  const float kr = std::fmod(5.0F + hue / 60.0F, 6.0F);
  const float kg = std::fmod(3.0F + hue / 60.0F, 6.0F);
  const float kb = std::fmod(1.0F + hue / 60.0F, 6.0F);
  rgba[0] = 1.0F - 1.0F * std::max(0.0F, std::min({kr, 4.0F - kr, 1.0F}));
  rgba[1] = 1.0F - 1.0F * std::max(0.0F, std::min({kg, 4.0F - kg, 1.0F}));
  rgba[2] = 1.0F - 1.0F * std::max(0.0F, std::min({kb, 4.0F - kb, 1.0F}));
  rgba[3] = alpha;
}
}  // namespace

int bb200_particle_cloud_markers(const bb200_cluster_cell* bins, uint64_t n_bins, bb200_marker_vertex* bodies, bb200_marker_vertex* heads,
                                 double* body_scale_x) {
  BB_REQUIRE((bins || n_bins == 0) && (bodies || n_bins == 0) && (heads || n_bins == 0) && body_scale_x);
  // beluga_ros/particle_cloud.hpp:212-294: arrows as a line list (bodies) and a triangle list (heads), sizes and
  // colours scaled by the bin weight relative to the heaviest bin.
  const double kArrowBodyLength = 0.5, kArrowHeadLength = 0.1, kArrowHeadWidth = kArrowHeadLength / 5.0;
  const double kArrowLength = kArrowBodyLength + kArrowHeadLength;
  double max_bin_weight = 1e-3;
  for (uint64_t k = 0; k < n_bins; ++k) max_bin_weight = bins[k].weight > max_bin_weight ? bins[k].weight : max_bin_weight;
  double min_scale_factor = 1.0;
  for (uint64_t k = 0; k < n_bins; ++k) {
    const double* st = bins[k].representative;  // {cos, sin, x, y}
    const double scale_factor = std::max(bins[k].weight / max_bin_weight, 1e-1);
    if (scale_factor < min_scale_factor) min_scale_factor = scale_factor;
    float rgba[4];
    alpha_hue_to_rgba(static_cast<float>((1.0 - scale_factor) * 270.0), static_cast<float>(0.25 + 0.75 * scale_factor), rgba);
    auto put = [&](bb200_marker_vertex& v, double lx, double ly) {  // state * (scale_factor * local point)
      const double px = scale_factor * lx, py = scale_factor * ly;
      v.x = (st[0] * px - st[1] * py) + st[2];
      v.y = (st[1] * px + st[0] * py) + st[3];
      v.z = 0.0;
      v.r = rgba[0], v.g = rgba[1], v.b = rgba[2], v.a = rgba[3];
    };
    put(bodies[2 * k], 0.0, 0.0);
    put(bodies[2 * k + 1], kArrowBodyLength, 0.0);
    put(heads[3 * k], kArrowBodyLength, kArrowHeadWidth / 2.0);
    put(heads[3 * k + 1], kArrowBodyLength, -(kArrowHeadWidth / 2.0));
    put(heads[3 * k + 2], kArrowLength, 0.0);
  }
  *body_scale_x = static_cast<double>(min_scale_factor * kArrowHeadWidth) * 0.8;
  return BB200_OK;
}

int bb200_likelihood_field_to_occupancy(const float* field, uint64_t n, int8_t* out) {
  BB_REQUIRE((field && out) || n == 0);
  if (n == 0) return BB200_OK;
  // beluga_ros/likelihood_field.hpp:44-79
  float min_val = field[0], max_val = field[0];
  for (uint64_t i = 1; i < n; ++i) {
    if (field[i] < min_val) min_val = field[i];
    if (max_val < field[i]) max_val = field[i];
  }
  const float range = max_val - min_val;
  if (range <= std::numeric_limits<float>::epsilon()) {
    for (uint64_t i = 0; i < n; ++i) out[i] = 0;
    return BB200_OK;
  }
  for (uint64_t i = 0; i < n; ++i) {
    const float normalized = (field[i] - min_val) / range;
    out[i] = static_cast<int8_t>(normalized * 100.0f);
  }
  return BB200_OK;
}

int bb200_filter_moments(bb200_filter* f, const double pivot_xy[2], double out[9]) {
  BB_REQUIRE(f && pivot_xy && out);
  return guarded(f->impl, [&] { return f->impl.moments(pivot_xy, out); });
}
int bb200_filter_set_timing(bb200_filter* f, int enabled) {
  BB_REQUIRE(f);
  f->impl.set_timing(enabled != 0);
  return BB200_OK;
}
int bb200_filter_clear_timings(bb200_filter* f) {
  BB_REQUIRE(f);
  f->impl.clear_timings();
  return BB200_OK;
}
int bb200_filter_last_timings(const bb200_filter* f, const char** names, float* ms, int capacity) {
  if (f == nullptr) return 0;
  return guarded(f->impl, [&] { return f->impl.last_timings(names, ms, capacity); });
}
uint64_t bb200_filter_launch_count(const bb200_filter* f) { return f != nullptr ? f->impl.launch_count() : 0; }
int bb200_filter_synchronize(bb200_filter* f) {
  BB_REQUIRE(f);
  return guarded(f->impl, [&] { return f->impl.synchronize(); });
}
int bb200_filter_device_pointer(bb200_filter* f, int which, void** ptr, uint64_t* bytes) {
  BB_REQUIRE(f && ptr && bytes);
  return guarded(f->impl, [&] { return f->impl.device_pointer(which, ptr, bytes); });
}

// ---- amcl ---------------------------------------------------------------------------------------

int bb200_amcl_create_with_motion(const bb200_amcl_param* p, const bb200_motion_param* motion, bb200_amcl** out) {
  if (p == nullptr || motion == nullptr || out == nullptr) {
    g_create_error = "null argument";
    return BB200_ERR_INVALID_ARGUMENT;
  }
  *out = nullptr;
  if (motion->model < BB200_MOTION_DIFFERENTIAL || motion->model > BB200_MOTION_STATIONARY) {
    g_create_error = "unknown motion model";
    return BB200_ERR_INVALID_ARGUMENT;
  }
  if (p->shard_capacity != 0 && p->shard_first_index + p->shard_capacity > p->max_particles) {
    g_create_error = "a shard must lie inside [0, max_particles)";
    return BB200_ERR_INVALID_ARGUMENT;
  }
  if (p->max_particles == 0 || p->min_particles > p->max_particles) {
    g_create_error = "need 0 < min_particles <= max_particles";
    return BB200_ERR_INVALID_ARGUMENT;
  }
  if (!(p->alpha_slow >= 0.0) || !(p->alpha_slow <= p->alpha_fast)) {  // thrun_recovery_probability_estimator.hpp:49-50
    g_create_error = "need 0 <= alpha_slow <= alpha_fast";
    return BB200_ERR_INVALID_ARGUMENT;
  }
  return guarded_create([&] {
    bb200_amcl* a = new bb200_amcl(*p, *motion);
    if (!a->impl.ok()) {
      g_create_error = a->impl.last_error();
      const int st = a->impl.create_status();
      delete a;
      return st;
    }
    a->filter_view = reinterpret_cast<bb200_filter*>(&a->impl.filter());
    *out = a;
    return static_cast<int>(BB200_OK);
  });
}
int bb200_amcl_create(const bb200_amcl_param* p, const bb200_diff_drive_param* motion, bb200_amcl** out) {
  if (motion == nullptr) {
    g_create_error = "null argument";
    return BB200_ERR_INVALID_ARGUMENT;
  }
  const bb200_motion_param m{BB200_MOTION_DIFFERENTIAL, motion->rotation_noise_from_rotation, motion->rotation_noise_from_translation,
                             motion->translation_noise_from_translation, motion->translation_noise_from_rotation, 0.0, motion->distance_threshold};
  return bb200_amcl_create_with_motion(p, &m, out);
}
void bb200_amcl_destroy(bb200_amcl* a) { delete a; }
const char* bb200_amcl_last_error(const bb200_amcl* a) { return a != nullptr ? a->impl.last_error() : "null amcl"; }
bb200_filter* bb200_amcl_filter(bb200_amcl* a) { return a != nullptr ? a->filter_view : nullptr; }
int bb200_amcl_initialize(bb200_amcl* a, const double mean_xytheta[3], const double cov[9]) {
  BB_REQUIRE(a && mean_xytheta && cov);
  return guarded(a->impl, [&] { return a->impl.initialize(mean_xytheta, cov); });
}
int bb200_amcl_initialize_from_map(bb200_amcl* a) {
  BB_REQUIRE(a);
  return guarded(a->impl, [&] { return a->impl.initialize_from_map(); });
}
int bb200_amcl_initialize_states(bb200_amcl* a, const double* states, const double* weights, uint64_t n) {
  BB_REQUIRE(a);
  return guarded(a->impl, [&] { return a->impl.initialize_states(states, weights, n); });
}
void bb200_amcl_force_update(bb200_amcl* a) {
  if (a != nullptr) a->impl.force_update();
}
int bb200_amcl_update(bb200_amcl* a, const double control_pose[4], const double* points_xy, uint64_t n_points, bb200_update_result* out) {
  BB_REQUIRE(a && control_pose && out && (points_xy || n_points == 0));
  static const double kNoPoints[2] = {0.0, 0.0};
  return guarded(a->impl, [&] { return a->impl.update(control_pose, points_xy != nullptr ? points_xy : kNoPoints, n_points, out); });
}
int bb200_take_evenly_indices(uint64_t size, uint64_t count, uint64_t* indices, uint64_t capacity, uint64_t* n_indices) {
  BB_REQUIRE(indices && n_indices);
  // take_evenly_view::size() and compute_offset() (views/take_evenly.hpp:47-57,118-145)
  uint64_t kept = size == 0 ? 0 : (count > size ? size : count);
  if (kept > capacity) return BB200_ERR_CAPACITY;
  for (uint64_t pos = 0; pos < kept; ++pos) {
    uint64_t idx;
    if (count > size) {
      idx = pos;
    } else if (pos == 0) {
      idx = 0;
    } else {  // count >= 2 here: pos < kept <= count
      const uint64_t a = pos * (size - 1), b = count - 1;
      idx = a / b + ((a % b == 0) ? 0 : 1);
    }
    indices[pos] = idx;
  }
  *n_indices = kept;
  return BB200_OK;
}

int bb200_scan_to_points(const bb200_laser_scan* scan, double* points_xy, uint64_t capacity, uint64_t* n_points) {
  BB_REQUIRE(scan && points_xy && n_points && (scan->ranges || scan->n_ranges == 0));
  const uint64_t size = scan->n_ranges;
  const uint64_t count = scan->max_beams == 0 ? size : scan->max_beams;
  const uint64_t kept = size == 0 ? 0 : (count > size ? size : count);
  uint64_t n = 0;
  for (uint64_t pos = 0; pos < kept; ++pos) {
    uint64_t i = pos;
    if (count <= size && pos != 0) {
      const uint64_t a = pos * (size - 1), b = count - 1;
      i = a / b + ((a % b == 0) ? 0 : 1);
    }
    const double range = static_cast<double>(scan->ranges[i]);
    // beluga_ros/laser_scan.hpp:73-74: float arithmetic, then widened.
    const double theta = static_cast<double>(scan->angle_min + static_cast<float>(static_cast<int>(i)) * scan->angle_increment);
    if (std::isnan(range) || !(range >= scan->min_range) || !(range <= scan->max_range)) continue;  // laser_scan.hpp:80-84
    const double x = range * std::cos(theta), y = range * std::sin(theta);                           // laser_scan.hpp:66-69
    double px = x, py = y;
    if (scan->laser_origin != nullptr) {  // origin * (x, y, 0), keep (x, y)  (beluga_ros/src/amcl.cpp:59-61)
      const double* m = scan->laser_origin;
      px = m[0] * x + m[1] * y + m[3];
      py = m[4] * x + m[5] * y + m[7];
    }
    if (n >= capacity) return BB200_ERR_CAPACITY;
    points_xy[2 * n] = px;
    points_xy[2 * n + 1] = py;
    ++n;
  }
  *n_points = n;
  return BB200_OK;
}

int bb200_amcl_update_scan(bb200_amcl* a, const double control_pose[4], const bb200_laser_scan* scan, bb200_update_result* out) {
  BB_REQUIRE(a && control_pose && scan && out);
  return guarded(a->impl, [&] {
    std::vector<double> points(2 * (scan->n_ranges + 1));
    uint64_t n = 0;
    const int st = bb200_scan_to_points(scan, points.data(), scan->n_ranges + 1, &n);
    if (st != BB200_OK) return st;
    return bb200_amcl_update(a, control_pose, points.data(), n, out);
  });
}

// ---- sharded filters ---------------------------------------------------------------------------------

int bb200_amcl_export_shard(bb200_amcl* a, void* out256) {
  BB_REQUIRE(a && out256);
  return guarded(a->impl, [&] { return a->impl.filter().export_shard(out256); });
}
int bb200_amcl_join_shards(bb200_amcl* a, int world, int rank, const void* handles) {
  BB_REQUIRE(a && handles);
  return guarded(a->impl, [&] { return a->impl.filter().join_shards_ipc(world, rank, handles); });
}
int bb200_amcl_leave_shards(bb200_amcl* a) {
  BB_REQUIRE(a);
  return guarded(a->impl, [&] { return a->impl.filter().leave_shards(); });
}
int bb200_filter_export_shard(bb200_filter* f, void* out256) {
  BB_REQUIRE(f && out256);
  return guarded(f->impl, [&] { return f->impl.export_shard(out256); });
}
int bb200_filter_join_shards(bb200_filter* f, int world, int rank, const void* handles) {
  BB_REQUIRE(f && handles);
  return guarded(f->impl, [&] { return f->impl.join_shards_ipc(world, rank, handles); });
}

int bb200_sharded_amcl_create(const bb200_amcl_param* p, const bb200_motion_param* motion, int n_shards, const int* devices, bb200_sharded_amcl** out) {
  if (p == nullptr || motion == nullptr || out == nullptr || devices == nullptr) {
    g_create_error = "null argument";
    return BB200_ERR_INVALID_ARGUMENT;
  }
  *out = nullptr;
  if (n_shards < 1 || n_shards > bb200::kMaxShards || p->max_particles == 0 || p->max_particles % static_cast<uint64_t>(n_shards) != 0) {
    g_create_error = "a filter splits into 1..8 equal shards: max_particles must be a multiple of n_shards";
    return BB200_ERR_INVALID_ARGUMENT;
  }
  return guarded_create([&] {
    auto group = std::make_unique<bb200_sharded_amcl>();
    const uint64_t shard = p->max_particles / static_cast<uint64_t>(n_shards);
    std::vector<bb200::Filter*> filters;
    for (int r = 0; r < n_shards; ++r) {
      bb200_amcl_param q = *p;
      q.device = devices[r];
      q.shard_capacity = shard;
      q.shard_first_index = static_cast<uint64_t>(r) * shard;
      bb200_amcl* a = nullptr;
      const int st = bb200_amcl_create_with_motion(&q, motion, &a);
      if (st != BB200_OK) return st;
      group->ranks.push_back(a);
      filters.push_back(&a->impl.filter());
    }
    const int st = bb200::Filter::join_shards_local(filters.data(), n_shards);
    if (st != BB200_OK) {
      g_create_error = "joining the shards failed (peer access between the devices?)";
      for (bb200::Filter* f : filters)
        if (f->last_error()[0] != '\0') g_create_error = f->last_error();
      return st;
    }
    *out = group.release();
    return static_cast<int>(BB200_OK);
  });
}
void bb200_sharded_amcl_destroy(bb200_sharded_amcl* g) { delete g; }
const char* bb200_sharded_amcl_last_error(const bb200_sharded_amcl* g) {
  if (g == nullptr) return "null sharded amcl";
  if (!g->error.empty()) return g->error.c_str();
  for (const bb200_amcl* a : g->ranks)
    if (a->impl.last_error()[0] != '\0') return a->impl.last_error();
  return "";
}
int bb200_sharded_amcl_shards(const bb200_sharded_amcl* g) { return g != nullptr ? static_cast<int>(g->ranks.size()) : 0; }
bb200_amcl* bb200_sharded_amcl_shard(bb200_sharded_amcl* g, int rank) {
  return (g != nullptr && rank >= 0 && rank < static_cast<int>(g->ranks.size())) ? g->ranks[static_cast<size_t>(rank)] : nullptr;
}
#define BB_EACH_SHARD(call)                                   \
  BB_REQUIRE(g);                                              \
  g->error.clear();                                           \
  return guarded(*g, [&] {                                    \
    for (bb200_amcl* a : g->ranks) {                          \
      const int st = (call);                                  \
      if (st != BB200_OK) return st;                          \
    }                                                         \
    return static_cast<int>(BB200_OK);                        \
  })
int bb200_sharded_amcl_set_likelihood_field_map(bb200_sharded_amcl* g, const bb200_likelihood_field_param* p, const bb200_occupancy_grid* grid, int prob) {
  BB_REQUIRE(p && grid);
  BB_EACH_SHARD(a->impl.filter().set_likelihood_field_map(*p, *grid, prob != 0));
}
int bb200_sharded_amcl_set_beam_map(bb200_sharded_amcl* g, const bb200_beam_param* p, const bb200_occupancy_grid* grid) {
  BB_REQUIRE(p && grid);
  BB_EACH_SHARD(a->impl.filter().set_beam_map(*p, *grid));
}
int bb200_sharded_amcl_initialize(bb200_sharded_amcl* g, const double mean_xytheta[3], const double cov[9]) {
  BB_REQUIRE(mean_xytheta && cov);
  BB_EACH_SHARD(a->impl.initialize(mean_xytheta, cov));
}
int bb200_sharded_amcl_initialize_from_map(bb200_sharded_amcl* g) { BB_EACH_SHARD(a->impl.initialize_from_map()); }
#undef BB_EACH_SHARD
void bb200_sharded_amcl_force_update(bb200_sharded_amcl* g) {
  if (g != nullptr)
    for (bb200_amcl* a : g->ranks) a->impl.force_update();
}
int bb200_sharded_amcl_update(bb200_sharded_amcl* g, const double control_pose[4], const double* points_xy, uint64_t n_points, bb200_update_result* out) {
  BB_REQUIRE(g && control_pose && out && (points_xy || n_points == 0));
  static const double kNoPoints[2] = {0.0, 0.0};
  g->error.clear();
  return guarded(*g, [&] {
    std::vector<Amcl*> ranks;
    for (bb200_amcl* a : g->ranks) ranks.push_back(&a->impl);
    return Amcl::update_group(ranks.data(), static_cast<int>(ranks.size()), control_pose, points_xy != nullptr ? points_xy : kNoPoints, n_points, out);
  });
}
int bb200_cluster_merge_host(const bb200_cluster_cell* const* shard_cells, const uint64_t* shard_counts, const uint64_t* shard_first_index, int shards,
                             bb200_cluster_cell* merged, uint64_t capacity, uint64_t* n_merged) {
  BB_REQUIRE(shard_cells && shard_counts && n_merged && shards >= 1);
  return guarded_create([&] {
    // The reference's cluster map is filled by ONE pass over the particles in order (cluster_based_estimation.hpp:141-161).
    // Shards hold contiguous index ranges, so going through the shards in rank order -- and through each shard's cells in
    // ITS first-occurrence order -- visits the cells in the global first-occurrence order; a cell met again adds its count,
    // weight and raw moments to the record opened by the lower rank (whose representative is the globally first particle).
    std::unordered_map<uint64_t, uint64_t> slot_of;
    uint64_t n = 0;
    for (int r = 0; r < shards; ++r) {
      for (uint64_t k = 0; k < shard_counts[r]; ++k) {
        const bb200_cluster_cell& c = shard_cells[r][k];
        const auto it = slot_of.find(c.hash);
        if (it == slot_of.end()) {
          if (merged != nullptr) {
            if (n >= capacity) return static_cast<int>(BB200_ERR_CAPACITY);
            merged[n] = c;
            merged[n].first_index = static_cast<uint32_t>(c.first_index + (shard_first_index != nullptr ? shard_first_index[r] : 0));
          }
          slot_of.emplace(c.hash, n++);
        } else if (merged != nullptr) {
          bb200_cluster_cell& m = merged[it->second];
          m.count += c.count;
          m.weight += c.weight;
          for (int j = 0; j < 9; ++j) m.moments[j] += c.moments[j];
        }
      }
    }
    *n_merged = n;
    return static_cast<int>(BB200_OK);
  });
}

int bb200_sharded_amcl_cluster_estimate(bb200_sharded_amcl* g, const bb200_cluster_param* p, bb200_estimate* out, uint32_t* n_cells,
                                        uint32_t* n_clusters) {
  BB_REQUIRE(g && p && out);
  g->error.clear();
  return guarded(*g, [&] {
    const int shards = static_cast<int>(g->ranks.size());
    std::vector<std::vector<bb200_cluster_cell>> cells(static_cast<size_t>(shards));
    std::vector<const bb200_cluster_cell*> ptrs;
    std::vector<uint64_t> counts, firsts;
    uint64_t total_cells = 0, n_particles = 0;
    for (int r = 0; r < shards; ++r) {
      Filter& f = g->ranks[static_cast<size_t>(r)]->impl.filter();
      uint64_t n = 0;
      double top = 0.0;
      if (f.size() > 0) {
        int st = f.particle_histogram(p->linear_hash_resolution, p->angular_hash_resolution, nullptr, 0, &n, &top);
        if (st != BB200_OK) return st;
        cells[static_cast<size_t>(r)].resize(n);
        st = f.particle_histogram(p->linear_hash_resolution, p->angular_hash_resolution, cells[static_cast<size_t>(r)].data(), n, &n, &top);
        if (st != BB200_OK) return st;
      }
      ptrs.push_back(cells[static_cast<size_t>(r)].data());
      counts.push_back(n);
      firsts.push_back(f.first_index());
      total_cells += n;
      n_particles += f.size();
    }
    if (n_particles == 0) {
      g->error = "no particles";
      return static_cast<int>(BB200_ERR_STATE);
    }
    std::vector<bb200_cluster_cell> merged(total_cells);
    uint64_t n_merged = 0;
    const int st = bb200_cluster_merge_host(ptrs.data(), counts.data(), firsts.data(), shards, merged.data(), merged.size(), &n_merged);
    if (st != BB200_OK) return st;
    static_assert(sizeof(bb200_cluster_cell) == sizeof(bb200::HostCell), "public and internal cell records must agree");
    const bb200::ClusterSelection sel = bb200::select_cluster(reinterpret_cast<const bb200::HostCell*>(merged.data()), n_merged, n_particles,
                                                              p->linear_hash_resolution, p->angular_hash_resolution, p->weight_cap_percentile);
    bb200::Filter::estimate_from_moments_static(sel.moments, g->ranks[0]->impl.filter().pivot(), out);  // every shard accumulates about the same pivot
    if (n_cells != nullptr) *n_cells = static_cast<uint32_t>(n_merged);
    if (n_clusters != nullptr) *n_clusters = sel.clusters;
    return static_cast<int>(BB200_OK);
  });
}

int bb200_sharded_amcl_get_particles(bb200_sharded_amcl* g, double* states, double* weights, uint64_t capacity) {
  BB_REQUIRE(g);
  g->error.clear();
  return guarded(*g, [&] {
    uint64_t done = 0;
    for (bb200_amcl* a : g->ranks) {  // rank order is global particle order
      Filter& f = a->impl.filter();
      const uint64_t n = std::min<uint64_t>(f.size(), capacity - done);
      const int st = f.get_particles(states != nullptr ? states + 4 * done : nullptr, weights != nullptr ? weights + done : nullptr, n);
      if (st != BB200_OK) return st;
      done += n;
      if (done >= capacity) break;
    }
    return static_cast<int>(BB200_OK);
  });
}

int bb200_amcl_plan_update(bb200_amcl* a, const double control_pose[4], bb200_step_plan* plan) {
  BB_REQUIRE(a && control_pose && plan);
  return guarded(a->impl, [&] { return a->impl.plan_update(control_pose, plan); });
}
void bb200_amcl_commit_update(bb200_amcl* a, int resampled, double random_state_probability) {
  if (a != nullptr) a->impl.commit_update(resampled, random_state_probability);
}
int bb200_motion_sampling_from_control(const bb200_motion_param* p, const double pose[4], const double previous_pose[4], bb200_motion_sampling* out) {
  BB_REQUIRE(p && pose && previous_pose && out);
  *out = bb200::motion_sampling(*p, bb200::Pose2{pose[0], pose[1], pose[2], pose[3]},
                                bb200::Pose2{previous_pose[0], previous_pose[1], previous_pose[2], previous_pose[3]});
  return BB200_OK;
}
int bb200_diff_drive_sampling_from_control(const bb200_diff_drive_param* p, const double pose[4], const double previous_pose[4], bb200_diff_drive_sampling* out) {
  BB_REQUIRE(p && pose && previous_pose && out);
  *out = bb200::diff_drive_sampling(*p, bb200::Pose2{pose[0], pose[1], pose[2], pose[3]},
                                    bb200::Pose2{previous_pose[0], previous_pose[1], previous_pose[2], previous_pose[3]});
  return BB200_OK;
}

}  // extern "C"
