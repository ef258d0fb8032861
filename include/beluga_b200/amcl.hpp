// beluga_b200/amcl.hpp -- beluga::Amcl shaped driver over the C ABI.
//
// Mirrors beluga::Amcl<MotionModel, SensorModel, ...> (beluga/algorithm/amcl_core.hpp:74-233):
// same constructor argument order (motion model, sensor model, [random state generator, spatial
// hasher,] params), same public methods -- particles(), initialize(pose, covariance),
// update_map(map), update(control_action, measurement) -> std::optional<std::pair<pose, covariance>>,
// force_update() -- and the same AmclParams members and defaults (:34-55).  The particle set lives
// on the GPU; particles() copies it out on demand.
//
// What differs from the template: the random state generator is always the map's free-cell uniform
// distribution (what beluga_ros::Amcl passes, beluga_ros/include/beluga_ros/amcl.hpp:269), the
// spatial hasher is spatial_hash<SE2d> with the resolutions in AmclParams, and the execution
// policy argument is accepted and ignored (everything per-particle runs on the device).
#pragma once

#include <execution>
#include <optional>
#include <utility>

#include "models.hpp"

namespace beluga_b200 {

/// beluga::AmclParams (amcl_core.hpp:34-55) plus the knobs the drop-in backend adds.
struct AmclParams {
  double update_min_d = 0.25;
  double update_min_a = 0.2;
  std::size_t resample_interval = 1UL;
  bool selective_resampling = false;
  std::size_t min_particles = 500UL;
  std::size_t max_particles = 2000UL;
  double alpha_slow = 0.001;
  double alpha_fast = 0.1;
  double kld_epsilon = 0.05;
  double kld_z = 3.0;
  // spatial_hash<SE2d>{x, y, theta} resolutions (beluga_ros/include/beluga_ros/amcl.hpp:91-97)
  double spatial_resolution_x = 0.5;
  double spatial_resolution_y = 0.5;
  double spatial_resolution_theta = 10.0 * 3.14159265358979323846 / 180.0;
  // backend
  bb200_resample_scheme resample_scheme = BB200_RESAMPLE_MULTINOMIAL;
  std::uint64_t seed = 0;  // the reference's engine is auto-seeded and cannot be set (amcl_core.hpp uses get_random_engine())
  int device = 0;
};

/// Particle set copy: the two vectors of beluga::TupleVector<std::tuple<SE2d, Weight>>.
struct ParticleSet {
  std::vector<SE2d> states;
  std::vector<double> weights;
  [[nodiscard]] std::size_t size() const { return states.size(); }
  [[nodiscard]] bool empty() const { return states.empty(); }
};

template <class MotionModel, class SensorModel, class ExecutionPolicy = std::execution::sequenced_policy>
class Amcl {
 public:
  using state_type = typename SensorModel::state_type;
  using measurement_type = typename SensorModel::measurement_type;
  using map_type = typename SensorModel::map_type;
  using estimation_type = std::pair<state_type, Matrix3d>;

  Amcl(MotionModel motion_model, SensorModel sensor_model, const AmclParams& params = AmclParams{}, ExecutionPolicy = ExecutionPolicy{})
      : params_{params}, motion_model_{std::move(motion_model)}, sensor_model_{std::move(sensor_model)} {
    bb200_amcl_param p{};
    p.update_min_d = params.update_min_d;
    p.update_min_a = params.update_min_a;
    p.resample_interval = params.resample_interval;
    p.selective_resampling = params.selective_resampling ? 1 : 0;
    p.min_particles = params.min_particles;
    p.max_particles = params.max_particles;
    p.alpha_slow = params.alpha_slow;
    p.alpha_fast = params.alpha_fast;
    p.kld_epsilon = params.kld_epsilon;
    p.kld_z = params.kld_z;
    p.spatial_resolution[0] = params.spatial_resolution_x;
    p.spatial_resolution[1] = params.spatial_resolution_y;
    p.spatial_resolution[2] = params.spatial_resolution_theta;
    p.resample_scheme = params.resample_scheme;
    p.seed = params.seed;
    p.device = params.device;
    const bb200_motion_param m = motion_model_.c_motion_param();
    const int st = bb200_amcl_create_with_motion(&p, &m, &handle_);
    if (st != BB200_OK) throw Error(st, bb200_create_error());
    check(sensor_model_.attach(bb200_amcl_filter(handle_)));
  }
  ~Amcl() { bb200_amcl_destroy(handle_); }
  Amcl(const Amcl&) = delete;
  Amcl& operator=(const Amcl&) = delete;
  Amcl(Amcl&& other) noexcept
      : params_{other.params_}, motion_model_{std::move(other.motion_model_)}, sensor_model_{std::move(other.sensor_model_)}, handle_{other.handle_} {
    other.handle_ = nullptr;
  }

  /// amcl_core.hpp:128 -- copies the particle set back from the device.
  [[nodiscard]] ParticleSet particles() const {
    bb200_filter* f = bb200_amcl_filter(handle_);
    std::uint64_t n = 0;
    check(bb200_filter_size(f, &n));
    ParticleSet out;
    out.states.resize(n);
    out.weights.resize(n);
    static_assert(sizeof(SE2d) == 4 * sizeof(double), "SE2d must be four packed doubles");
    if (n > 0) check(bb200_filter_get_particles(f, reinterpret_cast<double*>(out.states.data()), out.weights.data(), n));
    return out;
  }

  /// amcl_core.hpp:145-147 -- throws like MultivariateNormalDistribution on an invalid covariance.
  void initialize(const state_type& pose, const Matrix3d& covariance) {
    const double mean[3] = {pose.x(), pose.y(), pose.theta()};
    check(bb200_amcl_initialize(handle_, mean, covariance.data()));
  }
  /// amcl_core.hpp:131-137 with an explicit state list instead of a distribution.
  /// beluga_ros::Amcl::initialize_from_map() (beluga_ros/include/beluga_ros/amcl.hpp:209): max_particles samples of the
  /// map distribution -- uniform over the free cells of the sensor model's map.
  void initialize_from_map() { check(bb200_amcl_initialize_from_map(handle_)); }
  void initialize(const std::vector<state_type>& states) {
    check(bb200_amcl_initialize_states(handle_, states.empty() ? nullptr : reinterpret_cast<const double*>(states.data()), nullptr, states.size()));
  }

  /// amcl_core.hpp:150
  void update_map(const map_type& map) {
    sensor_model_.update_map(map);
    check(sensor_model_.attach(bb200_amcl_filter(handle_)));
  }

  /// amcl_core.hpp:165-201
  auto update(const state_type& control_action, measurement_type measurement) -> std::optional<estimation_type> {
    const measurement_token token = sensor_model_(std::move(measurement));
    static_assert(sizeof(std::pair<double, double>) == 2 * sizeof(double), "measurement points must be packed");
    bb200_update_result r{};
    check(bb200_amcl_update(handle_, control_action.data(), token.points.empty() ? nullptr : &token.points.front().first, token.points.size(), &r));
    if (!r.updated) return std::nullopt;
    Matrix3d cov;
    for (int i = 0; i < 9; ++i) cov[i] = r.estimate.cov[i];
    return std::make_pair(state_type::from_data(r.estimate.mean), cov);
  }

  /// amcl_core.hpp:204
  void force_update() { bb200_amcl_force_update(handle_); }

  [[nodiscard]] bb200_amcl* handle() const { return handle_; }

 private:
  void check(int status) const {
    if (status != BB200_OK) throw Error(status, bb200_amcl_last_error(handle_));
  }

  AmclParams params_;
  MotionModel motion_model_;
  SensorModel sensor_model_;
  bb200_amcl* handle_{nullptr};
};

template <class MotionModel, class SensorModel>
Amcl(MotionModel, SensorModel, const AmclParams&) -> Amcl<MotionModel, SensorModel>;

/// beluga::ParticleClusterizerParam (algorithm/cluster_based_estimation.hpp:259-276).
struct ParticleClusterizerParam {
  double linear_hash_resolution = 0.20;
  double angular_hash_resolution = 0.524;
  double weight_cap_percentile = 0.90;
};

/// beluga::cluster_based_estimate(states, weights, parameters) (cluster_based_estimation.hpp:415-432)
/// over the filter's device-resident particle set: what beluga_ros::Amcl::update returns
/// (beluga_ros/src/amcl.cpp:125).  Throws Error when the filter holds no particles.
template <class MotionModel, class SensorModel, class ExecutionPolicy>
[[nodiscard]] std::pair<SE2d, Matrix3d> cluster_based_estimate(
    const Amcl<MotionModel, SensorModel, ExecutionPolicy>& amcl,
    ParticleClusterizerParam parameters = {}) {
  const bb200_cluster_param p{parameters.linear_hash_resolution, parameters.angular_hash_resolution, parameters.weight_cap_percentile};
  bb200_estimate e{};
  bb200_filter* f = bb200_amcl_filter(amcl.handle());
  const int st = bb200_filter_cluster_estimate(f, &p, &e, nullptr, 0, nullptr, nullptr);
  if (st != BB200_OK) throw Error(st, bb200_last_error(f));
  Matrix3d cov;
  for (int i = 0; i < 9; ++i) cov[i] = e.cov[i];
  return std::make_pair(SE2d::from_data(e.mean), cov);
}

}  // namespace beluga_b200
