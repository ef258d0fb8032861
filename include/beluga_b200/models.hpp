// beluga_b200/models.hpp -- MotionModel / SensorModel shaped adaptors over the C ABI.
//
// Header-only C++17.  These types have the member types, constructor arguments and parameter
// structs of the reference models so that code written against
//   beluga::DifferentialDriveModel2d      (beluga/motion/differential_drive_model.hpp:77-174)
//   beluga::LikelihoodFieldModel<Grid>    (beluga/sensor/likelihood_field_model.hpp:41-92)
//   beluga::LikelihoodFieldProbModel<Grid>(beluga/sensor/likelihood_field_prob_model.hpp:41-92)
//   beluga::BeamSensorModel<Grid>         (beluga/sensor/beam_model.hpp:73-163)
// keeps compiling when the namespace is switched to beluga_b200.  The reference models return
// per-particle callables (state sampling / weighting functions); here `operator()` returns a small
// token that beluga_b200::Amcl hands to the device kernels -- the per-particle loop lives on the GPU.
//
// Sophus/Eigen are optional: define BELUGA_B200_WITH_SOPHUS before including to get converting
// constructors from/to Sophus::SE2d and Eigen::Matrix3d.
#pragma once

#include <array>
#include <cmath>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <tuple>
#include <type_traits>
#include <utility>
#include <vector>

#include "../beluga_b200.h"

#ifdef BELUGA_B200_WITH_SOPHUS
#include <Eigen/Core>
#include <sophus/se2.hpp>
#endif

namespace beluga_b200 {

/// 2-D pose with Sophus::SE2d's memory layout {cos, sin, x, y} (Sophus::SE2d::data()).
struct SE2d {
  std::array<double, 4> v{1.0, 0.0, 0.0, 0.0};

  SE2d() = default;
  /// Sophus::SE2d{theta, translation}: the rotation is normalised through hypot like SO2::exp.
  SE2d(double theta, double x, double y) {
    const double c = std::cos(theta), s = std::sin(theta), n = std::hypot(c, s);
    v = {c / n, s / n, x, y};
  }
  static SE2d from_data(const double* d) {
    SE2d p;
    p.v = {d[0], d[1], d[2], d[3]};
    return p;
  }
#ifdef BELUGA_B200_WITH_SOPHUS
  SE2d(const Sophus::SE2d& p) : v{p.data()[0], p.data()[1], p.data()[2], p.data()[3]} {}  // NOLINT(google-explicit-constructor)
  operator Sophus::SE2d() const {                                                           // NOLINT(google-explicit-constructor)
    Sophus::SE2d out;
    std::copy(v.begin(), v.end(), out.data());
    return out;
  }
#endif
  [[nodiscard]] const double* data() const { return v.data(); }
  [[nodiscard]] double x() const { return v[2]; }
  [[nodiscard]] double y() const { return v[3]; }
  [[nodiscard]] double theta() const { return std::atan2(v[1], v[0]); }  // so2().log()
};

/// 3x3 covariance over (x, y, theta), row-major (Sophus::Matrix3d of estimation.hpp:436).
using Matrix3d = std::array<double, 9>;

/// Thrown where the reference throws (std::runtime_error / std::invalid_argument) or a CUDA call fails.
class Error : public std::runtime_error {
 public:
  Error(int status, const std::string& what) : std::runtime_error(what), status_(status) {}
  [[nodiscard]] int status() const { return status_; }

 private:
  int status_;
};

// ---- motion --------------------------------------------------------------------------------------

/// Same members and defaults as beluga::DifferentialDriveModelParam (differential_drive_model.hpp:40-68).
struct DifferentialDriveModelParam {
  double rotation_noise_from_rotation;
  double rotation_noise_from_translation;
  double translation_noise_from_translation;
  double translation_noise_from_rotation;
  double distance_threshold = 0.01;
};

class DifferentialDriveModel {
 public:
  using state_type = SE2d;
  using control_type = std::tuple<state_type, state_type>;  // (current, previous) odometry poses
  using param_type = DifferentialDriveModelParam;

  explicit DifferentialDriveModel(const param_type& params) : params_{params} {}

  /// Token of a control action: the three normal distributions of sampling_fn_2d (:141-154).
  struct sampling_token {
    bb200_diff_drive_sampling sampling;
  };

  template <class Control>
  [[nodiscard]] sampling_token operator()(const Control& action) const {
    const auto& [pose, previous_pose] = action;
    sampling_token t{};
    const bb200_diff_drive_param p = c_param();
    bb200_diff_drive_sampling_from_control(&p, state_type(pose).data(), state_type(previous_pose).data(), &t.sampling);
    return t;
  }

  [[nodiscard]] bb200_diff_drive_param c_param() const {
    return bb200_diff_drive_param{params_.rotation_noise_from_rotation, params_.rotation_noise_from_translation,
                                  params_.translation_noise_from_translation, params_.translation_noise_from_rotation,
                                  params_.distance_threshold};
  }
  [[nodiscard]] bb200_motion_param c_motion_param() const {
    return bb200_motion_param{BB200_MOTION_DIFFERENTIAL, params_.rotation_noise_from_rotation, params_.rotation_noise_from_translation,
                              params_.translation_noise_from_translation, params_.translation_noise_from_rotation, 0.0, params_.distance_threshold};
  }

 private:
  param_type params_;
};
using DifferentialDriveModel2d = DifferentialDriveModel;

/// Same members and defaults as beluga::OmnidirectionalDriveModelParam (omnidirectional_drive_model.hpp:36-68).
struct OmnidirectionalDriveModelParam {
  double rotation_noise_from_rotation;
  double rotation_noise_from_translation;
  double translation_noise_from_translation;
  double translation_noise_from_rotation;
  double strafe_noise_from_translation;
  double distance_threshold = 0.01;
};

/// beluga::OmnidirectionalDriveModel (omnidirectional_drive_model.hpp:78-158).
class OmnidirectionalDriveModel {
 public:
  using state_type = SE2d;
  using control_type = std::tuple<state_type, state_type>;
  using param_type = OmnidirectionalDriveModelParam;
  explicit OmnidirectionalDriveModel(const param_type& params) : params_{params} {}
  [[nodiscard]] bb200_motion_param c_motion_param() const {
    return bb200_motion_param{BB200_MOTION_OMNIDIRECTIONAL, params_.rotation_noise_from_rotation, params_.rotation_noise_from_translation,
                              params_.translation_noise_from_translation, params_.translation_noise_from_rotation,
                              params_.strafe_noise_from_translation, params_.distance_threshold};
  }

 private:
  param_type params_;
};

/// beluga::StationaryModel (stationary_model.hpp:39-62).
class StationaryModel {
 public:
  using state_type = SE2d;
  using control_type = std::tuple<state_type, state_type>;
  [[nodiscard]] bb200_motion_param c_motion_param() const { return bb200_motion_param{BB200_MOTION_STATIONARY, 0, 0, 0, 0, 0, 0.01}; }
};

// ---- maps ----------------------------------------------------------------------------------------

/// Flattens any type satisfying the reference's OccupancyGrid2 requirements
/// (sensor/data/occupancy_grid.hpp: width(), height(), resolution(), origin(), data(), value_traits())
/// into the trinary int8 view of the C ABI.
struct GridSnapshot {
  std::vector<std::int8_t> cells;
  std::int32_t width{0}, height{0};
  double resolution{1.0};
  SE2d origin{};

  template <class OccupancyGrid>
  explicit GridSnapshot(const OccupancyGrid& grid)
      : width(static_cast<std::int32_t>(grid.width())), height(static_cast<std::int32_t>(grid.height())), resolution(grid.resolution()), origin(grid.origin()) {
    const auto traits = grid.value_traits();
    cells.reserve(grid.size());
    for (const auto& value : grid.data()) {
      cells.push_back(traits.is_occupied(value) ? std::int8_t{100} : (traits.is_free(value) ? std::int8_t{0} : std::int8_t{-1}));
    }
  }

  /// The C-ABI view; valid while this snapshot is alive and unmodified.
  [[nodiscard]] bb200_occupancy_grid view() const {
    bb200_occupancy_grid v{};
    v.cells = cells.data();
    v.width = width;
    v.height = height;
    v.resolution = resolution;
    for (int i = 0; i < 4; ++i) v.origin[i] = origin.data()[i];
    return v;
  }
};

/// A plain occupancy grid for callers without their own grid type (ROS trinary values).
class OccupancyGrid {
 public:
  struct ValueTraits {
    [[nodiscard]] static bool is_free(std::int8_t v) { return v == 0; }
    [[nodiscard]] static bool is_unknown(std::int8_t v) { return v == -1; }
    [[nodiscard]] static bool is_occupied(std::int8_t v) { return v == 100; }
  };
  OccupancyGrid(std::vector<std::int8_t> cells, std::size_t width, double resolution, SE2d origin = SE2d{})
      : cells_(std::move(cells)), width_(width), resolution_(resolution), origin_(origin) {}
  [[nodiscard]] std::size_t width() const { return width_; }
  [[nodiscard]] std::size_t height() const { return cells_.size() / width_; }
  [[nodiscard]] std::size_t size() const { return cells_.size(); }
  [[nodiscard]] double resolution() const { return resolution_; }
  [[nodiscard]] const SE2d& origin() const { return origin_; }
  [[nodiscard]] const std::vector<std::int8_t>& data() const { return cells_; }
  [[nodiscard]] ValueTraits value_traits() const { return {}; }

 private:
  std::vector<std::int8_t> cells_;
  std::size_t width_;
  double resolution_;
  SE2d origin_;
};

// ---- sensors -------------------------------------------------------------------------------------

/// Same members and defaults as beluga::LikelihoodFieldModelBaseParam (likelihood_field_model_base.hpp:42-64).
struct LikelihoodFieldModelParam {
  double max_obstacle_distance = 100.0;
  double max_laser_distance = 2.0;
  double z_hit = 0.5;
  double z_random = 0.5;
  double sigma_hit = 0.2;
  bool model_unknown_space = false;
  bool only_obstacle_boundaries = false;
};

/// Same members and defaults as beluga::BeamModelParam (beam_model.hpp:43-58).
struct BeamModelParam {
  double z_hit{0.5};
  double z_short{0.5};
  double z_max{0.05};
  double z_rand{0.05};
  double sigma_hit{0.2};
  double lambda_short{0.1};
  double beam_max_range{60};
};

/// Measurement token: the lidar hit points in the particle frame (measurement_type of the reference).
struct measurement_token {
  std::vector<std::pair<double, double>> points;
};

namespace detail {

template <class OccupancyGridT, bool kProb>
class LikelihoodFieldModelImpl {
 public:
  using state_type = SE2d;
  using weight_type = double;
  using measurement_type = std::vector<std::pair<double, double>>;
  using map_type = OccupancyGridT;
  using param_type = LikelihoodFieldModelParam;

  explicit LikelihoodFieldModelImpl(const param_type& params, const map_type& grid) : params_{params}, grid_{grid} {}
  void update_map(const map_type& grid) {
    grid_ = GridSnapshot{grid};
    ++map_version_;
  }
  [[nodiscard]] measurement_token operator()(measurement_type&& points) const { return measurement_token{std::move(points)}; }

  // -- used by beluga_b200::Amcl --
  [[nodiscard]] int attach(bb200_filter* f) const {
    const bb200_likelihood_field_param p{params_.max_obstacle_distance, params_.max_laser_distance, params_.z_hit,      params_.z_random,
                                         params_.sigma_hit,             params_.model_unknown_space ? 1 : 0, params_.only_obstacle_boundaries ? 1 : 0};
    const bb200_occupancy_grid g = grid_.view();
    return bb200_filter_set_likelihood_field_map(f, &p, &g, kProb ? 1 : 0);
  }
  [[nodiscard]] unsigned map_version() const { return map_version_; }

 private:
  param_type params_;
  GridSnapshot grid_;
  unsigned map_version_{0};
};

}  // namespace detail

template <class OccupancyGridT = OccupancyGrid>
using LikelihoodFieldModel = detail::LikelihoodFieldModelImpl<OccupancyGridT, false>;
template <class OccupancyGridT = OccupancyGrid>
using LikelihoodFieldProbModel = detail::LikelihoodFieldModelImpl<OccupancyGridT, true>;

template <class OccupancyGridT = OccupancyGrid>
class BeamSensorModel {
 public:
  using state_type = SE2d;
  using weight_type = double;
  using measurement_type = std::vector<std::pair<double, double>>;
  using map_type = OccupancyGridT;
  using param_type = BeamModelParam;

  explicit BeamSensorModel(const param_type& params, const map_type& grid) : params_{params}, grid_{grid} {}
  void update_map(const map_type& grid) {
    grid_ = GridSnapshot{grid};
    ++map_version_;
  }
  [[nodiscard]] measurement_token operator()(measurement_type&& points) const { return measurement_token{std::move(points)}; }

  [[nodiscard]] int attach(bb200_filter* f) const {
    const bb200_beam_param p{params_.z_hit, params_.z_short, params_.z_max, params_.z_rand, params_.sigma_hit, params_.lambda_short, params_.beam_max_range};
    const bb200_occupancy_grid g = grid_.view();
    return bb200_filter_set_beam_map(f, &p, &g);
  }
  [[nodiscard]] unsigned map_version() const { return map_version_; }

 private:
  param_type params_;
  GridSnapshot grid_;
  unsigned map_version_{0};
};

}  // namespace beluga_b200
