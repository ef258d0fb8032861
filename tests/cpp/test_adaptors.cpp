// Drop-in check of the C++ adaptors: the reference's own smoke tests of beluga::Amcl
// (beluga/test/beluga/algorithm/test_amcl_core.cpp:73-186) with the namespace switched to beluga_b200.
// Built by tests/test_cpp_adaptors.py with g++ against libbeluga_b200.so; needs a GPU to run.
#include <array>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "beluga_b200/amcl.hpp"
#include "beluga_b200/particle_cloud.hpp"

namespace {

int g_failures = 0;
#define ASSERT_TRUE(cond)                                                          \
  do {                                                                             \
    if (!(cond)) {                                                                 \
      std::fprintf(stderr, "%s:%d: assertion failed: %s\n", __FILE__, __LINE__, #cond); \
      ++g_failures;                                                                \
      return;                                                                      \
    }                                                                              \
  } while (0)
#define ASSERT_EQ(a, b) ASSERT_TRUE((a) == (b))
#define ASSERT_FALSE(cond) ASSERT_TRUE(!(cond))

/// beluga::testing::StaticOccupancyGrid<Rows, Cols, bool> (test/beluga/include/beluga/test/static_occupancy_grid.hpp:53-73).
template <std::size_t Rows, std::size_t Cols>
class StaticOccupancyGrid {
 public:
  struct ValueTraits {
    [[nodiscard]] static bool is_free(bool v) { return !v; }
    [[nodiscard]] static bool is_unknown(bool) { return false; }
    [[nodiscard]] static bool is_occupied(bool v) { return v; }
  };
  explicit StaticOccupancyGrid(std::array<bool, Rows * Cols> array, double resolution = 1.0, beluga_b200::SE2d origin = {})
      : grid_{array}, origin_{origin}, resolution_{resolution} {}
  [[nodiscard]] const beluga_b200::SE2d& origin() const { return origin_; }
  [[nodiscard]] const auto& data() const { return grid_; }
  [[nodiscard]] std::size_t size() const { return grid_.size(); }
  [[nodiscard]] std::size_t width() const { return Cols; }
  [[nodiscard]] std::size_t height() const { return Rows; }
  [[nodiscard]] double resolution() const { return resolution_; }
  [[nodiscard]] ValueTraits value_traits() const { return {}; }

 private:
  std::array<bool, Rows * Cols> grid_;
  beluga_b200::SE2d origin_;
  double resolution_;
};

using Grid = StaticOccupancyGrid<5, 5>;
const std::vector<std::pair<double, double>> kDummyMeasurement = {{0.0, 0.0}, {0.0, 0.0}, {0.0, 0.0}};  // as in the reference (zero-range beams)
const beluga_b200::SE2d kDummyControl{};
const beluga_b200::Matrix3d kIdentityCov = {1, 0, 0, 0, 1, 0, 0, 0, 1};

auto make_amcl(const beluga_b200::AmclParams& params = {}) {
  constexpr bool F = false, T = true;
  const auto map = Grid{{F, F, F, F, F, F, F, F, F, F, F, F, T, F, F, F, F, F, F, F, F, F, F, F, F}, 1.0};
  const beluga_b200::BeamModelParam param{};
  return beluga_b200::Amcl{beluga_b200::DifferentialDriveModel{beluga_b200::DifferentialDriveModelParam{}},
                           beluga_b200::BeamSensorModel<Grid>{param, map}, params};
}

void InitializeWithNoParticles() {
  auto amcl = make_amcl();
  ASSERT_EQ(amcl.particles().size(), 0u);
}
void InitializeFromPose() {
  auto amcl = make_amcl();
  amcl.initialize(beluga_b200::SE2d{}, kIdentityCov);
  ASSERT_EQ(amcl.particles().size(), beluga_b200::AmclParams{}.max_particles);
}
void UpdateWithNoParticles() {
  auto amcl = make_amcl();
  ASSERT_FALSE(amcl.update(kDummyControl, kDummyMeasurement).has_value());
}
void UpdateWithParticlesNoMotionAndForced() {
  auto amcl = make_amcl();
  amcl.initialize(beluga_b200::SE2d{}, kIdentityCov);
  ASSERT_TRUE(amcl.update(kDummyControl, kDummyMeasurement).has_value());
  ASSERT_FALSE(amcl.update(kDummyControl, kDummyMeasurement).has_value());
  amcl.force_update();
  ASSERT_TRUE(amcl.update(kDummyControl, kDummyMeasurement).has_value());
}
void LikelihoodFieldModelCanBeUsed() {
  constexpr bool F = false;
  const auto map = Grid{{F, F, F, F, F, F, F, F, F, F, F, F, F, F, F, F, F, F, F, F, F, F, F, F, F}, 0.5};
  beluga_b200::Amcl amcl{beluga_b200::DifferentialDriveModel{beluga_b200::DifferentialDriveModelParam{}},
                         beluga_b200::LikelihoodFieldModel<Grid>{beluga_b200::LikelihoodFieldModelParam{}, map}, beluga_b200::AmclParams{}};
  amcl.initialize(beluga_b200::SE2d{}, kIdentityCov);
  ASSERT_EQ(amcl.particles().size(), beluga_b200::AmclParams{}.max_particles);
  ASSERT_TRUE(amcl.update(kDummyControl, kDummyMeasurement).has_value());
}
void SelectiveResampleCanBeConstructed() {
  auto params = beluga_b200::AmclParams{};
  params.selective_resampling = true;
  auto amcl = make_amcl(params);
  amcl.initialize(beluga_b200::SE2d{}, kIdentityCov);
  ASSERT_TRUE(amcl.update(kDummyControl, kDummyMeasurement).has_value());
}
void TestRandomParticlesInserting() {
  auto params = beluga_b200::AmclParams{};
  params.min_particles = 2;
  params.max_particles = 100;
  params.alpha_slow = 0.0;
  params.alpha_fast = 100.0;  // Ensure we exercise random state generation
  auto amcl = make_amcl(params);
  amcl.initialize(beluga_b200::SE2d{0.0, 1.0, 1.0}, kIdentityCov);
  for (int i = 0; i < 30; ++i) {
    amcl.force_update();
    const auto estimate = amcl.update(kDummyControl, kDummyMeasurement);
    ASSERT_TRUE(estimate.has_value());
    const auto n = amcl.particles().size();
    ASSERT_TRUE(n >= 2 && n <= 100);
  }
}
void OtherMotionModelsCanBeUsed() {  // beluga_ros::Amcl::motion_model_variant (beluga_ros/amcl.hpp:108-111)
  constexpr bool F = false;
  const auto map = Grid{{F, F, F, F, F, F, F, F, F, F, F, F, F, F, F, F, F, F, F, F, F, F, F, F, F}, 0.5};
  beluga_b200::Amcl omni{beluga_b200::OmnidirectionalDriveModel{beluga_b200::OmnidirectionalDriveModelParam{0.1, 0.1, 0.1, 0.1, 0.1}},
                         beluga_b200::LikelihoodFieldModel<Grid>{beluga_b200::LikelihoodFieldModelParam{}, map}, beluga_b200::AmclParams{}};
  omni.initialize(beluga_b200::SE2d{}, kIdentityCov);
  ASSERT_TRUE(omni.update(beluga_b200::SE2d{0.1, 0.5, 0.2}, kDummyMeasurement).has_value());
  beluga_b200::Amcl still{beluga_b200::StationaryModel{}, beluga_b200::LikelihoodFieldModel<Grid>{beluga_b200::LikelihoodFieldModelParam{}, map},
                          beluga_b200::AmclParams{}};
  still.initialize(beluga_b200::SE2d{}, kIdentityCov);
  ASSERT_TRUE(still.update(kDummyControl, kDummyMeasurement).has_value());
}
void ClusterBasedEstimateCanBeUsed() {  // beluga_ros/src/amcl.cpp:125
  auto amcl = make_amcl();
  bool thrown = false;
  try {
    (void)beluga_b200::cluster_based_estimate(amcl);
  } catch (const beluga_b200::Error&) {
    thrown = true;  // no particles yet
  }
  ASSERT_TRUE(thrown);
  amcl.initialize(beluga_b200::SE2d{0.3, 1.0, 1.0}, beluga_b200::Matrix3d{0.01, 0, 0, 0, 0.01, 0, 0, 0, 0.01});
  const auto plain = amcl.update(kDummyControl, kDummyMeasurement);
  ASSERT_TRUE(plain.has_value());
  const auto [pose, covariance] = beluga_b200::cluster_based_estimate(amcl, beluga_b200::ParticleClusterizerParam{});
  // one tight blob: the heaviest cluster is (nearly) the whole set
  ASSERT_TRUE(std::abs(pose.x() - plain->first.x()) < 0.1 && std::abs(pose.y() - plain->first.y()) < 0.1);
  ASSERT_TRUE(covariance[0] > 0.0 && covariance[4] > 0.0);
}
void OutputBuildersCanBeUsed() {  // beluga_ros/particle_cloud.hpp:129-147,197-294, likelihood_field.hpp:44-79
  auto amcl = make_amcl();
  amcl.initialize(beluga_b200::SE2d{0.3, 1.0, 1.0}, beluga_b200::Matrix3d{0.01, 0, 0, 0, 0.01, 0, 0, 0, 0.01});
  ASSERT_TRUE(amcl.update(kDummyControl, kDummyMeasurement).has_value());
  const auto markers = beluga_b200::particle_cloud_markers(amcl);
  ASSERT_TRUE(!markers.bodies.empty() && markers.bodies.size() % 2 == 0 && markers.heads.size() == markers.bodies.size() / 2 * 3);
  ASSERT_TRUE(markers.body_scale_x > 0.0 && markers.body_scale_x <= 0.02 * 0.8 + 1e-12);
  const auto poses = beluga_b200::sample_poses(amcl, 25);
  ASSERT_TRUE(poses.size() == 25);
  for (const auto& p : poses) ASSERT_TRUE(std::abs(p.cos_yaw * p.cos_yaw + p.sin_yaw * p.sin_yaw - 1.0) < 1e-12);
#ifdef BELUGA_B200_WITH_SOPHUS
  const Sophus::SE2d sophus_pose = beluga_b200::SE2d{0.3, 1.0, 1.0};  // the conversion path, type-checked against the stub headers
  const beluga_b200::SE2d back{sophus_pose};
  ASSERT_TRUE(back.x() == 1.0 && back.y() == 1.0);
#endif
}
void InvalidCovarianceThrows() {  // multivariate_normal_distribution.hpp:114-116
  auto amcl = make_amcl();
  bool thrown = false;
  try {
    amcl.initialize(beluga_b200::SE2d{}, beluga_b200::Matrix3d{1, 0.5, 0, 0, 1, 0, 0, 0, 1});
  } catch (const std::runtime_error&) {
    thrown = true;
  }
  ASSERT_TRUE(thrown);
}

}  // namespace

int main() {
  InitializeWithNoParticles();
  InitializeFromPose();
  UpdateWithNoParticles();
  UpdateWithParticlesNoMotionAndForced();
  LikelihoodFieldModelCanBeUsed();
  SelectiveResampleCanBeConstructed();
  TestRandomParticlesInserting();
  OtherMotionModelsCanBeUsed();
  ClusterBasedEstimateCanBeUsed();
  OutputBuildersCanBeUsed();
  InvalidCovarianceThrows();
  if (g_failures == 0) std::printf("CPP_ADAPTORS_OK\n");
  return g_failures == 0 ? 0 : 1;
}
