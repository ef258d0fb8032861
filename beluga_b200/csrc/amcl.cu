#include "amcl.hpp"

#include <algorithm>
#include <cmath>
#include <initializer_list>
#include <limits>
#include <vector>

namespace bb200 {

namespace {

/// rotation_variance -- differential_drive_model.hpp:167-173: backward and forward motion are
/// treated symmetrically.
double rotation_variance(const Rot2& rotation) {
  static const Rot2 kFlip = rot_exp(3.14159265358979323846);
  const Rot2 flipped = rot_mul(rotation, kFlip);
  const double delta = std::min(std::fabs(rot_log(rotation)), std::fabs(rot_log(flipped)));
  return delta * delta;
}

double exponential_filter(double& output, double alpha, double input) {  // exponential_filter.hpp:41-44
  output += (output == 0.) ? input : alpha * (input - output);
  return output;
}

}  // namespace

bb200_diff_drive_sampling diff_drive_sampling(const bb200_diff_drive_param& p, const Pose2& pose, const Pose2& prev) {
  const double tx = pose.x - prev.x, ty = pose.y - prev.y;
  const double distance = std::sqrt(tx * tx + ty * ty);
  const double distance_variance = distance * distance;
  const Rot2 previous_orientation{prev.c, prev.s};
  const Rot2 current_orientation{pose.c, pose.s};
  const Rot2 heading_rotation = rot_exp(std::atan2(ty, tx));
  const Rot2 first_rotation = distance > p.distance_threshold ? rot_mul(heading_rotation, rot_inverse(previous_orientation)) : Rot2{1.0, 0.0};
  const Rot2 second_rotation = rot_mul(rot_mul(current_orientation, rot_inverse(previous_orientation)), rot_inverse(first_rotation));
  const double v1 = rotation_variance(first_rotation), v2 = rotation_variance(second_rotation);
  bb200_diff_drive_sampling s;
  s.rot1_mean = rot_log(first_rotation);
  s.rot1_std = std::sqrt(p.rotation_noise_from_rotation * v1 + p.rotation_noise_from_translation * distance_variance);
  s.trans_mean = distance;
  s.trans_std = std::sqrt(p.translation_noise_from_translation * distance_variance + p.translation_noise_from_rotation * (v1 + v2));
  s.rot2_mean = rot_log(second_rotation);
  s.rot2_std = std::sqrt(p.rotation_noise_from_rotation * v2 + p.rotation_noise_from_translation * distance_variance);
  return s;
}

bb200_motion_sampling motion_sampling(const bb200_motion_param& p, const Pose2& pose, const Pose2& prev) {
  bb200_motion_sampling out{};
  out.model = p.model;
  out.first_rotation[0] = 1.0;
  out.first_rotation[1] = 0.0;
  if (p.model == BB200_MOTION_STATIONARY) {  // stationary_model.hpp:55: N(0, 0.02) for theta, x, y
    for (int k = 0; k < 3; ++k) {
      out.mean[k] = 0.0;
      out.stddev[k] = 0.02;
    }
    return out;
  }
  if (p.model == BB200_MOTION_OMNIDIRECTIONAL) {  // omnidirectional_drive_model.hpp:101-129
    const double tx = pose.x - prev.x, ty = pose.y - prev.y;
    const double distance = std::sqrt(tx * tx + ty * ty);
    const double distance_variance = distance * distance;
    const Rot2 previous_orientation{prev.c, prev.s};
    const Rot2 current_orientation{pose.c, pose.s};
    const Rot2 rotation = rot_mul(current_orientation, rot_inverse(previous_orientation));
    const Rot2 heading_rotation = rot_exp(std::atan2(ty, tx));
    const Rot2 first_rotation = distance > p.distance_threshold ? rot_mul(heading_rotation, rot_inverse(previous_orientation)) : Rot2{1.0, 0.0};
    const double rv = rotation_variance(rotation);
    out.mean[0] = rot_log(rotation);
    out.stddev[0] = std::sqrt(p.rotation_noise_from_rotation * rv + p.rotation_noise_from_translation * distance_variance);
    out.mean[1] = distance;
    out.stddev[1] = std::sqrt(p.translation_noise_from_translation * distance_variance + p.translation_noise_from_rotation * rv);
    out.mean[2] = 0.0;
    out.stddev[2] = std::sqrt(p.strafe_noise_from_translation * distance_variance + p.translation_noise_from_rotation * rv);
    out.first_rotation[0] = first_rotation.c;
    out.first_rotation[1] = first_rotation.s;
    return out;
  }
  const bb200_diff_drive_param d{p.rotation_noise_from_rotation, p.rotation_noise_from_translation, p.translation_noise_from_translation,
                                 p.translation_noise_from_rotation, p.distance_threshold};
  const bb200_diff_drive_sampling s = diff_drive_sampling(d, pose, prev);
  out.mean[0] = s.rot1_mean, out.stddev[0] = s.rot1_std;
  out.mean[1] = s.trans_mean, out.stddev[1] = s.trans_std;
  out.mean[2] = s.rot2_mean, out.stddev[2] = s.rot2_std;
  return out;
}

Amcl::Amcl(const bb200_amcl_param& p, const bb200_motion_param& motion) : params_(p), motion_(motion) {
  bb200_filter_config c{};
  c.device = p.device;
  c.capacity = p.shard_capacity != 0 ? p.shard_capacity : p.max_particles;
  c.seed = p.seed;
  c.first_index = p.shard_first_index;
  c.global_count = p.max_particles;
  c.record_ancestors = p.record_ancestors;
  filter_ = std::make_unique<Filter>(c);
  if (params_.resample_interval == 0) params_.resample_interval = 1;
  if (sharded() && params_.min_particles < params_.max_particles && filter_->ok()) (void)filter_->enable_shard_kld();
}

int Amcl::initialize(const double mean[3], const double cov[9]) {
  error_.clear();
  const int st = filter_->initialize_normal(mean, cov, sharded() ? params_.shard_capacity : params_.max_particles);
  if (st == BB200_OK) {
    force_update_ = true;  // amcl_core.hpp:136
    initialized_ = true;
  }
  return st;
}

int Amcl::initialize_from_map() {
  error_.clear();
  const int st = filter_->initialize_uniform(sharded() ? params_.shard_capacity : params_.max_particles);
  if (st == BB200_OK) {
    force_update_ = true;  // beluga_ros/include/beluga_ros/amcl.hpp:193-199
    initialized_ = true;
  }
  return st;
}

int Amcl::initialize_states(const double* states, const double* weights, uint64_t n) {
  error_.clear();
  const int st = filter_->set_particles(states, weights, n);
  if (st == BB200_OK) {
    force_update_ = true;
    initialized_ = n > 0;
  }
  return st;
}

int Amcl::plan_update(const double control[4], bb200_step_plan* plan) {
  error_.clear();
  *plan = bb200_step_plan{};
  const Pose2 pose{control[0], control[1], control[2], control[3]};
  // Sharded filters gate on the GLOBAL particle count, which never drops to zero once initialised.
  const uint64_t n = sharded() ? (initialized_ ? filter_->global_size() : 0u) : filter_->size();
  if (n == 0) return BB200_OK;  // amcl_core.hpp:166-168 -> std::nullopt

  // update_policy_(control_action) -- on_motion vs the last ACCEPTED pose (on_motion.hpp:121-133)
  bool moved;
  if (!latest_pose_) {
    latest_pose_ = pose;
    moved = true;
  } else {
    const Pose2 delta = pose_mul(pose_inverse(*latest_pose_), pose);
    moved = std::sqrt(delta.x * delta.x + delta.y * delta.y) > params_.update_min_d ||
            std::fabs(rot_log(Rot2{delta.c, delta.s})) > params_.update_min_a;
    if (moved) latest_pose_ = pose;
  }
  if (!moved && !force_update_) return BB200_OK;

  // control_action_window_ << control (circular_array.hpp:473-480); reads clamp to size-1 (:353-361)
  window_[1] = window_[0];
  window_[0] = pose;
  window_size_ = std::min(window_size_ + 1, 2);
  const Pose2& previous = window_[std::min(1, window_size_ - 1)];
  plan->sampling = motion_sampling(motion_, window_[0], previous);
  plan->step = ++step_;

  // random_probability_estimator_(particles_) (thrun_recovery_probability_estimator.hpp:69-89) on the
  // normalised weights, whose total is 1 by construction: average = 1 / N.
  const double average_weight = 1.0 / static_cast<double>(n);
  const double fast_average = exponential_filter(fast_output_, params_.alpha_fast, average_weight);
  const double slow_average = exponential_filter(slow_output_, params_.alpha_slow, average_weight);
  double random_state_probability = 0.0;
  if (!(std::fabs(slow_average) < std::numeric_limits<double>::epsilon())) {
    random_state_probability = std::clamp(1.0 - fast_average / slow_average, 0.0, 1.0);
  }
  if (params_.recovery_probability_override > 0.0) random_state_probability = std::min(params_.recovery_probability_override, 1.0);
  plan->random_state_probability = random_state_probability;

  // resample_policy_: every_n (every_n.hpp:47-50); on_effective_size_drop is applied by the caller
  // once the effective sample size is known (commit_resample_decision).
  every_n_current_ = (every_n_current_ + 1) % params_.resample_interval;
  plan->resample = every_n_current_ == 0 ? 1 : 0;
  plan->needs_ess = (plan->resample && params_.selective_resampling) ? 1 : 0;

  bb200_resample_opts& o = plan->opts;
  o.scheme = params_.resample_scheme;
  o.step = step_;
  o.min_particles = params_.min_particles;
  o.max_particles = params_.max_particles;
  o.kld_epsilon = params_.kld_epsilon;
  o.kld_z = params_.kld_z;
  for (int k = 0; k < 3; ++k) o.spatial_resolution[k] = params_.spatial_resolution[k];
  o.random_state_probability = random_state_probability;
  plan->update = 1;
  return BB200_OK;
}

void Amcl::commit_update(int resampled, double random_state_probability) {
  if (resampled && random_state_probability > 0.0) fast_output_ = slow_output_ = 0.0;  // amcl_core.hpp:184-186
  force_update_ = false;
}

Amcl::HostState Amcl::snapshot() const {
  return HostState{latest_pose_, every_n_current_, slow_output_, fast_output_, {window_[0], window_[1]}, window_size_, force_update_, step_};
}

void Amcl::restore(const HostState& s) {
  latest_pose_ = s.latest_pose;
  every_n_current_ = s.every_n_current;
  slow_output_ = s.slow_output;
  fast_output_ = s.fast_output;
  window_[0] = s.window[0];
  window_[1] = s.window[1];
  window_size_ = s.window_size;
  force_update_ = s.force_update;
  step_ = s.step;
}

int Amcl::update_group(Amcl* const* ranks, int count, const double control[4], const double* points_xy, uint64_t n_points,
                       bb200_update_result* out) {
  *out = bb200_update_result{};
  if (ranks == nullptr || count < 1 || count > kMaxShards) return BB200_ERR_INVALID_ARGUMENT;
  Amcl& lead = *ranks[0];
  const bool kld = lead.params_.min_particles < lead.params_.max_particles;
  if (kld)
    for (int r = 0; r < count; ++r)
      if (!ranks[r]->filter_->shard_kld())
        return ranks[r]->filter_->fail_with(BB200_ERR_STATE, "KLD-adaptive resampling on shards: the hash arrays were not exported (create the shards with min_particles < max_particles before joining them)");
  for (int r = 0; r < count; ++r)
    if (ranks[r]->sharded() && ranks[r]->filter_->shard_world() <= 1)
      return ranks[r]->filter_->fail_with(BB200_ERR_STATE, "this shard has not joined its peers (bb200_amcl_join_shards / bb200_sharded_amcl_create)");
  std::vector<HostState> before;
  before.reserve(static_cast<size_t>(count));
  for (int r = 0; r < count; ++r) before.push_back(ranks[r]->snapshot());
  auto rollback = [&](int st) {
    for (int r = 0; r < count; ++r) {
      ranks[r]->restore(before[static_cast<size_t>(r)]);
      ranks[r]->filter_->step_abort();
    }
    return st;
  };
  // The host half is the same arithmetic on every shard (policies, control window, recovery estimator).
  bb200_step_plan plan{};
  for (int r = 0; r < count; ++r) {
    bb200_step_plan p{};
    const int st = ranks[r]->plan_update(control, &p);
    if (st != BB200_OK) return rollback(st);
    if (r == 0) plan = p;
    if (p.update != plan.update || p.step != plan.step) return rollback(lead.filter_->fail_with(BB200_ERR_STATE, "the shards of a filter disagree on the step plan"));
  }
  if (!plan.update) return BB200_OK;
  out->random_state_probability = plan.random_state_probability;

  auto run_phases = [&](std::initializer_list<int> phases) {
    for (const int phase : phases)
      for (int r = 0; r < count; ++r) {
        const int st = ranks[r]->filter_->step_phase(phase);
        if (st != BB200_OK) return st;
      }
    return static_cast<int>(BB200_OK);
  };
  auto close_batch = [&](double* sum_sq) {
    for (int r = 0; r < count; ++r) {
      bb200_estimate est{};
      double weight_sum = 0.0, sq = 0.0;
      uint64_t n = 0;
      const int st = ranks[r]->filter_->step_end(&est, &weight_sum, &n, &sq);
      if (st != BB200_OK) return st;
      if (r == 0) {  // the exchanged sums are the same bits on every shard
        out->estimate = est;
        out->weight_sum = weight_sum;
        out->n_particles = n;
        if (sum_sq != nullptr) *sum_sq = sq;
      }
    }
    return static_cast<int>(BB200_OK);
  };

  const bool resample_now = plan.resample != 0 && plan.needs_ess == 0;
  for (int r = 0; r < count; ++r) {
    const int st = ranks[r]->filter_->step_begin(plan.sampling, plan.step, points_xy, n_points, plan.opts, resample_now);
    if (st != BB200_OK) return rollback(st);
  }
  // views::take_while_kld over the shards (take_while_kld.hpp:72-137): the candidate stream is ordered by output slot, and
  // every rank sees the spatial hash of every candidate, so all ranks count distinct buckets over the same stream and
  // stop at the same slot.  Windows of slots double like the single-GPU chunks; one host decision per window.
  auto kld_size = [&](uint64_t* accepted) {
    const uint64_t max = lead.params_.max_particles;
    for (int r = 0; r < count; ++r) {
      const int st = ranks[r]->filter_->step_totals(nullptr, nullptr);  // enqueue the totals exchange on every shard
      if (st != BB200_OK) return st;
    }
    uint64_t begin = 0, k_before = 0;
    uint64_t end = std::min<uint64_t>(max, std::max<uint64_t>(2 * lead.params_.min_particles, 65536));
    *accepted = max;
    while (begin < max) {
      for (int r = 0; r < count; ++r) {
        const int st = ranks[r]->filter_->kld_sharded_candidates(begin, end);
        if (st != BB200_OK) return st;
      }
      for (int r = 0; r < count; ++r) {
        const int st = ranks[r]->filter_->kld_sharded_count(begin, end, k_before);
        if (st != BB200_OK) return st;
      }
      uint64_t cutoff = ~0ull, fresh = 0;
      for (int r = 0; r < count; ++r) {
        uint64_t c = 0, f = 0;
        const int st = ranks[r]->filter_->kld_sharded_read(&c, &f);
        if (st != BB200_OK) return st;
        if (r == 0) {
          cutoff = c;
          fresh = f;
        } else if (c != cutoff || f != fresh) {
          return lead.filter_->fail_with(BB200_ERR_STATE, "KLD on shards: the ranks disagree on the count");
        }
      }
      if (cutoff != ~0ull) {
        *accepted = std::min<uint64_t>(cutoff - 1, max);  // take_while drops the first element whose condition fails (:134-136)
        break;
      }
      k_before += fresh;
      begin = end;
      end = std::min<uint64_t>(max, end * 2);
    }
    return static_cast<int>(BB200_OK);
  };

  int st = BB200_OK;
  if (resample_now && kld) {
    st = run_phases({Filter::kPhaseReweight, Filter::kPhaseCdf});
    uint64_t accepted = 0;
    if (st == BB200_OK) st = kld_size(&accepted);
    if (st == BB200_OK) {
      for (int r = 0; r < count; ++r) ranks[r]->filter_->step_set_kld_accepted(accepted);
      st = run_phases({Filter::kPhaseResample, Filter::kPhaseFinish});
    }
  } else {
    st = resample_now ? run_phases({Filter::kPhaseReweight, Filter::kPhaseCdf, Filter::kPhaseResample, Filter::kPhaseFinish})
                      : run_phases({Filter::kPhaseReweight, Filter::kPhaseCdf, Filter::kPhaseNormalize, Filter::kPhaseFinish});
  }
  double sum_sq = 0.0;
  if (st == BB200_OK) st = close_batch(&sum_sq);
  if (st != BB200_OK) return rollback(st);
  bool resampled = resample_now;
  if (!resample_now && plan.resample != 0) {
    // on_effective_size_drop (on_effective_size_drop.hpp:45-49): ESS = 1 / sum w~^2 < N / 2
    const double ess = sum_sq > 0.0 ? 1.0 / sum_sq : 0.0;
    if (ess < static_cast<double>(lead.filter_->global_size()) * 0.5) {
      if (kld) {
        uint64_t accepted = 0;
        st = kld_size(&accepted);
        if (st != BB200_OK) return rollback(st);
        for (int r = 0; r < count; ++r) ranks[r]->filter_->step_set_kld_accepted(accepted);
      }
      st = run_phases({Filter::kPhaseResample, Filter::kPhaseFinish});
      if (st == BB200_OK) st = close_batch(nullptr);
      if (st != BB200_OK) return rollback(st);
      resampled = true;
    }
  }
  out->resampled = resampled ? 1 : 0;
  out->weights_degenerate = lead.filter_->last_weights_valid() ? 0 : 1;
  for (int r = 0; r < count; ++r) {
    ranks[r]->filter_->step_abort();
    ranks[r]->commit_update(out->resampled, plan.random_state_probability);
  }
  out->updated = 1;
  return BB200_OK;
}

int Amcl::update(const double control[4], const double* points_xy, uint64_t n_points, bb200_update_result* out) {
  if (sharded()) {  // this process drives one shard; the peers call update() themselves
    Amcl* self = this;
    return update_group(&self, 1, control, points_xy, n_points, out);
  }
  const HostState before = snapshot();
  const int st = update_device(control, points_xy, n_points, out);
  if (st != BB200_OK) restore(before);  // policies, control window and recovery estimator as if the call had not happened
  return st;
}

int Amcl::update_device(const double control[4], const double* points_xy, uint64_t n_points, bb200_update_result* out) {
  *out = bb200_update_result{};
  bb200_step_plan plan;
  int st = plan_update(control, &plan);
  if (st != BB200_OK || !plan.update) return st;
  out->random_state_probability = plan.random_state_probability;
  const uint64_t n = filter_->size();
  bool do_resample = plan.resample != 0;
  const bool kld_active = params_.min_particles < params_.max_particles;
  if (do_resample && !plan.needs_ess && !kld_active) {
    // The whole step in one stream-ordered sequence with a single host synchronisation.
    uint64_t new_size = 0;
    st = filter_->step_resample(plan.sampling, plan.step, points_xy, n_points, plan.opts, &out->estimate, &out->weight_sum, &new_size);
    if (st != BB200_OK) return st;
    out->resampled = 1;
    out->n_particles = new_size;
  } else {
    st = filter_->propagate_reweight(&plan.sampling, plan.step, points_xy, n_points);
    if (st != BB200_OK) return st;
    double factor = 0.0, sum_sq = 0.0;
    st = filter_->normalize(&factor, &sum_sq);
    if (st != BB200_OK) return st;
    out->weight_sum = factor;
    if (plan.needs_ess) {
      // on_effective_size_drop (on_effective_size_drop.hpp:45-49): ESS = 1 / sum w~^2 < N / 2
      const double ess = sum_sq > 0.0 ? 1.0 / sum_sq : 0.0;
      do_resample = ess < static_cast<double>(n) * 0.5;
    }
    if (do_resample) {
      uint64_t new_size = 0;
      st = filter_->resample(plan.opts, &new_size);
      if (st != BB200_OK) return st;
      out->resampled = 1;
    }
    st = filter_->estimate(&out->estimate);
    if (st != BB200_OK) return st;
    out->n_particles = filter_->size();
  }
  out->weights_degenerate = filter_->last_weights_valid() ? 0 : 1;
  commit_update(out->resampled, plan.random_state_probability);
  out->updated = 1;
  return BB200_OK;
}

}  // namespace bb200
