// Device-resident particle set of the B200 MCL backend: buffers, map uploads and the per-step
// kernel sequence behind the C ABI of include/beluga_b200.h.  Host-side counterpart of
// beluga::TupleVector<std::tuple<SE2d, Weight>> (reference: containers/tuple_vector.hpp:50-223)
// plus the actions/views that iterate it (propagate, reweight, normalize, sample, assign).
#pragma once

#include <cuda_runtime.h>

#include <cstdint>
#include <string>
#include <vector>

#include "../../include/beluga_b200.h"
#include "cluster.cuh"
#include "cluster_host.hpp"
#include "kernels.cuh"

namespace bb200 {

class Filter {
 public:
  explicit Filter(const bb200_filter_config& config);
  ~Filter();
  Filter(const Filter&) = delete;
  Filter& operator=(const Filter&) = delete;

  // Every method returns a bb200_status and records the message for last_error().
  int set_likelihood_field_map(const bb200_likelihood_field_param& p, const bb200_occupancy_grid& grid, bool prob);
  int set_beam_map(const bb200_beam_param& p, const bb200_occupancy_grid& grid);
  int get_likelihood_field(float* out, uint64_t capacity) const;

  int set_particles(const double* states, const double* weights, uint64_t n);
  int get_particles(double* states, double* weights, uint64_t capacity);
  int initialize_normal(const double mean[3], const double cov[9], uint64_t n);
  /// n samples of MultivariateUniformDistribution over the free cells of the current map (initialize_from_map).
  int initialize_uniform(uint64_t n);
  uint64_t size() const { return n_; }

  int propagate_reweight(const bb200_motion_sampling* sampling, uint32_t step, const double* points_xy, uint64_t n_points);
  int max_weight(double* wmax);
  int build_cdf(double global_wmax, uint64_t* local_total, int* exponent);
  int normalize_by(uint64_t global_total, double* local_sum_sq);
  int normalize(double* factor, double* sum_sq);
  int resample(const bb200_resample_opts& o, uint64_t* new_size);
  int resample_range(const bb200_resample_opts& o, uint64_t global_total, uint64_t cdf_offset, uint64_t slot_begin, uint64_t slot_end);
  int adopt(uint64_t n, int from_staging);
  int ancestors(int64_t* out, uint64_t capacity);
  int cdf(uint64_t* out, uint64_t capacity);
  int estimate(bb200_estimate* out);
  int moments(const double pivot[2], double out[9]);
  /// beluga::cluster_based_estimate (algorithm/cluster_based_estimation.hpp:415-432); cluster_ids
  /// (optional, one per particle) is the parity hook for ParticleClusterizer::operator() (:304-316).
  int cluster_estimate(const bb200_cluster_param& p, bb200_estimate* out, uint32_t* cluster_ids, uint64_t ids_capacity, uint32_t* n_cells,
                       uint32_t* n_clusters);
  /// Device-side histogram of the particle cloud over spatial-hash buckets (beluga_ros/particle_cloud.hpp:197-210): one
  /// record per occupied bucket in first-occurrence order.  bins may be null (count only).
  int particle_histogram(double linear_resolution, double angular_resolution, bb200_cluster_cell* bins, uint64_t capacity, uint64_t* n_bins,
                         double* max_bin_weight);
  /// `count` states drawn by weight without touching the set (views::sample | take_exactly, particle_cloud.hpp:141-147).
  int sample_states(uint64_t count, uint32_t step, double* states_out);
  static void estimate_from_moments_static(const double m[kMomentCount], const double pivot[2], bb200_estimate* out);

  /// Fused single-GPU step: propagate | reweight | normalize | resample | estimate with one
  /// host synchronisation at the end (the composition Amcl::update performs every step).
  int step_resample(const bb200_motion_sampling& sampling, uint32_t step, const double* points_xy, uint64_t n_points,
                    const bb200_resample_opts& o, bb200_estimate* est, double* weight_sum, uint64_t* new_size);

  // Stream-ordered variants for callers that interleave collectives on the same stream (sharded
  // filters): nothing here synchronises with the host or reads results back.
  int set_stream(void* stream);
  int enqueue_propagate_reweight(const bb200_motion_sampling* sampling, uint32_t step, const double* points_xy, uint64_t n_points);
  int enqueue_build_cdf();                      // exponent from the wmax in the device scalars (all-reduced by the caller)
  int enqueue_resample_range(const bb200_resample_opts& o, uint64_t global_total, uint64_t cdf_offset, uint64_t slot_begin, uint64_t slot_end);
  int enqueue_adopt(uint64_t n);
  // Peer-to-peer redistribution (one process per GPU; CUDA IPC): handles of the two state buffers,
  // mapping of the peers' buffers, and the fused produce-and-store step.
  int ipc_handles(void* out128);
  int open_peers(int world, int rank, const void* handles /* world x 128 bytes */);
  int enqueue_resample_push(const bb200_resample_opts& o, uint64_t global_total, uint64_t cdf_offset, uint64_t slot_begin, uint64_t slot_end,
                            uint64_t shard, const double pivot[2]);
  int enqueue_resample_push_device(const bb200_resample_opts& o, const uint64_t* rank_totals_device, int rank, int world, uint64_t shard,
                                   const double pivot[2]);
  int enqueue_flip_adopt(uint64_t n);
  int enqueue_reduce_moments();

  // ---- one filter over several shards, exchanges through peer memory (no NCCL, no host round trip) ----
  /// CUDA IPC handles of the two state buffers and the mail block (3 x 64 bytes) for peers in OTHER processes.
  int export_shard(void* out256);
  /// KLD-adaptive resampling on shards needs one hash array of the WHOLE filter's slot count on every rank, exported
  /// with the other buffers: call before export_shard / join_shards_local.
  int enable_shard_kld();
  /// Map the peers' buffers from their exported handles (world x 256 bytes, rank order); one process per GPU.
  int join_shards_ipc(int world, int rank, const void* handles);
  /// Shards living in THIS process (one per device, or several on one device): direct pointers, peer access enabled
  /// between distinct devices.  filters[r] becomes rank r.
  static int join_shards_local(Filter* const* filters, int world);
  int shard_world() const { return peer_world_; }
  /// Unmaps the peers' buffers (before any rank frees its own: CUDA IPC requires importers to close first).
  int leave_shards();
  /// The fused resampling step (propagate | reweight | normalize | resample | estimate), any number of shards.
  /// step_begin validates and stages the inputs; step_phase(1..4) only enqueues; step_end synchronises once.
  /// A driver holding several shards in one thread enqueues phase k of every shard before phase k+1 of any.
  int step_begin(const bb200_motion_sampling& sampling, uint32_t step, const double* points_xy, uint64_t n_points, const bb200_resample_opts& o,
                 bool resample_planned);
  enum StepPhase : int {
    kPhaseReweight = 1,   // begin_step | propagate | schedule | reweight                      (posts the shard's largest weight)
    kPhaseCdf = 2,        // common exponent | fixed-point CDF                                 (posts the shard's total)
    kPhaseResample = 3,   // resample (+ redistribution over peer memory) | moments of the new set (posts them)
    kPhaseFinish = 4,     // global moments, read-back
    kPhaseNormalize = 5   // instead of kPhaseResample on a step that keeps its particles: w /= S | moments
  };
  int step_phase(int phase);
  void step_abort();
  // KLD on shards (views/take_while_kld.hpp:72-137 over the globally ordered candidate stream).  After kPhaseCdf:
  //   step_totals            the ranks' fixed-point totals (also enqueues their exchange)
  //   kld_sharded_candidates hashes of this rank's candidates among the slots [begin, end) -> every rank's hash array
  //   kld_sharded_count      (after all ranks' candidates) distinct-bucket count over the window, on every rank alike
  //   kld_sharded_read       first failing count (~0: none in this window), new buckets of the window
  //   step_set_kld_accepted  the count the loop settled on; kPhaseResample then produces exactly those slots
  int step_totals(uint64_t* rank_totals, int* exponent);
  int kld_sharded_candidates(uint64_t begin, uint64_t end);
  int kld_sharded_count(uint64_t begin, uint64_t end, uint64_t k_before);
  int kld_sharded_read(uint64_t* cutoff, uint64_t* new_buckets);
  void step_set_kld_accepted(uint64_t accepted);
  bool shard_kld() const { return shard_kld_; }
  uint64_t global_size() const { return peer_world_ > 1 ? global_size_ : n_; }
  uint64_t first_index() const { return first_index_; }
  const double* pivot() const { return pivot_; }
  /// Synchronises and closes the batch of phases enqueued so far.  After kPhaseNormalize a caller may still run
  /// kPhaseResample + kPhaseFinish + step_end (selective resampling: the decision needs the effective sample size).
  int step_end(bb200_estimate* est, double* weight_sum, uint64_t* new_size, double* sum_sq);
  int enqueue_moments(const double pivot[2]);   // raw moments stay in the device result block

  int synchronize();
  int device_pointer(int which, void** ptr, uint64_t* bytes);

  void set_timing(bool on) { timing_ = on; }
  void clear_timings() { timings_.clear(); }
  int last_timings(const char** names, float* ms, int capacity) const;
  uint64_t launch_count() const { return launches_; }
  /// False when the last CDF was built from a weight set without any positive finite weight (uniform CDF substituted).
  bool last_weights_valid() const { return weights_valid_; }
  const char* last_error() const { return error_.c_str(); }
  void record_error(const std::string& message) const { error_ = message; }  // the C-ABI exception guard
  int fail_with(int status, const std::string& message) { return fail(status, message); }
  bool ok() const { return created_; }
  int create_status() const { return create_status_; }

 private:
  int fail(int status, const std::string& message);
  int check(cudaError_t e, const char* what);
  int upload_points(const double* points_xy, uint64_t n_points);
  int enqueue_propagate_reweight(const MotionSampling* sampling, uint32_t step, bool do_reweight, uint64_t n_points, bool counters_reset);
  bool predict_schedule(const MotionSampling& s, Schedule* grid) const;
  int ensure_cdf_ready();
  int resample_kld(const bb200_resample_opts& o, uint64_t* accepted);
  ResampleArgs make_resample_args(const bb200_resample_opts& o, uint64_t slot_begin, uint64_t slot_end, bool with_hashes) const;
  void estimate_from_moments(const double m[kMomentCount], bb200_estimate* out) const;
  void mark(const char* name);  // timing: record an event before the next kernel
  void finish_marks();
  bool use_device() const;

  struct StepContext {
    bool active{false};
    MotionSampling sampling{};
    uint32_t step{0};
    uint64_t n_points{0};
    bb200_resample_opts opts{};
    uint32_t partial_rows{1};
    bool resampled{false};   // kPhaseResample ran in the batch being closed
    bool normalized{false};  // kPhaseNormalize ran (weights are normalised, the CDF stays valid)
    bool resample_planned{false};  // kPhaseResample follows kPhaseCdf directly (every_n fired, no ESS decision in between)
    bool weights_filled{false};
    bool totals_exchanged{false};
    int poll_seq{0};  // != 0: the resample kernel stores this ticket to the pinned summary when the step is complete
    uint64_t kld_accepted{0};  // KLD on shards: particle count of the new set (0: fixed size)  // the ranks' totals of this step are in shard_totals_
    unsigned long long total{0};
    int exponent{0};
  };
  StepContext step_{};
  int enqueue_exchange(int kind, bool post, bool wait);
  void release_peers();

  bb200_filter_config config_{};
  bool created_{false};
  int create_status_{BB200_OK};
  mutable std::string error_;
  cudaStream_t stream_{nullptr};
  bool owns_stream_{true};
  int peer_world_{0}, peer_rank_{0};
  uint32_t pushed_blocks_{1};
  Pose2* peer_states_[2][8]{};  // [buffer][rank]: the peers' ping-pong state buffers mapped into this process
  ShardMail* mail_{nullptr};                 // this rank's mail block
  ShardMail* peer_mail_[kMaxShards]{};       // every rank's mail block as seen from here ([rank] = mail_)
  bool peers_ipc_{false};                    // the peer pointers came from cudaIpcOpenMemHandle
  bool split_posts_{false};                  // post and wait as separate launches (several shards enqueued by one thread)
  unsigned long long epoch_{0};              // sequence number of the next exchange (all ranks count alike)
  unsigned long long* shard_totals_{nullptr};  // device: the ranks' fixed-point totals of this step
  StepSummary* summary_{nullptr};            // device
  StepSummary* summary_host_{nullptr};       // pinned

  // particle set (ping-pong states for the resample gather)
  uint64_t capacity_{0}, n_{0};
  uint64_t first_index_{0};   // global index of local particle 0 (moves when a KLD-sized filter re-splits its particles)
  uint64_t global_size_{0};   // particles over all shards
  Pose2* states_[2]{nullptr, nullptr};
  int cur_{0};
  double* weights_{nullptr};
  unsigned long long* cdf_{nullptr};
  long long* ancestors_{nullptr};
  uint64_t ancestors_n_{0};
  unsigned long long* hashes_{nullptr};
  bool cdf_valid_{false};
  bool weights_valid_{true};

  // scratch
  Scalars* scalars_{nullptr};
  Scalars* scalars_host_{nullptr};  // pinned
  unsigned long long* tile_state_{nullptr};
  uint32_t tile_capacity_{0};
  double* partials_{nullptr};
  uint32_t partials_rows_{0};
  double* results_{nullptr};       // device: kMomentCount + extras
  double* results_host_{nullptr};  // pinned

  // clusterizer scratch (allocated on first use)
  int ensure_cluster_scratch(uint32_t cells);
  int cell_records(double linear_resolution, double angular_resolution, std::vector<HostCell>* host);
  ClusterScratch cluster_{};

  // KLD scratch
  unsigned long long* kld_keys_{nullptr};
  unsigned int* kld_vals_{nullptr};
  uint64_t kld_table_size_{0};
  uint32_t* kld_flags_{nullptr};
  uint32_t* kld_scan_{nullptr};
  unsigned long long* kld_tile_state_{nullptr};
  unsigned long long* kld_hashes_global_{nullptr};          // KLD on shards: spatial hash of every candidate slot (all ranks hold all)
  unsigned long long* peer_kld_hashes_[kMaxShards]{};
  bool shard_kld_{false};

  // measurement
  double* points_{nullptr};
  double* points_host_{nullptr};  // pinned staging
  uint64_t points_capacity_{0};
  double points_radius_{0.0};
  double points_mean_range_{1.0};

  // execution schedule (pose-sorted processing order of the reweight kernels)
  static constexpr uint64_t kScheduleMinParticles = 32768;
  bool schedule_enabled_{true};
  bool tiled_layout_{true};
  bool fixed_lookup_{true};
  bool param_points_{true};
  double schedule_per_bin_{4.0};
  double schedule_lever_{1.0};
  double schedule_x_split_{8.0};
  bool schedule_equal_mass_{true};
  Schedule* sched_{nullptr};
  uint32_t* bins_{nullptr};
  uint2* bin_rank_{nullptr};
  uint32_t* perm_{nullptr};
  uint32_t* counters_{nullptr};
  unsigned long long* sched_tiles_{nullptr};

  // maps
  int sensor_{-1};
  std::vector<float> field_host_;
  double* table_{nullptr};
  double* tiled_{nullptr};
  double* bordered_{nullptr};
  uint64_t bordered_bytes_{0};
  bool prefetch_table_{true};
  FieldView field_{};
  int8_t* occupancy_{nullptr};
  uint8_t* free_distance_{nullptr};
  uint8_t* free_padded_{nullptr};
  OccupancyView occupancy_view_{};
  BeamParams beam_{};
  double2* beam_eta_{nullptr};
  uint32_t* beam_hits_{nullptr};   // two-pass beam model: hit word per (beam, particle) of one pass
  uint64_t beam_hits_words_{0};
  bool beam_two_pass_{true};
  bool beam_eta_table_{true};
  uint32_t* free_cells_{nullptr};
  uint64_t n_free_{0};
  int grid_width_{0}, grid_height_{0};
  double grid_resolution_{1.0};
  Pose2 grid_origin_{1.0, 0.0, 0.0, 0.0};

  double pivot_[2]{0.0, 0.0};
  bb200_estimate cloud_{};     // last estimate, when it describes the raw (unit-weight) cloud
  bool cloud_known_{false};
  bool predict_schedule_{true};
  bool poll_completion_{true};
  int step_seq_{0};

  // timing
  bool timing_{false};
  struct Mark {
    const char* name;
    cudaEvent_t event;
  };
  std::vector<Mark> marks_;
  std::vector<cudaEvent_t> event_pool_;
  size_t events_used_{0};
  std::vector<std::pair<const char*, float>> timings_;
  uint64_t launches_{0};
};

/// Symmetric 3x3 covariance -> V * sqrt(Lambda) (multivariate_normal_distribution.hpp:109-126).
/// Returns false with a message for non-symmetric / negative-eigenvalue input.
bool normal_transform(const double cov[9], double transform[9], std::string* error);

}  // namespace bb200
