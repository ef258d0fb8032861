/*
 * beluga_b200 -- C ABI of the B200-native MCL particle-filter update.
 *
 * This is the drop-in boundary for ONE path of Ekumen-OS/beluga: the per-step
 * particle-filter update `beluga::Amcl::update`
 * (beluga/include/beluga/algorithm/amcl_core.hpp:165-201) and the models, actions, views and
 * reductions it composes.  The reference is a header-only C++17 template library without any
 * FFI; the C++ adaptors in include/beluga_b200/ give these entry points the reference's own
 * MotionModel / SensorModel / Amcl shapes.  Each entry point cites the
 * reference interface it replaces (paths relative to /root/reference/beluga/include/beluga).
 *
 * Conventions
 *   - Poses are `double[4]` in Sophus::SE2d::data() order {cos, sin, x, y}
 *     (what estimation.hpp:448-452 relies on); particle states are arrays of such quadruples.
 *   - Every function returns BB200_OK (0) or a negative bb200_status; bb200_last_error() gives
 *     the message of the last failure on that context.  Nothing throws across this boundary.
 *   - All pointers are HOST pointers unless the name says `_device`.  A context owns its
 *     device buffers and one CUDA stream; calls on one context must come from one thread at a
 *     time (same contract as the reference: no internal locking, amcl_node.cpp:581-603).
 *   - There is NO CPU fallback: creating a context without a usable CUDA device fails.
 */
#ifndef BELUGA_B200_H_
#define BELUGA_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define BB200_ABI_VERSION 2

typedef enum bb200_status {
  BB200_OK = 0,
  BB200_ERR_INVALID_ARGUMENT = -1,
  BB200_ERR_CUDA = -2,
  BB200_ERR_NO_DEVICE = -3,
  BB200_ERR_STATE = -4, /* call sequence error: no map, no particles, ... */
  BB200_ERR_CAPACITY = -5
} bb200_status;

typedef struct bb200_filter bb200_filter; /* device-resident particle set + kernels        */
typedef struct bb200_amcl bb200_amcl;     /* beluga::Amcl control flow on top of a filter   */

/* ---------------------------------------------------------------------------------------------
 * Parameter blocks (plain doubles/ints; same members and defaults as the reference structs)
 * ------------------------------------------------------------------------------------------- */

/* beluga::DifferentialDriveModelParam -- motion/differential_drive_model.hpp:40-68 */
typedef struct bb200_diff_drive_param {
  double rotation_noise_from_rotation;       /* alpha1 */
  double rotation_noise_from_translation;    /* alpha2 */
  double translation_noise_from_translation; /* alpha3 */
  double translation_noise_from_rotation;    /* alpha4 */
  double distance_threshold;                 /* default 0.01 */
} bb200_diff_drive_param;

/* The three std::normal_distribution parameter sets that sampling_fn_2d derives from a control
 * action -- motion/differential_drive_model.hpp:141-154. */
typedef struct bb200_diff_drive_sampling {
  double rot1_mean, rot1_std;
  double trans_mean, trans_std;
  double rot2_mean, rot2_std;
} bb200_diff_drive_sampling;

/* The motion models of beluga_ros::Amcl::motion_model_variant (beluga_ros/include/beluga_ros/amcl.hpp:108-111). */
typedef enum bb200_motion_model {
  BB200_MOTION_DIFFERENTIAL = 0,    /* motion/differential_drive_model.hpp */
  BB200_MOTION_OMNIDIRECTIONAL = 1, /* motion/omnidirectional_drive_model.hpp */
  BB200_MOTION_STATIONARY = 2       /* motion/stationary_model.hpp */
} bb200_motion_model;

/* DifferentialDriveModelParam / OmnidirectionalDriveModelParam (omnidirectional_drive_model.hpp:36-68)
 * in one block; strafe_noise_from_translation is only read by the omnidirectional model. */
typedef struct bb200_motion_param {
  int model; /* bb200_motion_model */
  double rotation_noise_from_rotation;
  double rotation_noise_from_translation;
  double translation_noise_from_translation;
  double translation_noise_from_rotation;
  double strafe_noise_from_translation;
  double distance_threshold; /* default 0.01 */
} bb200_motion_param;

/* What a motion model derives from one control action: three normal distributions, drawn per
 * particle in this order, and (omnidirectional only) the deterministic first rotation.
 *   differential:     (rot1, trans, rot2)            state * SE2(rot1, 0) * SE2(rot2, (trans, 0))
 *   omnidirectional:  (rotation, translation, strafe) state * SE2(first, 0) * SE2(SO2(rotation) * first^-1, (translation, -strafe))
 *   stationary:       (theta, x, y) ~ N(0, 0.02)      state * SE2(theta, (x, y)) */
typedef struct bb200_motion_sampling {
  int model;
  double mean[3];
  double stddev[3];
  double first_rotation[2]; /* {cos, sin} */
} bb200_motion_sampling;

/* beluga::LikelihoodFieldModelBaseParam -- sensor/likelihood_field_model_base.hpp:42-64 */
typedef struct bb200_likelihood_field_param {
  double max_obstacle_distance; /* default 100.0 */
  double max_laser_distance;    /* default 2.0 */
  double z_hit;                 /* default 0.5 */
  double z_random;              /* default 0.5 */
  double sigma_hit;             /* default 0.2 */
  int model_unknown_space;      /* default 0 */
  int only_obstacle_boundaries; /* default 0 */
} bb200_likelihood_field_param;

/* beluga::BeamModelParam -- sensor/beam_model.hpp:43-58 */
typedef struct bb200_beam_param {
  double z_hit, z_short, z_max, z_rand, sigma_hit, lambda_short, beam_max_range;
} bb200_beam_param;

/* Occupancy grid view -- the OccupancyGrid2 named requirement (sensor/data/occupancy_grid.hpp)
 * with the ROS trinary value traits (beluga_ros/include/beluga_ros/occupancy_grid.hpp:48-64):
 * 0 free, 100 occupied, -1 unknown; row-major, index = yi*width + xi (linear_grid.hpp:73-75). */
typedef struct bb200_occupancy_grid {
  const int8_t* cells;
  int32_t width, height;
  double resolution;
  double origin[4]; /* grid.origin() as {cos, sin, x, y} */
} bb200_occupancy_grid;

typedef enum bb200_sensor_model {
  BB200_SENSOR_LIKELIHOOD_FIELD = 0,      /* sensor/likelihood_field_model.hpp:69-90: 1 + sum pz^3 */
  BB200_SENSOR_LIKELIHOOD_FIELD_PROB = 1, /* sensor/likelihood_field_prob_model.hpp:69-90: exp(sum log pz) */
  BB200_SENSOR_BEAM = 2                   /* sensor/beam_model.hpp:104-150 */
} bb200_sensor_model;

typedef enum bb200_resample_scheme {
  BB200_RESAMPLE_MULTINOMIAL = 0, /* views/sample.hpp:128-135 in counter-RNG form */
  BB200_RESAMPLE_SYSTEMATIC = 1   /* not in the reference; low-variance comb over the same CDF */
} bb200_resample_scheme;

/* Pose estimate -- std::pair<Sophus::SE2d, Sophus::Matrix3d> of algorithm/estimation.hpp:436-475 */
typedef struct bb200_estimate {
  double mean[4]; /* {cos, sin, x, y}, rotation renormalised */
  double cov[9];  /* row-major 3x3 over (x, y, theta) */
} bb200_estimate;

/* ---------------------------------------------------------------------------------------------
 * Library
 * ------------------------------------------------------------------------------------------- */

int bb200_abi_version(void);
/* Number of CUDA devices visible; 0 when there is none (then every create call fails). */
int bb200_device_count(void);
/* Message of the last failure of a create call on this thread. */
const char* bb200_create_error(void);

/* ---------------------------------------------------------------------------------------------
 * bb200_filter -- the particle set on the device and the per-step kernels.
 * Replaces beluga::TupleVector<std::tuple<SE2d, Weight>> (containers/tuple_vector.hpp:50-223)
 * plus the range adaptors that iterate it.
 * ------------------------------------------------------------------------------------------- */

typedef struct bb200_filter_config {
  int device;             /* CUDA device ordinal */
  uint64_t capacity;      /* maximum number of LOCAL particles (max_particles, or the shard size) */
  uint64_t seed;          /* counter-RNG key (Philox4x32-10) */
  /* Sharding (single GPU: rank 0 of 1).  Global particle index = first_index + local index. */
  uint64_t first_index;   /* global index of local particle 0 */
  uint64_t global_count;  /* total particles over all ranks (0: same as local) */
  int record_ancestors;   /* keep the resample indices for bb200_filter_ancestors (parity hook) */
} bb200_filter_config;

int bb200_filter_create(const bb200_filter_config* config, bb200_filter** out);
void bb200_filter_destroy(bb200_filter* f);
const char* bb200_last_error(const bb200_filter* f);

/* Sensor models.  LikelihoodFieldModel{params, grid} (likelihood_field_model.hpp:58-59): builds
 * the likelihood field on the host exactly as make_likelihood_field
 * (likelihood_field_model_base.hpp:130-185, distance_map.hpp:55-98) and uploads it.
 * `prob` selects LikelihoodFieldProbModel.  Also the target of update_map() (:113-116). */
int bb200_filter_set_likelihood_field_map(bb200_filter* f, const bb200_likelihood_field_param* p, const bb200_occupancy_grid* grid, int prob);
/* BeamSensorModel{params, grid} (beam_model.hpp:96) / update_map (:156). */
int bb200_filter_set_beam_map(bb200_filter* f, const bb200_beam_param* p, const bb200_occupancy_grid* grid);
/* likelihood_field() accessor (likelihood_field_model_base.hpp:102): width*height floats. */
int bb200_filter_get_likelihood_field(const bb200_filter* f, float* out, uint64_t capacity);

/* Particle set access.  weights == NULL means 1.0 (make_from_state, particle_traits.hpp:105). */
int bb200_filter_set_particles(bb200_filter* f, const double* states, const double* weights, uint64_t n);
int bb200_filter_size(const bb200_filter* f, uint64_t* n);
/* particles() (amcl_core.hpp:128): copies min(n, capacity) particles out; either may be NULL. */
int bb200_filter_get_particles(bb200_filter* f, double* states, double* weights, uint64_t capacity);
/* Amcl::initialize(pose, covariance) (amcl_core.hpp:145-147) with
 * MultivariateNormalDistribution<SE2d> (random/multivariate_normal_distribution.hpp:96-126):
 * n particles ~ N(mean {x, y, theta}, cov 3x3), weights 1. */
int bb200_filter_initialize_normal(bb200_filter* f, const double mean_xytheta[3], const double cov[9], uint64_t n);
/* Global localisation: n particles from MultivariateUniformDistribution<SE2d, OccupancyGrid>
 * (random/multivariate_uniform_distribution.hpp:127-161) over the map set by bb200_filter_set_*_map: a free cell drawn
 * uniformly, its centroid in the global frame, yaw ~ U[-pi, pi); weights 1.  Counter RNG stream 6 at step 0,
 * keyed by the global particle index (same generator as the recovery injection of bb200_filter_resample). */
int bb200_filter_initialize_uniform(bb200_filter* f, uint64_t n);

/* actions::propagate(model(control)) (actions/propagate.hpp:57-79) for DifferentialDriveModel:
 * per particle 3 normals (counter RNG keyed by seed / global index / step) and the SE2 compose of
 * differential_drive_model.hpp:156-163. */
int bb200_filter_propagate(bb200_filter* f, const bb200_motion_sampling* s, uint32_t step);
/* actions::reweight(sensor_model(points)) (actions/reweight.hpp:54-60): w *= L(state). */
int bb200_filter_reweight(bb200_filter* f, const double* points_xy, uint64_t n_points);
/* Fused propagate | reweight (one pass over the particle set). */
int bb200_filter_propagate_reweight(bb200_filter* f, const bb200_motion_sampling* s, uint32_t step, const double* points_xy, uint64_t n_points);

/* Weight statistics of the local shard after reweight: the largest weight and, once an exponent
 * is fixed, the fixed-point total.  Single GPU: bb200_filter_normalize does all of it. */
int bb200_filter_max_weight(bb200_filter* f, double* wmax);
/* Fixes the quantisation exponent from the GLOBAL largest weight (all ranks pass the same value),
 * builds the local inclusive fixed-point CDF and returns the local total. */
int bb200_filter_build_cdf(bb200_filter* f, double global_wmax, uint64_t* local_total, int* exponent);
/* actions::normalize (actions/normalize.hpp:54-85): w /= S with S = global_total * 2^-exponent.
 * Also returns sum((w/S)^2) over the local shard (effective_sample_size.hpp:46-59). */
int bb200_filter_normalize_by(bb200_filter* f, uint64_t global_total, double* local_sum_sq);
/* Single-GPU convenience: max -> cdf -> normalize.  Returns the factor S and sum of squares. */
int bb200_filter_normalize(bb200_filter* f, double* factor, double* sum_sq);

/* views::sample | random_intersperse | take_while_kld | actions::assign
 * (views/sample.hpp:128-153, views/random_intersperse.hpp:93-100, views/take_while_kld.hpp:72-137,
 *  actions/assign.hpp:56-63) in counter-RNG form. */
typedef struct bb200_resample_opts {
  int scheme;                       /* bb200_resample_scheme */
  uint32_t step;
  uint64_t min_particles;           /* KLD lower bound (== max_particles disables KLD) */
  uint64_t max_particles;           /* number of output slots */
  double kld_epsilon, kld_z;
  double spatial_resolution[3];     /* spatial_hash<SE2d>{x, y, theta} (spatial_hash.hpp:160-197) */
  double random_state_probability;  /* recovery injection probability */
} bb200_resample_opts;
int bb200_filter_resample(bb200_filter* f, const bb200_resample_opts* o, uint64_t* new_size);
/* Sharded filter (one rank per GPU, contiguous global index ranges):
 * produce the output slots [slot_begin, slot_end) -- those whose systematic-comb position lies in this
 * rank's span [cdf_offset, cdf_offset + local_total) of the global fixed-point CDF -- in slot order
 * into the staging state buffer (bb200_filter_device_pointer(f, 3)).  The caller moves them to the
 * ranks that own the slots (all-to-all) and then calls bb200_filter_adopt on the receivers. */
int bb200_filter_resample_range(bb200_filter* f, const bb200_resample_opts* o, uint64_t global_total, uint64_t cdf_offset,
                                uint64_t slot_begin, uint64_t slot_end);
/* Declare that n particle states have been written into the current (from_staging = 0) or staging
 * (from_staging = 1) state buffer: they become the particle set with weights 1. */
int bb200_filter_adopt(bb200_filter* f, uint64_t n, int from_staging);
/* Host helpers of the sharded step: the systematic comb (stride = total / slots, offset in [0, stride)
 * from the counter RNG) and beluga::estimate from globally summed raw moments (bb200_filter_moments). */
int bb200_systematic_comb(uint64_t seed, uint32_t step, uint64_t global_total, uint64_t total_slots, uint64_t* stride, uint64_t* offset);
int bb200_estimate_from_moments(const double moments[9], const double pivot_xy[2], bb200_estimate* out);
/* Stream-ordered variants for sharded callers that put collectives on the same CUDA stream
 * (bb200_filter_set_stream, e.g. torch's current stream): they only enqueue work -- no host
 * synchronisation, no read-back; intermediate scalars stay on the device (device pointers 4 and 5).
 *   enqueue_propagate_reweight: begin_step | propagate | schedule | reweight; the shard's largest weight is left in scalars.wmax_bits
 *   enqueue_build_cdf:          exponent from scalars.wmax_bits (all-reduce it with MAX first), fixed-point scan; total in scalars.total
 *   enqueue_resample_range / enqueue_adopt / enqueue_moments: as the synchronous calls above. */
int bb200_filter_set_stream(bb200_filter* f, void* cuda_stream);
int bb200_filter_enqueue_propagate_reweight(bb200_filter* f, const bb200_motion_sampling* s, uint32_t step, const double* points_xy, uint64_t n_points);
int bb200_filter_enqueue_build_cdf(bb200_filter* f);
int bb200_filter_enqueue_resample_range(bb200_filter* f, const bb200_resample_opts* o, uint64_t global_total, uint64_t cdf_offset,
                                        uint64_t slot_begin, uint64_t slot_end);
int bb200_filter_enqueue_adopt(bb200_filter* f, uint64_t n);
int bb200_filter_enqueue_moments(bb200_filter* f, const double pivot_xy[2]);
/* Fused resample + redistribution over NVLink peer memory (one process per GPU, <= 8 ranks, equal
 * shards).  Each rank exports CUDA IPC handles of its two state buffers (128 bytes), the handles of all
 * ranks are gathered by the caller and opened with bb200_filter_open_peers.  bb200_filter_enqueue_resample_push
 * then produces this rank's slot range [slot_begin, slot_end) and stores every state straight into the
 * staging buffer of the rank that owns the slot (peer stores), accumulating the raw moments of what it
 * produced; bb200_filter_enqueue_reduce_moments leaves them in the result block for the caller's
 * all-reduce -- which is also the barrier after which every rank calls bb200_filter_enqueue_flip_adopt. */
int bb200_filter_ipc_handles(bb200_filter* f, void* out128);
int bb200_filter_open_peers(bb200_filter* f, int world, int rank, const void* handles);
int bb200_filter_enqueue_resample_push(bb200_filter* f, const bb200_resample_opts* o, uint64_t global_total, uint64_t cdf_offset,
                                       uint64_t slot_begin, uint64_t slot_end, uint64_t shard, const double pivot_xy[2]);
/* The same with the bookkeeping on the device: rank_totals_device points at the `world` all-gathered
 * fixed-point totals (uint64, device memory, rank order).  The kernel derives the global total, this
 * rank's CDF offset and (systematic) its slot range from them, so the host need not read the totals
 * back before launching -- one host synchronisation per step instead of two. */
int bb200_filter_enqueue_resample_push_device(bb200_filter* f, const bb200_resample_opts* o, const uint64_t* rank_totals_device, int rank,
                                              int world, uint64_t shard, const double pivot_xy[2]);
int bb200_filter_enqueue_reduce_moments(bb200_filter* f);
int bb200_filter_enqueue_flip_adopt(bb200_filter* f, uint64_t n);
/* Ancestor index of every particle produced by the last resample (-1: injected random state). */
int bb200_filter_ancestors(bb200_filter* f, int64_t* out, uint64_t capacity);
/* The local fixed-point CDF built by the last build_cdf / normalize (parity hook). */
int bb200_filter_cdf(bb200_filter* f, uint64_t* out, uint64_t capacity);

/* beluga::estimate(states, weights) (algorithm/estimation.hpp:436-475). */
int bb200_filter_estimate(bb200_filter* f, bb200_estimate* out);
/* beluga::ParticleClusterizerParam (algorithm/cluster_based_estimation.hpp:259-276), same defaults
 * when filled by bb200_cluster_param_default. */
typedef struct bb200_cluster_param {
  double linear_hash_resolution;  /* 0.20 */
  double angular_hash_resolution; /* 0.524 */
  double weight_cap_percentile;   /* 0.90 */
} bb200_cluster_param;
void bb200_cluster_param_default(bb200_cluster_param* p);
/* beluga::cluster_based_estimate(states, weights, parameters) (cluster_based_estimation.hpp:415-432),
 * what beluga_ros::Amcl::update returns (beluga_ros/src/amcl.cpp:125): mean and covariance of the
 * heaviest cluster with more than one particle, or of the whole set when there is none.
 * Optional outputs (may be NULL): cluster_ids[i] = ParticleClusterizer::operator() (:304-316) for
 * particle i (ids_capacity >= particle count), the number of occupied cells and of clusters. */
int bb200_filter_cluster_estimate(bb200_filter* f, const bb200_cluster_param* p, bb200_estimate* out, uint32_t* cluster_ids,
                                  uint64_t ids_capacity, uint32_t* n_cells, uint32_t* n_clusters);
/* The per-CELL half of the cluster-based estimate, host only (no device needed): what
 * bb200_filter_cluster_estimate runs on the cell records its kernels produce.  cells[] lists the occupied
 * spatial-hash cells in the order in which the particle sequence first touches them
 * (clusterizer_detail::make_cluster_map, cluster_based_estimation.hpp:141-161).  Outputs: the cluster id of
 * every cell (assign_clusters, :205-253), the number of clusters, whether a cluster with more than one
 * particle exists and, if so, the id of the heaviest (cluster_based_estimate, :420-431), and the raw moments
 * to estimate from (that cluster's, else the whole set's). */
typedef struct bb200_cluster_cell {
  double representative[4]; /* state of the first particle in the cell {cos, sin, x, y} */
  uint64_t hash;            /* spatial_hash<SE2d>{linear, linear, angular}(representative) */
  uint32_t first_index;     /* index of that first particle */
  uint32_t count;           /* particles in the cell */
  double weight;            /* sum of their weights, in particle order */
  double moments[9];        /* raw moments of the cell's particles (layout of bb200_filter_moments) */
} bb200_cluster_cell;
int bb200_cluster_select_host(const bb200_cluster_cell* cells, uint64_t n_cells, uint64_t n_particles, const bb200_cluster_param* p,
                              uint32_t* cluster_of_cell, uint32_t* n_clusters, int* found, uint32_t* best, double moments_out[9]);
/* ---- output side: what beluga_ros publishes after a step, without moving N x 40 bytes to the host ----
 * Device-side histogram of the particle cloud over spatial_hash<SE2d>{linear, linear, angular} buckets -- the
 * unordered_map pass of assign_particle_cloud(particles, linear_resolution, angular_resolution, MarkerArray)
 * (beluga_ros/include/beluga_ros/particle_cloud.hpp:197-210): one bin per occupied bucket, in the order in which the
 * particle sequence first touches them; representative = the first particle of the bucket, weight = the bucket's weights
 * added in particle order.  After a resample the million particles are copies of a few thousand candidates: the
 * publish costs a few-KB read-back.  (The reference's map additionally splits a bucket when two of its states are
 * not detail::almost_equal_to; with its 1 mm / 1 mrad defaults a bucket holds exact copies, and that refinement is not
 * reproduced.)  bins may be NULL to query the count.  max_bin_weight: max(1e-3, heaviest bin) (:197,205-207). */
int bb200_filter_particle_histogram(bb200_filter* f, double linear_resolution, double angular_resolution, bb200_cluster_cell* bins, uint64_t capacity,
                                    uint64_t* n_bins, double* max_bin_weight);
/* assign_particle_cloud(particles, size, PoseArray) (particle_cloud.hpp:129-147): `count` states drawn by weight
 * (views::sample | take_exactly) -- multinomial counter-RNG draws of `step` -- into states_out (count x {cos, sin, x, y});
 * the particle set is not modified. */
int bb200_filter_sample_states(bb200_filter* f, uint64_t count, uint32_t step, double* states_out);
/* Host only.  The marker geometry of assign_particle_cloud(..., MarkerArray) (particle_cloud.hpp:212-294) from the
 * histogram bins: 2 vertices per bin for the LINE_LIST "bodies" marker, 3 per bin for the TRIANGLE_LIST "heads"
 * marker, each with its RGBA colour (detail::alphaHueToRGBA, :56-70), and arrow_bodies.scale.x. */
typedef struct bb200_marker_vertex {
  double x, y, z;
  float r, g, b, a;
} bb200_marker_vertex;
int bb200_particle_cloud_markers(const bb200_cluster_cell* bins, uint64_t n_bins, bb200_marker_vertex* bodies, bb200_marker_vertex* heads,
                                 double* body_scale_x);
/* Host only.  assign_likelihood_field (beluga_ros/include/beluga_ros/likelihood_field.hpp:44-79): the likelihood field
 * (bb200_filter_get_likelihood_field) normalised to the [0, 100] int8 cells of a nav_msgs/OccupancyGrid. */
int bb200_likelihood_field_to_occupancy(const float* field, uint64_t n, int8_t* out);
/* Raw weighted moments of the local shard for a multi-rank estimate:
 * {sum w, sum w^2, sum w*cos, sum w*sin, sum w*dx, sum w*dy, sum w*dx^2, sum w*dx*dy, sum w*dy^2}
 * with (dx, dy) = (x, y) - pivot. */
int bb200_filter_moments(bb200_filter* f, const double pivot_xy[2], double out[9]);

/* Device-side timing (milliseconds, CUDA events on the filter's stream) of the kernels launched
 * since bb200_filter_clear_timings; names[i] are static strings.  Returns the count. */
int bb200_filter_set_timing(bb200_filter* f, int enabled);
int bb200_filter_clear_timings(bb200_filter* f);
int bb200_filter_last_timings(const bb200_filter* f, const char** names, float* ms, int capacity);
/* Total number of kernel launches issued by this filter since creation. */
uint64_t bb200_filter_launch_count(const bb200_filter* f);
/* Block until everything enqueued on the filter's stream has finished. */
int bb200_filter_synchronize(bb200_filter* f);
/* Raw device pointers (for torch.distributed / peer access plumbing): which = 0 states (double4),
 * 1 weights (double), 2 cdf (uint64), 3 staging states buffer (double4), 4 the scalar block
 * {u64 wmax_bits, u64 ticket, u64 total, i32 exponent, i32 valid, ...}, 5 the result block (9 raw moments). */
int bb200_filter_device_pointer(bb200_filter* f, int which, void** ptr, uint64_t* bytes);

/* ---------------------------------------------------------------------------------------------
 * bb200_amcl -- beluga::Amcl (algorithm/amcl_core.hpp:81-233) on top of a filter: update /
 * resample policies, rolling control window, recovery estimator, estimate.
 * ------------------------------------------------------------------------------------------- */

/* beluga::AmclParams (amcl_core.hpp:34-55) + spatial hasher resolutions + backend knobs. */
typedef struct bb200_amcl_param {
  double update_min_d;         /* 0.25 */
  double update_min_a;         /* 0.2 */
  uint64_t resample_interval;  /* 1 */
  int selective_resampling;    /* 0 */
  uint64_t min_particles;      /* 500 */
  uint64_t max_particles;      /* 2000 */
  double alpha_slow;           /* 0.001 */
  double alpha_fast;           /* 0.1 */
  double kld_epsilon;          /* 0.05 */
  double kld_z;                /* 3.0 */
  double spatial_resolution[3];/* spatial_hash<SE2d>: x, y, theta */
  int resample_scheme;         /* bb200_resample_scheme */
  uint64_t seed;
  int device;
  int record_ancestors;
  /* Sharding (both 0: single GPU).  max_particles is then the GLOBAL count; this rank holds
   * shard_capacity particles starting at global index shard_first_index. */
  uint64_t shard_first_index;
  uint64_t shard_capacity;
  /* > 0: views::random_intersperse runs with this probability on every resample instead of the recovery
   * estimator's output (which is identically 0 while the particle count is constant); 0: the estimator. */
  double recovery_probability_override;
} bb200_amcl_param;

/* What Amcl::update decided on the host for this step (policies, control window, recovery
 * estimator; amcl_core.hpp:166-190) -- lets a sharded caller run the device part with collectives
 * in between (see beluga_b200/distributed.py). */
typedef struct bb200_step_plan {
  int update;                         /* 0: std::nullopt, nothing to do */
  int resample;                       /* every_n fired */
  int needs_ess;                      /* selective resampling: resample only if ESS < N/2 */
  uint32_t step;
  double random_state_probability;
  bb200_motion_sampling sampling;
  bb200_resample_opts opts;
} bb200_step_plan;

typedef struct bb200_update_result {
  int updated;                 /* 0: std::nullopt (no motion / no particles) */
  int resampled;
  uint64_t n_particles;
  bb200_estimate estimate;
  double random_state_probability;
  double weight_sum;           /* normalisation factor S of this step */
  int weights_degenerate;      /* 1: no particle had a positive finite weight after the reweight (the reference would
                                  divide by zero in normalize.hpp:82); a uniform CDF was substituted, weights left as they were */
} bb200_update_result;

int bb200_amcl_create(const bb200_amcl_param* p, const bb200_diff_drive_param* motion, bb200_amcl** out);
void bb200_amcl_destroy(bb200_amcl* a);
const char* bb200_amcl_last_error(const bb200_amcl* a);
/* The filter owned by the driver (for set_*_map, get_particles, timings ...). */
bb200_filter* bb200_amcl_filter(bb200_amcl* a);
/* Amcl::initialize(pose, covariance) -- amcl_core.hpp:145-147 (max_particles samples). */
int bb200_amcl_initialize(bb200_amcl* a, const double mean_xytheta[3], const double cov[9]);
/* beluga_ros::Amcl::initialize_from_map() (beluga_ros/include/beluga_ros/amcl.hpp:192-209): max_particles samples of the
 * map distribution (uniform over the free cells), weights 1, next update forced. */
int bb200_amcl_initialize_from_map(bb200_amcl* a);
/* Amcl::initialize from explicit states (tests / custom distributions, amcl_core.hpp:131-137). */
int bb200_amcl_initialize_states(bb200_amcl* a, const double* states, const double* weights, uint64_t n);
/* Amcl::force_update -- amcl_core.hpp:204 */
void bb200_amcl_force_update(bb200_amcl* a);
/* Amcl::update(control_action, measurement) -- amcl_core.hpp:165-201.  `points_xy` is the
 * measurement_type std::vector<std::pair<double,double>> flattened (x0, y0, x1, y1, ...). */
int bb200_amcl_update(bb200_amcl* a, const double control_pose[4], const double* points_xy, uint64_t n_points, bb200_update_result* out);
/* Scan preprocessing in front of the path: what beluga_ros does between the sensor message
 * and Amcl::update -- beluga_ros::LaserScan (beluga_ros/include/beluga_ros/laser_scan.hpp:69-80:
 * take_evenly(max_beams) over ranges and angles, angle = float(angle_min + float(i) * angle_increment)),
 * BaseLaserScan::points_in_polar/cartesian_coordinates (beluga/sensor/data/laser_scan.hpp:64-91: drop
 * NaN and out-of-[min_range, max_range] readings, x = r cos a, y = r sin a) and the laser-to-base
 * transform of beluga_ros/src/amcl.cpp:57-62.  `laser_origin` is the 3x4 row-major [R | t] of the
 * laser frame in the base frame (NULL: identity); only x and y of the transformed point are kept. */
typedef struct bb200_laser_scan {
  const float* ranges;
  uint64_t n_ranges;
  float angle_min, angle_increment;
  double min_range, max_range; /* already combined with the message's range_min / range_max */
  uint64_t max_beams;          /* take_evenly count; 0 or >= n_ranges keeps every reading */
  const double* laser_origin;  /* 12 doubles or NULL */
} bb200_laser_scan;
/* Writes at most `capacity` points (x, y pairs) and their count. */
int bb200_scan_to_points(const bb200_laser_scan* scan, double* points_xy, uint64_t capacity, uint64_t* n_points);
/* take_evenly (beluga/views/take_evenly.hpp:118-145): the indices kept out of `size` elements. */
int bb200_take_evenly_indices(uint64_t size, uint64_t count, uint64_t* indices, uint64_t capacity, uint64_t* n_indices);
/* Amcl::update(base_pose_in_odom, laser_scan) -- beluga_ros/src/amcl.cpp:54-64. */
int bb200_amcl_update_scan(bb200_amcl* a, const double control_pose[4], const bb200_laser_scan* scan, bb200_update_result* out);

/* ---------------------------------------------------------------------------------------------
 * One filter over several GPUs.  Particles are split into equal contiguous global index
 * ranges ("shards"); map and scan are replicated; the counter RNG is keyed by the GLOBAL particle /
 * slot index and the CDF is an integer prefix sum, so the particle set does not depend on the number
 * of shards.  The three exchanges of a step (largest weight, fixed-point totals, raw moments) and the
 * post-resample redistribution all go through peer memory from inside the library's own kernels:
 * no NCCL, no host round trip, one host synchronisation per step.
 *
 * (1) One process, several devices (or several shards on one device): bb200_sharded_amcl -- the shape
 *     of beluga_ros::Amcl (one node, one thread calling update; beluga_ros/src/amcl.cpp:83-126).
 * (2) One process per GPU (torchrun, MPI): every rank creates a bb200_amcl with shard_first_index /
 *     shard_capacity set, exports 256 bytes of CUDA IPC handles (bb200_amcl_export_shard), gathers
 *     the handles of all ranks by whatever transport it has, maps them (bb200_amcl_join_shards) and
 *     from then on calls bb200_amcl_update in lock step with its peers.
 * Policies: every_n resampling and selective resampling (on_effective_size_drop), systematic or
 * multinomial, recovery injection, KLD-adaptive particle counts (min_particles < max_particles: every rank
 * sees the spatial hash of every candidate slot and counts distinct buckets over the same ordered stream, so
 * all ranks stop at the reference's count; the new set is re-split into equal contiguous shards).
 * ------------------------------------------------------------------------------------------- */
typedef struct bb200_sharded_amcl bb200_sharded_amcl;
/* p->max_particles is the GLOBAL particle count (a multiple of n_shards); devices[r] is the CUDA ordinal of
 * shard r (ordinals may repeat: several shards on one device). */
int bb200_sharded_amcl_create(const bb200_amcl_param* p, const bb200_motion_param* motion, int n_shards, const int* devices,
                              bb200_sharded_amcl** out);
void bb200_sharded_amcl_destroy(bb200_sharded_amcl* g);
const char* bb200_sharded_amcl_last_error(const bb200_sharded_amcl* g);
int bb200_sharded_amcl_shards(const bb200_sharded_amcl* g);
/* Shard r as a bb200_amcl (not owned by the caller): timings, per-shard particles, bb200_amcl_filter. */
bb200_amcl* bb200_sharded_amcl_shard(bb200_sharded_amcl* g, int rank);
int bb200_sharded_amcl_set_likelihood_field_map(bb200_sharded_amcl* g, const bb200_likelihood_field_param* p, const bb200_occupancy_grid* grid, int prob);
int bb200_sharded_amcl_set_beam_map(bb200_sharded_amcl* g, const bb200_beam_param* p, const bb200_occupancy_grid* grid);
int bb200_sharded_amcl_initialize(bb200_sharded_amcl* g, const double mean_xytheta[3], const double cov[9]);
int bb200_sharded_amcl_initialize_from_map(bb200_sharded_amcl* g);
void bb200_sharded_amcl_force_update(bb200_sharded_amcl* g);
/* Amcl::update (amcl_core.hpp:165-201) over all shards. */
int bb200_sharded_amcl_update(bb200_sharded_amcl* g, const double control_pose[4], const double* points_xy, uint64_t n_points,
                              bb200_update_result* out);
/* beluga::cluster_based_estimate over all shards: every shard builds its cell records on its device
 * (bb200_filter_particle_histogram with the clusterizer's resolutions), the records are merged on the host in rank
 * order (bb200_cluster_merge_host) and flooded like the single-GPU estimate.  Counts, weights and moments of a cell
 * that spans shards are added shard by shard: exact for unit weights (the state after a resample), otherwise equal to
 * the reference's single sequential sum up to the grouping of the additions. */
int bb200_sharded_amcl_cluster_estimate(bb200_sharded_amcl* g, const bb200_cluster_param* p, bb200_estimate* out, uint32_t* n_cells,
                                        uint32_t* n_clusters);
/* Host only: merges per-shard cell records (each list in its shard's first-occurrence order, shards in rank order) into
 * the global first-occurrence order; shard_first_index (may be NULL) turns local particle indices into global ones.
 * One process per GPU: gather the lists of bb200_filter_particle_histogram with the application's transport, then
 * bb200_cluster_merge_host + bb200_cluster_select_host + bb200_estimate_from_moments. */
int bb200_cluster_merge_host(const bb200_cluster_cell* const* shard_cells, const uint64_t* shard_counts, const uint64_t* shard_first_index, int shards,
                             bb200_cluster_cell* merged, uint64_t capacity, uint64_t* n_merged);
/* particles() in global index order (rank 0's shard first). */
int bb200_sharded_amcl_get_particles(bb200_sharded_amcl* g, double* states, double* weights, uint64_t capacity);
#define BB200_SHARD_HANDLE_BYTES 256
/* One process per GPU: CUDA IPC handles of this shard's two state buffers, mail block and (KLD) hash array (256 bytes) ... */
int bb200_amcl_export_shard(bb200_amcl* a, void* out256);
/* ... and the mapping of all ranks' handles (world x 256 bytes, rank order).  shard_capacity * world must equal
 * max_particles and shard_first_index must be rank * shard_capacity. */
int bb200_amcl_join_shards(bb200_amcl* a, int world, int rank, const void* handles);
/* Unmap the peers' buffers again.  CUDA IPC wants every importer to close before the exporter frees: call this on all
 * ranks, synchronise the ranks (barrier), then destroy. */
int bb200_amcl_leave_shards(bb200_amcl* a);
int bb200_filter_export_shard(bb200_filter* f, void* out256);
int bb200_filter_join_shards(bb200_filter* f, int world, int rank, const void* handles);

/* The two host halves of bb200_amcl_update for callers that drive the filter themselves. */
int bb200_amcl_plan_update(bb200_amcl* a, const double control_pose[4], bb200_step_plan* plan);
void bb200_amcl_commit_update(bb200_amcl* a, int resampled, double random_state_probability);
/* Amcl with any of the three motion models (bb200_amcl_create is the differential-drive shorthand). */
int bb200_amcl_create_with_motion(const bb200_amcl_param* p, const bb200_motion_param* motion, bb200_amcl** out);
/* MotionModel::operator()(control) host part for any model (differential_drive_model.hpp:129-154,
 * omnidirectional_drive_model.hpp:101-129, stationary_model.hpp:52). */
int bb200_motion_sampling_from_control(const bb200_motion_param* p, const double pose[4], const double previous_pose[4], bb200_motion_sampling* out);
/* DifferentialDriveModel::operator()(control) host part -- differential_drive_model.hpp:129-154. */
int bb200_diff_drive_sampling_from_control(const bb200_diff_drive_param* p, const double pose[4], const double previous_pose[4], bb200_diff_drive_sampling* out);

#ifdef __cplusplus
}
#endif
#endif /* BELUGA_B200_H_ */
