/*
 * dynoba.h -- C ABI of libdynoba: B200-native (sm_100a) batch nonlinear-least-squares hot path for DynOSAM.
 *
 * This is the drop-in boundary for the reference's solver call
 *     gtsam::LevenbergMarquardtOptimizer problem(graph, theta, params);
 *     gtsam::Values v = problem.optimize();  problem.getInnerIterations();  problem.iterations();
 * at /root/reference/dynosam/src/backend/RegularBackendModule.cc:405-428 (same shape at
 * dynosam_opt/src/SlidingWindowOptimization.cc:67-77, dynosam/test/internal/backend_runners.hpp:187-193,
 * dynosam/include/dynosam/frontend/vision/MotionSolver-inl.hpp:184-199).  The header-only C++ adapter
 * include/dynoba_gtsam_adapter.hpp flattens a gtsam::NonlinearFactorGraph / Values into these calls.
 *
 * Plain C types only: no torch, no CUDA types (streams are void*).  All functions return a
 * dynoba_status (0 = ok, < 0 = error); no exception crosses the ABI.  A handle is NOT thread-safe;
 * use one handle per optimiser instance (the reference calls from exactly one backend thread,
 * src/pipeline/PipelineManager.cc:175-250).  Host arrays are copied during the call; nothing is retained.
 * There is no CPU fallback: every compute entry point fails with DYNOBA_ERR_CUDA when no sm_100 device works.
 */
#ifndef DYNOBA_H
#define DYNOBA_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

typedef struct dynoba_solver* dynoba_handle;

typedef enum {
  DYNOBA_OK = 0,
  DYNOBA_ERR_BAD_ARG = -1,
  DYNOBA_ERR_STATE = -2,        /* call order violated (e.g. optimize before variables are set) */
  DYNOBA_ERR_CUDA = -3,         /* CUDA runtime / no usable device */
  DYNOBA_ERR_INDETERMINATE = -4,/* reduced system not SPD even at lambda upper bound
                                   (adapter maps this to gtsam::IndeterminantLinearSystemException) */
  DYNOBA_ERR_UNSUPPORTED = -5,  /* topology outside what the kernels handle */
  DYNOBA_ERR_COMM = -6          /* user all-reduce callback failed */
} dynoba_status;

/* Variable kinds (reference value types: gtsam::Pose3, gtsam::Point3, gtsam::Point2 optical-flow variable
 * of Pose3FlowProjectionFactor, dynosam/include/dynosam/factors/Pose3FlowProjectionFactor.h:45-71). */
typedef enum { DYNOBA_POSE6 = 0, DYNOBA_POINT3 = 1, DYNOBA_FLOW2 = 2 } dynoba_var_kind;

/* Factor types.  idx columns are the factor's keys in the reference's constructor order.
 *  type               reference factor (file:line)                                           keys (idx columns)            meas            */
typedef enum {
  DYNOBA_PRIOR6 = 0,        /* gtsam::PriorFactor<Pose3>   (Formulation-impl.hpp:523-533)              pose                 prior pose (12)   */
  DYNOBA_BETWEEN6 = 1,      /* gtsam::BetweenFactor<Pose3> (dynosam_opt/src/FactorGraphTools.cc:53-63) pose1, pose2         measured (12)     */
  DYNOBA_POSE2POINT3 = 2,   /* gtsam::PoseToPointFactor    (Formulation-impl.hpp:169-172)              pose, point          z (3)             */
  DYNOBA_STEREO3 = 3,       /* gtsam::GenericStereoFactor  (Formulation-impl.hpp:292-293)              pose, point          uL,uR,v (3)       */
  DYNOBA_TERNARY3 = 4,      /* LandmarkMotionTernaryFactor (src/factors/LandmarkMotionTernaryFactor.cc:41-72) prev pt, cur pt, motion   -      */
  DYNOBA_HYBRID3 = 5,       /* HybridMotionFactor          (src/factors/HybridFormulationFactors.cc:137-187) X_k, e_H_k, m_L   z (3) + aux L_e */
  DYNOBA_HYBRID_STEREO3 = 6,/* StereoHybridMotionFactor    (HybridFormulationFactors.cc:213-261)       X_k, e_H_k, m_L      uL,uR,v + aux L_e */
  DYNOBA_MOTIONPOSE3 = 7,   /* LandmarkMotionPoseFactor    (src/factors/LandmarkMotionPoseFactor.cc:42-105) prev pt, cur pt, prev L, cur L  - */
  DYNOBA_SMOOTH_HYBRID6 = 8,/* HybridSmoothingFactor       (HybridFormulationFactors.cc:274-322)       H_k-2, H_k-1, H_k    aux L_e           */
  DYNOBA_SMOOTH_POSE6 = 9,  /* LandmarkPoseSmoothingFactor (src/factors/LandmarkPoseSmoothingFactor.cc:37-95) L_k-2, L_k-1, L_k   -            */
  DYNOBA_FLOWPROJ2 = 10,    /* Pose3FlowProjectionFactor   (factors/Pose3FlowProjectionFactor.h:73-133) flow, pose          kp(2),depth,X_prev(12) */
  DYNOBA_NUM_FACTOR_TYPES = 11
} dynoba_factor_type;

/* gtsam::LevenbergMarquardtParams (defaults = GTSAM 4.2.0, see dynoba_lm_default_params) */
typedef struct {
  double lambda_initial, lambda_factor, lambda_upper_bound, lambda_lower_bound;
  double min_model_fidelity;
  double relative_error_tol, absolute_error_tol, error_tol;
  int32_t max_iterations;
  int32_t verbosity;            /* 0 silent, 1 per-iteration line on stderr */
} dynoba_lm_params;

typedef struct {
  int32_t iterations;           /* LevenbergMarquardtOptimizer::iterations()        */
  int32_t inner_iterations;     /* LevenbergMarquardtOptimizer::getInnerIterations() */
  double error_initial, error_final, lambda_final;
  int32_t reduced_dim, bandwidth;
  int64_t kernel_launches;      /* libdynoba kernels launched by this optimize() call */
  double ms_linearize, ms_schur, ms_factor, ms_backsub, ms_error, ms_total; /* CUDA-event timings */
} dynoba_lm_stats;

/* Sum-all-reduce of n doubles at device pointer `dev` on CUDA stream `stream`, in place.  Supplied by the
 * multi-GPU host (bench.py passes torch.distributed.all_reduce over NCCL).  Return 0 on success. */
typedef int (*dynoba_allreduce_fn)(void* ctx, double* dev, size_t n, void* stream);
/* Sum-reduce of n doubles at device pointer `dev` to rank `root` (in place there; the other ranks' buffers are left
 * unspecified), on CUDA stream `stream`.  bench.py passes torch.distributed.reduce over NCCL.  Return 0 on success. */
typedef int (*dynoba_reduce_fn)(void* ctx, double* dev, size_t n, int root, void* stream);

/* ---- lifecycle */
int dynoba_version(void);
const char* dynoba_status_string(int status);
const char* dynoba_last_error(dynoba_handle h);
int dynoba_create(int device, dynoba_handle* out);
int dynoba_destroy(dynoba_handle h);
void dynoba_lm_default_params(dynoba_lm_params* p);

/* ---- problem ingest (gtsam::Values / NonlinearFactorGraph equivalents) */
/* kind POSE6: data[n][12] = R row-major (9) | t (3);  POINT3: data[n][3];  FLOW2: data[n][2].
 * keys (optional, may be NULL): opaque gtsam::Key values, round-tripped by dynoba_get_keys. */
int dynoba_set_variables(dynoba_handle h, int kind, int64_t n, const uint64_t* keys, const double* data);
/* Fixed poses referenced through aux_idx (HybridMotionFactor's L_e): data[n][12]. */
int dynoba_set_aux_poses(dynoba_handle h, int64_t n, const double* data);
/* Cal3_S2Stereo: fx, fy, skew, u0, v0, baseline (also Cal3_S2 for FLOWPROJ2, baseline ignored). */
int dynoba_set_calibration(dynoba_handle h, const double calib[6]);
/* idx[n][arity] (int32 indices into the variable arrays of the slot's kind), meas[n][meas_dim] or NULL,
 * sigma: sigma_count == 1 -> one row [sigma_dim] shared by all n factors, else [n][sigma_dim];
 * sigma_dim is 1 (Isotropic) or the residual dimension (Diagonal); robust_k > 0 wraps the noise model in
 * noiseModel::Robust(mEstimator::Huber(k)) (BackendDefinitions.cc:170-193), <= 0 = Gaussian;
 * aux_idx[n] or NULL. */
int dynoba_add_factors(dynoba_handle h, int type, int64_t n, const int32_t* idx, const double* meas,
                       const double* sigma, int sigma_dim, int64_t sigma_count, double robust_k,
                       const int32_t* aux_idx);
/* gtsam::LinearContainerFactor holding a HessianFactor over n pose-like variables -- the form in which the reference's
 * sliding-window optimiser hands the information of the marginalised variables to the next window
 * (dynosam_opt/src/SlidingWindowOptimization.cc:67-121,165-190).  lin_poses[n][12] = linearisation point, G[6n][6n]
 * (row-major, symmetric) and g[6n] = information matrix / vector in the tangent space at the linearisation point
 * ([omega; v] per pose), f = constant:  error(x) = 1/2 d^T G d - g^T d + 1/2 f  with d_i = Logmap(lin_i^-1 x_i);
 * linearising at x gives HessianFactor(G, g - G d, ...) as GTSAM does.  A variable may appear once per prior.  With
 * several ranks, give a prior to exactly one of them. */
int dynoba_add_linear_prior(dynoba_handle h, int32_t n, const int32_t* pose_idx, const double* lin_poses,
                            const double* G, const double* g, double f);
/* Marginal information (G[6k][6k], g[6k]) of the k pose-like variables keep_pose_idx at the current values, every other
 * variable -- all landmarks and the older pose-like variables -- eliminated: what SlidingWindowOptimization::
 * CalculateMarginalFactors computes with eliminatePartialMultifrontal (:165-190), ready to be passed to
 * dynoba_add_linear_prior of the next window together with the current values of the kept variables as linearisation
 * point.  The kept variables must be the last ones of the elimination order (the most recent frames) and span no more
 * than the band of the reduced system (else DYNOBA_ERR_UNSUPPORTED).  Difference from the reference: it retains recent
 * LANDMARKS as well; here every landmark is eliminated into the prior.  Single GPU. */
int dynoba_marginal(dynoba_handle h, int32_t n_keep, const int32_t* keep_pose_idx, double* G, double* g);
/* Optional elimination-order hint for pose-like variables (e.g. the frame id encoded in the key):
 * variables are ordered by (rank, index).  Without it the given order is used. */
int dynoba_set_pose_order(dynoba_handle h, int64_t n, const int32_t* rank);
/* Multi-GPU: rank/world of the landmark shard held by this handle and the all-reduce used for the reduced
 * system and the scalar sums.  min_bandwidth forces a common band layout on all ranks (0 = local). */
int dynoba_set_shard(dynoba_handle h, int rank, int world, dynoba_allreduce_fn fn, void* ctx, int min_bandwidth);
/* Multi-GPU, optional: with a reduce callback the reduced solve itself is distributed -- the time axis of the banded
 * reduced system is cut into cells (one per rank by default), every rank factors its own cells and only receives their
 * tiles (a reduce per cell instead of an all-reduce of the whole band); the small boundary-separator system and the
 * pose update are all-reduced.  Without it the reduced system is all-reduced and solved on every rank. */
int dynoba_set_reduce(dynoba_handle h, dynoba_reduce_fn fn, void* ctx);
/* Number of cells the reduced solve is cut into (nested dissection in time; every cell is eliminated by two concurrent
 * chains, so 2*ncells chains run at once).  0 = automatic (one per rank; on one GPU by system length), -1 = one plain
 * band factorisation.  The request is clamped to what the system's length allows.  GTSAM equivalent: the elimination
 * ordering of LevenbergMarquardtParams (RegularBackendModule.cc:405-419 leaves it at COLAMD). */
int dynoba_set_partition(dynoba_handle h, int ncells);
/* The cells' extent on the pose axis for a reduced system of n_poses pose-like variables and the given scalar
 * half-bandwidth, cut for `world` ranks: first_position[r] = first pose position (in elimination order) of rank r's cells,
 * first_position[world] = n_poses.  A host that shards landmarks by the position of their first pose along these bounds
 * keeps every rank's contribution inside its own cells (plus a halo one co-visibility window wide).  Needs no handle
 * and no device. */
int dynoba_plan_partition(int32_t n_poses, int32_t bandwidth, int32_t world, int32_t* first_position);
/* Performance parameters of the reduced solve (results do not depend on them): "outer_weight" = relative length of the
 * two end chains, which carry no spike (default 4.5 on one GPU, 1 otherwise); "band_ctas_per_chain" = worker CTAs that
 * serve the band tiles of one chain (the others stream the spike updates; process-wide). */
int dynoba_set_tuning(dynoba_handle h, const char* name, double value);
/* Builds the device layout (sorting, CSR, band structure) and uploads.  Called implicitly by the
 * compute entry points; exposed so that uploads can be timed separately. */
int dynoba_finalize(dynoba_handle h);

/* ---- compute */
int dynoba_error(dynoba_handle h, double* out);                       /* graph.error(values) */
int dynoba_optimize(dynoba_handle h, const dynoba_lm_params* p, dynoba_lm_stats* stats);
int dynoba_get_variables(dynoba_handle h, int kind, int64_t n, double* out);
int dynoba_get_keys(dynoba_handle h, int kind, int64_t n, uint64_t* out);
int dynoba_num_variables(dynoba_handle h, int kind, int64_t* out);
/* Measured fp64 FMA throughput of the device in TFLOP/s (a few milliseconds of a register-only FMA kernel): the
 * roofline denominator bench.py prints next to the reduced solve (tcgen05 has no fp64; DMMA runs at the FMA rate). */
int dynoba_fp64_rate(dynoba_handle h, double* tflops);
int dynoba_problem_info(dynoba_handle h, int32_t* reduced_dim, int32_t* bandwidth, int64_t* jacobian_bytes);

/* ---- stepwise / parity hooks (what gtsam exposes as graph.linearize(values)) */
/* Materialising linearize of every factor at the current values (the roofline kernel).  Returns the
 * CUDA-event time of the linearize kernels alone in *ms (may be NULL). */
int dynoba_linearize(dynoba_handle h, float* ms);
/* Same, for the factor block of the block_index-th dynoba_add_factors call alone: one launch of the Jacobian-build
 * kernel of that factor type, CUDA-event time in *ms, its algorithmic bytes in *bytes. */
int dynoba_linearize_block(dynoba_handle h, int block_index, float* ms, int64_t* bytes);
/* Whitened Jacobian A[n][dim][jcols] and rhs b[n][dim] of the block_index-th dynoba_add_factors call,
 * in the caller's factor order (valid after dynoba_linearize / dynoba_optimize). */
int dynoba_get_linearization(dynoba_handle h, int block_index, double* A, double* b);
/* Per-factor nonlinear error of one block (0.5|r_w|^2 or Huber rho). */
int dynoba_get_factor_errors(dynoba_handle h, int block_index, double* err);
/* One damped solve (J^T J + lambda I) delta = J^T b through Schur + band Cholesky at the current
 * linearization; delta laid out [poses(6) in caller order | points(3) | flows(2)]. */
int dynoba_solve(dynoba_handle h, double lambda, double* delta);
/* Dense copy of the reduced camera/motion system S (dim x dim, row-major, caller pose order) and g_S. */
int dynoba_get_reduced_system(dynoba_handle h, double lambda, double* S, double* g);
/* values.retract(delta), same layout as dynoba_solve */
int dynoba_retract(dynoba_handle h, const double* delta);

/* ---- graph construction straight into SoA blocks: what Formulation<MAP>::updateStaticObservations /
 * updateDynamicObservations + the hybrid formulation's callbacks do with the Map
 * (dynosam/include/dynosam/backend/Formulation-impl.hpp:552-897, src/backend/rgbd/HybridEstimator.cc:573-830), without
 * allocating a factor object per observation: feed frames and observations, get the variable arrays and one homogeneous
 * block per factor type (read them back, or hand them to a solver handle with dynoba_builder_emit).  Host code only. */
typedef struct dynoba_builder* dynoba_builder_handle;
/* HYBRID: HybridEstimator.cc:573-830.  WCME: WorldMotionEstimator.cc:151-351 (point per (tracklet, frame), ternary motion factors,
 * motion variables initialised with the front-end translation and identity rotation).  WCPE: WorldPoseEstimator.cc:89-315 (object
 * POSE variables L_k, initialised motion * L_k-1 or centroid; dynoba_builder_set_keyframe_pose(object, frame, L) overrides the
 * initial value of L at that frame). */
enum { DYNOBA_FORMULATION_HYBRID = 0, DYNOBA_FORMULATION_WCME = 1, DYNOBA_FORMULATION_WCPE = 2 };
typedef struct {
  int32_t min_static_obs, min_dynamic_obs;   /* min_static_observations (2), min_dynamic_observations (3) */
  int32_t keyframe_gap;                      /* new object key-frame once unseen for more than this many frames (2) */
  double sigma_static, sigma_dynamic;        /* isotropic point noise (0.2) */
  double huber_k;                            /* 1e-4; <= 0: Gaussian */
  double odometry_sigma[6], smoothing_sigma[6];
  double prior_sigma;                        /* 1e-6 */
  int32_t formulation;                       /* DYNOBA_FORMULATION_* (hybrid) */
  double sigma_motion;                       /* motion_ternary_factor_noise_sigma (0.01): TERNARY3 / MOTIONPOSE3 of the world-centric formulations */
  int32_t backtrack;                         /* UpdateObservationParams::do_backtrack.  1 (default): every observation of a tracklet that reaches
                                                the minimum count enters the graph (the batch graphs of SURVEY 8d; ParallelHybridBackendModule.cc:427).
                                                0: RegularBackendModule.cc:139,197 -- a tracklet enters at the frame its count reaches the minimum
                                                with only what that update adds (static: that observation on; dynamic: the last pair on) */
} dynoba_builder_params;
void dynoba_builder_default_params(dynoba_builder_params* p);
int dynoba_builder_create(const dynoba_builder_params* p, dynoba_builder_handle* out);
int dynoba_builder_destroy(dynoba_builder_handle b);
const char* dynoba_builder_last_error(dynoba_builder_handle b);
/* camera pose estimate X[12] of a frame and (NULL for the first frame) the odometry measurement from the previous frame */
int dynoba_builder_add_frame(dynoba_builder_handle b, int32_t frame, const double* X, const double* odom_from_prev);
/* point measurements in the camera frame, z[n][3]; object ids start at 1 */
int dynoba_builder_add_static(dynoba_builder_handle b, int32_t frame, int64_t n, const int64_t* tracklet, const double* z);
int dynoba_builder_add_dynamic(dynoba_builder_handle b, int32_t frame, int64_t n, const int64_t* tracklet, const int32_t* object, const double* z);
/* initial value of the cumulative motion e_H_k (front-end estimate; identity at key-frames) / fixed key-frame pose L_e
 * (default: centroid of the key-frame's points, identity rotation) */
int dynoba_builder_set_motion_init(dynoba_builder_handle b, int32_t object, int32_t frame, const double* H);
int dynoba_builder_set_keyframe_pose(dynoba_builder_handle b, int32_t object, int32_t keyframe, const double* L_e);
int dynoba_builder_finalize(dynoba_builder_handle b);
int dynoba_builder_counts(dynoba_builder_handle b, int64_t* n_pose, int64_t* n_point, int64_t* n_aux, int32_t* n_blocks);
int dynoba_builder_get_variables(dynoba_builder_handle b, double* pose, double* point, double* aux, int32_t* pose_order,
                                 uint64_t* pose_keys, uint64_t* point_keys);
int dynoba_builder_block_info(dynoba_builder_handle b, int32_t block, int32_t* type, int64_t* n, int32_t* sigma_dim,
                              int64_t* sigma_count, double* robust_k, int32_t* has_aux);
int dynoba_builder_get_block(dynoba_builder_handle b, int32_t block, int32_t* idx, double* meas, double* sigma, int32_t* aux_idx);
int dynoba_builder_emit(dynoba_builder_handle b, dynoba_handle h);

/* ---- batched star problems --------------------------------------------------------------------------
 * The front end's per-object refinements, for ALL objects of a frame (or of many frames) in ONE launch: one CTA per problem
 * runs the whole Levenberg-Marquardt (GTSAM tryLambda semantics, as dynoba_optimize) and the outlier rounds on the device.
 * All pointers are HOST memory; poses are 12 doubles (row-major R, then t); calib5 = fx fy s u0 v0 (gtsam::Cal3_S2).
 * Problem p owns the features / tracklets [offsets[p], offsets[p+1]) of the concatenated per-feature arrays; offsets[0] = 0.
 *
 * (1) dynoba_flow_pose_batch replaces the loop over OpticalFlowAndPoseOptimizer::optimize
 * (dynosam/include/dynosam/frontend/vision/MotionSolver-inl.hpp:88-278): one Pose3 unknown and one Point2 flow unknown per
 * feature; every feature contributes
 *   Pose3FlowProjectionFactor(flow_i, pose; kp_prev_i, depth_i, pose_prev_p, K_p)   noise Robust(Huber(huber_k), Isotropic(flow_sigma))
 *   PriorFactor<Point2>(flow_i, flow_i^0)                                            noise Isotropic(flow_prior_sigma)
 * `flow` holds the measured flows: initial value and prior mean.  After the first LM, up to outlier_rounds times: the flow
 * factors whose Gaussian error exceeds outlier_threshold leave the graph (their prior stays), the pose is reset to pose_init,
 * LM runs again (:201-247).  Outputs: refined pose per problem, refined flow per feature, inlier_out[i] = 1 if the feature's
 * factor is still in the graph, error of the full graph at the initial values, error of the final graph at the result, LM
 * iterations / inner (lambda) iterations summed over the rounds, number of outlier rounds run. */
typedef struct {
  double flow_sigma, flow_prior_sigma, huber_k;   /* OpticalFlowAndPoseOptimizer::Params: 10, 3.33, 0.001; huber_k <= 0: Gaussian */
  int32_t outlier_rounds;                         /* 4 (params.outlier_reject); 0 = no outlier rejection */
  double outlier_threshold;                       /* <= 0: 0.5 * chi2inv(0.99, 2) (factor_graph_tools::determineFactorOutliers) */
  dynoba_lm_params lm;                            /* GTSAM defaults with max_iterations 10 */
} dynoba_flow_pose_params;
void dynoba_flow_pose_default_params(dynoba_flow_pose_params* p);
int dynoba_flow_pose_batch(int device, int32_t n_problems, const int32_t* offsets, const double* pose_init,
                           const double* pose_prev, const double* calib5, const double* kp_prev, const double* depth,
                           const double* flow, const dynoba_flow_pose_params* prm, double* pose_out, double* flow_out,
                           uint8_t* inlier_out, double* err_before, double* err_after, int32_t* iterations,
                           int32_t* inner_iterations, int32_t* rounds);
/* (2) dynoba_motion_refine_batch replaces the loop over MotionOnlyRefinementOptimizer::optimize (:291-470, ProjectionError
 * solver): unknowns are the two camera poses X_k-1, X_k (PriorFactor, Isotropic pose_prior_sigma), the object motion H and two
 * world points per tracklet (points_init[i] = m_k-1 | m_k, 6 doubles); every tracklet contributes
 *   GenericProjectionFactor(X_k-1, m_k-1; kp_prev_i, K)  and  (X_k, m_k; kp_cur_i, K)   Robust(Huber(huber_k), Isotropic(projection_sigma))
 *   LandmarkMotionTernaryFactor(m_k-1, m_k, H)                                           Robust(Huber(huber_k), Isotropic(landmark_motion_sigma))
 * One LM per problem (max_iterations 5).  motion_factor_error[i] is the Gaussian error of tracklet i's motion factor at the
 * result -- the quantity determineFactorOutliers<LandmarkMotionTernaryFactor> thresholds at 0.5 * chi2inv(0.99, 3); the
 * reference's re-optimisation after removing those factors re-inserts an existing key into gtsam::Values (:441) and is not
 * reproduced: call again with the inliers.  poses_out (24 per problem), points_out, motion_factor_error may be NULL. */
typedef struct {
  double landmark_motion_sigma, projection_sigma, huber_k;   /* MotionOnlyRefinementOptimizer::Params: 0.001, 2.0, 0.0001 */
  double pose_prior_sigma;                                   /* 0.00001 */
  dynoba_lm_params lm;                                       /* GTSAM defaults with max_iterations 5 */
} dynoba_motion_refine_params;
void dynoba_motion_refine_default_params(dynoba_motion_refine_params* p);
/* Both batch entry points keep one stream and one grow-only device workspace per device between calls (they run every frame:
 * no allocation on that path; calls on one device serialise).  dynoba_batch_release frees them. */
int dynoba_batch_release(int device);
int dynoba_motion_refine_batch(int device, int32_t n_problems, const int32_t* offsets, const double* pose_prev,
                               const double* pose_cur, const double* motion_init, const double* calib5,
                               const double* kp_prev, const double* kp_cur, const double* points_init,
                               const dynoba_motion_refine_params* prm, double* motion_out, double* poses_out,
                               double* points_out, double* motion_factor_error, double* err_before, double* err_after,
                               int32_t* iterations, int32_t* inner_iterations);

#ifdef __cplusplus
}
#endif
#endif /* DYNOBA_H */
