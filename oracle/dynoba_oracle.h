/*
 * oracle/dynoba_oracle.h -- CPU ORACLE (TEST INFRASTRUCTURE, NOT PRODUCT).
 *
 * Plain-C fp64 restatement of the reference's nonlinear-least-squares hot path:
 * DynOSAM's factors (in-tree sources, cited per function in dynoba_oracle.c) on
 * top of GTSAM 4.2.0 semantics (un-vendored third-party dependency pinned at
 * tags/4.2.0 by /root/reference/docker/Dockerfile.amd64:104-112; its algorithm is
 * restated from the published source).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
 * reference legs may load this library.  The product (libdynoba.so) never does.
 *
 * PARITY PINNING: every Jacobian known-answer test the reference holds for this
 * path (test_factors.cc:92-196, test_hybrid_motion.cc:71-343, the Schur KAT
 * test_factors.cc:278-450 and the LM recovery test :462-556) is re-run against
 * this oracle in tests/test_oracle_kat.py.  GTSAM's own LM chi^2 traces are NOT
 * pinned by any reference test ("parity unpinned" for the LM trace itself).
 */
#ifndef DYNOBA_ORACLE_H
#define DYNOBA_ORACLE_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

enum {
  ORC_PRIOR6 = 0, ORC_BETWEEN6 = 1, ORC_POSE2POINT3 = 2, ORC_STEREO3 = 3,
  ORC_TERNARY3 = 4, ORC_HYBRID3 = 5, ORC_HYBRID_STEREO3 = 6, ORC_MOTIONPOSE3 = 7,
  ORC_SMOOTH_HYBRID6 = 8, ORC_SMOOTH_POSE6 = 9, ORC_FLOWPROJ2 = 10, ORC_NUM_TYPES = 11
};

/* One homogeneous block of factors.  idx is [n][arity] (row-major), meas is
 * [n][meas_dim], sigma is [n][sigma_dim] or a single [sigma_dim] row when
 * sigma_bcast != 0.  robust_k <= 0 means plain Gaussian noise. */
typedef struct {
  int type, n;
  const int32_t* idx;
  const double* meas;
  const double* sigma;
  int sigma_dim, sigma_bcast;
  double robust_k;
  const int32_t* aux_idx; /* [n] index into aux_pose (L_e table) or NULL */
} orc_block;

typedef struct {
  int n_pose, n_point, n_flow;
  double* pose;   /* [n_pose][12]: R row-major (9) then t (3) */
  double* point;  /* [n_point][3] */
  double* flow;   /* [n_flow][2]  */
  int n_aux;
  const double* aux_pose; /* [n_aux][12] fixed poses (L_e) */
  double calib[6];        /* fx fy s u0 v0 baseline */
  int n_blocks;
  const orc_block* blocks;
  const int32_t* pose_order; /* optional [n_pose] ordering hint (e.g. frame id) */
} orc_problem;

typedef struct {
  double lambda_initial, lambda_factor, lambda_upper, lambda_lower;
  double min_model_fidelity, rel_tol, abs_tol, err_tol;
  int max_iterations;
  int verbose;
} orc_lm_params;

typedef struct {
  int iterations, inner_iterations;
  double error_initial, error_final, lambda_final;
  int bandwidth, reduced_dim;
  double t_linearize, t_schur, t_solve, t_backsub, t_error, t_total;
} orc_lm_stats;

/* static facts about a factor type */
int orc_type_arity(int type);
int orc_type_dim(int type);       /* residual rows */
int orc_type_meas_dim(int type);
int orc_type_jcols(int type);     /* total Jacobian columns (sum of key dims) */

void orc_lm_default_params(orc_lm_params* p);

/* Unwhitened residual r (dim) and Jacobian J [dim][jcols] of factor i of block b. */
void orc_factor_eval(const orc_problem* P, const orc_block* b, int i, double* r, double* J);
/* GTSAM NoiseModelFactor::linearize: whitened (+Huber re-weighted) A [n][dim][jcols]
 * and rhs bvec [n][dim] = -r_w * sqrt(w). */
void orc_linearize_block(const orc_problem* P, const orc_block* b, double* A, double* bvec);
/* per-factor nonlinear error (0.5|r_w|^2 or Huber rho) */
void orc_error_block(const orc_problem* P, const orc_block* b, double* err);
/* graph.error(values) */
double orc_error(const orc_problem* P);

/* Dense normal equations over [poses(6 each, natural order) | points(3) | flows(2)]:
 * H (n*n row-major) = sum A^T A, g = sum A^T b.  Small problems only. */
int orc_dense_dim(const orc_problem* P);
void orc_dense_normal(const orc_problem* P, double* H, double* g);

/* One damped solve (H + lambda I) delta = g through the Schur/band-Cholesky path.
 * delta in the same layout as orc_dense_normal.  Returns 0 ok, 1 not SPD. */
int orc_schur_solve(const orc_problem* P, double lambda, double* delta);
/* Reduced system in the solver ordering, dense (6*n_pose)^2 row-major + rhs; returns dim */
int orc_reduced_dense(const orc_problem* P, double lambda, double* S, double* gS, int32_t* pose_pos);

/* values.retract(delta): modifies P in place */
void orc_retract(orc_problem* P, const double* delta);

/* Levenberg-Marquardt, literal GTSAM 4.2 control flow.  Modifies values in place. */
int orc_lm_optimize(orc_problem* P, const orc_lm_params* prm, orc_lm_stats* st);

/* Lie helpers exported for the KAT tests */
void orc_se3_expmap(const double* xi, double* pose12);
void orc_se3_logmap(const double* pose12, double* xi);
void orc_se3_compose(const double* a, const double* b, double* out);
void orc_se3_inverse(const double* a, double* out);
void orc_se3_retract(const double* pose12, const double* xi, double* out);
void orc_hybrid_project_to_object3(const double* X, const double* E, const double* L,
                                   const double* Z, double* out, double* JX, double* JE, double* JL);

#ifdef __cplusplus
}
#endif
#endif
