// kernels_star.cu -- SURVEY.md 8f-2: many small "star" Levenberg-Marquardt problems in ONE launch.
//
// The front end refines, per frame and per object, a camera pose (or object motion) jointly with the optical flow of its
// features: OpticalFlowAndPoseOptimizer::optimize (dynosam/include/dynosam/frontend/vision/MotionSolver-inl.hpp:88-260) builds,
// for every tracked feature i, a Pose3FlowProjectionFactor(flow_i, pose; kp_i, depth_i, X_prev, K) with Robust(Huber) noise and
// a PriorFactor<Point2>(flow_i, measured flow) and runs gtsam::LevenbergMarquardtOptimizer with maxIterations 10 -- one tiny
// LM (6 + 2N unknowns, N <= a few hundred) per object per frame.  Here every problem gets one CTA that runs the WHOLE LM
// loop on the device (no host round trip per iteration): linearise, eliminate the flow variables (each has a scalar-diagonal
// 2x2 block: J_flow = I), factor the 6x6 pose system, back-substitute, retract, evaluate, and GTSAM's tryLambda control
// (LevenbergMarquardtOptimizer.cpp) -- literally the control loop of api.cu::dynoba_optimize.
#include <cfloat>
#include <algorithm>
#include "../../include/dynoba.h"
#include "internal.cuh"
#include "se3.cuh"

namespace dynoba {

struct StarBatch {
  int nprob; const int* off;                    // [nprob + 1] factor ranges
  const double* pose0; const double* pose_prev; const double* calib;    // [nprob][12], [nprob][12], [nprob][5] (fx fy s u0 v0)
  const double* kp; const double* depth; const double* flow0;           // [total][2], [total], [total][2] (prior mean = initial value)
  double isig_flow, isig_prior, huber_k;
  dynoba_lm_params prm;
  double* flow_cur; double* flow_cand;          // [total][2] work
  double* pose_out; double* flow_out; double* err_before; double* err_after; int* iterations; int* inner;
};

constexpr int STAR_THREADS = 256;

// deterministic block sums of NV values per thread (fixed tree order)
template <int NV>
__device__ __forceinline__ void block_sum(double (&v)[NV], double* red /* [NV][STAR_THREADS/32] */, double* out /* [NV] */) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < NV; k++) {
    double x = v[k];
    for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
    if (lane == 0) red[k*(STAR_THREADS/32) + warp] = x;
  }
  __syncthreads();
  if (threadIdx.x < NV) { double s = 0; for (int w = 0; w < STAR_THREADS/32; w++) s += red[threadIdx.x*(STAR_THREADS/32) + w]; out[threadIdx.x] = s; }
  __syncthreads();
}

// one feature: whitened, Huber-weighted flow-projection factor + Gaussian flow prior at (X, flow)
struct StarFactor { double A[2][6]; double a[2]; double b[2]; double bp[2]; double err; };
__device__ __forceinline__ void star_factor(const StarBatch& S, int i, const Pose& X, const Pose& Xp, const double* K6, const double* flow, StarFactor& F, bool want_j) {
  FVars v; v.pose[0] = X; v.pt[0][0] = flow[0]; v.pt[0][1] = flow[1]; v.pt[0][2] = 0.0;
  double z[15] = { S.kp[2*i], S.kp[2*i + 1], S.depth[i] };
  for (int k = 0; k < 9; k++) z[3 + k] = Xp.R[k];
  for (int k = 0; k < 3; k++) z[12 + k] = Xp.t[k];
  double r[2], J[16]; Pose none{};
  if (want_j) factor_eval<F_FLOWPROJ2, true>(v, z, none, K6, r, J); else factor_eval<F_FLOWPROJ2, false>(v, z, none, K6, r, J);
  double e = 0.0;
  const double isig = S.isig_flow;
  const double sw = whiten_weight<2>(r, &isig, 1, S.huber_k, &e);
  const double rp0 = (flow[0] - S.flow0[2*i])*S.isig_prior, rp1 = (flow[1] - S.flow0[2*i + 1])*S.isig_prior;
  F.err = e + 0.5*(rp0*rp0 + rp1*rp1);
  F.b[0] = -sw*r[0]; F.b[1] = -sw*r[1]; F.bp[0] = -rp0; F.bp[1] = -rp1;
  if (want_j) {
    const double sc = sw*isig;
    F.a[0] = sc*J[0*8 + 0]; F.a[1] = sc*J[1*8 + 1];
    for (int c = 0; c < 6; c++) { F.A[0][c] = sc*J[0*8 + 2 + c]; F.A[1][c] = sc*J[1*8 + 2 + c]; }
  }
}

__global__ void __launch_bounds__(STAR_THREADS) star_lm_kernel(StarBatch S) {
  __shared__ double red[28*(STAR_THREADS/32)];
  __shared__ double sums[28];
  __shared__ double sh_dp[6];
  __shared__ Pose sh_cur, sh_cand;
  __shared__ int sh_ok;
  const int pb = blockIdx.x;
  const int f0 = S.off[pb], f1 = S.off[pb + 1];
  const dynoba_lm_params P = S.prm;
  Pose Xp; double K6[6];
  for (int k = 0; k < 9; k++) Xp.R[k] = S.pose_prev[12*pb + k];
  for (int k = 0; k < 3; k++) Xp.t[k] = S.pose_prev[12*pb + 9 + k];
  for (int k = 0; k < 5; k++) K6[k] = S.calib[5*pb + k];
  K6[5] = 0.0;
  if (threadIdx.x == 0) { for (int k = 0; k < 9; k++) sh_cur.R[k] = S.pose0[12*pb + k]; for (int k = 0; k < 3; k++) sh_cur.t[k] = S.pose0[12*pb + 9 + k]; }
  for (int i = f0 + threadIdx.x; i < f1; i += STAR_THREADS) { S.flow_cur[2*i] = S.flow0[2*i]; S.flow_cur[2*i + 1] = S.flow0[2*i + 1]; }
  __syncthreads();
  auto total_error = [&](const Pose& X, const double* flows) {
    double e[1] = { 0.0 };
    for (int i = f0 + threadIdx.x; i < f1; i += STAR_THREADS) { StarFactor F; star_factor(S, i, X, Xp, K6, flows + 2*i, F, false); e[0] += F.err; }
    block_sum<1>(e, red, sums);
    return sums[0];
  };
  double err = total_error(sh_cur, S.flow_cur);
  const double err_initial = err;
  double lambda = P.lambda_initial;
  int iterations = 0, inner = 0;
  if (!(err <= P.error_tol) && P.max_iterations > 0) {
    double newError = err, currentError;
    do {
      currentError = newError;
      for (;;) {      // tryLambda
        // ---- reduced pose system at (cur, lambda): S6 = sum A^T A + lambda I - sum_k W_k W_k^T / v_k,  g6 likewise; 1/2 sum b^2
        double acc[28];
#pragma unroll
        for (int k = 0; k < 28; k++) acc[k] = 0.0;
        const Pose X = sh_cur;
        for (int i = f0 + threadIdx.x; i < f1; i += STAR_THREADS) {
          StarFactor F; star_factor(S, i, X, Xp, K6, S.flow_cur + 2*i, F, true);
          const double ip2 = S.isig_prior*S.isig_prior;
          int e = 0;
          for (int r = 0; r < 6; r++) for (int c = 0; c <= r; c++, e++) {
            double u = F.A[0][r]*F.A[0][c] + F.A[1][r]*F.A[1][c];
            for (int k = 0; k < 2; k++) { const double v = F.a[k]*F.a[k] + ip2 + lambda; u -= (F.A[k][r]*F.a[k])*(F.A[k][c]*F.a[k])/v; }
            acc[e] += u;
          }
          for (int r = 0; r < 6; r++) {
            double g = F.A[0][r]*F.b[0] + F.A[1][r]*F.b[1];
            for (int k = 0; k < 2; k++) { const double v = F.a[k]*F.a[k] + ip2 + lambda; const double gl = F.a[k]*F.b[k] + S.isig_prior*F.bp[k]; g -= (F.A[k][r]*F.a[k])*gl/v; }
            acc[21 + r] += g;
          }
          acc[27] += 0.5*(F.b[0]*F.b[0] + F.b[1]*F.b[1] + F.bp[0]*F.bp[0] + F.bp[1]*F.bp[1]);
        }
        block_sum<28>(acc, red, sums);
        if (threadIdx.x == 0) {
          double L[6][6]; int e = 0; bool ok = true;
          for (int r = 0; r < 6; r++) for (int c = 0; c <= r; c++, e++) L[r][c] = sums[e] + (r == c ? lambda : 0.0);
          for (int j = 0; j < 6 && ok; j++) {
            double d = L[j][j]; for (int k = 0; k < j; k++) d -= L[j][k]*L[j][k];
            if (!(d > 0.0)) { ok = false; break; }
            L[j][j] = sqrt(d);
            for (int i = j + 1; i < 6; i++) { double s = L[i][j]; for (int k = 0; k < j; k++) s -= L[i][k]*L[j][k]; L[i][j] = s/L[j][j]; }
          }
          if (ok) {
            double y[6];
            for (int i = 0; i < 6; i++) { double s = sums[21 + i]; for (int k = 0; k < i; k++) s -= L[i][k]*y[k]; y[i] = s/L[i][i]; }
            for (int i = 5; i >= 0; i--) { double s = y[i]; for (int k = i + 1; k < 6; k++) s -= L[k][i]*sh_dp[k]; sh_dp[i] = s/L[i][i]; }
            Pose cand; se3_retract(sh_cur, sh_dp, cand); sh_cand = cand;
          }
          sh_ok = ok ? 1 : 0;
        }
        __syncthreads();
        const bool solved0 = sh_ok != 0;
        const double oldLin = sums[27];
        double newLin = 0.0, nerr = INFINITY;
        if (solved0) {
          // ---- back-substitute the flows, linear model at delta, candidate values
          double m[1] = { 0.0 };
          const double ip2 = S.isig_prior*S.isig_prior;
          for (int i = f0 + threadIdx.x; i < f1; i += STAR_THREADS) {
            StarFactor F; star_factor(S, i, X, Xp, K6, S.flow_cur + 2*i, F, true);
            double dl[2], Ad[2];
            for (int k = 0; k < 2; k++) {
              Ad[k] = 0.0; for (int c = 0; c < 6; c++) Ad[k] += F.A[k][c]*sh_dp[c];
              const double v = F.a[k]*F.a[k] + ip2 + lambda, gl = F.a[k]*F.b[k] + S.isig_prior*F.bp[k];
              dl[k] = (gl - F.a[k]*Ad[k])/v;                 // W_k^T dp = a_k (A_k . dp)
            }
            for (int k = 0; k < 2; k++) {
              const double rf = Ad[k] + F.a[k]*dl[k] - F.b[k], rpk = S.isig_prior*dl[k] - F.bp[k];
              m[0] += 0.5*(rf*rf + rpk*rpk);
              S.flow_cand[2*i + k] = S.flow_cur[2*i + k] + dl[k];
            }
          }
          block_sum<1>(m, red, sums);
          newLin = sums[0];
          __syncthreads();
          nerr = total_error(sh_cand, S.flow_cand);
        }
        // ---- LevenbergMarquardtOptimizer::tryLambda (all threads take the same decisions from the same block sums)
        const double lin = oldLin - newLin;
        const bool solved = solved0 && isfinite(lin);
        bool success = false, stop = false;
        if (solved && lin >= 0) {
          const double cost = err - nerr;
          if (lin > DBL_EPSILON*oldLin) { const double fid = cost/lin; success = fid > P.min_model_fidelity; }
          if (fabs(cost) < P.relative_error_tol*err) stop = true;
        }
        __syncthreads();
        if (success) {
          if (threadIdx.x == 0) sh_cur = sh_cand;
          for (int i = f0 + threadIdx.x; i < f1; i += STAR_THREADS) { S.flow_cur[2*i] = S.flow_cand[2*i]; S.flow_cur[2*i + 1] = S.flow_cand[2*i + 1]; }
          __syncthreads();
          lambda = fmax(P.lambda_lower_bound, lambda/P.lambda_factor); err = nerr; iterations++; inner++;
          break;
        } else if (!stop) { lambda *= P.lambda_factor; inner++; if (lambda >= P.lambda_upper_bound) break; }
        else break;
      }
      newError = err;
    } while (iterations < P.max_iterations &&
             !((newError <= P.error_tol) ||
               ((P.relative_error_tol != 0.0) && ((currentError - newError)/currentError <= P.relative_error_tol)) ||
               ((currentError - newError) <= P.absolute_error_tol)) &&
             isfinite(currentError));
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 0; k < 9; k++) S.pose_out[12*pb + k] = sh_cur.R[k];
    for (int k = 0; k < 3; k++) S.pose_out[12*pb + 9 + k] = sh_cur.t[k];
    S.err_before[pb] = err_initial; S.err_after[pb] = err; S.iterations[pb] = iterations; S.inner[pb] = inner;
  }
  for (int i = f0 + threadIdx.x; i < f1; i += STAR_THREADS) { S.flow_out[2*i] = S.flow_cur[2*i]; S.flow_out[2*i + 1] = S.flow_cur[2*i + 1]; }
}

}  // namespace dynoba

using namespace dynoba;

extern "C" int dynoba_flow_pose_batch(int device, int32_t n_problems, const int32_t* offsets, const double* pose_init, const double* pose_prev,
                                      const double* calib5, const double* kp_prev, const double* depth, const double* flow, double flow_sigma,
                                      double flow_prior_sigma, double huber_k, const dynoba_lm_params* prm, double* pose_out, double* flow_out,
                                      double* err_before, double* err_after, int32_t* iterations, int32_t* inner_iterations) {
  if (n_problems == 0) return DYNOBA_OK;
  if (n_problems < 0 || !offsets || !pose_init || !pose_prev || !calib5 || !pose_out || !err_before || !err_after || !iterations || !inner_iterations ||
      !(flow_sigma > 0) || !(flow_prior_sigma > 0)) return DYNOBA_ERR_BAD_ARG;
  const int total = offsets[n_problems];
  if (total < 0 || (total > 0 && (!kp_prev || !depth || !flow || !flow_out))) return DYNOBA_ERR_BAD_ARG;
  for (int p = 0; p < n_problems; p++) if (offsets[p + 1] < offsets[p]) return DYNOBA_ERR_BAD_ARG;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count) return DYNOBA_ERR_CUDA;     // no CPU fallback
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess || prop.major < 10) return DYNOBA_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return DYNOBA_ERR_CUDA;
  const size_t np = (size_t)n_problems, nt = (size_t)std::max(total, 1);
  // one device block: [off | pose0 | pose_prev | calib | kp | depth | flow0 | flow_cur | flow_cand | pose_out | flow_out | e0 | e1 | it | inner]
  const size_t bytes = (np + 1)*4 + 8 + 8*(12*np*3 + 5*np + nt*(2 + 1 + 2 + 2 + 2 + 2) + 2*np) + 8*np;
  char* base = nullptr;
  if (cudaMalloc((void**)&base, bytes + 256) != cudaSuccess) { cudaGetLastError(); return DYNOBA_ERR_CUDA; }
  size_t o = 0; auto take = [&](size_t b) { char* p = base + o; o += (b + 15) & ~(size_t)15; return p; };
  StarBatch S{};
  S.nprob = n_problems;
  int* d_off = (int*)take((np + 1)*4);
  double* d_pose0 = (double*)take(96*np); double* d_prev = (double*)take(96*np); double* d_cal = (double*)take(40*np);
  double* d_kp = (double*)take(16*nt); double* d_depth = (double*)take(8*nt); double* d_flow0 = (double*)take(16*nt);
  S.flow_cur = (double*)take(16*nt); S.flow_cand = (double*)take(16*nt);
  S.pose_out = (double*)take(96*np); S.flow_out = (double*)take(16*nt); S.err_before = (double*)take(8*np); S.err_after = (double*)take(8*np);
  S.iterations = (int*)take(4*np); S.inner = (int*)take(4*np);
  cudaStream_t s; cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking);
  cudaMemcpyAsync(d_off, offsets, (np + 1)*4, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(d_pose0, pose_init, 96*np, cudaMemcpyHostToDevice, s); cudaMemcpyAsync(d_prev, pose_prev, 96*np, cudaMemcpyHostToDevice, s);
  cudaMemcpyAsync(d_cal, calib5, 40*np, cudaMemcpyHostToDevice, s);
  if (total > 0) { cudaMemcpyAsync(d_kp, kp_prev, 16*(size_t)total, cudaMemcpyHostToDevice, s); cudaMemcpyAsync(d_depth, depth, 8*(size_t)total, cudaMemcpyHostToDevice, s);
                   cudaMemcpyAsync(d_flow0, flow, 16*(size_t)total, cudaMemcpyHostToDevice, s); }
  S.off = d_off; S.pose0 = d_pose0; S.pose_prev = d_prev; S.calib = d_cal; S.kp = d_kp; S.depth = d_depth; S.flow0 = d_flow0;
  S.isig_flow = 1.0/flow_sigma; S.isig_prior = 1.0/flow_prior_sigma; S.huber_k = huber_k;
  if (prm) S.prm = *prm; else dynoba_lm_default_params(&S.prm);
  star_lm_kernel<<<n_problems, STAR_THREADS, 0, s>>>(S);
  cudaMemcpyAsync(pose_out, S.pose_out, 96*np, cudaMemcpyDeviceToHost, s);
  if (total > 0) cudaMemcpyAsync(flow_out, S.flow_out, 16*(size_t)total, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(err_before, S.err_before, 8*np, cudaMemcpyDeviceToHost, s); cudaMemcpyAsync(err_after, S.err_after, 8*np, cudaMemcpyDeviceToHost, s);
  cudaMemcpyAsync(iterations, S.iterations, 4*np, cudaMemcpyDeviceToHost, s); cudaMemcpyAsync(inner_iterations, S.inner, 4*np, cudaMemcpyDeviceToHost, s);
  const cudaError_t e1 = cudaStreamSynchronize(s), e2 = cudaGetLastError();
  cudaStreamDestroy(s); cudaFree(base);
  return (e1 == cudaSuccess && e2 == cudaSuccess) ? DYNOBA_OK : DYNOBA_ERR_CUDA;
}
