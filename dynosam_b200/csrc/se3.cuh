// se3.cuh -- fp64 SO(3)/SE(3) device math for the factor kernels.
// Semantics follow GTSAM 4.2.0 built with GTSAM_POSE3_EXPMAP/GTSAM_ROT3_EXPMAP (reference
// docker/Dockerfile.amd64:104-112): tangent [omega; v], retract(T, xi) = T * Expmap(xi).
// Everything is written for registers: fixed-size arrays, fully unrolled loops.
#pragma once
#include <cfloat>
#include <cmath>

namespace dynoba {

struct Pose {
  double R[9];  // row-major
  double t[3];
};

#define DYN_HD __host__ __device__ __forceinline__

DYN_HD void m3mul(const double* A, const double* B, double* C) {
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C[3*i+j] = A[3*i]*B[j] + A[3*i+1]*B[3+j] + A[3*i+2]*B[6+j];
}
DYN_HD void m3tmul(const double* A, const double* B, double* C) {  // A^T B
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) C[3*i+j] = A[i]*B[j] + A[3+i]*B[3+j] + A[6+i]*B[6+j];
}
DYN_HD void m3vec(const double* A, const double* v, double* o) {
#pragma unroll
  for (int i = 0; i < 3; i++) o[i] = A[3*i]*v[0] + A[3*i+1]*v[1] + A[3*i+2]*v[2];
}
DYN_HD void m3tvec(const double* A, const double* v, double* o) {
#pragma unroll
  for (int i = 0; i < 3; i++) o[i] = A[i]*v[0] + A[3+i]*v[1] + A[6+i]*v[2];
}
DYN_HD void skew3(const double* v, double* M) {
  M[0] = 0; M[1] = -v[2]; M[2] = v[1];
  M[3] = v[2]; M[4] = 0; M[5] = -v[0];
  M[6] = -v[1]; M[7] = v[0]; M[8] = 0;
}

DYN_HD void so3_expmap(const double* w, double* R) {
  const double th2 = w[0]*w[0] + w[1]*w[1] + w[2]*w[2];
  double W[9]; skew3(w, W);
  if (th2 <= DBL_EPSILON) {
#pragma unroll
    for (int i = 0; i < 9; i++) R[i] = W[i];
    R[0] += 1; R[4] += 1; R[8] += 1;
    return;
  }
  const double th = sqrt(th2), s = sin(th), s2 = sin(0.5*th), omc = 2.0*s2*s2;
  double K[9], KK[9];
#pragma unroll
  for (int i = 0; i < 9; i++) K[i] = W[i]/th;
  m3mul(K, K, KK);
#pragma unroll
  for (int i = 0; i < 9; i++) R[i] = s*K[i] + omc*KK[i];
  R[0] += 1; R[4] += 1; R[8] += 1;
}

DYN_HD void so3_logmap(const double* R, double* w) {
  const double tr = R[0] + R[4] + R[8];
  if (tr + 1.0 < 1e-10) {
    if (fabs(R[8] + 1.0) > 1e-5) {
      const double f = M_PI/sqrt(2.0 + 2.0*R[8]);
      w[0] = f*R[2]; w[1] = f*R[5]; w[2] = f*(1.0 + R[8]);
    } else if (fabs(R[4] + 1.0) > 1e-5) {
      const double f = M_PI/sqrt(2.0 + 2.0*R[4]);
      w[0] = f*R[1]; w[1] = f*(1.0 + R[4]); w[2] = f*R[7];
    } else {
      const double f = M_PI/sqrt(2.0 + 2.0*R[0]);
      w[0] = f*(1.0 + R[0]); w[1] = f*R[3]; w[2] = f*R[6];
    }
    return;
  }
  double mag; const double tr3 = tr - 3.0;
  if (tr3 < -1e-7) { const double th = acos((tr - 1.0)/2.0); mag = th/(2.0*sin(th)); }
  else mag = 0.5 - tr3/12.0;
  w[0] = mag*(R[7] - R[5]); w[1] = mag*(R[2] - R[6]); w[2] = mag*(R[3] - R[1]);
}

DYN_HD void se3_expmap(const double* xi, Pose& P) {
  const double* w = xi; const double* v = xi + 3;
  so3_expmap(w, P.R);
  const double th2 = w[0]*w[0] + w[1]*w[1] + w[2]*w[2];
  if (th2 > DBL_EPSILON) {
    const double wv = w[0]*v[0] + w[1]*v[1] + w[2]*v[2];
    const double c[3] = { w[1]*v[2] - w[2]*v[1], w[2]*v[0] - w[0]*v[2], w[0]*v[1] - w[1]*v[0] };
    double Rc[3]; m3vec(P.R, c, Rc);
#pragma unroll
    for (int i = 0; i < 3; i++) P.t[i] = (c[i] - Rc[i] + w[i]*wv)/th2;
  } else { P.t[0] = v[0]; P.t[1] = v[1]; P.t[2] = v[2]; }
}

DYN_HD void se3_logmap(const Pose& P, double* xi) {
  double w[3]; so3_logmap(P.R, w);
  const double t = sqrt(w[0]*w[0] + w[1]*w[1] + w[2]*w[2]);
  xi[0] = w[0]; xi[1] = w[1]; xi[2] = w[2];
  if (t < 1e-10) { xi[3] = P.t[0]; xi[4] = P.t[1]; xi[5] = P.t[2]; return; }
  const double wn[3] = { w[0]/t, w[1]/t, w[2]/t };
  double W[9]; skew3(wn, W);
  const double Tan = tan(0.5*t);
  double WT[3], WWT[3]; m3vec(W, P.t, WT); m3vec(W, WT, WWT);
#pragma unroll
  for (int i = 0; i < 3; i++) xi[3+i] = P.t[i] - (0.5*t)*WT[i] + (1.0 - t/(2.0*Tan))*WWT[i];
}

DYN_HD void se3_compose(const Pose& a, const Pose& b, Pose& o) {
  double R[9], t[3]; m3mul(a.R, b.R, R); m3vec(a.R, b.t, t);
#pragma unroll
  for (int i = 0; i < 9; i++) o.R[i] = R[i];
#pragma unroll
  for (int i = 0; i < 3; i++) o.t[i] = t[i] + a.t[i];
}
DYN_HD void se3_inverse(const Pose& a, Pose& o) {
  double R[9], t[3];
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) R[3*i+j] = a.R[3*j+i];
  m3vec(R, a.t, t);
#pragma unroll
  for (int i = 0; i < 9; i++) o.R[i] = R[i];
#pragma unroll
  for (int i = 0; i < 3; i++) o.t[i] = -t[i];
}
DYN_HD void se3_between(const Pose& a, const Pose& b, Pose& o) {  // a^-1 b
  Pose ai; se3_inverse(a, ai); se3_compose(ai, b, o);
}
DYN_HD void se3_retract(const Pose& P, const double* xi, Pose& o) {
  Pose E; se3_expmap(xi, E); se3_compose(P, E, o);
}
DYN_HD void se3_local(const Pose& a, const Pose& b, double* xi) {  // Logmap(a^-1 b)
  Pose d; se3_between(a, b, d); se3_logmap(d, xi);
}
DYN_HD void se3_transform_from(const Pose& P, const double* p, double* o) {
  m3vec(P.R, p, o); o[0] += P.t[0]; o[1] += P.t[1]; o[2] += P.t[2];
}
DYN_HD void se3_transform_to(const Pose& P, const double* p, double* o) {
  const double d[3] = { p[0]-P.t[0], p[1]-P.t[1], p[2]-P.t[2] }; m3tvec(P.R, d, o);
}
// Ad(T) = [[R,0],[[t]x R, R]], row-major 6x6
DYN_HD void se3_adjoint(const Pose& P, double* Ad) {
  double tx[9], txR[9]; skew3(P.t, tx); m3mul(tx, P.R, txR);
#pragma unroll
  for (int i = 0; i < 36; i++) Ad[i] = 0.0;
#pragma unroll
  for (int i = 0; i < 3; i++)
#pragma unroll
    for (int j = 0; j < 3; j++) {
      Ad[6*i+j] = P.R[3*i+j];
      Ad[6*(i+3)+j] = txR[3*i+j];
      Ad[6*(i+3)+j+3] = P.R[3*i+j];
    }
}

}  // namespace dynoba
