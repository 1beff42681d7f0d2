// kernels_band.cu -- K5: Cholesky factorisation and solve of the reduced camera/object-motion system.
//
// Replaces GTSAM's multifrontal Cholesky on the COLAMD ordering ([GTSAM-ext]) for the
// reduced system that is left after the landmarks are eliminated.  With pose-like variables ordered by frame
// the system is banded (half-width = the co-visibility window, <= max track age), stored as 32x32 tiles.
// tcgen05/UMMA has no fp64 path, so the tile kernels are fp64 FMA code.
//
// Factorisation = one persistent DATAFLOW kernel (no grid-wide barriers), band_cholesky_dataflow_kernel_v3:
//   * worker warps own one tile task at a time: every left-looking update T_IK -= L_IJ L_KJ^T (fp64 tensor-core MMA,
//     operands prefetched with cp.async) as soon as the per-tile "done" flags of the operand tiles are released, then
//     the TRSM against L_KK;
//   * one spine CTA per band problem owns the sequential chain potrf(K,K) -> trsm(K+1,K) -> update+potrf(K+1,K+1) and
//     the two tiles below it, in shared memory / registers, so the chain never waits on an L2 round trip.
// band_cholesky_dataflow_kernel (4-warp spine, DYNOBA_SPINE=2) is the previous generation, kept for A/B runs.
// Solve = explicit inverses of the diagonal tiles (one warp each, fully parallel); the forward substitution is folded
// into the factorisation as one task per column, the backward sweep runs on a cluster of 8 CTAs per band problem.
#include <algorithm>
#include <cstdlib>
#include <cstdio>
#include <vector>
#include <cooperative_groups.h>
#include "internal.cuh"

namespace cg = cooperative_groups;

namespace dynoba {

constexpr int CH_WARPS = 4;
constexpr int TS = 40, TSZ = TILE*TS;      // column stride of shared-memory tiles that feed MMA operand fragments: conflict-free loads

__device__ __forceinline__ int ld_acquire(const int* p) {
  int v;
  asm volatile("ld.acquire.gpu.global.s32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void st_release(int* p, int v) {
  asm volatile("st.release.gpu.global.s32 [%0], %1;" :: "l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ void wait_flag(const int* f, int lane) {
  if (lane == 0) {
    int spins = 0;
    while (ld_acquire(f) == 0) { if (++spins > 8) __nanosleep(64); }
  }
  __syncwarp();
}
// optional timeline trace (DYNOBA_CHOL_TRACE): global timer at the moment a flag of problem 0 is released
__device__ long long* g_trace = nullptr;
__device__ int* g_trace_base = nullptr;
__device__ long long g_trace_len = 0;
__device__ __forceinline__ void trace_flag(int* f) {
  if (g_trace) {
    const long long off = f - g_trace_base;
    if (off >= 0 && off < g_trace_len) { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); g_trace[off] = t; }
  }
}
__device__ long long* g_strace = nullptr;     // [NT][16] spine event times of problem 0 (DYNOBA_CHOL_TRACE)
__device__ __forceinline__ void strace(int col, int slot, int lane) {
  if (g_strace && blockIdx.x == 0 && lane == 0 && col >= 0) { long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); g_strace[(size_t)col*16 + slot] = t; }
}
__device__ __forceinline__ void set_flag(int* f, int lane) {
  // the lanes' tile stores are ordered before lane 0's release by the warp barrier (release is cumulative).  No
  // __threadfence(): that is MEMBAR.SC.GPU + ERRBAR + CCTL.IVALL, microseconds per flag on the two-die part.
  __syncwarp();
  if (lane == 0) { st_release(f, 1); trace_flag(f); }
}

// warp-level Cholesky of a 32x32 tile held one row per lane; returns false when a pivot is not positive
__device__ __forceinline__ bool warp_potrf(double (&row)[TILE], int lane) {
  bool ok = true;
#pragma unroll
  for (int k = 0; k < TILE; k++) {
    const double d = __shfl_sync(0xffffffffu, row[k], k);
    if (!(d > 0.0)) ok = false;
    const double inv = rsqrt(d);
    const double l = (lane == k) ? d*inv : row[k]*inv;   // column k of L, valid for lane >= k
    row[k] = l;
#pragma unroll
    for (int c = k + 1; c < TILE; c++) {
      const double lc = __shfl_sync(0xffffffffu, l, c);
      row[c] -= l*lc;     // only the lower triangle (lane >= c) is meaningful
    }
  }
  return ok;
}

// acc(row = lane, 32 cols) -= A(row = lane, 32 k) * B^T with B staged in shared memory as sB[k*32 + c] = B[c][k]
__device__ __forceinline__ void tile_gemm_sub(double (&acc)[TILE], const double (&a)[TILE], const double* sB) {
#pragma unroll
  for (int k = 0; k < TILE; k++) {
#pragma unroll
    for (int c = 0; c < TILE; c += 2) {
      const double2 b = *reinterpret_cast<const double2*>(sB + k*TILE + c);
      acc[c] -= a[k]*b.x; acc[c + 1] -= a[k]*b.y;
    }
  }
}
// x(row = lane) <- x * L^-T with L staged as sL[k*32 + c] = L[c][k]; sinv[c] = 1 / L[c][c].  Blocked by 8 columns: the
// serial chain is 8 x (mul, fma) per panel, the rank-8 update of the columns behind a panel has 24/16/8 independent chains.
__device__ __forceinline__ void tile_trsm(double (&x)[TILE], const double* sL, const double* sinv) {
#pragma unroll
  for (int p = 0; p < 4; p++) {
#pragma unroll
    for (int kk = 0; kk < 8; kk++) {
      const int k = 8*p + kk;
      const double l = x[k]*sinv[k];
      x[k] = l;
#pragma unroll
      for (int c = k + 1; c < 8*p + 8; c++) x[c] -= l*sL[k*TILE + c];
    }
    if (p < 3) {
#pragma unroll
      for (int c = 8*p + 8; c < TILE; c += 2) {
#pragma unroll
        for (int j = 0; j < 8; j++) {
          const double2 b = *reinterpret_cast<const double2*>(sL + (8*p + j)*TILE + c);
          x[c] -= x[8*p + j]*b.x; x[c + 1] -= x[8*p + j]*b.y;
        }
      }
    }
  }
}
__device__ __forceinline__ void tile_load(const double* t, double (&r)[TILE], int lane) {
#pragma unroll
  for (int c = 0; c < TILE; c++) r[c] = __ldcg(t + c*TILE + lane);
}
__device__ __forceinline__ void tile_store(double* t, const double (&r)[TILE], int lane) {
#pragma unroll
  for (int c = 0; c < TILE; c++) t[c*TILE + lane] = r[c];
}
// stage a tile held one row per lane into shared memory (s[c*32 + row]) + reciprocals of its diagonal
__device__ __forceinline__ void tile_stage(double* s, const double (&r)[TILE], int lane) {
  __syncwarp();
#pragma unroll
  for (int c = 0; c < TILE; c++) s[c*TILE + lane] = r[c];
  __syncwarp();
}


// Blocked variant for the spine: 4 panels of 8 columns.  Inside a panel the right-looking updates only touch the
// panel's columns (<= 7 shuffles per step); the rank-8 update of the trailing columns reads the panel from
// shared memory (sP[row*8 + j], 4 x LDS.128 per trailing column) instead of 8 shuffles per column.
__device__ __forceinline__ bool warp_potrf_blocked(double (&row)[TILE], int lane, double* sP) {
  bool ok = true;
#pragma unroll
  for (int p = 0; p < 4; p++) {
#pragma unroll
    for (int kk = 0; kk < 8; kk++) {
      const int k = 8*p + kk;
      const double d = __shfl_sync(0xffffffffu, row[k], k);
      if (!(d > 0.0)) ok = false;
      const double inv = rsqrt(d);
      const double l = (lane == k) ? d*inv : row[k]*inv;
      row[k] = l;
#pragma unroll
      for (int c = k + 1; c < 8*p + 8; c++) {
        const double lc = __shfl_sync(0xffffffffu, l, c);
        row[c] -= l*lc;
      }
    }
    if (p < 3) {
      __syncwarp();
#pragma unroll
      for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2*>(sP + lane*8 + j) = make_double2(row[8*p + j], row[8*p + j + 1]);
      __syncwarp();
#pragma unroll
      for (int c = 8*p + 8; c < TILE; c++) {
        double acc = row[c];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const double2 b = *reinterpret_cast<const double2*>(sP + c*8 + j);
          acc -= row[8*p + j]*b.x; acc -= row[8*p + j + 1]*b.y;
        }
        row[c] = acc;
      }
    }
  }
  return ok;
}
constexpr int LT_STRIDE = 34;   // padded row stride of the row-major copy of L_KK used by the spine TRSM
// x(row = lane) <- x * L^-T with L staged row-major: sLt[c*LT_STRIDE + k] = L[c][k]; sinv[c] = 1 / L[c][c]
__device__ __forceinline__ void tile_trsm_rm(double (&x)[TILE], const double* sLt, const double* sinv) {
#pragma unroll
  for (int c = 0; c < TILE; c++) {
    double s = x[c];
#pragma unroll
    for (int k = 0; k + 1 < c; k += 2) {
      const double2 l = *reinterpret_cast<const double2*>(sLt + c*LT_STRIDE + k);
      s -= x[k]*l.x; s -= x[k + 1]*l.y;
    }
    if (c & 1) s -= x[c - 1]*sLt[c*LT_STRIDE + c - 1];
    x[c] = s*sinv[c];
  }
}
// acc[j] (16 columns c0..c0+15 of row = lane) -= sum_k a[k] * B[c0+j][k], B staged as sB[k*32 + c]
__device__ __forceinline__ void tile_gemm_sub_half(double (&acc)[16], const double (&a)[TILE], const double* sB, int c0) {
#pragma unroll
  for (int k = 0; k < TILE; k++) {
#pragma unroll
    for (int j = 0; j < 16; j += 2) {
      const double2 b = *reinterpret_cast<const double2*>(sB + k*TILE + c0 + j);
      acc[j] -= a[k]*b.x; acc[j + 1] -= a[k]*b.y;
    }
  }
}
__device__ __forceinline__ void named_bar(int id, int nthreads) {
  asm volatile("bar.sync %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}
__device__ __forceinline__ void named_arrive(int id, int nthreads) {
  asm volatile("bar.arrive %0, %1;" :: "r"(id), "r"(nthreads) : "memory");
}

// Panel-fused spine: warp 0 factors the diagonal tile in 4 panels of 8 columns and PUBLISHES each finished panel
// (sPan[p][row*8 + j] = L[row][8p + j], sIv[k] = 1 / L[k][k]) with a non-blocking barrier arrival; the TRSM warps
// follow one panel behind (warp_trsm_follow), so the triangular solves of the sub-diagonal tiles overlap the potrf
// instead of starting after it.
__device__ __forceinline__ bool warp_potrf_publish(double (&row)[TILE], int lane, double* sPan, double* sIv) {
  bool ok = true;
#pragma unroll
  for (int p = 0; p < 4; p++) {
    double* sP = sPan + p*256;
#pragma unroll
    for (int kk = 0; kk < 8; kk++) {
      const int k = 8*p + kk;
      const double d = __shfl_sync(0xffffffffu, row[k], k);
      if (!(d > 0.0)) ok = false;
      const double inv = rsqrt(d);
      const double l = (lane == k) ? d*inv : row[k]*inv;
      row[k] = l;
      if (lane == k) sIv[k] = inv;
#pragma unroll
      for (int c = k + 1; c < 8*p + 8; c++) {
        const double lc = __shfl_sync(0xffffffffu, l, c);
        row[c] -= l*lc;
      }
    }
    __syncwarp();
#pragma unroll
    for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2*>(sP + lane*8 + j) = make_double2(row[8*p + j], row[8*p + j + 1]);
    __syncwarp();
    named_arrive(4 + p, 96);
    if (p < 3) {
#pragma unroll
      for (int c = 8*p + 8; c < TILE; c++) {
        double acc = row[c];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const double2 b = *reinterpret_cast<const double2*>(sP + c*8 + j);
          acc -= row[8*p + j]*b.x; acc -= row[8*p + j + 1]*b.y;
        }
        row[c] = acc;
      }
    }
  }
  return ok;
}
// x(row = lane) <- x * L^-T, consuming the panels warp 0 publishes
__device__ __forceinline__ void warp_trsm_follow(double (&x)[TILE], const double* sPan, const double* sIv) {
#pragma unroll
  for (int p = 0; p < 4; p++) {
    const double* sP = sPan + p*256;
    named_bar(4 + p, 96);
#pragma unroll
    for (int kk = 0; kk < 8; kk++) {
      const int k = 8*p + kk;
      const double l = x[k]*sIv[k];
      x[k] = l;
#pragma unroll
      for (int c = k + 1; c < 8*p + 8; c++) x[c] -= l*sP[c*8 + kk];
    }
    if (p < 3) {
#pragma unroll
      for (int c = 8*p + 8; c < TILE; c++) {
        double acc = x[c];
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const double2 b = *reinterpret_cast<const double2*>(sP + c*8 + j);
          acc -= x[8*p + j]*b.x; acc -= x[8*p + j + 1]*b.y;
        }
        x[c] = acc;
      }
    }
  }
}

__device__ long long g_spine_dbg[16];
#define TS(i) do { const long long t_ = clock64(); if (lane == 0) dbgacc[i] += t_ - tlast; tlast = t_; } while (0)

// One band problem handed to the factorisation kernel: columns [Kbeg, Kend) are factored; tiles in columns >= Kend
// only receive the updates from the factored columns (their Schur complement), updates from columns < Kbeg are
// assumed applied already (second phase of the two-directional scheme).
struct CholProb { double* tiles; double* rhs; int NT, Kbeg, Kend; int* done; int* pre; int* ydone; };
struct CholJob { CholProb p[2]; int np, WB, skew; };

__device__ __forceinline__ void chol_worker(const struct CholJob& job, int wid, int nworkers, double* sb, double* sinv, double* stage, int lane, int dd0_lag);

// Tile roles (dd = I - K):
//   dd == 0, 1 : workers apply the updates from columns J <= K-2 ("pre"), the spine applies J = K-1 and finishes
//   dd == 2    : workers apply J <= K-1 ("pre"), the spine does the TRSM
//   dd >= 3    : workers apply J <= K-1 and do the TRSM once done(K,K) is released
// plus one "rhs" task per column that folds the forward substitution y_K = L_KK^-1 (g_K - sum_J L_KJ y_J) in.
// flags: done[o], pre[o] for tile o = K*(WB+1) + dd;  ydone[K]
__global__ void __launch_bounds__(CH_WARPS*32)
band_cholesky_dataflow_kernel(CholJob job, int* __restrict__ fail) {
  extern __shared__ __align__(16) double chol_smem[];   // [CH_WARPS + 1][TILE2] + [CH_WARPS][32] (+ padding that pins CTAs/SM)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int WB = job.WB, W1 = WB + 1;
  double* sb = chol_smem + (size_t)warp*TILE2;
  double* sinv = chol_smem + (size_t)(CH_WARPS + 1)*TILE2 + warp*TILE;

  if ((int)blockIdx.x < job.np) {
    // ------------------------------------------------------------------ spine CTA (4 warps, block-wide barriers)
    //   warp 0: potrf(K,K);  warp 1: TRSM (K+1,K);  warp 2: TRSM (K+2,K);  then all four warps split the two
    //   tile updates T(K+1,K+1) -= L(K+1,K) L(K+1,K)^T and T(K+2,K+1) -= L(K+2,K) L(K+1,K)^T by column halves.
    const CholProb P = job.p[blockIdx.x];
    int* done = P.done; int* pre = P.pre;
    const int NT = P.NT;
    double* sLt  = chol_smem;                       // [32*LT_STRIDE] row-major L_KK
    double* sX1  = chol_smem + 1152;                // [1024] L(K+1,K), sX1[k*32 + r]
    double* sX2  = sX1 + TILE2;                     // [1024] L(K+2,K)
    double* sD   = sX2 + TILE2;                     // [1024] next diagonal tile hand-off
    double* sXn  = sD + TILE2;                      // [1024] next x1 hand-off
    double* sP   = sXn + TILE2;                     // [256]  potrf panel
    double* sIv  = sP + 256;                        // [32]   1 / diag(L_KK)
    double r[TILE];                                 // warp 0: diagonal tile row; warp 1: x1 row; warp 2: x2 row
    long long dbgacc[12] = {0,0,0,0,0,0,0,0,0,0,0,0}; long long tlast = clock64();
    if (P.Kbeg >= P.Kend) return;
    if (warp == 0) { wait_flag(pre + (size_t)P.Kbeg*W1, lane); tile_load(P.tiles + (size_t)P.Kbeg*W1*TILE2, r, lane); }
    if (warp == 1 && P.Kbeg + 1 < NT) { wait_flag(pre + (size_t)P.Kbeg*W1 + 1, lane); tile_load(P.tiles + ((size_t)P.Kbeg*W1 + 1)*TILE2, r, lane); }
    double* sPan = sLt;                             // [4][256] published potrf panels (the row-major L_KK copy is gone)
    for (int K = P.Kbeg; K < P.Kend; K++) {
      const size_t oD = (size_t)K*W1;
      const bool last = K + 1 >= NT;
      const bool has2 = (K + 2 < NT) && (WB >= 2), hasn = (K + 2 < NT);
      double h[16];
      if (warp == 0) {
        TS(11);
        if (!warp_potrf_publish(r, lane, sPan, sIv) && lane == 0) atomicOr(fail, 2);
        TS(0);
        tile_store(P.tiles + oD*TILE2, r, lane);
        set_flag(done + oD, lane);
        TS(1);
        if (!last) {
          wait_flag(pre + oD + W1, lane);
          const double* t = P.tiles + (oD + W1)*TILE2;
#pragma unroll
          for (int j = 0; j < 16; j++) h[j] = __ldcg(t + (16 + j)*TILE + lane);
        }
        TS(2);
      } else if (!last && warp == 1) {
        warp_trsm_follow(r, sPan, sIv);
        tile_store(P.tiles + (oD + 1)*TILE2, r, lane);
        tile_stage(sX1, r, lane);
        set_flag(done + oD + 1, lane);
      } else if (!last && warp == 2) {
        if (has2) { wait_flag(pre + oD + 2, lane); tile_load(P.tiles + (oD + 2)*TILE2, r, lane); }
        else {
#pragma unroll
          for (int c = 0; c < TILE; c++) r[c] = 0.0;
        }
        warp_trsm_follow(r, sPan, sIv);                 // always follows: the panel barriers count three warps
        if (has2) {
          tile_store(P.tiles + (oD + 2)*TILE2, r, lane);
          tile_stage(sX2, r, lane);
          set_flag(done + oD + 2, lane);
        }
      } else if (!last) {
        wait_flag(pre + oD + W1, lane);
        const double* t = P.tiles + (oD + W1)*TILE2;
#pragma unroll
        for (int j = 0; j < 16; j++) h[j] = __ldcg(t + j*TILE + lane);
      }
      if (last) break;
      named_bar(2, 128);
      if (warp == 0) TS(4);
      if (warp == 0 || warp == 3) {
        const int c0 = warp == 0 ? 16 : 0;
        double a[TILE];
#pragma unroll
        for (int k = 0; k < TILE; k++) a[k] = sX1[k*TILE + lane];
        tile_gemm_sub_half(h, a, sX1, c0);
#pragma unroll
        for (int j = 0; j < 16; j++) sD[(c0 + j)*TILE + lane] = h[j];
      } else if (hasn) {
        const int c0 = warp == 1 ? 16 : 0;
        wait_flag(pre + oD + W1 + 1, lane);
        const double* t = P.tiles + (oD + W1 + 1)*TILE2;
#pragma unroll
        for (int j = 0; j < 16; j++) h[j] = __ldcg(t + (c0 + j)*TILE + lane);
        if (has2) {
          double a[TILE];
#pragma unroll
          for (int k = 0; k < TILE; k++) a[k] = sX2[k*TILE + lane];
          tile_gemm_sub_half(h, a, sX1, c0);
        }
#pragma unroll
        for (int j = 0; j < 16; j++) sXn[(c0 + j)*TILE + lane] = h[j];
      }
      if (warp == 0) TS(5);
      named_bar(3, 128);
      if (warp == 0) {
#pragma unroll
        for (int c = 0; c < TILE; c++) r[c] = sD[c*TILE + lane];
        TS(6);
      } else if (warp == 1 && hasn) {
#pragma unroll
        for (int c = 0; c < TILE; c++) r[c] = sXn[c*TILE + lane];
      }
    }
    // columns >= Kend are not factored here: hand the two partially updated tiles of column Kend back
    if (P.Kend < NT) {
      if (warp == 0) tile_store(P.tiles + (size_t)P.Kend*W1*TILE2, r, lane);
      if (warp == 1 && P.Kend + 1 < NT) tile_store(P.tiles + ((size_t)P.Kend*W1 + 1)*TILE2, r, lane);
    }
    if (blockIdx.x == 0 && warp == 0 && lane == 0) for (int i = 0; i < 12; i++) g_spine_dbg[i] = dbgacc[i];
    return;
  }
  // -------------------------------------------------------------------- workers
  chol_worker(job, ((int)blockIdx.x - job.np)*CH_WARPS + warp, ((int)gridDim.x - job.np)*CH_WARPS, sb, sinv, nullptr, lane, 2);
}

// Worker warp `wid` of `nworkers`: tile tasks in column-major order (see the role table above the kernels).
// ---- fp64 tensor-core tile update for the workers: T(32x32) -= LI LK^T as 128 x mma.m8n8k4 (DMMA).  The FMA version
// reads its B operand from shared memory (one LDS.128 per two FMAs) and saturates the SM's shared-memory pipe with four
// warps at about half the fp64 rate; the MMA fragments live in registers, so the inner loop has no memory operation.
// Fragment maps (g = lane >> 2, q = lane & 3):  A/B operand block b, k-step kb: L[8b + g][4kb + q];
// accumulator block (rb, cb): T[8rb + g][8cb + 2q + {0, 1}].  Tiles are column-major (element (r, c) at c*32 + r).
__device__ __forceinline__ void dmma(double& c0, double& c1, double a, double b) {
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0, %1}, {%2}, {%3}, {%0, %1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}
__device__ __forceinline__ void frag_load(const double* t, double (&f)[TILE], int lane) {
  const double* p = t + (lane & 3)*TILE + (lane >> 2);
#pragma unroll
  for (int b = 0; b < 4; b++)
#pragma unroll
    for (int kb = 0; kb < 8; kb++) f[b*8 + kb] = __ldcg(p + 4*kb*TILE + 8*b);
}
__device__ __forceinline__ void cfrag_load(const double* t, double (&c)[TILE], int lane) {
  const double* p = t + 2*(lane & 3)*TILE + (lane >> 2);
#pragma unroll
  for (int rb = 0; rb < 4; rb++)
#pragma unroll
    for (int cb = 0; cb < 4; cb++) { c[(rb*4 + cb)*2] = __ldcg(p + 8*cb*TILE + 8*rb); c[(rb*4 + cb)*2 + 1] = __ldcg(p + (8*cb + 1)*TILE + 8*rb); }
}
__device__ __forceinline__ void cfrag_store(double* t, const double (&c)[TILE], int lane) {
  double* p = t + 2*(lane & 3)*TILE + (lane >> 2);
#pragma unroll
  for (int rb = 0; rb < 4; rb++)
#pragma unroll
    for (int cb = 0; cb < 4; cb++) { p[8*cb*TILE + 8*rb] = c[(rb*4 + cb)*2]; p[(8*cb + 1)*TILE + 8*rb] = c[(rb*4 + cb)*2 + 1]; }
}
// c -= A B^T with A, B given as operand fragments
__device__ __forceinline__ void frag_gemm_sub(double (&c)[TILE], const double (&a)[TILE], const double (&b)[TILE]) {
#pragma unroll
  for (int kb = 0; kb < 8; kb++)
#pragma unroll
    for (int rb = 0; rb < 4; rb++) {
      const double na = -a[rb*8 + kb];
#pragma unroll
      for (int cb = 0; cb < 4; cb++) dmma(c[(rb*4 + cb)*2], c[(rb*4 + cb)*2 + 1], na, b[cb*8 + kb]);
    }
}

// asynchronous copy of one 32x32 tile (column-major, 8 KB) into a shared-memory tile with column stride TS
__device__ __forceinline__ void tile_prefetch(double* s, const double* t, int lane) {
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const int ch = lane + 32*i, c = ch >> 4, part = ch & 15;
    const unsigned dst = (unsigned)__cvta_generic_to_shared(s + c*TS + 2*part);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(dst), "l"(t + c*TILE + 2*part) : "memory");
  }
}
__device__ __forceinline__ void sfrag_load(const double* t, double (&f)[TILE], int lane) {
  const double* p = t + (lane & 3)*TS + (lane >> 2);
#pragma unroll
  for (int b = 0; b < 4; b++)
#pragma unroll
    for (int kb = 0; kb < 8; kb++) f[b*8 + kb] = p[4*kb*TS + 8*b];
}
// both flags set?  (non-blocking, acquire)
__device__ __forceinline__ bool flags_ready(const int* fa, const int* fb, int lane) {
  int ok = 0;
  if (lane == 0) ok = ld_acquire(fa) != 0 && ld_acquire(fb) != 0;
  return __shfl_sync(0xffffffffu, ok, 0) != 0;
}

// dd0_lag: the diagonal tile (dd == 0) receives the worker updates from columns J <= K - dd0_lag only (the spine owns the rest).
// stage: per-warp [2][2][TSZ] shared-memory operand buffers for the asynchronous prefetch of the next update, or nullptr.
__device__ __forceinline__ void chol_worker(const CholJob& job, int wid, int nworkers, double* sb, double* sinv, double* stage, int lane, int dd0_lag) {
  const int WB = job.WB, W1 = WB + 1;
  const int W2 = W1 + 1;                                 // tile tasks + the rhs task of the column
  int ncols = 0;
  for (int q = 0; q < job.np; q++) ncols = max(ncols, job.p[q].NT - job.p[q].Kbeg);
  // Task order: skewed wavefronts s = skew*K + dd instead of column-major.  A worker processes its tasks in order and
  // blocks on operands, so a task's lead over the spine has to cover its own serial work; the (WB - dd) updates of a tile
  // near the diagonal need more lead than the short tasks far from it.  With the skew the tiles of one column are handed
  // out over WB/skew columns, longest first.  Every operand of task (K, dd) has a strictly smaller s, so in-order
  // blocking cannot deadlock.  (Measured on C5: skew 2 is 1-2 % faster than column-major or a near/far split of the
  // worker pool; the column period is set by the row-chain hop latency, see DESIGN.md.)
  const int skew = job.skew, nj = (W2 + skew - 1)/skew;
  const long long per_s = (long long)job.np*nj;
  const long long ntasks = ((long long)skew*ncols + W2)*per_s;
  for (long long t = wid; t < ntasks; t += nworkers) {
    const int sidx = (int)(t/per_s); const int u = (int)(t - (long long)sidx*per_s);
    const int q = u/nj, j = u - q*nj;
    const int dd = sidx%skew + skew*j, c = sidx/skew - j;
    if (dd > W1 || c < 0 || c >= ncols) continue;
    const CholProb P = job.p[q];
    int* done = P.done; int* pre = P.pre; int* ydone = P.ydone;
    const int NT = P.NT, K = P.Kbeg + c, I = K + dd;
    if (K >= NT) continue;
    if (dd == W1) {
      // ---- rhs task: y_K = L_KK^-1 (g_K - sum_{J<K} L_KJ y_J); rows >= Kend only collect the factored columns' part
      double v = P.rhs[(size_t)K*TILE + lane];
      const int Jend = min(K, P.Kend);
      for (int J = max(P.Kbeg, K - WB); J < Jend; J++) {
        const size_t oK = (size_t)J*W1 + (K - J);
        wait_flag(done + oK, lane);
        double a[TILE];
        tile_load(P.tiles + oK*TILE2, a, lane);
        wait_flag(ydone + J, lane);
        const double yj = __ldcg(P.rhs + (size_t)J*TILE + lane);
#pragma unroll
        for (int k = 0; k < TILE; k++) v -= a[k]*__shfl_sync(0xffffffffu, yj, k);
      }
      if (K >= P.Kend) { P.rhs[(size_t)K*TILE + lane] = v; continue; }
      wait_flag(done + (size_t)K*W1, lane);
      double l[TILE];
      tile_load(P.tiles + (size_t)K*W1*TILE2, l, lane);
      double y = 0.0, mydiag = 1.0;
#pragma unroll
      for (int cc = 0; cc < TILE; cc++) if (lane == cc) mydiag = l[cc];
      const double rinv = 1.0/mydiag;
#pragma unroll
      for (int cc = 0; cc < TILE; cc++) {
        const double yc = __shfl_sync(0xffffffffu, v, cc)*__shfl_sync(0xffffffffu, rinv, cc);
        if (lane == cc) y = yc;
        if (lane > cc) v -= l[cc]*yc;
      }
      P.rhs[(size_t)K*TILE + lane] = y;
      set_flag(ydone + K, lane);
      continue;
    }
    if (I >= NT) continue;
    const size_t o = (size_t)K*W1 + dd;
    double acc[TILE];                       // accumulator fragments while the updates run, one row per lane afterwards
    double* tp = P.tiles + o*TILE2;
    cfrag_load(tp, acc, lane);
    const int Jlo = max(P.Kbeg, I - WB);
    // the spine finishes the tiles of columns <= Kend itself: it owns the last (dd0_lag - 1) updates of a diagonal tile and
    // the last update of a first sub-diagonal tile
    const int Jhi = min(dd == 0 ? (K <= P.Kend ? K - dd0_lag : K - 1) : (dd == 1 ? K - 2 : K - 1), P.Kend - 1);
    if (stage) {
      // software pipeline: while update J runs on the tensor pipe, the operand tiles of J+1 (when their flags are already
      // up) stream into the other shared-memory stage with cp.async -- the L2 latency leaves the critical path
      int st = 0; bool have = false;
      for (int J = Jlo; J <= Jhi; J++) {
        const size_t oI = (size_t)J*W1 + (I - J), oK = (size_t)J*W1 + (K - J);
        if (!have) {
          wait_flag(done + oK, lane);
          if (dd != 0) wait_flag(done + oI, lane);
          tile_prefetch(stage + (size_t)(st*2 + 1)*TSZ, P.tiles + oK*TILE2, lane);
          if (dd != 0) tile_prefetch(stage + (size_t)(st*2)*TSZ, P.tiles + oI*TILE2, lane);
          asm volatile("cp.async.commit_group;" ::: "memory");
        }
        bool next = false;
        if (J + 1 <= Jhi) {
          const size_t nI = (size_t)(J + 1)*W1 + (I - J - 1), nK = (size_t)(J + 1)*W1 + (K - J - 1);
          next = flags_ready(done + nK, done + (dd != 0 ? nI : nK), lane);
          if (next) {
            tile_prefetch(stage + (size_t)((st ^ 1)*2 + 1)*TSZ, P.tiles + nK*TILE2, lane);
            if (dd != 0) tile_prefetch(stage + (size_t)((st ^ 1)*2)*TSZ, P.tiles + nI*TILE2, lane);
            asm volatile("cp.async.commit_group;" ::: "memory");
          }
        }
        if (next) asm volatile("cp.async.wait_group 1;" ::: "memory"); else asm volatile("cp.async.wait_group 0;" ::: "memory");
        __syncwarp();
        double fb[TILE];
        sfrag_load(stage + (size_t)(st*2 + 1)*TSZ, fb, lane);
        if (dd == 0) frag_gemm_sub(acc, fb, fb);
        else {
          double fa[TILE];
          sfrag_load(stage + (size_t)(st*2)*TSZ, fa, lane);
          frag_gemm_sub(acc, fa, fb);
        }
        __syncwarp();            // the stage is free for the prefetch after next
        st ^= 1; have = next;
      }
    } else
    for (int J = Jlo; J <= Jhi; J++) {
      const size_t oI = (size_t)J*W1 + (I - J), oK = (size_t)J*W1 + (K - J);
      double fb[TILE];
      wait_flag(done + oK, lane);
      frag_load(P.tiles + oK*TILE2, fb, lane);
      if (dd == 0) frag_gemm_sub(acc, fb, fb);
      else {
        double fa[TILE];
        wait_flag(done + oI, lane);
        frag_load(P.tiles + oI*TILE2, fa, lane);
        frag_gemm_sub(acc, fa, fb);
      }
    }
    if (dd <= 2 || K >= P.Kend) {
      cfrag_store(tp, acc, lane);
      set_flag(pre + o, lane);
    } else {
      // accumulator fragments -> one row per lane, through the warp's shared-memory tile
      __syncwarp();
      cfrag_store(sb, acc, lane);
      __syncwarp();
#pragma unroll
      for (int cc = 0; cc < TILE; cc++) acc[cc] = sb[cc*TILE + lane];
      const size_t oD = (size_t)K*W1;
      wait_flag(done + oD, lane);
      __syncwarp();
#pragma unroll
      for (int cc = 0; cc < TILE; cc++) sb[cc*TILE + lane] = __ldcg(P.tiles + oD*TILE2 + cc*TILE + lane);
      __syncwarp();
      sinv[lane] = 1.0/sb[lane*TILE + lane];
      __syncwarp();
      tile_trsm(acc, sb, sinv);
      tile_store(tp, acc, lane);
      set_flag(done + o, lane);
    }
  }
}

// =====================================================================================================================
// Spine v3.  One CTA of 8 warps per band problem; the compute warps work out of shared memory / registers only and
// never touch global memory, fences or flags -- three I/O warps do that one column ahead / behind:
// (warp w issues on scheduler w % 4: each A warp shares its scheduler with a mostly sleeping I/O warp)
//   A pair (warps 0, 1)  alternate per column between  potrf(K,K)  and  accumulating the next diagonal tile
//                        D(K+1) = T(K+1,K+1) - X2(K-1) X2(K-1)^T - X1(K) X1(K)^T  (rank-8 updates as X1 panels appear)
//   B pair (warps 2, 3)  alternate between the TRSM of X1(K) = L(K+1,K), one panel behind the potrf, and accumulating
//                        Xn = T(K+2,K+1) - X2(K) X1(K)^T
//   C      (warp 6)      TRSM of X2(K) = L(K+2,K) and its store; runs on its own clock (its input arrives late from the workers)
//   IO0    (warp 7)      L(K,K), X1(K) -> global, done flags
//   IO1    (warp 5)      T(K+1,K+1), T(K+2,K+1) (worker-updated) -> shared memory, one column ahead
// Hand-offs are monotone event counters in shared memory (value c+1 = "done for column c").
constexpr int SP_WARPS = 8;
constexpr int PSTR = 10, PANSZ = TILE*PSTR;     // potrf panel [row][8], row stride 10 doubles: conflict-free LDS.128 per row
enum { EV_PAN = 0, EV_X1P = 4, EV_X2 = 8, EV_DIN = 9, EV_XNIN = 10, EV_ST_L = 11, EV_ST_X1 = 12, EV_ST_X2 = 13, EV_TK_D = 14, EV_TK_XN = 15, EV_N = 16 };

__device__ __forceinline__ void ev_signal(volatile int* ev, int i, int v, int lane) {
  __syncwarp();
  if (lane == 0) asm volatile("st.release.cta.shared.s32 [%0], %1;" :: "r"((unsigned)__cvta_generic_to_shared((const void*)(ev + i))), "r"(v) : "memory");
}
__device__ __forceinline__ void ev_wait(volatile int* ev, int i, int v, int lane) {
  if (lane == 0) {
    const unsigned a = (unsigned)__cvta_generic_to_shared((const void*)(ev + i));
    int cur, spins = 0;
    for (;;) {
      asm volatile("ld.acquire.cta.shared.s32 %0, [%1];" : "=r"(cur) : "r"(a) : "memory");
      if (cur >= v) break;
      if (++spins > 16) __nanosleep(20);
    }
  }
  __syncwarp();
}

// The spine's code must stay small: a warp that runs thousands of straight-line instructions once per column is bound by
// instruction fetch, not by the arithmetic (tools/ubench/potrf_bench.cu: 5.8k cycles per tile alone, 14-17k inside a
// 25k-instruction kernel).  So the panel loops below are ROLLED; the register tile keeps static indices through
// block-of-8 switches on the (warp-uniform) panel number.
#define SPINE_BLOCK_SWITCH(p, BODY) \
  do { if ((p) == 0) { constexpr int B8 = 0; BODY } else if ((p) == 1) { constexpr int B8 = 8; BODY } \
       else if ((p) == 2) { constexpr int B8 = 16; BODY } else { constexpr int B8 = 24; BODY } } while (0)

// Cholesky of a 32x32 tile held one row per lane, in 4 panels of 8 columns.  The 8x8 diagonal block of a panel is read
// back from shared memory by EVERY lane and factored redundantly in registers, each lane solving its own row against it
// on the way (no shuffles); the serial chain per column is rsqrt -> scale -> one FMA.  Finished panels are published:
// sPan[p][row*PSTR + j] = L[row][8p + j], sIv[k] = 1 / L[k][k], event EV_PAN + p.
__device__ __forceinline__ bool spine_potrf(double (&row)[TILE], int lane, double* sPan, double* sIv, volatile int* ev, int cval, long long (&prof)[3]) {
  bool ok = true;
#pragma unroll 1
  for (int p = 0; p < 4; p++) {
    const long long tp0 = clock64();
    double* sP = sPan + p*PANSZ;
    double x[8];
    SPINE_BLOCK_SWITCH(p, {
#pragma unroll
      for (int j = 0; j < 8; j++) x[j] = row[B8 + j];
    });
#pragma unroll
    for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2*>(sP + lane*PSTR + j) = make_double2(x[j], x[j + 1]);
    __syncwarp();
    double G[36];                                   // packed lower triangle, G[i(i+1)/2 + j] = block[i][j]
#pragma unroll
    for (int i = 0; i < 8; i++)
#pragma unroll
      for (int j = 0; j <= i; j++) G[i*(i + 1)/2 + j] = sP[(8*p + i)*PSTR + j];
    __syncwarp();
    const long long tp1 = clock64();
    double myinv = 0.0;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const double d = G[k*(k + 1)/2 + k];
      if (!(d > 0.0)) ok = false;
      const double inv = rsqrt(d);
      if ((lane & 7) == k) myinv = inv;      // published after the loop: a store inside it makes ptxas clone the rsqrt
      x[k] *= inv;
#pragma unroll
      for (int i = k + 1; i < 8; i++) G[i*(i + 1)/2 + k] *= inv;
#pragma unroll
      for (int j = k + 1; j < 8; j++) {
        const double ljk = G[j*(j + 1)/2 + k];
        x[j] -= x[k]*ljk;
#pragma unroll
        for (int i = j; i < 8; i++) G[i*(i + 1)/2 + j] -= G[i*(i + 1)/2 + k]*ljk;
      }
    }
    const long long tp2 = clock64();
    if (lane < 8) sIv[8*p + lane] = myinv;
#pragma unroll
    for (int j = 0; j < 8; j += 2) *reinterpret_cast<double2*>(sP + lane*PSTR + j) = make_double2(x[j], x[j + 1]);
    ev_signal(ev, EV_PAN + p, cval, lane);
    // rank-8 update of the columns behind the panel (blocks of 8 columns, static register indices)
#pragma unroll
    for (int blk = 1; blk < 4; blk++) {
      if (blk > p) {
#pragma unroll
        for (int cc = 0; cc < 8; cc++) {
          const int c = 8*blk + cc;
          double acc = row[c];
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            const double2 b = *reinterpret_cast<const double2*>(sP + c*PSTR + j);
            acc -= x[j]*b.x; acc -= x[j + 1]*b.y;
          }
          row[c] = acc;
        }
      }
    }
    const long long tp3 = clock64();
    prof[0] += tp1 - tp0; prof[1] += tp2 - tp1; prof[2] += tp3 - tp2;
  }
  return ok;
}
// x(row = lane) <- x * L^-T from the published panels; every finished panel of x is staged k-major (sX[k*TS + row]) and
// announced on ev_out + p (ev_out < 0: no per-panel events)
__device__ __forceinline__ void spine_trsm(double (&x)[TILE], int lane, const double* sPan, const double* sIv, volatile int* ev, int cval,
                                           double* sX, int ev_out) {
#pragma unroll 1
  for (int p = 0; p < 4; p++) {
    const double* sP = sPan + p*PANSZ;
    ev_wait(ev, EV_PAN + p, cval, lane);
    double xs[8];
    SPINE_BLOCK_SWITCH(p, {
#pragma unroll
      for (int j = 0; j < 8; j++) xs[j] = x[B8 + j];
    });
#pragma unroll
    for (int kk = 0; kk < 8; kk++) {
      const double l = xs[kk]*sIv[8*p + kk];
      xs[kk] = l;
#pragma unroll
      for (int j = kk + 1; j < 8; j++) xs[j] -= l*sP[(8*p + j)*PSTR + kk];
    }
#pragma unroll
    for (int j = 0; j < 8; j++) sX[(8*p + j)*TS + lane] = xs[j];
    if (ev_out >= 0) ev_signal(ev, ev_out + p, cval, lane);
#pragma unroll
    for (int blk = 1; blk < 4; blk++) {
      if (blk > p) {
#pragma unroll
        for (int cc = 0; cc < 8; cc++) {
          const int c = 8*blk + cc;
          double acc = x[c];
#pragma unroll
          for (int j = 0; j < 8; j += 2) {
            const double2 b = *reinterpret_cast<const double2*>(sP + c*PSTR + j);
            acc -= xs[j]*b.x; acc -= xs[j + 1]*b.y;
          }
          x[c] = acc;
        }
      }
    }
  }
}
// Rank-32 update of an accumulator tile held as MMA fragments (see frag_gemm_sub): c -= XA XB^T, XA / XB k-major tiles in
// shared memory with column stride TS, applied in four rank-8 steps (two k-steps of the m8n8k4 MMA each) as the panels
// of XB appear (ev_base >= 0: wait for event ev_base + p first).  16 LDS.64 + 32 DMMA per step: next to nothing on the
// shared-memory pipe, which the FMA formulation (one LDS.128 per two FMAs) saturates when several spine warps run it.
__device__ __forceinline__ void spine_rank32(double (&c)[TILE], const double* xa, const double* xb, bool same,
                                             volatile int* ev, int ev_base, int cval, int lane) {
  const int off = (lane & 3)*TS + (lane >> 2);
#pragma unroll 1
  for (int p = 0; p < 4; p++) {
    if (ev_base >= 0) ev_wait(ev, ev_base + p, cval, lane);
#pragma unroll
    for (int kk = 0; kk < 2; kk++) {
      const int kb = 2*p + kk;
      double fa[4], fb[4];
#pragma unroll
      for (int b = 0; b < 4; b++) fb[b] = xb[off + 4*kb*TS + 8*b];
#pragma unroll
      for (int b = 0; b < 4; b++) fa[b] = same ? -fb[b] : -xa[off + 4*kb*TS + 8*b];
#pragma unroll
      for (int rb = 0; rb < 4; rb++)
#pragma unroll
        for (int cb = 0; cb < 4; cb++) dmma(c[(rb*4 + cb)*2], c[(rb*4 + cb)*2 + 1], fa[rb], fb[cb]);
    }
  }
}
// accumulator fragments <-> shared-memory tile with column stride S
template <int S>
__device__ __forceinline__ void cfrag_load_s(const double* t, double (&c)[TILE], int lane) {
  const double* p = t + 2*(lane & 3)*S + (lane >> 2);
#pragma unroll
  for (int rb = 0; rb < 4; rb++)
#pragma unroll
    for (int cb = 0; cb < 4; cb++) { c[(rb*4 + cb)*2] = p[8*cb*S + 8*rb]; c[(rb*4 + cb)*2 + 1] = p[(8*cb + 1)*S + 8*rb]; }
}
template <int S>
__device__ __forceinline__ void cfrag_store_s(double* t, const double (&c)[TILE], int lane) {
  double* p = t + 2*(lane & 3)*S + (lane >> 2);
#pragma unroll
  for (int rb = 0; rb < 4; rb++)
#pragma unroll
    for (int cb = 0; cb < 4; cb++) { p[8*cb*S + 8*rb] = c[(rb*4 + cb)*2]; p[(8*cb + 1)*S + 8*rb] = c[(rb*4 + cb)*2 + 1]; }
}
// accumulator fragments -> one row per lane through a private shared-memory tile
__device__ __forceinline__ void cfrag_to_rows(double (&r)[TILE], double* scr, int lane) {
  __syncwarp();
  cfrag_store_s<TILE>(scr, r, lane);
  __syncwarp();
#pragma unroll
  for (int col = 0; col < TILE; col++) r[col] = scr[col*TILE + lane];
  __syncwarp();
}

__global__ void __launch_bounds__(SP_WARPS*32)
band_cholesky_dataflow_kernel_v3(CholJob job, int wk_warps, int* __restrict__ fail) {
  extern __shared__ __align__(16) double chol_smem[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((int)blockIdx.x >= job.np) {
    if (warp >= wk_warps) return;
    if (wk_warps <= 4) {       // room for the prefetch buffers: [warp][2 stages][2 tiles][TSZ] | [TILE2] | [TILE]
      double* base = chol_smem + (size_t)warp*(4*TSZ + TILE2 + TILE);
      chol_worker(job, ((int)blockIdx.x - job.np)*wk_warps + warp, ((int)gridDim.x - job.np)*wk_warps,
                  base + 4*TSZ, base + 4*TSZ + TILE2, base, lane, 3);
    } else {
      chol_worker(job, ((int)blockIdx.x - job.np)*wk_warps + warp, ((int)gridDim.x - job.np)*wk_warps,
                  chol_smem + (size_t)warp*TILE2, chol_smem + (size_t)SP_WARPS*TILE2 + warp*TILE, nullptr, lane, 3);
    }
    return;
  }
  const CholProb P = job.p[blockIdx.x];
  const int WB = job.WB, W1 = WB + 1, NT = P.NT, Kbeg = P.Kbeg, Kend = P.Kend;
  if (Kbeg >= Kend) return;
  double* sPan = chol_smem;                   // [2][4][PANSZ]
  double* sIv = sPan + 2*4*PANSZ;             // [2][32]
  double* sX1 = sIv + 2*TILE;                 // [2][TSZ]  k-major, column stride TS
  double* sX2 = sX1 + 2*TSZ;                  // [2][TSZ]
  double* sDin = sX2 + 2*TSZ;                 // [2][TSZ]  column-major, column stride TS
  double* sXnin = sDin + 2*TSZ;               // [2][TSZ]
  double* sScr = sXnin + 2*TSZ;               // [4][TILE2] layout-conversion scratch of the A / B warps
  volatile int* ev = reinterpret_cast<volatile int*>(sScr + 4*TILE2);
  if (threadIdx.x < EV_N) ev[threadIdx.x] = Kbeg;
  __syncthreads();
  int* done = P.done; int* pre = P.pre;
  const bool x2 = WB >= 2;
  double r[TILE];
  if (warp == 0 || warp == 1) {
    // ---------------------------------------------------------------- A pair: potrf / next diagonal tile
    const int i = warp;
    long long dbg[4] = {0, 0, 0, 0}, prof[3] = {0, 0, 0};
    for (int c = Kbeg - 1; c < Kend; c++) {
      if (c >= Kbeg && (c & 1) == i) {
        const long long t0 = clock64();
        ev_wait(ev, EV_ST_L, c - 1, lane);
        const long long t1 = clock64();
        strace(c, 0, lane);
        if (!spine_potrf(r, lane, sPan + (c & 1)*4*PANSZ, sIv + (c & 1)*TILE, ev, c + 1, prof) && lane == 0) atomicOr(fail, 2);
        dbg[0] += clock64() - t1; dbg[1] += t1 - t0;
        strace(c, 1, lane);
      } else if (((c + 1) & 1) == i && c + 1 < NT) {
        const long long t0 = clock64();
        ev_wait(ev, EV_DIN, c + 2, lane);
        cfrag_load_s<TS>(sDin + ((c + 1) & 1)*TSZ, r, lane);      // r: accumulator fragments until cfrag_to_rows
        ev_signal(ev, EV_TK_D, c + 2, lane);
        const long long t1 = clock64();
        strace(c + 1, 2, lane);
        // term 0: X2(c-1) = L(c+1, c-1), staged two columns ago and still in its buffer; term 1: X1(c), panel by panel
#pragma unroll 1
        for (int t = 0; t < 2; t++) {
          if (t == 0 ? !(x2 && c - 1 >= Kbeg) : !(c >= Kbeg)) continue;
          if (t == 0) { ev_wait(ev, EV_X2, c, lane); strace(c + 1, 3, lane); } else strace(c + 1, 4, lane);
          const double* xa = t == 0 ? sX2 + ((c - 1) & 1)*TSZ : sX1 + (c & 1)*TSZ;
          spine_rank32(r, xa, xa, true, ev, t == 0 ? -1 : EV_X1P, c + 1, lane);
        }
        cfrag_to_rows(r, sScr + i*TILE2, lane);
        const long long t2 = t1;
        strace(c + 1, 5, lane);
        dbg[2] += t1 - t0; dbg[3] += clock64() - t2;
        if (c + 1 == Kend) tile_store(P.tiles + (size_t)Kend*W1*TILE2, r, lane);   // not factored here: hand it back
      }
    }
    if (blockIdx.x == 0 && lane == 0) { for (int k = 0; k < 4; k++) g_spine_dbg[i*4 + k] = dbg[k]; if (i == 0) for (int k = 0; k < 3; k++) g_spine_dbg[8 + k] = prof[k]; }
  } else if (warp == 2 || warp == 3) {
    // ---------------------------------------------------------------- B pair: X1 TRSM / next X1 input
    const int i = warp - 2;
    for (int c = Kbeg - 1; c < Kend; c++) {
      if (c >= Kbeg && (c & 1) == i) {
        if (c + 1 < NT) {
          ev_wait(ev, EV_ST_X1, c - 1, lane);
          strace(c, 6, lane);
          spine_trsm(r, lane, sPan + (c & 1)*4*PANSZ, sIv + (c & 1)*TILE, ev, c + 1, sX1 + (c & 1)*TSZ, EV_X1P);
          strace(c, 7, lane);
        }
      } else if (((c + 1) & 1) == i && c + 2 < NT) {
        ev_wait(ev, EV_XNIN, c + 2, lane);
        cfrag_load_s<TS>(sXnin + ((c + 1) & 1)*TSZ, r, lane);
        ev_signal(ev, EV_TK_XN, c + 2, lane);
        strace(c + 1, 8, lane);
        if (x2 && c >= Kbeg) {
          ev_wait(ev, EV_X2, c + 1, lane);
          strace(c + 1, 9, lane);
          spine_rank32(r, sX2 + (c & 1)*TSZ, sX1 + (c & 1)*TSZ, false, ev, EV_X1P, c + 1, lane);
          strace(c + 1, 10, lane);
        }
        cfrag_to_rows(r, sScr + (2 + i)*TILE2, lane);
        if (c + 1 == Kend) tile_store(P.tiles + ((size_t)Kend*W1 + 1)*TILE2, r, lane);
      }
    }
  } else if (warp == 6) {
    // ---------------------------------------------------------------- C: X2 TRSM
    if (x2) for (int c = Kbeg; c < Kend; c++) if (c + 2 < NT) {
      wait_flag(pre + (size_t)c*W1 + 2, lane);
      tile_load(P.tiles + ((size_t)c*W1 + 2)*TILE2, r, lane);
      strace(c, 11, lane);
      spine_trsm(r, lane, sPan + (c & 1)*4*PANSZ, sIv + (c & 1)*TILE, ev, c + 1, sX2 + (c & 1)*TSZ, -1);
      ev_signal(ev, EV_X2, c + 1, lane);
      strace(c, 12, lane);
      // publish X2(c) from its staged copy (this warp runs on its own clock: the store is off the spine's critical path)
      {
        const double* sx = sX2 + (c & 1)*TSZ;
        double* t = P.tiles + ((size_t)c*W1 + 2)*TILE2;
#pragma unroll
        for (int col = 0; col < TILE; col++) t[col*TILE + lane] = sx[col*TS + lane];
      }
      set_flag(done + (size_t)c*W1 + 2, lane);
    }
  } else if (warp == 7) {
    // ---------------------------------------------------------------- IO0: L(c,c), X1(c) -> global
    for (int c = Kbeg; c < Kend; c++) {
      ev_wait(ev, EV_PAN + 3, c + 1, lane);
      const double* pan = sPan + (c & 1)*4*PANSZ;
      double* t = P.tiles + (size_t)c*W1*TILE2;
#pragma unroll
      for (int p = 0; p < 4; p++)
#pragma unroll
        for (int j = 0; j < 8; j += 2) {
          const double2 v = *reinterpret_cast<const double2*>(pan + p*PANSZ + lane*PSTR + j);
          t[(8*p + j)*TILE + lane] = v.x; t[(8*p + j + 1)*TILE + lane] = v.y;
        }
      set_flag(done + (size_t)c*W1, lane);
      ev_signal(ev, EV_ST_L, c + 1, lane);
      if (c + 1 < NT) {
        ev_wait(ev, EV_X1P + 3, c + 1, lane);
        const double* sx = sX1 + (c & 1)*TSZ;
#pragma unroll
        for (int col = 0; col < TILE; col++) t[TILE2 + col*TILE + lane] = sx[col*TS + lane];
        set_flag(done + (size_t)c*W1 + 1, lane);
        ev_signal(ev, EV_ST_X1, c + 1, lane);
      }
    }
  } else if (warp == 5) {
    // ---------------------------------------------------------------- IO1: inputs of column c+1 -> shared memory
    for (int c = Kbeg - 1; c < Kend; c++) if (c + 1 < NT) {
      const double* t = P.tiles + (size_t)(c + 1)*W1*TILE2;
      ev_wait(ev, EV_TK_D, c, lane);
      wait_flag(pre + (size_t)(c + 1)*W1, lane);
      tile_load(t, r, lane);
      double* sd = sDin + ((c + 1) & 1)*TSZ;
#pragma unroll
      for (int col = 0; col < TILE; col++) sd[col*TS + lane] = r[col];
      ev_signal(ev, EV_DIN, c + 2, lane);
      if (c + 2 < NT) {
        ev_wait(ev, EV_TK_XN, c, lane);
        wait_flag(pre + (size_t)(c + 1)*W1 + 1, lane);
        tile_load(t + TILE2, r, lane);
        double* sx = sXnin + ((c + 1) & 1)*TSZ;
#pragma unroll
        for (int col = 0; col < TILE; col++) sx[col*TS + lane] = r[col];
        ev_signal(ev, EV_XNIN, c + 2, lane);
      }
    }
  }
  // warp 4 has no role: it leaves scheduler 0 to the A warp that issues there
}

// sums the two Schur complements left in the middle separator: A.M += flip(B.M), A.rhs_M += flip(B.rhs_M)
__global__ void merge_middle_kernel(DevBand B) {
  const int np1 = B.n_pad - 1, lo = B.split_lo, hi = B.split_hi, s = hi - lo;
  const long long total = (long long)s*s;
  for (long long e = (long long)blockIdx.x*blockDim.x + threadIdx.x; e < total; e += (long long)gridDim.x*blockDim.x) {
    const int i = lo + (int)(e/s), j = lo + (int)(e%s);
    if (j > i) continue;
    B.tiles[band_index_wb(B.WB, i, j)] += B.tiles2[band_index_wb(B.WB, np1 - j, np1 - i)];
  }
  for (int p = lo + blockIdx.x*blockDim.x + threadIdx.x; p < hi; p += gridDim.x*blockDim.x) B.rhs[p] += B.rhs2[np1 - p];
}
// x_M (solved in A) -> the reversed copy that primes B's backward sweep
__global__ void prime_back_kernel(DevBand B) {
  const int np1 = B.n_pad - 1;
  for (int p = B.split_lo + blockIdx.x*blockDim.x + threadIdx.x; p < B.split_hi; p += gridDim.x*blockDim.x) B.rhs2[np1 - p] = B.rhs[p];
}
__global__ void gather_solution_kernel(DevBand B) {
  const int np1 = B.n_pad - 1;
  for (int p = blockIdx.x*blockDim.x + threadIdx.x; p < B.n_pad; p += gridDim.x*blockDim.x)
    B.dp[p] = p < B.split_hi ? B.rhs[p] : B.rhs2[np1 - p];
}

// explicit inverse of every diagonal tile: Linv[K] = L_KK^-1 (lower triangular), one warp per tile
__global__ void __launch_bounds__(128) diag_inverse_kernel(const double* __restrict__ tiles, int NT, int WB, double* __restrict__ linv) {
  __shared__ double sL[4][TILE2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int K = blockIdx.x*4 + warp;
  if (K >= NT) return;
  const double* t = tiles + (size_t)K*(WB + 1)*TILE2;
  double* s = sL[warp];
#pragma unroll
  for (int c = 0; c < TILE; c++) s[c*TILE + lane] = t[c*TILE + lane];
  __syncwarp();
  // lane j solves L x = e_j  -> column j of L^-1
  double x[TILE];
#pragma unroll
  for (int r = 0; r < TILE; r++) {
    double v = (r == lane) ? 1.0 : 0.0;
#pragma unroll
    for (int k = 0; k < r; k++) v -= s[k*TILE + r]*x[k];      // L[r][k]
    x[r] = v/s[r*TILE + r];
  }
  double* o = linv + (size_t)K*TILE2;
#pragma unroll
  for (int r = 0; r < TILE; r++) o[lane*TILE + r] = (r >= lane) ? x[r] : 0.0;   // element (r, j=lane) at j*32 + r
}

// butterfly: on entry lane r holds v[c] (c = 0..31); on exit every lane c returns sum over r of v_r[c]
__device__ __forceinline__ double warp_transpose_sum(double (&v)[TILE], int lane) {
#pragma unroll
  for (int o = 16; o >= 1; o >>= 1) {
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < o; i++) {
      const double send = up ? v[i] : v[i + o];
      const double recv = __shfl_xor_sync(0xffffffffu, send, o);
      v[i] = (up ? v[i + o] : v[i]) + recv;
    }
  }
  return v[0];
}

// backward sweep  x_J = Linv_JJ^T (y_J - sum_{I>J} L_IJ^T x_I) for J = Jtop-1 ... Jbot; rhs overwritten.
// One thread-block CLUSTER of 8 CTAs (8 SMs) per band problem: the off-diagonal tiles of a column are spread over
// 8 x 4 warps so the 240 KB a column reads come through eight SMs' load paths; per-CTA partial sums are all-gathered
// through distributed shared memory and every CTA finishes x_J itself (one cluster barrier per column).
// Columns >= Jtop are already solved (the middle separator of the two-directional scheme): their x primes the ring.
struct BackProb { const double* tiles; double* rhs; const double* linv; int NT, Jtop, Jbot; };
struct BackJob { BackProb p[2]; int WB; };
constexpr int BW_CL = 8, BW_TW = 4;      // cluster size, tile warps per CTA
constexpr int BW_PS = 34;                // padded column stride of warp 0's shared-memory tiles (conflict-free LDS.128 per column)
// dot product of 32 register values with a 32-vector in shared memory (broadcast reads)
__device__ __forceinline__ double dot32(const double (&a)[TILE], const double* x) {
  double s0 = 0.0, s1 = 0.0, s2 = 0.0, s3 = 0.0;
#pragma unroll
  for (int r = 0; r < TILE; r += 4) {
    const double2 u = *reinterpret_cast<const double2*>(x + r), w = *reinterpret_cast<const double2*>(x + r + 2);
    s0 += a[r]*u.x; s1 += a[r + 1]*u.y; s2 += a[r + 2]*w.x; s3 += a[r + 3]*w.y;
  }
  return (s0 + s1) + (s2 + s3);
}
// Backward sweep x_J = L_JJ^-T (y_J - sum_{I>J} L_IJ^T x_I), TWO columns per cluster barrier: the off-diagonal terms of
// columns J and J-1 that involve x_I, I > J, are independent of each other (and use the same x_I: tile (I,J) and tile
// (I,J-1) go to the same warp), only L(J,J-1)^T x_J has to wait for x_J and is done by warp 0 between the two solves.
// Every tile GEMV^T is COLUMN-per-lane: lane c holds column c of the tile (32 contiguous doubles) and reads x as
// broadcast LDS.128 -- 32 FMAs and no shuffles (the row-per-lane form needs a 31-shuffle transpose-sum per tile).
__global__ void __cluster_dims__(BW_CL, 1, 1) __launch_bounds__((BW_TW + 1)*32)
band_backward_cluster_kernel(BackJob job) {
  extern __shared__ __align__(16) double sm[];
  cg::cluster_group cl = cg::this_cluster();
  const int rank = (int)cl.block_rank();
  const BackProb P = job.p[blockIdx.x/BW_CL];
  const int NT = P.NT, WB = job.WB, W1 = WB + 1, ring = WB + 2;
  double* xs = sm;                                   // [ring][32] solved blocks (replicated in every CTA)
  double* lpart = sm + (size_t)ring*TILE;            // [2][BW_TW][32] partial sums of this CTA's tile warps, per column
  double* cpart = lpart + 2*BW_TW*TILE;              // [2][2][BW_CL][32] all-gathered per-CTA partials (pair parity, column)
  double* sv = cpart + 4*BW_CL*TILE;                 // [32] warp 0: right-hand side to broadcast
  double* sW0 = sv + TILE;                           // [2 stages][3][32*BW_PS] warp 0: Linv(J), Linv(J-1), L(J,J-1) of a pair, fetched with
                                                     // cp.async one pair ahead
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int I = P.Jtop + warp; I < min(NT, P.Jtop + WB); I += BW_TW + 1) xs[(size_t)(I % ring)*TILE + lane] = P.rhs[(size_t)I*TILE + lane];
  __syncthreads();
  const int myq = rank*BW_TW + (warp - 1);           // this warp's tile of column J; of column J-1 it takes tile myq + 1
  double tA[TILE], tB[TILE];                          // column `lane` of the prefetched tiles (warp 0: tA = Linv(J))
  auto load_col = [&](const double* tile, double (&t)[TILE]) {
    const double2* p = reinterpret_cast<const double2*>(tile + lane*TILE);
#pragma unroll
    for (int r = 0; r < TILE/2; r++) { const double2 u = p[r]; t[2*r] = u.x; t[2*r + 1] = u.y; }
  };
  // software prefetch: the tiles of the next pair are loaded while this pair is processed
  auto prefetchA = [&](int J) { if (J >= P.Jbot && myq < min(WB, NT - 1 - J)) load_col(P.tiles + (size_t)J*W1*TILE2 + (size_t)(myq + 1)*TILE2, tA); };
  auto prefetchB = [&](int J) { if (J - 1 >= P.Jbot && myq + 1 < min(WB, NT - J)) load_col(P.tiles + (size_t)(J - 1)*W1*TILE2 + (size_t)(myq + 2)*TILE2, tB); };
  auto fetch_w0 = [&](int J, int stage) {              // warp 0: the three tiles of pair (J, J-1) -> shared memory stage
    if (J >= P.Jbot) {
      double* dst = sW0 + (size_t)stage*3*TILE*BW_PS;
      const double* l0 = P.linv + (size_t)J*TILE2;
      const double* l1 = P.linv + (size_t)(J - 1)*TILE2;
      const double* tp = P.tiles + (size_t)(J - 1)*W1*TILE2 + TILE2;
      const bool two = J - 1 >= P.Jbot;
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const int ch = lane + 32*i, c = ch >> 4, part = ch & 15;
        const unsigned d = (unsigned)__cvta_generic_to_shared(dst + c*BW_PS + 2*part);
        asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(l0 + 2*ch) : "memory");
        if (two) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d + (unsigned)(TILE*BW_PS*8)), "l"(l1 + 2*ch) : "memory");
        if (two && WB >= 1) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d + (unsigned)(2*TILE*BW_PS*8)), "l"(tp + 2*ch) : "memory");
      }
    }
    asm volatile("cp.async.commit_group;" ::: "memory");   // always: keeps the group count in step with the pair count
  };
  // the factor is far larger than L2 and was written long ago: pull the tiles of the pair after next into L2 early, so
  // that the register / cp.async prefetches above see L2 latency instead of HBM latency
  auto l2_line = [&](const double* p) { asm volatile("prefetch.global.L2 [%0];" :: "l"(p)); };
  auto warm = [&](int J) {
    if (J < P.Jbot) return;
    if (warp >= 1) {
      if (myq < min(WB, NT - 1 - J)) { const double* t = P.tiles + (size_t)J*W1*TILE2 + (size_t)(myq + 1)*TILE2; l2_line(t + lane*16); l2_line(t + 512 + lane*16); }
      if (J - 1 >= P.Jbot && myq + 1 < min(WB, NT - J)) { const double* t = P.tiles + (size_t)(J - 1)*W1*TILE2 + (size_t)(myq + 2)*TILE2; l2_line(t + lane*16); l2_line(t + 512 + lane*16); }
    } else {
      const double* t = P.linv + (size_t)J*TILE2; l2_line(t + lane*16); l2_line(t + 512 + lane*16);
      if (J - 1 >= P.Jbot) {
        const double* u = P.linv + (size_t)(J - 1)*TILE2; l2_line(u + lane*16); l2_line(u + 512 + lane*16);
        const double* w = P.tiles + (size_t)(J - 1)*W1*TILE2 + TILE2; l2_line(w + lane*16); l2_line(w + 512 + lane*16);
      }
    }
  };
  if (warp >= 1) { prefetchA(P.Jtop - 1); prefetchB(P.Jtop - 1); } else fetch_w0(P.Jtop - 1, 0);
  warm(P.Jtop - 3); warm(P.Jtop - 5); warm(P.Jtop - 7);
  int par = 0;
  for (int J = P.Jtop - 1; J >= P.Jbot; J -= 2, par ^= 1) {
    const bool hasB = J - 1 >= P.Jbot;
    const int nbA = min(WB, NT - 1 - J), nbB = min(WB, NT - J);       // tiles below the diagonal in columns J, J-1
    warm(J - 8);
    if (warp >= 1) {
      double sA = 0.0, sB = 0.0;
      if (myq < nbA) {
        const double* xv = xs + (size_t)((J + 1 + myq) % ring)*TILE;
        sA += dot32(tA, xv);                                          // lane c: sum_r L[r][c] x[r]
        if (hasB && myq + 1 < nbB) sB += dot32(tB, xv);
      }
      for (int q = myq + BW_CL*BW_TW; q < nbA; q += BW_CL*BW_TW) {    // only when WB > 32 (tA / tB double as scratch)
        const double* xv = xs + (size_t)((J + 1 + q) % ring)*TILE;
        load_col(P.tiles + (size_t)J*W1*TILE2 + (size_t)(q + 1)*TILE2, tA);
        sA += dot32(tA, xv);
        if (hasB && q + 1 < nbB) { load_col(P.tiles + (size_t)(J - 1)*W1*TILE2 + (size_t)(q + 2)*TILE2, tB); sB += dot32(tB, xv); }
      }
      prefetchA(J - 2);                                                // in flight across the cluster barrier
      prefetchB(J - 2);
      lpart[(warp - 1)*TILE + lane] = sA;
      lpart[(BW_TW + warp - 1)*TILE + lane] = sB;
    }
    __syncthreads();
    if (warp == 0) {
      double sA = 0.0, sB = 0.0;
#pragma unroll
      for (int w = 0; w < BW_TW; w++) { sA += lpart[w*TILE + lane]; sB += lpart[(BW_TW + w)*TILE + lane]; }
#pragma unroll
      for (int r = 0; r < BW_CL; r++) {
        double* cp = cl.map_shared_rank(cpart, r) + (size_t)par*2*BW_CL*TILE;
        cp[rank*TILE + lane] = sA; cp[(BW_CL + rank)*TILE + lane] = sB;
      }
      fetch_w0(J - 2, par ^ 1);                                        // next pair's tiles
      tA[0] = P.rhs[(size_t)J*TILE + lane];                            // this pair's right-hand sides: latency under the barrier
      tA[1] = hasB ? P.rhs[(size_t)(J - 1)*TILE + lane] : 0.0;
    }
    cl.sync();
    if (warp == 0) {
      const double* cp = cpart + (size_t)par*2*BW_CL*TILE;
      const double* st = sW0 + (size_t)par*3*TILE*BW_PS;
      double v = tA[0], v2 = tA[1];
#pragma unroll
      for (int r = 0; r < BW_CL; r++) { v -= cp[r*TILE + lane]; v2 -= cp[(BW_CL + r)*TILE + lane]; }
      asm volatile("cp.async.wait_group 1;" ::: "memory");           // everything but the group issued above has landed
      sv[lane] = v;
      __syncwarp();
      auto col = [&](const double* tile) {
        const double2* lc = reinterpret_cast<const double2*>(tile + lane*BW_PS);
#pragma unroll
        for (int r = 0; r < TILE/2; r++) { const double2 u = lc[r]; tB[2*r] = u.x; tB[2*r + 1] = u.y; }
      };
      col(st);
      const double x = dot32(tB, sv);                                 // lane c: sum_r Linv[r][c] v[r]
      double* xj = xs + (size_t)(J % ring)*TILE;
      xj[lane] = x;
      if (rank == 0) P.rhs[(size_t)J*TILE + lane] = x;
      if (hasB) {
        __syncwarp();
        if (WB >= 1) { col(st + 2*TILE*BW_PS); v2 -= dot32(tB, xj); }   // (L(J,J-1)^T x_J)[lane]
        __syncwarp();
        sv[lane] = v2;
        __syncwarp();
        col(st + TILE*BW_PS);
        const double x2 = dot32(tB, sv);
        xs[(size_t)((J - 1) % ring)*TILE + lane] = x2;
        if (rank == 0) P.rhs[(size_t)(J - 1)*TILE + lane] = x2;
      }
    }
    __syncthreads();
  }
}

static int g_max_blocks = 0, g_spine_ver = 3, g_wk_warps = 4, g_skew = 2;
static size_t g_chol_smem = 0;

static void chol_init() {
  if (g_max_blocks) return;
  int dev = 0, sms = 0, per = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  // CTAs per SM: 1 keeps each spine CTA alone on its SM (DYNOBA_CHOL_BPS overrides for experiments)
  int bps = 1; if (const char* e = getenv("DYNOBA_CHOL_BPS")) bps = atoi(e) > 0 ? atoi(e) : 1;
  if (const char* e = getenv("DYNOBA_SPINE")) g_spine_ver = atoi(e) == 2 ? 2 : 3;
  if (const char* e = getenv("DYNOBA_SKEW")) g_skew = std::max(2, atoi(e));
  if (const char* e = getenv("DYNOBA_WK_WARPS")) g_wk_warps = std::max(1, std::min(SP_WARPS, atoi(e)));
  if (g_spine_ver == 2) g_wk_warps = CH_WARPS;
  const size_t need = g_spine_ver == 2 ? (size_t)(1152 + 4*TILE2 + 256 + 64)*sizeof(double)
                                       : std::max((size_t)(2*4*PANSZ + 2*TILE + 8*TSZ + 4*TILE2 + 64), (size_t)4*(4*TSZ + TILE2 + TILE))*sizeof(double);   // spine | workers with prefetch stages
  g_chol_smem = std::max(need, (size_t)(220*1024)/bps - 2048);
  const void* kern = g_spine_ver == 2 ? (const void*)band_cholesky_dataflow_kernel : (const void*)band_cholesky_dataflow_kernel_v3;
  const int nthr = g_spine_ver == 2 ? CH_WARPS*32 : SP_WARPS*32;
  cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)g_chol_smem);
  cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, kern, nthr, g_chol_smem);
  g_max_blocks = sms*(per > 0 ? per : 1);
}
static void chol_launch(const CholJob& job, int* fail, cudaStream_t s) {
  long long ntile = 0;
  for (int q = 0; q < job.np; q++) ntile += (long long)(job.p[q].NT - job.p[q].Kbeg)*(job.WB + 2);
  int grid = g_max_blocks;
  const long long want = (ntile + g_wk_warps - 1)/g_wk_warps + job.np;
  if (grid > want) grid = (int)want;
  if (grid < job.np + 1) grid = job.np + 1;
  CholJob j = job; j.skew = g_skew;
  // cooperative launch only for its co-residency guarantee (the flag waits need every warp resident)
  if (g_spine_ver == 2) {
    void* args[] = { (void*)&j, (void*)&fail };
    cudaLaunchCooperativeKernel((void*)band_cholesky_dataflow_kernel, dim3(grid), dim3(CH_WARPS*32), args, g_chol_smem, s);
  } else {
    int wk = g_wk_warps;
    void* args[] = { (void*)&j, (void*)&wk, (void*)&fail };
    cudaLaunchCooperativeKernel((void*)band_cholesky_dataflow_kernel_v3, dim3(grid), dim3(SP_WARPS*32), args, g_chol_smem, s);
  }
}

int launch_band_cholesky(const DevBand& B, int* flags, double* linv, int* fail, cudaStream_t s) {
  chol_init();
  const int W1 = B.WB + 1;
  const int NTA = B.two ? B.NTA : B.NT, NTB = B.two ? B.NTB : 0;
  const size_t nfl = (size_t)(NTA + NTB)*(2*W1 + 1);
  cudaMemsetAsync(flags, 0, nfl*sizeof(int), s);
  int* fA = flags; int* fB = flags + (size_t)NTA*(2*W1 + 1);
  auto prob = [&](double* tiles, double* rhs, int NT, int kb, int ke, int* f) {
    CholProb p; p.tiles = tiles; p.rhs = rhs; p.NT = NT; p.Kbeg = kb; p.Kend = ke; p.done = f; p.pre = f + (size_t)NT*W1; p.ydone = f + (size_t)2*NT*W1; return p; };
  int launches = 0;
  long long* dtrace = nullptr; long long* dstrace = nullptr; const long long trace_len = (long long)NTA*(2*W1 + 1);
  if (getenv("DYNOBA_CHOL_TRACE")) {
    cudaMalloc(&dtrace, trace_len*sizeof(long long)); cudaMemsetAsync(dtrace, 0, trace_len*sizeof(long long), s);
    cudaMemcpyToSymbolAsync(g_trace, &dtrace, sizeof(dtrace), 0, cudaMemcpyHostToDevice, s);
    cudaMemcpyToSymbolAsync(g_trace_base, &fA, sizeof(fA), 0, cudaMemcpyHostToDevice, s);
    cudaMemcpyToSymbolAsync(g_trace_len, &trace_len, sizeof(trace_len), 0, cudaMemcpyHostToDevice, s);
    cudaMalloc(&dstrace, (size_t)NTA*16*sizeof(long long)); cudaMemsetAsync(dstrace, 0, (size_t)NTA*16*sizeof(long long), s);
    cudaMemcpyToSymbolAsync(g_strace, &dstrace, sizeof(dstrace), 0, cudaMemcpyHostToDevice, s);
  }
  auto dump_trace = [&]() {     // after the first factorisation launch: flag-release times of columns in the middle of problem 0
    if (!dtrace) return;
    cudaStreamSynchronize(s);
    std::vector<long long> ht(trace_len); cudaMemcpy(ht.data(), dtrace, trace_len*sizeof(long long), cudaMemcpyDeviceToHost);
    long long* nul = nullptr; cudaMemcpyToSymbol(g_trace, &nul, sizeof(nul)); cudaFree(dtrace);
    const int K0 = std::max(4, (B.two ? B.split_lo/TILE : NTA)/2), nd = getenv("DYNOBA_CHOL_TRACE_ALL") ? W1 : std::min(W1, 8);
    const long long t0 = ht[(size_t)K0*W1];
    fprintf(stderr, "[trace] ns relative to done(K0,K0), K0 = %d; columns: done dd=0..%d | pre dd=0..2\n", K0, nd - 1);
    for (int K = K0 - 2; K < K0 + 10; K++) {
      fprintf(stderr, "[trace] K=%d done:", K);
      for (int d = 0; d < nd; d++) fprintf(stderr, " %7lld", ht[(size_t)K*W1 + d] ? ht[(size_t)K*W1 + d] - t0 : -1);
      fprintf(stderr, "  pre:");
      for (int d = 0; d < std::min(W1, 3); d++) fprintf(stderr, " %7lld", ht[(size_t)NTA*W1 + (size_t)K*W1 + d] ? ht[(size_t)NTA*W1 + (size_t)K*W1 + d] - t0 : -1);
      fprintf(stderr, "\n");
    }
    std::vector<long long> hs((size_t)NTA*16); cudaMemcpy(hs.data(), dstrace, hs.size()*sizeof(long long), cudaMemcpyDeviceToHost);
    cudaMemcpyToSymbol(g_strace, &nul, sizeof(nul)); cudaFree(dstrace);
    fprintf(stderr, "[strace] potrf start,end | D: taken, X2 seen, X1 term start, done | X1 trsm start,end | Xn: taken, X2 seen, done | X2 trsm start,end\n");
    for (int K = K0 - 1; K < K0 + 8; K++) {
      fprintf(stderr, "[strace] K=%d", K);
      for (int e = 0; e < 13; e++) { if (e == 2 || e == 6 || e == 8 || e == 11) fprintf(stderr, " |"); fprintf(stderr, " %7lld", hs[(size_t)K*16 + e] ? hs[(size_t)K*16 + e] - t0 : -1); }
      fprintf(stderr, "\n");
    }
  };
  if (!B.two) {
    CholJob job; job.np = 1; job.WB = B.WB; job.p[0] = prob(B.tiles, B.rhs, B.NT, 0, B.NT, fA); job.p[1] = job.p[0];
    chol_launch(job, fail, s); launches++;
    dump_trace();
    diag_inverse_kernel<<<(B.NT + 3)/4, 128, 0, s>>>(B.tiles, B.NT, B.WB, linv); launches++;
  } else {
    const int KmA = B.split_lo/TILE, KmB = NTB - (B.split_hi - B.split_lo)/TILE;
    CholJob job; job.np = 2; job.WB = B.WB;
    job.p[0] = prob(B.tiles, B.rhs, NTA, 0, KmA, fA); job.p[1] = prob(B.tiles2, B.rhs2, NTB, 0, KmB, fB);
    chol_launch(job, fail, s); launches++;                               // both halves, towards the middle
    dump_trace();
    merge_middle_kernel<<<64, 256, 0, s>>>(B); launches++;
    cudaMemsetAsync(fA, 0, (size_t)NTA*(2*W1 + 1)*sizeof(int), s);
    CholJob mid; mid.np = 1; mid.WB = B.WB; mid.p[0] = prob(B.tiles, B.rhs, NTA, KmA, NTA, fA); mid.p[1] = mid.p[0];
    chol_launch(mid, fail, s); launches++;                               // the middle separator
    diag_inverse_kernel<<<(NTA + 3)/4, 128, 0, s>>>(B.tiles, NTA, B.WB, linv); launches++;
    diag_inverse_kernel<<<(KmB + 3)/4, 128, 0, s>>>(B.tiles2, KmB, B.WB, linv + (size_t)NTA*TILE2); launches++;
  }
  if (getenv("DYNOBA_SPINE_DBG")) {
    long long hd[16]; cudaStreamSynchronize(s); cudaMemcpyFromSymbol(hd, g_spine_dbg, sizeof(hd));
    const char* nm2[12] = {"potrf", "store+stage+flag D", "barA", "wait+load Dnext half", "barB (trsm)", "gemm half + stage", "barC + reload", "-", "-", "-", "-", "loop"};
    const char* nm3[12] = {"A0 potrf", "A0 wait L stored", "A0 wait D input", "A0 X1 rank-8 (+waits)", "A1 potrf", "A1 wait L stored", "A1 wait D input", "A1 X1 rank-8 (+waits)", "A0 potrf: stage + block load", "A0 potrf: 8x8 factor", "A0 potrf: publish + trailing", "-"};
    const char** nm = g_spine_ver == 2 ? nm2 : nm3;
    for (int i = 0; i < 12; i++) fprintf(stderr, "[spine] %-22s %10.3f ms\n", nm[i], hd[i]/1.965e6);
  }
  return launches;
}

int launch_band_solve(const DevBand& B, const double* linv, cudaStream_t s) {
  const size_t smem = ((size_t)(B.WB + 2)*TILE + (size_t)(2*BW_TW + 4*BW_CL + 1)*TILE + 6*TILE*BW_PS)*sizeof(double);
  static bool attr = false;
  if (!attr) { cudaFuncSetAttribute(band_backward_cluster_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200*1024); attr = true; }
  const int nthr = (BW_TW + 1)*32;
  if (!B.two) {
    BackJob job; job.WB = B.WB; job.p[0] = BackProb{ B.tiles, B.rhs, linv, B.NT, B.NT, 0 }; job.p[1] = job.p[0];
    band_backward_cluster_kernel<<<BW_CL, nthr, smem, s>>>(job);          // forward sweep: folded into the factorisation
    return 1;
  }
  const int NTA = B.NTA, NTB = B.NTB, KmA = B.split_lo/TILE, KmB = NTB - (B.split_hi - B.split_lo)/TILE;
  BackJob mid; mid.WB = B.WB; mid.p[0] = BackProb{ B.tiles, B.rhs, linv, NTA, NTA, KmA }; mid.p[1] = mid.p[0];
  band_backward_cluster_kernel<<<BW_CL, nthr, smem, s>>>(mid);            // x of the middle separator
  prime_back_kernel<<<8, 256, 0, s>>>(B);
  BackJob job; job.WB = B.WB;
  job.p[0] = BackProb{ B.tiles, B.rhs, linv, NTA, KmA, 0 };
  job.p[1] = BackProb{ B.tiles2, B.rhs2, linv + (size_t)NTA*TILE2, NTB, KmB, 0 };
  band_backward_cluster_kernel<<<2*BW_CL, nthr, smem, s>>>(job);          // both halves, away from the middle
  gather_solution_kernel<<<148, 256, 0, s>>>(B);
  return 4;
}

}  // namespace dynoba
