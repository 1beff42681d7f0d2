// kernels_band.cu -- K5: Cholesky factorisation and solve of the reduced camera/object-motion system.
//
// Replaces GTSAM's multifrontal Cholesky on the COLAMD ordering ([GTSAM-ext]) for the
// reduced system that is left after the landmarks are eliminated.  With pose-like variables ordered by frame
// the system is banded (half-width = the co-visibility window, <= max track age), stored as 32x32 tiles.
// tcgen05/UMMA has no fp64 path, so the tile kernels are fp64 FMA code; one persistent cooperative kernel
// walks the tile columns (right-looking), with the forward substitution of g_S folded in as an extra row.
#include <cooperative_groups.h>
#include "internal.cuh"

namespace cg = cooperative_groups;

namespace dynoba {

constexpr int CH_WARPS = 4;

// warp-level Cholesky of a 32x32 tile held one row per lane; returns false when a pivot is not positive
__device__ __forceinline__ bool warp_potrf(double (&row)[TILE], int lane) {
  bool ok = true;
#pragma unroll
  for (int k = 0; k < TILE; k++) {
    const double d = __shfl_sync(0xffffffffu, row[k], k);
    if (!(d > 0.0)) { ok = false; }
    const double s = sqrt(d), inv = 1.0/s;
    const double l = (lane == k) ? s : row[k]*inv;   // column k of L, valid for lane >= k
    row[k] = l;
#pragma unroll
    for (int c = k + 1; c < TILE; c++) {
      const double lc = __shfl_sync(0xffffffffu, l, c);
      row[c] -= l*lc;     // only the lower triangle (lane >= c) is meaningful
    }
  }
  return ok;
}

__global__ void __launch_bounds__(CH_WARPS*32) band_cholesky_kernel(DevBand B, int* __restrict__ fail) {
  cg::grid_group grid = cg::this_grid();
  __shared__ double sD[TILE2];                 // L_JJ (diagonal tile of the current column)
  __shared__ double sK[CH_WARPS][TILE2];       // per-warp staging of L_KJ
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int gw = blockIdx.x*CH_WARPS + warp, nw = gridDim.x*CH_WARPS;
  const int NT = B.NT, WB = B.WB;
  const size_t cs = (size_t)(WB + 1)*TILE2;    // tile-column stride

  if (gw == 0) {   // prologue: factor the first diagonal tile
    double row[TILE];
    double* t = B.tiles;
#pragma unroll
    for (int c = 0; c < TILE; c++) row[c] = t[c*TILE + lane];
    if (!warp_potrf(row, lane) && lane == 0) atomicOr(fail, 2);
#pragma unroll
    for (int c = 0; c < TILE; c++) t[c*TILE + lane] = row[c];
  }
  grid.sync();

  for (int J = 0; J < NT; J++) {
    double* colJ = B.tiles + (size_t)J*cs;
    const int nbelow = min(WB, NT - 1 - J);
    // ---- phase 1: L_IJ = T_IJ L_JJ^-T for the tiles below the diagonal, and y_J = L_JJ^-1 y_J
    for (int i = threadIdx.x; i < TILE2; i += blockDim.x) sD[i] = colJ[i];
    __syncthreads();
    for (int task = gw; task <= nbelow; task += nw) {
      if (task == nbelow) {   // rhs "row"
        double x = 0.0;
        // forward substitution done by the whole warp: lane c finalises y[c] in turn
        double yv = B.rhs[J*TILE + lane];
#pragma unroll
        for (int c = 0; c < TILE; c++) {
          const double yc = __shfl_sync(0xffffffffu, yv, c)/sD[c*TILE + c];
          if (lane == c) x = yc;
          if (lane > c) yv -= sD[c*TILE + lane]*yc;
        }
        B.rhs[J*TILE + lane] = x;
      } else {
        double* t = colJ + (size_t)(task + 1)*TILE2;
        double x[TILE];
#pragma unroll
        for (int c = 0; c < TILE; c++) x[c] = t[c*TILE + lane];
#pragma unroll
        for (int c = 0; c < TILE; c++) {
          double s = x[c];
#pragma unroll
          for (int k = 0; k < c; k++) s -= x[k]*sD[k*TILE + c];   // L_JJ[c][k]
          x[c] = s/sD[c*TILE + c];
        }
#pragma unroll
        for (int c = 0; c < TILE; c++) t[c*TILE + lane] = x[c];
      }
    }
    grid.sync();
    // ---- phase 2: trailing update T_IK -= L_IJ L_KJ^T (J < K <= I <= J+nbelow), y_K -= L_KJ y_J
    const int ntile = nbelow*(nbelow + 1)/2;
    for (int task = gw; task < ntile + nbelow; task += nw) {
      if (task >= ntile) {   // rhs update for K = J + 1 + (task - ntile)
        const int kk = task - ntile + 1;
        const double* lk = colJ + (size_t)kk*TILE2;
        double s = 0;
#pragma unroll 8
        for (int k = 0; k < TILE; k++) s += lk[k*TILE + lane]*B.rhs[J*TILE + k];
        B.rhs[(J + kk)*TILE + lane] -= s;
        continue;
      }
      // decode (ii >= kk) in 1..nbelow from the triangular index
      int ii = (int)((sqrt(8.0*task + 1.0) - 1.0)*0.5);
      while (ii*(ii + 1)/2 > task) ii--;
      while ((ii + 1)*(ii + 2)/2 <= task) ii++;
      const int kk = task - ii*(ii + 1)/2 + 1;
      ii += 1;
      const double* li = colJ + (size_t)ii*TILE2;
      const double* lk = colJ + (size_t)kk*TILE2;
      double* dst = B.tiles + (size_t)(J + kk)*cs + (size_t)(ii - kk)*TILE2;
      double* sk = sK[warp];
      __syncwarp();
#pragma unroll
      for (int c = 0; c < TILE; c++) sk[c*TILE + lane] = lk[c*TILE + lane];
      __syncwarp();
      double a[TILE], acc[TILE];
#pragma unroll
      for (int k = 0; k < TILE; k++) a[k] = li[k*TILE + lane];       // row `lane` of L_IJ
#pragma unroll
      for (int c = 0; c < TILE; c++) acc[c] = dst[c*TILE + lane];
#pragma unroll
      for (int k = 0; k < TILE; k++) {
#pragma unroll
        for (int c = 0; c < TILE; c++) acc[c] -= a[k]*sk[k*TILE + c];  // L_KJ[c][k], broadcast read
      }
      if (ii == 1 && kk == 1) {   // next diagonal tile is complete: factor it now
        if (!warp_potrf(acc, lane) && lane == 0) atomicOr(fail, 2);
      }
#pragma unroll
      for (int c = 0; c < TILE; c++) dst[c*TILE + lane] = acc[c];
    }
    grid.sync();
  }
}

int launch_band_cholesky(const DevBand& B, int* fail, cudaStream_t s) {
  static int max_blocks = 0;
  if (!max_blocks) {
    int dev = 0, sms = 0, per = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per, band_cholesky_kernel, CH_WARPS*32, 0);
    max_blocks = sms*(per > 0 ? per : 1);
  }
  const int ntask = B.WB*(B.WB + 1)/2 + B.WB + 1;
  int grid = (ntask + CH_WARPS - 1)/CH_WARPS;
  if (grid > max_blocks) grid = max_blocks;
  if (grid < 1) grid = 1;
  DevBand Bc = B;
  void* args[] = { (void*)&Bc, (void*)&fail };
  cudaLaunchCooperativeKernel((void*)band_cholesky_kernel, dim3(grid), dim3(CH_WARPS*32), args, 0, s);
  return 1;
}

// ---- backward substitution  L^T x = y  (single CTA; x overwrites rhs)
constexpr int BS_WARPS = 16;
__global__ void __launch_bounds__(BS_WARPS*32) band_backsolve_kernel(DevBand B) {
  __shared__ double part[BS_WARPS][TILE];
  __shared__ double sD[TILE2];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int NT = B.NT, WB = B.WB;
  const size_t cs = (size_t)(WB + 1)*TILE2;
  for (int J = NT - 1; J >= 0; J--) {
    const double* colJ = B.tiles + (size_t)J*cs;
    const int nbelow = min(WB, NT - 1 - J);
    double s = 0;
    for (int ii = 1 + warp; ii <= nbelow; ii += BS_WARPS) {
      const double* t = colJ + (size_t)ii*TILE2 + (size_t)lane*TILE;   // column `lane` of L_IJ
      const double* x = B.rhs + (size_t)(J + ii)*TILE;
#pragma unroll 8
      for (int r = 0; r < TILE; r++) s += t[r]*x[r];
    }
    part[warp][lane] = s;
    for (int i = threadIdx.x; i < TILE2; i += blockDim.x) sD[i] = colJ[i];
    __syncthreads();
    if (warp == 0) {
      double acc = B.rhs[J*TILE + lane];
#pragma unroll
      for (int w = 0; w < BS_WARPS; w++) acc -= part[w][lane];
      double x = 0;
#pragma unroll
      for (int c = TILE - 1; c >= 0; c--) {
        const double xc = __shfl_sync(0xffffffffu, acc, c)/sD[c*TILE + c];
        if (lane == c) x = xc;
        if (lane < c) acc -= sD[lane*TILE + c]*xc;     // L_JJ[c][lane]
      }
      B.rhs[J*TILE + lane] = x;
    }
    __syncthreads();
  }
}
int launch_band_backsolve(const DevBand& B, cudaStream_t s) {
  band_backsolve_kernel<<<1, BS_WARPS*32, 0, s>>>(B);
  return 1;
}

}  // namespace dynoba
