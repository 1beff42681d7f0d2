// kernels_window.cu -- K4 (round-1 final): landmark Schur scatter WITHOUT atomics in the inner loop.
//
// Landmarks of a factor block are cut (on the host, DevWindows) into chunks whose cliques live in a window of at most
// NLOC pose-like variables.  A CTA owns a stripe of the window's lower-triangular 6x6 blocks, ONE BLOCK PER THREAD,
// held in registers (output-stationary).  The chunk's landmarks are streamed through shared memory in batches: for
// every clique variable the staging warp leaves  What_a = A_a^T B M^T  (M = chol(V + lambda I)^-1, so
// What_a What_b^T = W_a Vinv W_b^T) and the whitened Jacobian tile A_a; every thread then adds
//     S_ab += [same factor] A_a^T A_b - What_a What_b^T
// for the landmarks whose clique contains both of its variables.  Each block is flushed once per chunk into the tiled
// band storage.  Arithmetic is the same as kernels_schur.cu (GTSAM's landmark elimination).
#include "internal.cuh"

namespace dynoba {

constexpr int WIN_THREADS = 256;
constexpr int WIN_BATCH = 8;          // landmarks staged per batch (one per warp)
constexpr int WIN_TMAX = 24;          // factors per landmark handled by this path (larger ones use kernels_schur.cu)

__device__ __forceinline__ double wsum(double x) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

// shared-memory record of one clique variable of one landmark
struct WinSlot { double What[18]; double A[18]; double pad; };   // row-major 6x3 / 3x6; 37 doubles: odd stride, no bank conflicts

template <int NP, int PCOL0, int LCOL>
__global__ void __launch_bounds__(WIN_THREADS)
schur_window_kernel(DevBlock blk, DevWindows Wn, DevBand B, double lambda, int* __restrict__ fail) {
  constexpr int JC = NP*6 + 3, SLOTS = WIN_TMAX*NP;
  extern __shared__ unsigned char wsm[];
  WinSlot* slots = reinterpret_cast<WinSlot*>(wsm);                              // [WIN_BATCH][SLOTS]
  signed char* map = reinterpret_cast<signed char*>(slots + WIN_BATCH*SLOTS);    // [WIN_BATCH][NLOC_MAX] local var -> slot
  unsigned char* sfac = reinterpret_cast<unsigned char*>(map + WIN_BATCH*WIN_NLOC_MAX);  // [WIN_BATCH][SLOTS] slot -> factor
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int4 job = Wn.jobs[blockIdx.x];                 // (chunk, stripe, first group, end group)
  const int chunk = job.x, g0 = job.z, g1 = job.w;
  const int nloc = Wn.chunk_nloc[chunk];
  const int* cvars = Wn.cvars + (size_t)chunk*WIN_NLOC_MAX;
  // my block (a >= b) of the window's lower triangle
  const int q = job.y*WIN_THREADS + threadIdx.x;
  int a = (int)((sqrt(8.0*q + 1.0) - 1.0)*0.5);
  while (a*(a + 1)/2 > q) a--;
  while ((a + 1)*(a + 2)/2 <= q) a++;
  const int b = q - a*(a + 1)/2;
  const bool active = a < nloc;
  // block rows of this stripe: a landmark whose clique does not reach them is not staged at all
  int a_lo, a_hi;
  { const int q0 = job.y*WIN_THREADS, q1 = min(q0 + WIN_THREADS - 1, nloc*(nloc + 1)/2 - 1);
    a_lo = (int)((sqrt(8.0*q0 + 1.0) - 1.0)*0.5); while (a_lo*(a_lo + 1)/2 > q0) a_lo--; while ((a_lo + 1)*(a_lo + 2)/2 <= q0) a_lo++;
    a_hi = (int)((sqrt(8.0*q1 + 1.0) - 1.0)*0.5); while (a_hi*(a_hi + 1)/2 > q1) a_hi--; while ((a_hi + 1)*(a_hi + 2)/2 <= q1) a_hi++; }
  double acc[36];
#pragma unroll
  for (int i = 0; i < 36; i++) acc[i] = 0.0;
  bool touched = false;

  for (int gb = g0; gb < g1; gb += WIN_BATCH) {
    __syncthreads();
    // ---------------- staging: warp w prepares landmark gb + w
    {
      const int g = gb + warp;
      signed char* mp = map + warp*WIN_NLOC_MAX;
      for (int i = lane; i < WIN_NLOC_MAX; i += 32) mp[i] = -1;
      __syncwarp();
      if (g < g1 && Wn.grp_win[g] == 1 && (job.y == 0 || ((int)Wn.grp_lmax[g] >= a_lo && (int)Wn.grp_lmin[g] <= a_hi))) {
        const int f0 = blk.grp_ptr[g], T = blk.grp_ptr[g + 1] - f0;   // T <= WIN_TMAX guaranteed by the host
        double V[9], gl[3], Bm[9], bb[3];
#pragma unroll
        for (int k = 0; k < 9; k++) V[k] = 0.0;
        gl[0] = gl[1] = gl[2] = 0.0;
        const bool have = lane < T;
        const int f = f0 + lane;
        double Aall[NP*18];     // every load of the staging step is issued up front: one memory latency per landmark
        int pidx[NP];
        if (have) {
#pragma unroll
          for (int r = 0; r < 3; r++) {
            bb[r] = blk.b[(size_t)r*blk.stride + f];
#pragma unroll
            for (int c = 0; c < 3; c++) Bm[r*3 + c] = blk.J[(size_t)(r*JC + LCOL + c)*blk.stride + f];
          }
#pragma unroll
          for (int s = 0; s < NP; s++) {
            pidx[s] = blk.idx[(size_t)s*blk.stride + f];
#pragma unroll
            for (int r = 0; r < 3; r++)
#pragma unroll
              for (int c = 0; c < 6; c++) Aall[s*18 + r*6 + c] = blk.J[(size_t)(r*JC + PCOL0 + 6*s + c)*blk.stride + f];
          }
#pragma unroll
          for (int c1 = 0; c1 < 3; c1++) {
#pragma unroll
            for (int r = 0; r < 3; r++) gl[c1] += Bm[r*3 + c1]*bb[r];
#pragma unroll
            for (int c2 = 0; c2 <= c1; c2++)
#pragma unroll
              for (int r = 0; r < 3; r++) V[c1*3 + c2] += Bm[r*3 + c1]*Bm[r*3 + c2];
          }
        }
#pragma unroll
        for (int c1 = 0; c1 < 3; c1++) {
          gl[c1] = wsum(gl[c1]);
#pragma unroll
          for (int c2 = 0; c2 <= c1; c2++) V[c1*3 + c2] = wsum(V[c1*3 + c2]);
          V[c1*3 + c1] += lambda;
        }
        // M = chol(V)^-1 (lower)
        bool ok = V[0] > 0;
        const double l00 = sqrt(V[0]), l10 = V[3]/l00, l20 = V[6]/l00;
        const double d1 = V[4] - l10*l10; ok = ok && d1 > 0;
        const double l11 = sqrt(d1), l21 = (V[7] - l20*l10)/l11;
        const double d2 = V[8] - l20*l20 - l21*l21; ok = ok && d2 > 0;
        const double l22 = sqrt(d2);
        if (!ok) { if (lane == 0) atomicOr(fail, 1); }   // landmark skipped: the trial step is rejected anyway
        else {
          const double m00 = 1.0/l00, m11 = 1.0/l11, m22 = 1.0/l22;
          const double m10 = -l10*m00*m11, m21 = -l21*m11*m22, m20 = -(l20*m00 + l21*m10)*m22;
          // y = M g_l
          const double y0 = m00*gl[0], y1 = m10*gl[0] + m11*gl[1], y2 = m20*gl[0] + m21*gl[1] + m22*gl[2];
          if (have) {
            // Bh = B M^T (3x3): Bh[r][c] = sum_k B[r][k] M[c][k]
            double Bh[9];
#pragma unroll
            for (int r = 0; r < 3; r++) {
              Bh[r*3 + 0] = Bm[r*3]*m00;
              Bh[r*3 + 1] = Bm[r*3]*m10 + Bm[r*3 + 1]*m11;
              Bh[r*3 + 2] = Bm[r*3]*m20 + Bm[r*3 + 1]*m21 + Bm[r*3 + 2]*m22;
            }
            // rb = b - Bh y
            double rb[3];
#pragma unroll
            for (int r = 0; r < 3; r++) rb[r] = bb[r] - (Bh[r*3]*y0 + Bh[r*3 + 1]*y1 + Bh[r*3 + 2]*y2);
#pragma unroll
            for (int s = 0; s < NP; s++) {
              const int sl = lane*NP + s;
              WinSlot& ws = slots[warp*SLOTS + sl];
              const double* A = Aall + s*18;
#pragma unroll
              for (int i = 0; i < 18; i++) ws.A[i] = A[i];
#pragma unroll
              for (int c = 0; c < 6; c++) {
#pragma unroll
                for (int k = 0; k < 3; k++) ws.What[c*3 + k] = A[c]*Bh[k] + A[6 + c]*Bh[3 + k] + A[12 + c]*Bh[6 + k];
                if (job.y == 0) atomicAdd(rhs_at(B, pidx[s]*6 + c), A[c]*rb[0] + A[6 + c]*rb[1] + A[12 + c]*rb[2]);
              }
              mp[Wn.lvar[(size_t)s*blk.stride + f]] = (signed char)sl;
              sfac[warp*SLOTS + sl] = (unsigned char)lane;
            }
          }
        }
      }
    }
    __syncthreads();
    // ---------------- accumulation: my block over the batch's landmarks
    if (active) {
#pragma unroll 1
      for (int w = 0; w < WIN_BATCH; w++) {
        const signed char* mp = map + w*WIN_NLOC_MAX;
        const int sa = mp[a], sb = mp[b];
        if (sa < 0 || sb < 0) continue;
        touched = true;
        const WinSlot& wa = slots[w*SLOTS + sa];
        const WinSlot& wb = slots[w*SLOTS + sb];
        double wbv[18];
#pragma unroll
        for (int i = 0; i < 18; i++) wbv[i] = wb.What[i];
#pragma unroll
        for (int r = 0; r < 6; r++) {
          const double x0 = wa.What[r*3], x1 = wa.What[r*3 + 1], x2 = wa.What[r*3 + 2];
#pragma unroll
          for (int c = 0; c < 6; c++) acc[r*6 + c] -= x0*wbv[c*3] + x1*wbv[c*3 + 1] + x2*wbv[c*3 + 2];
        }
        if (sfac[w*SLOTS + sa] == sfac[w*SLOTS + sb]) {
#pragma unroll
          for (int i = 0; i < 18; i++) wbv[i] = wb.A[i];
#pragma unroll
          for (int r = 0; r < 6; r++) {
            const double x0 = wa.A[r], x1 = wa.A[6 + r], x2 = wa.A[12 + r];
#pragma unroll
            for (int c = 0; c < 6; c++) acc[r*6 + c] += x0*wbv[c] + x1*wbv[6 + c] + x2*wbv[12 + c];
          }
        }
      }
    }
  }
  if (active && touched) {
    const int pa = cvars[a], pb = cvars[b];   // cvars ascending => pa >= pb
#pragma unroll
    for (int r = 0; r < 6; r++)
#pragma unroll
      for (int c = 0; c < 6; c++) {
        if (a == b && c > r) continue;
        atomicAdd(band_at(B, pa*6 + r, pb*6 + c), acc[r*6 + c]);
      }
  }
}

static size_t win_smem(int NP) {
  return (size_t)WIN_BATCH*(WIN_TMAX*NP*sizeof(WinSlot) + WIN_NLOC_MAX + WIN_TMAX*NP);
}

int launch_schur_window(const DevBlock& blk, const DevWindows& Wn, const DevBand& B, double lambda, int* fail, cudaStream_t s) {
  if (Wn.n_jobs == 0) return 0;
  static bool attr = false;
  if (!attr) {
    cudaFuncSetAttribute(schur_window_kernel<1, 0, 6>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)win_smem(1));
    cudaFuncSetAttribute(schur_window_kernel<2, 0, 12>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)win_smem(2));
    attr = true;
  }
  switch (blk.type) {
    case F_POSE2POINT3: case F_STEREO3:
      schur_window_kernel<1, 0, 6><<<Wn.n_jobs, WIN_THREADS, win_smem(1), s>>>(blk, Wn, B, lambda, fail); return 1;
    case F_HYBRID3: case F_HYBRID_STEREO3:
      schur_window_kernel<2, 0, 12><<<Wn.n_jobs, WIN_THREADS, win_smem(2), s>>>(blk, Wn, B, lambda, fail); return 1;
    default: return 0;
  }
}

}  // namespace dynoba
