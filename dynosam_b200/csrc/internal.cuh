// internal.cuh -- device data model shared by the kernel translation units (DESIGN.md "Data layout in HBM").
#pragma once
#include <cuda_runtime.h>
#include <cstdint>
#include "factors.cuh"

namespace dynoba {

constexpr int TILE = 32;           // band-Cholesky tile edge
constexpr int TILE2 = TILE*TILE;

// Variables, SoA, fp64.  Pose-like variables are stored in *solver order* (ordered by the frame hint),
// landmarks in group order (sorted by the first pose they touch).  strides are padded to 32.
struct DevVars {
  int np, np_stride;   double* pose;   // [12][np_stride]  rows 0-8 R (row-major), 9-11 t
  int nl, nl_stride;   double* point;  // [3][nl_stride]
  int nf, nf_stride;   double* flow;   // [2][nf_stride]
  int naux, naux_stride; const double* aux;  // [12][naux_stride]
  double K[6];
};

// One homogeneous factor block, sorted by landmark group, SoA.
struct DevBlock {
  int type, n, stride;
  const int* idx;       // [arity][stride]   pose slots: solver position; point/flow slots: device index
  const double* meas;   // [meas][stride]
  const double* isig;   // [sigma_dim][stride]  1/sigma
  int sigma_dim;
  const int* aux;       // [stride] or nullptr
  double robust_k;
  double* J;            // [dim*jcols][stride]  whitened, Huber-weighted Jacobian, element e = row*jcols + col
  double* b;            // [dim][stride]        rhs = -sqrt(w) * r_w
  // landmark groups (simple groups: exactly one landmark, all of its factors in this block)
  int n_groups;
  const int* grp_ptr;   // [n_groups+1] into the sorted factor range
  const int* grp_lmk;   // [n_groups]   device landmark index, or -1 when the group is handled by the general path
  double* num_scratch;  // numeric-Jacobian factor types: per-CTA partial sums of the column-parallel linearize kernel
};

// Band storage of the reduced (camera + object-motion) system, lower triangle, TILE x TILE tiles:
// tile (I,J), J <= I <= J+WB at tiles[(J*(WB+1) + (I-J))*TILE2], element (r,c) at c*TILE + r.
//
// Two-directional ("twisted") layout: the system may be stored as TWO band problems that share a middle separator M =
// positions [split_lo, split_hi):  problem A = positions [0, split_hi) in natural order, problem B = positions
// [split_lo, n_pad) REVERSED (local index n_pad-1-p).  A is eliminated forward, B is eliminated from the far end
// backwards at the same time; both leave their Schur complement in M, which is summed and factored last.
// An entry (i >= j) lives in A when i < split_hi, else in B at (n_pad-1-j, n_pad-1-i).
struct DevBand {
  int n, n_pad, NT, WB, bw;
  double* tiles;        // problem A (the whole system when two == 0): [NT_A*(WB+1)*TILE2]
  double* rhs;          // [NT_A*32]   g_S, then y = L^-1 g_S, then x
  size_t tile_count;    // tiles of A + tiles of B (one allocation, tiles2 follows tiles, then rhs, rhs2)
  int two, split_lo, split_hi, NTA, NTB;
  double* tiles2;       // problem B
  double* rhs2;
  double* dp;           // [n_pad] solution delta_p in solver order (== rhs when two == 0)
};
__host__ __device__ __forceinline__ size_t band_index_wb(int WB, int i, int j) {  // requires i >= j
  const int I = i >> 5, Jt = j >> 5;
  return ((size_t)Jt*(WB + 1) + (I - Jt))*TILE2 + (size_t)(j & 31)*TILE + (i & 31);
}
__host__ __device__ __forceinline__ size_t band_index(const DevBand& B, int i, int j) { return band_index_wb(B.WB, i, j); }
// address of entry (i >= j) / of rhs element p of the reduced system
__host__ __device__ __forceinline__ double* band_at(const DevBand& B, int i, int j) {
  if (!B.two || i < B.split_hi) return B.tiles + band_index_wb(B.WB, i, j);
  const int np1 = B.n_pad - 1;
  return B.tiles2 + band_index_wb(B.WB, np1 - j, np1 - i);
}
__host__ __device__ __forceinline__ double* rhs_at(const DevBand& B, int p) {
  if (!B.two || p < B.split_hi) return B.rhs + p;
  return B.rhs2 + (B.n_pad - 1 - p);
}

// Landmark groups the per-landmark kernels do not cover (chains of points, landmarks spanning factor blocks):
// factor references (block, sorted index) per group, points of a group contiguous from gl0 (gnl of them).
struct GeneralGroups {
  int n_groups;
  const DevBlock* blocks;   // device copy of every factor block descriptor
  const int* gptr;          // [n_groups+1] into refs
  const void* refs;         // [gptr[n_groups]] (int blk, int idx)
  const int* gl0;           // [n_groups] first device point index
  const int* gnl;           // [n_groups] number of points (<= 21)
};

// Window decomposition of a factor block for kernels_window.cu (built by the host in finalize)
constexpr int WIN_NLOC_MAX = 48;          // largest window (local pose-like variables of a chunk)
constexpr int WIN_TMAX = 24;              // factors per landmark on the window path (larger ones: kernels_schur.cu)
constexpr int WIN_SLOT_DOUBLES = 42;      // one staged clique variable of one factor: Wt[3][6] | A[3][6] | rb[3] | meta | pad[2]
constexpr int WIN_BATCH_SLOTS = 256;      // slots per shared-memory buffer of the accumulate kernel
constexpr int WIN_BATCH_LMK = 32;         // landmarks per batch
// The accumulate kernel tiles the window's lower triangle of 6x6 blocks into 8x4 block tiles, one tile per warp (a
// landmark's clique is a contiguous range of local variables, so whole tiles are active or idle together).
__host__ __device__ __forceinline__ int win_ntiles(int nloc) {
  int n = 0;
  for (int tr = 0; tr*8 < nloc; tr++) { const int rmax = (8*tr + 7 < nloc - 1) ? 8*tr + 7 : nloc - 1; n += rmax/4 + 1; }
  return n;
}
constexpr int WIN_ACC_WARPS = 16;         // tiles per CTA (stripe) of the accumulate kernel
struct DevWindows {
  int n_jobs;
  const int4* jobs;              // (chunk, stripe of the window's lower-triangular blocks, first batch, end batch)
  const int4* batches;           // (first group, end group, first factor, end factor): runs of window-path landmarks,
                                 // at most WIN_BATCH_LMK landmarks / WIN_BATCH_SLOTS slots, never crossing a chunk
  const int* chunk_nloc;         // [n_chunks]
  const int* cvars;              // [n_chunks][WIN_NLOC_MAX] solver positions, ascending
  const unsigned char* lvar;     // [npose_slots][stride] local variable of each factor's pose slot
  const unsigned char* grp_win;  // [n_groups] 0: general path, 1: window path, 2: per-landmark atomics path
  double* slots;                 // [n*npose_slots][WIN_SLOT_DOUBLES] staged slots in factor order (HBM scratch)
};

// ---- launchers (each returns the number of kernels it launched)
int launch_linearize(const DevBlock& blk, const DevVars& v, double* partials, cudaStream_t s);
int launch_error(const DevBlock& blk, const DevVars& v, double* partials, double* per_factor, cudaStream_t s);
int launch_sum(const double* partials, int n, double* out, cudaStream_t s);
int linearize_grid(int n);          // number of partial sums a linearize/error launch of n factors writes
int numeric_grid(int type, int n);  // CTAs of the column-parallel linearize kernel (0 for analytic factor types)

int launch_band_clear(const DevBand& B, double lambda, int add_damping, cudaStream_t s);
int launch_schur_simple(const DevBlock& blk, const unsigned char* grp_win, const DevBand& B, double lambda, int* fail, cudaStream_t s);
int launch_schur_window(const DevBlock& blk, const DevWindows& Wn, const DevBand& B, double lambda, int* fail, cudaStream_t s);
int launch_schur_general(const GeneralGroups& G, const DevBand& B, double lambda, int* fail, cudaStream_t s);
int launch_backsub_general(const GeneralGroups& G, const DevBand& B, double lambda, double* dl_point, int nl_stride,
                           double* partials, cudaStream_t s);
int general_grid(int n_groups);
int launch_pose_factors(const DevBlock& blk, const DevBand& B, cudaStream_t s);
int launch_band_cholesky(const DevBand& B, int* flags /*[2*NT*(WB+1) + NT]*/, double* linv /*[NT*TILE2]*/, int* fail, cudaStream_t s);
int launch_band_solve(const DevBand& B, const double* linv, cudaStream_t s);
int launch_backsub_simple(const DevBlock& blk, const DevBand& B, double lambda, double* dl_point, int nl_stride,
                          double* dl_flow, int nf_stride, double* partials, cudaStream_t s);
int backsub_grid(int n_groups);
int launch_pose_model(const DevBlock& blk, const DevBand& B, double* partials, cudaStream_t s);
int launch_pose_delta_norm(const DevBand& B, double lambda, double* partials, cudaStream_t s);
int pose_norm_grid(int n);
int launch_retract(const DevVars& cur, const DevVars& cand, const DevBand& B, const double* dl_point,
                   const double* dl_flow, cudaStream_t s);

}  // namespace dynoba
