// Micro-benchmark of the spine's 32x32 potrf variants in isolation (one warp), optionally with spinning sibling warps.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -Xcompiler -fopenmp -I../../dynosam_b200/csrc potrf_bench.cu -o potrf_bench
#include <cstdio>
#include "../../dynosam_b200/csrc/kernels_band.cu"
using namespace dynoba;

__global__ void potrf_bench(double* out, long long* cyc, int mode, int iters) {
  __shared__ double sPan[4*PANSZ];
  __shared__ double sIv[TILE];
  __shared__ double sP[256];
  __shared__ volatile int ev[EV_N];
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x < EV_N) ev[threadIdx.x] = 0;
  if (threadIdx.x == 0) stop = 0;
  __syncthreads();
  if (warp != 0) {   // sibling warps: the same polling loop the spine's waiting warps run
    if (lane == 0) { int spins = 0; while (stop == 0) { if (++spins > 16) __nanosleep(20); } }
    __syncwarp();
    return;
  }
  double row0[TILE], row[TILE];
#pragma unroll
  for (int c = 0; c < TILE; c++) { const int d = lane > c ? lane - c : c - lane; row0[c] = (d == 0 ? 8.0 : 0.0) + 1.0/(1.0 + d); }
  long long total = 0; bool ok = true;
  for (int it = 0; it < iters; it++) {
#pragma unroll
    for (int c = 0; c < TILE; c++) row[c] = row0[c];
    __syncwarp();
    const long long t0 = clock64();
    if (mode == 0) ok &= spine_potrf(row, lane, sPan, sIv, ev, it + 1);
    else if (mode == 1) ok &= warp_potrf_publish(row, lane, sPan, sIv);
    else if (mode == 2) ok &= warp_potrf_blocked(row, lane, sP);
    else ok &= warp_potrf(row, lane);
    total += clock64() - t0;
  }
  double s = 0.0;
#pragma unroll
  for (int c = 0; c < TILE; c++) if (c <= lane) s += row[c];
  out[lane] = s + (ok ? 0.0 : 1e300);
  if (lane == 0) { cyc[0] = total/iters; stop = 1; }
}

int main() {
  double* out; long long* cyc; cudaMalloc(&out, 32*8); cudaMalloc(&cyc, 8);
  const char* names[4] = {"spine_potrf (v3: 8x8 in registers, rolled)", "warp_potrf_publish (v2: shuffles, unrolled)", "warp_potrf_blocked", "warp_potrf (unblocked)"};
  for (int nw = 1; nw <= 8; nw *= 8) for (int mode = 0; mode < 4; mode++) {
    potrf_bench<<<1, 32*nw>>>(out, cyc, mode, 200);
    cudaDeviceSynchronize();
    long long h; double ho[32]; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); cudaMemcpy(ho, out, 256, cudaMemcpyDeviceToHost);
    double chk = 0; for (int i = 0; i < 32; i++) chk += ho[i];
    printf("warps %d  %-46s %8lld cycles/tile   checksum %.12f  (%s)\n", nw, names[mode], h, chk, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
