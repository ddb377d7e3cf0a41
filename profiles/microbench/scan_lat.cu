// Isolated copy of the K2 scanner's per-chunk work (one warp, shared memory, 16 rows per chunk): cycles per chunk.
#include <cstdio>
#include <cuda_runtime.h>
constexpr int kR = 16;
template <int MODE>
__global__ void scan_chunks(double* out, long long* cyc, double gl, int zero, int n_chunks, int extra_warps_spin) {
  __shared__ double sd[9][kR * 32];
  __shared__ unsigned char sm[9][kR * 32];
  __shared__ unsigned long long bar;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < 9 * kR * 32; i += blockDim.x) {
    (&sd[0][0])[i] = 1e-3 * (i % 97);
    (&sm[0][0])[i] = (i % 13) != 0;
  }
  __syncthreads();
  if (warp != 0) return;
  double g = 0.0;
  long long t0 = clock64();
  for (int c = 0; c < n_chunks; ++c) {
    double* d_ = sd[c % 9];
    const unsigned char* m_ = sm[c % 9];
    double d[kR], m[kR];
    unsigned int landed = 0u;
#pragma unroll
    for (int r = 0; r < kR; ++r) {
      d[r] = d_[r * 32 + lane];
      const unsigned int f = m_[r * 32 + lane];
      m[r] = f ? gl : 0.0;
      landed |= f | static_cast<unsigned int>(__double2loint(d[r]));
    }
    if (MODE != 2) g = __hiloint2double(__double2hiint(g), __double2loint(g) + static_cast<int>(landed & static_cast<unsigned int>(zero)));
#pragma unroll
    for (int r = kR - 1; r >= 0; --r) {
      g = __dadd_rn(d[r], __dmul_rn(m[r], g));
      if (MODE != 1) d_[r * 32 + lane] = g;
    }
    __syncwarp();
  }
  long long t1 = clock64();
  out[threadIdx.x] = g;
  if (threadIdx.x == 0) cyc[0] = (t1 - t0) / n_chunks;
}
int main() {
  double* d_out; long long* d_cyc;
  cudaMalloc(&d_out, 4096 * 8); cudaMalloc(&d_cyc, 8);
  const char* names[] = {"loads first + chain + STS.64 (product)", "no stores", "no load-first dependency"};
  for (int threads : {32, 576}) {
    for (int mode = 0; mode < 3; ++mode) {
      for (int rep = 0; rep < 2; ++rep) {
        if (mode == 0) scan_chunks<0><<<1, threads>>>(d_out, d_cyc, 0.94, 0, 900, 0);
        if (mode == 1) scan_chunks<1><<<1, threads>>>(d_out, d_cyc, 0.94, 0, 900, 0);
        if (mode == 2) scan_chunks<2><<<1, threads>>>(d_out, d_cyc, 0.94, 0, 900, 0);
      }
      long long c = 0; cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
      printf("threads=%3d %-42s %lld cycles per 16-row chunk (%.1f per row)\n", threads, names[mode], c, c / 16.0);
    }
  }
  printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
  return 0;
}
