// Microbenchmarks behind the kernel design decisions in DESIGN.md (run on the B200 box):
//   dependent-chain latency of DMUL / DADD / DFMA / FFMA, LDS latency, graph launch cost of an empty kernel.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void chain_f64(double* out, long long* cyc, double m, double d, int n, int mode) {
  double g = out[0];
  long long t0 = clock64();
  if (mode == 0) {
#pragma unroll 16
    for (int i = 0; i < n; ++i) g = __dadd_rn(d, __dmul_rn(m, g));
  } else if (mode == 1) {
#pragma unroll 16
    for (int i = 0; i < n; ++i) g = __dadd_rn(d, g);
  } else if (mode == 2) {
#pragma unroll 16
    for (int i = 0; i < n; ++i) g = __dmul_rn(m, g);
  } else {
#pragma unroll 16
    for (int i = 0; i < n; ++i) g = __fma_rn(m, g, d);
  }
  long long t1 = clock64();
  out[threadIdx.x] = g;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void chain_f32(float* out, long long* cyc, float m, float d, int n) {
  float g = out[0];
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) g = __fmaf_rn(m, g, d);
  long long t1 = clock64();
  out[threadIdx.x] = g;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void lds_chain(int* out, long long* cyc, int n) {
  __shared__ int s[1024];
  for (int i = threadIdx.x; i < 1024; i += blockDim.x) s[i] = (i + 32) & 1023;
  __syncthreads();
  int p = threadIdx.x;
  long long t0 = clock64();
#pragma unroll 16
  for (int i = 0; i < n; ++i) p = s[p];
  long long t1 = clock64();
  out[threadIdx.x] = p;
  if (threadIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void empty_kernel(int* p) { if (p && threadIdx.x == 9999) *p = 1; }

int main() {
  double* d_out; long long* d_cyc; float* f_out; int* i_out;
  cudaMalloc(&d_out, 1024 * 8); cudaMalloc(&d_cyc, 8); cudaMalloc(&f_out, 4096); cudaMalloc(&i_out, 4096);
  cudaMemset(d_out, 0, 1024 * 8); cudaMemset(f_out, 0, 4096);
  const int n = 4096;
  const char* names[] = {"DMUL+DADD", "DADD", "DMUL", "DFMA"};
  for (int threads : {8, 32}) {
    for (int mode = 0; mode < 4; ++mode) {
      long long c = 0;
      chain_f64<<<1, threads>>>(d_out, d_cyc, 0.9603, 0.125, n, mode);
      chain_f64<<<1, threads>>>(d_out, d_cyc, 0.9603, 0.125, n, mode);
      cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
      printf("threads=%2d %-10s %.2f cycles/iter\n", threads, names[mode], double(c) / n);
    }
  }
  long long c = 0;
  chain_f32<<<1, 32>>>(f_out, d_cyc, 0.96f, 0.125f, n); chain_f32<<<1, 32>>>(f_out, d_cyc, 0.96f, 0.125f, n);
  cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
  printf("FFMA chain %.2f cycles/iter\n", double(c) / n);
  lds_chain<<<1, 32>>>(i_out, d_cyc, n); lds_chain<<<1, 32>>>(i_out, d_cyc, n);
  cudaMemcpy(&c, d_cyc, 8, cudaMemcpyDeviceToHost);
  printf("LDS chain %.2f cycles/iter\n", double(c) / n);

  // graph launch cost: 64 dependent empty kernels in one graph
  cudaStream_t st; cudaStreamCreate(&st);
  for (int grid : {1, 148, 512}) {
    cudaGraph_t g; cudaGraphExec_t ge;
    cudaStreamBeginCapture(st, cudaStreamCaptureModeGlobal);
    for (int i = 0; i < 64; ++i) empty_kernel<<<grid, 256, 0, st>>>(nullptr);
    cudaStreamEndCapture(st, &g);
    cudaGraphInstantiate(&ge, g, 0);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int w = 0; w < 3; ++w) cudaGraphLaunch(ge, st);
    cudaEventRecord(a, st);
    for (int r = 0; r < 20; ++r) cudaGraphLaunch(ge, st);
    cudaEventRecord(b, st); cudaStreamSynchronize(st);
    float ms = 0; cudaEventElapsedTime(&ms, a, b);
    printf("graph of 64 serial empty kernels (grid %3d x 256): %.2f us per kernel\n", grid, ms * 1000 / (20 * 64));
    // same launched directly on the stream
    cudaEventRecord(a, st);
    for (int r = 0; r < 20 * 64; ++r) empty_kernel<<<grid, 256, 0, st>>>(nullptr);
    cudaEventRecord(b, st); cudaStreamSynchronize(st);
    cudaEventElapsedTime(&ms, a, b);
    printf("stream launches of empty kernels   (grid %3d x 256): %.2f us per kernel\n", grid, ms * 1000 / (20 * 64));
  }
  int clk = 0; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
  printf("clock rate attr %d kHz\n", clk);
  return 0;
}
