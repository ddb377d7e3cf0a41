// How expensive are fp32<->fp64 conversions (F2F) on B200?  The GAE scan needs ~6 per row-lane.
#include <cstdio>
#include <cuda_runtime.h>

__global__ void f2f_chain(float* out, long long* cyc, int n) {
  float x = out[threadIdx.x];
  double acc = 0.0;
  long long t0 = clock64();
#pragma unroll 8
  for (int i = 0; i < n; ++i) {
    double d = static_cast<double>(x);     // F2F.F64.F32
    acc = __dadd_rn(acc, d);
    x = static_cast<float>(acc);           // F2F.F32.F64
  }
  long long t1 = clock64();
  out[threadIdx.x] = x;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

// independent conversions: throughput per warp
__global__ void f2f_tput(const float* in, double* out, long long* cyc, int n) {
  float x0 = in[threadIdx.x], x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < n; ++i) {
    a0 = __dadd_rn(a0, static_cast<double>(x0));
    a1 = __dadd_rn(a1, static_cast<double>(x1));
    a2 = __dadd_rn(a2, static_cast<double>(x2));
    a3 = __dadd_rn(a3, static_cast<double>(x3));
    x0 += 1.f; x1 += 1.f; x2 += 1.f; x3 += 1.f;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__device__ __forceinline__ double f32_to_f64_bits(float f) {
  const unsigned int b = __float_as_uint(f);
  const unsigned int e = (b >> 23) & 0xffu;
  unsigned int hi = (b & 0x80000000u) | ((e + 896u) << 20) | ((b & 0x007fffffu) >> 3);
  const unsigned int lo = b << 29;
  if (e == 0u) hi = b & 0x80000000u;  // zero (denormals handled by the caller's slow path)
  return __hiloint2double(static_cast<int>(hi), static_cast<int>(lo));
}

__global__ void bits_tput(const float* in, double* out, long long* cyc, int n) {
  float x0 = in[threadIdx.x], x1 = x0 + 1.f, x2 = x0 + 2.f, x3 = x0 + 3.f;
  double a0 = 0, a1 = 0, a2 = 0, a3 = 0;
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < n; ++i) {
    a0 = __dadd_rn(a0, f32_to_f64_bits(x0));
    a1 = __dadd_rn(a1, f32_to_f64_bits(x1));
    a2 = __dadd_rn(a2, f32_to_f64_bits(x2));
    a3 = __dadd_rn(a3, f32_to_f64_bits(x3));
    x0 += 1.f; x1 += 1.f; x2 += 1.f; x3 += 1.f;
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

__global__ void dadd_tput(double* out, long long* cyc, int n) {
  double a0 = out[threadIdx.x], a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
  long long t0 = clock64();
#pragma unroll 4
  for (int i = 0; i < n; ++i) {
    a0 = __dadd_rn(a0, 1.5); a1 = __dadd_rn(a1, 1.5); a2 = __dadd_rn(a2, 1.5); a3 = __dadd_rn(a3, 1.5);
    a4 = __dadd_rn(a4, 1.5); a5 = __dadd_rn(a5, 1.5); a6 = __dadd_rn(a6, 1.5); a7 = __dadd_rn(a7, 1.5);
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

int main() {
  float* f; double* d; long long* c; long long h;
  cudaMalloc(&f, 1 << 16); cudaMalloc(&d, 1 << 20); cudaMalloc(&c, 8);
  cudaMemset(f, 0, 1 << 16); cudaMemset(d, 0, 1 << 20);
  const int n = 2048;
  for (int threads : {8, 32}) {
    f2f_chain<<<1, threads>>>(f, c, n); f2f_chain<<<1, threads>>>(f, c, n);
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("chain F2F.F64.F32 -> DADD -> F2F.F32.F64, %2d threads: %.1f cycles/iter (DADD alone is 8.4)\n", threads, double(h) / n);
  }
  for (int warps : {1, 4, 8, 16}) {
    f2f_tput<<<1, 32 * warps>>>(f, d, c, n); f2f_tput<<<1, 32 * warps>>>(f, d, c, n);
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("F2F.F64.F32 + DADD, 4 independent per thread, %2d warps/SM: %.1f cycles per (F2F+DADD) per warp-slot\n", warps, double(h) / (4.0 * n));
    bits_tput<<<1, 32 * warps>>>(f, d, c, n); bits_tput<<<1, 32 * warps>>>(f, d, c, n);
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("integer f32->f64 + DADD, 4 independent per thread,   %2d warps/SM: %.1f cycles\n", warps, double(h) / (4.0 * n));
    dadd_tput<<<1, 32 * warps>>>(d, c, n); dadd_tput<<<1, 32 * warps>>>(d, c, n);
    cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("DADD only, 8 independent per thread,                 %2d warps/SM: %.1f cycles per DADD\n", warps, double(h) / (8.0 * n));
  }
  return 0;
}
