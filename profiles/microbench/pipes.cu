// Issue throughput of the instructions the GAE / loss kernels lean on (B200, sm_100a):
// warp-instructions per clock per SM for FADD, DADD, DMUL, DFMA, F2F.F64.F32, F2F.F32.F64, MUFU.EX2.
// 8 independent chains per thread, 32 warps per SM, one CTA per SM.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipes pipes.cu && ./pipes
#include <cstdio>
#include <cuda_runtime.h>

constexpr int kIters = 4096;
constexpr int kChains = 8;

template <int OP>
__global__ void __launch_bounds__(1024) tput(const float* in, double* out, long long* cyc) {
  float f[kChains];
  double d[kChains];
#pragma unroll
  for (int k = 0; k < kChains; ++k) {
    f[k] = in[threadIdx.x + k];
    d[k] = static_cast<double>(f[k]) + 1.0;
  }
  const double dm = static_cast<double>(in[1]) + 1.0000001, da = static_cast<double>(in[2]);
  const float fm = in[3] + 1.0000001f;
  __syncthreads();
  const long long t0 = clock64();
  for (int i = 0; i < kIters; ++i) {
#pragma unroll
    for (int k = 0; k < kChains; ++k) {
      if (OP == 0) f[k] = __fadd_rn(f[k], fm);
      if (OP == 1) d[k] = __dadd_rn(d[k], da);
      if (OP == 2) d[k] = __dmul_rn(d[k], dm);
      if (OP == 3) d[k] = __fma_rn(d[k], dm, da);
      if (OP == 4) {  // one F2F.F64.F32 + one FADD per link
        d[k] = static_cast<double>(f[k]);
        f[k] = __fadd_rn(f[k], __double2float_rn(0.0) + fm) ;
        asm volatile("" : "+d"(d[k]));
      }
      if (OP == 5) {  // one F2F.F32.F64 + one DADD per link
        f[k] = __double2float_rn(d[k]);
        d[k] = __dadd_rn(d[k], da);
        asm volatile("" : "+f"(f[k]));
      }
      if (OP == 6) f[k] = exp2f(f[k]);
      if (OP == 7) {  // round trip F2F.F64.F32 -> F2F.F32.F64 (dependent pair)
        d[k] = static_cast<double>(f[k]);
        asm volatile("" : "+d"(d[k]));
        f[k] = __double2float_rn(d[k]);
        asm volatile("" : "+f"(f[k]));
      }
    }
  }
  const long long t1 = clock64();
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < kChains; ++k) s += d[k] + static_cast<double>(f[k]);
  out[blockIdx.x * blockDim.x + threadIdx.x] = s;
  if (threadIdx.x == 0 && blockIdx.x == 0) cyc[0] = t1 - t0;
}

template <int OP>
void run(const char* name, int per_link, const float* in, double* out, long long* cyc) {
  int sms = 0;
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
  tput<OP><<<sms, 1024>>>(in, out, cyc);
  tput<OP><<<sms, 1024>>>(in, out, cyc);
  cudaDeviceSynchronize();
  long long c = 0;
  cudaMemcpy(&c, cyc, sizeof(c), cudaMemcpyDeviceToHost);
  const double warp_instr = 32.0 * kIters * kChains * per_link;  // per SM (32 warps)
  printf("%-34s %9lld cycles  %6.3f warp-instr/clk/SM  (%5.1f lanes/clk/SM, links of %d instr)\n", name, c,
         warp_instr / c, 32.0 * warp_instr / c, per_link);
}

int main() {
  float* in;
  double* out;
  long long* cyc;
  cudaMalloc(&in, 4096 * sizeof(float));
  cudaMemset(in, 0, 4096 * sizeof(float));
  cudaMalloc(&out, 148 * 1024 * sizeof(double) * 2);
  cudaMalloc(&cyc, 64);
  run<0>("FADD", 1, in, out, cyc);
  run<1>("DADD", 1, in, out, cyc);
  run<2>("DMUL", 1, in, out, cyc);
  run<3>("DFMA", 1, in, out, cyc);
  run<4>("F2F.F64.F32 + FADD", 2, in, out, cyc);
  run<5>("F2F.F32.F64 + DADD", 2, in, out, cyc);
  run<6>("MUFU.EX2 (+FMUL)", 1, in, out, cyc);
  run<7>("F2F.F64.F32 -> F2F.F32.F64", 2, in, out, cyc);
  printf("%s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
