// n-step return (sibling of the GAE scan on the same batch layout).
// Replaces modules.n_step_return (legacy/algorithm/modules/n_step_return.py:11-50), used by the DQN / QMIX trainers:
//   ret = 0, disc = 1
//   for i in 0..n-1:  ret  += reward[t+i] * disc
//                     ret  += disc * gamma * nex_truncated[t+i] * nex_value[t+i]      (bootstrap at a truncation)
//                     disc *= gamma * (1 - nex_done[t+i]) * (1 - nex_truncated[t+i])
//   out[t] = float(ret + disc * nex_value[t+n-1])
// in float64 with the reference's operation order (explicit roundings, no FMA contraction) -> bit-identical.
// One thread per (t, lane); a warp reads 32 adjacent lanes of row t+i (coalesced), and the n rows of a window
// overlap with the next thread row's window, so every input byte comes from HBM once and from L1/L2 n-1 times.
#include "common.cuh"

namespace srl {
namespace {

__global__ void __launch_bounds__(256) n_step_return_kernel(const float* __restrict__ reward,
                                                            const float* __restrict__ nex_value,
                                                            const uint8_t* __restrict__ nex_done,
                                                            const uint8_t* __restrict__ nex_truncated, int n, int T, int N,
                                                            double gamma, float* __restrict__ out) {
  const long long W = static_cast<long long>(T) * N;
  for (long long e = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; e < W;
       e += static_cast<long long>(gridDim.x) * blockDim.x) {
    double ret = 0.0, disc = 1.0;
    long long g = e;  // element (t + i, lane)
    double last_value = 0.0;
    for (int i = 0; i < n; ++i, g += N) {
      const double r = static_cast<double>(__ldg(reward + g));
      const double v = static_cast<double>(__ldg(nex_value + g));
      const double dn = __ldg(nex_done + g) ? 1.0 : 0.0;
      const double tr = __ldg(nex_truncated + g) ? 1.0 : 0.0;
      ret = __dadd_rn(ret, __dmul_rn(r, disc));
      ret = __dadd_rn(ret, __dmul_rn(__dmul_rn(__dmul_rn(disc, gamma), tr), v));
      disc = __dmul_rn(disc, __dmul_rn(__dmul_rn(gamma, __dsub_rn(1.0, dn)), __dsub_rn(1.0, tr)));
      last_value = v;
    }
    out[e] = static_cast<float>(__dadd_rn(ret, __dmul_rn(disc, last_value)));
  }
}

}  // namespace
}  // namespace srl

extern "C" int srl_n_step_return(const float* reward, const float* nex_value, const uint8_t* nex_done,
                                 const uint8_t* nex_truncated, int n, int rows, int N, double gamma, float* out,
                                 srl_stream_t stream) {
  using namespace srl;
  SRL_REQUIRE(reward && nex_value && nex_done && nex_truncated && out, SRL_ERR_INVALID_ARG,
              "srl_n_step_return: null pointer");
  SRL_REQUIRE(n >= 1 && N >= 1 && rows >= n, SRL_ERR_INVALID_ARG,
              "srl_n_step_return: need n >= 1, N >= 1 and rows >= n (got n=%d rows=%d N=%d)", n, rows, N);
  const int T = rows - n + 1;
  const long long W = static_cast<long long>(T) * N;
  long long grid = (W + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (grid > cap) grid = cap;
  n_step_return_kernel<<<static_cast<int>(grid), 256, 0, static_cast<cudaStream_t>(stream)>>>(
      reward, nex_value, nex_done, nex_truncated, n, T, N, gamma, out);
  SRL_CUDA(cudaGetLastError());
  return SRL_OK;
}
