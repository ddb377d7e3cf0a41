// Shared between the two GAE scan implementations (gae_scan.cu: general tile kernel; gae_scan_tma.cu: TMA pipeline).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "perm.cuh"

namespace srl {

struct GaeParams {
  const float* reward;
  const float* value;
  const uint8_t* done;
  const uint8_t* truncated;
  const uint8_t* on_reset;
  const float* vt_new_logp;
  const float* vt_old_logp;
  const double* popart;  // {mean, std} or null
  const float* old_logp;  // [L, N] or null (only for the pack)
  float* adv;
  float* ret;
  double* lane_part;
  double* lane_aos;  // [N][4] f64 {sum mask, sum adv*mask, sum (adv*mask)^2, 0} or null: the loss kernel's gather form
  float* pack;  // pair-interleaved [ceil(L/2)][N][2] float4 or null (include/srl_b200.h)
  int L, N, row_lo, row_hi;
  double gamma, gamma_lmbda, rho, c;
  PermJob perm;  // srl_gae_scan_perm: a permutation computed on the side (gae_scan_ws.cu only); out == nullptr: none
};

// Position (in float4 items) of transition (t, lane) in the loss pack: the two rows of a row pair sit next to each other,
// so a permuted minibatch fetches 32 contiguous, 32-byte aligned bytes per lane and row pair.
__host__ __device__ inline size_t pack_index(int t, int N, int lane) {
  return (static_cast<size_t>(t >> 1) * N + lane) * 2 + (t & 1);
}

bool gae_tma_eligible(const GaeParams& p);
int launch_gae_tma(const GaeParams& p, cudaStream_t st);
bool gae_ws_eligible(const GaeParams& p);
int launch_gae_ws(const GaeParams& p, cudaStream_t st);

}  // namespace srl
