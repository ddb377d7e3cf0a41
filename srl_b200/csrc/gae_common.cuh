// Shared between the two GAE scan implementations (gae_scan.cu: general tile kernel; gae_scan_tma.cu: TMA pipeline).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

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
  float* pack;  // [L, N, 4] or null
  int L, N, row_lo, row_hi;
  double gamma, gamma_lmbda, rho, c;
};

bool gae_tma_eligible(const GaeParams& p);
int launch_gae_tma(const GaeParams& p, cudaStream_t st);
bool gae_ws_eligible(const GaeParams& p);
int launch_gae_ws(const GaeParams& p, cudaStream_t st);

}  // namespace srl
