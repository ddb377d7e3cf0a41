// K5a: Philox-keyed permutation of the environment axis (new capability).
// Spec = oracle/ref_math.py:philox_perm_ref (numpy); this file must match it bit for bit.
#include "perm.cuh"

namespace srl {

namespace {

__global__ void __launch_bounds__(256) philox_perm_kernel(uint32_t seed_lo, uint32_t seed_hi, uint32_t epoch0, int n_env,
                                                         int group, int bits, int32_t* __restrict__ out) {
  __shared__ PermKeys keys;
  if (threadIdx.x == 0) SRL_TL(0, blockIdx.y * gridDim.x + blockIdx.x, 0);
  pdl_launch_dependents();  // the scan that follows on the stream does not read the permutation: let it start now
  const uint32_t epoch = epoch0 + blockIdx.y;  // one grid row per epoch
  out += static_cast<size_t>(blockIdx.y) * n_env * group;
  if (threadIdx.x < 2) {  // two Philox blocks -> eight round keys
    Philox4 ctr;
    ctr.c[0] = threadIdx.x;
    ctr.c[1] = epoch;
    ctr.c[2] = 0x53524C50u;  // 'SRLP'
    ctr.c[3] = 0u;
    const Philox4 o = philox4x32_10(ctr, seed_lo, seed_hi);
    for (int i = 0; i < 4; ++i) keys.rk[threadIdx.x * 4 + i] = o.c[i];
  }
  if (threadIdx.x == 2) {
    const int lb = bits / 2, hb = bits - lb;
    keys.lb = lb;
    keys.lmask = (1u << lb) - 1u;
    keys.hmask = (1u << hb) - 1u;
  }
  __syncthreads();
  const PermKeys k = keys;
  for (int e = blockIdx.x * blockDim.x + threadIdx.x; e < n_env; e += gridDim.x * blockDim.x) {
    uint32_t x = feistel8(static_cast<uint32_t>(e), k);
    while (x >= static_cast<uint32_t>(n_env)) x = feistel8(x, k);  // cycle-walk back into [0, n_env)
    for (int a = 0; a < group; ++a) out[static_cast<size_t>(e) * group + a] = static_cast<int32_t>(x) * group + a;
  }
  if (threadIdx.x == 0) SRL_TL(0, blockIdx.y * gridDim.x + blockIdx.x, 1);
  // launched programmatically behind a scan (common.cuh): this grid only completes once the scan has
  if (threadIdx.x == 0) pdl_wait();
}

__global__ void philox_blocks_kernel(const uint32_t* __restrict__ counter, const uint32_t* __restrict__ key, int n,
                                     uint32_t* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  Philox4 c;
  for (int q = 0; q < 4; ++q) c.c[q] = counter[4 * i + q];
  const Philox4 o = philox4x32_10(c, key[2 * i], key[2 * i + 1]);
  for (int q = 0; q < 4; ++q) out[4 * i + q] = o.c[q];
}

}  // namespace
}  // namespace srl

SRL_TL_SETTER(srl_tl_set_perm)

extern "C" int srl_philox_perm(uint64_t seed, uint32_t epoch, int n_epochs, int n_env, int group, int32_t* out,
                               srl_stream_t stream) {
  using namespace srl;
  SRL_REQUIRE(n_env >= 0 && group >= 1 && n_epochs >= 0 && n_epochs <= 65535, SRL_ERR_INVALID_ARG,
              "srl_philox_perm: need n_env >= 0, group >= 1 and 0 <= n_epochs <= 65535");
  if (n_env == 0 || n_epochs == 0) return SRL_OK;
  SRL_REQUIRE(out != nullptr, SRL_ERR_INVALID_ARG, "srl_philox_perm: null output");
  SRL_REQUIRE(static_cast<long long>(n_env) * group < (1ll << 31), SRL_ERR_UNSUPPORTED,
              "srl_philox_perm: n_env * group must fit int32");
  const int bits = perm_bits(n_env);
  const int threads = 256;
  int grid = (n_env + threads - 1) / threads;
  const int cap = sm_count() * 8;
  if (grid > cap) grid = cap;
  const cudaStream_t st = static_cast<cudaStream_t>(stream);
  const bool behind_scan = pdl_take_scan(st);  // directly behind a scan on this stream: start beside it
  SRL_CUDA(launch_maybe_pdl(behind_scan, philox_perm_kernel, dim3(grid, n_epochs), dim3(threads), 0, st,
                            static_cast<uint32_t>(seed & 0xffffffffull), static_cast<uint32_t>(seed >> 32), epoch, n_env, group,
                            bits, out));
  if (!behind_scan) pdl_note_perm(st);  // a scan launched next on this stream may start beside this kernel
  return SRL_OK;
}

extern "C" int srl_philox4x32_10(const uint32_t* counter, const uint32_t* key, int n_blocks, uint32_t* out,
                                 srl_stream_t stream) {
  using namespace srl;
  SRL_REQUIRE(n_blocks >= 0, SRL_ERR_INVALID_ARG, "srl_philox4x32_10: negative block count");
  if (n_blocks == 0) return SRL_OK;
  SRL_REQUIRE(counter && key && out, SRL_ERR_INVALID_ARG, "srl_philox4x32_10: null pointer");
  philox_blocks_kernel<<<(n_blocks + 127) / 128, 128, 0, static_cast<cudaStream_t>(stream)>>>(counter, key, n_blocks,
                                                                                             out);
  SRL_CUDA(cudaGetLastError());
  return SRL_OK;
}
