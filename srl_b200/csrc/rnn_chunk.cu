// what a recurrent policy does to a minibatch's reset flags and hidden states before its RNN
// core, in ONE launch instead of ~10 ATen kernels and a blocking nonzero():
//   * the reference chunks every leaf (recursive_apply(sample, to_chunk), actor_critic_policy.py:348-350), takes the hidden
//     state at each chunk's first step and transposes it to layer-major (x[0].transpose(0, 1), :362-363);
//   * AutoResetRNN.forward (legacy/algorithm/modules/autoreset_rnn.py:42-60) then finds the rows in which ANY lane resets
//     ((masks[1:] == 0).any(dim=1).nonzero() -- a host synchronisation) to cut the sequence into segments, and multiplies the
//     hidden state by (1 - on_reset) at each segment's start.
// Here: reset_chunk = to_chunk(on_reset[:, env_idx], C); row_any[t'] = "some lane of chunk-row t' resets" (the segment
// boundaries: a Tc-byte vector the host reads once); hx0 = the chunk-start hidden states of the minibatch's lanes, layer-major,
// already multiplied by the first row's mask (the first segment's `hxs * masks[0]`).  Byte and float copies with one exact
// multiply by 0 or 1: bit-identical to the torch expressions (tests/test_gpu_family.py).
#include "common.cuh"

namespace srl {
namespace {

constexpr int kChunkThreads = 256;

// blocks [0, Tc): one chunk-row each (flags + the row's OR); blocks [Tc, gridDim.x): the hidden states, a warp per (column, layer)
__global__ void __launch_bounds__(kChunkThreads) rnn_chunk_prep_kernel(const uint8_t* __restrict__ on_reset,
                                                                      const float* __restrict__ hx,
                                                                      const int32_t* __restrict__ env_idx, int Tc, int B, int n,
                                                                      int C, int layers, int H,
                                                                      uint8_t* __restrict__ reset_chunk,
                                                                      uint8_t* __restrict__ row_any, float* __restrict__ hx0) {
  const int cols = C * n;
  if (static_cast<int>(blockIdx.x) < Tc) {
    const int t = blockIdx.x;
    int any = 0;
    for (int col = threadIdx.x; col < cols; col += kChunkThreads) {
      const int c = col / n, j = col - c * n;
      const uint8_t r = on_reset[(static_cast<size_t>(c) * Tc + t) * B + env_idx[j]];
      reset_chunk[static_cast<size_t>(t) * cols + col] = r;
      any |= r != 0 ? 1 : 0;
    }
    any = __syncthreads_or(any);
    if (threadIdx.x == 0) row_any[t] = any ? 1 : 0;
    return;
  }
  if (hx == nullptr) return;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarps = kChunkThreads / 32;
  const long long items = static_cast<long long>(cols) * layers;
  const long long stride = static_cast<long long>(gridDim.x - Tc) * kWarps;
  for (long long it = static_cast<long long>(blockIdx.x - Tc) * kWarps + warp; it < items; it += stride) {
    const int col = static_cast<int>(it / layers), l = static_cast<int>(it - static_cast<long long>(col) * layers);
    const int c = col / n, j = col - c * n;
    const size_t slot = static_cast<size_t>(c) * Tc * B + env_idx[j];  // time row c * Tc: the chunk's first step
    const float keep = on_reset[slot] != 0 ? 0.f : 1.f;                // masks[0] = 1 - on_reset[0]
    const float* s = hx + (slot * layers + l) * H;
    float* d = hx0 + (static_cast<size_t>(l) * cols + col) * H;
    for (int h = lane; h < H; h += 32) d[h] = __fmul_rn(s[h], keep);
  }
}

}  // namespace
}  // namespace srl

extern "C" int srl_rnn_chunk_prep(const uint8_t* on_reset, const float* hx, const int32_t* env_idx, int T, int B, int n,
                                  int num_chunks, int layers, int H, uint8_t* reset_chunk, uint8_t* row_any, float* hx0,
                                  srl_stream_t stream) {
  using namespace srl;
  SRL_REQUIRE(T >= 1 && B >= 1 && n >= 0 && num_chunks >= 1, SRL_ERR_INVALID_ARG,
              "srl_rnn_chunk_prep: need T >= 1, B >= 1, n >= 0, num_chunks >= 1 (got %d, %d, %d, %d)", T, B, n, num_chunks);
  SRL_REQUIRE(T % num_chunks == 0, SRL_ERR_INVALID_ARG,
              "srl_rnn_chunk_prep: the time dimension %d must be a multiple of num_chunks %d", T, num_chunks);
  SRL_REQUIRE(static_cast<long long>(n) * num_chunks <= 0x7fffffffll, SRL_ERR_INVALID_ARG,
              "srl_rnn_chunk_prep: n * num_chunks must fit int32");
  SRL_REQUIRE((hx == nullptr) == (hx0 == nullptr) && (hx == nullptr || (layers >= 1 && H >= 1)), SRL_ERR_INVALID_ARG,
              "srl_rnn_chunk_prep: hx and hx0 come together, with layers >= 1 and H >= 1");
  if (n == 0) return SRL_OK;
  SRL_REQUIRE(on_reset && env_idx && reset_chunk && row_any, SRL_ERR_INVALID_ARG, "srl_rnn_chunk_prep: null pointer");
  const int Tc = T / num_chunks;
  long long hx_blocks = 0;
  if (hx != nullptr) {
    const long long items = static_cast<long long>(n) * num_chunks * layers;
    hx_blocks = (items + kChunkThreads / 32 - 1) / (kChunkThreads / 32);
    const long long cap = static_cast<long long>(sm_count()) * 16;
    if (hx_blocks > cap) hx_blocks = cap;
  }
  rnn_chunk_prep_kernel<<<static_cast<unsigned>(Tc + hx_blocks), kChunkThreads, 0, static_cast<cudaStream_t>(stream)>>>(
      on_reset, hx, env_idx, Tc, B, n, num_chunks, layers, H, reset_chunk, row_any, hx0);
  SRL_CUDA(cudaGetLastError());
  return SRL_OK;
}
