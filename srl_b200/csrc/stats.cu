// Group statistics over K2's per-lane partial sums, and the PopArt / RunningMeanStd update (K3).
// See include/srl_b200.h: replaces the reductions + 3+3 single-element all-reduces of
// utils.py:54-61 (masked_normalization) and utils.py:113-130 (RunningMeanStd.update).
#include "common.cuh"
#include "xchg.cuh"

namespace srl {
namespace {

constexpr int kThreads = 256;
constexpr int kChunkLanes = 512;   // lanes one CTA sums: 2 per thread, all 14 loads of a thread in flight together
constexpr int kMaxChunks = 256;    // chunks per output row (longer rows use longer chunks)

__host__ __device__ inline int chunks_of(int lanes) {
  int c = (lanes + kChunkLanes - 1) / kChunkLanes;
  return c < 1 ? 1 : (c > kMaxChunks ? kMaxChunks : c);
}

// Workspace: [rows] tickets (uint32, padded to 256 B) then [rows][kMaxChunks][8] float64 chunk partials.
struct GroupWs {
  unsigned int* ticket;     // [rows] chunk tickets, then one "rows finished" counter (fused exchange only)
  double* partial;
};
__host__ __device__ inline size_t ticket_bytes(int rows) { return (static_cast<size_t>(rows + 1) * 4 + 255) / 256 * 256; }

// Fused exchange (srl_group_stats_xchg): the CTA that completes the LAST row of the table runs the peer-memory exchange
// right away -- the table goes out over NVLink from the kernel that produced it, one launch and one dependency edge
// less between K2 and the loss than group_stats -> exchange kernel -> loss.
struct FusedXchg {
  XchgView view;
  double* global_out;  // null = no exchange
  int n_rows;
};

// grid = (chunks, rows).  CTA (c, r) sums chunk c of output row r in a fixed order (lane-strided per thread, warp
// shuffle tree, warps in order); rows of one chunk are written directly, longer rows go through per-chunk partials
// that the last CTA to arrive (atomic ticket) adds in chunk order: deterministic, no float atomics.  The first
// version ran 8 CTAs per row whatever its length -- 68 us for the 65536-lane batch row of cfg5 (profiles/).
__global__ void __launch_bounds__(kThreads) group_stats_kernel(const double* __restrict__ lane_part, int N,
                                                               const int32_t* __restrict__ idx, int per_group,
                                                               int whole_first, double* __restrict__ out, GroupWs ws,
                                                               const FusedXchg fx) {
  pdl_launch_dependents();  // a loss kernel launched behind this one may become resident (it waits before reading)
  // output row blockIdx.y; with whole_first, row 0 is the identity group over all N lanes
  const bool whole = whole_first && blockIdx.y == 0;
  const int g = static_cast<int>(blockIdx.y) - (whole_first ? 1 : 0);
  const int per = whole ? N : per_group;
  if (whole) idx = nullptr;
  const int n_chunks = chunks_of(per);
  if (static_cast<int>(blockIdx.x) >= n_chunks) return;
  const int chunk = (per + n_chunks - 1) / n_chunks;
  const int j0 = blockIdx.x * chunk, j1 = min(per, j0 + chunk);

  double acc[SRL_LANE_PART];
#pragma unroll
  for (int k = 0; k < SRL_LANE_PART; ++k) acc[k] = 0.0;
  // two lanes per thread and iteration: both index loads, then all 14 value loads, are issued before the first use
  // (one exposed round trip for the indices, one for the values)
  for (int j = j0 + threadIdx.x; j < j1; j += 2 * kThreads) {
    const int ja = j, jb = j + kThreads;
    const bool has_b = jb < j1;
    const size_t base = static_cast<size_t>(g) * per;
    const int la = whole ? ja : (idx ? __ldg(idx + base + ja) : g * per + ja);
    const int lb = !has_b ? la : (whole ? jb : (idx ? __ldg(idx + base + jb) : g * per + jb));
    double va[SRL_LANE_PART - 1], vb[SRL_LANE_PART - 1];
#pragma unroll
    for (int k = 0; k < SRL_LANE_PART - 1; ++k) {
      va[k] = __ldg(lane_part + static_cast<size_t>(k) * N + la);
      vb[k] = __ldg(lane_part + static_cast<size_t>(k) * N + lb);
    }
#pragma unroll
    for (int k = 0; k < SRL_LANE_PART - 1; ++k) {
      acc[k] += va[k];
      if (has_b) acc[k] += vb[k];
    }
  }
  __shared__ double wsum[SRL_LANE_PART][kThreads / 32];
  __shared__ bool is_last, table_done;
  const int warp = threadIdx.x >> 5, ln = threadIdx.x & 31;
#pragma unroll
  for (int k = 0; k < SRL_LANE_PART; ++k) {
    const double s = warp_sum(acc[k]);
    if (ln == 0) wsum[k][warp] = s;
  }
  __syncthreads();
  double mine = 0.0;
  if (threadIdx.x < SRL_LANE_PART) {
    for (int w = 0; w < kThreads / 32; ++w) mine += wsum[threadIdx.x][w];
    if (n_chunks == 1) {
      out[static_cast<size_t>(blockIdx.y) * SRL_LANE_PART + threadIdx.x] = mine;
      if (fx.global_out) __threadfence();
    } else {
      ws.partial[(static_cast<size_t>(blockIdx.y) * kMaxChunks + blockIdx.x) * SRL_LANE_PART + threadIdx.x] = mine;
      __threadfence();
    }
  }
  if (n_chunks > 1) {
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(&ws.ticket[blockIdx.y], 1u) == static_cast<unsigned int>(n_chunks) - 1u;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    if (threadIdx.x < SRL_LANE_PART) {
      double s = 0.0;
      for (int c = 0; c < n_chunks; ++c)
        s += __ldcg(ws.partial + (static_cast<size_t>(blockIdx.y) * kMaxChunks + c) * SRL_LANE_PART + threadIdx.x);
      out[static_cast<size_t>(blockIdx.y) * SRL_LANE_PART + threadIdx.x] = s;
      if (fx.global_out) __threadfence();
    }
    if (threadIdx.x == 0) ws.ticket[blockIdx.y] = 0u;  // ready for the next launch
  }
  if (fx.global_out == nullptr) return;
  // this CTA completed row blockIdx.y; the one that completes the table's last row exchanges it with the peers
  __syncthreads();
  if (threadIdx.x == 0) table_done = atomicAdd(&ws.ticket[gridDim.y], 1u) == static_cast<unsigned int>(fx.n_rows) - 1u;
  __syncthreads();
  if (!table_done) return;
  __threadfence();
  if (threadIdx.x == 0) ws.ticket[gridDim.y] = 0u;
  xchg_exchange(fx.view, out, fx.global_out, fx.n_rows * SRL_LANE_PART);
}

// Per-lane partial sums from adv / ret that already exist (a re-served sample whose host copy carries them,
// mappo.py:224-225 with recompute_adv_on_reuse=False): the same eight rows K2 writes, one thread per lane,
// rows walked newest-first exactly like the scan so both producers sum in the same order.
__global__ void __launch_bounds__(256) lane_stats_kernel(const float* __restrict__ adv, const float* __restrict__ ret,
                                                         const uint8_t* __restrict__ done,
                                                         const uint8_t* __restrict__ truncated,
                                                         const uint8_t* __restrict__ on_reset, int N, int row_lo,
                                                         int row_hi, double* __restrict__ lane_part) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= N) return;
  double s0 = 0, s1 = 0, s2 = 0, s3 = 0, s4 = 0, s5 = 0, s6 = 0;
  for (int t = row_hi - 1; t >= row_lo; --t) {
    const size_t g = static_cast<size_t>(t) * N + col;
    const double mk = on_reset[g + N] ? 0.0 : 1.0;  // 1 - on_reset[t+1]   mappo.py:260-261
    const double x = __dmul_rn(static_cast<double>(adv[g]), mk);
    const double y = __dmul_rn(static_cast<double>(ret[g]), mk);
    s0 += mk;
    s1 += x;
    s2 = __dadd_rn(s2, __dmul_rn(x, x));
    s3 += y;
    s4 = __dadd_rn(s4, __dmul_rn(y, y));
    s5 += done[g] ? 1.0 : 0.0;
    s6 += truncated[g] ? 1.0 : 0.0;
  }
  double* o = lane_part + col;
  o[0] = s0;
  o[static_cast<size_t>(1) * N] = s1;
  o[static_cast<size_t>(2) * N] = s2;
  o[static_cast<size_t>(3) * N] = s3;
  o[static_cast<size_t>(4) * N] = s4;
  o[static_cast<size_t>(5) * N] = s5;
  o[static_cast<size_t>(6) * N] = s6;
  o[static_cast<size_t>(7) * N] = 0.0;
}

__global__ void popart_update_kernel(const double* __restrict__ bs, double* __restrict__ state, double beta,
                                     double eps, double* __restrict__ ms) {
  pdl_launch_dependents();
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double mean = state[0], mean_sq = state[1], debias = state[2];
  // statistics before the update (PopArtValueHead.update reads them for the head rescale, popart.py:43)
  {
    const double d = fmax(debias, eps);
    const double mu = mean / d;
    ms[2] = mu;
    ms[3] = sqrt(fmax(__dsub_rn(mean_sq / d, __dmul_rn(mu, mu)), 1e-2));
  }
  const double cnt = bs[0];
  const double bm = bs[3] / cnt, bsq = bs[4] / cnt;  // utils.py:125-126
  const double omb = 1.0 - beta;
  mean = __dadd_rn(__dmul_rn(beta, mean), __dmul_rn(bm, omb));         // utils.py:128
  mean_sq = __dadd_rn(__dmul_rn(beta, mean_sq), __dmul_rn(bsq, omb));  // utils.py:129
  debias = __dsub_rn(__dadd_rn(__dmul_rn(beta, debias), 1.0), beta);   // utils.py:130
  state[0] = mean;
  state[1] = mean_sq;
  state[2] = debias;
  state[3] += 1.0;
  const double d = fmax(debias, eps);  // utils.py:134-137
  const double mu = mean / d;
  ms[0] = mu;
  ms[1] = sqrt(fmax(__dsub_rn(mean_sq / d, __dmul_rn(mu, mu)), 1e-2));
}

}  // namespace
}  // namespace srl


namespace srl {
namespace {
int launch_group_stats(const char* fn, const double* lane_part, int N, const int32_t* idx, int G, int per, int whole_first,
                       double* out, void* workspace, size_t workspace_bytes, const srl_xchg* x, double* global_out,
                       srl_stream_t stream) {
  SRL_REQUIRE(lane_part && out, SRL_ERR_INVALID_ARG, "%s: null pointer", fn);
  SRL_REQUIRE(N >= 1 && G >= (whole_first ? 0 : 1) && per >= 1, SRL_ERR_INVALID_ARG, "%s: need N, G, per >= 1", fn);
  SRL_REQUIRE(idx != nullptr || static_cast<long long>(G) * per <= N, SRL_ERR_INVALID_ARG, "%s: G*per=%lld exceeds N=%d", fn,
              static_cast<long long>(G) * per, N);
  SRL_REQUIRE(G <= 65534, SRL_ERR_UNSUPPORTED, "%s: at most 65534 groups", fn);
  const int rows = G + (whole_first ? 1 : 0);
  int max_chunks = G > 0 ? chunks_of(per) : 1;
  if (whole_first && chunks_of(N) > max_chunks) max_chunks = chunks_of(N);
  GroupWs ws{nullptr, nullptr};
  FusedXchg fx;
  fx.global_out = nullptr;
  fx.n_rows = rows;
  if (x != nullptr) {
    const XchgView* v = xchg_view(x);
    SRL_REQUIRE(v != nullptr && global_out != nullptr, SRL_ERR_INVALID_ARG, "%s: exchange not connected or null output", fn);
    SRL_REQUIRE(rows * SRL_LANE_PART <= v->cap, SRL_ERR_INVALID_ARG, "%s: table of %d doubles exceeds the exchange capacity %d",
                fn, rows * SRL_LANE_PART, v->cap);
    fx.view = *v;
    fx.global_out = global_out;
  }
  if (max_chunks > 1 || x != nullptr) {
    SRL_REQUIRE(workspace != nullptr && workspace_bytes >= srl_group_stats_workspace_bytes(G, whole_first) &&
                    aligned(workspace, 8),
                SRL_ERR_INVALID_ARG, "%s: rows longer than %d lanes (or a fused exchange) need a workspace of %zu bytes", fn,
                kChunkLanes, srl_group_stats_workspace_bytes(G, whole_first));
    ws.ticket = static_cast<unsigned int*>(workspace);
    ws.partial = reinterpret_cast<double*>(static_cast<char*>(workspace) + ticket_bytes(rows));
  }
  pdl_forget();
  group_stats_kernel<<<dim3(max_chunks, rows), kThreads, 0, static_cast<cudaStream_t>(stream)>>>(lane_part, N, idx, per,
                                                                                                 whole_first, out, ws, fx);
  SRL_CUDA(cudaGetLastError());
  return SRL_OK;
}
}  // namespace
}  // namespace srl

extern "C" size_t srl_group_stats_workspace_bytes(int G, int whole_first) {
  const int rows = G + (whole_first ? 1 : 0);
  return srl::ticket_bytes(rows) + static_cast<size_t>(rows) * srl::kMaxChunks * SRL_LANE_PART * sizeof(double);
}

extern "C" int srl_group_stats(const double* lane_part, int N, const int32_t* idx, int G, int per, int whole_first,
                               double* out, void* workspace, size_t workspace_bytes, srl_stream_t stream) {
  return srl::launch_group_stats("srl_group_stats", lane_part, N, idx, G, per, whole_first, out, workspace, workspace_bytes,
                                 nullptr, nullptr, stream);
}

extern "C" int srl_group_stats_xchg(const double* lane_part, int N, const int32_t* idx, int G, int per, int whole_first,
                                    double* local_out, double* global_out, void* workspace, size_t workspace_bytes,
                                    srl_xchg* x, srl_stream_t stream) {
  using namespace srl;
  SRL_REQUIRE(x != nullptr, SRL_ERR_INVALID_ARG, "srl_group_stats_xchg: null exchange handle");
  return launch_group_stats("srl_group_stats_xchg", lane_part, N, idx, G, per, whole_first, local_out, workspace,
                            workspace_bytes, x, global_out, stream);
}

extern "C" int srl_popart_update(const double* batch_stats, double* state, double beta, double eps,
                                 double* mean_std_out, srl_stream_t stream) {
  using namespace srl;
  SRL_REQUIRE(batch_stats && state && mean_std_out, SRL_ERR_INVALID_ARG, "srl_popart_update: null pointer");
  popart_update_kernel<<<1, 32, 0, static_cast<cudaStream_t>(stream)>>>(batch_stats, state, beta, eps, mean_std_out);
  SRL_CUDA(cudaGetLastError());
  return SRL_OK;
}

extern "C" int srl_lane_stats(const float* adv, const float* ret, const uint8_t* done, const uint8_t* truncated,
                              const uint8_t* on_reset, int L, int N, int row_lo, int row_hi, double* lane_part,
                              srl_stream_t stream) {
  using namespace srl;
  SRL_REQUIRE(adv && ret && done && truncated && on_reset && lane_part, SRL_ERR_INVALID_ARG, "srl_lane_stats: null pointer");
  SRL_REQUIRE(L >= 2 && N >= 1 && row_lo >= 0 && row_lo <= row_hi && row_hi <= L - 1, SRL_ERR_INVALID_ARG,
              "srl_lane_stats: need L >= 2, N >= 1 and 0 <= row_lo <= row_hi <= L-1 (got L=%d N=%d rows [%d, %d))", L, N,
              row_lo, row_hi);
  lane_stats_kernel<<<(N + 255) / 256, 256, 0, static_cast<cudaStream_t>(stream)>>>(adv, ret, done, truncated, on_reset,
                                                                                      N, row_lo, row_hi, lane_part);
  SRL_CUDA(cudaGetLastError());
  return SRL_OK;
}
