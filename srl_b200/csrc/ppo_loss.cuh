// K4 device code shared by the loss translation units (ppo_loss.cu: C-ABI, logits variant, finaliser;
// ppo_loss_dense.cu / ppo_loss_gather.cu / ppo_loss_pack.cu: the kernel instantiated per sample-side form, one file
// each so nvcc builds them in parallel).  See ppo_loss.cu for the design notes.
#pragma once
#include <float.h>

#include "common.cuh"

namespace srl {
namespace loss {

#ifndef SRL_LOSS_UNROLL
#define SRL_LOSS_UNROLL 1
#endif
#ifndef SRL_LOSS_NORM_FP32
#define SRL_LOSS_NORM_FP32 0  // 1: advantage normalisation in float32 (a measurement variant; default float64 as the reference)
#endif
#ifndef SRL_LOSS_MIN_BLOCKS
#define SRL_LOSS_MIN_BLOCKS 3  // CTAs of 256 threads per SM the register allocator must leave room for
#endif
constexpr int kLossUnroll = SRL_LOSS_UNROLL;  // rows of a tile whose loads are in flight together
constexpr int kMaxGrid = 2048;                // partial rows per workspace slot
constexpr int kNumSums = 8;
constexpr size_t kPartialsOffset = 64;
constexpr int kMaxBatch = SRL_MAX_LOSS_BATCH;
constexpr int kFlushRows = 8;  // rows whose masked terms are summed in fp32 before they enter the float64 sums

// First 64 bytes of a workspace slot; the partial rows [n_rows][8] f64 follow.
struct SlotHeader {
  unsigned int ticket;  // immediate mode only; zero between launches
  unsigned int n_rows;  // CTAs that wrote a partial row
  double mask_sum;      // this rank's sum(mask) for the minibatch
  double wv, we;        // loss weights, so the finaliser needs nothing but the slot
  double pad[4];
};
static_assert(sizeof(SlotHeader) == kPartialsOffset, "slot header must stay 64 bytes");

struct LossHyperDev {
  float clip_lo, clip_hi;  // (float)(1 -/+ eps_clip)
  float veps;
  float c_clip;
  float wv, we;
  float vl_param;
  double adv_eps;
  int value_loss, clip_value, dual_clip, normalize_old_value;
};

// Per-minibatch pointers.
struct Problem {
  const float* new_logp;
  const float* v_pred;
  const float* entropy;
  const int32_t* lane_idx;
  const double* norm_stats;
  const double* local_stats;
  float* g_logp;
  float* g_value;
  float* g_entropy;
  double* out;
  float* out_f32;
  SlotHeader* slot;
};

// What every problem of a launch shares.
struct LossShared {
  const float* old_logp;
  const float* old_value;
  const float* ret;
  const float* adv;
  const uint8_t* reset_next;
  const float4* pack;
  const double* popart;
  const double* lane_part;  // [SRL_LANE_PART, lane_part_n] from K2, or null; see self_uniforms()
  int lane_part_n;
  long long ld_pol, ld_grad, ld_smp;
  int T, n;
  int rows_per_tile, col_tiles, n_tiles;
  int smp_vec_ok;  // sample-side leaves allow 128-bit loads at 4-aligned lanes (alignment + row stride)
  int prefetch_rows;  // policy-side rows a CTA asks into L2 before it waits for the kernel ahead of it (0: none)
  LossHyperDev h;
};

struct LossBatch {
  LossShared s;
  Problem prob[kMaxBatch];
};

// Sample-side forms
constexpr int kDense = 0;   // separate leaves, lanes in policy order, 128-bit loads
constexpr int kGather = 1;  // separate leaves through lane_idx (or any alignment)
constexpr int kPack = 2;    // K2's float4 pack through lane_idx (or in order)

// Hyper-parameter configuration of the element math.  The hot kernels are instantiated per configuration
// (StaticCfg): ncu showed ~25 % of the executed instructions of the first batched kernel were uniform branches,
// constant-bank reloads and reconvergence barriers around `if (h.clip_value)`-style tests inside the element loop
// (profiles/r1c_notes.md).  RuntimeCfg keeps one general instantiation for the rare combinations.
template <int VL, bool CLIP, bool DUAL, bool POPART>
struct StaticCfg {
  __device__ static __forceinline__ int vl(const LossHyperDev&) { return VL; }
  __device__ static __forceinline__ bool clip(const LossHyperDev&) { return CLIP; }
  __device__ static __forceinline__ bool dual(const LossHyperDev&) { return DUAL; }
  __device__ static __forceinline__ bool popart(bool) { return POPART; }
};
struct RuntimeCfg {
  __device__ static __forceinline__ int vl(const LossHyperDev& h) { return h.value_loss; }
  __device__ static __forceinline__ bool clip(const LossHyperDev& h) { return h.clip_value != 0; }
  __device__ static __forceinline__ bool dual(const LossHyperDev& h) { return h.dual_clip != 0; }
  __device__ static __forceinline__ bool popart(bool have) { return have; }
};

// torch.nn.{MSELoss,HuberLoss,SmoothL1Loss}(reduction='none') value and derivative wrt the input.
// Inside a kind both branches are evaluated and selected (the quadratic / linear choice is per element, and a
// divergent branch costs more than the three spare flops).
__device__ __forceinline__ void pointwise_loss(int kind, float prm, float d, float& l, float& dl) {
  if (kind == SRL_VL_MSE) {
    l = d * d;
    dl = 2.f * d;
  } else if (kind == SRL_VL_HUBER) {
    const float z = fabsf(d);
    const bool quad = z < prm;
    const float lq = 0.5f * z * z, ll = prm * (z - 0.5f * prm);
    l = quad ? lq : ll;
    dl = quad ? d : copysignf(prm, d);  // d != 0 here (|d| >= prm > 0), so this is prm * sign(d)
  } else {
    const float z = fabsf(d);
    const bool quad = z < prm;
    const float lq = 0.5f * z * z / prm, ll = z - 0.5f * prm;
    l = quad ? lq : ll;
    dl = quad ? d / prm : copysignf(1.f, d);
  }
}

struct Uniforms {
  double mean, denom, rdenom;   // advantage normalisation: (x - mean) / denom, rdenom = 1 / denom
  double pa_mu, pa_sd, pa_rsd;  // popart: (x - mu) / sd
  float inv_m;                  // 1 / local sum(mask)
  bool popart;
};

__device__ __forceinline__ Uniforms load_uniforms(const double* norm_stats, const double* local_stats,
                                                  const double* popart, double adv_eps, double& mask_sum) {
  Uniforms u;
  const double cnt = __ldg(norm_stats), s1 = __ldg(norm_stats + 1), s2 = __ldg(norm_stats + 2);
  mask_sum = __ldg(local_stats);
  u.popart = popart != nullptr;
  u.pa_mu = u.popart ? __ldg(popart) : 0.0;
  u.pa_sd = u.popart ? __ldg(popart + 1) : 1.0;
  u.pa_rsd = 1.0 / u.pa_sd;
  u.mean = s1 / cnt;
  const double var = s2 / cnt - u.mean * u.mean;  // biased variance, utils.py:62-64
  u.denom = sqrt(var) + adv_eps;                  // eps outside the sqrt, utils.py:67
  u.rdenom = 1.0 / u.denom;
  u.inv_m = 1.f / static_cast<float>(mask_sum);
  return u;
}

__device__ __forceinline__ Uniforms uniforms_from(double cnt, double s1, double s2, double m_local, const double* popart,
                                                  double adv_eps) {
  Uniforms u;
  u.popart = popart != nullptr;
  u.pa_mu = u.popart ? __ldg(popart) : 0.0;
  u.pa_sd = u.popart ? __ldg(popart + 1) : 1.0;
  u.pa_rsd = 1.0 / u.pa_sd;
  u.mean = s1 / cnt;
  const double var = s2 / cnt - u.mean * u.mean;
  u.denom = sqrt(var) + adv_eps;
  u.rdenom = 1.0 / u.denom;
  u.inv_m = 1.f / static_cast<float>(m_local);
  return u;
}

// Self-computed statistics (one GPU, no PopArt, the CTA's column tile spans the whole minibatch): the CTA adds K2's
// per-lane sums over ITS minibatch's lanes -- the indices are in registers already -- instead of waiting for a
// srl_group_stats launch between K2 and K4.  Every CTA of a minibatch adds the same values in the same order (lanes of
// a thread, warp shuffle tree, warps in order), so they all normalise with bit-identical statistics.
template <int LANES>
__device__ __forceinline__ Uniforms self_uniforms(const double* __restrict__ lane_part, int N, const int (&c)[LANES],
                                                  bool active, double adv_eps, double& mask_sum) {
  __shared__ double s_part[3][8];
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  if (active) {
    double v0[LANES], v1[LANES], v2[LANES];
#pragma unroll
    for (int q = 0; q < LANES; ++q) {
      v0[q] = __ldg(lane_part + c[q]);
      v1[q] = __ldg(lane_part + static_cast<size_t>(N) + c[q]);
      v2[q] = __ldg(lane_part + 2 * static_cast<size_t>(N) + c[q]);
    }
#pragma unroll
    for (int q = 0; q < LANES; ++q) {
      a0 += v0[q];
      a1 += v1[q];
      a2 += v2[q];
    }
  }
  a0 = warp_sum(a0);
  a1 = warp_sum(a1);
  a2 = warp_sum(a2);
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    s_part[0][warp] = a0;
    s_part[1][warp] = a1;
    s_part[2][warp] = a2;
  }
  __syncthreads();
  double cnt = 0.0, s1 = 0.0, s2 = 0.0;
  for (int w = 0; w < nwarps; ++w) {
    cnt += s_part[0][w];
    s1 += s_part[1][w];
    s2 += s_part[2][w];
  }
  mask_sum = cnt;
  return uniforms_from(cnt, s1, s2, cnt, nullptr, adv_eps);
}

// a / d in float64 given rd = 1 / d (correctly rounded): product, exact residual, one correction.  This is the
// tail of the IEEE division sequence without its reciprocal refinement and special-case handling; d is a
// positive finite scale here (sqrt(var) + eps, or PopArt's sigma >= 0.1).
__device__ __forceinline__ double div_by(double a, double d, double rd) {
  const double q = __dmul_rn(a, rd);
  const double r = __fma_rn(-d, q, a);
  return __fma_rn(r, rd, q);
}

__device__ __forceinline__ float popart_normalize(float x, const Uniforms& u) {
  // RunningMeanStd.normalize: ((x.double() - mean) / std).clip(-5, 5).float()   utils.py:139-144
  // (clipping after the cast gives the same float: +-5 are exact and rounding is monotonic)
  const float z = static_cast<float>(div_by(__dsub_rn(static_cast<double>(x), u.pa_mu), u.pa_sd, u.pa_rsd));
  return fminf(fmaxf(z, -5.f), 5.f);
}

// fp32 partial sums of a few rows (kFlushRows x 4 lanes); folded into the thread's float64 accumulators in blocks.
struct RowSums {
  float pl = 0.f, vl = 0.f, en = 0.f, adv = 0.f, ratio = 0.f, vt = 0.f, ret = 0.f;
  int clip = 0;
};

struct Acc {
  double pl = 0, vl = 0, en = 0, adv = 0, ratio = 0, vt = 0, ret = 0;
  int clip = 0;
  __device__ __forceinline__ void add(const RowSums& r) {
    pl += static_cast<double>(r.pl);
    vl += static_cast<double>(r.vl);
    en += static_cast<double>(r.en);
    adv += static_cast<double>(r.adv);
    ratio += static_cast<double>(r.ratio);
    vt += static_cast<double>(r.vt);
    ret += static_cast<double>(r.ret);
    clip += r.clip;
  }
};

// One transition.  Masked transitions (valid == false) produce zero gradients and enter no sum: in the reference
// every term is multiplied by mask before it is summed (mappo.py:184,197,199) and the stats go through
// masked_select (mappo.py:206-216).
template <class CFG>
__device__ __forceinline__ void element(const LossHyperDev& h, const Uniforms& u, float nl, float vp, float en,
                                        float ol, float ov, float rt, float ad, bool valid, float& g_lp,
                                        float& g_v, float& g_en, RowSums& rs) {
  const float mk = valid ? 1.f : 0.f;
  const float scale = valid ? u.inv_m : 0.f;  // d(masked mean)/d(element) = mask / M
  ad = valid ? ad : 0.f;                      // the pack marks masked transitions with a NaN advantage

  // ---- critic: mappo.py:172-184, utils.py:228-239 ---------------------------------------------
  const float vt = CFG::popart(u.popart) ? popart_normalize(rt, u) : rt;
  float l, dl;
  pointwise_loss(CFG::vl(h), h.vl_param, vp - vt, l, dl);
  float vl = l, gv = dl;
  if (CFG::clip(h)) {
    const float ovn = (CFG::popart(u.popart) && h.normalize_old_value) ? popart_normalize(ov, u) : ov;
    const float dv = vp - ovn;
    const float vc = ovn + fminf(fmaxf(dv, -h.veps), h.veps);
    const bool in = fabsf(dv) <= h.veps;  // clamp passes grad on the closed interval
    float l2, dl2;
    pointwise_loss(CFG::vl(h), h.vl_param, vc - vt, l2, dl2);
    dl2 = in ? dl2 : 0.f;
    vl = fmaxf(l, l2);
    gv = l > l2 ? dl : (l < l2 ? dl2 : 0.5f * (dl + dl2));  // torch.max splits ties evenly
  }
  g_v = h.wv * scale * gv;

  // ---- actor: mappo.py:157-158,186-197 ---------------------------------------------------------
  const float ratio = expf(nl - ol);
  // masked_normalization (utils.py:38-67) in float64, cast to float at the end; masked entries are centred
  // zeros there (x = adv * mask) and here
#if SRL_LOSS_NORM_FP32
  // build variant for profiles/ (profiles/r1d_notes.md, candidate 2): ~1e-7 relative instead of the reference's float64
  const float nadv = (ad - static_cast<float>(u.mean)) * static_cast<float>(u.rdenom);
#else
  const float nadv = static_cast<float>(div_by(__dsub_rn(static_cast<double>(ad), u.mean), u.denom, u.rdenom));
#endif
  const float s1 = ratio * nadv;
  const float s2 = fminf(fmaxf(ratio, h.clip_lo), h.clip_hi) * nadv;
  const bool in_clip = ratio >= h.clip_lo && ratio <= h.clip_hi;
  // d min(s1, s2) / d ratio / nadv: 1 where s1 is the smaller, the clamp's pass-through where s2 is, the mean of
  // the two on ties (torch.min splits ties evenly)
  const float pass = in_clip ? 1.f : 0.f;
  const float g12 = s1 < s2 ? 1.f : (s1 > s2 ? pass : 0.5f + 0.5f * pass);
  float obj = fminf(s1, s2);
  float gsum = g12 * nadv * ratio;
  if (CFG::dual(h)) {
    const float s3 = -h.c_clip * fabsf(nadv);  // -sign(nadv) * c * nadv
    gsum = obj > s3 ? gsum : (obj < s3 ? 0.f : 0.5f * gsum);
    obj = fmaxf(obj, s3);
  }
  g_lp = -scale * gsum;
  g_en = -h.we * scale;  // entropy_loss = -sum(entropy * mask) / M   mappo.py:199

  // masked sums: multiply by the mask as the reference does (mappo.py:184,197,199).  Like there, a non-finite
  // term of a masked transition would poison the sum; the inputs of masked transitions are finite in practice
  // (dead agents carry new_logp = -inf, i.e. ratio = 0).
  rs.pl -= mk * obj;
  rs.vl += mk * vl;
  rs.en += mk * en;
  rs.adv += mk * ad;
  rs.ratio += mk * ratio;
  rs.vt += mk * vt;
  rs.ret += mk * rt;
  rs.clip += (valid && s2 < s1) ? 1 : 0;
}

// Folds n_rows partial rows (fixed order: lane-strided, then the warp-shuffle tree) and writes the results.
// Called by one CTA; sred is [kNumSums][8] shared scratch.
__device__ __forceinline__ void fold_rows_and_write(const double* __restrict__ partials, int n_rows, double M, double wv,
                                                    double we, double (*sred)[8], double* __restrict__ o,
                                                    float* __restrict__ o32) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  for (int k = warp; k < kNumSums; k += nwarps) {
    double s = 0.0;
    for (int b = lane; b < n_rows; b += 32) s += __ldcg(partials + static_cast<size_t>(b) * kNumSums + k);
    s = warp_sum(s);
    if (lane == 0) sred[k][0] = s;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const double pl = sred[0][0] / M, vl = sred[1][0] / M, el = -sred[2][0] / M;
    const double loss = pl + wv * vl + we * el;
    o[SRL_OUT_LOSS] = loss;
    o[SRL_OUT_POLICY_LOSS] = pl;
    o[SRL_OUT_VALUE_LOSS] = vl;
    o[SRL_OUT_ENTROPY_LOSS] = el;
    o[SRL_OUT_ADVANTAGE] = sred[3][0] / M;
    o[SRL_OUT_IMPORTANCE_WEIGHT] = sred[4][0] / M;
    o[SRL_OUT_CLIP_RATIO] = sred[5][0] / M;
    o[SRL_OUT_VALUE_TARGETS] = sred[6][0] / M;
    o[SRL_OUT_DENORM_VALUE] = sred[7][0] / M;
    o[SRL_OUT_MASK_SUM] = M;
    for (int k = SRL_OUT_MASK_SUM + 1; k < SRL_LOSS_OUT_LEN; ++k) o[k] = 0.0;
    if (o32) {
      o32[0] = static_cast<float>(loss);
      o32[1] = static_cast<float>(pl);
      o32[2] = static_cast<float>(vl);
      o32[3] = static_cast<float>(el);
    }
  }
}

// Block reduction of the 8 masked sums -> this CTA's partial row; then either done (deferred) or ticket.
// `row` / `n_rows`: this CTA's row and the number of CTAs working on the same problem.
__device__ __forceinline__ void reduce_and_finalize(const Problem& pr, const LossHyperDev& h, const Acc& acc,
                                                    double mask_sum, int row, int n_rows) {
  __shared__ double sred[kNumSums][8];
  __shared__ bool is_last;
  double* partials = reinterpret_cast<double*>(reinterpret_cast<char*>(pr.slot) + kPartialsOffset);
  double v[kNumSums] = {acc.pl, acc.vl, acc.en, acc.adv, acc.ratio, static_cast<double>(acc.clip), acc.vt, acc.ret};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31, nwarps = blockDim.x >> 5;
  // butterfly level by level over all 8 sums: 8 independent shuffle+add chains per level hide each other's latency
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double w[kNumSums];
#pragma unroll
    for (int k = 0; k < kNumSums; ++k) w[k] = __shfl_xor_sync(0xffffffffu, v[k], o);
#pragma unroll
    for (int k = 0; k < kNumSums; ++k) v[k] += w[k];
  }
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < kNumSums; ++k) sred[k][warp] = v[k];
  }
  __syncthreads();
  if (threadIdx.x < kNumSums) {
    double s = 0.0;
    for (int w = 0; w < nwarps; ++w) s += sred[threadIdx.x][w];
    partials[static_cast<size_t>(row) * kNumSums + threadIdx.x] = s;
  }
  if (pr.out == nullptr) {  // deferred: publish what the finaliser needs and leave
    if (row == 0 && threadIdx.x == 0) {
      pr.slot->n_rows = n_rows;
      pr.slot->mask_sum = mask_sum;
      pr.slot->wv = static_cast<double>(h.wv);
      pr.slot->we = static_cast<double>(h.we);
    }
    return;
  }
  if (threadIdx.x < kNumSums) __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(&pr.slot->ticket, 1u);
    is_last = (done == static_cast<unsigned int>(n_rows) - 1u);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  fold_rows_and_write(partials, n_rows, mask_sum, static_cast<double>(h.wv), static_cast<double>(h.we), sred, pr.out,
                      pr.out_f32);
  if (threadIdx.x == 0) pr.slot->ticket = 0u;  // ready for the next launch on this slot
}

__device__ __forceinline__ void unpack4(const float4 v, float (&a)[4]) { a[0] = v.x, a[1] = v.y, a[2] = v.z, a[3] = v.w; }

// Policy-side loads of the register path.  Default: streaming loads (ld.global.cs, evict-first).  SRL_LOSS_POLICY_NC=1
// (a build variant for profiles/, with SRL_LOSS_UNROLL=2): non-coherent loads (ld.global.nc), which the compiler may hoist
// above the gradient stores of the previous row -- the policy outputs are never written by this kernel.
#ifndef SRL_LOSS_POLICY_NC
#define SRL_LOSS_POLICY_NC 0
#endif
template <class V>
__device__ __forceinline__ V ld_policy(const V* p) {
#if SRL_LOSS_POLICY_NC
  return __ldg(p);
#else
  return ldg_stream(p);
#endif
}

// ---- asynchronous global -> shared copies (LDGSTS): the loads of the next rows are in flight while the current
// row is computed, without holding registers for them -----------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
               "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;" ::"r"(static_cast<uint32_t>(__cvta_generic_to_shared(smem_dst))),
               "l"(gmem_src)
               : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

#ifndef SRL_LOSS_STAGES
#define SRL_LOSS_STAGES 2
#endif
#ifndef SRL_LOSS_PIPE
#define SRL_LOSS_PIPE 0  // measured on B200 (profiles/r1c_notes.md): no gain over the register path at cfg2 or cfg5
#endif
constexpr bool kUsePipe = SRL_LOSS_PIPE != 0;
constexpr int kStages = SRL_LOSS_STAGES;  // rows in flight per thread in the pipelined kernels
constexpr int kPlanes = 7;                // 16-byte items per thread and row: 3 policy + 4 sample side

// Dynamic shared memory of the pipelined kernels: [kStages][kPlanes][threads] float4 + [kStages][threads] uint32.
__host__ __device__ constexpr size_t loss_smem_bytes(int threads) {
  return static_cast<size_t>(kStages) * threads * (kPlanes * 16 + 4);
}

// Work decomposition: grid = (column tiles, row groups, problems).  A CTA owns `blockDim * LANES` adjacent policy-side
// lanes and the rows [r0, r1) of one problem: its gather indices are loaded once, and it walks its rows in order.
// LANES = 4: a thread owns four adjacent lanes (n % 4 == 0, 16-byte aligned rows); LANES = 1: any shape.
// PIPE (dense and pack forms, LANES = 4): row t + kStages is copied global -> shared (16 bytes per cp.async, each
// thread into its own slots, so no block barrier) while row t is computed.
// Policy-side rows a CTA asks into L2 before it waits for the scan.  Only for programmatic launches of short tiles
// (cfg2: 5 rows per CTA, the whole tile): measured on one box (profiles/r1d_notes.md), the requests cost the kernel
// ~2 us when it runs alone (26.6 -> 28.7 us at cfg2, 76.4 -> 78.6 us at cfg5: the demand loads follow at once and find
// their lines still in flight), and win 1.3 us per step when they are issued under the scan (cfg2: 43.1 -> 41.7 us).
#ifndef SRL_LOSS_PREFETCH_ROWS
#define SRL_LOSS_PREFETCH_ROWS 6
#endif
constexpr int kPrefetchRows = SRL_LOSS_PREFETCH_ROWS;

template <int LANES, int MODE, class CFG>
__global__ void __launch_bounds__(256, SRL_LOSS_MIN_BLOCKS) ppo_loss_kernel(const __grid_constant__ LossBatch b) {
  constexpr bool PIPE = kUsePipe && (LANES == 4) && (MODE == kDense || MODE == kPack);
  const LossShared& s = b.s;
  const LossHyperDev& h = s.h;
  const Problem& pr = b.prob[blockIdx.z];
  const int T = s.T, n = s.n;
  const int r0 = blockIdx.y * s.rows_per_tile;
  const int r1 = min(T, r0 + s.rows_per_tile);
  Acc acc;
  RowSums rs;
  int pending = 0;
  double mask_sum = 0.0;
  Uniforms u;
  bool have_u = false;

  // Launched programmatically behind the scan (common.cuh), this CTA may be resident while the scan still runs.  The
  // policy-side rows do not depend on it: ask for this CTA's share (three tensors x its rows, one request per 128-byte
  // line) to be brought into L2 under the scan, then wait for the scan's results.
  if constexpr (LANES >= 2 && kPrefetchRows > 0) {
    if (s.prefetch_rows > 0 && (threadIdx.x & (32 / LANES - 1)) == 0) {  // one thread per 128-byte line
      const int j = (blockIdx.x * blockDim.x + threadIdx.x) * LANES;
      if (j < n) {
        long long o = static_cast<long long>(r0) * s.ld_pol + j;
        const int r_pf = min(r1, r0 + s.prefetch_rows);
        for (int t = r0; t < r_pf; ++t, o += s.ld_pol) {
          prefetch_l2(pr.new_logp + o);
          prefetch_l2(pr.v_pred + o);
          prefetch_l2(pr.entropy + o);
        }
      }
    }
  }
  pdl_wait();

  // normally one column tile per CTA (gridDim.x == col_tiles); only batches wider than kMaxGrid tiles loop here
  for (int ct = blockIdx.x; ct < s.col_tiles; ct += gridDim.x) {
  const int j = (ct * blockDim.x + threadIdx.x) * LANES;
  const bool active = j < n;
  // gather indices of this thread's lanes: once per column tile
  int c[LANES];
  if (pr.lane_idx && active) {
    if constexpr (LANES == 4) {
      const int4 q = __ldg(reinterpret_cast<const int4*>(pr.lane_idx + j));
      c[0] = q.x, c[1] = q.y, c[2] = q.z, c[3] = q.w;
    } else if constexpr (LANES == 2) {
      const int2 q = __ldg(reinterpret_cast<const int2*>(pr.lane_idx + j));
      c[0] = q.x, c[1] = q.y;
    } else {
      c[0] = __ldg(pr.lane_idx + j);
    }
  } else {
#pragma unroll
    for (int q = 0; q < LANES; ++q) c[q] = j + q;
  }

  if constexpr (PIPE) {
    extern __shared__ __align__(16) unsigned char ring_raw[];
    float4* ring = reinterpret_cast<float4*>(ring_raw);
    uint32_t* fring = reinterpret_cast<uint32_t*>(ring_raw + static_cast<size_t>(kStages) * kPlanes * blockDim.x * 16);
    const int tid = threadIdx.x, nthr = blockDim.x;
    auto slot = [&](int st, int plane) { return ring + (st * kPlanes + plane) * nthr + tid; };
    auto issue = [&](int t, int st) {  // all loads of row t into stage st
      if (active && t < r1) {
        const long long op = static_cast<long long>(t) * s.ld_pol + j;
        cp_async16(slot(st, 0), pr.new_logp + op);
        cp_async16(slot(st, 1), pr.v_pred + op);
        cp_async16(slot(st, 2), pr.entropy + op);
        const long long ob = static_cast<long long>(t) * s.ld_smp;
        if constexpr (MODE == kPack) {
#pragma unroll
          for (int q = 0; q < 4; ++q) cp_async16(slot(st, 3 + q), s.pack + ob + c[q]);
        } else {
          cp_async16(slot(st, 3), s.old_logp + ob + j);
          cp_async16(slot(st, 4), s.ret + ob + j);
          cp_async16(slot(st, 5), s.adv + ob + j);
          if (CFG::clip(h)) cp_async16(slot(st, 6), s.old_value + ob + j);
          cp_async4(fring + st * nthr + tid, s.reset_next + ob + j);
        }
      }
      cp_async_commit();  // one group per row, issued or not, so wait_group counts rows
    };
#pragma unroll
    for (int k = 0; k < kStages; ++k) issue(r0 + k, k);
    // the statistics loads and the float64 divisions / sqrt on them run under the first rows' copies
    if (!have_u) {
      if (s.lane_part != nullptr)
        u = self_uniforms<LANES>(s.lane_part, s.lane_part_n, c, active, h.adv_eps, mask_sum);
      else
        u = load_uniforms(pr.norm_stats, pr.local_stats, s.popart, h.adv_eps, mask_sum);
      have_u = true;
    }
    float* glp_row = pr.g_logp + static_cast<long long>(r0) * s.ld_grad + j;
    float* gv_row = pr.g_value + static_cast<long long>(r0) * s.ld_grad + j;
    float* ge_row = pr.g_entropy + static_cast<long long>(r0) * s.ld_grad + j;
    int st = 0;
    for (int t = r0; t < r1; ++t, glp_row += s.ld_grad, gv_row += s.ld_grad, ge_row += s.ld_grad) {
      cp_async_wait<kStages - 1>();  // row t has landed (rows t+1 .. t+kStages-1 may still be in flight)
      float nl[4], vp[4], en[4], ol[4], ov[4], rt[4], ad[4];
      bool valid[4];
      unpack4(*slot(st, 0), nl);
      unpack4(*slot(st, 1), vp);
      unpack4(*slot(st, 2), en);
      if constexpr (MODE == kPack) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const float4 k = *slot(st, 3 + q);
          ol[q] = k.x;
          ov[q] = k.y;
          rt[q] = k.z;
          ad[q] = k.w;
          valid[q] = (k.w == k.w);  // K2 stores NaN in the advantage slot of masked transitions
        }
      } else {
        unpack4(*slot(st, 3), ol);
        unpack4(*slot(st, 4), rt);
        unpack4(*slot(st, 5), ad);
        float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
        if (CFG::clip(h)) o = *slot(st, 6);
        unpack4(o, ov);
        const uint32_t m = fring[st * nthr + tid];
#pragma unroll
        for (int q = 0; q < 4; ++q) valid[q] = ((m >> (8 * q)) & 0xffu) == 0u;
      }
      issue(t + kStages, st);  // refill the stage just drained (its values are in registers now)
      st = (st + 1 == kStages) ? 0 : st + 1;
      if (active) {
        float glp[4], gv[4], ge[4];
#pragma unroll
        for (int q = 0; q < 4; ++q)
          element<CFG>(h, u, nl[q], vp[q], en[q], ol[q], ov[q], rt[q], ad[q], valid[q], glp[q], gv[q], ge[q], rs);
        if (++pending == kFlushRows) {  // fp32 partial sums of at most kFlushRows * LANES terms, then float64
          acc.add(rs);
          rs = RowSums();
          pending = 0;
        }
        stg_stream(reinterpret_cast<float4*>(glp_row), make_float4(glp[0], glp[1], glp[2], glp[3]));
        stg_stream(reinterpret_cast<float4*>(gv_row), make_float4(gv[0], gv[1], gv[2], gv[3]));
        stg_stream(reinterpret_cast<float4*>(ge_row), make_float4(ge[0], ge[1], ge[2], ge[3]));
      }
    }
    cp_async_wait<0>();
  } else {
    // a run of four consecutive, 4-aligned lanes (agents of one environment, sector-aligned environment blocks) is
    // fetched with one 128-bit load per leaf
    bool run4 = false;
    if constexpr (LANES == 4 && MODE == kGather)
      run4 = s.smp_vec_ok && (c[1] == c[0] + 1) && (c[2] == c[0] + 2) && (c[3] == c[0] + 3) && ((c[0] & 3) == 0);
    // row base pointers are warp-uniform (64-bit, advanced once per row); per-thread offsets stay 32-bit
    const float* nl_row = pr.new_logp + static_cast<long long>(r0) * s.ld_pol;
    const float* vp_row = pr.v_pred + static_cast<long long>(r0) * s.ld_pol;
    const float* en_row = pr.entropy + static_cast<long long>(r0) * s.ld_pol;
    float* glp_row = pr.g_logp + static_cast<long long>(r0) * s.ld_grad;
    float* gv_row = pr.g_value + static_cast<long long>(r0) * s.ld_grad;
    float* ge_row = pr.g_entropy + static_cast<long long>(r0) * s.ld_grad;
    long long ob = static_cast<long long>(r0) * s.ld_smp;  // sample-side row base (elements)
    float nl[LANES], vp[LANES], en[LANES], ol[LANES], ov[LANES], rt[LANES], ad[LANES];
    bool valid[LANES];
    // every global load of the row at (nl_row, vp_row, en_row, ob)
    auto load_row = [&]() {
      if constexpr (LANES == 4) {
        unpack4(ld_policy(reinterpret_cast<const float4*>(nl_row + j)), nl);
        unpack4(ld_policy(reinterpret_cast<const float4*>(vp_row + j)), vp);
        unpack4(ld_policy(reinterpret_cast<const float4*>(en_row + j)), en);
      } else if constexpr (LANES == 2) {
        const float2 a = ld_policy(reinterpret_cast<const float2*>(nl_row + j));
        const float2 b2 = ld_policy(reinterpret_cast<const float2*>(vp_row + j));
        const float2 e2 = ld_policy(reinterpret_cast<const float2*>(en_row + j));
        nl[0] = a.x, nl[1] = a.y, vp[0] = b2.x, vp[1] = b2.y, en[0] = e2.x, en[1] = e2.y;
      } else {
        nl[0] = ld_policy(nl_row + j);
        vp[0] = ld_policy(vp_row + j);
        en[0] = ld_policy(en_row + j);
      }
      if constexpr (MODE == kPack) {
        const float4* pack_row = s.pack + ob;
#pragma unroll
        for (int q = 0; q < LANES; ++q) {
          const float4 k = __ldg(pack_row + c[q]);
          ol[q] = k.x;
          ov[q] = k.y;
          rt[q] = k.z;
          ad[q] = k.w;
          valid[q] = (k.w == k.w);
        }
      } else if (run4) {
        if constexpr (LANES == 4) {
          unpack4(ldg_stream(reinterpret_cast<const float4*>(s.old_logp + ob + c[0])), ol);
          unpack4(ldg_stream(reinterpret_cast<const float4*>(s.ret + ob + c[0])), rt);
          unpack4(ldg_stream(reinterpret_cast<const float4*>(s.adv + ob + c[0])), ad);
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
          if (CFG::clip(h)) o = ldg_stream(reinterpret_cast<const float4*>(s.old_value + ob + c[0]));
          unpack4(o, ov);
          const uint32_t m = ldg_stream(reinterpret_cast<const uint32_t*>(s.reset_next + ob + c[0]));
#pragma unroll
          for (int q = 0; q < LANES; ++q) valid[q] = ((m >> (8 * q)) & 0xffu) == 0u;
        }
      } else {
        const float* ol_row = s.old_logp + ob;
        const float* rt_row = s.ret + ob;
        const float* ad_row = s.adv + ob;
        const float* ov_row = CFG::clip(h) ? s.old_value + ob : nullptr;
        const uint8_t* rs_row = s.reset_next + ob;
#pragma unroll
        for (int q = 0; q < LANES; ++q) {
          ol[q] = __ldg(ol_row + c[q]);
          rt[q] = __ldg(rt_row + c[q]);
          ad[q] = __ldg(ad_row + c[q]);
          ov[q] = CFG::clip(h) ? __ldg(ov_row + c[q]) : 0.f;
          valid[q] = __ldg(rs_row + c[q]) == 0;
        }
      }
    };
    // (Issuing the first row's loads here, ahead of the statistics prologue, and each later row right after the stores of
    // the row before it was built and measured: it loses, 26.7 -> 33.0 us at cfg2 -- profiles/r1d_notes.md.)
    if (!have_u) {
      if (s.lane_part != nullptr)
        u = self_uniforms<LANES>(s.lane_part, s.lane_part_n, c, active, h.adv_eps, mask_sum);
      else
        u = load_uniforms(pr.norm_stats, pr.local_stats, s.popart, h.adv_eps, mask_sum);
      have_u = true;
    }
    if (active) {
#pragma unroll kLossUnroll
      for (int t = r0; t < r1; ++t) {
        load_row();
        float glp[LANES], gv[LANES], ge[LANES];
#pragma unroll
        for (int q = 0; q < LANES; ++q)
          element<CFG>(h, u, nl[q], vp[q], en[q], ol[q], ov[q], rt[q], ad[q], valid[q], glp[q], gv[q], ge[q], rs);
        if (++pending == kFlushRows) {
          acc.add(rs);
          rs = RowSums();
          pending = 0;
        }
        if constexpr (LANES == 4) {
          stg_stream(reinterpret_cast<float4*>(glp_row + j), make_float4(glp[0], glp[1], glp[2], glp[3]));
          stg_stream(reinterpret_cast<float4*>(gv_row + j), make_float4(gv[0], gv[1], gv[2], gv[3]));
          stg_stream(reinterpret_cast<float4*>(ge_row + j), make_float4(ge[0], ge[1], ge[2], ge[3]));
        } else if constexpr (LANES == 2) {
          stg_stream(reinterpret_cast<float2*>(glp_row + j), make_float2(glp[0], glp[1]));
          stg_stream(reinterpret_cast<float2*>(gv_row + j), make_float2(gv[0], gv[1]));
          stg_stream(reinterpret_cast<float2*>(ge_row + j), make_float2(ge[0], ge[1]));
        } else {
          stg_stream(glp_row + j, glp[0]);
          stg_stream(gv_row + j, gv[0]);
          stg_stream(ge_row + j, ge[0]);
        }
        nl_row += s.ld_pol, vp_row += s.ld_pol, en_row += s.ld_pol;
        glp_row += s.ld_grad, gv_row += s.ld_grad, ge_row += s.ld_grad;
        ob += s.ld_smp;
      }
    }
  }
  }  // column tiles
  acc.add(rs);
  const int row = blockIdx.y * gridDim.x + blockIdx.x;
  reduce_and_finalize(pr, h, acc, mask_sum, row, static_cast<int>(gridDim.x * gridDim.y));
}

template <int LANES, int MODE, class CFG>
struct LossLauncher {
  static constexpr bool PIPE = kUsePipe && (LANES == 4) && (MODE == kDense || MODE == kPack);

  // CTAs of `threads` threads one SM holds (registers / shared memory), asked once per device and block size
  static int resident(int threads) {
    static int cached[64][2] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 2;
    int& c = cached[dev][threads >= 256 ? 1 : 0];
    if (c == 0) {
      auto kern = ppo_loss_kernel<LANES, MODE, CFG>;
      const size_t smem = PIPE ? loss_smem_bytes(256) : 0;
      if (PIPE) cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
      int n = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, kern, threads, PIPE ? loss_smem_bytes(threads) : 0) !=
              cudaSuccess ||
          n < 1)
        n = 1;
      c = n;
    }
    return c;
  }

  static int launch(LossBatch& b, int n_problems, cudaStream_t st) {
    LossShared& s = b.s;
    const int per_row = (s.n + LANES - 1) / LANES;  // threads one row needs
    const int threads = per_row <= 128 ? 128 : 256;
    s.col_tiles = (per_row + threads - 1) / threads;
    SRL_REQUIRE(s.lane_part == nullptr || s.col_tiles == 1, SRL_ERR_UNSUPPORTED,
                "ppo loss: self-computed statistics need the whole minibatch in one column tile (n <= %d lanes)",
                256 * LANES);
    const long long capacity = static_cast<long long>(sm_count()) * resident(threads);  // CTAs resident at once
    // row groups: as many as keep the whole grid resident in ONE wave (a second, partial wave would double the
    // kernel's duration); every CTA then walks ceil(T / groups) consecutive rows of its column tile
    long long groups = capacity / (static_cast<long long>(s.col_tiles) * n_problems);
    if (groups < 1) groups = 1;
    if (groups > s.T) groups = s.T;
    int rows = static_cast<int>((s.T + groups - 1) / groups);
    groups = (s.T + rows - 1) / rows;
    long long gx = s.col_tiles;  // CTAs along the lanes; wider batches loop over column tiles inside the kernel
    if (gx * groups > kMaxGrid) gx = kMaxGrid / groups;  // one partial row per CTA in the workspace slot
    s.rows_per_tile = rows;
    s.n_tiles = static_cast<int>(groups);
    s.prefetch_rows = (pdl_enabled() && rows <= kPrefetchRows) ? rows : 0;
    const dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(groups), static_cast<unsigned>(n_problems));
    SRL_CUDA(launch_pdl(ppo_loss_kernel<LANES, MODE, CFG>, grid, dim3(threads), PIPE ? loss_smem_bytes(threads) : 0, st, b));
    return SRL_OK;
  }
};

// Picks the instantiation for the launch's hyper-parameters.  Static configurations: {mse, huber} x clip_value x
// dual_clip x popart with four lanes per thread; everything else (smoothl1, odd shapes) runs the general instantiation.
template <int LANES, int MODE>
int launch_loss_static(LossBatch& b, int n_problems, cudaStream_t st) {
  const LossHyperDev& h = b.s.h;
  const bool popart = b.s.popart != nullptr;
  const int key = (h.value_loss == SRL_VL_MSE ? 0 : h.value_loss == SRL_VL_HUBER ? 1 : 2) * 8 + (h.clip_value ? 4 : 0) +
                  (h.dual_clip ? 2 : 0) + (popart ? 1 : 0);
#define SRL_LOSS_CASE(k, vl, clip, dual, pa) \
  case k:                                    \
    return LossLauncher<LANES, MODE, StaticCfg<vl, clip, dual, pa>>::launch(b, n_problems, st)
  switch (key) {
    SRL_LOSS_CASE(0, SRL_VL_MSE, false, false, false);
    SRL_LOSS_CASE(1, SRL_VL_MSE, false, false, true);
    SRL_LOSS_CASE(2, SRL_VL_MSE, false, true, false);
    SRL_LOSS_CASE(3, SRL_VL_MSE, false, true, true);
    SRL_LOSS_CASE(4, SRL_VL_MSE, true, false, false);
    SRL_LOSS_CASE(5, SRL_VL_MSE, true, false, true);
    SRL_LOSS_CASE(6, SRL_VL_MSE, true, true, false);
    SRL_LOSS_CASE(7, SRL_VL_MSE, true, true, true);
    SRL_LOSS_CASE(8, SRL_VL_HUBER, false, false, false);
    SRL_LOSS_CASE(9, SRL_VL_HUBER, false, false, true);
    SRL_LOSS_CASE(10, SRL_VL_HUBER, false, true, false);
    SRL_LOSS_CASE(11, SRL_VL_HUBER, false, true, true);
    SRL_LOSS_CASE(12, SRL_VL_HUBER, true, false, false);
    SRL_LOSS_CASE(13, SRL_VL_HUBER, true, false, true);
    SRL_LOSS_CASE(14, SRL_VL_HUBER, true, true, false);
    SRL_LOSS_CASE(15, SRL_VL_HUBER, true, true, true);
    default:
      return LossLauncher<LANES, MODE, RuntimeCfg>::launch(b, n_problems, st);
  }
#undef SRL_LOSS_CASE
}

template <int MODE>
int launch_loss_mode(LossBatch& b, int n_problems, bool lanes4, cudaStream_t st) {
  if (!lanes4) {
    if constexpr (MODE == kDense)
      return SRL_ERR_INVALID_ARG;  // the dense form needs 128-bit rows; callers route odd shapes to kGather
    else
      return LossLauncher<1, MODE, RuntimeCfg>::launch(b, n_problems, st);
  }
  return launch_loss_static<4, MODE>(b, n_problems, st);
}

// defined in ppo_loss_dense.cu / ppo_loss_gather.cu / ppo_loss_pack.cu
int launch_loss_dense(LossBatch& b, int n_problems, bool lanes4, cudaStream_t st);
int launch_loss_gather(LossBatch& b, int n_problems, bool lanes4, cudaStream_t st);
int launch_loss_pack(LossBatch& b, int n_problems, bool lanes4, cudaStream_t st);
// ppo_loss_pack2.cu: the pack form with TWO lanes per thread (twice the threads per row, half the registers a row's loads
// hold): an experiment for small minibatches, reached only with SRL_LOSS_LANES=2 in the environment (ppo_loss.cu)
int launch_loss_pack2(LossBatch& b, int n_problems, cudaStream_t st);

}  // namespace loss
}  // namespace srl
