// K4: fused PPO / MAPPO loss forward + backward (see include/srl_b200.h; replaces mappo.py:146-217,
// utils.py:10-67,228-265 and the autograd backward to new_logp / state_values / entropy).
//
// One pass over the [T, n] minibatch: each element reads 29 B (25 B without value clipping) and
// writes 12 B of gradients.  The advantage-normalisation statistics arrive pre-reduced (and, on
// several GPUs, pre-all-reduced) from K2 + srl_group_stats, so there is no second pass and no grid
// barrier.
//
// Work decomposition (r1c; profiles/r1c_notes.md has the ncu evidence behind each choice):
//  * a launch covers up to SRL_MAX_LOSS_BATCH minibatches ("problems") of one shape -- blockIdx.y picks the
//    problem -- so the 32 tiny per-minibatch launches of a 4 x 8 PPO schedule become one grid that fills the
//    machine;
//  * inside a problem a CTA walks tiles of `rows_per_tile` rows x (blockDim * 4) lanes; a thread owns four
//    adjacent lanes (128-bit accesses on the dense policy side) and keeps its four gather indices in
//    registers for all rows of the tile, so index loads and the run-of-4 test are paid once per tile;
//  * the sample side comes in one of three forms: dense leaves (whole batch, no permutation), leaves
//    gathered through lane_idx (any alignment), or K2's `pack` ([T, N] float4 = old_logp, old_value, ret,
//    adv-or-NaN): one 16-byte gather per element instead of five 1..4-byte ones -- the L1 wavefront count,
//    which is what bounds a per-lane gather, drops 5x;
//  * per-element arithmetic is fp32 except the advantage / PopArt normalisation, which stays in float64
//    (reference: utils.py:38-67,139-144) as subtract + multiply by a correctly rounded reciprocal + one
//    Newton residual step (3 fp64 instructions instead of a ~25-instruction IEEE division; the quotient is
//    the correctly rounded one, then cast to float as the reference does);
//  * eight masked sums: four lanes are summed in fp32, then accumulated in float64 per thread, reduced
//    warp-shuffle -> shared memory -> one float64 partial row per CTA in the problem's workspace slot.
// Finalisation: immediate (out != NULL): the last CTA of the problem (atomic ticket) folds the rows in a fixed
// order; deferred (out == NULL): srl_ppo_loss_finalize folds any number of slots later in ONE launch.
// Either way the result is deterministic for a given launch shape and nothing syncs with the host.
#include <float.h>

#include "common.cuh"

namespace srl {
namespace {

#ifndef SRL_LOSS_UNROLL
#define SRL_LOSS_UNROLL 2
#endif
constexpr int kLossUnroll = SRL_LOSS_UNROLL;  // rows of a tile whose loads are in flight together
constexpr int kMaxGrid = 2048;  // partial rows per workspace slot
constexpr int kNumSums = 8;
constexpr size_t kPartialsOffset = 64;

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
  long long ld_pol, ld_grad, ld_smp;
  int T, n;
  int rows_per_tile, col_tiles, n_tiles;
  int smp_vec_ok;  // sample-side leaves allow 128-bit loads at 4-aligned lanes (alignment + row stride)
  LossHyperDev h;
};

template <int NB>
struct LossBatch {
  LossShared s;
  Problem prob[NB];
};

// torch.nn.{MSELoss,HuberLoss,SmoothL1Loss}(reduction='none') value and derivative wrt the input.
// `kind` is uniform across the launch; inside a kind both branches are evaluated and selected (the quadratic /
// linear choice is per element, and a divergent branch costs more than the three spare flops).
__device__ __forceinline__ void pointwise_loss(int kind, float prm, float d, float& l, float& dl) {
  if (kind == SRL_VL_MSE) {
    l = d * d;
    dl = 2.f * d;
  } else if (kind == SRL_VL_HUBER) {
    const float z = fabsf(d);
    const bool quad = z < prm;
    const float lq = 0.5f * z * z, ll = prm * (z - 0.5f * prm);
    l = quad ? lq : ll;
    dl = quad ? d : (d > 0.f ? prm : -prm);
  } else {
    const float z = fabsf(d);
    const bool quad = z < prm;
    const float lq = 0.5f * z * z / prm, ll = z - 0.5f * prm;
    l = quad ? lq : ll;
    dl = quad ? d / prm : (d > 0.f ? 1.f : -1.f);
  }
}

struct Uniforms {
  double mean, denom, rdenom;    // advantage normalisation: (x - mean) / denom, rdenom = 1 / denom
  double pa_mu, pa_sd, pa_rsd;   // popart: (x - mu) / sd
  float inv_m;                   // 1 / local sum(mask)
  bool popart;
};

// The six scalars every thread needs.  Loading them is split from the math on them so that the (long) fp64
// divide / sqrt latency overlaps the element loads instead of preceding them.
struct RawStats {
  double cnt, s1, s2, m_local, pa_mu, pa_sd;
  bool popart;
};

__device__ __forceinline__ RawStats load_raw_stats(const double* norm_stats, const double* local_stats,
                                                   const double* popart) {
  RawStats r;
  r.cnt = __ldg(norm_stats);
  r.s1 = __ldg(norm_stats + 1);
  r.s2 = __ldg(norm_stats + 2);
  r.m_local = __ldg(local_stats);
  r.popart = popart != nullptr;
  r.pa_mu = r.popart ? __ldg(popart) : 0.0;
  r.pa_sd = r.popart ? __ldg(popart + 1) : 1.0;
  return r;
}

__device__ __forceinline__ Uniforms make_uniforms(const RawStats& r, double adv_eps) {
  Uniforms u;
  u.popart = r.popart;
  u.pa_mu = r.pa_mu;
  u.pa_sd = r.pa_sd;
  u.pa_rsd = 1.0 / r.pa_sd;
  u.mean = r.s1 / r.cnt;
  const double var = r.s2 / r.cnt - u.mean * u.mean;  // biased variance, utils.py:62-64
  u.denom = sqrt(var) + adv_eps;                      // eps outside the sqrt, utils.py:67
  u.rdenom = 1.0 / u.denom;
  u.inv_m = 1.f / static_cast<float>(r.m_local);
  return u;
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

// Per-row (4 lanes) fp32 partial sums; folded into the thread's float64 accumulators once per row.
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
__device__ __forceinline__ void element(const LossHyperDev& h, const Uniforms& u, float nl, float vp, float en,
                                        float ol, float ov, float rt, float ad, bool valid, float& g_lp,
                                        float& g_v, float& g_en, RowSums& rs) {
  const float scale = valid ? u.inv_m : 0.f;  // d(masked mean)/d(element) = mask / M

  // ---- critic: mappo.py:172-184, utils.py:228-239 ---------------------------------------------
  const float vt = u.popart ? popart_normalize(rt, u) : rt;
  float l, dl;
  pointwise_loss(h.value_loss, h.vl_param, vp - vt, l, dl);
  float vl = l, gv = dl;
  if (h.clip_value) {
    const float ovn = h.normalize_old_value ? popart_normalize(ov, u) : ov;
    const float dv = vp - ovn;
    const float vc = ovn + fminf(fmaxf(dv, -h.veps), h.veps);
    const float in = (dv >= -h.veps && dv <= h.veps) ? 1.f : 0.f;  // clamp passes grad on the closed interval
    float l2, dl2;
    pointwise_loss(h.value_loss, h.vl_param, vc - vt, l2, dl2);
    dl2 *= in;
    vl = fmaxf(l, l2);
    gv = l > l2 ? dl : (l < l2 ? dl2 : 0.5f * (dl + dl2));  // torch.max splits ties evenly
  }
  g_v = valid ? h.wv * scale * gv : 0.f;

  // ---- actor: mappo.py:157-158,186-197 ---------------------------------------------------------
  const float ratio = expf(nl - ol);
  // masked_normalization (utils.py:38-67) in float64, cast to float at the end; masked entries are centred
  // zeros there (x = adv * mask) but they never reach the loss or the gradients
  const float nadv = static_cast<float>(div_by(__dsub_rn(static_cast<double>(ad), u.mean), u.denom, u.rdenom));
  const float s1 = ratio * nadv;
  const float s2 = fminf(fmaxf(ratio, h.clip_lo), h.clip_hi) * nadv;
  const float in_clip = (ratio >= h.clip_lo && ratio <= h.clip_hi) ? 1.f : 0.f;
  const float w1 = s1 < s2 ? 1.f : (s1 > s2 ? 0.f : 0.5f);  // torch.min splits ties evenly
  float obj = fminf(s1, s2);
  float gsum = (w1 + (1.f - w1) * in_clip) * nadv * ratio;
  if (h.dual_clip) {
    const float sgn = nadv > 0.f ? 1.f : (nadv < 0.f ? -1.f : 0.f);
    const float s3 = -sgn * h.c_clip * nadv;
    gsum *= obj > s3 ? 1.f : (obj < s3 ? 0.f : 0.5f);
    obj = fmaxf(obj, s3);
  }
  g_lp = valid ? -scale * gsum : 0.f;
  g_en = -h.we * scale;  // entropy_loss = -sum(entropy * mask) / M   mappo.py:199

  rs.pl += valid ? -obj : 0.f;
  rs.vl += valid ? vl : 0.f;
  rs.en += valid ? en : 0.f;
  rs.adv += valid ? ad : 0.f;
  rs.ratio += valid ? ratio : 0.f;
  rs.vt += valid ? vt : 0.f;
  rs.ret += valid ? rt : 0.f;
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

__global__ void __launch_bounds__(256) loss_finalize_kernel(const unsigned char* __restrict__ ws, size_t slot_bytes,
                                                            double* __restrict__ out, float* __restrict__ out_f32) {
  __shared__ double sred[kNumSums][8];
  const unsigned char* base = ws + static_cast<size_t>(blockIdx.x) * slot_bytes;
  const SlotHeader* hd = reinterpret_cast<const SlotHeader*>(base);
  fold_rows_and_write(reinterpret_cast<const double*>(base + kPartialsOffset), static_cast<int>(hd->n_rows),
                      hd->mask_sum, hd->wv, hd->we, sred, out + static_cast<size_t>(blockIdx.x) * SRL_LOSS_OUT_LEN,
                      out_f32 ? out_f32 + static_cast<size_t>(blockIdx.x) * 4 : nullptr);
}

__device__ __forceinline__ void unpack4(const float4 v, float (&a)[4]) { a[0] = v.x, a[1] = v.y, a[2] = v.z, a[3] = v.w; }

// Sample-side forms
constexpr int kDense = 0;   // separate leaves, lanes in policy order, 128-bit loads
constexpr int kGather = 1;  // separate leaves through lane_idx (or any alignment)
constexpr int kPack = 2;    // K2's float4 pack through lane_idx (or in order)

// LANES = 4: a thread owns four adjacent policy-side lanes (n % 4 == 0, 16-byte aligned rows); LANES = 1: any shape.
template <int LANES, int MODE, int NB>
__global__ void __launch_bounds__(256) ppo_loss_kernel(const __grid_constant__ LossBatch<NB> b) {
  const LossShared& s = b.s;
  const LossHyperDev& h = s.h;
  const Problem& pr = b.prob[blockIdx.y];
  const RawStats raw = load_raw_stats(pr.norm_stats, pr.local_stats, s.popart);  // loads only
  Uniforms u;
  bool have_u = false;
  Acc acc;
  const int T = s.T, n = s.n;
  const int tile_lanes = blockDim.x * LANES;

  for (int tile = blockIdx.x; tile < s.n_tiles; tile += gridDim.x) {
    const int rt_i = tile / s.col_tiles;
    const int ct_i = tile - rt_i * s.col_tiles;
    const int j = ct_i * tile_lanes + threadIdx.x * LANES;
    if (j >= n) continue;
    const int t0 = rt_i * s.rows_per_tile;
    const int t1 = min(T, t0 + s.rows_per_tile);

    // gather indices of this thread's lanes: once per tile
    int c[LANES];
    if (pr.lane_idx) {
      if constexpr (LANES == 4) {
        const int4 q = __ldg(reinterpret_cast<const int4*>(pr.lane_idx + j));
        c[0] = q.x, c[1] = q.y, c[2] = q.z, c[3] = q.w;
      } else {
        c[0] = __ldg(pr.lane_idx + j);
      }
    } else {
#pragma unroll
      for (int q = 0; q < LANES; ++q) c[q] = j + q;
    }
    // a run of four consecutive, 4-aligned lanes (dense batches, agents of one environment, sector-aligned
    // environment blocks) is fetched with one 128-bit load per leaf
    bool run4 = false;
    if constexpr (LANES == 4 && MODE == kGather)
      run4 = s.smp_vec_ok && (c[1] == c[0] + 1) && (c[2] == c[0] + 2) && (c[3] == c[0] + 3) && ((c[0] & 3) == 0);

    const float* p_nl = pr.new_logp + static_cast<long long>(t0) * s.ld_pol + j;
    const float* p_vp = pr.v_pred + static_cast<long long>(t0) * s.ld_pol + j;
    const float* p_en = pr.entropy + static_cast<long long>(t0) * s.ld_pol + j;
    float* p_glp = pr.g_logp + static_cast<long long>(t0) * s.ld_grad + j;
    float* p_gv = pr.g_value + static_cast<long long>(t0) * s.ld_grad + j;
    float* p_ge = pr.g_entropy + static_cast<long long>(t0) * s.ld_grad + j;
    long long ob = static_cast<long long>(t0) * s.ld_smp;

#pragma unroll kLossUnroll
    for (int t = t0; t < t1; ++t) {
      float nl[LANES], vp[LANES], en[LANES], ol[LANES], ov[LANES], rt[LANES], ad[LANES];
      bool valid[LANES];
      // ---- policy side: dense ------------------------------------------------------------------------
      if constexpr (LANES == 4) {
        unpack4(ldg_stream(reinterpret_cast<const float4*>(p_nl)), nl);
        unpack4(ldg_stream(reinterpret_cast<const float4*>(p_vp)), vp);
        unpack4(ldg_stream(reinterpret_cast<const float4*>(p_en)), en);
      } else {
        nl[0] = ldg_stream(p_nl);
        vp[0] = ldg_stream(p_vp);
        en[0] = ldg_stream(p_en);
      }
      // ---- sample side -------------------------------------------------------------------------------
      bool vec_rows = false;
      if constexpr (LANES == 4) vec_rows = (MODE == kDense) || run4;
      if constexpr (MODE == kPack) {
#pragma unroll
        for (int q = 0; q < LANES; ++q) {
          const float4 k = __ldg(s.pack + ob + c[q]);
          ol[q] = k.x;
          ov[q] = k.y;
          rt[q] = k.z;
          ad[q] = k.w;
          valid[q] = (k.w == k.w);  // K2 stores NaN in the advantage slot of masked transitions
        }
      } else if (vec_rows) {
        if constexpr (LANES == 4) {
          const long long os = ob + c[0];
          unpack4(ldg_stream(reinterpret_cast<const float4*>(s.old_logp + os)), ol);
          unpack4(ldg_stream(reinterpret_cast<const float4*>(s.ret + os)), rt);
          unpack4(ldg_stream(reinterpret_cast<const float4*>(s.adv + os)), ad);
          float4 o = make_float4(0.f, 0.f, 0.f, 0.f);
          if (h.clip_value) o = ldg_stream(reinterpret_cast<const float4*>(s.old_value + os));
          unpack4(o, ov);
          const uint32_t m = ldg_stream(reinterpret_cast<const uint32_t*>(s.reset_next + os));
#pragma unroll
          for (int q = 0; q < LANES; ++q) valid[q] = ((m >> (8 * q)) & 0xffu) == 0u;
        }
      } else {
#pragma unroll
        for (int q = 0; q < LANES; ++q) {
          const long long os = ob + c[q];
          ol[q] = __ldg(s.old_logp + os);
          rt[q] = __ldg(s.ret + os);
          ad[q] = __ldg(s.adv + os);
          ov[q] = h.clip_value ? __ldg(s.old_value + os) : 0.f;
          valid[q] = __ldg(s.reset_next + os) == 0;
        }
      }
      if (!have_u) {
        u = make_uniforms(raw, h.adv_eps);
        have_u = true;
      }
      float glp[LANES], gv[LANES], ge[LANES];
      RowSums rs;
#pragma unroll
      for (int q = 0; q < LANES; ++q)
        element(h, u, nl[q], vp[q], en[q], ol[q], ov[q], rt[q], ad[q], valid[q], glp[q], gv[q], ge[q], rs);
      acc.add(rs);
      if constexpr (LANES == 4) {
        stg_stream(reinterpret_cast<float4*>(p_glp), make_float4(glp[0], glp[1], glp[2], glp[3]));
        stg_stream(reinterpret_cast<float4*>(p_gv), make_float4(gv[0], gv[1], gv[2], gv[3]));
        stg_stream(reinterpret_cast<float4*>(p_ge), make_float4(ge[0], ge[1], ge[2], ge[3]));
      } else {
        stg_stream(p_glp, glp[0]);
        stg_stream(p_gv, gv[0]);
        stg_stream(p_ge, ge[0]);
      }
      p_nl += s.ld_pol;
      p_vp += s.ld_pol;
      p_en += s.ld_pol;
      p_glp += s.ld_grad;
      p_gv += s.ld_grad;
      p_ge += s.ld_grad;
      ob += s.ld_smp;
    }
  }
  reduce_and_finalize(pr, h, acc, raw.m_local, blockIdx.x, gridDim.x);
}

// ---- K4b: the same loss starting from the actor head's logits -----------------------------------
// (actor_critic_policy.py:303-324: Categorical(logits=slice).log_prob / .entropy per head, summed.)
// A CTA takes 256 consecutive transitions; their [256, sumK] logits block is contiguous in memory, so
// it is staged through shared memory with coalesced loads (row stride sumK+1 words: conflict-free
// per-thread row walks), turned into d loss / d logits in place, and streamed back out coalesced.
struct LogitsParams {
  LossShared s;
  Problem pr;             // new_logp / entropy / g_logp / g_entropy are unused here
  const float* logits;    // [T*n, SK]
  const int32_t* action;  // [T*n, heads]
  float* g_logits;        // [T*n, SK]
  float* logp_out;        // [T*n] or null
  float* entropy_out;     // [T*n] or null
  int heads, SK;
  int head_size[SRL_MAX_HEADS];
};

__global__ void __launch_bounds__(256) ppo_loss_logits_kernel(const __grid_constant__ LogitsParams q) {
  extern __shared__ float srow[];  // [256][SK + 1]
  const LossShared& p = q.s;
  const Problem& pr = q.pr;
  const LossHyperDev& h = p.h;
  const RawStats raw = load_raw_stats(pr.norm_stats, pr.local_stats, p.popart);
  const Uniforms u = make_uniforms(raw, h.adv_eps);
  Acc acc;
  const int SK = q.SK, stride = SK + 1;
  const long long W = static_cast<long long>(p.T) * p.n;
  const long long tiles = (W + 255) / 256;
  for (long long tile = blockIdx.x; tile < tiles; tile += gridDim.x) {
    const long long i0 = tile * 256;
    const int cnt = static_cast<int>(min(256ll, W - i0));
    // coalesced stage-in of cnt*SK contiguous floats
    const float* src = q.logits + i0 * SK;
    for (int e = threadIdx.x; e < cnt * SK; e += 256) {
      const int r = e / SK;
      srow[r * stride + (e - r * SK)] = ldg_stream(src + e);
    }
    __syncthreads();
    if (static_cast<int>(threadIdx.x) < cnt) {
      const long long i = i0 + threadIdx.x;
      const int t = static_cast<int>(i / p.n);
      const int j = static_cast<int>(i - static_cast<long long>(t) * p.n);
      const int c = pr.lane_idx ? pr.lane_idx[j] : j;
      const long long os = t * p.ld_smp + c;
      float* z = srow + threadIdx.x * stride;
      float lse[SRL_MAX_HEADS], hent[SRL_MAX_HEADS];
      int act[SRL_MAX_HEADS];
      float logp = 0.f, ent = 0.f;
      int off = 0;
#pragma unroll
      for (int hd = 0; hd < SRL_MAX_HEADS; ++hd) {
        if (hd < q.heads) {
          const int K = q.head_size[hd];
          float mx = -INFINITY;
          for (int k = 0; k < K; ++k) mx = fmaxf(mx, z[off + k]);
          float se = 0.f;
          for (int k = 0; k < K; ++k) se += expf(z[off + k] - mx);
          const float l = mx + logf(se);
          float hh = 0.f;
          for (int k = 0; k < K; ++k) {
            const float lp = fmaxf(z[off + k] - l, -FLT_MAX);  // Categorical.entropy clamps at finfo.min
            hh -= expf(lp) * lp;
          }
          const int a = q.action[i * q.heads + hd];
          lse[hd] = l;
          hent[hd] = hh;
          act[hd] = a;
          logp += z[off + a] - l;
          ent += hh;
          off += K;
        }
      }
      const float vp = ldg_stream(pr.v_pred + i);
      const float ol = __ldg(p.old_logp + os), rt = __ldg(p.ret + os), ad = __ldg(p.adv + os);
      const float ov = h.clip_value ? __ldg(p.old_value + os) : 0.f;
      const bool valid = __ldg(p.reset_next + os) == 0;
      float g_lp, g_v, g_en;
      RowSums rs;
      element(h, u, logp, vp, ent, ol, ov, rt, ad, valid, g_lp, g_v, g_en, rs);
      acc.add(rs);
      stg_stream(pr.g_value + i, g_v);
      if (q.logp_out) q.logp_out[i] = logp;
      if (q.entropy_out) q.entropy_out[i] = ent;
      off = 0;
#pragma unroll
      for (int hd = 0; hd < SRL_MAX_HEADS; ++hd) {
        if (hd < q.heads) {
          const int K = q.head_size[hd];
          for (int k = 0; k < K; ++k) {
            const float lp = z[off + k] - lse[hd];
            const float pk = expf(lp);
            // d logp / d z_k = [k == a] - p_k ;  d H / d z_k = -p_k (lp_k + H)
            z[off + k] = g_lp * ((k == act[hd] ? 1.f : 0.f) - pk) - g_en * pk * (fmaxf(lp, -FLT_MAX) + hent[hd]);
          }
          off += K;
        }
      }
    }
    __syncthreads();
    float* dst = q.g_logits + i0 * SK;
    for (int e = threadIdx.x; e < cnt * SK; e += 256) {
      const int r = e / SK;
      stg_stream(dst + e, srow[r * stride + (e - r * SK)]);
    }
    __syncthreads();
  }
  reduce_and_finalize(pr, h, acc, raw.m_local, blockIdx.x, gridDim.x);
}

int fill_loss_hyper(const srl_ppo_hyper* hyper, const double* popart_mean_std, LossHyperDev& h) {
  SRL_REQUIRE(hyper != nullptr, SRL_ERR_INVALID_ARG, "ppo loss: null hyper pointer");
  SRL_REQUIRE(hyper->value_loss >= SRL_VL_MSE && hyper->value_loss <= SRL_VL_SMOOTHL1, SRL_ERR_INVALID_ARG,
              "ppo loss: unknown value_loss %d (0 mse, 1 huber, 2 smoothl1)", hyper->value_loss);
  SRL_REQUIRE(!(hyper->normalize_old_value && popart_mean_std == nullptr), SRL_ERR_INVALID_ARG,
              "ppo loss: normalize_old_value needs popart statistics");
  h.clip_lo = static_cast<float>(1.0 - hyper->eps_clip);  // python: 1 - self.eps_clip, then cast by torch.clamp
  h.clip_hi = static_cast<float>(1.0 + hyper->eps_clip);
  h.veps = static_cast<float>(hyper->value_eps_clip);
  h.c_clip = static_cast<float>(hyper->c_clip);
  h.wv = static_cast<float>(hyper->value_loss_weight);
  h.we = static_cast<float>(hyper->entropy_bonus_weight);
  h.vl_param = static_cast<float>(hyper->vl_param);
  h.adv_eps = hyper->adv_eps;
  h.value_loss = hyper->value_loss;
  h.clip_value = hyper->clip_value;
  h.dual_clip = hyper->dual_clip;
  h.normalize_old_value = hyper->normalize_old_value;
  return SRL_OK;
}

template <int LANES, int MODE, int NB>
struct LossLauncher {
  // CTAs of `threads` threads one SM holds (register-limited), asked once per device and block size
  static int resident(int threads) {
    static int cached[64][2] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 2;
    int& c = cached[dev][threads >= 256 ? 1 : 0];
    if (c == 0) {
      int n = 0;
      if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&n, ppo_loss_kernel<LANES, MODE, NB>, threads, 0) != cudaSuccess ||
          n < 1)
        n = 1;
      c = n;
    }
    return c;
  }

  static int launch(LossBatch<NB>& b, int n_problems, cudaStream_t st) {
    LossShared& s = b.s;
    const int per_row = (s.n + LANES - 1) / LANES;  // threads one row needs
    const int threads = per_row <= 128 ? 128 : 256;
    s.col_tiles = (per_row + threads - 1) / threads;
    const long long capacity = static_cast<long long>(sm_count()) * resident(threads);
    // rows per tile: the tallest tile (index loads amortised over more rows) whose tile count still fills the
    // machine about as evenly as the best choice does
    int best_rows = 1;
    double best_eff = -1.0;
    for (int rows = 8; rows >= 1; rows >>= 1) {
      const long long tiles = static_cast<long long>(s.col_tiles) * ((s.T + rows - 1) / rows) * n_problems;
      const long long waves = (tiles + capacity - 1) / capacity;
      const double eff = static_cast<double>(tiles) / static_cast<double>(waves * capacity);
      if (eff > best_eff + 0.05) {
        best_eff = eff;
        best_rows = rows;
      }
    }
    s.rows_per_tile = best_rows;
    const long long tiles_pp = static_cast<long long>(s.col_tiles) * ((s.T + best_rows - 1) / best_rows);
    SRL_REQUIRE(tiles_pp < (1ll << 31), SRL_ERR_UNSUPPORTED, "ppo loss: problem too large (%lld tiles)", tiles_pp);
    s.n_tiles = static_cast<int>(tiles_pp);
    long long gx = (capacity + n_problems - 1) / n_problems;  // CTAs per problem when the grid is persistent
    if (gx > tiles_pp) gx = tiles_pp;
    if (gx > kMaxGrid) gx = kMaxGrid;
    if (gx < 1) gx = 1;
    const dim3 grid(static_cast<unsigned>(gx), static_cast<unsigned>(n_problems));
    ppo_loss_kernel<LANES, MODE, NB><<<grid, threads, 0, st>>>(b);
    SRL_CUDA(cudaGetLastError());
    return SRL_OK;
  }
};

template <int NB>
int launch_batch(LossBatch<NB>& b, int n_problems, bool lanes4, int mode, cudaStream_t st) {
  if (lanes4) {
    if (mode == kDense) return LossLauncher<4, kDense, NB>::launch(b, n_problems, st);
    if (mode == kGather) return LossLauncher<4, kGather, NB>::launch(b, n_problems, st);
    return LossLauncher<4, kPack, NB>::launch(b, n_problems, st);
  }
  if (mode == kPack) return LossLauncher<1, kPack, NB>::launch(b, n_problems, st);
  return LossLauncher<1, kGather, NB>::launch(b, n_problems, st);
}

}  // namespace
}  // namespace srl

extern "C" size_t srl_ppo_loss_workspace_bytes(int, int) {
  return srl::kPartialsOffset + static_cast<size_t>(srl::kMaxGrid) * srl::kNumSums * sizeof(double);
}

extern "C" int srl_ppo_loss_finalize(const void* workspace, size_t slot_bytes, int n_slots, double* out, float* out_f32,
                                     srl_stream_t stream) {
  using namespace srl;
  SRL_REQUIRE(n_slots >= 0, SRL_ERR_INVALID_ARG, "srl_ppo_loss_finalize: negative slot count");
  if (n_slots == 0) return SRL_OK;
  SRL_REQUIRE(workspace && out, SRL_ERR_INVALID_ARG, "srl_ppo_loss_finalize: null pointer");
  SRL_REQUIRE(slot_bytes >= srl_ppo_loss_workspace_bytes(1, 1) && slot_bytes % 8 == 0, SRL_ERR_INVALID_ARG,
              "srl_ppo_loss_finalize: slot_bytes=%zu smaller than one workspace or not 8-byte aligned", slot_bytes);
  loss_finalize_kernel<<<n_slots, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<const unsigned char*>(workspace), slot_bytes, out, out_f32);
  SRL_CUDA(cudaGetLastError());
  return SRL_OK;
}

extern "C" int srl_ppo_loss_fwd_bwd_batched(const srl_loss_problem* problems, int n_problems, int64_t ld_pol,
                                            int64_t ld_grad, const float* old_logp, const float* old_value,
                                            const float* ret, const float* adv, const uint8_t* on_reset_next,
                                            int64_t ld_smp, const float* pack, int T, int n,
                                            const double* popart_mean_std, const srl_ppo_hyper* hyper,
                                            size_t workspace_bytes, srl_stream_t stream) {
  using namespace srl;
  SRL_REQUIRE(n_problems >= 0, SRL_ERR_INVALID_ARG, "srl_ppo_loss_fwd_bwd_batched: negative problem count");
  if (n_problems == 0) return SRL_OK;
  SRL_REQUIRE(problems != nullptr, SRL_ERR_INVALID_ARG, "srl_ppo_loss_fwd_bwd_batched: null problem table");
  SRL_REQUIRE(T >= 1 && n >= 1, SRL_ERR_INVALID_ARG, "srl_ppo_loss_fwd_bwd: need T >= 1 and n >= 1 (got %d, %d)", T, n);
  SRL_REQUIRE(workspace_bytes >= srl_ppo_loss_workspace_bytes(T, n), SRL_ERR_INVALID_ARG,
              "srl_ppo_loss_fwd_bwd: workspace too small (%zu bytes)", workspace_bytes);
  LossShared s;
  int rc = fill_loss_hyper(hyper, popart_mean_std, s.h);
  if (rc != SRL_OK) return rc;
  if (pack == nullptr) {
    SRL_REQUIRE(old_logp && ret && adv && on_reset_next, SRL_ERR_INVALID_ARG, "srl_ppo_loss_fwd_bwd: null pointer");
    SRL_REQUIRE(!(s.h.clip_value && old_value == nullptr), SRL_ERR_INVALID_ARG,
                "srl_ppo_loss_fwd_bwd: clip_value needs old_value");
  } else {
    SRL_REQUIRE(aligned(pack, 16), SRL_ERR_INVALID_ARG, "srl_ppo_loss_fwd_bwd: pack must be 16-byte aligned");
  }
  bool any_idx = false, all_idx = true;
  bool dense_ok = (n % 4 == 0) && (ld_pol % 4 == 0) && (ld_grad % 4 == 0);
  for (int k = 0; k < n_problems; ++k) {
    const srl_loss_problem& q = problems[k];
    SRL_REQUIRE(q.new_logp && q.v_pred && q.entropy && q.g_logp && q.g_value && q.g_entropy && q.norm_stats &&
                    q.local_stats && q.workspace,
                SRL_ERR_INVALID_ARG, "srl_ppo_loss_fwd_bwd: null pointer");
    any_idx = any_idx || q.lane_idx != nullptr;
    all_idx = all_idx && q.lane_idx != nullptr;
    dense_ok = dense_ok && aligned(q.new_logp, 16) && aligned(q.v_pred, 16) && aligned(q.entropy, 16) &&
               aligned(q.g_logp, 16) && aligned(q.g_value, 16) && aligned(q.g_entropy, 16) &&
               (q.lane_idx == nullptr || aligned(q.lane_idx, 16));
  }
  SRL_REQUIRE(any_idx == all_idx, SRL_ERR_INVALID_ARG,
              "srl_ppo_loss_fwd_bwd_batched: either every problem has a lane_idx or none has");
  SRL_REQUIRE(ld_pol >= n && ld_grad >= n && ld_smp >= (any_idx ? 1 : n), SRL_ERR_INVALID_ARG,
              "srl_ppo_loss_fwd_bwd: row strides smaller than the row");
  s.old_logp = old_logp;
  s.old_value = old_value;
  s.ret = ret;
  s.adv = adv;
  s.reset_next = on_reset_next;
  s.pack = reinterpret_cast<const float4*>(pack);
  s.popart = popart_mean_std;
  s.ld_pol = ld_pol;
  s.ld_grad = ld_grad;
  s.ld_smp = ld_smp;
  s.T = T;
  s.n = n;
  const bool smp_vec = pack == nullptr && (ld_smp % 4 == 0) && aligned(old_logp, 16) && aligned(ret, 16) &&
                       aligned(adv, 16) && (!s.h.clip_value || aligned(old_value, 16)) && aligned(on_reset_next, 4);
  s.smp_vec_ok = smp_vec ? 1 : 0;
  const int mode = pack ? kPack : ((!any_idx && dense_ok && smp_vec) ? kDense : kGather);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  auto fill = [](Problem& p, const srl_loss_problem& q) {
    p.new_logp = q.new_logp;
    p.v_pred = q.v_pred;
    p.entropy = q.entropy;
    p.lane_idx = q.lane_idx;
    p.norm_stats = q.norm_stats;
    p.local_stats = q.local_stats;
    p.g_logp = q.g_logp;
    p.g_value = q.g_value;
    p.g_entropy = q.g_entropy;
    p.out = q.out;
    p.out_f32 = q.out_f32;
    p.slot = reinterpret_cast<SlotHeader*>(q.workspace);
  };
  for (int k0 = 0; k0 < n_problems;) {
    const int nb = n_problems - k0 < SRL_MAX_LOSS_BATCH ? n_problems - k0 : SRL_MAX_LOSS_BATCH;
    if (nb == 1) {
      LossBatch<1> b;
      b.s = s;
      fill(b.prob[0], problems[k0]);
      rc = launch_batch<1>(b, 1, dense_ok, mode, st);
    } else {
      LossBatch<SRL_MAX_LOSS_BATCH> b;
      b.s = s;
      for (int k = 0; k < nb; ++k) fill(b.prob[k], problems[k0 + k]);
      for (int k = nb; k < SRL_MAX_LOSS_BATCH; ++k) b.prob[k] = b.prob[0];
      rc = launch_batch<SRL_MAX_LOSS_BATCH>(b, nb, dense_ok, mode, st);
    }
    if (rc != SRL_OK) return rc;
    k0 += nb;
  }
  return SRL_OK;
}

extern "C" int srl_ppo_loss_fwd_bwd(const float* new_logp, const float* v_pred, const float* entropy, int64_t ld_pol,
                                    const float* old_logp, const float* old_value, const float* ret,
                                    const float* adv, const uint8_t* on_reset_next, int64_t ld_smp,
                                    const int32_t* lane_idx, int T, int n, const double* norm_stats,
                                    const double* local_stats, const double* popart_mean_std,
                                    const srl_ppo_hyper* hyper, float* g_logp, float* g_value, float* g_entropy,
                                    int64_t ld_grad, double* out, float* out_f32, void* workspace,
                                    size_t workspace_bytes, srl_stream_t stream) {
  srl_loss_problem q;
  q.new_logp = new_logp;
  q.v_pred = v_pred;
  q.entropy = entropy;
  q.lane_idx = lane_idx;
  q.norm_stats = norm_stats;
  q.local_stats = local_stats;
  q.g_logp = g_logp;
  q.g_value = g_value;
  q.g_entropy = g_entropy;
  q.out = out;
  q.out_f32 = out_f32;
  q.workspace = workspace;
  return srl_ppo_loss_fwd_bwd_batched(&q, 1, ld_pol, ld_grad, old_logp, old_value, ret, adv, on_reset_next, ld_smp,
                                      nullptr, T, n, popart_mean_std, hyper, workspace_bytes, stream);
}

extern "C" int srl_ppo_loss_from_logits(const float* logits, const int32_t* action, const int32_t* head_sizes_host,
                                        int heads, const float* v_pred, const float* old_logp,
                                        const float* old_value, const float* ret, const float* adv,
                                        const uint8_t* on_reset_next, int64_t ld_smp, const int32_t* lane_idx, int T,
                                        int n, const double* norm_stats, const double* local_stats,
                                        const double* popart_mean_std, const srl_ppo_hyper* hyper, float* g_logits,
                                        float* g_value, float* logp_out, float* entropy_out, double* out,
                                        float* out_f32, void* workspace, size_t workspace_bytes,
                                        srl_stream_t stream) {
  using namespace srl;
  SRL_REQUIRE(T >= 1 && n >= 1, SRL_ERR_INVALID_ARG, "srl_ppo_loss_from_logits: need T >= 1 and n >= 1");
  SRL_REQUIRE(heads >= 1 && heads <= SRL_MAX_HEADS && head_sizes_host, SRL_ERR_INVALID_ARG,
              "srl_ppo_loss_from_logits: heads=%d outside [1, %d]", heads, SRL_MAX_HEADS);
  SRL_REQUIRE(logits && action && v_pred && old_logp && ret && adv && on_reset_next && g_logits && g_value &&
                  norm_stats && local_stats && workspace,
              SRL_ERR_INVALID_ARG, "srl_ppo_loss_from_logits: null pointer");
  SRL_REQUIRE(ld_smp >= (lane_idx ? 1 : n), SRL_ERR_INVALID_ARG, "srl_ppo_loss_from_logits: ld_smp smaller than the row");
  SRL_REQUIRE(workspace_bytes >= srl_ppo_loss_workspace_bytes(T, n), SRL_ERR_INVALID_ARG,
              "srl_ppo_loss_from_logits: workspace too small (%zu bytes)", workspace_bytes);
  LogitsParams q;
  int rc = fill_loss_hyper(hyper, popart_mean_std, q.s.h);
  if (rc != SRL_OK) return rc;
  SRL_REQUIRE(!(q.s.h.clip_value && old_value == nullptr), SRL_ERR_INVALID_ARG,
              "srl_ppo_loss_from_logits: clip_value needs old_value");
  int sk = 0;
  for (int i = 0; i < SRL_MAX_HEADS; ++i) {
    q.head_size[i] = i < heads ? head_sizes_host[i] : 0;
    SRL_REQUIRE(i >= heads || head_sizes_host[i] >= 1, SRL_ERR_INVALID_ARG,
                "srl_ppo_loss_from_logits: head %d has no actions", i);
    sk += q.head_size[i];
  }
  const size_t smem = static_cast<size_t>(256) * (sk + 1) * sizeof(float);
  SRL_REQUIRE(smem <= 200 * 1024, SRL_ERR_UNSUPPORTED, "srl_ppo_loss_from_logits: sum K = %d too wide (max 199)", sk);
  q.heads = heads;
  q.SK = sk;
  q.logits = logits;
  q.action = action;
  q.g_logits = g_logits;
  q.logp_out = logp_out;
  q.entropy_out = entropy_out;
  LossShared& p = q.s;
  p.old_logp = old_logp;
  p.old_value = old_value;
  p.ret = ret;
  p.adv = adv;
  p.reset_next = on_reset_next;
  p.pack = nullptr;
  p.popart = popart_mean_std;
  p.ld_pol = n;
  p.ld_grad = n;
  p.ld_smp = ld_smp;
  p.T = T;
  p.n = n;
  p.rows_per_tile = p.col_tiles = p.n_tiles = 0;
  p.smp_vec_ok = 0;
  Problem& pr = q.pr;
  pr.new_logp = pr.entropy = nullptr;
  pr.v_pred = v_pred;
  pr.lane_idx = lane_idx;
  pr.norm_stats = norm_stats;
  pr.local_stats = local_stats;
  pr.g_logp = pr.g_entropy = nullptr;
  pr.g_value = g_value;
  pr.out = out;
  pr.out_f32 = out_f32;
  pr.slot = reinterpret_cast<SlotHeader*>(workspace);
  static bool opted_in[64] = {};
  int dev = 0;
  SRL_CUDA(cudaGetDevice(&dev));
  if (dev < 64 && !opted_in[dev]) {
    SRL_CUDA(cudaFuncSetAttribute(ppo_loss_logits_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    opted_in[dev] = true;
  }
  const long long tiles = (static_cast<long long>(T) * n + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 8 < kMaxGrid ? static_cast<long long>(sm_count()) * 8 : kMaxGrid;
  const int grid = static_cast<int>(tiles < cap ? tiles : cap);
  ppo_loss_logits_kernel<<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(q);
  SRL_CUDA(cudaGetLastError());
  return SRL_OK;
}
