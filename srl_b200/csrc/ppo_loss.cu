// K4: fused PPO / MAPPO loss forward + backward (see include/srl_b200.h; replaces mappo.py:146-217,
// utils.py:10-67,228-265 and the autograd backward to new_logp / state_values / entropy).
//
// One pass over the [T, n] minibatch: each element reads 29 B (25 B without value clipping) and
// writes 12 B of gradients.  The advantage-normalisation statistics arrive pre-reduced (and, on
// several GPUs, pre-all-reduced) from K2 + srl_group_stats, so there is no second pass and no grid
// barrier.  Eight masked sums are reduced warp-shuffle -> shared memory -> one float64 partial row per
// CTA in the caller's workspace slot.  Two ways to turn the rows into the loss scalars + stats vector:
//   immediate (out != NULL): the last CTA to finish (atomic ticket) folds the rows in a fixed order;
//   deferred  (out == NULL): the kernel ends right after its stores -- the gradients are all the backward
//     pass needs -- and srl_ppo_loss_finalize folds any number of slots later in ONE launch (the reference
//     pays eleven .item() syncs per epoch for the same numbers, mappo.py:293-299).
// Either way the result is deterministic for a given launch shape and nothing syncs with the host.
#include <float.h>

#include "common.cuh"

namespace srl {
namespace {

constexpr int kMaxGrid = 2048;
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

struct LossParams {
  const float* new_logp;
  const float* v_pred;
  const float* entropy;
  long long ld_pol;
  const float* old_logp;
  const float* old_value;
  const float* ret;
  const float* adv;
  const uint8_t* reset_next;
  long long ld_smp;
  const int32_t* lane_idx;
  int T, n;
  int smp_vec_ok;  // sample-side leaves allow 128-bit loads at 4-aligned lanes (alignment + row stride)
  const double* norm_stats;
  const double* local_stats;
  const double* popart;
  float* g_logp;
  float* g_value;
  float* g_entropy;
  long long ld_grad;
  double* out;
  float* out_f32;
  double* partials;
  SlotHeader* slot;
  LossHyperDev h;
};

// torch.nn.{MSELoss,HuberLoss,SmoothL1Loss}(reduction='none') value and derivative wrt the input.
__device__ __forceinline__ void pointwise_loss(int kind, float prm, float d, float& l, float& dl) {
  if (kind == SRL_VL_MSE) {
    l = d * d;
    dl = 2.f * d;
  } else if (kind == SRL_VL_HUBER) {
    const float z = fabsf(d);
    if (z < prm) {
      l = 0.5f * z * z;
      dl = d;
    } else {
      l = prm * (z - 0.5f * prm);
      dl = d > 0.f ? prm : -prm;
    }
  } else {
    const float z = fabsf(d);
    if (z < prm) {
      l = 0.5f * z * z / prm;
      dl = d / prm;
    } else {
      l = z - 0.5f * prm;
      dl = d > 0.f ? 1.f : -1.f;
    }
  }
}

struct Uniforms {
  double mean, denom;   // advantage normalisation: (x - mean) / denom
  double pa_mu, pa_sd;  // popart
  float inv_m;          // 1 / local sum(mask)
  bool popart;
};

struct Acc {
  double pl = 0, vl = 0, en = 0, adv = 0, ratio = 0, clip = 0, vt = 0, ret = 0;
};

// The six scalars every thread needs.  Loading them is split from the math on them so that the (long) fp64
// divide / sqrt latency overlaps the element loads instead of preceding them.
struct RawStats {
  double cnt, s1, s2, m_local, pa_mu, pa_sd;
  bool popart;
};

__device__ __forceinline__ RawStats load_raw_stats(const LossParams& p) {
  RawStats r;
  r.cnt = __ldg(p.norm_stats);
  r.s1 = __ldg(p.norm_stats + 1);
  r.s2 = __ldg(p.norm_stats + 2);
  r.m_local = __ldg(p.local_stats);
  r.popart = p.popart != nullptr;
  r.pa_mu = r.popart ? __ldg(p.popart) : 0.0;
  r.pa_sd = r.popart ? __ldg(p.popart + 1) : 1.0;
  return r;
}

__device__ __forceinline__ Uniforms make_uniforms(const RawStats& r, double adv_eps) {
  Uniforms u;
  u.popart = r.popart;
  u.pa_mu = r.pa_mu;
  u.pa_sd = r.pa_sd;
  u.mean = r.s1 / r.cnt;
  const double var = r.s2 / r.cnt - u.mean * u.mean;  // biased variance, utils.py:62-64
  u.denom = sqrt(var) + adv_eps;                      // eps outside the sqrt, utils.py:67
  u.inv_m = 1.f / static_cast<float>(r.m_local);
  return u;
}

__device__ __forceinline__ float popart_normalize(float x, const Uniforms& u) {
  // RunningMeanStd.normalize: ((x.double() - mean) / std).clip(-5, 5).float()   utils.py:139-144
  double z = (static_cast<double>(x) - u.pa_mu) / u.pa_sd;
  z = fmin(fmax(z, -5.0), 5.0);
  return static_cast<float>(z);
}

__device__ __forceinline__ void element(const LossHyperDev& h, const Uniforms& u, float nl, float vp, float en,
                                        float ol, float ov, float rt, float ad, uint32_t rs, float& g_lp,
                                        float& g_v, float& g_en, Acc& acc) {
  const bool valid = (rs == 0);
  const float mk = valid ? 1.f : 0.f;
  const float scale = mk * u.inv_m;  // d(masked mean)/d(element) = mask / M

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
  g_v = h.wv * scale * gv;

  // ---- actor: mappo.py:157-158,186-197 ---------------------------------------------------------
  const float ratio = expf(nl - ol);
  const double x = static_cast<double>(ad) * static_cast<double>(mk);  // masked BEFORE centring (utils.py:54)
  const float nadv = static_cast<float>((x - u.mean) / u.denom);
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
  g_lp = -scale * gsum;
  g_en = -h.we * scale;  // entropy_loss = -sum(entropy * mask) / M   mappo.py:199

  if (valid) {
    acc.pl += static_cast<double>(-obj);
    acc.vl += static_cast<double>(vl);
    acc.en += static_cast<double>(en);
    acc.adv += static_cast<double>(ad);
    acc.ratio += static_cast<double>(ratio);
    acc.clip += (s2 < s1) ? 1.0 : 0.0;
    acc.vt += static_cast<double>(vt);
    acc.ret += static_cast<double>(rt);
  }
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
__device__ __forceinline__ void reduce_and_finalize(const LossParams& p, const Acc& acc, double mask_sum) {
  const LossHyperDev& h = p.h;
  __shared__ double sred[kNumSums][8];
  __shared__ bool is_last;
  double v[kNumSums] = {acc.pl, acc.vl, acc.en, acc.adv, acc.ratio, acc.clip, acc.vt, acc.ret};
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
    p.partials[static_cast<size_t>(blockIdx.x) * kNumSums + threadIdx.x] = s;
  }
  if (p.out == nullptr) {  // deferred: publish what the finaliser needs and leave
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      p.slot->n_rows = gridDim.x;
      p.slot->mask_sum = mask_sum;
      p.slot->wv = static_cast<double>(h.wv);
      p.slot->we = static_cast<double>(h.we);
    }
    return;
  }
  if (threadIdx.x < kNumSums) __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int done = atomicAdd(&p.slot->ticket, 1u);
    is_last = (done == gridDim.x - 1);
  }
  __syncthreads();
  if (!is_last) return;
  __threadfence();
  fold_rows_and_write(p.partials, static_cast<int>(gridDim.x), mask_sum, static_cast<double>(h.wv),
                      static_cast<double>(h.we), sred, p.out, p.out_f32);
  if (threadIdx.x == 0) p.slot->ticket = 0u;  // ready for the next launch on this slot
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

template <bool VEC4>
__global__ void __launch_bounds__(256) ppo_loss_kernel(const LossParams p) {
  const LossHyperDev& h = p.h;
#ifdef SRL_DEBUG_PHASES
  const long long dbg0 = clock64();
#endif
  const RawStats raw = load_raw_stats(p);  // loads only; the math on them runs under the element loads
  Uniforms u;
  bool have_u = false;
  const double mask_sum = raw.m_local;
#ifdef SRL_DEBUG_PHASES
  const long long dbg1 = clock64();
#endif
  Acc acc;
  const int n = p.n, T = p.T;
  if (VEC4) {
    // four consecutive lanes per thread: 128-bit loads/stores on the dense policy side; the sample side is either
    // dense too (128-bit) or gathered through lane_idx (four independent 32-bit gathers per leaf, all in flight)
    const int n4 = n >> 2;
    const long long W = static_cast<long long>(T) * n4;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < W;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      const int t = static_cast<int>(i / n4);
      const int j = static_cast<int>(i - static_cast<long long>(t) * n4) << 2;
      const long long op = t * p.ld_pol + j, og = t * p.ld_grad + j;
      int4 c = make_int4(j, j + 1, j + 2, j + 3);
      if (p.lane_idx) c = __ldg(reinterpret_cast<const int4*>(p.lane_idx + j));
      const float4 nl = ldg_stream(reinterpret_cast<const float4*>(p.new_logp + op));
      const float4 vp = ldg_stream(reinterpret_cast<const float4*>(p.v_pred + op));
      const float4 en = ldg_stream(reinterpret_cast<const float4*>(p.entropy + op));
      float4 ol, rt, ad, ov = make_float4(0.f, 0.f, 0.f, 0.f);
      uint32_t rs;
      const long long ob = t * p.ld_smp;
      // a run of four consecutive, 4-aligned lanes (dense batches, or permutations of >= 4-lane blocks: agents of
      // one environment, sector-aligned environment blocks) is fetched with one 128-bit load per leaf
      const bool run4 = (c.y == c.x + 1) && (c.z == c.x + 2) && (c.w == c.x + 3) && ((c.x & 3) == 0) && p.smp_vec_ok;
      if (run4) {
        const long long os = ob + c.x;
        ol = ldg_stream(reinterpret_cast<const float4*>(p.old_logp + os));
        rt = ldg_stream(reinterpret_cast<const float4*>(p.ret + os));
        ad = ldg_stream(reinterpret_cast<const float4*>(p.adv + os));
        if (h.clip_value) ov = ldg_stream(reinterpret_cast<const float4*>(p.old_value + os));
        rs = ldg_stream(reinterpret_cast<const uint32_t*>(p.reset_next + os));
      } else {
        ol = make_float4(__ldg(p.old_logp + ob + c.x), __ldg(p.old_logp + ob + c.y), __ldg(p.old_logp + ob + c.z),
                         __ldg(p.old_logp + ob + c.w));
        rt = make_float4(__ldg(p.ret + ob + c.x), __ldg(p.ret + ob + c.y), __ldg(p.ret + ob + c.z), __ldg(p.ret + ob + c.w));
        ad = make_float4(__ldg(p.adv + ob + c.x), __ldg(p.adv + ob + c.y), __ldg(p.adv + ob + c.z), __ldg(p.adv + ob + c.w));
        if (h.clip_value)
          ov = make_float4(__ldg(p.old_value + ob + c.x), __ldg(p.old_value + ob + c.y), __ldg(p.old_value + ob + c.z),
                           __ldg(p.old_value + ob + c.w));
        rs = static_cast<uint32_t>(__ldg(p.reset_next + ob + c.x)) | (static_cast<uint32_t>(__ldg(p.reset_next + ob + c.y)) << 8) |
             (static_cast<uint32_t>(__ldg(p.reset_next + ob + c.z)) << 16) |
             (static_cast<uint32_t>(__ldg(p.reset_next + ob + c.w)) << 24);
      }
      if (!have_u) {
        u = make_uniforms(raw, h.adv_eps);
        have_u = true;
      }
      float4 glp, gv, ge;
      element(h, u, nl.x, vp.x, en.x, ol.x, ov.x, rt.x, ad.x, rs & 0xffu, glp.x, gv.x, ge.x, acc);
      element(h, u, nl.y, vp.y, en.y, ol.y, ov.y, rt.y, ad.y, (rs >> 8) & 0xffu, glp.y, gv.y, ge.y, acc);
      element(h, u, nl.z, vp.z, en.z, ol.z, ov.z, rt.z, ad.z, (rs >> 16) & 0xffu, glp.z, gv.z, ge.z, acc);
      element(h, u, nl.w, vp.w, en.w, ol.w, ov.w, rt.w, ad.w, (rs >> 24) & 0xffu, glp.w, gv.w, ge.w, acc);
      stg_stream(reinterpret_cast<float4*>(p.g_logp + og), glp);
      stg_stream(reinterpret_cast<float4*>(p.g_value + og), gv);
      stg_stream(reinterpret_cast<float4*>(p.g_entropy + og), ge);
    }
  } else {
    const long long W = static_cast<long long>(T) * n;
    for (long long i = blockIdx.x * static_cast<long long>(blockDim.x) + threadIdx.x; i < W;
         i += static_cast<long long>(gridDim.x) * blockDim.x) {
      const int t = static_cast<int>(i / n);
      const int j = static_cast<int>(i - static_cast<long long>(t) * n);
      const int c = p.lane_idx ? __ldg(p.lane_idx + j) : j;  // minibatch gather fused into the load
      const long long op = t * p.ld_pol + j, os = t * p.ld_smp + c, og = t * p.ld_grad + j;
      const float nl = ldg_stream(p.new_logp + op), vp = ldg_stream(p.v_pred + op), en = ldg_stream(p.entropy + op);
      const float ol = __ldg(p.old_logp + os), rt = __ldg(p.ret + os), ad = __ldg(p.adv + os);
      const float ov = h.clip_value ? __ldg(p.old_value + os) : 0.f;
      const uint32_t rs = __ldg(p.reset_next + os);
      if (!have_u) {
        u = make_uniforms(raw, h.adv_eps);
        have_u = true;
      }
      float glp, gv, ge;
      element(h, u, nl, vp, en, ol, ov, rt, ad, rs, glp, gv, ge, acc);
      stg_stream(p.g_logp + og, glp);
      stg_stream(p.g_value + og, gv);
      stg_stream(p.g_entropy + og, ge);
    }
  }
#ifdef SRL_DEBUG_PHASES
  const long long dbg2 = clock64() + (acc.pl == 123.0);
#endif
  reduce_and_finalize(p, acc, mask_sum);
#ifdef SRL_DEBUG_PHASES
  if (threadIdx.x == 0 && (blockIdx.x == 0 || blockIdx.x == gridDim.x - 1))
    printf("loss block %d/%d (%d thr): uniforms %lld  elements %lld  reduce %lld cycles\n", blockIdx.x, gridDim.x, blockDim.x,
           dbg1 - dbg0, dbg2 - dbg1, clock64() - dbg2);
#endif
}


// ---- K4b: the same loss starting from the actor head's logits -----------------------------------
// (actor_critic_policy.py:303-324: Categorical(logits=slice).log_prob / .entropy per head, summed.)
// A CTA takes 256 consecutive transitions; their [256, sumK] logits block is contiguous in memory, so
// it is staged through shared memory with coalesced loads (row stride sumK+1 words: conflict-free
// per-thread row walks), turned into d loss / d logits in place, and streamed back out coalesced.
struct LogitsParams {
  LossParams c;           // policy-side pointers new_logp/entropy/g_logp/g_entropy are unused here
  const float* logits;    // [T*n, SK]
  const int32_t* action;  // [T*n, heads]
  float* g_logits;        // [T*n, SK]
  float* logp_out;        // [T*n] or null
  float* entropy_out;     // [T*n] or null
  int heads, SK;
  int head_size[SRL_MAX_HEADS];
};

__global__ void __launch_bounds__(256) ppo_loss_logits_kernel(const LogitsParams q) {
  extern __shared__ float srow[];  // [256][SK + 1]
  const LossParams& p = q.c;
  const LossHyperDev& h = p.h;
  const RawStats raw = load_raw_stats(p);
  const double mask_sum = raw.m_local;
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
      const int c = p.lane_idx ? p.lane_idx[j] : j;
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
      const float vp = ldg_stream(p.v_pred + i);
      const float ol = __ldg(p.old_logp + os), rt = __ldg(p.ret + os), ad = __ldg(p.adv + os);
      const float ov = h.clip_value ? __ldg(p.old_value + os) : 0.f;
      const uint32_t rs = __ldg(p.reset_next + os);
      float g_lp, g_v, g_en;
      element(h, u, logp, vp, ent, ol, ov, rt, ad, rs, g_lp, g_v, g_en, acc);
      stg_stream(p.g_value + i, g_v);
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
  reduce_and_finalize(p, acc, mask_sum);
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

namespace srl {
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
}  // namespace srl

extern "C" int srl_ppo_loss_fwd_bwd(const float* new_logp, const float* v_pred, const float* entropy, int64_t ld_pol,
                                    const float* old_logp, const float* old_value, const float* ret,
                                    const float* adv, const uint8_t* on_reset_next, int64_t ld_smp,
                                    const int32_t* lane_idx, int T, int n, const double* norm_stats,
                                    const double* local_stats, const double* popart_mean_std,
                                    const srl_ppo_hyper* hyper, float* g_logp, float* g_value, float* g_entropy,
                                    int64_t ld_grad, double* out, float* out_f32, void* workspace,
                                    size_t workspace_bytes, srl_stream_t stream) {
  using namespace srl;
  SRL_REQUIRE(T >= 1 && n >= 1, SRL_ERR_INVALID_ARG, "srl_ppo_loss_fwd_bwd: need T >= 1 and n >= 1 (got %d, %d)", T, n);
  SRL_REQUIRE(new_logp && v_pred && entropy && old_logp && ret && adv && on_reset_next && g_logp && g_value &&
                  g_entropy && norm_stats && local_stats && workspace,
              SRL_ERR_INVALID_ARG, "srl_ppo_loss_fwd_bwd: null pointer");
  SRL_REQUIRE(ld_pol >= n && ld_grad >= n && ld_smp >= (lane_idx ? 1 : n), SRL_ERR_INVALID_ARG,
              "srl_ppo_loss_fwd_bwd: row strides smaller than the row");
  SRL_REQUIRE(workspace_bytes >= srl_ppo_loss_workspace_bytes(T, n), SRL_ERR_INVALID_ARG,
              "srl_ppo_loss_fwd_bwd: workspace too small (%zu bytes)", workspace_bytes);
  LossParams p;
  int rc = fill_loss_hyper(hyper, popart_mean_std, p.h);
  if (rc != SRL_OK) return rc;
  SRL_REQUIRE(!(p.h.clip_value && old_value == nullptr), SRL_ERR_INVALID_ARG,
              "srl_ppo_loss_fwd_bwd: clip_value needs old_value");
  p.new_logp = new_logp;
  p.v_pred = v_pred;
  p.entropy = entropy;
  p.ld_pol = ld_pol;
  p.old_logp = old_logp;
  p.old_value = old_value;
  p.ret = ret;
  p.adv = adv;
  p.reset_next = on_reset_next;
  p.ld_smp = ld_smp;
  p.lane_idx = lane_idx;
  p.T = T;
  p.n = n;
  p.norm_stats = norm_stats;
  p.local_stats = local_stats;
  p.popart = popart_mean_std;
  p.g_logp = g_logp;
  p.g_value = g_value;
  p.g_entropy = g_entropy;
  p.ld_grad = ld_grad;
  p.out = out;
  p.out_f32 = out_f32;
  p.slot = reinterpret_cast<SlotHeader*>(workspace);
  p.partials = reinterpret_cast<double*>(static_cast<char*>(workspace) + kPartialsOffset);

  // four lanes per thread whenever the dense side allows 128-bit accesses; the sample side may be gathered
  const bool dense_ok = (n % 4 == 0) && (ld_pol % 4 == 0) && (ld_grad % 4 == 0) && aligned(new_logp, 16) &&
                        aligned(v_pred, 16) && aligned(entropy, 16) && aligned(g_logp, 16) && aligned(g_value, 16) &&
                        aligned(g_entropy, 16);
  const bool smp_vec = (ld_smp % 4 == 0) && aligned(old_logp, 16) && aligned(ret, 16) && aligned(adv, 16) &&
                       (!p.h.clip_value || aligned(old_value, 16)) && aligned(on_reset_next, 4);
  p.smp_vec_ok = smp_vec ? 1 : 0;
  const bool use_vec4 = dense_ok && (lane_idx == nullptr || aligned(lane_idx, 16));
  const int sms = sm_count();
  const long long W = static_cast<long long>(T) * (use_vec4 ? n / 4 : n);
  // The kernel is latency-bound at minibatch sizes (profiles/r1_notes.md): what matters is how many elements
  // are in flight per SM, so threads carry four elements and CTAs stay small enough for several launches to
  // share the machine when the caller runs them on parallel streams / graph branches.
  const int threads = (W <= static_cast<long long>(sms) * 256) ? 128 : 256;
  long long grid = (W + threads - 1) / threads;
  const long long cap = static_cast<long long>(sms) * 8 < kMaxGrid ? static_cast<long long>(sms) * 8 : kMaxGrid;
  if (grid > cap) grid = cap;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (use_vec4)
    ppo_loss_kernel<true><<<static_cast<int>(grid), threads, 0, st>>>(p);
  else
    ppo_loss_kernel<false><<<static_cast<int>(grid), threads, 0, st>>>(p);
  SRL_CUDA(cudaGetLastError());
  return SRL_OK;
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
  int rc = fill_loss_hyper(hyper, popart_mean_std, q.c.h);
  if (rc != SRL_OK) return rc;
  SRL_REQUIRE(!(q.c.h.clip_value && old_value == nullptr), SRL_ERR_INVALID_ARG,
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
  LossParams& p = q.c;
  p.new_logp = p.entropy = nullptr;
  p.v_pred = v_pred;
  p.ld_pol = n;
  p.old_logp = old_logp;
  p.old_value = old_value;
  p.ret = ret;
  p.adv = adv;
  p.reset_next = on_reset_next;
  p.ld_smp = ld_smp;
  p.lane_idx = lane_idx;
  p.T = T;
  p.n = n;
  p.norm_stats = norm_stats;
  p.local_stats = local_stats;
  p.popart = popart_mean_std;
  p.g_logp = p.g_entropy = nullptr;
  p.smp_vec_ok = 0;
  p.g_value = g_value;
  p.ld_grad = n;
  p.out = out;
  p.out_f32 = out_f32;
  p.slot = reinterpret_cast<SlotHeader*>(workspace);
  p.partials = reinterpret_cast<double*>(static_cast<char*>(workspace) + kPartialsOffset);
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
