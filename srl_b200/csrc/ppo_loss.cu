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
//    gathered through lane_idx (any alignment), or K2's `pack` (float4 = old_logp, old_value, ret, adv-or-NaN per
//    transition, the two rows of a row pair next to each other): one gather per lane and row pair instead of five
//    1..4-byte ones per transition -- permuted minibatches of even width run ppo_loss_pair.cu, this file's row-tile
//    kernel keeps the odd shapes of that form;
//  * per-element arithmetic is fp32 except the advantage / PopArt normalisation, which stays in float64
//    (reference: utils.py:38-67,139-144) as subtract + multiply by a correctly rounded reciprocal + one
//    Newton residual step (3 fp64 instructions instead of a ~25-instruction IEEE division; the quotient is
//    the correctly rounded one, then cast to float as the reference does);
//  * eight masked sums: four lanes are summed in fp32, then accumulated in float64 per thread, reduced
//    warp-shuffle -> shared memory -> one float64 partial row per CTA in the problem's workspace slot.
// Finalisation: immediate (out != NULL): the last CTA of the problem (atomic ticket) folds the rows in a fixed
// order; deferred (out == NULL): srl_ppo_loss_finalize folds any number of slots later in ONE launch.
// Either way the result is deterministic for a given launch shape and nothing syncs with the host.
#include <stdlib.h>
#include <string.h>

#include "ppo_loss.cuh"

namespace srl {
namespace loss {
namespace {

__global__ void __launch_bounds__(256) loss_finalize_kernel(unsigned char* __restrict__ ws, size_t slot_bytes,
                                                            double* __restrict__ out, float* __restrict__ out_f32) {
  __shared__ double sred[kNumSums][8];
  unsigned char* base = ws + static_cast<size_t>(blockIdx.x) * slot_bytes;
  const SlotHeader* hd = reinterpret_cast<const SlotHeader*>(base);
  fold_rows_and_write(reinterpret_cast<double*>(base + kPartialsOffset), static_cast<int>(hd->n_rows),
                      hd->mask_sum, hd->wv, hd->we, sred, out + static_cast<size_t>(blockIdx.x) * SRL_LOSS_OUT_LEN,
                      out_f32 ? out_f32 + static_cast<size_t>(blockIdx.x) * 4 : nullptr);
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

// CACHE_E: a second shared-memory row per transition keeps e_k = exp(z_k - max) of the log-sum-exp pass, so that the entropy
// and the gradient passes need no further exponentials (1 expf per logit instead of 3: the kernel is bound by the special
// function unit, not by memory -- 72 us for 90 MB at cfg2's shape before, profiles/r2_notes.md); used when both rows fit.
template <bool CACHE_E>
__global__ void __launch_bounds__(256) ppo_loss_logits_kernel(const __grid_constant__ LogitsParams q) {
  extern __shared__ float srow[];  // [256][SK + 1] (+ [256][SK + 1] with CACHE_E)
  const LossShared& p = q.s;
  const Problem& pr = q.pr;
  const LossHyperDev& h = p.h;
  double mask_sum;
  const Uniforms u = load_uniforms(pr.norm_stats, pr.local_stats, p.popart, h.adv_eps, mask_sum);
  Acc acc;
  const int SK = q.SK, stride = SK + 1;
  float* erow = srow + 256 * stride;
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
      float* ez = erow + threadIdx.x * stride;
      float lse[SRL_MAX_HEADS], hent[SRL_MAX_HEADS], rse[SRL_MAX_HEADS];
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
          for (int k = 0; k < K; ++k) {
            const float ek = expf(z[off + k] - mx);
            if (CACHE_E) ez[off + k] = ek;
            se += ek;
          }
          const float l = mx + logf(se);
          const float inv = 1.f / se;
          float hh = 0.f;
          for (int k = 0; k < K; ++k) {
            const float lp = fmaxf(z[off + k] - l, -FLT_MAX);  // Categorical.entropy clamps at finfo.min
            const float pk = CACHE_E ? ez[off + k] * inv : expf(lp);
            hh -= pk * lp;
          }
          const int a = q.action[i * q.heads + hd];
          lse[hd] = l;
          rse[hd] = inv;
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
      element<RuntimeCfg>(h, u, logp, vp, ent, ol, ov, rt, ad, valid, g_lp, g_v, g_en, rs);
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
            const float pk = CACHE_E ? ez[off + k] * rse[hd] : expf(lp);
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
  reduce_and_finalize(pr, h, acc, mask_sum, blockIdx.x, gridDim.x);
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

}  // namespace
}  // namespace loss
}  // namespace srl

extern "C" size_t srl_ppo_loss_workspace_bytes(int, int) {
  return srl::loss::kPartialsOffset + static_cast<size_t>(srl::loss::kMaxGrid) * srl::loss::kNumSums * sizeof(double);
}

extern "C" int srl_ppo_loss_finalize(void* workspace, size_t slot_bytes, int n_slots, double* out, float* out_f32,
                                     srl_stream_t stream) {
  using namespace srl;
  using namespace srl::loss;
  SRL_REQUIRE(n_slots >= 0, SRL_ERR_INVALID_ARG, "srl_ppo_loss_finalize: negative slot count");
  if (n_slots == 0) return SRL_OK;
  SRL_REQUIRE(workspace && out, SRL_ERR_INVALID_ARG, "srl_ppo_loss_finalize: null pointer");
  SRL_REQUIRE(slot_bytes >= srl_ppo_loss_workspace_bytes(1, 1) && slot_bytes % 8 == 0, SRL_ERR_INVALID_ARG,
              "srl_ppo_loss_finalize: slot_bytes=%zu smaller than one workspace or not 8-byte aligned", slot_bytes);
  loss_finalize_kernel<<<n_slots, 256, 0, static_cast<cudaStream_t>(stream)>>>(
      static_cast<unsigned char*>(workspace), slot_bytes, out, out_f32);
  SRL_CUDA(cudaGetLastError());
  return SRL_OK;
}

extern "C" int srl_ppo_loss_fwd_bwd_batched(const srl_loss_problem* problems, int n_problems, int64_t ld_pol,
                                            int64_t ld_grad, const float* old_logp, const float* old_value,
                                            const float* ret, const float* adv, const uint8_t* on_reset_next,
                                            int64_t ld_smp, const float* pack, int pack_row_lo, const double* lane_aos,
                                            int T, int n, const double* popart_mean_std, const srl_ppo_hyper* hyper,
                                            size_t workspace_bytes, srl_xchg* xchg, srl_stream_t stream) {
  using namespace srl;
  using namespace srl::loss;
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
    SRL_REQUIRE(aligned(pack, 32) && pack_row_lo >= 0, SRL_ERR_INVALID_ARG,
                "srl_ppo_loss_fwd_bwd: pack must be 32-byte aligned and pack_row_lo >= 0");
  }
  bool any_idx = false, all_idx = true;
  bool dense_ok = (n % 4 == 0) && (ld_pol % 4 == 0) && (ld_grad % 4 == 0);
  for (int k = 0; k < n_problems; ++k) {
    const srl_loss_problem& q = problems[k];
    SRL_REQUIRE(q.new_logp && q.v_pred && q.entropy && q.g_logp && q.g_value && q.g_entropy && q.workspace &&
                    (lane_aos != nullptr || (q.norm_stats && q.local_stats)),
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
  SRL_REQUIRE(lane_aos == nullptr || (popart_mean_std == nullptr && pack != nullptr && all_idx && aligned(lane_aos, 32)),
              SRL_ERR_INVALID_ARG,
              "srl_ppo_loss_fwd_bwd_batched: self-computed statistics (lane_aos) need the pack form with lane indices, no "
              "PopArt and a 32-byte aligned table");
  s.lane_aos = lane_aos;
  s.row_lo = pack ? pack_row_lo : 0;
  s.ld_pol = ld_pol;
  s.ld_grad = ld_grad;
  s.ld_smp = ld_smp;
  s.T = T;
  s.n = n;
  const bool smp_vec = pack == nullptr && (ld_smp % 4 == 0) && aligned(old_logp, 16) && aligned(ret, 16) &&
                       aligned(adv, 16) && (!s.h.clip_value || aligned(old_value, 16)) && aligned(on_reset_next, 4);
  s.smp_vec_ok = smp_vec ? 1 : 0;
  bool aligned8 = true;  // the pair kernel's 64-bit policy-side accesses
  for (int k = 0; k < n_problems; ++k) {
    const srl_loss_problem& q = problems[k];
    aligned8 = aligned8 && aligned(q.new_logp, 8) && aligned(q.v_pred, 8) && aligned(q.entropy, 8) && aligned(q.g_logp, 8) &&
               aligned(q.g_value, 8) && aligned(q.g_entropy, 8) && (q.lane_idx == nullptr || aligned(q.lane_idx, 8));
  }
  const bool pair = loss_pair_eligible(s, aligned8);
  SRL_REQUIRE(lane_aos == nullptr || pair, SRL_ERR_UNSUPPORTED,
              "srl_ppo_loss_fwd_bwd_batched: self-computed statistics need an even minibatch width <= 1024 lanes, even row "
              "strides and 8-byte aligned policy-side tensors (n = %d)", n);
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
  const XchgView* xv = xchg_view(xchg);
  SRL_REQUIRE(xchg == nullptr || (xv != nullptr && lane_aos != nullptr && xv->cap >= 3 * kMaxBatch), SRL_ERR_INVALID_ARG,
              "srl_ppo_loss_fwd_bwd_batched: the in-kernel exchange needs a connected srl_xchg of >= %d doubles and the "
              "self-computed statistics (lane_aos)", 3 * kMaxBatch);
  for (int k = 0; k < n_problems && xv != nullptr; ++k)
    SRL_REQUIRE(problems[k].out != nullptr, SRL_ERR_INVALID_ARG,
                "srl_ppo_loss_fwd_bwd_batched: the in-kernel exchange needs immediate finalisation (out != NULL): the last CTA "
                "of each problem ends the exchange round");
  for (int k0 = 0; k0 < n_problems; k0 += kMaxBatch) {
    const int nb = n_problems - k0 < kMaxBatch ? n_problems - k0 : kMaxBatch;
    LossBatch b;
    b.s = s;
    if (xv != nullptr && xv->world > 1) {
      b.xv = *xv;
    } else {
      memset(&b.xv, 0, sizeof(b.xv));
    }
    for (int k = 0; k < kMaxBatch; ++k) fill(b.prob[k], problems[k0 + (k < nb ? k : 0)]);
    if (mode == kPack && pair)
      rc = launch_loss_pair(b, nb, st);
    else if (mode == kPack)
      rc = launch_loss_pack(b, nb, dense_ok, st);
    else if (mode == kDense)
      rc = launch_loss_dense(b, nb, dense_ok, st);
    else
      rc = launch_loss_gather(b, nb, dense_ok, st);
    if (rc != SRL_OK) return rc;
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
                                      nullptr, 0, nullptr, T, n, popart_mean_std, hyper, workspace_bytes, nullptr, stream);
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
  using namespace srl::loss;
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
  const size_t row_bytes = static_cast<size_t>(256) * (sk + 1) * sizeof(float);
  SRL_REQUIRE(row_bytes <= 200 * 1024, SRL_ERR_UNSUPPORTED, "srl_ppo_loss_from_logits: sum K = %d too wide (max 199)", sk);
  const bool cache_e = 2 * row_bytes <= 200 * 1024;  // both shared-memory rows fit: one exponential per logit
  const size_t smem = cache_e ? 2 * row_bytes : row_bytes;
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
  p.lane_aos = nullptr;
  p.row_lo = 0;
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
    SRL_CUDA(cudaFuncSetAttribute(ppo_loss_logits_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    SRL_CUDA(cudaFuncSetAttribute(ppo_loss_logits_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    opted_in[dev] = true;
  }
  const long long tiles = (static_cast<long long>(T) * n + 255) / 256;
  const long long cap = static_cast<long long>(sm_count()) * 8 < kMaxGrid ? static_cast<long long>(sm_count()) * 8 : kMaxGrid;
  // every CTA the same number of tiles (2048 tiles on 1184 slots would leave most CTAs one tile and the rest two)
  const long long per_cta = (tiles + cap - 1) / cap;
  const int grid = static_cast<int>((tiles + per_cta - 1) / per_cta);
  if (cache_e)
    ppo_loss_logits_kernel<true><<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(q);
  else
    ppo_loss_logits_kernel<false><<<grid, 256, smem, static_cast<cudaStream_t>(stream)>>>(q);
  SRL_CUDA(cudaGetLastError());
  return SRL_OK;
}
