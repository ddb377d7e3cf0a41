// K4, pack form: the kernel behind every permuted minibatch (4 epochs x 8 minibatches of cfg2 in one launch, or one
// minibatch per launch in the trainer's dependency order).  It replaced the row-tile kernel of round 1 for this form; what the
// profiles said along the way (profiles/r1d_ncu_cfg2.txt, profiles/r2_notes.md, the device timeline of
// profiles/microbench/timeline.py) and what the kernel does about it:
//
//  * round 1: every CTA of the one-wave grid in lockstep -- a burst of loads for one row, then ~500 instructions of
//    arithmetic per warp with nothing in flight, five times over (26.6 us for 86 MB).  Here a thread owns 2 adjacent lanes x
//    2 rows per step, the next step's loads are in registers before this step's arithmetic, and 24 threads of the CTA
//    request the policy-side lines two steps ahead from L2 (prefetch.global.L2).
//  * the pack was one 16-byte item per transition: a per-environment permutation fetched a 32-byte sector for 16 bytes.
//    K2 interleaves row pairs -- pack2[t / 2][lane] = 32 bytes = {row t, row t+1} of {old_logp, value, ret, mask ? adv :
//    NaN} -- and a thread fetches its lane's pair as one 256-bit load: every sector that moves is used whole.
//  * the minibatch statistics: where the scan computed the permutation itself it also left every scan CTA's share of every
//    minibatch's sums (minibatch_part): a loss CTA adds N / 32 contiguous items -- one round of loads.  Otherwise ONE CTA per
//    problem adds K2's per-lane sums (lane_aos, one 256-bit gather per lane behind the index load) and publishes them in
//    the workspace slot; the others wait.  With several ranks that CTA also exchanges the sums before it publishes them.
//  * the grid is (problems, slices): a problem's (column tile, row pair) units are numbered in one line and dealt to its
//    CTAs in equal runs (+-1 unit), for any shape; 64-thread CTAs keep the units fine; the whole grid is one wave.
//  * the tail: one block barrier, one partial row per CTA, no fence, no atomic -- published words validate themselves
//    (ppo_loss.cuh: box / unbox) and the problem's first CTA folds the rows by looking at them, in CTA order
//    (deterministic for a given launch shape).
#include <stdlib.h>

#include "ppo_loss.cuh"

namespace srl {
namespace loss {
namespace {

#ifndef SRL_PAIR_THREADS
#define SRL_PAIR_THREADS 64
#endif
#ifndef SRL_PAIR_MIN_BLOCKS
#define SRL_PAIR_MIN_BLOCKS 8
#endif
#ifndef SRL_PAIR_PF_AHEAD
#define SRL_PAIR_PF_AHEAD 2
#endif
constexpr int kPairThreads = SRL_PAIR_THREADS;
constexpr int kPairLanes = 2 * kPairThreads;  // lanes per column tile
constexpr int kStatMax = 1024 / kPairThreads; // lanes per thread in the statistics prologue (n <= 1024)

struct PairSched {
  int pair_lo;    // absolute index of the first row pair that holds a loss row: row_lo >> 1
  int pairs;      // row pairs per column
  int col_tiles;  // column tiles per problem
  int cpp;        // units per problem = col_tiles * pairs
  int zero;       // 0 at run time, unknown at compile time: see the scoreboard note in the kernel
};

// Both rows' policy outputs of this thread's two lanes, and its two lanes' row pairs of the pack.
struct Stage {
  float2 nl[2], vp[2], en[2];
  float pk[2][8];
};

__device__ __forceinline__ float2 ld_cs2(const float* p) {
  float2 v;
  asm volatile("ld.global.cs.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "l"(p));
  return v;
}
__device__ __forceinline__ void ld_nc256(const float4* p, float (&v)[8]) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.f32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=f"(v[0]), "=f"(v[1]), "=f"(v[2]), "=f"(v[3]), "=f"(v[4]), "=f"(v[5]), "=f"(v[6]), "=f"(v[7])
               : "l"(p));
}
__device__ __forceinline__ void ld_nc256(const double* p, double (&v)[4]) {
  asm volatile("ld.global.nc.v4.f64 {%0, %1, %2, %3}, [%4];" : "=d"(v[0]), "=d"(v[1]), "=d"(v[2]), "=d"(v[3]) : "l"(p));
}

// Minibatch statistics from K2's per-lane sums (lane_aos[lane] = {sum mask, sum adv*mask, sum (adv*mask)^2, 0}), added
// over ALL lanes of the problem: thread i adds lanes i, i + THREADS, ... in that order, then the warp tree, then the warps
// in order.  `share` (immediate mode: the problem has a ticket, so somebody can clear the flag afterwards): only the
// problem's first CTA (`sender`) does this -- and, with several ranks, the exchange -- and publishes the sums in the
// workspace slot; the other CTAs of the problem wait for them.  (Every CTA adding them for itself -- the first version --
// put 1184 x 512 gathers of the same 16 K lane items through L2 at the same moment: the device timeline showed 4.6 us between
// griddepcontrol.wait and the first row, profiles/r2_notes.md.)  Without `share` every CTA adds them itself, in the same
// order, so all CTAs of a problem normalise with bit-identical statistics either way.
//
// `part` (K2's per-CTA shares of THIS problem's sums, srl_gae_scan_perm; [part_ctas][4]): the scan computed the permutation,
// so it knew every lane's minibatch and left one {count, sum, sum of squares} item per scan CTA and minibatch.  The sums are
// then part_ctas contiguous items -- ONE round of coalesced loads (128 items at cfg2) instead of the index round and the
// dependent round of 512 scattered 32-byte gathers behind it.
__device__ __forceinline__ Uniforms self_uniforms_aos(const double* __restrict__ lane_aos, const int32_t* __restrict__ idx,
                                                      const double* __restrict__ part, int part_ctas, int n, double adv_eps, double& mask_sum, double (*s_part)[8],
                                                      const XchgView& xv, int slot, bool sender, SlotHeader* hdr, bool share,
                                                      unsigned int xchg_seq) {
  if (share && !sender) {
    if (threadIdx.x == 0) {
      // the sender is resident (dispatched first, launch_pair) and waits for nothing but -- with several ranks -- its peers,
      // for at most spin_limit ticks, after which it publishes NaN: twice that bound is never reached unless the sender died
      const long long limit = 2 * (xv.world > 1 ? xv.spin_limit : kDefaultSpinLimit);
      const long long t0 = clock64();
      bool ok = true;
      unsigned long long w[4];
      const volatile unsigned long long* src = hdr->bc;
      while (true) {
#pragma unroll
        for (int q = 0; q < 4; ++q) w[q] = src[q];
        if (w[0] != 0ull && w[1] != 0ull && w[2] != 0ull && w[3] != 0ull) break;
        if (clock64() - t0 > limit) {
          ok = false;
          break;
        }
        __nanosleep(32);  // (a first long sleep or slower polling changed nothing: 30.24 - 30.42 us per step)
      }
      const double nan = __longlong_as_double(0x7ff8000000000000ll);
#pragma unroll
      for (int q = 0; q < 4; ++q) s_part[q][0] = ok ? unbox(w[q]) : nan;  // no statistics: nothing but NaN leaves this launch
    }
    __syncthreads();
    mask_sum = s_part[3][0];
    return uniforms_from(s_part[0][0], s_part[1][0], s_part[2][0], mask_sum, nullptr, adv_eps);
  }
  double a0 = 0.0, a1 = 0.0, a2 = 0.0;
  if (part != nullptr) {
    constexpr int kPartBatch = 4;  // items in flight per thread
    for (int i0 = threadIdx.x; i0 < part_ctas; i0 += kPairThreads * kPartBatch) {
      double v[kPartBatch][4];
#pragma unroll
      for (int q = 0; q < kPartBatch; ++q) {
        const int i = i0 + q * kPairThreads;
        if (i < part_ctas) {
          ld_nc256(part + 4 * static_cast<size_t>(i), v[q]);
        } else {
          v[q][0] = v[q][1] = v[q][2] = v[q][3] = 0.0;
        }
      }
#pragma unroll
      for (int q = 0; q < kPartBatch; ++q) {
        a0 += v[q][0];
        a1 += v[q][1];
        a2 += v[q][2];
      }
    }
  } else {
    // kStatBatch gathers in flight per thread at a time (all index loads first): two dependent rounds for 1024 lanes on
    // 64 threads instead of sixteen, without holding 16 x 4 doubles in registers
    constexpr int kStatBatch = kStatMax < 8 ? kStatMax : 8;
    int ci[kStatMax];
#pragma unroll
    for (int q = 0; q < kStatMax; ++q) {
      const int i = threadIdx.x + q * kPairThreads;
      ci[q] = i < n ? __ldg(idx + i) : -1;
    }
#pragma unroll
    for (int q0 = 0; q0 < kStatMax; q0 += kStatBatch) {
      if (q0 * kPairThreads >= n) break;  // uniform: the minibatch ends before this batch
      double v[kStatBatch][4];
#pragma unroll
      for (int q = 0; q < kStatBatch; ++q) {
        if (ci[q0 + q] >= 0) {
          ld_nc256(lane_aos + 4 * static_cast<size_t>(ci[q0 + q]), v[q]);
        } else {
          v[q][0] = v[q][1] = v[q][2] = v[q][3] = 0.0;
        }
      }
#pragma unroll
      for (int q = 0; q < kStatBatch; ++q) {
        a0 += v[q][0];
        a1 += v[q][1];
        a2 += v[q][2];
      }
    }
  }
  a0 = warp_sum(a0);
  a1 = warp_sum(a1);
  a2 = warp_sum(a2);
  const int warp = threadIdx.x >> 5;
  __syncthreads();  // a previous problem's readers are done with s_part
  if ((threadIdx.x & 31) == 0) {
    s_part[0][warp] = a0;
    s_part[1][warp] = a1;
    s_part[2][warp] = a2;
  }
  __syncthreads();
  double cnt = 0.0, s1 = 0.0, s2 = 0.0;
#pragma unroll
  for (int w = 0; w < kPairThreads / 32; ++w) {
    cnt += s_part[0][w];
    s1 += s_part[1][w];
    s2 += s_part[2][w];
  }
  mask_sum = cnt;  // the masked means of the loss stay rank-local
  double x[3] = {cnt, s1, s2};
  if (threadIdx.x == 0) SRL_TL(2, blockIdx.y * gridDim.x + blockIdx.x, 6);
  if (xv.world > 1) {
    // several ranks: the normalisation statistics are the sums over ALL ranks' minibatches (utils.py:58-61).  The problem's
    // first CTA sends this rank's three sums to every rank's mailbox; it (share) or every CTA collects the ranks' sums from
    // its own rank's mailbox and adds them in rank order.
    __shared__ unsigned int s_words[6 * kMaxWorld];
    xchg_problem_sums(xv, slot, sender, x, s_words, xchg_seq);
  }
  if (threadIdx.x == 0) SRL_TL(2, blockIdx.y * gridDim.x + blockIdx.x, 7);
  if (share && threadIdx.x < 4) {  // sender: four self-validating words, no fence
    volatile unsigned long long* dst = hdr->bc;
    dst[threadIdx.x] = box(threadIdx.x < 3 ? x[threadIdx.x] : cnt);
  }
  return uniforms_from(x[0], x[1], x[2], cnt, nullptr, adv_eps);
}

// Block reduction of the eight masked sums for the pair kernel: every warp's tree, then warp 0 alone adds the warps, writes
// the CTA's partial row and takes the problem's ticket -- the other warps are done after ONE block barrier (the round-1
// epilogue held the whole CTA through three barriers, eight fences and the atomic's round trip).  The last CTA's warp 0
// folds the problem's rows in row order (deterministic for a given launch shape).
__device__ __forceinline__ void pair_reduce_and_finalize(const Problem& pr, const LossHyperDev& h, const Acc& acc,
                                                         double mask_sum, int row, int n_rows, double (*sred)[8],
                                                         const XchgView& xv, unsigned int n_problems) {
  double* partials = reinterpret_cast<double*>(reinterpret_cast<char*>(pr.slot) + kPartialsOffset);
  double v[kNumSums] = {acc.pl, acc.vl, acc.en, acc.adv, acc.ratio, static_cast<double>(acc.clip), acc.vt, acc.ret};
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarps = kPairThreads / 32;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    double w[kNumSums];
#pragma unroll
    for (int k = 0; k < kNumSums; ++k) w[k] = __shfl_xor_sync(0xffffffffu, v[k], o);
#pragma unroll
    for (int k = 0; k < kNumSums; ++k) v[k] += w[k];
  }
  if (kWarps > 1) {
    if (lane == 0) {
#pragma unroll
      for (int k = 0; k < kNumSums; ++k) sred[k][warp] = v[k];
    }
    __syncthreads();
    if (warp != 0) return;
  }
  if (lane < kNumSums) {
    double s = 0.0;
    if (kWarps > 1) {
#pragma unroll
      for (int w = 0; w < kWarps; ++w) s += sred[lane][w];
    } else {
      // lane k keeps sum k: all lanes hold all sums after the butterfly
#pragma unroll
      for (int k = 0; k < kNumSums; ++k) s = (lane == k) ? v[k] : s;
    }
    // deferred mode: plain float64 rows for srl_ppo_loss_finalize.  Immediate mode: boxed words (ppo_loss.cuh: 0 = not there
    // yet), so the row needs no fence and no ticket -- the problem's first CTA waits for every row by looking at it.
    if (pr.out == nullptr)
      partials[static_cast<size_t>(row) * kNumSums + lane] = s;
    else
      reinterpret_cast<volatile unsigned long long*>(partials)[static_cast<size_t>(row) * kNumSums + lane] = box(s);
  }
  if (pr.out == nullptr) {  // deferred: publish what the finaliser needs and leave
    if (row == 0 && lane == 0) {
      pr.slot->n_rows = n_rows;
      pr.slot->mask_sum = mask_sum;
      pr.slot->wv = static_cast<double>(h.wv);
      pr.slot->we = static_cast<double>(h.we);
    }
    return;
  }
  if (lane == 0) SRL_TL(2, row * gridDim.x + blockIdx.x, 4);
  if (row != 0) return;
  // The problem's first CTA folds the rows (the round-2 first version: every CTA fenced its row and took a ticket, the last
  // one folded -- 1.5 us of fence + atomic round trip per CTA behind its last stores, on the kernel's tail).  Lane k (and
  // k + 8, k + 16, k + 24) walks the rows of sum k in row order, waiting for each word to appear and clearing it behind
  // itself; then the four quarter sums are added -- the same order of additions as before, the same bits.  Every CTA of
  // the problem is resident (one wave, launch_pair), so the wait ends; its bound is the statistics wait's.
  const int k = lane & 7, part = lane >> 3;
  double s = 0.0;
  constexpr int kFoldBatch = 16;
  volatile unsigned long long* rows = reinterpret_cast<volatile unsigned long long*>(partials);
  const long long limit = 2 * (xv.world > 1 ? xv.spin_limit : kDefaultSpinLimit);
  const long long t0 = clock64();
  for (int r0 = part; r0 < n_rows; r0 += 4 * kFoldBatch) {
    unsigned long long w[kFoldBatch];
    bool all = false, ok = true;
    while (!all) {
      all = true;
#pragma unroll
      for (int q = 0; q < kFoldBatch; ++q) {
        const int r = r0 + 4 * q;
        w[q] = r < n_rows ? rows[static_cast<size_t>(r) * kNumSums + k] : 1ull;
        all = all && w[q] != 0ull;
      }
      if (!all && clock64() - t0 > limit) {
        ok = false;
        break;
      }
    }
#pragma unroll
    for (int q = 0; q < kFoldBatch; ++q) {
      const int r = r0 + 4 * q;
      if (r < n_rows) {
        s += ok ? unbox(w[q]) : __longlong_as_double(0x7ff8000000000000ll);  // a CTA never arrived: NaN, not a partial sum
        rows[static_cast<size_t>(r) * kNumSums + k] = 0ull;
      }
    }
  }
  s += __shfl_down_sync(0xffffffffu, s, 16);
  s += __shfl_down_sync(0xffffffffu, s, 8);
  // lanes 0..7 hold the eight sums: each divides its own, then lane 0 gathers the quotients
  const double M = mask_sum;
  const double quo = s / M;
  double t[kNumSums];
#pragma unroll
  for (int q = 0; q < kNumSums; ++q) t[q] = __shfl_sync(0xffffffffu, quo, q);
  if (lane == 0) {
    const double wv = static_cast<double>(h.wv), we = static_cast<double>(h.we);
    const double pl = t[0], vl = t[1], el = -t[2];
    const double loss = pl + wv * vl + we * el;
    double* o = pr.out;
    o[SRL_OUT_LOSS] = loss;
    o[SRL_OUT_POLICY_LOSS] = pl;
    o[SRL_OUT_VALUE_LOSS] = vl;
    o[SRL_OUT_ENTROPY_LOSS] = el;
    o[SRL_OUT_ADVANTAGE] = t[3];
    o[SRL_OUT_IMPORTANCE_WEIGHT] = t[4];
    o[SRL_OUT_CLIP_RATIO] = t[5];
    o[SRL_OUT_VALUE_TARGETS] = t[6];
    o[SRL_OUT_DENORM_VALUE] = t[7];
    o[SRL_OUT_MASK_SUM] = M;
    for (int q = SRL_OUT_MASK_SUM + 1; q < SRL_LOSS_OUT_LEN; ++q) o[q] = 0.0;
    if (pr.out_f32) {
      pr.out_f32[0] = static_cast<float>(loss);
      pr.out_f32[1] = static_cast<float>(pl);
      pr.out_f32[2] = static_cast<float>(vl);
      pr.out_f32[3] = static_cast<float>(el);
    }
    // every CTA of the problem has written its row, so it has read the published statistics: ready for the next launch
    pr.slot->bc[0] = pr.slot->bc[1] = pr.slot->bc[2] = pr.slot->bc[3] = 0ull;
    // ... and is past the exchange: the last PROBLEM to get here ends the launch's exchange round (n_problems atomics per
    // launch instead of one per CTA)
    if (xv.world > 1) xchg_launch_done(xv, n_problems);
    SRL_TL(2, row * gridDim.x + blockIdx.x, 5);
  }
}

// grid = (problems, slices): CTA (k, x) owns units [x * cpp / slices, (x + 1) * cpp / slices) of problem k, a unit being one
// row pair of one column tile, numbered column tile by column tile.  A CTA never leaves its problem: the statistics
// prologue and the reduction epilogue (a few microseconds of dependent latency each) are paid exactly once per CTA, and
// a problem's partial rows are its `slices` CTAs.
template <class CFG>
__global__ void __launch_bounds__(kPairThreads, SRL_PAIR_MIN_BLOCKS) ppo_loss_pair_kernel(const __grid_constant__ LossBatch b,
                                                                                          const PairSched sc) {
  __shared__ double s_part[8][8];
  const LossShared& s = b.s;
  const LossHyperDev& h = s.h;
  const Problem& pr = b.prob[blockIdx.x];
  const int n = s.n, T = s.T, row_lo = s.row_lo;
  const long long N2 = 2 * s.ld_smp;  // float4 items per row pair of the pack
  const int u_begin = static_cast<int>(static_cast<long long>(blockIdx.y) * sc.cpp / gridDim.y);
  const int u_end = static_cast<int>(static_cast<long long>(blockIdx.y + 1) * sc.cpp / gridDim.y);
  int ct = u_begin / sc.pairs;
  int pi = u_begin - ct * sc.pairs;

  Stage sA = {}, sB = {};
  const int tl_cta = blockIdx.y * gridDim.x + blockIdx.x;
  if (threadIdx.x == 0) SRL_TL(2, tl_cta, 0);
  // Scoreboards.  ptxas put every load of this loop on ONE hardware scoreboard (SB5; decoded from the SASS control words),
  // so the first use of step i's data ALSO waited for the loads of step i+1 issued just before it -- 20 % of all warp
  // samples sat on that one instruction (ncu, warm L2) and the software pipeline hid nothing.  The fix is an ordering the
  // scheduler cannot undo: the addresses of step i+1's loads depend on a value derived from step i's data (`& sc.zero`, 0
  // at run time), so the wait for step i happens BEFORE step i+1 is issued, when nothing younger is outstanding.
  auto landed = [&](const Stage& st) -> int {
    return (__float_as_int(st.nl[0].x) | __float_as_int(st.vp[0].x) | __float_as_int(st.en[0].x) | __float_as_int(st.nl[1].x) |
            __float_as_int(st.vp[1].x) | __float_as_int(st.en[1].x) | __float_as_int(st.pk[0][0]) |
            __float_as_int(st.pk[1][0])) & sc.zero;
  };
  // Addresses.  The first version recomputed every address of a step from (row, lane) -- 64-bit multiplies and the
  // problem's pointers re-read from the parameter bank through blockIdx.x: ~125 of the 451 instructions of a step (ncu
  // --page source).  Now: one running byte pointer per side (policy outputs, gradients) and one per pack lane, advanced by a
  // constant per step; the other arrays and the pair's second row are that pointer plus a UNIFORM byte offset (the distance
  // between the problem's arrays), so a load or store costs one 64-bit add.
  typedef const char* cptr;
  const long long pol_row = s.ld_pol * 4, grad_row = s.ld_grad * 4;  // bytes per row
  const long long d_vp = reinterpret_cast<cptr>(pr.v_pred) - reinterpret_cast<cptr>(pr.new_logp);
  const long long d_en = reinterpret_cast<cptr>(pr.entropy) - reinterpret_cast<cptr>(pr.new_logp);
  const long long d_gv = reinterpret_cast<cptr>(pr.g_value) - reinterpret_cast<cptr>(pr.g_logp);
  const long long d_ge = reinterpret_cast<cptr>(pr.g_entropy) - reinterpret_cast<cptr>(pr.g_logp);
  const long long pack_step = N2 * static_cast<long long>(sizeof(float4));
  // policy-side loads of the row pair whose first loss row is r0 (-1: before the first loss row); q = &new_logp[r0][j]
  auto issue_policy = [&](Stage& st, int r0, cptr q) {
    if (r0 >= 0) {
      st.nl[0] = ld_cs2(reinterpret_cast<const float*>(q));
      st.vp[0] = ld_cs2(reinterpret_cast<const float*>(q + d_vp));
      st.en[0] = ld_cs2(reinterpret_cast<const float*>(q + d_en));
    }
    if (r0 + 1 < T) {
      st.nl[1] = ld_cs2(reinterpret_cast<const float*>(q + pol_row));
      st.vp[1] = ld_cs2(reinterpret_cast<const float*>(q + (pol_row + d_vp)));
      st.en[1] = ld_cs2(reinterpret_cast<const float*>(q + (pol_row + d_en)));
    }
  };
  auto issue_pack = [&](Stage& st, cptr k0, cptr k1) {
    ld_nc256(reinterpret_cast<const float4*>(k0), st.pk[0]);
    ld_nc256(reinterpret_cast<const float4*>(k1), st.pk[1]);
  };
  auto policy_ptr = [&](int r0, int j) { return reinterpret_cast<cptr>(pr.new_logp + static_cast<long long>(r0) * s.ld_pol + j); };

  // Launched programmatically behind the scan, this CTA may be resident while the scan still runs: the policy outputs do
  // not depend on it, so the first step's policy-side loads are in flight before the wait.
  {
    const int j = (ct * kPairThreads + threadIdx.x) * 2;
    const int r0 = 2 * (sc.pair_lo + pi) - row_lo;
    if (j < n) issue_policy(sA, r0, policy_ptr(r0, j));
  }
  pdl_wait();
  if (threadIdx.x == 0) SRL_TL(2, tl_cta, 1);
  // the exchange's sequence number (it moved when the previous exchanging launch ended): loaded here, used after the local
  // sums are added -- its latency hides under theirs
  const unsigned int xchg_seq = b.xv.world > 1 ? *reinterpret_cast<volatile unsigned int*>(b.xv.seq) : 0u;

  // L2 requests for the policy-side rows a few steps ahead (prefetch.global.L2: no register, no scoreboard).  The loop keeps
  // ONE step of loads in flight, which hides an L2 hit but not a DRAM round trip under load, and the policy outputs are the
  // only DRAM-resident operand of the loop (the pack was just written by the scan).  A row pair of a column tile is 24
  // lines (2 rows x 3 arrays x 512 bytes): thread t < 24 requests line t of the pair kPfAhead steps ahead, every step.
  // (All of a CTA's rows requested at once -- before the wait, or right behind it -- was measured first: the burst of
  // 25 MB delayed the scan's own traffic or the statistics prologue's dependent loads by as much as the loop gained.)
  constexpr int kPfAhead = SRL_PAIR_PF_AHEAD;
  const bool pf_thread = kPfAhead > 0 && threadIdx.x < 24;
  const int pf_rr = (threadIdx.x % 24) / 12, pf_line = threadIdx.x & 3;
  const long long pf_arr = ((threadIdx.x % 12) >> 2) == 0 ? 0 : (((threadIdx.x % 12) >> 2) == 1 ? d_vp : d_en);
  auto request_pair = [&](int p, int ct_) {  // pair index within the column, column tile
    const int r = 2 * (sc.pair_lo + p) - row_lo + pf_rr, lane0 = ct_ * kPairLanes + pf_line * 32;
    if (r >= 0 && r < T && lane0 < n) prefetch_l2(policy_ptr(r, lane0) + pf_arr);
  };

  Acc acc;
  RowSums rs;
  int pending = 0;
  bool first = true;
  double mask_sum = 0.0;
  Uniforms uf;
  bool have_u = false;
  for (int u = u_begin; u < u_end;) {  // the column segments of this CTA's range
    const int j = (ct * kPairThreads + threadIdx.x) * 2;
    const bool active = j < n;
    const int p_end = min(sc.pairs, pi + (u_end - u));
    int c[2] = {j, j + 1};
    if (pr.lane_idx && active) {
      const int2 q = __ldg(reinterpret_cast<const int2*>(pr.lane_idx + j));
      c[0] = q.x, c[1] = q.y;
    }
    // the next pair to load / to store: loss row of its first row and the running pointers
    int r_ld = 2 * (sc.pair_lo + pi) - row_lo;
    cptr q_ld = policy_ptr(r_ld, j);
    cptr k0 = reinterpret_cast<cptr>(s.pack + static_cast<long long>(sc.pair_lo + pi) * N2 + 2 * c[0]);
    cptr k1 = reinterpret_cast<cptr>(s.pack + static_cast<long long>(sc.pair_lo + pi) * N2 + 2 * c[1]);
    int r_st = r_ld;
    char* g_st = reinterpret_cast<char*>(pr.g_logp + static_cast<long long>(r_st) * s.ld_grad + j);
    if (active) {
      if (!first) issue_policy(sA, r_ld, q_ld);
      issue_pack(sA, k0, k1);
    }
    if (pf_thread) {
#pragma unroll
      for (int a = 2; a < kPfAhead; ++a)
        if (pi + a < p_end) request_pair(pi + a, ct);
    }
    r_ld += 2, q_ld += 2 * pol_row, k0 += pack_step, k1 += pack_step;
    first = false;
    if (!have_u) {  // once per CTA; the first step's loads are in flight underneath
      if (s.lane_aos != nullptr) {
        const double* part =
            s.part != nullptr ? s.part + static_cast<size_t>(s.part_first + static_cast<int>(blockIdx.x)) * s.part_ctas * 4 : nullptr;
        // with the scan's shares and one rank, every CTA adds them itself: 4 KB of contiguous L2 hits per CTA, no
        // publication by the problem's first CTA and no polling by the others (measured equal to publishing them, 30.2 us
        // either way, profiles/r2_notes.md r2s; with several ranks the first CTA exchanges and publishes as before)
        const bool each = part != nullptr && b.xv.world <= 1;
        uf = self_uniforms_aos(s.lane_aos, pr.lane_idx, part, s.part_ctas, n, h.adv_eps, mask_sum, s_part, b.xv,
                               static_cast<int>(blockIdx.x), blockIdx.y == 0, pr.slot,
                               pr.out != nullptr && gridDim.y > 1 && !each, xchg_seq);
      }
      else
        uf = load_uniforms(pr.norm_stats, pr.local_stats, s.popart, h.adv_eps, mask_sum);
      have_u = true;
      if (threadIdx.x == 0) SRL_TL(2, tl_cta, 2);
    }
    if (active) {
      // one step: both rows of the pair, both lanes; masked or out-of-range rows produce nothing
      auto compute = [&](const Stage& st) {
#pragma unroll
        for (int rr = 0; rr < 2; ++rr) {
          const int r = r_st + rr;
          if (r < 0 || r >= T) continue;
          float glp[2], gv[2], ge[2];
          const float nl[2] = {st.nl[rr].x, st.nl[rr].y}, vp[2] = {st.vp[rr].x, st.vp[rr].y},
                      en[2] = {st.en[rr].x, st.en[rr].y};
#pragma unroll
          for (int q = 0; q < 2; ++q) {
            const float ad = st.pk[q][4 * rr + 3];
            element<CFG>(h, uf, nl[q], vp[q], en[q], st.pk[q][4 * rr + 0], st.pk[q][4 * rr + 1], st.pk[q][4 * rr + 2], ad,
                         ad == ad, glp[q], gv[q], ge[q], rs);
          }
          char* g = g_st + (rr ? grad_row : 0);
          stg_stream(reinterpret_cast<float2*>(g), make_float2(glp[0], glp[1]));
          stg_stream(reinterpret_cast<float2*>(g + d_gv), make_float2(gv[0], gv[1]));
          stg_stream(reinterpret_cast<float2*>(g + d_ge), make_float2(ge[0], ge[1]));
        }
        r_st += 2, g_st += 2 * grad_row;
        if (++pending == kFlushRows) {  // fp32 partial sums of at most kFlushRows * 4 terms, then float64
          acc.add(rs);
          rs = RowSums();
          pending = 0;
        }
      };
      int p = pi;
      while (true) {
        if (pf_thread && p + kPfAhead < p_end) request_pair(p + kPfAhead, ct);
        if (p + 1 < p_end) {
          const long long dep = landed(sA);
          issue_policy(sB, r_ld, q_ld + dep);
          issue_pack(sB, k0 + dep, k1 + dep);
          r_ld += 2, q_ld += 2 * pol_row, k0 += pack_step, k1 += pack_step;
        }
        compute(sA);
        if (++p >= p_end) break;
        if (pf_thread && p + kPfAhead < p_end) request_pair(p + kPfAhead, ct);
        if (p + 1 < p_end) {
          const long long dep = landed(sB);
          issue_policy(sA, r_ld, q_ld + dep);
          issue_pack(sA, k0 + dep, k1 + dep);
          r_ld += 2, q_ld += 2 * pol_row, k0 += pack_step, k1 += pack_step;
        }
        compute(sB);
        if (++p >= p_end) break;
      }
    }
    u += p_end - pi;
    pi = 0;
    ++ct;
  }
  acc.add(rs);
  if (threadIdx.x == 0) SRL_TL(2, tl_cta, 3);
  pair_reduce_and_finalize(pr, h, acc, mask_sum, static_cast<int>(blockIdx.y), static_cast<int>(gridDim.y), s_part, b.xv,
                           gridDim.x);
}

template <class CFG>
int launch_pair(LossBatch& b, int n_problems, cudaStream_t st) {
  const LossShared& s = b.s;
  auto kern = ppo_loss_pair_kernel<CFG>;
  static int resident[64] = {};
  int dev = 0;
  SRL_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64) dev = 0;
  if (resident[dev] == 0) {
    int nb = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kern, kPairThreads, 0) != cudaSuccess || nb < 1) nb = 1;
    resident[dev] = nb;
  }
  PairSched sc;
  sc.pair_lo = s.row_lo >> 1;
  sc.pairs = ((s.row_lo + s.T - 1) >> 1) - sc.pair_lo + 1;
  sc.col_tiles = (s.n + kPairLanes - 1) / kPairLanes;
  sc.cpp = sc.col_tiles * sc.pairs;
  sc.zero = 0;
  // slices per problem: the whole grid resident at once (one wave), every CTA inside ONE problem
  long long slices = static_cast<long long>(sm_count()) * resident[dev] / n_problems;
  if (slices > sc.cpp) slices = sc.cpp;
  if (slices > kMaxGrid) slices = kMaxGrid;  // one partial row per CTA in the workspace slot
  if (slices < 1) slices = 1;
  // problems along x: the first n_problems CTAs to be dispatched are every problem's slice 0 -- the senders of the in-kernel
  // exchange -- so they are resident (and sending) whatever else shares the machine; no CTA ever waits for a CTA that is
  // not resident yet
  SRL_CUDA(launch_pdl(kern, dim3(static_cast<unsigned>(n_problems), static_cast<unsigned>(slices)), dim3(kPairThreads), 0, st,
                      b, sc));
  return SRL_OK;
}

}  // namespace
}  // namespace loss
}  // namespace srl
SRL_TL_SETTER(srl_tl_set_loss)
namespace srl {
namespace loss {

bool loss_pair_eligible(const LossShared& s, bool aligned8) {
  return s.pack != nullptr && aligned8 && (s.n % 2 == 0) && (s.ld_pol % 2 == 0) && (s.ld_grad % 2 == 0) &&
         (s.lane_aos == nullptr || s.n <= 1024);
}

int launch_loss_pair(LossBatch& b, int n_problems, cudaStream_t st) {
  const LossHyperDev& h = b.s.h;
  const bool popart = b.s.popart != nullptr;
  const int key = (h.value_loss == SRL_VL_MSE ? 0 : h.value_loss == SRL_VL_HUBER ? 1 : 2) * 8 + (h.clip_value ? 4 : 0) +
                  (h.dual_clip ? 2 : 0) + (popart ? 1 : 0);
#define SRL_PAIR_CASE(k, vl, clip, dual, pa) \
  case k:                                    \
    return launch_pair<StaticCfg<vl, clip, dual, pa>>(b, n_problems, st)
  switch (key) {
    SRL_PAIR_CASE(0, SRL_VL_MSE, false, false, false);
    SRL_PAIR_CASE(1, SRL_VL_MSE, false, false, true);
    SRL_PAIR_CASE(2, SRL_VL_MSE, false, true, false);
    SRL_PAIR_CASE(3, SRL_VL_MSE, false, true, true);
    SRL_PAIR_CASE(4, SRL_VL_MSE, true, false, false);
    SRL_PAIR_CASE(5, SRL_VL_MSE, true, false, true);
    SRL_PAIR_CASE(6, SRL_VL_MSE, true, true, false);
    SRL_PAIR_CASE(7, SRL_VL_MSE, true, true, true);
    SRL_PAIR_CASE(8, SRL_VL_HUBER, false, false, false);
    SRL_PAIR_CASE(9, SRL_VL_HUBER, false, false, true);
    SRL_PAIR_CASE(10, SRL_VL_HUBER, false, true, false);
    SRL_PAIR_CASE(11, SRL_VL_HUBER, false, true, true);
    SRL_PAIR_CASE(12, SRL_VL_HUBER, true, false, false);
    SRL_PAIR_CASE(13, SRL_VL_HUBER, true, false, true);
    SRL_PAIR_CASE(14, SRL_VL_HUBER, true, true, false);
    SRL_PAIR_CASE(15, SRL_VL_HUBER, true, true, true);
    default:
      return launch_pair<RuntimeCfg>(b, n_problems, st);
  }
#undef SRL_PAIR_CASE
}

}  // namespace loss
}  // namespace srl
