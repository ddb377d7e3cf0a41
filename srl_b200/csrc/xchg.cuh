// Device side of the peer-memory statistics exchange (see xchg.cu for the protocol), shared by the stand-alone kernel
// (xchg.cu) and the fused group-statistics + exchange kernel (stats.cu).
#pragma once
#include "common.cuh"

namespace srl {

constexpr int kMaxWorld = 16;
// A rank waits for its peers like a collective does: minutes, not seconds -- a peer that autotunes its first step, saves a
// checkpoint or waits for samples is late, not gone.  Only after `spin_limit` clock64 ticks (default ~10 min, the order of
// NCCL's watchdog; srl_xchg_set_timeout) does it give up, and then it produces NO sums: the whole output table is NaN and
// the sticky status says why, so a caller that does not look at the status still cannot train on stale numbers.
constexpr long long kDefaultSpinLimit = 1200000000000ll;

struct XchgView {
  unsigned long long* mailbox[kMaxWorld];  // peer p's mailbox base (mailbox[rank] = local); words = seq << 32 | payload
  unsigned int* seq;          // local: exchanges so far
  int* status;                // local: 0 ok, 1 timed out
  unsigned int* done;         // local: CTAs of the current launch that have finished (in-kernel exchange of the loss kernel)
  long long spin_limit;       // clock64 ticks a rank waits for a peer's words
  int world, rank, cap;
};

// The exchange itself, executed by ONE whole CTA (any block size).  `local` may have been written by other CTAs of the
// same grid (srl_group_stats_xchg): it is read through L2 (__ldcg) after the caller's fence + ticket.
__device__ __forceinline__ void xchg_exchange(const XchgView& v, const double* __restrict__ local,
                                              double* __restrict__ global, int n) {
  __shared__ int s_bad;
  if (threadIdx.x == 0) s_bad = 0;
  unsigned int seq = *v.seq + 1u;
  if (seq == 0u) seq = 1u;  // 0 is the mailbox's initial state
  const int par = static_cast<int>(seq & 1u);
  __syncthreads();  // everybody has read the old sequence number before thread 0 updates it
  // 1. my table, two {payload, seq} words per float64, into every mailbox (mine included), slot [par][rank]
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    const double x = __ldcg(local + i);
    const unsigned long long tag = static_cast<unsigned long long>(seq) << 32;
    const unsigned long long lo = tag | static_cast<unsigned int>(__double2loint(x));
    const unsigned long long hi = tag | static_cast<unsigned int>(__double2hiint(x));
    for (int p = 0; p < v.world; ++p) {
      volatile unsigned long long* dst = v.mailbox[p] + ((static_cast<size_t>(par) * v.world + v.rank) * v.cap + i) * 2;
      dst[0] = lo;
      dst[1] = hi;
    }
  }
  // 2. collect: spin on each word of each rank's slot in my own mailbox, add in rank order (the same order on every
  //    rank -> bit-identical tables)
  const unsigned long long* mine = v.mailbox[v.rank] + static_cast<size_t>(par) * v.world * v.cap * 2;
  const long long t0 = clock64();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    double s = 0.0;
    for (int q = 0; q < v.world; ++q) {
      const volatile unsigned long long* src = mine + (static_cast<size_t>(q) * v.cap + i) * 2;
      unsigned long long lo = src[0], hi = src[1];
      while (static_cast<unsigned int>(lo >> 32) != seq || static_cast<unsigned int>(hi >> 32) != seq) {
        if (clock64() - t0 > v.spin_limit) {
          s_bad = 1;
          break;
        }
        lo = src[0];
        hi = src[1];
      }
      s += __hiloint2double(static_cast<int>(static_cast<unsigned int>(hi)), static_cast<int>(static_cast<unsigned int>(lo)));
    }
    global[i] = s;
  }
  __syncthreads();
  if (s_bad) {  // a peer never arrived: no partial sums leave this kernel
    const double nan = __longlong_as_double(0x7ff8000000000000ll);
    for (int i = threadIdx.x; i < n; i += blockDim.x) global[i] = nan;
  }
  if (threadIdx.x == 0) {
    *v.seq = seq;
    if (s_bad) *v.status = 1;
  }
}


// The loss kernel's own exchange (ppo_loss_pair.cu): every CTA of problem `slot` holds the problem's LOCAL sums x[0..2];
// the problem's first CTA (`sender`) stores them into every rank's mailbox, every CTA collects the ranks' sums from its own
// rank's mailbox and adds them in rank order (bit-identical on every rank and in every CTA).  Same word format and double
// buffering as xchg_exchange; the sequence number advances once per LAUNCH (xchg_launch_done, by the last CTA to finish).
// Called by all threads of the CTA after griddepcontrol.wait (the previous launch has advanced the sequence number by then).
// `seq_now`: the mailbox's sequence number as read by the caller AFTER its griddepcontrol.wait (the number moves at the
// end of a launch that exchanged), sparing this call a dependent load.
__device__ __forceinline__ bool xchg_problem_sums(const XchgView& v, int slot, bool sender, double (&x)[3], unsigned int* sh_words,
                                                  unsigned int seq_now) {
  unsigned int seq = seq_now + 1u;
  if (seq == 0u) seq = 1u;
  const int par = static_cast<int>(seq & 1u);
  const unsigned long long tag = static_cast<unsigned long long>(seq) << 32;
  const int words = 6 * v.world;
  if (sender) {
    for (int i = threadIdx.x; i < words; i += blockDim.x) {
      const int p = i / 6, w = i - 6 * p;
      const double val = x[w >> 1];
      const unsigned int half = (w & 1) ? static_cast<unsigned int>(__double2hiint(val)) : static_cast<unsigned int>(__double2loint(val));
      volatile unsigned long long* dst =
          v.mailbox[p] + ((static_cast<size_t>(par) * v.world + v.rank) * v.cap + slot * 3 + (w >> 1)) * 2 + (w & 1);
      *dst = tag | half;
    }
  }
  const unsigned long long* mine = v.mailbox[v.rank] + static_cast<size_t>(par) * v.world * v.cap * 2;
  bool bad = false;
  const long long t0 = clock64();
  for (int i = threadIdx.x; i < words; i += blockDim.x) {
    const int q = i / 6, w = i - 6 * q;
    const volatile unsigned long long* src = mine + (static_cast<size_t>(q) * v.cap + slot * 3 + (w >> 1)) * 2 + (w & 1);
    unsigned long long word = *src;
    while (static_cast<unsigned int>(word >> 32) != seq) {
      if (clock64() - t0 > v.spin_limit) {
        bad = true;
        break;
      }
      word = *src;  // (one CTA per problem polls: no back-off needed)
    }
    sh_words[i] = static_cast<unsigned int>(word);
  }
  bad = __syncthreads_or(bad);
  const double nan = __longlong_as_double(0x7ff8000000000000ll);
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    double s = 0.0;
    for (int q = 0; q < v.world; ++q)
      s += __hiloint2double(static_cast<int>(sh_words[q * 6 + 2 * k + 1]), static_cast<int>(sh_words[q * 6 + 2 * k]));
    x[k] = bad ? nan : s;  // a peer never arrived: no partial sums leave this function
  }
  if (bad && threadIdx.x == 0) *v.status = 1;
  return !bad;
}

// End of a launch that used xchg_problem_sums: the last CTA to get here advances the sequence number (one thread per CTA).
__device__ __forceinline__ void xchg_launch_done(const XchgView& v, unsigned int total_ctas) {
  const unsigned int d = atomicAdd(v.done, 1u);
  if (d == total_ctas - 1u) {
    unsigned int seq = *reinterpret_cast<volatile unsigned int*>(v.seq) + 1u;
    if (seq == 0u) seq = 1u;
    *v.done = 0u;
    __threadfence();
    *reinterpret_cast<volatile unsigned int*>(v.seq) = seq;
  }
}

struct Xchg;                           // host handle (xchg.cu)
const XchgView* xchg_view(const srl_xchg* h);  // nullptr if not connected

}  // namespace srl
