// K1 / K5b: batch assembly and minibatch gather over many leaves in one call (bit-exact byte copy).
// dst[t, j, :] = src(t, idx[j], :) with src either [L, slots, row] (a slot slab) or [slots, L, row] (per-sample
// staging: the np.stack(axis=1) transpose itself) -- replaces recursive_aggregate(np.stack(axis=1)) of
// PriorityQueueBuffer.put (base/buffer.py:118-126, base/namedarray.py:598-633) and
// SharedMemoryDock.get (base/shared_memory.py:85-99).
//
// Leaves differ by four orders of magnitude in row size (1 B flags .. 28 224 B Atari frames), so
// leaves are bucketed by row size and each bucket gets the thread-group that keeps its accesses
// coalesced and its index arithmetic off the critical path:
//   CTA  per (t, j) item for rows >= 4 KiB : 256 threads stream the row with 128-bit loads/stores
//   warp per item        for rows >= 64 B
//   thread per item      for short rows (scalars): coalesced writes, sector-granular reads
// The item -> (leaf, t, j) decode (two integer divisions) happens once per item, not per 16 bytes.
#include "common.cuh"

namespace srl {
namespace {

struct GatherLeaf {
  const unsigned char* src;
  unsigned char* dst;
  long long row_bytes;
  long long t_stride;     // bytes from (t, slot) to (t + 1, slot) in src
  long long slot_stride;  // bytes from (t, slot) to (t, slot + 1) in src
  long long item_begin;  // prefix sum of items (L * B per leaf) inside this bucket
  int unit;              // copy granule: 16, 4 or 1 bytes
};

struct GatherParams {
  GatherLeaf leaf[SRL_MAX_LEAVES];
  int n_leaves;
  int L, B;
  long long total_items;
  const int32_t* idx;
};

template <int UNIT>
struct Granule;
template <>
struct Granule<16> {
  using type = int4;
};
template <>
struct Granule<4> {
  using type = uint32_t;
};
template <>
struct Granule<1> {
  using type = uint8_t;
};

template <int UNIT, int GS>
__device__ __forceinline__ void copy_row(const unsigned char* __restrict__ s, unsigned char* __restrict__ d,
                                         long long row_bytes, int member) {
  using G = typename Granule<UNIT>::type;
  const G* sp = reinterpret_cast<const G*>(s);
  G* dp = reinterpret_cast<G*>(d);
  const long long n = row_bytes / UNIT;
  long long i = member;
  // 4 independent granules in flight per thread before the first store
  for (; i + 3 * GS < n; i += 4 * GS) {
    const G a = sp[i], b = sp[i + GS], c = sp[i + 2 * GS], e = sp[i + 3 * GS];
    dp[i] = a;
    dp[i + GS] = b;
    dp[i + 2 * GS] = c;
    dp[i + 3 * GS] = e;
  }
  for (; i < n; i += GS) dp[i] = sp[i];
}

// GS = threads cooperating on one item: 1, 32 or blockDim (256)
template <int GS>
__global__ void __launch_bounds__(256) gather_kernel(const __grid_constant__ GatherParams p) {
  const long long groups_per_block = 256 / GS;
  const long long group0 = blockIdx.x * groups_per_block + threadIdx.x / GS;
  const long long group_stride = static_cast<long long>(gridDim.x) * groups_per_block;
  const int member = threadIdx.x % GS;
  const long long per_leaf = static_cast<long long>(p.L) * p.B;
  for (long long item = group0; item < p.total_items; item += group_stride) {
    int li = 0;
    while (li + 1 < p.n_leaves && item >= p.leaf[li + 1].item_begin) ++li;
    const GatherLeaf& lf = p.leaf[li];
    const long long local = item - lf.item_begin;
    const long long t = local / p.B;
    const int j = static_cast<int>(local - t * p.B);
    const long long slot = p.idx ? p.idx[j] : j;
    const unsigned char* s = lf.src + t * lf.t_stride + slot * lf.slot_stride;
    unsigned char* d = lf.dst + local * lf.row_bytes;
    if (lf.unit == 16)
      copy_row<16, GS>(s, d, lf.row_bytes, member);
    else if (lf.unit == 4)
      copy_row<4, GS>(s, d, lf.row_bytes, member);
    else
      copy_row<1, GS>(s, d, lf.row_bytes, member);
    (void)per_leaf;
  }
}

template <int GS>
int launch_bucket(GatherParams& p, cudaStream_t st) {
  if (p.n_leaves == 0) return SRL_OK;
  const long long groups_per_block = 256 / GS;
  long long grid = (p.total_items + groups_per_block - 1) / groups_per_block;
  const long long cap = static_cast<long long>(sm_count()) * 16;
  if (grid > cap) grid = cap;
  gather_kernel<GS><<<static_cast<int>(grid), 256, 0, st>>>(p);
  SRL_CUDA(cudaGetLastError());
  return SRL_OK;
}

}  // namespace
}  // namespace srl

extern "C" int srl_batch_gather(const srl_leaf_desc* leaves, int n_leaves, const int32_t* idx, int L, int B,
                                srl_stream_t stream) {
  using namespace srl;
  SRL_REQUIRE(n_leaves >= 0 && n_leaves <= SRL_MAX_LEAVES, SRL_ERR_INVALID_ARG,
              "srl_batch_gather: n_leaves=%d outside [0, %d]", n_leaves, SRL_MAX_LEAVES);
  SRL_REQUIRE(L >= 0 && B >= 0, SRL_ERR_INVALID_ARG, "srl_batch_gather: negative L or B");
  if (n_leaves == 0 || L == 0 || B == 0) return SRL_OK;
  SRL_REQUIRE(leaves != nullptr, SRL_ERR_INVALID_ARG, "srl_batch_gather: null leaf table");
  GatherParams bucket[3];  // 0: thread per item, 1: warp per item, 2: CTA per item
  for (auto& b : bucket) {
    b.n_leaves = 0;
    b.L = L;
    b.B = B;
    b.total_items = 0;
    b.idx = idx;
  }
  for (int i = 0; i < n_leaves; ++i) {
    const srl_leaf_desc& d = leaves[i];
    SRL_REQUIRE(d.src && d.dst && d.row_bytes > 0 && d.src_slots > 0, SRL_ERR_INVALID_ARG,
                "srl_batch_gather: leaf %d has a null pointer or non-positive size", i);
    SRL_REQUIRE(idx != nullptr || d.src_slots >= B, SRL_ERR_INVALID_ARG,
                "srl_batch_gather: leaf %d has %lld slots < B=%d", i, static_cast<long long>(d.src_slots), B);
    const long long t_stride = d.src_t_stride ? d.src_t_stride : d.row_bytes * d.src_slots;
    const long long slot_stride = d.src_slot_stride ? d.src_slot_stride : d.row_bytes;
    SRL_REQUIRE(t_stride > 0 && slot_stride > 0, SRL_ERR_INVALID_ARG, "srl_batch_gather: leaf %d has a negative stride", i);
    int unit = 1;
    if (d.row_bytes % 16 == 0 && aligned(d.src, 16) && aligned(d.dst, 16) && t_stride % 16 == 0 && slot_stride % 16 == 0)
      unit = 16;
    else if (d.row_bytes % 4 == 0 && aligned(d.src, 4) && aligned(d.dst, 4) && t_stride % 4 == 0 && slot_stride % 4 == 0)
      unit = 4;
    const int which = d.row_bytes >= 4096 ? 2 : (d.row_bytes >= 64 ? 1 : 0);
    GatherParams& b = bucket[which];
    GatherLeaf& g = b.leaf[b.n_leaves++];
    g.src = static_cast<const unsigned char*>(d.src);
    g.dst = static_cast<unsigned char*>(d.dst);
    g.row_bytes = d.row_bytes;
    g.t_stride = t_stride;
    g.slot_stride = slot_stride;
    g.item_begin = b.total_items;
    g.unit = unit;
    b.total_items += static_cast<long long>(L) * B;
  }
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  int rc = launch_bucket<1>(bucket[0], st);
  if (rc != SRL_OK) return rc;
  rc = launch_bucket<32>(bucket[1], st);
  if (rc != SRL_OK) return rc;
  return launch_bucket<256>(bucket[2], st);
}
