// One-shot SUM all-reduce of a small float64 table over NVLink peer memory (one process per GPU, one node).
//
// The only data-path collective of the step is the statistics table ([1 + E*M, 8] float64, <= a few KB): the reference
// issues 3 + 3 one-element NCCL all-reduces per loss call for it (legacy/algorithm/modules/utils.py:58-61,121-124).
// A tree/ring collective is pure latency at this size (NCCL: ~20 us between two graph launches at 2 GPUs, measured),
// so every rank instead STORES its table straight into every peer's mailbox (P2P stores through NVSwitch), publishes
// a sequence number, waits for the peers' numbers in its OWN memory, and adds the world's tables in rank order --
// bit-identical on every rank, one kernel, capturable in the step's CUDA graph.
//
// Protocol ("LL", as in low-latency collectives): every 8-byte store carries 4 bytes of payload and the 4-byte
// sequence number of the exchange, and an aligned 8-byte store is delivered atomically -- so there is no separate flag,
// no system-wide fence and no second round: the receiver spins on each word until its sequence number matches.  A
// float64 travels as two such words.
//
// Memory: each rank cudaMalloc's  mailbox[2][world][cap][2] x uint64  +  {seq, status}; the allocation is exported with
// cudaIpcGetMemHandle and opened by the peers (host side: srl_b200/xchg.py exchanges the 64-byte handles with
// torch.distributed.all_gather_object).  Double buffering by sequence parity makes the mailbox of exchange s+1
// independent of stragglers still reading exchange s (a rank can only reach s+2 after everybody published s+1, i.e.
// finished reading s).
#include <string.h>

#include <new>

#include "xchg.cuh"

namespace srl {
namespace {

__global__ void __launch_bounds__(256) xchg_allreduce_kernel(const XchgView v, const double* __restrict__ local,
                                                             double* __restrict__ global, int n) {
  xchg_exchange(v, local, global, n);
}

}  // namespace

struct Xchg {
  XchgView view;
  void* local_base = nullptr;             // my allocation
  void* peer_base[kMaxWorld] = {};        // opened peer allocations (null for myself)
  size_t bytes = 0;
  bool connected = false;
};

const XchgView* xchg_view(const srl_xchg* h) {
  const Xchg* x = reinterpret_cast<const Xchg*>(h);
  return (x != nullptr && x->connected) ? &x->view : nullptr;
}

namespace {
size_t mailbox_bytes(int world, int cap) {
  return (static_cast<size_t>(2) * world * cap * 2 * sizeof(unsigned long long) + 255) / 256 * 256;
}

void fill_view_for(XchgView& v, int p, void* base, int, int) { v.mailbox[p] = static_cast<unsigned long long*>(base); }
}  // namespace
}  // namespace srl

extern "C" int srl_xchg_create(int world, int rank, int capacity_doubles, srl_xchg** out) {
  using namespace srl;
  SRL_REQUIRE(out != nullptr, SRL_ERR_INVALID_ARG, "srl_xchg_create: null output");
  SRL_REQUIRE(world >= 1 && world <= kMaxWorld && rank >= 0 && rank < world && capacity_doubles >= 1,
              SRL_ERR_INVALID_ARG, "srl_xchg_create: need 1 <= world <= %d, 0 <= rank < world, capacity >= 1", kMaxWorld);
  Xchg* x = new (std::nothrow) Xchg();
  SRL_REQUIRE(x != nullptr, SRL_ERR_CUDA, "srl_xchg_create: out of host memory");
  x->bytes = mailbox_bytes(world, capacity_doubles) + 256;
  cudaError_t e = cudaMalloc(&x->local_base, x->bytes);
  if (e == cudaSuccess) e = cudaMemset(x->local_base, 0, x->bytes);
  if (e != cudaSuccess) {
    set_error("srl_xchg_create: cudaMalloc/cudaMemset of %zu bytes failed: %s", x->bytes, cudaGetErrorString(e));
    if (x->local_base) cudaFree(x->local_base);
    delete x;
    return SRL_ERR_CUDA;
  }
  x->view.world = world;
  x->view.rank = rank;
  x->view.cap = capacity_doubles;
  x->view.spin_limit = kDefaultSpinLimit;
  for (int p = 0; p < kMaxWorld; ++p) x->view.mailbox[p] = nullptr;
  fill_view_for(x->view, rank, x->local_base, world, capacity_doubles);
  char* tail = static_cast<char*>(x->local_base) + mailbox_bytes(world, capacity_doubles);
  x->view.seq = reinterpret_cast<unsigned int*>(tail);
  x->view.status = reinterpret_cast<int*>(tail + 64);
  x->view.done = reinterpret_cast<unsigned int*>(tail + 128);
  x->connected = (world == 1);
  *out = reinterpret_cast<srl_xchg*>(x);
  return SRL_OK;
}

extern "C" int srl_xchg_local_handle(srl_xchg* h, void* handle_out) {
  using namespace srl;
  SRL_REQUIRE(h && handle_out, SRL_ERR_INVALID_ARG, "srl_xchg_local_handle: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == SRL_XCHG_HANDLE_BYTES, "IPC handle size");
  Xchg* x = reinterpret_cast<Xchg*>(h);
  cudaIpcMemHandle_t mh;
  SRL_CUDA(cudaIpcGetMemHandle(&mh, x->local_base));
  memcpy(handle_out, &mh, sizeof(mh));
  return SRL_OK;
}

extern "C" int srl_xchg_connect(srl_xchg* h, const void* handles) {
  using namespace srl;
  SRL_REQUIRE(h && handles, SRL_ERR_INVALID_ARG, "srl_xchg_connect: null pointer");
  Xchg* x = reinterpret_cast<Xchg*>(h);
  const char* hb = static_cast<const char*>(handles);
  for (int p = 0; p < x->view.world; ++p) {
    if (p == x->view.rank || x->peer_base[p] != nullptr) continue;
    cudaIpcMemHandle_t mh;
    memcpy(&mh, hb + static_cast<size_t>(p) * SRL_XCHG_HANDLE_BYTES, sizeof(mh));
    void* base = nullptr;
    SRL_CUDA(cudaIpcOpenMemHandle(&base, mh, cudaIpcMemLazyEnablePeerAccess));
    x->peer_base[p] = base;
    fill_view_for(x->view, p, base, x->view.world, x->view.cap);
  }
  x->connected = true;
  return SRL_OK;
}

extern "C" int srl_xchg_allreduce_sum(srl_xchg* h, const double* local, double* global, int n, srl_stream_t stream) {
  using namespace srl;
  SRL_REQUIRE(h && local && global, SRL_ERR_INVALID_ARG, "srl_xchg_allreduce_sum: null pointer");
  Xchg* x = reinterpret_cast<Xchg*>(h);
  SRL_REQUIRE(x->connected, SRL_ERR_INVALID_ARG, "srl_xchg_allreduce_sum: peers not connected (srl_xchg_connect)");
  SRL_REQUIRE(n >= 0 && n <= x->view.cap, SRL_ERR_INVALID_ARG, "srl_xchg_allreduce_sum: n=%d exceeds the capacity %d", n,
              x->view.cap);
  if (n == 0) return SRL_OK;
  xchg_allreduce_kernel<<<1, 256, 0, static_cast<cudaStream_t>(stream)>>>(x->view, local, global, n);
  SRL_CUDA(cudaGetLastError());
  return SRL_OK;
}

extern "C" int srl_xchg_status(srl_xchg* h, int* status_out) {
  using namespace srl;
  SRL_REQUIRE(h && status_out, SRL_ERR_INVALID_ARG, "srl_xchg_status: null pointer");
  Xchg* x = reinterpret_cast<Xchg*>(h);
  SRL_CUDA(cudaMemcpy(status_out, x->view.status, sizeof(int), cudaMemcpyDeviceToHost));
  return SRL_OK;
}

extern "C" int srl_xchg_status_async(srl_xchg* h, int* pinned_status_out, srl_stream_t stream) {
  using namespace srl;
  SRL_REQUIRE(h && pinned_status_out, SRL_ERR_INVALID_ARG, "srl_xchg_status_async: null pointer");
  Xchg* x = reinterpret_cast<Xchg*>(h);
  SRL_CUDA(cudaMemcpyAsync(pinned_status_out, x->view.status, sizeof(int), cudaMemcpyDeviceToHost,
                           static_cast<cudaStream_t>(stream)));
  return SRL_OK;
}

extern "C" int srl_xchg_set_timeout(srl_xchg* h, double seconds) {
  using namespace srl;
  SRL_REQUIRE(h && seconds > 0.0, SRL_ERR_INVALID_ARG, "srl_xchg_set_timeout: need a handle and seconds > 0");
  Xchg* x = reinterpret_cast<Xchg*>(h);
  x->view.spin_limit = static_cast<long long>(seconds * 2.0e9);  // clock64 runs at the SM clock, ~2 GHz
  return SRL_OK;
}

extern "C" int srl_xchg_destroy(srl_xchg* h) {
  using namespace srl;
  if (h == nullptr) return SRL_OK;
  Xchg* x = reinterpret_cast<Xchg*>(h);
  for (int p = 0; p < kMaxWorld; ++p)
    if (x->peer_base[p]) cudaIpcCloseMemHandle(x->peer_base[p]);
  if (x->local_base) cudaFree(x->local_base);
  delete x;
  return SRL_OK;
}
