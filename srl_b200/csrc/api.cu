// Library-level plumbing: status/error text, device queries.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <atomic>

#include "common.cuh"

namespace srl {
namespace {
thread_local char g_err[512] = "";
}

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

namespace {
std::atomic<int> g_pdl{-1};  // -1: not decided yet (SRL_PDL in the environment, default on)
}

bool pdl_enabled() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("SRL_PDL");
    v = (e == nullptr || e[0] != '0') ? 1 : 0;
    g_pdl.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}

namespace {
thread_local cudaStream_t g_note_stream = nullptr;
thread_local int g_note_kind = 0;  // the last library launch of this thread: 0 anything else, 1 permutation, 2 scan
bool take(cudaStream_t st, int kind) {
  const bool yes = g_note_kind == kind && g_note_stream == st;
  g_note_kind = 0;
  return yes;
}
}  // namespace

void pdl_note_perm(cudaStream_t st) {
  g_note_stream = st;
  g_note_kind = 1;
}
bool pdl_take_perm(cudaStream_t st) { return take(st, 1); }
void pdl_note_scan(cudaStream_t st) {
  g_note_stream = st;
  g_note_kind = 2;
}
bool pdl_take_scan(cudaStream_t st) { return take(st, 2); }

void pdl_forget() { g_note_kind = 0; }

int sm_count() {
  static int cached[64];
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}
}  // namespace srl

extern "C" const char* srl_last_error(void) { return srl::g_err; }

extern "C" int srl_abi_version(void) { return SRL_B200_ABI_VERSION; }

extern "C" int srl_pdl_enabled(void) { return srl::pdl_enabled() ? 1 : 0; }

extern "C" int srl_set_pdl(int on) {
  const int before = srl::pdl_enabled() ? 1 : 0;
  srl::g_pdl.store(on ? 1 : 0, std::memory_order_relaxed);
  return before;
}

extern "C" int srl_device_info(int* sm_count, int* cc_major, int* cc_minor) {
  using namespace srl;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  SRL_REQUIRE(e == cudaSuccess && n > 0, SRL_ERR_NO_DEVICE, "no CUDA device visible (%s)", cudaGetErrorString(e));
  int dev = 0;
  SRL_CUDA(cudaGetDevice(&dev));
  int sms = 0, maj = 0, min = 0;
  SRL_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  SRL_CUDA(cudaDeviceGetAttribute(&maj, cudaDevAttrComputeCapabilityMajor, dev));
  SRL_CUDA(cudaDeviceGetAttribute(&min, cudaDevAttrComputeCapabilityMinor, dev));
  if (sm_count) *sm_count = sms;
  if (cc_major) *cc_major = maj;
  if (cc_minor) *cc_minor = min;
  return SRL_OK;
}
