// Host-side check of perm.cuh (the functions are __host__ __device__): perm_at is a permutation of [0, n_env) for every epoch
// key, and perm_pos_of is its inverse -- what the scan kernel relies on when it works out, per lane, the minibatch that lane
// falls into (gae_scan_ws.cu).  Built and run by tests/test_properties.py with nvcc; no GPU needed.
#include <cstdint>
#include <cstdio>
#include <vector>

#include "perm.cuh"

int main() {
  using namespace srl;
  long long bad = 0, checked = 0;
  const int sizes[] = {1, 2, 3, 34, 37, 100, 128, 512, 2048, 4096, 5000};
  for (int n_env : sizes) {
    for (uint32_t ep = 0; ep < 5; ++ep) {
      const PermKeys k = perm_keys(77u, 3u, ep, perm_bits(n_env));
      std::vector<int> seen(n_env, 0);
      for (uint32_t e = 0; e < static_cast<uint32_t>(n_env); ++e) {
        const uint32_t x = perm_at(e, static_cast<uint32_t>(n_env), k);
        ++checked;
        if (x >= static_cast<uint32_t>(n_env)) {
          ++bad;
          continue;
        }
        ++seen[x];
        if (perm_pos_of(x, static_cast<uint32_t>(n_env), k) != e) ++bad;
      }
      for (int v : seen) bad += v != 1;
    }
  }
  std::printf("checked=%lld bad=%lld\n", checked, bad);
  return bad != 0;
}
