// K4 instantiations for the "pack" sample-side form with two lanes per thread (see ppo_loss.cuh, ppo_loss.cu:
// SRL_LOSS_LANES=2).  Same kernel template, LANES = 2: 64-bit policy-side loads and stores, two pack gathers per row.
#include "ppo_loss.cuh"

namespace srl {
namespace loss {
int launch_loss_pack2(LossBatch& b, int n_problems, cudaStream_t st) {
  return launch_loss_static<2, kPack>(b, n_problems, st);
}
}  // namespace loss
}  // namespace srl
