// K4 instantiations for the "dense" sample-side form (see ppo_loss.cuh; one translation unit per form so the
// nine hyper-parameter configurations of each build in parallel).
#include "ppo_loss.cuh"

namespace srl {
namespace loss {
int launch_loss_dense(LossBatch& b, int n_problems, bool lanes4, cudaStream_t st) {
  return launch_loss_mode<kDense>(b, n_problems, lanes4, st);
}
}  // namespace loss
}  // namespace srl
