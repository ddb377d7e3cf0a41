"""srl_b200 -- B200-native (sm_100a) trainer-side sample-batch hot path for SRL.

Scope: batch assembly gather, GAE / returns scan, Philox-keyed minibatch permutation
+ gather, fused PPO/MAPPO loss forward+backward, PopArt statistics -- behind the reference's
`api.trainer.Trainer` plugin interface (`srl_b200.trainer.MultiAgentPPOB200`).

The compute path is hand-written CUDA reached through the C-ABI in `include/srl_b200.h`
(`libsrl_b200.so`, loaded with ctypes by `srl_b200._lib`).  There is no CPU fallback: every op raises
`srl_b200.MissingCudaLibrary` if the library has not been built.
"""
from srl_b200._lib import MissingCudaLibrary, SrlCudaError, lib_path, load_library  # noqa: F401

__version__ = "0.1.0"
