"""Where MultiAgentPPOB200.step's wall clock goes (bench.py's trainer_step workload): cProfile of 3 calls + CUDA time."""
import cProfile, io, os, pstats, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import torch
import bench
from srl_b200 import synth
cfg = synth.CONFIGS["cfg2_atari_large"]
dev = torch.device("cuda", 0)
# reuse bench's own workload builder by running it under the profiler
pr = cProfile.Profile()
pr.enable()
line = bench.trainer_step_run(cfg, dev, steps=3)
pr.disable()
print({k: v for k, v in line.items() if k != "what"})
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue()[:9000])
