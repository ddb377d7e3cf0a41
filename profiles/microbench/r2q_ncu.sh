#!/bin/bash
# ncu capture of one kernel of the cfg2 step with dense PC sampling: r2q_ncu.sh <kernel regex> <tag>
cd "$(dirname "$0")/../.."
K=${1:-gae_scan_ws}; TAG=${2:-k2}
ncu --set full --import-source on --clock-control none --warp-sampling-interval 0 --kernel-name regex:$K --launch-skip 4 --launch-count 1 \
    -f -o gpurun_out/r2q_ncu_$TAG python bench.py --steps 4 --warmup 3 --e2e-steps 3 --no-cpu-baseline --no-extras --no-parity-check > gpurun_out/r2q_ncu_$TAG.log 2>&1
tail -3 gpurun_out/r2q_ncu_$TAG.log | cut -c1-200
ls -la gpurun_out/r2q_ncu_$TAG.ncu-rep
