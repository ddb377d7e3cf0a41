#!/bin/bash
# tests + timeline (condensed) + same-box A/B against libsrl_head.so: r2q2.sh <tag>
TAG=${1:-x}
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3
timeout 300 python profiles/microbench/timeline.py --tag tl > gpurun_out/r2q_${TAG}_timeline.log 2>&1
grep -A75 "cold L2 (again)" gpurun_out/r2q_${TAG}_timeline.log | grep "^==\|K5a\|K2 scan  \|K4 loss\|scanner" | cut -c1-150
bash profiles/microbench/r2q_ab.sh $TAG srl_b200/libsrl_head.so srl_b200/libsrl_b200.so
