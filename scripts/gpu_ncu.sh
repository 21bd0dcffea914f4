#!/usr/bin/env bash
# ncu --set full capture of the main tc_topk launch (8th tc_topk launch = main pass of step 4)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:tc_topk -s 7 -c 1 -f -o gpurun_out/${1:-prof_bf16} $CMD > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
