#!/usr/bin/env bash
# ncu --set full of one rescore_kernel launch (config c1, fp32 mode)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
CMD="python bench.py --mode fp32 --bank 138360 --queries 34590 --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || exit 1
ncu --set full --clock-control none --import-source on -k regex:rescore -s 3 -c 1 -f -o gpurun_out/prof_rescore_r1 $CMD > gpurun_out/ncu3.log 2>&1
tail -2 gpurun_out/ncu3.log
