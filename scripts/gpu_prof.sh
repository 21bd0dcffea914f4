#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== experiments"; timeout 600 python scripts/exp_tc.py bf16 18944 2>&1 | tail -14 | tee gpurun_out/exp_bf16.log
CMD="python bench.py --mode bf16 --queries 18944 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_bf16.csv $CMD > gpurun_out/ncu1.log 2>&1
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_topk -s 1 -c 1 -o gpurun_out/prof_bf16 $CMD > gpurun_out/ncu2.log 2>&1
tail -3 gpurun_out/ncu2.log; ls -la gpurun_out
