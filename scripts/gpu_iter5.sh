#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q -s 2>&1 | grep -E "uncertified|passed|failed|Error|error|assert" | tail -60 | tee gpurun_out/test_all.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
for m in fp32 bf16x3 fp32_tf32; do
echo "== bench $m"; timeout 600 python bench.py --mode $m --queries 37888 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_$m.log
done
CMD="python bench.py --mode fp32 --queries 37888 --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_topk|vote|prepare|decode|merge|rescore|exact|norm|key_sim" -c 40 --csv --log-file gpurun_out/launches_fp32.csv $CMD > gpurun_out/ncu1.log 2>&1
