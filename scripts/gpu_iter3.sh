#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/test_all.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== bench default"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_default.log
echo "== bench fp32"; timeout 600 python bench.py --mode fp32 --queries 37888 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_fp32.log
echo "== bench B=64"; timeout 300 python bench.py --queries 64 --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_b64.log
CMD="python bench.py --mode fp32 --queries 18944 --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_topk|vote|prepare|decode|merge|rescore|exact|norm" -c 40 --csv --log-file gpurun_out/launches_fp32.csv $CMD > gpurun_out/ncu1.log 2>&1
CMD="python bench.py --queries 64 --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_topk|vote|prepare|decode|merge|rescore|exact|norm" -c 40 --csv --log-file gpurun_out/launches_b64.csv $CMD > gpurun_out/ncu1b.log 2>&1
