#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/test_all.log
echo "== bench default"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_default.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_topk|vote|prepare|decode|merge|rescore|exact|norm|key_sim" -c 40 --csv --log-file gpurun_out/launches_bf16.csv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_topk -s 7 -c 1 -f -o gpurun_out/prof_bf16_r1d $CMD > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
