#!/usr/bin/env bash
# the driver's own commands on a fresh box, then an ncu --set full of the bf16 main pass (final code)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
echo "== pytest -m gpu"; timeout 1500 python -m pytest tests/ -x -q -m gpu 2>&1 | tail -3 | tee gpurun_out/test_all.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
echo "== bench.py --impl reference"; timeout 600 python bench.py --impl reference 2>&1 | tail -1 | tee gpurun_out/bench_reference.log | cut -c1-400
echo "== bench.py"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_default.log | cut -c1-300
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-fp32-line"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_topk -s 7 -c 1 -f -o gpurun_out/prof_bf16_r1e $CMD > gpurun_out/ncu2.log 2>&1; tail -1 gpurun_out/ncu2.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_topk|vote|prepare|merge|key_sim|rescore" -c 40 --csv --log-file gpurun_out/launches_bf16.csv $CMD > gpurun_out/ncu1.log 2>&1
