#!/usr/bin/env bash
# full GPU test-suite + smoke, then ncu --set full of the fp32 mode's two dominant kernels (F16X2 main pass, rescore_dot)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
echo "== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 | tee gpurun_out/test_all.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 | tee gpurun_out/smoke.log
CMD="python bench.py --mode fp32 --queries 37888 --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain.log 2>&1 || exit 1
# per step: SAMPLE tc_topk + main tc_topk -> 8th tc_topk launch = main pass of step 4; rescore_dot: 4th launch
timeout 600 ncu --set full --clock-control none --import-source on -k regex:tc_topk -s 7 -c 1 -f -o gpurun_out/prof_f16x2_r1 $CMD > gpurun_out/ncu4.log 2>&1; tail -1 gpurun_out/ncu4.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:rescore_dot -s 3 -c 1 -f -o gpurun_out/prof_rescore_dot_r1 $CMD > gpurun_out/ncu5.log 2>&1; tail -1 gpurun_out/ncu5.log
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_fp32.csv $CMD > gpurun_out/ncu6.log 2>&1
