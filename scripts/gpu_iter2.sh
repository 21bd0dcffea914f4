#!/usr/bin/env bash
# iteration call: tc tests in their own process, short benches, pipeline ablation
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== tc tests"; timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -x 2>&1 | tail -25 | tee gpurun_out/test_tc.log
echo "== exact tests"; timeout 900 python -m pytest tests/test_gpu_exact.py -m gpu -q -x 2>&1 | tail -5 | tee gpurun_out/test_exact.log
echo "== bench bf16"; timeout 600 python bench.py --mode bf16 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_bf16.log
echo "== bench bf16 1-CTA"; B200KNN_NO_PAIR=1 timeout 600 python bench.py --mode bf16 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_bf16_nopair.log


echo "== prepass exp"; timeout 600 python scripts/exp_prepass.py 2>&1 | tail -8 | tee gpurun_out/exp_prepass.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -k regex:"tc_topk|vote|prepare|decode|merge|rescore" -c 60 --csv --log-file gpurun_out/launches_bf16.csv $CMD > gpurun_out/ncu1.log 2>&1
