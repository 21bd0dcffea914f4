#!/usr/bin/env bash
# rescore rewrite: parity (both kernels vs seqfma oracle, fp32 modes), then c1 / north-star fp32 bench per chunk size
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== rescore tests"; timeout 600 python -m pytest tests/test_gpu_rescore.py -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/test_rescore.log
echo "== fp32 tc tests"; timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -k "fp32" 2>&1 | tail -8 | tee gpurun_out/test_tc_fp32.log
for c in 128 256; do
echo "== c1 fp32 chunk $c"; B200KNN_RESCORE_CHUNK=$c timeout 600 python bench.py --mode fp32 --bank 138360 --queries 34590 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c1_fp32_chunk$c.log | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"uncertified_rows_last_step": [0-9]*'
done
echo "== north-star fp32"; timeout 600 python bench.py --mode fp32 --queries 37888 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_fp32_g1.log | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"uncertified_rows_last_step": [0-9]*'
CMD="python bench.py --mode fp32 --bank 138360 --queries 34590 --steps 2 --warmup 3 --no-cpu-baseline"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:"rescore|tc_topk|vote|merge|prepare" -c 60 --csv --log-file gpurun_out/launches_fp32_c1.csv $CMD > gpurun_out/ncu1c.log 2>&1
grep -E "rescore" gpurun_out/launches_fp32_c1.csv | tail -8 | cut -c1-260
