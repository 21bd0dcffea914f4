#!/usr/bin/env bash
# quick iteration: parity tests, ablation experiments, bench
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
echo "== exact tests"; timeout 900 python -m pytest tests/test_gpu_exact.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/test_exact.log
echo "== tc tests"; timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -s 2>&1 | tail -30 | tee gpurun_out/test_tc.log
echo "== experiments"; timeout 600 python scripts/exp_tc.py bf16 18944 2>&1 | tail -12 | tee gpurun_out/exp_bf16.log
echo "== bench bf16"; timeout 600 python bench.py --mode bf16 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_bf16.log
echo "== bench tf32x3"; timeout 600 python bench.py --mode tf32x3 --queries 18944 --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -2 | tee gpurun_out/bench_tf32x3.log
