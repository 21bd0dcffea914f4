#!/usr/bin/env bash
# One gpurun call: parity tests, then the tensor-core tests in their own process (a trap
# there must not take the exact-mode results with it), then short bench runs.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== exact tests"; timeout 900 python -m pytest tests/test_gpu_exact.py -m gpu -x -q 2>&1 | tail -25 | tee gpurun_out/test_exact.log
echo "== tc tests"; timeout 900 python -m pytest tests/test_gpu_tc.py -m gpu -q -s 2>&1 | tail -60 | tee gpurun_out/test_tc.log
echo "== bench exact"; timeout 600 python bench.py --mode exact --queries 18944 --steps 2 --warmup 1 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_exact.log
echo "== bench bf16"; timeout 600 python bench.py --mode bf16 --steps 5 --warmup 3 2>&1 | tail -3 | tee gpurun_out/bench_bf16.log
echo "== bench tf32x3"; timeout 600 python bench.py --mode tf32x3 --queries 18944 --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -3 | tee gpurun_out/bench_tf32x3.log
