#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== widen tests"; timeout 900 python -m pytest tests/test_gpu_widen.py -m gpu -x -q 2>&1 | tail -30 | tee gpurun_out/test_widen.log
echo "== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -6 | tee gpurun_out/test_all.log
echo "== bench default"; timeout 900 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_default.log
