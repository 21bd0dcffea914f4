#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/test_all.log
echo "== bench default"; timeout 900 python bench.py --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_default.log
echo "== bench B=64"; timeout 300 python bench.py --queries 64 --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_b64.log
echo "== bench B=1024"; timeout 300 python bench.py --queries 1024 --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_b1024.log
