#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== tc tests"; timeout 1500 python -m pytest tests/test_gpu_tc.py -m gpu -x -q 2>&1 | tail -40 | tee gpurun_out/test_tc.log
