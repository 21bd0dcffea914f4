#!/usr/bin/env bash
# class-nested candidate append: parity + c1 / north-star benches
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 | tee gpurun_out/test_all.log
F='"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"uncertified_rows_last_step": [0-9]*\|"frac": [0-9.]*'
echo "== c1 bf16"; timeout 600 python bench.py --bank 138360 --queries 34590 --steps 10 --warmup 3 --no-cpu-baseline --no-fp32-line 2>&1 | tail -1 | tee gpurun_out/bench_c1_bf16.log | grep -o "$F"
echo "== c1 fp32"; timeout 600 python bench.py --mode fp32 --bank 138360 --queries 34590 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c1_fp32.log | grep -o "$F"
echo "== north-star default line"; timeout 900 python bench.py --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_default.log | grep -o "$F\|\"fp32_mode\": {[^}]*}"
