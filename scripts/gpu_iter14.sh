#!/usr/bin/env bash
# F16 first cascade level: parity (tc + rescore + widen tests), fp32 benches at north-star / c1, default line
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/test_all.log
F='"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"uncertified_rows_last_step": [0-9]*\|"frac": [0-9.]*\|"keys_bitwise_equal_exact_mode": [a-z]*'
echo "== north-star fp32 (f16 -> f16x2 cascade)"; timeout 600 python bench.py --mode fp32 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_fp32_g1_q75k.log | grep -o "$F"
echo "== north-star fp32_f16x2 (previous default)"; timeout 600 python bench.py --mode fp32_f16x2 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_fp32_f16x2_g1_q75k.log | grep -o "$F"
echo "== c1 fp32"; timeout 600 python bench.py --mode fp32 --bank 138360 --queries 34590 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c1_fp32.log | grep -o "$F"
echo "== north-star f16 raw"; timeout 600 python bench.py --mode f16 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_f16_g1.log | grep -o "$F\|\"recall_at_k\": [0-9.]*"
