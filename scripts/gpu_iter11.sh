#!/usr/bin/env bash
# F16X2 candidate mode: parity tests, then fp32 benches (c1, north-star) + bf16 default with quality block
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== f16x2 tile tests"; timeout 600 python -m pytest tests/test_gpu_tc.py -m gpu -x -q -k "f16x2" 2>&1 | tail -15 | tee gpurun_out/test_tc_f16x2.log
echo "== all tc+rescore tests"; timeout 1500 python -m pytest tests/test_gpu_tc.py tests/test_gpu_rescore.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/test_tc.log
F='"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"uncertified_rows_last_step": [0-9]*\|"quality": {[^}]*}\|"executed_frac": [0-9.]*'
echo "== c1 fp32"; timeout 600 python bench.py --mode fp32 --bank 138360 --queries 34590 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c1_fp32.log | grep -o "$F"
echo "== north-star fp32"; timeout 600 python bench.py --mode fp32 --queries 37888 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_fp32_g1.log | grep -o "$F"
echo "== north-star fp32 Q=75776"; timeout 600 python bench.py --mode fp32 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_fp32_g1_q75k.log | grep -o "$F"
echo "== north-star f16x2 raw"; timeout 600 python bench.py --mode f16x2 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_f16x2_g1.log | grep -o "$F"
echo "== north-star bf16"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_bf16_g1.log | grep -o "$F"
