#!/usr/bin/env bash
# epilogue with software-pipelined TMEM loads: parity, then bf16 / fp32 benches at the north-star, c1, c2
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== tc+exact+rescore tests"; timeout 1500 python -m pytest tests/test_gpu_tc.py tests/test_gpu_exact.py tests/test_gpu_rescore.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/test_tc.log
F='"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"uncertified_rows_last_step": [0-9]*\|"frac": [0-9.]*\|"sm_mhz": [0-9.]*'
echo "== north-star bf16"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp32-line 2>&1 | tail -1 | tee gpurun_out/bench_bf16_g1.log | grep -o "$F"
echo "== north-star fp32"; timeout 600 python bench.py --mode fp32 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_fp32_g1_q75k.log | grep -o "$F"
echo "== c1 bf16"; timeout 600 python bench.py --bank 138360 --queries 34590 --steps 10 --warmup 3 --no-cpu-baseline --no-fp32-line 2>&1 | tail -1 | tee gpurun_out/bench_c1_bf16.log | grep -o "$F"
echo "== c1 fp32"; timeout 600 python bench.py --mode fp32 --bank 138360 --queries 34590 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c1_fp32.log | grep -o "$F"
echo "== c2 bf16"; timeout 600 python bench.py --bank 138360 --dim 384 --queries 34590 --steps 10 --warmup 3 --no-cpu-baseline --no-fp32-line 2>&1 | tail -1 | tee gpurun_out/bench_c2_bf16.log | grep -o "$F"
echo "== north-star bf16 again"; timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-fp32-line 2>&1 | tail -1 | tee gpurun_out/bench_bf16_g1_b.log | grep -o "$F"
