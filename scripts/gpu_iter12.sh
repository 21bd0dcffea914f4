#!/usr/bin/env bash
# merge (multi-warp rows) parity + B=64 call; default bench line with quality + fp32_mode; configs c2..c4 on 1 GPU
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== exact/merge tests"; timeout 900 python -m pytest tests/test_gpu_exact.py -m gpu -x -q 2>&1 | tail -8 | tee gpurun_out/test_exact.log
F='"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"uncertified_rows_last_step": [0-9]*\|"quality": {[^}]*}\|"frac": [0-9.]*\|"fp32_mode": {[^}]*}'
echo "== B=64 bf16"; timeout 300 python bench.py --queries 64 --steps 50 --warmup 5 --no-cpu-baseline --no-fp32-line 2>&1 | tail -1 | tee gpurun_out/bench_b64.log | grep -o "$F"
echo "== B=64 fp32"; timeout 300 python bench.py --mode fp32 --queries 64 --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_b64_fp32.log | grep -o "$F"
echo "== default"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_default.log | grep -o "$F"
for m in bf16 fp32; do
echo "== c2 $m: D=384 N=138360 Q=34590"; timeout 600 python bench.py --mode $m --bank 138360 --dim 384 --queries 34590 --steps 10 --warmup 3 --no-cpu-baseline --no-fp32-line 2>&1 | tail -1 | tee gpurun_out/bench_c2_$m.log | grep -o "$F"
echo "== c3 $m: 38 classes N=30000 Q=5703 k=20"; timeout 600 python bench.py --mode $m --bank 30000 --queries 5703 --k 20 --classes 38 --steps 20 --warmup 3 --no-cpu-baseline --no-fp32-line 2>&1 | tail -1 | tee gpurun_out/bench_c3_$m.log | grep -o "$F"
echo "== c4 $m: N=Q=811457 k=10"; timeout 900 python bench.py --mode $m --queries 811457 --k 10 --steps 2 --warmup 1 --no-cpu-baseline --no-fp32-line 2>&1 | tail -1 | tee gpurun_out/bench_c4_$m.log | grep -o "$F"
done
CMD="python bench.py --queries 64 --steps 3 --warmup 3 --no-cpu-baseline --no-fp32-line"
timeout 600 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum --clock-control none -k regex:"tc_topk|vote|prepare|merge|key_sim" -c 30 --csv --log-file gpurun_out/launches_b64.csv $CMD > gpurun_out/ncu1b.log 2>&1
grep -E "merge" gpurun_out/launches_b64.csv | grep gpu__time | tail -4 | cut -c1-30,100-260
