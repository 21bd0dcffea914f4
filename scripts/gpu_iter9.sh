#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== bench default"; timeout 900 python bench.py --steps 10 --no-cpu-baseline 2>&1 | tail -1 | cut -c1-200 | tee gpurun_out/bench_default.log
echo "== B=64 pair"; timeout 300 python bench.py --queries 64 --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"plan": {[^}]*}' | tee gpurun_out/b64_pair.log
echo "== B=64 1-CTA"; B200KNN_NO_PAIR=1 timeout 300 python bench.py --queries 64 --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | grep -o '"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"plan": {[^}]*}' | tee gpurun_out/b64_1cta.log
echo "== c5 shard shape: N=2097152 D=768 Q=65536 bf16"; timeout 600 python bench.py --bank 2097152 --dim 768 --queries 65536 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c5shard_bf16.log | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"plan": {[^}]*}\|"achieved": [0-9.]*'
echo "== c2: D=384 N=138360 Q=34590"; timeout 600 python bench.py --bank 138360 --dim 384 --queries 34590 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c2_bf16.log | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"plan": {[^}]*}\|"achieved": [0-9.]*'
echo "== c1 fp32: D=512 N=138360 Q=34590"; timeout 600 python bench.py --mode fp32 --bank 138360 --queries 34590 --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_c1_fp32.log | grep -o '"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"plan": {[^}]*}\|"achieved": [0-9.]*'
CMD="python bench.py --queries 64 --steps 3 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:"tc_topk|vote|prepare|merge|key_sim" -c 30 --csv --log-file gpurun_out/launches_b64.csv $CMD > gpurun_out/ncu1b.log 2>&1
