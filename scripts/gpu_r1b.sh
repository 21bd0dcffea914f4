#!/usr/bin/env bash
# Round-1 measurement call: GPU parity tests, default bench (+cpu baseline), reference arm,
# B=64 reference-shaped call, ncu launch list and one --set full capture of the main kernel.
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -15 | tee gpurun_out/test_all.log
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -3 | tee gpurun_out/smoke.log
echo "== bench default"; timeout 900 python bench.py 2>&1 | tail -1 | tee gpurun_out/bench_default.log
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_reference.log
echo "== bench B=64"; timeout 300 python bench.py --queries 64 --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_b64.log
echo "== bench B=1024"; timeout 300 python bench.py --queries 1024 --steps 50 --warmup 5 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_b1024.log
echo "== bench fp32_bf16"; timeout 300 python bench.py --mode fp32_bf16 --queries 18944 --steps 3 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_fp32_bf16.log
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_bf16.csv $CMD > gpurun_out/ncu1.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:tc_topk -s 7 -c 1 -f -o gpurun_out/prof_bf16_r1b $CMD > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log; ls -la gpurun_out | head -40
