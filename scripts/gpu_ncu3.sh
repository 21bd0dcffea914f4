#!/usr/bin/env bash
# ncu --set full of the main bf16 tc_topk launch at config c1 (N=138,360: candidate-bound epilogue)
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
CMD="python bench.py --bank 138360 --queries 34590 --steps 2 --warmup 3 --no-cpu-baseline --no-fp32-line"
$CMD > gpurun_out/plain.log 2>&1 || exit 1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_topk -s 7 -c 1 -f -o gpurun_out/prof_bf16_c1 $CMD > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
