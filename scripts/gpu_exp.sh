#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
L=self-supervised-wafermaps_b200/lib
for v in e1 e4 ""; do
  lib=$L/libb200knn${v:+_$v}.so
  echo "== variant ${v:-e2} ($lib)"
  B200KNN_LIB=$PWD/$lib timeout 300 python scripts/exp_tc.py bf16 18944 2>&1 | grep -E "flags= (0|1) " | tee -a gpurun_out/exp_variants.log
  B200KNN_LIB=$PWD/$lib timeout 300 python bench.py --mode bf16 --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('bench', d['value'], d['roofline']['frac'])" | tee -a gpurun_out/exp_variants.log
done
CMD="python bench.py --mode bf16 --queries 18944 --steps 1 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:tc_topk -s 1 -c 1 -o gpurun_out/prof_bf16_v2 $CMD > gpurun_out/ncu2.log 2>&1
tail -2 gpurun_out/ncu2.log
