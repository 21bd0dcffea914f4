#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
G=${1:-1}
echo "== all gpu tests"; timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -6 | tee gpurun_out/test_all.log
echo "== bench bf16 g1"; timeout 600 python bench.py --mode bf16 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_bf16_g1.log
echo "== bench bf16 g1 noprepass"; B200KNN_PREPASS=0 timeout 600 python bench.py --mode bf16 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_bf16_g1_nopre.log
echo "== bench fp32 g1"; timeout 600 python bench.py --mode fp32 --queries 18944 --steps 3 --warmup 2 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_fp32_g1.log
if [ $G -gt 1 ]; then
  echo "== sharded check"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 scripts/sharded_check.py 2>&1 | grep -E "rank 0|Error|error|Traceback" | tail -12 | tee gpurun_out/sharded_check.log
  echo "== bench bf16 g$G"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --mode bf16 --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_bf16_g$G.log
fi
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/bench_*g*.log")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split('/')[-1], round(d["value"]), "ms/step",round(d["ms_per_step"],2), "kern", round(d["roofline"]["kernel_ms"],2), "frac", round(d["roofline"]["frac"],3), "e2e", round(d["e2e"]["value"]), d["clocks"].get("sm_mhz"), d["config"].get("prepass"), d["config"].get("uncertified_rows_last_step"))
    except Exception as e: print(f, "ERR", e)
PY
