#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
G=${1:-2}
nvidia-smi -L | head -3
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== sharded check"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 scripts/sharded_check.py 2>&1 | grep -E "rank 0|Error|error|Traceback" | tail -20 | tee gpurun_out/sharded_check_g$G.log
for n in ${2:-1 $G}; do
  echo "== bench bf16 gpus=$n"
  if [ $n -eq 1 ]; then timeout 600 python bench.py --gpus 1 --mode bf16 --steps 5 --warmup 3 --no-cpu-baseline 2>&1 | tail -1 | tee gpurun_out/bench_bf16_g1.log
  else timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n --mode bf16 --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_bf16_g$n.log; fi
done
