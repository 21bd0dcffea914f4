#!/usr/bin/env bash
# 8-GPU validation: bitwise check, bf16 (fused exchange) and fp32 benches
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
G=${1:-8}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== sharded check"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 scripts/sharded_check.py 2>&1 | grep -E "rank 0|Error|error|Traceback" | tail -20 | tee gpurun_out/sharded_check_g$G.log
F='"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"phases_ms": {[^}]*}'
echo "== bench bf16 gpus=$G"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --mode bf16 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_bf16_g${G}.log | grep -o "$F"
echo "== bench fp32 gpus=$G"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $G --mode fp32 --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_fp32_g${G}.log | grep -o "$F"
echo "== reference arm under torchrun"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29515 bench.py --impl reference --gpus $G --steps 2 --warmup 1 2>&1 | tail -1 | cut -c1-300 | tee gpurun_out/bench_reference_g${G}.log
