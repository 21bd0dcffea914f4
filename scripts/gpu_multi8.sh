#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
G=${1:-8}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1
echo "== sharded check"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 scripts/sharded_check.py 2>&1 | grep -E "rank 0|Error|error|Traceback" | tail -20 | tee gpurun_out/sharded_check_g$G.log
echo "== bench bf16 gpus=$G fused (P2P stores)"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --mode bf16 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_bf16_g${G}_fused.log
echo "== bench bf16 gpus=$G NCCL all-to-all"
B200KNN_FUSED_EXCHANGE=0 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus $G --mode bf16 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_bf16_g${G}_nccl.log
echo "== bench fp32 gpus=$G"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $G --mode fp32 --steps 5 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_fp32_g${G}.log
