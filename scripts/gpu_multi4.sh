#!/usr/bin/env bash
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
G=${1:-4}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
F='"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"phases_ms": {[^}]*}\|"uncertified_rows_last_step": [0-9]*'
echo "== bench bf16 gpus=$G (driver command line)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $G --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_bf16_g${G}.log | grep -o "$F"
echo "== bench fp32 gpus=$G"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $G --mode fp32 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_fp32_g${G}.log | grep -o "$F"
