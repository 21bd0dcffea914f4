#!/usr/bin/env bash
# N-GPU: bitwise check + fp32 bench with the level-by-level sharded cascade
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
G=${1:-2}
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?"
echo "== sharded check"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29511 scripts/sharded_check.py > gpurun_out/sharded_check_g$G.log 2>&1; echo "rc=$? True: $(grep -o 'sharded==single True' gpurun_out/sharded_check_g$G.log | wc -l) False: $(grep -o 'sharded==single False' gpurun_out/sharded_check_g$G.log | wc -l)"; grep -E "Error|Traceback|assert" gpurun_out/sharded_check_g$G.log | head -5
F='"value": [0-9.]*\|"ms_per_step": [0-9.]*\|"kernel_ms": [0-9.]*\|"phases_ms": {[^}]*}\|"uncertified_rows_last_step": [0-9]*'
echo "== bench fp32 gpus=$G"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $G --master-addr 127.0.0.1 --master-port 29514 bench.py --gpus $G --mode fp32 --steps 10 --warmup 3 2>&1 | tail -1 | tee gpurun_out/bench_fp32_g${G}.log | grep -o "$F"
