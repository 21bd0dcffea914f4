#!/usr/bin/env bash
# One parameterised launcher for everything that runs on the GPU box (under gpurun):
#   scripts/gpu.sh build                      build the library + oracle, print rc
#   scripts/gpu.sh test [pytest args]         pytest -m gpu (default: all GPU tests)
#   scripts/gpu.sh smoke
#   scripts/gpu.sh bench NAME [bench args]    python bench.py ... -> gpurun_out/bench_NAME.json
#   scripts/gpu.sh mbench G NAME [bench args] torchrun --nproc-per-node G bench.py ...
#   scripts/gpu.sh launches NAME [bench args] ncu launch list (gpu__time_duration) of a short bench run
#   scripts/gpu.sh ncu NAME REGEX SKIP [bench args]   ncu --set full of one launch matching REGEX
#   scripts/gpu.sh sanitizer TOOL             compute-sanitizer (memcheck|racecheck) on the small cases
#   scripts/gpu.sh sharded G                  scripts/sharded_check.py on G GPUs
# Several steps can be chained in one gpurun call:  scripts/gpu.sh build -- test -- bench default
cd "${GRAFT_REPO_ROOT:-/root/repo}"
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
run_one() {
  local cmd="$1"; shift
  case "$cmd" in
    build)
      python -c "import __graft_entry__ as g; g.build()" > gpurun_out/build.log 2>&1; echo "build rc=$?" ;;
    test)
      echo "== pytest -m gpu $*"
      timeout 2400 python -m pytest tests/ -x -q -m gpu "$@" 2>&1 | tail -15 | tee gpurun_out/test.log ;;
    smoke)
      timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2 ;;
    bench)
      local name="$1"; shift
      echo "== bench $name: $*"
      timeout 1200 python bench.py "$@" 2> gpurun_out/bench_$name.err | tail -1 > gpurun_out/bench_$name.json
      echo "rc=$?"; cut -c1-1500 gpurun_out/bench_$name.json; tail -3 gpurun_out/bench_$name.err ;;
    mbench)
      local g="$1" name="$2"; shift 2
      echo "== bench $name on $g GPUs: $*"
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$g" --master-addr 127.0.0.1 \
        --master-port $((29500 + RANDOM % 400)) bench.py --gpus "$g" "$@" 2> gpurun_out/bench_$name.err | tail -1 > gpurun_out/bench_$name.json
      echo "rc=$?"; cut -c1-1500 gpurun_out/bench_$name.json; grep -v "^W\|^\*\*\*\|OMP_NUM" gpurun_out/bench_$name.err | tail -5 ;;
    launches)
      local name="$1"; shift
      timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv \
        --log-file gpurun_out/launches_$name.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library "$@" \
        > gpurun_out/launches_$name.log 2>&1; echo "launches rc=$?" ;;
    ncu)
      local name="$1" regex="$2" skip="$3"; shift 3
      timeout 1200 ncu --set full --clock-control none --import-source on -k regex:"$regex" -s "$skip" -c 1 -f \
        -o gpurun_out/prof_$name python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-library "$@" \
        > gpurun_out/ncu_$name.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/ncu_$name.log ;;
    sanitizer)
      local tool="$1"; shift
      echo "== compute-sanitizer --tool $tool"
      timeout 2400 compute-sanitizer --tool "$tool" --error-exitcode 7 python scripts/sanitizer_cases.py "$@" \
        > gpurun_out/sanitizer_$tool.log 2>&1; echo "sanitizer $tool rc=$?"; tail -6 gpurun_out/sanitizer_$tool.log ;;
    sharded)
      local g="$1"; shift
      timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node "$g" --master-addr 127.0.0.1 \
        --master-port $((29500 + RANDOM % 400)) scripts/sharded_check.py "$@" > gpurun_out/sharded_check_${g}gpu.log 2>&1
      echo "sharded_check rc=$?"; grep -c "sharded==single True" gpurun_out/sharded_check_${g}gpu.log; grep -v "True" gpurun_out/sharded_check_${g}gpu.log | grep -v "^W\|^\*\*\*\|OMP_NUM" | tail -8 ;;
    py)
      timeout 1200 python "$@" 2>&1 | tail -40 ;;
    *) echo "unknown step $cmd"; return 2 ;;
  esac
}
args=()
for a in "$@"; do
  if [ "$a" == "--" ]; then run_one "${args[@]}"; args=(); else args+=("$a"); fi
done
[ ${#args[@]} -gt 0 ] && run_one "${args[@]}"
