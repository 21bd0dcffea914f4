#!/usr/bin/env bash
# Builds libb200knn.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
out="$here/../lib"
mkdir -p "$out" "$here/build"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
       --expt-relaxed-constexpr -Xptxas -v)
pids=()
for f in api exact select vote prepare tc_topk; do
  ( "$NVCC" "${FLAGS[@]}" -c "$here/$f.cu" -o "$here/build/$f.o" > "$here/build/$f.log" 2>&1 ) &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=1; done
if [ $rc -ne 0 ]; then
  grep -h -E "error|Error" "$here"/build/*.log || true
  exit 1
fi
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$out/libb200knn.so" \
  "$here"/build/{api,exact,select,vote,prepare,tc_topk}.o -cudart static
echo "built $out/libb200knn.so"
