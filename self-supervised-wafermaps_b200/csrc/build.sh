#!/usr/bin/env bash
# Builds libb200knn.so in-tree for sm_100a (cross-compiles without a GPU).
set -euo pipefail
here="$(cd "$(dirname "$0")" && pwd)"
out="$here/../lib"
mkdir -p "$out"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
# EXTRA_NVCC_FLAGS / LIBNAME: experiment variants (e.g. -DB200KNN_EPI_PER_QUARTER=4)
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC
       --expt-relaxed-constexpr -Xptxas -v ${EXTRA_NVCC_FLAGS:-})
LIBNAME="${LIBNAME:-libb200knn.so}"
BUILD="${BUILD_DIR:-build}"
mkdir -p "$here/$BUILD"
pids=()
for f in api exact select vote prepare rescore tc_topk; do
  ( "$NVCC" "${FLAGS[@]}" -c "$here/$f.cu" -o "$here/$BUILD/$f.o" > "$here/$BUILD/$f.log" 2>&1 ) &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=1; done
if [ $rc -ne 0 ]; then
  grep -h -E "error|Error" "$here"/$BUILD/*.log || true
  exit 1
fi
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$out/$LIBNAME" \
  "$here"/$BUILD/{api,exact,select,vote,prepare,rescore,tc_topk}.o -cudart static
echo "built $out/$LIBNAME"
