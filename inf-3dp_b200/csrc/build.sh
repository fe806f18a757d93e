#!/bin/bash
# Build libinf3dp.so (sm_100a only) next to the Python package.  nvcc cross-compiles without a GPU.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../inf3dp/libinf3dp.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xptxas -v
       --expt-relaxed-constexpr)
mkdir -p "${HERE}/build"
pids=()
for f in api siren_simt siren_tc siren_rev contours intersection volume collision; do
  ( "${NVCC}" "${FLAGS[@]}" -c "${HERE}/${f}.cu" -o "${HERE}/build/${f}.o" > "${HERE}/build/${f}.log" 2>&1 ) &
  pids+=($!)
done
rc=0
for p in "${pids[@]}"; do wait "$p" || rc=1; done
if [ $rc -ne 0 ]; then cat "${HERE}"/build/*.log; exit 1; fi
"${NVCC}" -shared -o "${OUT}" "${HERE}"/build/{api,siren_simt,siren_tc,siren_rev,contours,intersection,volume,collision}.o -lcudart_static -lpthread -ldl -lrt
echo "built ${OUT}"
