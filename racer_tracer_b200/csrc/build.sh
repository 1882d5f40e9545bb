#!/usr/bin/env bash
# Builds libracer_cuda.so (the C-ABI shared library) for sm_100a, in-tree.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
OUT="${HERE}/../libracer_cuda.so"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
"${NVCC}" -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 \
    -Xptxas -v -Xcompiler -fPIC -Xcompiler -fvisibility=hidden \
    -shared -o "${OUT}" "${HERE}/rc_api.cu" -lcudart -ldl "$@"
echo "built ${OUT}"
