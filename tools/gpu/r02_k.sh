#!/usr/bin/env bash
# Round 2, call K: why the sample-split sum differs from the whole render (debug)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
echo "--- default library"; timeout 200 python tools/debug/partition_diff.py 2>&1 | tail -8
echo "--- without the all-pixel loop copy"; RC_CUDA_LIB=$PWD/gpurun_in_noallpix.so timeout 200 python tools/debug/partition_diff.py 2>&1 | tail -8
