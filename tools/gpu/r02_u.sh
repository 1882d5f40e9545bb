#!/usr/bin/env bash
# Round 2, call U: v26 register caps (CTAs per SM) and the stealing threshold
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() {  # label, env...
    local label=$1; shift
    env "$@" timeout 150 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02u_$label.json 2> gpurun_out/r02u_$label.err
    python - "$label" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r02u_{sys.argv[1]}.json"))
    print(f"{sys.argv[1]:>16}: {d['value']:.4e} samples/s  {d['ms_per_step']:.3f} ms")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run base RC_X=0
run mb4 RC_MIN_BLOCKS=4
run mb5 RC_MIN_BLOCKS=5
run mb7 RC_MIN_BLOCKS=7
run mb8 RC_MIN_BLOCKS=8
run mb9 RC_MIN_BLOCKS=9
run mb10 RC_MIN_BLOCKS=10
run smin2 RC_STEAL_MIN=2
run smin8 RC_STEAL_MIN=8
run smin16 RC_STEAL_MIN=16
run base2 RC_X=0
