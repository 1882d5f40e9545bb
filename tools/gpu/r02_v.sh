#!/usr/bin/env bash
# Round 2, call V: A/B on one box — the first rectangle test with / without its t <= RT_NO_HIT compare
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() {  # label, env...
    local label=$1; shift
    env "$@" timeout 150 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02v_$label.json 2> gpurun_out/r02v_$label.err
    python - "$label" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r02v_{sys.argv[1]}.json"))
    print(f"{sys.argv[1]:>16}: {d['value']:.4e} samples/s  {d['ms_per_step']:.3f} ms")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run norange RC_FIRST_TEST_RANGE=0
run range RC_FIRST_TEST_RANGE=1
run norange_mb5 RC_FIRST_TEST_RANGE=0 RC_MIN_BLOCKS=5
run norange_mb8 RC_FIRST_TEST_RANGE=0 RC_MIN_BLOCKS=8
run norange2 RC_FIRST_TEST_RANGE=0
run range2 RC_FIRST_TEST_RANGE=1
