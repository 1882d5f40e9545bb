#!/usr/bin/env bash
# Round 2, call W: harmless perturbations of the generated kernel on one box (ptxas schedules differ by +-2 %)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
run() {  # label, env...
    local label=$1; shift
    env "$@" timeout 150 python bench.py --steps 8 --warmup 3 --no-cpu-baseline > gpurun_out/r02w_$label.json 2> gpurun_out/r02w_$label.err
    python - "$label" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r02w_{sys.argv[1]}.json"))
    print(f"{sys.argv[1]:>22}: {d['value']:.4e} samples/s  {d['ms_per_step']:.3f} ms")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run base RC_X=0
run philox_late RC_PHILOX_LATE=1
run philox_late_norange RC_PHILOX_LATE=1 RC_FIRST_TEST_RANGE=0
run kc3 RC_KCONST_MASK=3
run kc5 RC_KCONST_MASK=5
run kc6 RC_KCONST_MASK=6
run kc0 RC_KCONST_MASK=0
run kc6_norange RC_KCONST_MASK=6 RC_FIRST_TEST_RANGE=0
run noregc_norange RC_SPEC_NO_REG_CONSTS=1 RC_FIRST_TEST_RANGE=0
run mb5_norange RC_MIN_BLOCKS=5 RC_FIRST_TEST_RANGE=0
run base2 RC_X=0
