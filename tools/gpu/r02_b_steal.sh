#!/usr/bin/env bash
# Round 2, call B: in-warp sample stealing — parity suite, then A/B timings (steal off/on, sample slices at N=1,
# CTAs per SM).  Every line is one `bench.py` run (CUDA events, 256 MiB L2 flush between steps).
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r02b_pytest.log 2>&1
echo "pytest rc=$?"; tail -5 gpurun_out/r02b_pytest.log
run() {  # label, env...
    local label=$1; shift
    env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02b_$label.json 2> gpurun_out/r02b_$label.err
    python - "$label" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r02b_{sys.argv[1]}.json"))
    print(f"{sys.argv[1]:>16}: {d['value']:.4e} samples/s  {d['ms_per_step']:.3f} ms  e2e {d['e2e']['value']:.4e}  launches {d['gpu_launches']}  kernel {d['config']['kernel']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run steal0 RC_STEAL=0
run steal1 RC_STEAL=1
run steal1_sl2 RC_STEAL=1 RC_SLICES=2
run steal1_sl3 RC_STEAL=1 RC_SLICES=3
run steal1_sl4 RC_STEAL=1 RC_SLICES=4
run steal1_sl6 RC_STEAL=1 RC_SLICES=6
run steal0_sl4 RC_STEAL=0 RC_SLICES=4
run steal1_mb8 RC_STEAL=1 RC_MIN_BLOCKS=8
run steal1_mb10 RC_STEAL=1 RC_MIN_BLOCKS=10
run steal1_mb12 RC_STEAL=1 RC_MIN_BLOCKS=12
run steal1_sl4_mb10 RC_STEAL=1 RC_SLICES=4 RC_MIN_BLOCKS=10
