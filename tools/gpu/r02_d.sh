#!/usr/bin/env bash
# Round 2, call D: v23 with the edge-tile fix — parity suite (bounded), CTAs-per-SM sweep.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 400 python -m pytest tests -m gpu -q -x --timeout 120 > gpurun_out/r02d_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r02d_pytest.log | cut -c1-300
run() {  # label, env...
    local label=$1; shift
    env "$@" timeout 120 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02d_$label.json 2> gpurun_out/r02d_$label.err
    python - "$label" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r02d_{sys.argv[1]}.json"))
    print(f"{sys.argv[1]:>16}: {d['value']:.4e} samples/s  {d['ms_per_step']:.3f} ms  e2e {d['e2e']['value']:.4e}  launches {d['gpu_launches']}  kernel {d['config']['kernel']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run v23 RC_STEAL=1
run v23_mb4 RC_MIN_BLOCKS=4
run v23_mb5 RC_MIN_BLOCKS=5
run v23_mb6 RC_MIN_BLOCKS=6
run v23_mb7 RC_MIN_BLOCKS=7
run v23_mb9 RC_MIN_BLOCKS=9
