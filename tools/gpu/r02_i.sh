#!/usr/bin/env bash
# Round 2, call I: v24 (161-instruction path loop) — whole GPU suite, then the bench line and the N = 1 slice alternatives.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
PYTHONUNBUFFERED=1 timeout 700 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider > gpurun_out/r02i_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Timeout|Error" gpurun_out/r02i_pytest.log | tail -12 | cut -c1-300
run() {  # label, env...
    local label=$1; shift
    env "$@" timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02i_$label.json 2> gpurun_out/r02i_$label.err
    python - "$label" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r02i_{sys.argv[1]}.json"))
    print(f"{sys.argv[1]:>16}: {d['value']:.4e} samples/s  {d['ms_per_step']:.3f} ms  e2e {d['e2e']['value']:.4e}  e2e_cancel {d['e2e_cancel']['value']:.4e} launches {d['gpu_launches']}  kernel {d['config']['kernel']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run v24 RC_X=0
run v24_sl3 RC_SLICES=3
run v24_mb4 RC_MIN_BLOCKS=4
run v24_mb8 RC_MIN_BLOCKS=8
