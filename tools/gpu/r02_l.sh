#!/usr/bin/env bash
# Round 2, call L: v26 (one lean iteration body in every loop) — partition debug, GPU suite, bench.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 200 python tools/debug/partition_diff.py 2>&1 | tail -6
PYTHONUNBUFFERED=1 timeout 700 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider > gpurun_out/r02l_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Timeout|Error" gpurun_out/r02l_pytest.log | tail -12 | cut -c1-300
run() {  # label, env...
    local label=$1; shift
    env "$@" timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02l_$label.json 2> gpurun_out/r02l_$label.err
    python - "$label" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r02l_{sys.argv[1]}.json"))
    print(f"{sys.argv[1]:>16}: {d['value']:.4e} samples/s  {d['ms_per_step']:.3f} ms  e2e {d['e2e']['value']:.4e}  e2e_cancel {d['e2e_cancel']['value']:.4e} launches {d['gpu_launches']}  kernel {d['config']['kernel']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run v26 RC_X=0
run v26_steal0 RC_STEAL=0
