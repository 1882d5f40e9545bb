#!/usr/bin/env bash
# Round 2, call J: v25 (147-instruction hot loop: scene-origin shift, register constants, all-pixel hot loop, v-step) —
# whole GPU suite, bench line (with cpu baseline), a few switches off one at a time, preview / upload latency.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
PYTHONUNBUFFERED=1 timeout 700 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider > gpurun_out/r02j_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Timeout|Error" gpurun_out/r02j_pytest.log | tail -12 | cut -c1-300
run() {  # label, env...
    local label=$1; shift
    env "$@" timeout 150 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02j_$label.json 2> gpurun_out/r02j_$label.err
    python - "$label" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r02j_{sys.argv[1]}.json"))
    print(f"{sys.argv[1]:>16}: {d['value']:.4e} samples/s  {d['ms_per_step']:.3f} ms  e2e {d['e2e']['value']:.4e}  e2e_cancel {d['e2e_cancel']['value']:.4e} launches {d['gpu_launches']}  kernel {d['config']['kernel']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run v25 RC_X=0
run v25_noshift RC_SPEC_NO_SHIFT=1
run v25_noregc RC_SPEC_NO_REG_CONSTS=1
run v25_packall RC_SPEC_PACK_ALL=1
run v25_sl3 RC_SLICES=3
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02j_bench_full.json 2> gpurun_out/r02j_bench_full.err
echo "full bench rc=$?"; tail -c 600 gpurun_out/r02j_bench_full.json
timeout 200 python tools/preview_latency.py gpurun_out/r02j_preview_latency.json > gpurun_out/r02j_preview_latency.log 2>&1
echo "preview latency rc=$?"; tail -5 gpurun_out/r02j_preview_latency.log | cut -c1-400
