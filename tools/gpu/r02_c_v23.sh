#!/usr/bin/env bash
# Round 2, call C: v23 = in-warp sample stealing + the Philox block scheduled inside the closest hit.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/r02c_pytest.log 2>&1
echo "pytest rc=$?"; tail -8 gpurun_out/r02c_pytest.log
run() {  # label, env...
    local label=$1; shift
    env "$@" python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02c_$label.json 2> gpurun_out/r02c_$label.err
    python - "$label" <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r02c_{sys.argv[1]}.json"))
    print(f"{sys.argv[1]:>16}: {d['value']:.4e} samples/s  {d['ms_per_step']:.3f} ms  e2e {d['e2e']['value']:.4e}  launches {d['gpu_launches']}  kernel {d['config']['kernel']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run v23 RC_STEAL=1
run v23_steal0 RC_STEAL=0
run v23_mb10 RC_MIN_BLOCKS=10
run v23_mb8 RC_MIN_BLOCKS=8
run v23_mb7 RC_MIN_BLOCKS=7
ncu --set full --clock-control none --import-source on -k spec_megakernel -c 1 -f -o gpurun_out/r02c_mega_v23_1024 \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02c_ncu_mega.log 2>&1
echo "ncu mega rc=$?"
