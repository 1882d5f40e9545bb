#!/usr/bin/env bash
# Round 2, call Q: the whole GPU suite on the final build, smoke(), the default bench line.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
PYTHONUNBUFFERED=1 timeout 900 python -m pytest tests -m gpu -q --timeout 250 -p no:cacheprovider > gpurun_out/r02q_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Timeout|Error" gpurun_out/r02q_pytest.log | tail -12 | cut -c1-300
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3
timeout 300 python bench.py > gpurun_out/r02q_bench_default.json 2> gpurun_out/r02q_bench_default.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02q_bench_default.json"))
print(f"{d['value']:.4e} samples/s {d['ms_per_step']:.3f} ms steps {d['steps']} e2e {d['e2e']['value']:.4e} ({d['e2e']['steps']} steps) e2e_cancel {d['e2e_cancel']['value']:.4e} frac {d['roofline']['frac']:.4f} frac_executed {d['roofline']['frac_executed']:.4f} traffic {d['roofline']['traffic']} launches {d['gpu_launches']} cpu {d['cpu_baseline']['value']:.3e}")
PY
