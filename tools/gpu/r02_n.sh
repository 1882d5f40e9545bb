#!/usr/bin/env bash
# Round 2, call N: slice shapes for one rank's share of an 8- / 4-way split; the library SAH re-split test; per-workload lines with v26.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
timeout 200 python tools/time_share.py 8 2>&1 | tail -12
timeout 200 python tools/time_share.py 4 2>&1 | tail -12
PYTHONUNBUFFERED=1 timeout 400 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider -k "resplits or sliced" > gpurun_out/r02n_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Timeout|Error|^E " gpurun_out/r02n_pytest.log | tail -12 | cut -c1-300
for wl in three_balls_600_200spp emissive_600_200spp noise_and_textures_600_200spp random_1080p_256spp clown_4k_4096spp; do
    timeout 400 python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/r02n_bench_${wl}_n1.json 2> gpurun_out/r02n_bench_${wl}.err
    echo "$wl rc=$?"
    python - $wl <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r02n_bench_{sys.argv[1]}_n1.json"))
    print(f"  {d['value']:.4e} samples/s  {d['ms_per_step']:.3f} ms  e2e {d['e2e']['value']:.4e} e2e_cancel {d['e2e_cancel']['value']:.4e} frac {d.get('roofline',{}).get('frac')}  cpu {d['cpu_baseline']['value']:.3e}")
except Exception as e:
    print("  FAILED", e)
PY
done
