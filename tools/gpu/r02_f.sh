#!/usr/bin/env bash
# Round 2, call F: the whole GPU suite (all failures listed), then full bench lines for every BASELINE workload at N = 1.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
PYTHONUNBUFFERED=1 timeout 900 python -m pytest tests -m gpu -v --timeout 200 -p no:cacheprovider > gpurun_out/r02f_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Timeout" gpurun_out/r02f_pytest.log | tail -12 | cut -c1-300
for wl in three_balls_600_200spp emissive_600_200spp noise_and_textures_600_200spp random_1080p_256spp clown_4k_4096spp; do
    timeout 400 python bench.py --workload $wl --steps 5 --warmup 3 > gpurun_out/r02f_bench_${wl}_n1.json 2> gpurun_out/r02f_bench_${wl}.err
    echo "$wl rc=$?"
    python - $wl <<'PY'
import json, sys
try:
    d = json.load(open(f"gpurun_out/r02f_bench_{sys.argv[1]}_n1.json"))
    print(f"  {d['value']:.4e} samples/s  {d['ms_per_step']:.3f} ms  e2e {d['e2e']['value']:.4e}  frac {d.get('roofline',{}).get('frac')}  A {d.get('roofline',{}).get('flops_per_sample')}  cpu {d['cpu_baseline']['value']:.3e}")
except Exception as e:
    print("  FAILED", e)
PY
done
