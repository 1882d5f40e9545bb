#!/usr/bin/env bash
# Round 2: Philox2x32-7 in the specialised kernel — test, and the bench line (reported separately from the 10-round headline)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
PYTHONUNBUFFERED=1 timeout 400 python -m pytest tests/test_gpu_bench_config.py -m gpu -q --timeout 250 -p no:cacheprovider -k "seven or benched or cancel" 2>&1 | tail -4
timeout 300 python bench.py --rng-rounds 7 --steps 10 --warmup 3 > gpurun_out/r02_bench_rounds7.json 2> gpurun_out/r02_bench_rounds7.err
python - <<'PY'
import json
d = json.load(open("gpurun_out/r02_bench_rounds7.json"))
print(f"rounds 7: {d['value']:.4e} samples/s {d['ms_per_step']:.3f} ms e2e {d['e2e']['value']:.4e} frac {d['roofline']['frac']:.4f} rng {d['config']['rng']} kernel {d['config']['kernel']}")
PY
timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02_bench_rounds10.json 2>/dev/null
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_rounds10.json')); print('rounds 10:', d['value'], d['ms_per_step'], d['roofline'] if 'roofline' in d else '')"
