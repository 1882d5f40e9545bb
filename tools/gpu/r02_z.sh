#!/usr/bin/env bash
# Round 2, call Z: two sample slices for big single-GPU frames — GPU suite, bench (full line), launch list, ncu capture
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
PYTHONUNBUFFERED=1 timeout 900 python -m pytest tests -m gpu -q --timeout 250 -p no:cacheprovider > gpurun_out/r02z_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Timeout|Error" gpurun_out/r02z_pytest.log | tail -8 | cut -c1-300
timeout 300 python bench.py --steps 10 --warmup 3 > gpurun_out/r02z_bench_full.json 2> gpurun_out/r02z_bench_full.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02z_bench_full.json"))
print(f"{d['value']:.4e} samples/s {d['ms_per_step']:.3f} ms e2e {d['e2e']['value']:.4e} e2e_cancel {d['e2e_cancel']['value']:.4e} frac {d['roofline']['frac']:.4f} launches {d['gpu_launches']}")
PY
RC_SLICES=1 timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02z_bench_unsliced.json 2> gpurun_out/r02z_bench_unsliced.err
python -c "
import json; d=json.load(open('gpurun_out/r02z_bench_unsliced.json')); print('unsliced:', d['value'], d['ms_per_step'])"
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02z_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02z_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k spec_megakernel -c 1 -f -o gpurun_out/r02z_mega_v27_1024 \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02z_ncu_mega.log 2>&1
echo "ncu mega rc=$?"
