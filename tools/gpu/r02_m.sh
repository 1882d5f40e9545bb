#!/usr/bin/env bash
# Round 2, call M: cancel word in device memory (tests + e2e_cancel), one rank's share of an 8- and 4-way split
# for several slice counts, launch list and ncu capture of the v26 kernel at the bench configuration.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
PYTHONUNBUFFERED=1 timeout 400 python -m pytest tests -m gpu -q --timeout 200 -p no:cacheprovider -k "cancel or frame or bench_config" > gpurun_out/r02m_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Timeout|Error" gpurun_out/r02m_pytest.log | tail -12 | cut -c1-300
timeout 200 python bench.py --steps 10 --warmup 3 > gpurun_out/r02m_bench_full.json 2> gpurun_out/r02m_bench_full.err
echo "bench rc=$?"; python - <<'PY'
import json
d = json.load(open("gpurun_out/r02m_bench_full.json"))
print(f"{d['value']:.4e} samples/s {d['ms_per_step']:.3f} ms e2e {d['e2e']['value']:.4e} e2e_cancel {d['e2e_cancel']['value']:.4e} frac {d['roofline']['frac']:.4f}")
PY
timeout 200 python tools/time_share.py 8 2>&1 | tail -8
timeout 200 python tools/time_share.py 4 2>&1 | tail -8
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02m_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02m_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k spec_megakernel -c 1 -f -o gpurun_out/r02m_mega_v26_1024 \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02m_ncu_mega.log 2>&1
echo "ncu mega rc=$?"; ls -la gpurun_out/*.ncu-rep | tail -3
