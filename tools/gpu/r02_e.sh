#!/usr/bin/env bash
# Round 2, call E: stealing (all edge cases), cancel relay, frame API, baseline-size parity tests, bench line.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
PYTHONUNBUFFERED=1 timeout 600 python -m pytest tests -m gpu -v -x --timeout 150 -p no:cacheprovider > gpurun_out/r02e_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "passed|failed|FAILED|Timeout|Error" gpurun_out/r02e_pytest.log | tail -8 | cut -c1-300
timeout 300 python bench.py --steps 5 --warmup 3 > gpurun_out/r02e_bench.json 2> gpurun_out/r02e_bench.err
echo "bench rc=$?"; tail -c 2500 gpurun_out/r02e_bench.json; tail -3 gpurun_out/r02e_bench.err
