#!/usr/bin/env bash
# Round 2, call O (2 GPUs): frame protocol with the one-frame-ahead release, two-device tests, N = 2 line.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $TR --nproc-per-node 2 --master-port 29611 tools/ipc_tiles_check.py > gpurun_out/r02o_ipc2.log 2>&1
echo "ipc check (2 ranks) rc=$?"; tail -4 gpurun_out/r02o_ipc2.log | cut -c1-400
PYTHONUNBUFFERED=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -v --timeout 200 -p no:cacheprovider -k "two_devices or processes_store" > gpurun_out/r02o_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "PASSED|FAILED|SKIPPED|passed|failed|^E  " gpurun_out/r02o_pytest.log | tail -8 | cut -c1-300
timeout 300 $TR --nproc-per-node 2 --master-port 29620 bench.py --gpus 2 --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r02o_n2.json 2> gpurun_out/r02o_n2.err
echo "bench n2 rc=$?"; python - <<'PY'
import json
line = [l for l in open("gpurun_out/r02o_n2.json") if l.startswith("{")][-1]
d = json.loads(line)
open("gpurun_out/r02o_n2.json", "w").write(line)
print(f"{d['value']:.4e} samples/s {d['ms_per_step']:.3f} ms kernel_rank0 {d.get('kernel_ms_rank0')} e2e {d['e2e']['value']:.4e} e2e_cancel {d['e2e_cancel']['value']:.4e}")
PY
