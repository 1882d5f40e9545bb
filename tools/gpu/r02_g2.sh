#!/usr/bin/env bash
# Round 2, call G2 (2 GPUs): the frame API with two ranks, the two-device tests, the N = 2 bench line.
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $TR --nproc-per-node 2 --master-port 29611 tools/ipc_tiles_check.py > gpurun_out/r02g_ipc2.log 2>&1
echo "ipc check (2 ranks) rc=$?"; tail -4 gpurun_out/r02g_ipc2.log | cut -c1-400
PYTHONUNBUFFERED=1 timeout 300 python -m pytest tests/test_gpu_parity.py -m gpu -v --timeout 200 -p no:cacheprovider -k "two_devices or processes_store" > gpurun_out/r02g_pytest.log 2>&1
echo "pytest rc=$?"; grep -E "PASSED|FAILED|SKIPPED|passed|failed|^E  " gpurun_out/r02g_pytest.log | tail -8 | cut -c1-300
timeout 300 $TR --nproc-per-node 2 --master-port 29620 bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r02g_n2.json 2> gpurun_out/r02g_n2.err
echo "bench n2 rc=$?"; tail -c 1500 gpurun_out/r02g_n2.json; tail -5 gpurun_out/r02g_n2.err | cut -c1-300
