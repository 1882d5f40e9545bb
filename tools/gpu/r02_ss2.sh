#!/usr/bin/env bash
# Round 2 (2 GPUs): the sample split through rc_render_frame — IPC check, clown N = 2 line, and the NCCL form for comparison
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 240 $TR --nproc-per-node 2 --master-port 29611 tools/ipc_tiles_check.py > gpurun_out/r02ss_ipc2.log 2>&1
echo "ipc check (2 ranks) rc=$?"; tail -3 gpurun_out/r02ss_ipc2.log | cut -c1-400
PYTHONUNBUFFERED=1 timeout 300 python -m pytest tests/test_gpu_bench_config.py -m gpu -q --timeout 200 -p no:cacheprovider -k "frame_api" 2>&1 | tail -3
for mode in frame nccl; do
    if [ $mode = nccl ]; then export RC_BENCH_NCCL_GATHER=1; else unset RC_BENCH_NCCL_GATHER; fi
    timeout 300 $TR --nproc-per-node 2 --master-port 29620 bench.py --gpus 2 --workload clown_4k_4096spp --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02ss_clown_n2_$mode.json 2> gpurun_out/r02ss_clown_n2_$mode.err
    python - $mode <<'PY'
import json, sys
try:
    line = [l for l in open(f"gpurun_out/r02ss_clown_n2_{sys.argv[1]}.json") if l.startswith("{")][-1]
    d = json.loads(line)
    print(f"clown n2 {sys.argv[1]}: {d['value']:.4e} samples/s {d['ms_per_step']:.3f} ms e2e {d['e2e']['value']:.4e} e2e_cancel {(d.get('e2e_cancel') or {}).get('value')} call {d['e2e']['call']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e); print(open(f"gpurun_out/r02ss_clown_n2_{sys.argv[1]}.err").read()[-800:])
PY
done
