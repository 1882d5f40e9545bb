#!/usr/bin/env bash
# Round 2, call H (8 GPUs): the N = 8 / 4 bench lines through rc_render_frame, one slice-count alternative, clown (sample split).
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
run() {  # label, nproc, env..., -- bench args
    local label=$1 n=$2; shift 2
    local envs=()
    while [ "$1" != "--" ]; do envs+=("$1"); shift; done
    shift
    env "${envs[@]}" timeout 240 $TR --nproc-per-node $n --master-port 29620 bench.py --gpus $n "$@" > gpurun_out/r02h_$label.json 2> gpurun_out/r02h_$label.err
    python - "$label" <<'PY'
import json, sys
try:
    line = [l for l in open(f"gpurun_out/r02h_{sys.argv[1]}.json") if l.startswith("{")][-1]
    d = json.loads(line)
    open(f"gpurun_out/r02h_{sys.argv[1]}.json", "w").write(line)
    print(f"{sys.argv[1]:>14}: {d['value']:.4e} samples/s  {d['ms_per_step']:.3f} ms  kernel_rank0 {d.get('kernel_ms_rank0')} ms  e2e {d['e2e']['value']:.4e}  e2e_cancel {(d.get('e2e_cancel') or {}).get('value')}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
}
run n8 8 RC_X=0 -- --steps 20 --warmup 5 --no-cpu-baseline
run n2 2 RC_X=0 -- --steps 20 --warmup 3 --no-cpu-baseline
run n4 4 RC_X=0 -- --steps 20 --warmup 3 --no-cpu-baseline
