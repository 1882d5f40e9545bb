#!/usr/bin/env bash
# Round 2, call X (8 GPUs): FULL lines (roofline + cpu_baseline + e2e) at N = 8 — the headline workload and clown (sample split)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
for wl in cornell_box_1080p_1024spp clown_4k_4096spp; do
    timeout 400 $TR --nproc-per-node 8 --master-port 29620 bench.py --gpus 8 --workload $wl --steps 10 --warmup 3 > gpurun_out/r02x_${wl}_n8.json 2> gpurun_out/r02x_${wl}_n8.err
    python - "$wl" <<'PY'
import json, sys
try:
    line = [l for l in open(f"gpurun_out/r02x_{sys.argv[1]}_n8.json") if l.startswith("{")][-1]
    d = json.loads(line)
    open(f"gpurun_out/r02x_{sys.argv[1]}_n8.json", "w").write(line)
    r = d.get("roofline") or {}
    print(f"{sys.argv[1]}: {d['value']:.4e} samples/s {d['ms_per_step']:.3f} ms e2e {d['e2e']['value']:.4e} e2e_cancel {(d.get('e2e_cancel') or {}).get('value')} frac {r.get('frac')} peak {r.get('peak')} cpu {(d.get('cpu_baseline') or {}).get('value')} clocks {d['clocks']}")
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
done
