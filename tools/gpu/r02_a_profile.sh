#!/usr/bin/env bash
# Round 2, first GPU call: the round-1 kernels as they are — bench line, launch list, and ncu captures at the
# BENCH configuration (1024 spp), of the wavefront kernels and of the Random scene.
set -u
mkdir -p gpurun_out
cd "$(dirname "$0")/../.."
python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench.json 2> gpurun_out/r02a_bench.err
echo "bench rc=$?"; tail -c 600 gpurun_out/r02a_bench.json
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02a_launches.csv \
    python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_launches.log 2>&1
echo "launch list rc=$?"
ncu --set full --clock-control none --import-source on -k spec_megakernel -c 1 -f -o gpurun_out/r02a_mega_v22_1024 \
    python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02a_ncu_mega.log 2>&1
echo "ncu mega rc=$?"
ncu --set full --clock-control none --import-source on -k regex:wf_ --launch-skip 60 -c 24 -f -o gpurun_out/r02a_wavefront \
    python bench.py --variant wavefront --spp 16 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02a_ncu_wf.log 2>&1
echo "ncu wavefront rc=$?"
ncu --set full --clock-control none --import-source on -k spec_megakernel -c 1 -f -o gpurun_out/r02a_random_v3 \
    python bench.py --workload random_1080p_256spp --spp 32 --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02a_ncu_random.log 2>&1
echo "ncu random rc=$?"
python bench.py --workload random_1080p_256spp --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/r02a_bench_random.json 2>&1
ls -la gpurun_out
