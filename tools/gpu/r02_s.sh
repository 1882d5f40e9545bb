#!/usr/bin/env bash
# Round 2, call S: ncu capture of the Random scene's kernel at ITS bench configuration (256 spp) and of clown's (4096 spp, 4K)
set -u
cd "$(dirname "$0")/../.."
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on -k spec_megakernel -c 1 -f -o gpurun_out/r02s_random_256 \
    python bench.py --workload random_1080p_256spp --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02s_ncu_random.log 2>&1
echo "ncu random rc=$?"
ncu --set full --clock-control none --import-source on -k spec_megakernel -c 1 -f -o gpurun_out/r02s_three_balls \
    python bench.py --workload three_balls_600_200spp --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/r02s_ncu_tb.log 2>&1
echo "ncu three_balls rc=$?"
ls -la gpurun_out/r02s_*.ncu-rep
