#!/bin/bash
# Run on the GPU box (under gpurun).  1) per-launch device times of one forward; 2) full capture of the top kernels.
# Usage: bash scripts/profile_ncu.sh <tag>     -> gpurun_out/launches_<tag>.csv, gpurun_out/prof_<tag>_*.ncu-rep
set -u
TAG=${1:-r01}
mkdir -p gpurun_out
L=$(python bench.py --quick --warmup 1 --steps 1 | python -c "import sys,json; print(json.loads(sys.stdin.readline())['launches_per_step'])")
echo "launches per step: $L"
# torch launches a few kernels of its own per step (copies); capture generously and post-filter by name
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches_${TAG}.csv \
    python bench.py --quick --warmup 1 --steps 1 > gpurun_out/ncu_list_${TAG}.log 2>&1
for K in attention3_kernel gemm_tcgen05_kernel; do
  ncu --set full --clock-control none --import-source on -k regex:${K} -s 300 -c 4 -f -o gpurun_out/prof_${TAG}_${K} \
      python bench.py --quick --warmup 1 --steps 1 > gpurun_out/ncu_full_${TAG}_${K}.log 2>&1
done
ls -la gpurun_out/
