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
# --launch-skip counts MATCHING launches.  Attention: 72 per forward (24 DINO, then frame / global alternating), so
# skipping 96 lands on the first aggregator frame + global pair of the second forward; GEMMs: any three mid-forward.
for KS in "attention3_kernel 96 2" "gemm_tcgen05_kernel 400 3"; do
  set -- $KS
  ncu --set full --clock-control none --import-source on -k regex:$1 -s $2 -c $3 -f -o gpurun_out/prof_${TAG}_$1 \
      python bench.py --quick --warmup 1 --steps 1 > gpurun_out/ncu_full_${TAG}_$1.log 2>&1
done
ls -la gpurun_out/
