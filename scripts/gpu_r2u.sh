#!/bin/bash
# Round 2, GPU call U: ncu --set full of the camera head after the chunk-major change (8 and 16 rows) and of LayerNorm.
set -u
TAG=r02u
O=gpurun_out; mkdir -p $O
timeout 240 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/prof_${TAG} \
    python scripts/ncu_targets.py --only camera_head_one_launch,camera_head_one_launch_16_rows,layernorm > $O/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"; tail -3 $O/${TAG}_ncu.log
ncu -i /tmp/prof_${TAG}.ncu-rep --page raw --csv > $O/${TAG}_ncu_raw.csv 2>/dev/null
ls -la $O | grep ${TAG}
