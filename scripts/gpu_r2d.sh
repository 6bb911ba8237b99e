#!/bin/bash
# Round 2, GPU call D (2 GPUs): view sharding with the NCCL gather vs the fused TMA-store gather; camera-head timeline.
set -u
O=gpurun_out; mkdir -p $O
IGGT_CAMERA_DEBUG=1 timeout 120 python scripts/camera_debug.py > $O/r02d_camera_debug.log 2>&1; tail -32 $O/r02d_camera_debug.log
bash scripts/check_fused_gather.sh 2>&1 | tee $O/r02d_fused_gather.log | tail -30
