#!/bin/bash
# Round 2, GPU call K (2 GPUs): fused K|V gather with several scenes (direct peer stores for straddling rows); camera v4.
set -u
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
echo "== sharded vs unsharded, fused gather, B = 2"; timeout 200 $TR scripts/check_sharded.py > $O/r02k_sharded.log 2>&1; grep -E "SHARDED|rror" $O/r02k_sharded.log | tail -4
timeout 200 python -m pytest tests/test_camera_gpu.py -m gpu -q -x > $O/r02k_pytest_camera.log 2>&1; echo "camera pytest rc=$?"; tail -6 $O/r02k_pytest_camera.log | cut -c1-300
IGGT_CAMERA_DEBUG=1 timeout 120 python scripts/camera_debug.py > $O/r02k_camera_debug.log 2>&1; sed -n 1,14p $O/r02k_camera_debug.log; tail -1 $O/r02k_camera_debug.log
IGGT_CAMERA_DEBUG=2 timeout 120 python scripts/camera_debug.py > $O/r02k_camera_debug_nomath.log 2>&1; sed -n 1,1p $O/r02k_camera_debug_nomath.log; tail -1 $O/r02k_camera_debug_nomath.log
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline --scenes 2 --views 4 > $O/r02k_bench_2gpu_2x4.json 2> $O/r02k_bench_2gpu_2x4.err; tail -1 $O/r02k_bench_2gpu_2x4.json | cut -c1-200; grep -iE "error|capture failed" $O/r02k_bench_2gpu_2x4.err | tail -2
