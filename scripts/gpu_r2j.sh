#!/bin/bash
set -u
O=gpurun_out; mkdir -p $O
IGGT_CAMERA_DEBUG=1 timeout 120 python scripts/camera_debug.py > $O/r02j_camera_debug1.log 2>&1; sed -n 1,12p $O/r02j_camera_debug1.log; tail -1 $O/r02j_camera_debug1.log
IGGT_CAMERA_DEBUG=2 timeout 120 python scripts/camera_debug.py > $O/r02j_camera_debug2.log 2>&1; sed -n 1,12p $O/r02j_camera_debug2.log; tail -1 $O/r02j_camera_debug2.log
