#!/bin/bash
# Round 2, GPU call I: camera head, k-block order rotated per CTA (HBM channel camping), timeline + tests + bench.
set -u
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_camera_gpu.py -m gpu -q -x > $O/r02i_pytest_camera.log 2>&1; echo "camera pytest rc=$?"; tail -8 $O/r02i_pytest_camera.log | cut -c1-300
IGGT_CAMERA_DEBUG=1 timeout 120 python scripts/camera_debug.py > $O/r02i_camera_debug.log 2>&1; tail -30 $O/r02i_camera_debug.log
timeout 300 python bench.py --no-cpu-baseline > $O/r02i_bench.json 2> $O/r02i_bench.err; echo "bench rc=$?"; cut -c1-160 $O/r02i_bench.json; tail -3 $O/r02i_bench.err
python - <<PY
import json
j = json.loads(open("$O/r02i_bench.json").read().strip().splitlines()[-1])
print(round(j["ms_per_step"], 2), "ms; e2e", round(j["e2e"]["ms_per_step"], 2), {k: round(v["ms_per_step"], 2) for k, v in j["kernel_shares"].items() if v["ms_per_step"] > 0.25})
PY
