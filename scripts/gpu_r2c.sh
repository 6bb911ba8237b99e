#!/bin/bash
# Round 2, GPU call C: one-launch camera head, rewritten upsample kernel, refitted split planner.
set -u
TAG=r02c
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_camera_gpu.py -m gpu -q -x > $O/${TAG}_pytest_camera.log 2>&1; echo "camera pytest rc=$?"; tail -12 $O/${TAG}_pytest_camera.log | cut -c1-300
timeout 600 python -m pytest tests/test_kernels_gpu.py tests/test_model_gpu.py tests/test_fullsize_gpu.py -m gpu -q > $O/${TAG}_pytest_kernels.log 2>&1; echo "kernels pytest rc=$?"; tail -5 $O/${TAG}_pytest_kernels.log | cut -c1-300
timeout 300 python bench.py --no-cpu-baseline > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-200 $O/${TAG}_bench.json; tail -3 $O/${TAG}_bench.err
IGGT_CAMERA_FUSED=0 timeout 300 python bench.py --no-cpu-baseline > $O/${TAG}_bench_camera_layers.json 2> $O/${TAG}_bench_camera_layers.err; echo "bench(layer camera) rc=$?"; cut -c1-200 $O/${TAG}_bench_camera_layers.json
python - <<PY
import json
for f in ("${TAG}_bench", "${TAG}_bench_camera_layers"):
    try:
        j = json.loads(open("$O/%s.json" % f).read().strip().splitlines()[-1])
        print(f, round(j["ms_per_step"], 2), "ms; graph", j.get("cuda_graph"), {k: round(v["ms_per_step"], 2) for k, v in j["kernel_shares"].items() if v["ms_per_step"] > 0.25})
    except Exception as e:
        print(f, "ERR", e)
PY
ls -la $O | grep ${TAG}
