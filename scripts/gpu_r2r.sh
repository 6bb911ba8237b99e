#!/bin/bash
# Round 2, GPU call R (2 GPUs): camera head refined per scene on separate ranks (poses all-gathered) - sharded == unsharded,
# and its effect on a two-scene step; the ReLU NaN test.   gpurun --gpus 2 -- bash scripts/gpu_r2r.sh
set -u
TAG=r02r
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k relu > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 $O/${TAG}_pytest.log | cut -c1-200
timeout 200 $TR scripts/check_sharded.py > $O/${TAG}_sharded.log 2>&1; grep -E "SHARDED|rror|pose_enc" $O/${TAG}_sharded.log | tail -6
for f in 1 0; do
  IGGT_CAMERA_BY_SCENE=$f timeout 300 $TR bench.py --gpus 2 --scenes 2 --views 16 --dtype bf16 --steps 5 --warmup 3 --no-cpu-baseline \
      > $O/${TAG}_bench_2x16_2gpu_byscene$f.json 2> $O/${TAG}_bench_2x16_2gpu_byscene$f.err
  tail -1 $O/${TAG}_bench_2x16_2gpu_byscene$f.json | cut -c1-160; tail -2 $O/${TAG}_bench_2x16_2gpu_byscene$f.err
done
python - <<PY
import json
for f in (1, 0):
    try:
        j = json.loads(open("$O/${TAG}_bench_2x16_2gpu_byscene%d.json" % f).read().strip().splitlines()[-1])
        print("by scene", f, round(j["ms_per_step"], 2), "ms; e2e", round(j["e2e"]["ms_per_step"], 2), {k: round(v["ms_per_step"], 2) for k, v in j["kernel_shares"].items() if "camera" in k or "nccl" in k or "symm" in k})
    except Exception as e:
        print("by scene", f, "ERR", e)
PY
