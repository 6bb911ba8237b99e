#!/bin/bash
# Round 2, GPU call S: the BASELINE configurations that fit one GPU (C1, C2 with the part path, C4, and C3 / C5 unsharded),
# and the C4 bench line (8 views 1036^2, IGGT incl. the part path).
set -u
TAG=r02s
O=gpurun_out; mkdir -p $O
timeout 400 python scripts/run_configs.py C1 C2 C2part C2bf16 C4 C3local C5local > $O/${TAG}_configs.log 2>&1; echo "configs rc=$?"; cut -c1-200 $O/${TAG}_configs.log | tail -8
mv $O/configs.json $O/${TAG}_configs_single_gpu.json
timeout 400 python bench.py --size 1036 --part --steps 5 --warmup 3 --no-cpu-baseline > $O/${TAG}_bench_c4.json 2> $O/${TAG}_bench_c4.err; echo "bench rc=$?"; cut -c1-200 $O/${TAG}_bench_c4.json; tail -2 $O/${TAG}_bench_c4.err
python - <<PY
import json
j = json.loads(open("$O/${TAG}_bench_c4.json").read().strip().splitlines()[-1])
print(round(j["ms_per_step"], 2), "ms; e2e", round(j["e2e"]["ms_per_step"], 2), "launches", j["gpu_launches"], {k: round(v["ms_per_step"], 2) for k, v in j["kernel_shares"].items() if v["ms_per_step"] > 1.0})
print(j.get("gpu_eager_baseline"))
PY
