#!/bin/bash
# Round 2, GPU call N (8 GPUs): the fused K|V gather with several scenes at world 8 (check + C5), C2 again.
set -u
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519"
timeout 240 $TR scripts/check_sharded.py > $O/r02n_sharded.log 2>&1; grep -E "SHARDED|rror" $O/r02n_sharded.log | tail -3
timeout 300 $TR bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline --scenes 4 --views 16 --dtype bf16 > $O/r02n_c5_n8_fused.json 2> $O/r02n_c5_n8_fused.err; tail -1 $O/r02n_c5_n8_fused.json | cut -c1-160; grep -iE "error|capture failed" $O/r02n_c5_n8_fused.err | tail -2
timeout 300 $TR bench.py --gpus 8 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02n_c2_n8_fused.json 2> $O/r02n_c2_n8_fused.err; tail -1 $O/r02n_c2_n8_fused.json | cut -c1-160
python - <<PY
import json
for f in ("r02n_c5_n8_fused", "r02n_c2_n8_fused"):
    try:
        j = json.loads(open("$O/%s.json" % f).read().strip().splitlines()[-1])
        print(f, round(j["value"], 1), "views/s", round(j["ms_per_step"], 2), "ms; e2e", round(j["e2e"]["value"], 1), {k: round(v["ms_per_step"], 2) for k, v in j["kernel_shares"].items() if v["ms_per_step"] > 0.3})
    except Exception as e:
        print(f, "ERR", e)
PY
