#!/bin/bash
# Round 2, GPU call H (8 GPUs): sharded == unsharded at world 8, C2 / C3 / C5 bench lines (fused K|V gather vs NCCL).
#   gpurun --gpus 8 -- bash scripts/gpu_r2h.sh
set -u
O=gpurun_out; mkdir -p $O
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
echo "== sharded vs unsharded forward, world $N (fused gather)"; timeout 300 $TR scripts/check_sharded.py 2>&1 | grep -E "SHARDED|Error|error" | tail -3
echo "== same with the NCCL gather"; IGGT_FUSED_GATHER=0 timeout 300 $TR scripts/check_sharded.py 2>&1 | grep -E "SHARDED|Error|error" | tail -3
run() {  # tag, env, bench args
  local tag=$1; shift; local env=$1; shift
  env $env timeout 400 $TR bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline "$@" > $O/r02h_${tag}.json 2> $O/r02h_${tag}.err
  echo "$tag rc=$? $(tail -1 $O/r02h_${tag}.json | cut -c1-150)"; grep -iE "error|Traceback|capture failed" $O/r02h_${tag}.err | tail -3
}
run c2_n${N}_fused IGGT_FUSED_GATHER=1
run c2_n${N}_nccl IGGT_FUSED_GATHER=0
run c3_n${N}_fused IGGT_FUSED_GATHER=1 --views 32
run c3_n${N}_nccl IGGT_FUSED_GATHER=0 --views 32
run c5_n${N}_fused IGGT_FUSED_GATHER=1 --scenes 4 --views 16 --dtype bf16
run c5_n${N}_nccl IGGT_FUSED_GATHER=0 --scenes 4 --views 16 --dtype bf16
python - <<PY
import json, glob
for f in sorted(glob.glob("$O/r02h_*.json")):
    try:
        j = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("/")[-1], round(j["value"], 1), "views/s", round(j["ms_per_step"], 2), "ms; e2e", round(j["e2e"]["value"], 1), "graph", j.get("cuda_graph"),
              {k: round(v["ms_per_step"], 2) for k, v in j["kernel_shares"].items() if v["ms_per_step"] > 0.3})
    except Exception as e:
        print(f, "ERR", e)
PY
