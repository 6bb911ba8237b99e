#!/bin/bash
# 2-GPU check of view sharding: sharded == unsharded forward (NCCL gather, then the fused TMA-store gather), bench lines
# of both, and the exchange step's trace rows.   gpurun --gpus 2 -- bash scripts/check_fused_gather.sh
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
for f in 0 1; do
  echo "== IGGT_FUSED_GATHER=$f"
  IGGT_FUSED_GATHER=$f timeout 200 $TR scripts/check_sharded.py 2>&1 | tail -4
  IGGT_FUSED_GATHER=$f timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02d_bench_2gpu_fused$f.json 2> $O/r02d_bench_2gpu_fused$f.err
  tail -1 $O/r02d_bench_2gpu_fused$f.json | cut -c1-200; tail -2 $O/r02d_bench_2gpu_fused$f.err
done
python - <<PY
import json
for f in (0, 1):
    try:
        j = json.loads(open("$O/r02d_bench_2gpu_fused%d.json" % f).read().strip().splitlines()[-1])
        print("fused", f, round(j["ms_per_step"], 2), "ms", {k: round(v["ms_per_step"], 2) for k, v in j["kernel_shares"].items() if v["ms_per_step"] > 0.2})
    except Exception as e:
        print("fused", f, "ERR", e)
PY
