#!/bin/bash
# Round 2, GPU call M (2 GPUs): separate QKV-gather instantiation re-verified; C1 test; C2 bench.
set -u
O=gpurun_out; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
timeout 200 $TR scripts/check_sharded.py > $O/r02m_sharded.log 2>&1; grep -E "SHARDED|rror" $O/r02m_sharded.log | tail -3
timeout 300 python -m pytest tests/test_c1_plumbing_gpu.py tests/test_kernels_gpu.py -m gpu -q -k "c1 or demo1 or qkv" > $O/r02m_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 $O/r02m_pytest.log | cut -c1-200
timeout 300 python bench.py --no-cpu-baseline > $O/r02m_bench.json 2> $O/r02m_bench.err; cut -c1-160 $O/r02m_bench.json
timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-cpu-baseline > $O/r02m_bench_2gpu.json 2> $O/r02m_bench_2gpu.err; tail -1 $O/r02m_bench_2gpu.json | cut -c1-160
python - <<PY
import json
for f in ("r02m_bench", "r02m_bench_2gpu"):
    try:
        j = json.loads(open("$O/%s.json" % f).read().strip().splitlines()[-1])
        print(f, round(j["ms_per_step"], 2), "ms; e2e", round(j["e2e"]["ms_per_step"], 2), {k: round(v["ms_per_step"], 2) for k, v in j["kernel_shares"].items() if v["ms_per_step"] > 0.3})
    except Exception as e:
        print(f, "ERR", e)
PY
