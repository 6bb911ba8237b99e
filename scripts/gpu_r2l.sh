#!/bin/bash
# Round 2, GPU call L: everything green? full GPU suite, smoke, final bench lines, tensor-core window attentions A/B,
# ncu of the kernels added this round.
set -u
TAG=r02l
O=gpurun_out; mkdir -p $O
timeout 200 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "ocab or window_attention" > $O/${TAG}_pytest_winattn.log 2>&1; echo "winattn pytest rc=$?"; tail -4 $O/${TAG}_pytest_winattn.log | cut -c1-300
timeout 1200 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/${TAG}_pytest.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/${TAG}_smoke.log 2>&1; tail -6 $O/${TAG}_smoke.log
timeout 400 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-200 $O/${TAG}_bench.json; tail -2 $O/${TAG}_bench.err
timeout 300 python bench.py --part --size 532 --no-cpu-baseline > $O/${TAG}_bench_part532.json 2> $O/${TAG}_bench_part532.err; echo "bench(part, TC windows) rc=$?"; cut -c1-160 $O/${TAG}_bench_part532.json
IGGT_WINATTN_TC=0 timeout 300 python bench.py --part --size 532 --no-cpu-baseline > $O/${TAG}_bench_part532_scalar.json 2> $O/${TAG}_bench_part532_scalar.err; echo "bench(part, scalar windows) rc=$?"; cut -c1-160 $O/${TAG}_bench_part532_scalar.json
python - <<PY
import json
for f in ("${TAG}_bench", "${TAG}_bench_part532", "${TAG}_bench_part532_scalar"):
    try:
        j = json.loads(open("$O/%s.json" % f).read().strip().splitlines()[-1])
        print(f, round(j["ms_per_step"], 2), "ms; e2e", round(j["e2e"]["ms_per_step"], 2), {k: round(v["ms_per_step"], 2) for k, v in j["kernel_shares"].items() if v["ms_per_step"] > 0.3})
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 600 ncu --set full --clock-control none --profile-from-start off -f -o /tmp/prof_${TAG} \
    python scripts/ncu_targets.py --only camera_head_one_launch,attention_global_1of8_split3,dpt_tail_fused,upsample_bilinear_pe_296_518,upsample_bilinear_74_148,ocab_attention,window_attention > $O/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
ncu -i /tmp/prof_${TAG}.ncu-rep --page raw --csv > $O/${TAG}_ncu_raw.csv 2>/dev/null
ls -la $O | grep ${TAG}
