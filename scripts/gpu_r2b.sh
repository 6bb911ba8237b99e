#!/bin/bash
# Round 2, GPU call B: attention variants (packed exp2 emulation, split-KV), GELU form A/B, small-M microbench.
set -u
TAG=r02b
O=gpurun_out
mkdir -p $O
timeout 600 python -m pytest tests/test_kernels_gpu.py -m gpu -q -k "attention or gemm_store16" > $O/${TAG}_pytest_attn.log 2>&1; echo "pytest rc=$?"; tail -3 $O/${TAG}_pytest_attn.log
timeout 900 python scripts/attn_sweep.py 0,1,2,3,4 > $O/${TAG}_attn_sweep.log 2>&1; echo "sweep rc=$?"; cat $O/${TAG}_attn_sweep.log | cut -c1-1500
cp $O/attn_sweep.json $O/${TAG}_attn_sweep.json 2>/dev/null
for G in 1 2; do
  IGGT_GELU=$G timeout 200 python scripts/microbench.py > $O/${TAG}_mb_gelu$G.json 2> $O/${TAG}_mb_gelu$G.err
  python -c "import json; d=json.load(open('$O/${TAG}_mb_gelu$G.json')); print('gelu$G fc1 us', round(d['fc1']['ms']*1e3,1), 'cublas', round(d['fc1']['cublas_ms']*1e3,1))"
done
timeout 200 python scripts/microbench.py --views 1 > $O/${TAG}_mb_1view.json 2> $O/${TAG}_mb_1view.err
timeout 200 python scripts/microbench.py --views 1 --noflush > $O/${TAG}_mb_1view_noflush.json 2> $O/${TAG}_mb_1view_noflush.err
python - <<PY
import json
for f in ("${TAG}_mb_1view", "${TAG}_mb_1view_noflush"):
    try:
        d = json.load(open("$O/%s.json" % f)); print(f, {k: (round(v["ms"] * 1000, 1), round(v.get("cublas_ms", v.get("sdpa_ms", 0)) * 1000, 1)) for k, v in d.items()})
    except Exception as e:
        print(f, "ERR", e)
PY
for E in 2 3; do
  IGGT_ATTN_EMU=$E timeout 300 python bench.py --no-cpu-baseline > $O/${TAG}_bench_emu$E.json 2> $O/${TAG}_bench_emu$E.err; echo "bench emu$E rc=$?"; cut -c1-160 $O/${TAG}_bench_emu$E.json
done
IGGT_GELU=2 timeout 300 python bench.py --no-cpu-baseline > $O/${TAG}_bench_gelu2.json 2> $O/${TAG}_bench_gelu2.err; echo "bench gelu2 rc=$?"; cut -c1-160 $O/${TAG}_bench_gelu2.json
ls -la $O | grep ${TAG}
