#!/bin/bash
# Round-end verification on the GPU box: every GPU test, smoke(), the bench line (with CPU baseline), the reference
# arm, the microbenchmarks and the ncu captures that profiles/ is refreshed from.   bash scripts/final_check.sh <tag>
set -u
TAG=${1:-r01final}
mkdir -p gpurun_out
timeout 500 python -m pytest tests -m gpu -q > gpurun_out/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/${TAG}_pytest.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" 2>&1 | tail -1
timeout 240 python bench.py > gpurun_out/${TAG}_bench.json 2> gpurun_out/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-330 gpurun_out/${TAG}_bench.json
timeout 240 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${TAG}_ref.json 2> gpurun_out/${TAG}_ref.err; echo "ref rc=$?"; cut -c1-300 gpurun_out/${TAG}_ref.json
timeout 100 python scripts/microbench.py > gpurun_out/${TAG}_mb.json 2>/dev/null
IGGT_PAIR=15 timeout 100 python scripts/microbench.py > gpurun_out/${TAG}_mb_pair15.json 2>/dev/null
python - <<PY
import json
for f in ("${TAG}_mb", "${TAG}_mb_pair15"):
    try:
        d = json.load(open("gpurun_out/%s.json" % f)); print(f, {k: round(v["ms"] * 1000, 1) for k, v in d.items()})
    except Exception as e:
        print(f, "ERR", e)
PY
timeout 400 bash scripts/profile_ncu.sh ${TAG} > gpurun_out/${TAG}_ncu.log 2>&1; echo "ncu rc=$?"
ls gpurun_out | grep ${TAG} | head -20
