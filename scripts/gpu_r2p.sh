#!/bin/bash
# Round 2, GPU call P: final verification - camera head (chunk-major), the whole GPU suite, smoke, the bench lines.
set -u
TAG=r02p
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_camera_gpu.py -m gpu -q -x > $O/${TAG}_pytest_camera.log 2>&1; echo "camera pytest rc=$?"; tail -4 $O/${TAG}_pytest_camera.log | cut -c1-300
IGGT_CAMERA_DEBUG=1 timeout 120 python scripts/camera_debug.py > $O/${TAG}_camera_debug.log 2>&1; head -1 $O/${TAG}_camera_debug.log; tail -1 $O/${TAG}_camera_debug.log
python - <<PY
import torch, sys, os
sys.path.insert(0, os.getcwd())
from iggt_official_b200 import ops
from iggt_official_b200.models.vggt import VGGT
m = VGGT().eval().cuda(); head = m.camera_head
for (B, S) in ((1, 8), (1, 16), (2, 8), (4, 16)):
    tok = torch.randn(B, S, 7, 2048, device="cuda")
    toks = [None] * 23 + [tok]
    for _ in range(2): head(toks, compute_dtype=torch.float16)
    torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); head(toks, compute_dtype=torch.float16); e1.record(); torch.cuda.synchronize()
    print(f"camera head B={B} S={S}: {e0.elapsed_time(e1) * 1e3:.0f} us")
PY
timeout 1200 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 $O/${TAG}_pytest.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
timeout 400 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-200 $O/${TAG}_bench.json; tail -2 $O/${TAG}_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > $O/${TAG}_ref.json 2> $O/${TAG}_ref.err; echo "ref rc=$?"; cut -c1-160 $O/${TAG}_ref.json
python - <<PY
import json
j = json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
print(round(j["ms_per_step"], 2), "ms; e2e", round(j["e2e"]["ms_per_step"], 2), "launches", j["gpu_launches"], {k: round(v["ms_per_step"], 2) for k, v in j["kernel_shares"].items() if v["ms_per_step"] > 0.3})
print({k: j[k] for k in ("gpu_eager_baseline", "cpu_baseline") if k in j})
PY
ls -la $O | grep ${TAG}
