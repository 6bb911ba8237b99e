#!/bin/bash
# Round 2, GPU call Q: verification after the camera producer-order fix and the NaN-keeping ReLU; camera head fused vs layer
# path beyond 16 rows; launch list of one forward.
set -u
TAG=r02q
O=gpurun_out; mkdir -p $O
timeout 300 python -m pytest tests/test_camera_gpu.py tests/test_kernels_gpu.py -m gpu -q -k "camera or relu" > $O/${TAG}_pytest_camera.log 2>&1; echo "camera pytest rc=$?"; tail -3 $O/${TAG}_pytest_camera.log | cut -c1-300
IGGT_CAMERA_DEBUG=1 timeout 120 python scripts/camera_debug.py > $O/${TAG}_camera_debug.log 2>&1; head -2 $O/${TAG}_camera_debug.log
python - <<PY 2>&1 | tee $O/${TAG}_camera_paths.log
import torch, sys, os
sys.path.insert(0, os.getcwd())
from iggt_official_b200.heads import camera_head as CH
from iggt_official_b200.models.vggt import VGGT
m = VGGT().eval().cuda(); head = m.camera_head
for (B, S) in ((1, 8), (1, 16), (2, 8), (2, 16), (4, 8), (4, 16), (8, 4)):
    tok = torch.randn(B, S, 7, 2048, device="cuda")
    toks = [None] * 23 + [tok]
    row = {}
    for fused in (True, False):
        CH.FUSED = fused
        for _ in range(2): head(toks, compute_dtype=torch.float16)
        torch.cuda.synchronize(); e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3): head(toks, compute_dtype=torch.float16)
        e1.record(); torch.cuda.synchronize()
        row["fused" if fused else "layers"] = round(e0.elapsed_time(e1) * 1e3 / 3)
    CH.FUSED = True
    print(f"camera head B={B} S={S} (us):", row)
PY
timeout 1200 python -m pytest tests -m gpu -q > $O/${TAG}_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 $O/${TAG}_pytest.log | cut -c1-300
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('SMOKE_OK')" > $O/${TAG}_smoke.log 2>&1; tail -2 $O/${TAG}_smoke.log
timeout 400 python bench.py > $O/${TAG}_bench.json 2> $O/${TAG}_bench.err; echo "bench rc=$?"; cut -c1-200 $O/${TAG}_bench.json; tail -2 $O/${TAG}_bench.err
python - <<PY
import json
j = json.loads(open("$O/${TAG}_bench.json").read().strip().splitlines()[-1])
print(round(j["ms_per_step"], 2), "ms; e2e", round(j["e2e"]["ms_per_step"], 2), "launches", j["gpu_launches"], {k: round(v["ms_per_step"], 2) for k, v in j["kernel_shares"].items() if v["ms_per_step"] > 0.3})
PY
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $O/launches_${TAG}.csv \
    python bench.py --quick --warmup 1 --steps 1 > $O/${TAG}_ncu_list.log 2>&1; echo "launch list rc=$?"
ls -la $O | grep ${TAG}
