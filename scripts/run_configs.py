"""Run the BASELINE.json configurations that fit one GPU (timing with CUDA events, finiteness checks).
usage: python scripts/run_configs.py C2 C2part C4 C5 ...   (results -> gpurun_out/configs.json)"""
import json, os, sys, time
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from iggt_official_b200.models.vggt import IGGT, VGGT
from iggt_official_b200.graphs import GraphedForward

CONFIGS = {
    "C1": dict(model="iggt", B=1, S=3, H=336, W=504, dtype="bfloat16"),        # demo.py shape
    "C2": dict(model="vggt", B=1, S=8, H=518, W=518, dtype="float16"),
    "C2part": dict(model="iggt", B=1, S=8, H=532, W=532, dtype="float16"),     # even grid: part head runs
    "C2bf16": dict(model="vggt", B=1, S=8, H=518, W=518, dtype="bfloat16"),
    "C3local": dict(model="vggt", B=1, S=32, H=518, W=518, dtype="float16"),   # C3 without sharding (fits one GPU)
    "C4": dict(model="iggt", B=1, S=8, H=1036, W=1036, dtype="float16"),
    "C5local": dict(model="vggt", B=4, S=16, H=518, W=518, dtype="bfloat16"),
}


def main():
    names = sys.argv[1:] or ["C2"]
    out = {}
    models = {}
    for name in names:
        c = CONFIGS[name]
        if c["model"] not in models:
            torch.manual_seed(0)
            models[c["model"]] = (IGGT if c["model"] == "iggt" else VGGT)().eval().cuda()
        m = models[c["model"]]
        m.compute_dtype = getattr(torch, c["dtype"])
        g = torch.Generator().manual_seed(0)
        x = torch.rand(c["B"], c["S"], 3, c["H"], c["W"], generator=g).cuda()
        torch.cuda.reset_peak_memory_stats()
        try:
            fwd = GraphedForward(lambda im: m(im))
            o = fwd(x); torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            reps = 3
            e0.record()
            for _ in range(reps):
                o = fwd(x)
            e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / reps
            finite = all(bool(torch.isfinite(v).all()) for k, v in o.items() if torch.is_tensor(v))
            out[name] = dict(c, ms=ms, views_per_s=c["B"] * c["S"] / ms * 1e3, finite=finite,
                             peak_mem_gb=torch.cuda.max_memory_allocated() / 2 ** 30,
                             keys={k: list(v.shape) for k, v in o.items() if torch.is_tensor(v)})
        except Exception as e:
            out[name] = dict(c, error=f"{type(e).__name__}: {e}"[:500])
        print(name, json.dumps(out[name])[:600], flush=True)
        del x
        torch.cuda.empty_cache()
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/configs.json", "w"), indent=1)


if __name__ == "__main__":
    main()
