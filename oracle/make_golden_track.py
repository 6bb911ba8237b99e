"""Generates tests/golden/track_vggt_s3_140x154.pt by running the UNMODIFIED reference `VGGT.forward(images,
query_points)` (imported from /root/reference with oracle/shims.py) on seeded inputs and the synthetic checkpoint of
oracle/weights.py.  Run here only:   python oracle/make_golden_track.py"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import shims, weights  # noqa: E402

CASE = dict(name="track_vggt_s3_140x154", B=1, S=3, H=140, W=154, kind="stress", wseed=5, iseed=21, N=9)


def query_points(case):
    g = torch.Generator().manual_seed(case["iseed"] + 1)
    xy = torch.rand(case["N"], 2, generator=g) * torch.tensor([case["W"] - 1.0, case["H"] - 1.0])
    xy[0] = torch.tensor([0.0, 0.0])                       # a corner and a border point: clamped / zero-padded sampling
    xy[1] = torch.tensor([case["W"] - 1.0, 3.5])
    return xy


def main():
    shims.install()
    from iggt.models.vggt import VGGT
    torch.set_grad_enabled(False)
    c = CASE
    m = VGGT().eval()
    sd = weights.make_state_dict(c["wseed"], c["kind"])
    missing, _ = m.load_state_dict(sd, strict=False)
    assert not missing, missing
    g = torch.Generator().manual_seed(c["iseed"])
    images = torch.rand(c["S"], 3, c["H"], c["W"], generator=g)
    qp = query_points(c)
    lists = {"x_in": [], "delta": []}
    hook = m.track_head.register_forward_hook(lambda mod, inp, out: lists.update(all_iters=torch.stack(out[0], 0)))
    fh = m.track_head.feature_extractor.register_forward_hook(lambda mod, inp, out: lists.update(fmaps=out.clone()))
    # the refinement loop is chaotic on synthetic weights: record the transformer input / output of EVERY iteration so
    # that the restatement can be pinned teacher-forced (tests/test_oracle_track.py)
    def record(mod, inp, out):                  # (a hook that returns something would replace the module's output)
        lists["x_in"].append(inp[0].clone())
        lists["delta"].append(out[0].clone())

    uh = m.track_head.tracker.updateformer.register_forward_hook(record)
    out = m(images, query_points=qp)
    hook.remove(); fh.remove(); uh.remove()
    rec = {"case": c, "query_points": qp, "track": out["track"], "vis": out["vis"], "conf": out["conf"],
           "track_all_iters": lists["all_iters"], "fmaps_mean": lists["fmaps"].mean((3, 4)),
           "fmaps_corner": lists["fmaps"][..., :4, :4].clone(),
           "x_in": torch.stack(lists["x_in"], 0), "delta": torch.stack(lists["delta"], 0)}
    path = os.path.join(ROOT, "tests", "golden", c["name"] + ".pt")
    torch.save(rec, path)
    print(path, os.path.getsize(path), "track", tuple(out["track"].shape), "fmaps", tuple(lists["fmaps"].shape))
    print("track range", float(out["track"].min()), float(out["track"].max()), "vis", float(out["vis"].mean()))


if __name__ == "__main__":
    main()
