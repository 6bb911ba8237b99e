"""Generates tests/golden/demo1_ref.json: the UNMODIFIED reference loader `iggt.utils.load_fn.load_and_preprocess_images`
(PIL bicubic resize + ToTensor on the CPU, iggt/utils/load_fn.py:12-128) applied to the three demo1 views exactly as
demo.py:181-186 does (mode="resize", DEFAULT_IMAGE_SIZE), reduced to a SHA-256 of the 8-bit batch plus per-view means.
The three JPEGs themselves are copied (data, not source) to tests/golden/demo1/ so that the GPU test can run config C1
(BASELINE.json configs[0]) without /root/reference.  Run here only:   python oracle/make_golden_demo1.py"""
import glob
import hashlib
import json
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import shims  # noqa: E402

SIZE = (504, 336)          # demo.py DEFAULT_IMAGE_SIZE (width, height)


def main():
    shims.install()
    from iggt.utils.load_fn import load_and_preprocess_images
    paths = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "demo1", "*.jpg")))
    assert len(paths) == 3
    for p in paths:                                                         # byte-identical copies of the reference's files
        ref = os.path.join("/root/reference/iggt_demo/demo1/images", os.path.basename(p))
        assert open(p, "rb").read() == open(ref, "rb").read()
    batch = load_and_preprocess_images(paths, mode="resize", resize_target_size=SIZE)
    u8 = (batch * 255.0).round().to(torch.uint8)
    assert torch.equal(u8.float() / 255.0, batch)                          # ToTensor: every value is k / 255
    rec = {"files": [os.path.basename(p) for p in paths], "mode": "resize", "resize_target_size": list(SIZE),
           "shape": list(batch.shape), "sha256_u8": hashlib.sha256(u8.numpy().tobytes()).hexdigest(),
           "view_means": [float(batch[i].double().mean()) for i in range(3)]}
    json.dump(rec, open(os.path.join(ROOT, "tests", "golden", "demo1_ref.json"), "w"), indent=1)
    print(rec)


if __name__ == "__main__":
    main()
