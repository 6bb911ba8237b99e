"""GPU: BASELINE.json configs[0] (C1, "plumbing") end to end on the device: the three demo1 views
(iggt_demo/demo1/images, byte-identical copies under tests/golden/demo1/) -> `utils.load_fn.load_and_preprocess_images`
exactly as demo.py:181-186 calls it -> `IGGT.forward` under autocast as demo.py:191-195 -> `postprocess.*` as
demo.py:333-355, with the keys / shapes / dtypes demo.py:333-363 consumes.  The loader's batch must be BIT-IDENTICAL to
the unmodified reference loader's (SHA-256 of the 8-bit batch, oracle/make_golden_demo1.py); the model runs on synthetic
weights (no checkpoint is reachable offline) and is checked against the oracle on the same batch."""
import glob
import hashlib
import json
import os

import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_demo1_images_through_loader_model_and_postprocess():
    from oracle import ref_model, weights
    from iggt_official_b200 import postprocess
    from iggt_official_b200.models.vggt import IGGT
    from iggt_official_b200.utils.load_fn import load_and_preprocess_images
    rec = json.load(open(os.path.join(ROOT, "tests", "golden", "demo1_ref.json")))
    paths = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "demo1", "*.jpg")))
    assert [os.path.basename(p) for p in paths] == rec["files"]
    images = load_and_preprocess_images(paths, mode=rec["mode"], resize_target_size=tuple(rec["resize_target_size"]))
    assert images.is_cuda and images.dtype == torch.float32 and list(images.shape) == rec["shape"]
    u8 = (images * 255.0).round().to(torch.uint8).cpu()
    assert hashlib.sha256(u8.numpy().tobytes()).hexdigest() == rec["sha256_u8"]          # bit-exact with the reference loader
    sd = weights.make_state_dict(4, "default")
    model = IGGT()
    model.load_state_dict(sd, strict=False)
    model.eval().to("cuda")
    with torch.no_grad(), torch.amp.autocast("cuda", dtype=torch.bfloat16):             # demo.py:191-195 on sm >= 80
        predictions = model(images)
    S, H, W = 3, 336, 504
    assert predictions["images"].shape == (1, S, 3, H, W)
    assert len(predictions["pose_enc"]) == 4 and predictions["pose_enc"][-1].shape == (1, S, 9)
    for k, shape in (("depth", (1, S, H, W, 1)), ("depth_conf", (1, S, H, W)), ("world_points", (1, S, H, W, 3)),
                     ("world_points_conf", (1, S, H, W)), ("part_feat", (1, S, 8, H, W))):
        assert predictions[k].shape == shape and predictions[k].dtype == torch.float32 and torch.isfinite(predictions[k]).all(), k
    # demo.py:333-355
    pose = predictions["pose_enc"][-1]
    extrinsic, intrinsic = postprocess.pose_encoding_to_extri_intri(pose, images.shape[-2:])
    assert extrinsic.shape == (1, S, 3, 4) and intrinsic.shape == (1, S, 3, 3)
    world = postprocess.unproject_depth_map_to_point_map(predictions["depth"][0], extrinsic[0], intrinsic[0])
    assert world.shape == (S, H, W, 3) and torch.isfinite(world).all()
    part = predictions["part_feat"][0].permute(0, 2, 3, 1)                              # demo.py:363
    assert part.shape == (S, H, W, 8)
    # same batch through the oracle (bf16 autocast policy): the dict demo.py consumes carries the reference's values
    sdg = {k: v.cuda() for k, v in sd.items() if not k.startswith("track_head.")}
    ref = ref_model.forward(sdg, images, model="iggt", amp=torch.bfloat16, frames_chunk=3)
    for k in ("depth", "depth_conf", "world_points", "world_points_conf", "part_feat"):
        err = ((predictions[k] - ref[k]).norm() / ref[k].norm()).item()
        assert err < 2e-2, (k, err)                                                    # bf16 trunk: see test_fullsize_parity_gpu.py
    we, wi = ref_model.pose_encoding_to_extri_intri(pose, (H, W))
    assert (extrinsic - we).abs().max().item() < 1e-5
    # synthetic weights can put the ReLU'd field of view at exactly 0, i.e. an infinite focal length - in both
    fin = torch.isfinite(wi)
    assert torch.equal(torch.isfinite(intrinsic), fin) and torch.equal(intrinsic[~fin], wi[~fin])
    assert ((intrinsic[fin] - wi[fin]).abs() / wi[fin].abs().clamp_min(1.0)).max().item() < 1e-5
