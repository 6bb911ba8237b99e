"""tests/golden/postprocess.pt from the UNMODIFIED reference helpers (run here only; needs /root/reference):
iggt.utils.pose_enc.pose_encoding_to_extri_intri and iggt.utils.geometry.unproject_depth_map_to_point_map."""
import os
import sys
import types

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

if __name__ == "__main__":
    # iggt/utils/__init__.py pulls in matplotlib / evo; load only the three files on this path through a synthetic
    # package whose __path__ is the reference directory (the files themselves are executed unmodified).
    import importlib
    for name in ("iggt", "iggt.utils", "iggt.utils.misc", "iggt.utils.device"):   # geometry.py:12-13 imports helpers
        sys.modules[name] = types.ModuleType(name)                                 # that this path never calls
    sys.modules["iggt.utils.misc"].invalid_to_zeros = sys.modules["iggt.utils.misc"].invalid_to_nans = None
    sys.modules["iggt.utils.device"].to_numpy = None
    pkg = types.ModuleType("refutils")
    pkg.__path__ = ["/root/reference/iggt/utils"]
    sys.modules["refutils"] = pkg
    pose_encoding_to_extri_intri = importlib.import_module("refutils.pose_enc").pose_encoding_to_extri_intri
    unproject_depth_map_to_point_map = importlib.import_module("refutils.geometry").unproject_depth_map_to_point_map
    g = torch.Generator().manual_seed(7)
    S, H, W = 3, 12, 20
    pose = torch.randn(1, S, 9, generator=g)
    pose[..., 7:] = pose[..., 7:].abs() * 0.3 + 0.6            # field of view (rad), positive like the ReLU'd output
    depth = torch.rand(S, H, W, 1, generator=g) * 5 + 0.1
    depth[0, 0, 0, 0] = 0.0                                     # invalid (masked) pixel
    depth[1, 3, 4, 0] = 150.0                                   # beyond z_far
    extr, intr = pose_encoding_to_extri_intri(pose, (H, W))
    world = unproject_depth_map_to_point_map(depth.numpy(), extr[0].numpy(), intr[0].numpy())
    torch.save({"pose": pose, "depth": depth, "H": H, "W": W, "extrinsic": extr, "intrinsic": intr,
                "world": torch.from_numpy(world).float()}, os.path.join(ROOT, "tests", "golden", "postprocess_ref.pt"))
    print("ok", extr.shape, intr.shape, world.shape)
