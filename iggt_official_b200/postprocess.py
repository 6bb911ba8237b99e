"""Device-side post-processing of a prediction dict.

Same names and argument meaning as the reference helpers `iggt.utils.pose_enc.pose_encoding_to_extri_intri` and
`iggt.utils.geometry.unproject_depth_map_to_point_map`, but on CUDA tensors, so demo.py's
"predictions -> .cpu().numpy() -> per-frame numpy loop" (demo.py:340-355) becomes two kernel launches."""
import torch

from . import ops


def pose_encoding_to_extri_intri(pose_encoding: torch.Tensor, image_size_hw=None, pose_encoding_type="absT_quaR_FoV",
                                 build_intrinsics=True):
    if pose_encoding_type != "absT_quaR_FoV":
        raise NotImplementedError
    H, W = image_size_hw if image_size_hw is not None else (0, 0)
    return ops.pose_to_cameras(pose_encoding.float(), int(H), int(W), build_intrinsics and image_size_hw is not None)


def unproject_depth_map_to_point_map(depth_map: torch.Tensor, extrinsics_cam: torch.Tensor, intrinsics_cam: torch.Tensor):
    """depth [S,H,W,1] or [S,H,W]; extrinsics [S,3,4]; intrinsics [S,3,3]  ->  world points [S,H,W,3] (CUDA)."""
    world, _ = ops.unproject_depth(depth_map.float(), extrinsics_cam.float(), intrinsics_cam.float())
    return world
