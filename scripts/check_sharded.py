"""torchrun --nproc-per-node N scripts/check_sharded.py : view-sharded forward == single-GPU forward.

Every rank runs the sharded forward on its views; rank r then runs the full (unsharded) forward locally and
compares its slice.  Differences come only from the softmax accumulation order (K/V arrive in the same order, so
they are expected to be bit-identical or ~1e-6)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    from iggt_official_b200.models.vggt import IGGT, VGGT
    from iggt_official_b200.parallel import forward_sharded
    B, S, H, W = 2, 4 * world // 2 if world > 1 else 4, 42, 56
    S = 2 * world
    torch.manual_seed(0)
    ok = True
    for cls, (H, W) in ((VGGT, (42, 56)), (IGGT, (28, 56))):
        torch.manual_seed(0)
        model = cls().eval().cuda()
        model.compute_dtype = torch.float16
        g = torch.Generator().manual_seed(1)
        images = torch.rand(B, S, 3, H, W, generator=g).cuda()
        S_loc = S // world
        mine = images[:, rank * S_loc:(rank + 1) * S_loc].contiguous()
        out = forward_sharded(model, mine, rank, world)
        model.aggregator.process_group = None
        ref = model(images)
        for k in ("depth", "depth_conf", "world_points", "world_points_conf", "part_feat"):
            if k not in ref:
                continue
            a, b = out[k], ref[k][:, rank * S_loc:(rank + 1) * S_loc]
            err = ((a - b).abs().max() / b.abs().max()).item()
            ok = ok and err < 1e-4
            print(f"[rank {rank}] {cls.__name__} {k}: max rel diff sharded vs single = {err:.2e}")
        pe = (torch.stack(out["pose_enc"]) - torch.stack(ref["pose_enc"])).abs().max().item()
        ok = ok and pe < 1e-4
        print(f"[rank {rank}] {cls.__name__} pose_enc max abs diff = {pe:.2e}")
    t = torch.tensor([1.0 if ok else 0.0], device="cuda")
    dist.all_reduce(t, op=dist.ReduceOp.MIN)
    if rank == 0:
        print("SHARDED_OK" if t.item() == 1.0 else "SHARDED_MISMATCH")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
