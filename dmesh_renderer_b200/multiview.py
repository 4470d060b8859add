"""Camera-sharded multi-view optimisation step (BASELINE.json north_star).

The reference has no distributed code at all; views are independent, and only
three gradient tensors are summed over views (dL_dverts, dL_dvcolor,
dL_dfopacity: cuda_rasterizer/backward.cu:389-407,415 carry no batch offset).
So: the scene is replicated on every rank, rank r renders its slice of the
cameras, and ONE all-reduce(SUM) over a single packed fp32 buffer
[dL_dverts | dL_dvcolor | dL_dfopacity] = (6P + F) floats exchanges everything
that has to be exchanged.  Per-view gradients (verts_depth, faces_intense rows)
stay rank-local.  One process per GPU, torch.distributed (NCCL over NVLink 5 /
NVSwitch on the B200 box, gloo in the CPU tests); nothing else communicates.
"""
from typing import Callable, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def shard_views(num_views: int, rank: int, world_size: int) -> range:
    """Contiguous slice of the cameras owned by `rank` (rank r renders
    views r*B/G .. (r+1)*B/G - 1; remainders go to the first ranks)."""
    base, rem = divmod(num_views, world_size)
    start = rank * base + min(rank, rem)
    return range(start, start + base + (1 if rank < rem else 0))


class PackedSceneGrads:
    """Scene leaves (verts, verts_color, faces_opacity) whose .grad tensors are
    views into ONE contiguous fp32 buffer, so the step needs a single collective."""

    def __init__(self, verts: torch.Tensor, verts_color: torch.Tensor, faces_opacity: torch.Tensor):
        self.leaves = [verts, verts_color, faces_opacity]
        n = sum(t.numel() for t in self.leaves)
        self.flat = torch.zeros(n, dtype=torch.float32, device=verts.device)
        o = 0
        for t in self.leaves:
            if not t.requires_grad:
                t.requires_grad_(True)
            t.grad = self.flat[o:o + t.numel()].view_as(t)
            o += t.numel()

    def zero_(self):
        self.flat.zero_()

    def all_reduce(self, group=None, async_op=False):
        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            return dist.all_reduce(self.flat, op=dist.ReduceOp.SUM, group=group, async_op=async_op)
        return None


def multiview_step(render: Callable, scene_grads: PackedSceneGrads, faces: torch.Tensor, mv_mats: torch.Tensor,
                   proj_mats: torch.Tensor, verts_depth: torch.Tensor, faces_intense: torch.Tensor,
                   cotangent_fn: Callable[[torch.Tensor, torch.Tensor], Tuple[torch.Tensor, torch.Tensor]],
                   group=None, views_per_call: Optional[int] = None) -> Sequence[torch.Tensor]:
    """One optimisation step on this rank's cameras.

    render            a TriRenderer-like callable (verts, faces, verts_color, faces_opacity,
                      mv, proj, verts_depth, faces_intense) -> (color, depth)
    mv_mats ...       THIS RANK's slice: [B_local,4,4], [B_local,P], [B_local,F]
    cotangent_fn      (color, depth) -> (dL_dcolor, dL_ddepth)   (e.g. gradient of an image loss)
    views_per_call    split the local views into calls of this many views (None = one call)

    After the call every rank holds the view-summed gradients of the whole job in
    scene_grads.leaves[i].grad.  Returns the list of (color, depth) outputs.
    """
    verts, verts_color, faces_opacity = scene_grads.leaves
    scene_grads.zero_()
    B = mv_mats.shape[0]
    step = views_per_call or max(B, 1)
    outs = []
    for s in range(0, B, step):
        e = min(B, s + step)
        color, depth = render(verts, faces, verts_color, faces_opacity, mv_mats[s:e], proj_mats[s:e], verts_depth[s:e],
                              faces_intense[s:e])
        gc, gd = cotangent_fn(color, depth)
        torch.autograd.backward([color, depth], [gc, gd])
        outs.append((color.detach(), depth.detach()))
    scene_grads.all_reduce(group)
    return outs
