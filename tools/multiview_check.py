"""Multi-GPU end-to-end check: one optimisation step of B views sharded over the ranks
(multiview_step: real TriRenderer kernels, direct gradient sink, ONE all-reduce of the packed scene gradients --
NVLS kernel when the node has NVSwitch multicast, else NCCL) against the same B views rendered in one call on one GPU.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/multiview_check.py [B=8]
"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dmesh_renderer_b200 import TriRenderer, TriRenderSettings, scenes  # noqa: E402
from dmesh_renderer_b200.multiview import PackedSceneGrads, multiview_step, shard_views  # noqa: E402

rank, local, ws = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 8


def rel_l2(a, b):
    a, b = a.double(), b.double()
    return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()


for label, cpu in (("small", scenes.random_tri_scene("mvc", 77, 6000, 0.06, 192, 160, B=B)),
                   ("C4-like", scenes.random_tri_scene("mvc4", 78, 200_000, 0.01, 512, 512, B=B))):
    s = scenes.to_device(cpu, dev)
    gc, gd = [t.to(dev) for t in scenes.cotangents(cpu)]
    renderer = TriRenderer(TriRenderSettings(s.H, s.W, s.bg))

    # ---- reference: all B views in one call on this GPU, ordinary autograd accumulation
    ref = [s.verts.clone().requires_grad_(), s.verts_color.clone().requires_grad_(), s.faces_opacity.clone().requires_grad_()]
    vd0, fi0 = s.verts_depth.clone().requires_grad_(), s.faces_intense.clone().requires_grad_()
    c0, d0 = renderer(ref[0], s.faces, ref[1], ref[2], s.mv_mats, s.proj_mats, vd0, fi0)
    torch.autograd.backward([c0, d0], [gc, gd])

    # ---- sharded: this rank's views, two calls per step, direct sink, one all-reduce
    mine = shard_views(B, rank, ws)
    sl = slice(mine.start, mine.stop)
    g = PackedSceneGrads(s.verts.clone(), s.verts_color.clone(), s.faces_opacity.clone())
    vd1, fi1 = s.verts_depth[sl].clone().requires_grad_(), s.faces_intense[sl].clone().requires_grad_()
    calls = []

    def cotangents(c, d):
        first = mine.start + sum(calls)
        calls.append(c.shape[0])
        return gc[first:first + c.shape[0]], gd[first:first + c.shape[0]]

    outs = multiview_step(renderer, g, s.faces, s.mv_mats[sl], s.proj_mats[sl], vd1, fi1, cotangents,
                          views_per_call=max(1, len(mine) // 2))
    torch.cuda.synchronize()
    color = torch.cat([o[0] for o in outs])
    assert torch.equal(color, c0[sl]), "sharded forward differs from the single-GPU forward"
    errs = {n: rel_l2(leaf.grad, r.grad) for n, leaf, r in zip(("verts", "verts_color", "faces_opacity"), g.leaves, ref)}
    errs["verts_depth"] = rel_l2(vd1.grad, vd0.grad[sl])
    errs["faces_intense"] = rel_l2(fi1.grad, fi0.grad[sl])
    # every rank must hold the same reduced gradients
    mx = g.flat.clone()
    dist.all_reduce(mx, op=dist.ReduceOp.MAX)
    same = bool(torch.equal(mx, g.flat))
    if rank == 0:
        print("%s: B=%d over %d ranks, collective=%s, rel L2 vs one GPU: %s, identical on all ranks: %s" %
              (label, B, ws, g.collective, {k: "%.2e" % v for k, v in errs.items()}, same), flush=True)
    assert all(v <= 1e-4 for v in errs.values()), errs
    same_all = torch.tensor([1 if same else 0], device=dev)
    dist.all_reduce(same_all, op=dist.ReduceOp.MIN)
    assert int(same_all.item()) == 1, "ranks hold different reduced gradients"
    del g
if rank == 0:
    print("multiview_check ok", flush=True)
dist.destroy_process_group()
