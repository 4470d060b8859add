"""CPU, world_size 2, gloo: the camera-sharded step (dmesh_renderer_b200/multiview.py) must give
every rank the same view-summed scene gradients as one process rendering all views -- the only
collective is one all-reduce of the packed (6P+F) buffer."""
import os
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _scene():
    from dmesh_renderer_b200 import scenes
    return scenes.random_tri_scene("mv", 31, 400, 0.12, 48, 64, B=4)


def _cot(color, depth):
    return color.detach() * 0.5 - 0.25, depth.detach() - 0.5


def _run_rank(rank, ws, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=ws)
    try:
        from oracle_renderer import OracleTriRenderer
        from dmesh_renderer_b200.multiview import PackedSceneGrads, multiview_step, shard_views
        s = _scene()
        mine = shard_views(s.mv_mats.shape[0], rank, ws)
        sl = slice(mine.start, mine.stop)
        g = PackedSceneGrads(s.verts.clone(), s.verts_color.clone(), s.faces_opacity.clone())
        multiview_step(OracleTriRenderer(s.H, s.W, s.bg), g, s.faces, s.mv_mats[sl], s.proj_mats[sl], s.verts_depth[sl],
                       s.faces_intense[sl], _cot, views_per_call=1)
        torch.save(g.flat.clone(), os.path.join(out, "rank%d.pt" % rank))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_two_ranks_match_single_process(tmp_path):
    import socket
    with socket.socket() as sk:
        sk.bind(("127.0.0.1", 0))
        port = sk.getsockname()[1]
    mp.spawn(_run_rank, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    r0, r1 = torch.load(tmp_path / "rank0.pt"), torch.load(tmp_path / "rank1.pt")
    assert torch.equal(r0, r1)                       # all-reduce leaves identical buffers everywhere

    from oracle_renderer import OracleTriRenderer
    from dmesh_renderer_b200.multiview import PackedSceneGrads, multiview_step
    s = _scene()
    g = PackedSceneGrads(s.verts.clone(), s.verts_color.clone(), s.faces_opacity.clone())
    multiview_step(OracleTriRenderer(s.H, s.W, s.bg), g, s.faces, s.mv_mats, s.proj_mats, s.verts_depth, s.faces_intense, _cot)
    rel = ((r0 - g.flat).norm() / g.flat.norm()).item()
    assert g.flat.abs().sum() > 0
    assert rel < 1e-5, rel
    P, F = s.verts.shape[0], s.faces.shape[0]
    assert g.flat.numel() == 6 * P + F               # the packed buffer of SURVEY.md 8e
