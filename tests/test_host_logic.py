"""CPU: host-side logic of the drop-in boundary -- validation messages of the `_C` shim
(reference render.cu:49-79, 237-277), the CUDA-only guard (no CPU fallback), camera sharding,
seeded scene generation."""
import pytest
import torch

from dmesh_renderer_b200 import (TetRenderer, TetRenderSettings, TriRenderer, TriRenderSettings, _C, scenes)
from dmesh_renderer_b200.multiview import shard_views


def tri_args(s):
    mv, pj = s.mv_mats.transpose(1, 2), s.proj_mats.transpose(1, 2)
    return [s.bg, s.verts, s.faces, s.verts_color, s.faces_opacity, mv, pj, torch.inverse(mv), torch.inverse(pj),
            s.verts_depth, s.faces_intense, s.H, s.W]


@pytest.mark.parametrize("idx,bad,msg", [
    (1, lambda t: t[:, :2], "verts must have dimensions"),
    (2, lambda t: t[:, :2], "faces must have dimensions"),
    (3, lambda t: t[:5], "vert color must have dimensions"),
    (4, lambda t: t[:5], "face opacity must have dimensions"),
    (5, lambda t: t[:, :3], "mv_mats must have dimensions"),
    (6, lambda t: t[0], "proj_mats must have dimensions"),
    (7, lambda t: t[:, :, :2], "inv_mv_mats must have dimensions"),
    (9, lambda t: t[:, :7], "verts_depth must have dimensions"),
    (10, lambda t: t[:, :7], "faces_intense must have dimensions"),
])
def test_tri_validation_messages(idx, bad, msg):
    a = tri_args(scenes.config("tiny_tri"))
    a[idx] = bad(a[idx])
    with pytest.raises(RuntimeError, match=msg):
        _C.render_tris(*a)


def test_tet_validation_messages():
    s = scenes.config("tiny_tet")
    mv, pj = s.mv_mats.transpose(1, 2), s.proj_mats.transpose(1, 2)
    a = [s.bg, s.verts, s.faces, s.verts_color, s.faces_opacity, mv, pj, torch.inverse(mv), torch.inverse(pj),
         s.verts_depth, s.faces_intense, s.tets, s.face_tets, s.tet_faces, s.H, s.W, 0]
    for idx, bad, msg in [(11, lambda t: t[:, :3], "tets must have dimensions"),
                          (12, lambda t: t[:4], "face_tets must have dimensions"),
                          (13, lambda t: t[:4], "tet_faces must have dimensions"),
                          (3, lambda t: t[:, :2], "vert_color must have dimensions")]:
        b = list(a)
        b[idx] = bad(b[idx])
        with pytest.raises(RuntimeError, match=msg):
            _C.render_tets(*b)


def test_no_cpu_fallback():
    a = tri_args(scenes.config("tiny_tri"))
    with pytest.raises(RuntimeError, match="CUDA-only"):
        _C.render_tris(*a)
    s = scenes.config("tiny_tri")
    with pytest.raises(RuntimeError, match="CUDA-only"):
        TriRenderer(TriRenderSettings(s.H, s.W, s.bg))(s.verts, s.faces, s.verts_color, s.faces_opacity, s.mv_mats,
                                                       s.proj_mats, s.verts_depth, s.faces_intense)


def test_api_surface_matches_reference_names():
    import dmesh_renderer_b200 as m
    for n in ["TriRenderSettings", "render_tri", "TriRenderer", "TetRenderSettings", "render_tet", "TetRenderer"]:
        assert hasattr(m, n)
    assert TriRenderSettings._fields == ("image_height", "image_width", "bg")
    assert TetRenderSettings._fields == ("image_height", "image_width", "bg", "ray_random_seed")
    for n in ["render_tris", "render_tris_backward", "render_tets", "render_tets_backward"]:
        assert callable(getattr(_C, n))
    assert isinstance(TetRenderer(TetRenderSettings(8, 8, torch.ones(3), 0)), torch.nn.Module)


def test_shard_views_partitions_all_cameras():
    for n, ws in [(64, 1), (64, 2), (64, 8), (7, 4), (3, 8)]:
        got = [v for r in range(ws) for v in shard_views(n, r, ws)]
        assert got == list(range(n))
        sizes = [len(shard_views(n, r, ws)) for r in range(ws)]
        assert max(sizes) - min(sizes) <= 1


def test_scenes_are_deterministic_and_sized():
    a, b = scenes.config("tiny_tri"), scenes.config("tiny_tri")
    assert torch.equal(a.verts, b.verts) and torch.equal(a.faces_intense, b.faces_intense)
    s = scenes.config("C1")
    assert s.faces.shape == (10_000, 3) and s.verts.shape == (30_000, 3) and (s.H, s.W) == (256, 256)
    t = scenes.config("tiny_tet")
    T, F = t.tets.shape[0], t.faces.shape[0]
    assert T == 6 * 4 ** 3 and t.tet_faces.shape == (T, 4) and t.face_tets.shape == (F, 2)
    # every interior face has two tets, every boundary face one
    assert int((t.face_tets[:, 1] < 0).sum()) == 6 * 2 * 4 * 4


def test_sync_free_inverses_match_torch_inverse_and_raise_on_singular():
    """_Inverses = the reference's two th.inverse calls (__init__.py:62-63) without their device syncs."""
    from dmesh_renderer_b200 import _C
    g = torch.Generator().manual_seed(0)
    a, b = torch.randn(3, 4, 4, generator=g), torch.randn(3, 4, 4, generator=g)
    inv = _C._Inverses(a, b)
    assert torch.equal(inv.inv_mv, torch.inverse(a)) and torch.equal(inv.inv_proj, torch.inverse(b))
    a[1] = 0
    with pytest.raises(RuntimeError, match="singular"):
        _C._Inverses(a, b)


def test_cpu_binding_helper_is_a_no_op_without_a_gpu():
    """multiview.bind_to_gpu_cpus must never fail a job: without NVML / a GPU it returns None and leaves the
    process's CPU affinity alone."""
    import os
    from dmesh_renderer_b200.multiview import bind_to_gpu_cpus
    before = os.sched_getaffinity(0)
    if torch.cuda.is_available():
        pytest.skip("CPU-only check")
    assert bind_to_gpu_cpus(0) is None
    assert os.sched_getaffinity(0) == before


def test_gradient_sink_matching_rules():
    """_C.grad_sink_for: a renderer call is redirected into a PackedSceneGrads buffer only while the sink is active,
    only for the sink's own leaves with live .grad views, and the sink stack unwinds on exceptions."""
    from dmesh_renderer_b200 import _C
    from dmesh_renderer_b200.multiview import PackedSceneGrads
    v, c, o = torch.randn(5, 3), torch.rand(5, 3), torch.rand(4)
    g = PackedSceneGrads(v, c, o)
    assert g.flat.numel() == 34 and all(leaf.grad.data_ptr() >= g.flat.data_ptr() for leaf in g.leaves)
    assert _C.grad_sink_for(*g.leaves) is None                      # not active
    with g.direct():
        assert _C.grad_sink_for(*g.leaves) is g
        assert _C.grad_sink_for(v.clone(), c, o) is None             # a copy is not the leaf
        assert _C.grad_sink_for(v * 1.0, c, o) is None               # neither is a function of it
        assert _C.grad_sink_for(c, v, o) is None                     # order matters
        other = PackedSceneGrads(v.detach().clone(), c.detach().clone(), o.detach().clone())
        with other.direct():                                         # innermost matching sink wins, outer one still found
            assert _C.grad_sink_for(*other.leaves) is other
            assert _C.grad_sink_for(*g.leaves) is g
        assert _C.grad_sink_for(*other.leaves) is None
        saved = v.grad
        v.grad = None                                                # e.g. optimizer.zero_grad(set_to_none=True)
        assert _C.grad_sink_for(*g.leaves) is None
        v.grad = saved
    assert not _C._grad_sinks
    with pytest.raises(ValueError):
        with g.direct():
            raise ValueError("boom")
    assert not _C._grad_sinks


def test_packed_scene_grads_rebinds_dropped_grads():
    """optimizer.zero_grad(set_to_none=True) (or any code that replaces .grad) must not silently
    disconnect a leaf from the packed buffer that all_reduce() exchanges."""
    from dmesh_renderer_b200.multiview import PackedSceneGrads
    g = PackedSceneGrads(torch.zeros(5, 3), torch.zeros(5, 3), torch.zeros(7))
    assert g.bind() == 0
    opt = torch.optim.SGD(g.leaves, lr=0.1)
    g.leaves[0].grad.fill_(1.0)
    opt.zero_grad()                                   # set_to_none=True: all three .grad become None
    assert all(leaf.grad is None for leaf in g.leaves)
    g.zero_()                                         # re-attaches
    for leaf in g.leaves:
        assert leaf.grad is not None and g.flat.data_ptr() <= leaf.grad.data_ptr() < g.flat.data_ptr() + 4 * g.flat.numel()
    # a gradient that autograd put into a foreign tensor is moved into the packed buffer, not lost
    g.leaves[2].grad = torch.full((7,), 2.5)
    assert g.bind() == 1
    assert torch.equal(g.flat[30:], torch.full((7,), 2.5)) and g.leaves[2].grad.data_ptr() == g.flat[30:].data_ptr()
    g.all_reduce()                                    # single process: no collective, still binds
    assert g.collective.startswith("none")


def test_alias_package_exposes_the_reference_names():
    import dmesh_renderer
    import dmesh_renderer_b200
    for n in ("TriRenderSettings", "render_tri", "TriRenderer", "TetRenderSettings", "render_tet", "TetRenderer"):
        assert getattr(dmesh_renderer, n) is getattr(dmesh_renderer_b200, n)
    assert dmesh_renderer._C is dmesh_renderer_b200._C


def test_tet_record_cache_keys_on_identity_and_version():
    from dmesh_renderer_b200._C import _TetRecordCache

    class Lib:
        @staticmethod
        def dmr_tet_records_bytes(T):
            return 128 * T + 256
    _TetRecordCache.clear()
    dev = torch.device("cpu")
    _TetRecordCache._entries.clear()
    ts = tuple(torch.zeros(4, 3) for _ in range(5))
    import unittest.mock as mock
    with mock.patch("torch.cuda.current_device", return_value=0):
        b0, valid0 = _TetRecordCache.get(Lib, ts, 10, dev)
        b1, valid1 = _TetRecordCache.get(Lib, ts, 10, dev)
        assert (valid0, valid1) == (0, 1) and b1 is b0
        ts[0].add_(1.0)                               # in-place edit of the geometry: version bump -> rebuild
        b2, valid2 = _TetRecordCache.get(Lib, ts, 10, dev)
        assert valid2 == 0 and b2 is not b0
        other = tuple(t.clone() for t in ts)          # equal values, different tensors -> no hit
        assert _TetRecordCache.get(Lib, other, 10, dev)[1] == 0
    _TetRecordCache.clear()
