"""GPU parity proper: the CUDA path (through the `_C` shim -> C ABI) against
  (1) the committed golden fixtures produced by the reference extension,
  (2) the CPU oracle on further seeded scenes (edge cases: non-multiple-of-16 images, B>1,
      huge / off-screen / degenerate / behind-camera triangles, shared vertices),
  (3) size-independent properties at BASELINE.json's full sizes (C2, C5, C3).
Bit-exact for integer work, 1e-5 max-abs for images, 1e-4 relative L2 for gradients.
"""
import os
import sys

import numpy as np
import pytest

import binning_checks
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402
from dmesh_renderer_b200 import _C, debug, scenes  # noqa: E402

pytestmark = pytest.mark.gpu
GOLD = os.path.join(ROOT, "tests", "golden")
IMG_TOL, GRAD_TOL = 1e-5, 1e-4


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def dev_mats(s, mats=None):
    if mats is not None:
        return [torch.from_numpy(np.ascontiguousarray(m)).cuda() for m in mats]
    mv, pj = s.mv_mats.transpose(1, 2).contiguous(), s.proj_mats.transpose(1, 2).contiguous()
    return [mv, pj, torch.inverse(mv), torch.inverse(pj)]


def run_tri(s, mats, gc=None, gd=None):
    a = (s.bg, s.verts, s.faces, s.verts_color, s.faces_opacity, *mats, s.verts_depth, s.faces_intense)
    out = _C.render_tris(*a, s.H, s.W)
    grads = None
    if gc is not None:
        grads = _C.render_tris_backward(*a, gc, gd, out[0], out[3], out[4], out[5], out[6])
    return out, grads


def run_tet(s, mats, gc=None, gd=None, seed=0):
    a = (s.bg, s.verts, s.faces, s.verts_color, s.faces_opacity, *mats, s.verts_depth, s.faces_intense, s.tets,
         s.face_tets, s.tet_faces)
    out = _C.render_tets(*a, s.H, s.W, seed)
    grads = None
    if gc is not None:
        grads = _C.render_tets_backward(*a, gc, gd, out[3], out[4], out[5], out[6], seed)
    return out, grads


def tri_views(s, out):
    B, P, F = s.mv_mats.shape[0], s.verts.shape[0], s.faces.shape[0]
    d = dict(B=B, P=P, F=F, W=s.W, H=s.H, R=out[0])
    v = {k: debug.view("tri", k, out[4], **d) for k in ("tiles_touched", "offsets", "depth_keys", "face_order")}
    v.update({k: debug.view("tri", k, out[5], face_buffer=out[4], **d)
              for k in ("keys_unsorted", "values_unsorted", "keys_sorted", "values_sorted")})
    v.update({k: debug.view("tri", k, out[6], **d) for k in ("ranges", "n_contrib", "final_T")})
    v["verts_image"] = debug.view("tri", "verts_image", out[3], **d)
    return v


def check_tri_against(v, out, grads, ref, ref_grads, exact_T=False):
    np.testing.assert_array_equal(v["tiles_touched"], ref["tiles_touched"])
    binning_checks.check_face_order_and_offsets(v["face_order"], v["depth_keys"], v["tiles_touched"], v["offsets"],
                                                ref["offsets"])
    assert out[0] == int(ref["R"])
    live = ref["tiles_touched"] > 0
    np.testing.assert_array_equal(v["depth_keys"][live], ref["depth_keys"][live])
    np.testing.assert_array_equal(v["keys_sorted"], ref["keys_sorted"])
    np.testing.assert_array_equal(v["values_sorted"], ref["values_sorted"])
    np.testing.assert_array_equal(v["ranges"], ref["ranges"])
    np.testing.assert_array_equal(v["n_contrib"], ref["n_contrib"])
    assert np.abs(out[1].cpu().numpy() - ref["color"]).max() <= IMG_TOL
    assert np.abs(out[2].cpu().numpy() - ref["depth"]).max() <= IMG_TOL
    if grads is not None:
        for i, (g, r) in enumerate(zip(grads, ref_grads)):
            e = rel_l2(g.cpu().numpy(), r)
            assert e <= GRAD_TOL, "grad %d rel L2 %.3e" % (i, e)


# ------------------------------------------------------------------ (1) golden fixtures
@pytest.mark.parametrize("name", ["tiny_tri", "small_tri"])
def test_tri_cuda_matches_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cpu = scenes.config(name)
    s = scenes.to_device(cpu, "cuda")
    gc, gd = [t.cuda() for t in scenes.cotangents(cpu)]
    out, grads = run_tri(s, dev_mats(s, g["mats"]), gc, gd)
    v = tri_views(s, out)
    np.testing.assert_array_equal(v["verts_image"][:, :2].view(np.uint32), g["verts_image"].view(np.uint32))
    np.testing.assert_array_equal(v["final_T"].view(np.uint32), g["final_T"].view(np.uint32))
    check_tri_against(v, out, grads, g, [g[k] for k in ("g_verts", "g_verts_color", "g_faces_opacity", "g_verts_depth", "g_faces_intense")])


@pytest.mark.parametrize("name", ["tiny_tet", "small_tet"])
def test_tet_cuda_matches_golden(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    cpu = scenes.config(name)
    s = scenes.to_device(cpu, "cuda")
    gc, gd = [t.cuda() for t in scenes.cotangents(cpu)]
    out, grads = run_tet(s, dev_mats(s, g["mats"]), gc, gd)
    B, P, F, T = s.mv_mats.shape[0], s.verts.shape[0], s.faces.shape[0], s.tets.shape[0]
    d = dict(B=B, P=P, F=F, W=s.W, H=s.H, R=int(g["R"]), T=T)
    np.testing.assert_array_equal(debug.view("tet", "tiles_touched", out[4], **d), g["tiles_touched"])
    np.testing.assert_array_equal(debug.view("tet", "keys_sorted", out[5], face_buffer=out[4], **d), g["keys_sorted"])
    np.testing.assert_array_equal(debug.view("tet", "values_sorted", out[5], **d), g["values_sorted"])
    np.testing.assert_array_equal(debug.view("tet", "ranges", out[6], **d), g["ranges"])
    np.testing.assert_array_equal(debug.view("tet", "first_face", out[6], **d), g["first_face"])
    np.testing.assert_array_equal(debug.view("tet", "first_tet", out[6], **d), g["first_tet"])
    np.testing.assert_array_equal(debug.view("tet", "n_contrib", out[6], **d), g["n_contrib"])
    assert np.array_equal(out[2].cpu().numpy() > 0.5, g["active"] > 0.5)
    assert np.abs(out[0].cpu().numpy() - g["color"]).max() <= IMG_TOL
    assert np.abs(out[1].cpu().numpy() - g["depth"]).max() <= IMG_TOL
    assert rel_l2(grads[0].cpu().numpy(), g["g_verts_color"]) <= GRAD_TOL
    assert rel_l2(grads[1].cpu().numpy(), g["g_faces_opacity"]) <= GRAD_TOL


# ------------------------------------------------------------------ (2) CPU oracle, edge cases
def edge_scene():
    """Shared vertices, a full-screen triangle, slivers, zero-area and off-screen / behind-camera faces,
    opacity exactly 1 and 0, image size not a multiple of 16, two views."""
    g = torch.Generator().manual_seed(77)
    base = scenes.random_tri_scene("edge", 78, 600, 0.2, 90, 110, B=2)
    verts = base.verts.clone()
    faces = base.faces.clone()
    faces[1::3, 0] = faces[0::3, 0][: faces[1::3].shape[0]]          # shared vertices between neighbours
    verts[0:3] = torch.tensor([[-4.0, -4.0, 0.0], [4.0, -4.0, 0.0], [0.0, 5.0, 0.2]])   # covers the whole screen
    verts[3:6] = torch.tensor([[0.1, 0.1, 0.1], [0.1, 0.1, 0.1], [0.3, 0.2, 0.0]])      # zero area
    verts[6:9] = torch.tensor([[0.0, 0.0, 9.0], [0.2, 0.0, 9.0], [0.0, 0.2, 9.0]])      # behind / beyond planes
    verts[9:12] = torch.tensor([[30.0, 0.0, 0.0], [31.0, 0.0, 0.0], [30.0, 1.0, 0.0]])  # far off screen
    verts[12:15] = torch.tensor([[-0.5, 0.0, 0.0], [0.5, 1e-4, 0.0], [0.0, 5e-5, 0.0]])  # sliver
    op = base.faces_opacity.clone()
    op[5], op[7], op[20:40] = 1.0, 0.0, 1.0
    depth = scenes.ndc_depth(verts, base.mv_mats, base.proj_mats)
    return base._replace(verts=verts, faces=faces, faces_opacity=op, verts_depth=depth)


@pytest.mark.parametrize("make", [edge_scene, lambda: scenes.random_tri_scene("mid", 5, 6000, 0.06, 176, 208, B=3),
                                  # 8 views: B*F >= 2*P, the per-vertex vector accumulators of tri_grad_finish
                                  lambda: scenes.random_tri_scene("mv8", 6, 2500, 0.07, 96, 112, B=8)])
def test_tri_cuda_matches_oracle(make):
    cpu = make()
    s = scenes.to_device(cpu, "cuda")
    gc, gd = [t.cuda() for t in scenes.cotangents(cpu)]
    mats = dev_mats(s)
    out, grads = run_tri(s, mats, gc, gd)
    o = oracle.TriOracle(cpu, mats=[m.contiguous().cpu().numpy() for m in mats])
    ref = o.outputs()
    v = tri_views(s, out)
    binning_checks.check_same_pairs(v["keys_unsorted"], v["values_unsorted"], ref["keys_unsorted"], ref["values_unsorted"])
    check_tri_against(v, out, grads, ref, o.backward(gc.cpu(), gd.cpu()))


def test_tet_cuda_matches_oracle_odd_image_two_views():
    cpu = scenes.tet_grid_scene("tet_odd", 41, 6, 72, 88, B=2, opacity=(0.0, 0.6))
    op = cpu.faces_opacity.clone()
    op[::17] = 1.0                                                   # exercises the log(T_EPS*0.1) branch
    cpu = cpu._replace(faces_opacity=op)
    s = scenes.to_device(cpu, "cuda")
    gc, gd = [t.cuda() for t in scenes.cotangents(cpu)]
    mats = dev_mats(s)
    out, grads = run_tet(s, mats, gc, gd)
    o = oracle.TetOracle(cpu, mats=[m.contiguous().cpu().numpy() for m in mats])
    ref = o.outputs()
    B, P, F, T = s.mv_mats.shape[0], s.verts.shape[0], s.faces.shape[0], s.tets.shape[0]
    d = dict(B=B, P=P, F=F, W=s.W, H=s.H, R=ref["R"], T=T)
    np.testing.assert_array_equal(debug.view("tet", "first_face", out[6], **d), ref["first_face"])
    np.testing.assert_array_equal(debug.view("tet", "first_tet", out[6], **d), ref["first_tet"])
    np.testing.assert_array_equal(debug.view("tet", "n_contrib", out[6], **d), ref["n_contrib"])
    assert np.array_equal(out[2].cpu().numpy() > 0.5, ref["active"] > 0.5)
    assert np.abs(out[0].cpu().numpy() - ref["color"]).max() <= IMG_TOL
    assert np.abs(out[1].cpu().numpy() - ref["depth"]).max() <= IMG_TOL
    og = o.backward(gc.cpu(), gd.cpu())
    assert rel_l2(grads[0].cpu().numpy(), og[0]) <= GRAD_TOL
    assert rel_l2(grads[1].cpu().numpy(), og[1]) <= GRAD_TOL


# ------------------------------------------------------------------ (3) properties at full size
@pytest.mark.parametrize("name", ["C2", "C5"])
def test_tri_full_size_properties(name):
    cpu = scenes.config(name)
    s = scenes.to_device(cpu, "cuda")
    gc, gd = [t.cuda() for t in scenes.cotangents(cpu)]
    mats = dev_mats(s)
    out, grads = run_tri(s, mats, gc, gd)
    v = tri_views(s, out)
    keys, R = v["keys_sorted"], out[0]
    assert keys.size == R == int(v["offsets"][-1])
    np.testing.assert_array_equal(np.cumsum(v["tiles_touched"][v["face_order"]], dtype=np.uint64).astype(np.uint32), v["offsets"])
    assert np.all(keys[1:] >= keys[:-1])                                             # sortedness
    # the sort is a permutation: order-independent checksums of (key, value) pairs agree
    mix = lambda k, val: int(np.bitwise_xor.reduce((k * np.uint64(0x9E3779B97F4A7C15)) ^ (val.astype(np.uint64) << np.uint64(7))))
    assert mix(v["keys_unsorted"], v["values_unsorted"]) == mix(keys, v["values_sorted"])
    assert int(v["keys_unsorted"].sum(dtype=np.uint64)) == int(keys.sum(dtype=np.uint64))
    # stability: equal keys keep emission (face-major) order -> values ascending inside runs of equal keys
    same = keys[1:] == keys[:-1]
    assert np.all(v["values_sorted"][1:][same] >= v["values_sorted"][:-1][same])
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    counts = np.bincount(tiles, minlength=v["ranges"].shape[0])
    np.testing.assert_array_equal(v["ranges"][:, 1] - v["ranges"][:, 0], counts)      # ranges partition the list
    color, depth = out[1], out[2]
    assert torch.isfinite(color).all() and torch.isfinite(depth).all()
    assert color.min() >= -1e-5 and color.max() <= 1 + 1e-4
    assert np.all((v["final_T"] >= 0) & (v["final_T"] <= 1))
    # idempotence: a second forward is bit-identical (no atomics in forward)
    out2, _ = run_tri(s, mats)
    assert torch.equal(out2[1], color) and torch.equal(out2[2], depth)
    # backward is linear in the cotangent
    _, g2 = run_tri(s, mats, 2 * gc, 2 * gd)
    for a, b in zip(grads, g2):
        assert rel_l2(b.cpu().numpy(), 2 * a.cpu().numpy()) <= GRAD_TOL
    assert all(torch.isfinite(t).all() for t in grads)


def test_tri_views_are_independent():
    """B=3 in one call == three B=1 calls (images bit-identical, view-summed grads within tolerance)."""
    cpu = scenes.random_tri_scene("ind", 6, 20000, 0.04, 256, 320, B=3)
    s = scenes.to_device(cpu, "cuda")
    gc, gd = [t.cuda() for t in scenes.cotangents(cpu)]
    mats = dev_mats(s)
    out, grads = run_tri(s, mats, gc, gd)
    acc = None
    for b in range(3):
        sb = s._replace(mv_mats=s.mv_mats[b:b + 1], proj_mats=s.proj_mats[b:b + 1], verts_depth=s.verts_depth[b:b + 1],
                        faces_intense=s.faces_intense[b:b + 1])
        ob, gb = run_tri(sb, [m[b:b + 1].contiguous() for m in mats], gc[b:b + 1].contiguous(), gd[b:b + 1].contiguous())
        assert torch.equal(ob[1][0], out[1][b]) and torch.equal(ob[2][0], out[2][b])
        assert rel_l2(gb[3][0].cpu().numpy(), grads[3][b].cpu().numpy()) <= GRAD_TOL
        acc = [x.clone() for x in gb[:3]] if acc is None else [x + y for x, y in zip(acc, gb[:3])]
    for a, g in zip(acc, grads[:3]):
        assert rel_l2(a.cpu().numpy(), g.cpu().numpy()) <= GRAD_TOL


def test_tet_full_size_properties():
    cpu = scenes.config("C3")
    s = scenes.to_device(cpu, "cuda")
    gc, gd = [t.cuda() for t in scenes.cotangents(cpu)]
    mats = dev_mats(s)
    out, grads = run_tet(s, mats, gc, gd)
    color, depth, active = out[0], out[1], out[2]
    assert torch.isfinite(color).all() and torch.isfinite(depth).all()
    act = active > 0.5
    assert act.float().mean().item() > 0.3
    # inactive pixels are pure background (cuda_renderer/forward.cu:807-814)
    assert torch.equal(color.permute(0, 2, 3, 1)[~act], torch.ones_like(color.permute(0, 2, 3, 1)[~act]))
    assert torch.equal(depth[:, 0][~act], torch.ones_like(depth[:, 0][~act]))
    out2, _ = run_tet(s, mats)
    assert torch.equal(out2[0], color) and torch.equal(out2[2], active)
    _, g2 = run_tet(s, mats, 2 * gc, 2 * gd)
    for a, b in zip(grads, g2):
        assert rel_l2(b.cpu().numpy(), 2 * a.cpu().numpy()) <= GRAD_TOL
