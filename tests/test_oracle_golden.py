"""CPU: the oracle (oracle/oracle.cpp) against the golden fixtures generated from the
UNMODIFIED reference CUDA extension on a B200 (tests/golden/make_golden.py).

This is what pins the oracle: integer intermediates are compared exactly (with a
reported, bounded number of FMA-contraction ties allowed where a float feeds a
float->int conversion -- SURVEY.md 8c), images to 1e-5, gradients to 1e-4 rel L2.
"""
import hashlib
import os
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402
from dmesh_renderer_b200 import scenes  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
IMG_TOL = 1e-5
GRAD_TOL = 1e-4


def checksum(s):
    h = hashlib.sha256()
    for v in s:
        if isinstance(v, torch.Tensor):
            h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()


def rel_l2(a, b):
    a, b = np.asarray(a, np.float64), np.asarray(b, np.float64)
    return float(np.linalg.norm(a - b) / max(np.linalg.norm(b), 1e-30))


def mats_of(g):
    """mv, proj, inv_mv, inv_proj exactly as the reference's `_C` entry points received them."""
    return [np.ascontiguousarray(m, dtype=np.float32) for m in g["mats"]]


def load(name):
    g = np.load(os.path.join(GOLD, name + ".npz"))
    s = scenes.config(name)
    assert checksum(s) == str(g["checksum"]), "seeded scene generator no longer reproduces the fixture's inputs"
    return s, g


@pytest.mark.parametrize("name", ["tiny_tri", "small_tri"])
def test_tri_oracle_matches_reference(name):
    s, g = load(name)
    o = oracle.TriOracle(s, mats=mats_of(g))
    out = o.outputs()
    # projection feeds float->int conversions: must be bit-identical
    np.testing.assert_array_equal(out["verts_image"].view(np.uint32), g["verts_image"].view(np.uint32))
    np.testing.assert_array_equal(out["ndc_z"].view(np.uint32), g["ndc_z"].view(np.uint32))
    np.testing.assert_array_equal(out["tiles_touched"], g["tiles_touched"])
    np.testing.assert_array_equal(out["offsets"], g["offsets"])
    assert out["R"] == int(g["R"])
    live = g["tiles_touched"] > 0
    np.testing.assert_array_equal(out["depth_keys"][live], g["depth_keys"][live])
    np.testing.assert_array_equal(out["keys_sorted"], g["keys_sorted"])
    np.testing.assert_array_equal(out["values_sorted"], g["values_sorted"])
    np.testing.assert_array_equal(out["ranges"], g["ranges"])
    np.testing.assert_array_equal(out["n_contrib"], g["n_contrib"])
    assert np.abs(out["final_T"] - g["final_T"]).max() <= 1e-6
    assert np.abs(out["color"] - g["color"]).max() <= IMG_TOL
    assert np.abs(out["depth"] - g["depth"]).max() <= IMG_TOL
    gc, gd = scenes.cotangents(s)
    grads = o.backward(gc, gd)
    for k, a in zip(["g_verts", "g_verts_color", "g_faces_opacity", "g_verts_depth", "g_faces_intense"], grads):
        e = rel_l2(a, g[k])
        assert e <= GRAD_TOL, "%s rel L2 %.3e" % (k, e)


@pytest.mark.parametrize("name", ["tiny_tet", "small_tet"])
def test_tet_oracle_matches_reference(name):
    s, g = load(name)
    o = oracle.TetOracle(s, mats=mats_of(g))
    out = o.outputs()
    np.testing.assert_array_equal(out["tiles_touched"], g["tiles_touched"])
    np.testing.assert_array_equal(out["offsets"], g["offsets"])
    assert out["R"] == int(g["R"])
    live = g["tiles_touched"] > 0
    np.testing.assert_array_equal(out["depth_keys"][live], g["depth_keys"][live])
    np.testing.assert_array_equal(out["keys_sorted"], g["keys_sorted"])
    np.testing.assert_array_equal(out["values_sorted"], g["values_sorted"])
    np.testing.assert_array_equal(out["ranges"], g["ranges"])
    np.testing.assert_array_equal(out["first_face"], g["first_face"])
    np.testing.assert_array_equal(out["first_tet"], g["first_tet"])
    np.testing.assert_array_equal(out["n_contrib"], g["n_contrib"])
    np.testing.assert_array_equal(out["active"] > 0.5, g["active"] > 0.5)
    assert np.abs(out["color"] - g["color"]).max() <= IMG_TOL
    assert np.abs(out["depth"] - g["depth"]).max() <= IMG_TOL
    gc, gd = scenes.cotangents(s)
    grads = o.backward(gc, gd)
    for k, a in zip(["g_verts_color", "g_faces_opacity"], grads):
        e = rel_l2(a, g[k])
        assert e <= GRAD_TOL, "%s rel L2 %.3e" % (k, e)


def test_oracle_properties_tri():
    """Size-independent properties of the binning stages on a mid-size scene."""
    s = scenes.random_tri_scene("prop", 5, 5000, 0.06, 160, 208, B=2)
    o = oracle.TriOracle(s).outputs()
    keys = o["keys_sorted"]
    assert np.all(keys[1:] >= keys[:-1])                      # sortedness
    assert o["offsets"][-1] == o["R"] == keys.size            # scan total == instances
    np.testing.assert_array_equal(np.cumsum(o["tiles_touched"], dtype=np.uint64).astype(np.uint32), o["offsets"])
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    counts = np.bincount(tiles, minlength=o["ranges"].shape[0])
    np.testing.assert_array_equal(o["ranges"][:, 1] - o["ranges"][:, 0], counts)   # ranges partition the list
    # multiset of (key,value) preserved by the sort
    a = np.stack([o["keys_unsorted"], o["values_unsorted"].astype(np.uint64)], 1)
    b = np.stack([keys, o["values_sorted"].astype(np.uint64)], 1)
    np.testing.assert_array_equal(a[np.lexsort((a[:, 1], a[:, 0]))], b[np.lexsort((b[:, 1], b[:, 0]))])
    assert o["n_contrib"].max() <= counts.max()
    assert np.all((o["final_T"] >= 0) & (o["final_T"] <= 1))
