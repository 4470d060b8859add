"""Test-side harness around the UNMODIFIED reference extension (oracle/_ref).

Calls the four `_C` entry points exactly the way the reference's Python wrapper
does (/root/reference/dmesh_renderer/__init__.py:62-88, 124-149, 298-328,
373-401) and unpacks the reference's opaque state buffers with the layouts of
its fromChunk functions (cuda_rasterizer/rasterizer_impl.cu:127-171,
cuda_renderer/renderer_impl.cu:129-190).

Test infrastructure only -- never imported by dmesh_renderer_b200.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import build_ref  # noqa: E402

_ref = None


def ref_module():
    global _ref
    if _ref is None:
        _ref = build_ref.load()
    return _ref


def _A(x):
    return (x + 127) // 128 * 128


def _view(buf, off, count, dtype):
    nbytes = count * np.dtype(dtype).itemsize
    return buf[off:off + nbytes].cpu().numpy().view(dtype)


def ref_tri_forward(s):
    """s: Scene on cuda.  Returns dict with outputs, num_rendered and the raw buffers."""
    C = ref_module()
    mv = s.mv_mats.transpose(1, 2)
    pj = s.proj_mats.transpose(1, 2)
    imv, ipj = torch.inverse(mv), torch.inverse(pj)
    R, color, depth, pb, fb, bb, ib = C.render_tris(s.bg, s.verts, s.faces, s.verts_color, s.faces_opacity, mv, pj, imv,
                                                    ipj, s.verts_depth, s.faces_intense, s.H, s.W)
    return dict(R=R, color=color, depth=depth, bufs=(pb, fb, bb, ib), mats=(mv, pj, imv, ipj))


def ref_tri_backward(s, fwd, gc, gd):
    C = ref_module()
    mv, pj, imv, ipj = fwd["mats"]
    pb, fb, bb, ib = fwd["bufs"]
    return C.render_tris_backward(s.bg, s.verts, s.faces, s.verts_color, s.faces_opacity, mv, pj, imv, ipj,
                                  s.verts_depth, s.faces_intense, gc, gd, fwd["R"], pb, fb, bb, ib)


def ref_tri_intermediates(s, fwd):
    B, P, F = s.mv_mats.shape[0], s.verts.shape[0], s.faces.shape[0]
    BP, BF, BI = B * P, B * F, B * s.H * s.W
    R = fwd["R"]
    pb, fb, bb, ib = fwd["bufs"]
    out = {}
    ndc = _view(pb, 0, 3 * BP, np.float32).reshape(BP, 3)
    out["ndc_z"] = ndc[:, 2].copy()
    out["verts_image"] = _view(pb, _A(12 * BP), 2 * BP, np.float32).reshape(BP, 2)
    out["depths"] = _view(fb, 0, BF, np.float32)
    out["tiles_touched"] = _view(fb, _A(4 * BF), BF, np.uint32)
    out["offsets"] = _view(fb, fb.numel() - 128 - 4 * BF, BF, np.uint32)
    o = 0
    out["values_sorted"] = _view(bb, o, R, np.uint32); o = _A(o + 4 * R)
    out["values_unsorted"] = _view(bb, o, R, np.uint32); o = _A(o + 4 * R)
    out["keys_sorted"] = _view(bb, o, R, np.uint64); o = _A(o + 8 * R)
    out["keys_unsorted"] = _view(bb, o, R, np.uint64)
    o = 0
    out["final_T"] = _view(ib, o, BI, np.float32); o = _A(o + 4 * BI)
    out["final_prev_T"] = _view(ib, o, BI, np.float32); o = _A(o + 4 * BI)
    out["n_contrib"] = _view(ib, o, BI, np.uint32); o = _A(o + 4 * BI)
    tiles = B * ((s.W + 15) // 16) * ((s.H + 15) // 16)
    out["ranges"] = _view(ib, o, 2 * tiles, np.uint32).reshape(tiles, 2)
    return out


def ref_tet_forward(s, seed=0):
    C = ref_module()
    mv = s.mv_mats.transpose(1, 2)
    pj = s.proj_mats.transpose(1, 2)
    imv, ipj = torch.inverse(mv), torch.inverse(pj)
    color, depth, active, pb, fb, bb, ib = C.render_tets(s.bg, s.verts, s.faces, s.verts_color, s.faces_opacity, mv, pj,
                                                         imv, ipj, s.verts_depth, s.faces_intense, s.tets, s.face_tets,
                                                         s.tet_faces, s.H, s.W, seed)
    return dict(color=color, depth=depth, active=active, bufs=(pb, fb, bb, ib), mats=(mv, pj, imv, ipj))


def ref_tet_backward(s, fwd, gc, gd):
    C = ref_module()
    mv, pj, imv, ipj = fwd["mats"]
    pb, fb, bb, ib = fwd["bufs"]
    return C.render_tets_backward(s.bg, s.verts, s.faces, s.verts_color, s.faces_opacity, mv, pj, imv, ipj,
                                  s.verts_depth, s.faces_intense, s.tets, s.face_tets, s.tet_faces, gc, gd, pb, fb, bb,
                                  ib)


def ref_tet_intermediates(s, fwd):
    B, P, F = s.mv_mats.shape[0], s.verts.shape[0], s.faces.shape[0]
    BP, BF, BI = B * P, B * F, B * s.H * s.W
    pb, fb, bb, ib = fwd["bufs"]
    out = {}
    out["verts_image"] = _view(pb, _A(12 * BP), 2 * BP, np.float32).reshape(BP, 2)
    o = 0
    out["depths"] = _view(fb, o, BF, np.float32); o = _A(o + 4 * BF)
    out["min_depths"] = _view(fb, o, BF, np.float32); o = _A(o + 4 * BF)
    out["max_depths"] = _view(fb, o, BF, np.float32); o = _A(o + 4 * BF)
    out["tiles_touched"] = _view(fb, o, BF, np.uint32)
    out["offsets"] = _view(fb, fb.numel() - 128 - 4 * BF, BF, np.uint32)
    R = int(out["offsets"][-1]) if BF else 0
    out["R"] = R
    o = 0
    out["values_sorted"] = _view(bb, o, R, np.uint32); o = _A(o + 4 * R)
    out["values_unsorted"] = _view(bb, o, R, np.uint32); o = _A(o + 4 * R)
    out["keys_sorted"] = _view(bb, o, R, np.uint64); o = _A(o + 8 * R)
    out["keys_unsorted"] = _view(bb, o, R, np.uint64)
    tiles = B * ((s.W + 15) // 16) * ((s.H + 15) // 16)
    o = 0
    out["n_contrib"] = _view(ib, o, BI, np.uint32); o = _A(o + 4 * BI)
    out["ranges"] = _view(ib, o, 2 * tiles, np.uint32).reshape(tiles, 2); o = _A(o + 8 * BI)
    out["final_log_T"] = _view(ib, o, BI, np.float32); o = _A(o + 4 * BI)
    out["final_prev_log_T"] = _view(ib, o, BI, np.float32); o = _A(o + 4 * BI)
    o = _A(o + 12 * BI)   # ray_o
    o = _A(o + 12 * BI)   # ray_d
    out["first_face"] = _view(ib, o, BI, np.int32); o = _A(o + 4 * BI)
    out["first_tet"] = _view(ib, o, BI, np.int32); o = _A(o + 4 * BI)
    out["last_face"] = _view(ib, o, BI, np.int32); o = _A(o + 4 * BI)
    out["last_tet"] = _view(ib, o, BI, np.int32)
    return out
