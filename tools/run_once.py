"""Dev tool: run fwd+bwd of the CUDA path N times on one config (for ncu launch lists)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dmesh_renderer_b200 import _C, scenes  # noqa: E402

name = sys.argv[1]
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 2
views = int(sys.argv[3]) if len(sys.argv) > 3 else None
s = scenes.to_device(scenes.config(name, views=views) if views else scenes.config(name), "cuda")
gc, gd = [t.cuda() for t in scenes.cotangents(s)]
mv, pj = s.mv_mats.transpose(1, 2).contiguous(), s.proj_mats.transpose(1, 2).contiguous()
imv, ipj = torch.inverse(mv), torch.inverse(pj)
for _ in range(iters):
    if s.kind == "tri":
        a = (s.bg, s.verts, s.faces, s.verts_color, s.faces_opacity, mv, pj, imv, ipj, s.verts_depth, s.faces_intense)
        o = _C.render_tris(*a, s.H, s.W)
        _C.render_tris_backward(*a, gc, gd, o[0], o[3], o[4], o[5], o[6])
    else:
        a = (s.bg, s.verts, s.faces, s.verts_color, s.faces_opacity, mv, pj, imv, ipj, s.verts_depth, s.faces_intense,
             s.tets, s.face_tets, s.tet_faces)
        o = _C.render_tets(*a, s.H, s.W, 0)
        _C.render_tets_backward(*a, gc, gd, o[3], o[4], o[5], o[6], 0)
torch.cuda.synchronize()
print("done", name)
