import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch, ref_harness
from dmesh_renderer_b200 import scenes, TetRenderer, TetRenderSettings, _lib
lib = _lib.load()
def rel_l2(a, b): return ((a - b).norm() / b.norm().clamp_min(1e-30)).item()
for name in ["tiny_tet", "small_tet", "C3"]:
    s = scenes.to_device(scenes.config(name), "cuda")
    gc, gd = [t.cuda() for t in scenes.cotangents(s)]
    ref = ref_harness.ref_tet_forward(s, 0)
    rg = ref_harness.ref_tet_backward(s, ref, gc, gd)
    for cap in (0, 1, 7):
        lib.dmr_debug_set_tet_trail_cap(cap)
        vc = s.verts_color.clone().requires_grad_(); fo = s.faces_opacity.clone().requires_grad_()
        r = TetRenderer(TetRenderSettings(s.H, s.W, s.bg, 0))
        color, depth, active = r(s.verts, s.faces, vc, fo, s.mv_mats, s.proj_mats, s.verts_depth, s.faces_intense, s.tets, s.face_tets, s.tet_faces)
        torch.autograd.backward([color, depth], [gc, gd])
        print(name, "cap", cap, "B", s.mv_mats.shape[0], "vc", rel_l2(vc.grad, rg[0]), "fo", rel_l2(fo.grad, rg[1]), "maxabs fo", (fo.grad-rg[1]).abs().max().item(), "nbad", ((fo.grad-rg[1]).abs()>1e-5).sum().item())
    lib.dmr_debug_set_tet_trail_cap(0)
