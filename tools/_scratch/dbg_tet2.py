import sys, os
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import torch, ref_harness
from dmesh_renderer_b200 import scenes, TetRenderer, TetRenderSettings, _lib
lib = _lib.load()
s = scenes.to_device(scenes.config("C3"), "cuda")
gc, gd = [t.cuda() for t in scenes.cotangents(s)]
ref = ref_harness.ref_tet_forward(s, 0)
rg = ref_harness.ref_tet_backward(s, ref, gc, gd)
res = {}
for cap in (0, 1):
    lib.dmr_debug_set_tet_trail_cap(cap)
    vc = s.verts_color.clone().requires_grad_(); fo = s.faces_opacity.clone().requires_grad_()
    r = TetRenderer(TetRenderSettings(s.H, s.W, s.bg, 0))
    color, depth, active = r(s.verts, s.faces, vc, fo, s.mv_mats, s.proj_mats, s.verts_depth, s.faces_intense, s.tets, s.face_tets, s.tet_faces)
    torch.autograd.backward([color, depth], [gc, gd])
    res[cap] = fo.grad.clone()
lib.dmr_debug_set_tet_trail_cap(0)
print("|ref| max", rg[1].abs().max().item(), "mean", rg[1].abs().mean().item())
for a, b, n in ((res[0], rg[1], "trail-ref"), (res[1], rg[1], "remarch-ref"), (res[0], res[1], "trail-remarch")):
    d = (a - b).abs()
    top = torch.topk(d, 8)
    print(n, "top diffs:")
    for v, i in zip(top.values.tolist(), top.indices.tolist()):
        print("   face %d diff %.3e  a %.6e b %.6e rel %.2e" % (i, v, a[i].item(), b[i].item(), v / max(abs(b[i].item()), 1e-30)))
    print("   hist of rel diff (|b|>1e-3):", torch.histc(torch.log10((d / b.abs().clamp_min(1e-3)).clamp_min(1e-12)), bins=12, min=-12, max=0).tolist())
