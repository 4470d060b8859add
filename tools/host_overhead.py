"""Dev tool: where does the step time of the public API go (C2)?"""
import os, sys, time
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dmesh_renderer_b200 import TriRenderer, TriRenderSettings, _C, scenes
from dmesh_renderer_b200.multiview import PackedSceneGrads

s = scenes.to_device(scenes.config("C2"), "cuda")
gc, gd = [t.cuda() for t in scenes.cotangents(s)]
mv, pj = s.mv_mats.transpose(1, 2).contiguous(), s.proj_mats.transpose(1, 2).contiguous()
imv, ipj = torch.inverse(mv), torch.inverse(pj)
a = (s.bg, s.verts, s.faces, s.verts_color, s.faces_opacity, mv, pj, imv, ipj, s.verts_depth, s.faces_intense)
flushbuf = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

def timeit(fn, n=30, flush=False):
    for _ in range(5): fn()
    tot = 0; wall = 0
    for _ in range(n):
        if flush: flushbuf.fill_(1)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter(); e0.record(); fn(); e1.record(); t_host = time.perf_counter() - t0
        torch.cuda.synchronize(); tot += e0.elapsed_time(e1); wall += t_host
    return tot / n, wall / n * 1e3

def direct():
    o = _C.render_tris(*a, s.H, s.W)
    _C.render_tris_backward(*a, gc, gd, o[0], o[3], o[4], o[5], o[6])

leaves = PackedSceneGrads(s.verts.clone(), s.verts_color.clone(), s.faces_opacity.clone())
verts, vcol, fopa = leaves.leaves
vdep = s.verts_depth.clone().requires_grad_(); fint = s.faces_intense.clone().requires_grad_()
r = TriRenderer(TriRenderSettings(s.H, s.W, s.bg))
def api_fixed_cot():
    leaves.zero_(); vdep.grad = None; fint.grad = None
    c, d = r(verts, s.faces, vcol, fopa, s.mv_mats, s.proj_mats, vdep, fint)
    torch.autograd.backward([c, d], [gc, gd])
def api_loss_cot():
    leaves.zero_(); vdep.grad = None; fint.grad = None
    c, d = r(verts, s.faces, vcol, fopa, s.mv_mats, s.proj_mats, vdep, fint)
    torch.autograd.backward([c, d], [c.detach() - gc, d.detach() - gd])
def inverses():
    torch.inverse(mv); torch.inverse(pj)
def fwd_only():
    _C.render_tris(*a, s.H, s.W)

for name, fn in [("_C direct fwd+bwd", direct), ("_C fwd only", fwd_only), ("2x torch.inverse", inverses), ("API fixed cotangent", api_fixed_cot), ("API loss cotangent", api_loss_cot)]:
    g, h = timeit(fn)
    gf, hf = timeit(fn, flush=True)
    print("%-22s gpu %.3f ms (host-side issue %.3f ms) | with L2 flush: gpu %.3f ms" % (name, g, h, gf))

# ---- forward-only / backward-only through the public API
st = {}
def api_fwd():
    st["o"] = r(verts, s.faces, vcol, fopa, s.mv_mats, s.proj_mats, vdep, fint)
def api_bwd():
    leaves.zero_(); vdep.grad = None; fint.grad = None
    torch.autograd.backward(list(st["o"]), [gc, gd], retain_graph=True)
def c_bwd():
    o = st["c"]
    _C.render_tris_backward(*a, gc, gd, o[0], o[3], o[4], o[5], o[6])
st["c"] = _C.render_tris(*a, s.H, s.W)
api_fwd()
for name, fn in [("API fwd only", api_fwd), ("API bwd only", api_bwd), ("_C bwd only", c_bwd)]:
    g, h = timeit(fn)
    print("%-22s gpu %.3f ms (host-side issue %.3f ms)" % (name, g, h))
