"""Dev tool: fwd / bwd / fwd+bwd ms of the CUDA path and of the reference extension on
the same seeded scene (CUDA events, median of N).  usage: time_compare.py C1 C2 ..."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import ref_harness  # noqa: E402
from dmesh_renderer_b200 import _C, scenes  # noqa: E402


def timeit(fn, n=20, warm=5):
    for _ in range(warm):
        fn()
    ts = []
    for _ in range(n):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    names = [a for a in sys.argv[1:] if not a.startswith("--")]
    no_ref = "--no-ref" in sys.argv
    for name in names:
        views = None
        if ":" in name:
            name, views = name.split(":")
            views = int(views)
        s = scenes.to_device(scenes.config(name, views=views) if views else scenes.config(name), "cuda")
        gc, gd = [t.cuda() for t in scenes.cotangents(s)]
        mv, pj = s.mv_mats.transpose(1, 2).contiguous(), s.proj_mats.transpose(1, 2).contiguous()
        imv, ipj = torch.inverse(mv), torch.inverse(pj)
        ref = None if no_ref else ref_harness.ref_module()
        if s.kind == "tri":
            fargs = (s.bg, s.verts, s.faces, s.verts_color, s.faces_opacity, mv, pj, imv, ipj, s.verts_depth,
                     s.faces_intense, s.H, s.W)
            for label, C in (("ours", _C), ("ref", ref)):
                if C is None:
                    continue
                st = {}

                def fwd():
                    st["o"] = C.render_tris(*fargs)

                def bwd():
                    o = st["o"]
                    C.render_tris_backward(*fargs[:11], gc, gd, o[0], o[3], o[4], o[5], o[6])

                def both():
                    fwd()
                    bwd()
                f = timeit(fwd)
                bw = timeit(bwd)
                fb = timeit(both)
                print("%s %-5s R=%d fwd %.3f (min %.3f)  bwd %.3f (min %.3f)  fwd+bwd %.3f (min %.3f) ms" %
                      (name, label, st["o"][0], f[0], f[1], bw[0], bw[1], fb[0], fb[1]), flush=True)
        else:
            fargs = (s.bg, s.verts, s.faces, s.verts_color, s.faces_opacity, mv, pj, imv, ipj, s.verts_depth,
                     s.faces_intense, s.tets, s.face_tets, s.tet_faces, s.H, s.W, 0)
            for label, C in (("ours", _C), ("ref", ref)):
                if C is None:
                    continue
                st = {}

                def fwd():
                    st["o"] = C.render_tets(*fargs)

                def bwd():
                    o = st["o"]
                    C.render_tets_backward(*fargs[:14], gc, gd, o[3], o[4], o[5], o[6])

                def both():
                    fwd()
                    bwd()
                f = timeit(fwd)
                bw = timeit(bwd)
                fb = timeit(both)
                print("%s %-5s fwd %.3f (min %.3f)  bwd %.3f (min %.3f)  fwd+bwd %.3f (min %.3f) ms" %
                      (name, label, f[0], f[1], bw[0], bw[1], fb[0], fb[1]), flush=True)


if __name__ == "__main__":
    main()
