"""Dev tool: per-kernel durations (the library's stage events, dmr_profile_*) of fwd and bwd at the
_C level on one or more configs, plus total fwd / bwd / fwd+bwd CUDA-event times.

    python tools/stage_times.py C2 C3 C5 [C4:8]
"""
import ctypes
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dmesh_renderer_b200 import _C, _lib, scenes  # noqa: E402


def med(ts):
    ts = sorted(ts)
    return ts[len(ts) // 2]


def main():
    lib = _lib.load()
    nst = lib.dmr_profile_stage_count()
    names = [lib.dmr_profile_stage_name(i).decode() for i in range(nst)]
    buf = (ctypes.c_float * nst)()
    for name in sys.argv[1:]:
        views = None
        if ":" in name:
            name, views = name.split(":")
            views = int(views)
        s = scenes.to_device(scenes.config(name, views=views) if views else scenes.config(name), "cuda")
        gc, gd = [t.cuda() for t in scenes.cotangents(s)]
        mv, pj = s.mv_mats.transpose(1, 2).contiguous(), s.proj_mats.transpose(1, 2).contiguous()
        imv, ipj = torch.inverse(mv), torch.inverse(pj)
        a = [s.bg, s.verts, s.faces, s.verts_color, s.faces_opacity, mv, pj, imv, ipj, s.verts_depth, s.faces_intense]
        if s.kind == "tet":
            a += [s.tets, s.face_tets, s.tet_faces]
        st = {}

        def fwd():
            st["o"] = _C.render_tris(*a, s.H, s.W) if s.kind == "tri" else _C.render_tets(*a, s.H, s.W, 0)

        def bwd():
            o = st["o"]
            if s.kind == "tri":
                _C.render_tris_backward(*a, gc, gd, o[0], o[3], o[4], o[5], o[6])
            else:
                _C.render_tets_backward(*a, gc, gd, o[3], o[4], o[5], o[6], 0)

        def ev(fn):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize()
            e0.record()
            fn()
            e1.record()
            torch.cuda.synchronize()
            return e0.elapsed_time(e1)

        for _ in range(5):
            fwd(); bwd()
        tf = med([ev(fwd) for _ in range(15)])
        tb = med([ev(bwd) for _ in range(15)])
        tfb = med([ev(lambda: (fwd(), bwd())) for _ in range(15)])
        acc = {}
        lib.dmr_profile_enable(1)
        for _ in range(10):
            for fn in (fwd, bwd):
                torch.cuda.synchronize()
                fn()
                torch.cuda.synchronize()
                lib.dmr_profile_read(buf)
                for i in range(nst):
                    if buf[i] >= 0:
                        acc.setdefault(names[i], []).append(buf[i])
                lib.dmr_profile_enable(1)   # clears the 'used' flags
        lib.dmr_profile_enable(0)
        print("== %s  fwd %.3f  bwd %.3f  fwd+bwd %.3f ms" % (name, tf, tb, tfb))
        tot = sum(med(v) for v in acc.values())
        for k, v in acc.items():
            print("   %-22s %9.1f us  %5.1f%%" % (k, med(v) * 1e3, 100 * med(v) / tot))
        sys.stdout.flush()


if __name__ == "__main__":
    main()
