"""Test-only TriRenderer stand-in whose forward/backward run on the CPU oracle.  Lets the
multi-rank logic (camera sharding + packed all-reduce) be exercised on CPU with gloo.
Never imported by the product."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import oracle  # noqa: E402
from dmesh_renderer_b200.scenes import Scene  # noqa: E402


class _OracleTri(torch.autograd.Function):
    @staticmethod
    def forward(ctx, verts, faces, verts_color, faces_opacity, mv, proj, verts_depth, faces_intense, H, W, bg):
        s = Scene("x", "tri", H, W, verts, faces, verts_color, faces_opacity, mv, proj, verts_depth, faces_intense, bg)
        o = oracle.TriOracle(s)
        out = o.outputs()
        ctx.o = o
        return torch.from_numpy(out["color"]), torch.from_numpy(out["depth"])

    @staticmethod
    def backward(ctx, gc, gd):
        g = ctx.o.backward(gc.contiguous(), gd.contiguous())
        f = lambda a: torch.from_numpy(np.asarray(a, np.float32))
        return f(g[0]), None, f(g[1]), f(g[2]), None, None, f(g[3]), f(g[4]), None, None, None


class OracleTriRenderer:
    def __init__(self, H, W, bg):
        self.H, self.W, self.bg = H, W, bg

    def __call__(self, verts, faces, verts_color, faces_opacity, mv_mats, proj_mats, verts_depth, faces_intense):
        return _OracleTri.apply(verts, faces, verts_color, faces_opacity, mv_mats, proj_mats, verts_depth,
                                faces_intense, self.H, self.W, self.bg)
