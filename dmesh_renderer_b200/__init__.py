"""dmesh_renderer_b200 -- B200-native drop-in for the `dmesh_renderer` package.

Public surface = the reference's (/root/reference/dmesh_renderer/__init__.py):
    TriRenderSettings, render_tri, TriRenderer          (:13-225)
    TetRenderSettings, render_tet, TetRenderer          (:237-488)
with identical argument order, dtype handling, return values and gradient
positions.  `import dmesh_renderer_b200 as dmesh_renderer` is the whole
migration.

All compute runs in hand-written sm_100a CUDA kernels behind the C ABI of
libdmesh_b200.so (include/dmesh_b200.h) through the `_C` shim; PyTorch provides
device memory, the current stream and autograd plumbing only.
"""
from typing import NamedTuple

import torch as th

from . import _C

__all__ = ["TriRenderSettings", "render_tri", "TriRenderer", "TetRenderSettings", "render_tet", "TetRenderer"]


# =============================================================================
# TriRenderer: semi-transparent triangles, mean-depth ordered compositing.
# =============================================================================
class TriRenderSettings(NamedTuple):      # reference __init__.py:13-16
    image_height: int
    image_width: int
    bg: th.Tensor


_deterministic_default = False
# TriRenderer launches phase 2 of a forward call before num_rendered has reached the host (see
# _C.tri_forward_finish); DMESH_B200_NO_SPECULATIVE=1 restores wait-then-launch (debugging / A-B timing)
import os as _os
_speculative_phase2 = _os.environ.get("DMESH_B200_NO_SPECULATIVE") != "1"


def set_deterministic(flag: bool) -> bool:
    """Process-wide default for the backward passes of both renderers (extension): True = run-to-run
    reproducible gradients (64-bit fixed-point accumulation instead of fp32 atomics; same values within fp32
    accumulation noise, slower).  `TriRenderer(settings, deterministic=...)` / `render_tri(..., deterministic=...)`
    override it per renderer / call (TetRenderer / render_tet likewise).  Returns the previous value."""
    global _deterministic_default
    old, _deterministic_default = _deterministic_default, bool(flag)
    return old


def render_tri(verts, faces, verts_color, faces_opacity, mv_mats, proj_mats, verts_depth, faces_intense,
               render_settings: TriRenderSettings, deterministic=None):
    """reference __init__.py:18-43 (`deterministic`: see set_deterministic)"""
    det = _deterministic_default if deterministic is None else bool(deterministic)
    return _RenderTri.apply(verts, faces, verts_color, faces_opacity, mv_mats, proj_mats, verts_depth, faces_intense,
                            render_settings, det)


class _RenderTri(th.autograd.Function):
    """reference __init__.py:45-170"""

    @staticmethod
    def forward(ctx, verts, faces, verts_color, faces_opacity, mv_mats, proj_mats, verts_depth, faces_intense,
                render_settings, deterministic=False):
        ctx.deterministic = deterministic
        try:
            # Same computation as _C.render_tris(*13 args) (reference __init__.py:62-88), issued in two halves
            # (phase 1: preprocess + scan; phase 2 after the num_rendered read-back).
            # depth in [-1, 1]: -1 near, 1 far
            inv = _C._Inverses(mv_mats, proj_mats)   # th.inverse x2 (reference :62-63): one launch, no device sync
            inv_mv_mats, inv_proj_mats = inv.inv_mv, inv.inv_proj
            mv_mats, proj_mats = inv.mv, inv.proj    # contiguous copies written by the same launch
            pending = _C.tri_forward_begin(render_settings.bg, verts, faces, verts_color, faces_opacity, mv_mats,
                                           proj_mats, verts_depth, faces_intense, render_settings.image_height,
                                           render_settings.image_width)
            # speculative: phase 2 is enqueued before num_rendered reaches the host; `num_rendered` is then the count
            # the binning buffer was laid out for (>= the real one), which is what backward needs
            num_rendered, color, depth, pointBuffer, faceBuffer, binningBuffer, imgBuffer = \
                _C.tri_forward_finish(pending, inv_mv_mats, inv_proj_mats, inv, speculative=_speculative_phase2)
        except Exception as ex:
            print("\nAn error occured in forward.")
            print(ex)
            raise ex
        ctx.render_settings = render_settings
        ctx.num_rendered = num_rendered
        ctx.fused_depth = verts_depth is None
        ctx.grad_sink = _C.grad_sink_for(verts, verts_color, faces_opacity)   # multiview.PackedSceneGrads.direct()
        ctx.save_for_backward(verts, faces, verts_color, faces_opacity, mv_mats, proj_mats, inv_mv_mats, inv_proj_mats,
                              verts_depth, faces_intense, pointBuffer, faceBuffer, binningBuffer, imgBuffer)
        return color, depth

    @staticmethod
    def backward(ctx, grad_out_color, grad_out_depth):
        num_rendered = ctx.num_rendered
        render_settings = ctx.render_settings
        (verts, faces, verts_color, faces_opacity, mv_mats, proj_mats, inv_mv_mats, inv_proj_mats, verts_depth,
         faces_intense, pointBuffer, faceBuffer, binningBuffer, imgBuffer) = ctx.saved_tensors
        args = (render_settings.bg, verts, faces, verts_color, faces_opacity, mv_mats, proj_mats, inv_mv_mats,
                inv_proj_mats, verts_depth, faces_intense, grad_out_color, grad_out_depth, num_rendered, pointBuffer,
                faceBuffer, binningBuffer, imgBuffer)
        sink = ctx.grad_sink
        into = None
        if sink is not None and sink.bind() == 0:
            into = tuple(leaf.grad for leaf in sink.leaves)
        # (bind() != 0: a leaf's .grad was dropped or replaced between forward and backward -- e.g. zero_grad with
        # set_to_none=True; the sink has re-attached it and this call hands its gradients to autograd the ordinary way)
        try:
            grad_verts, grad_verts_color, grad_faces_opacity, grad_verts_depth, grad_faces_intense = \
                _C.render_tris_backward(*args, accumulate_into=into, deterministic=ctx.deterministic)
        except Exception as ex:
            print("\nAn error occured in backward.\n")
            raise ex
        if ctx.fused_depth:
            # verts_depth=None: the depth was the vertex's own NDC z -> chain its gradient into the vertex positions
            _C.tri_depth_chain(verts, mv_mats, proj_mats, grad_verts_depth, grad_verts)
            grad_verts_depth = None
        if into is not None:
            # already added to the leaves' .grad by the kernels
            grad_verts = grad_verts_color = grad_faces_opacity = None
        # gradient positions: reference __init__.py:156-168
        return (grad_verts, None, grad_verts_color, grad_faces_opacity, None, None, grad_verts_depth,
                grad_faces_intense, None, None)


class TriRenderer(th.nn.Module):
    """reference __init__.py:172-225"""

    def __init__(self, render_settings: TriRenderSettings, deterministic=None):
        super().__init__()
        self.render_settings = render_settings
        self.deterministic = deterministic     # None: the process-wide default (set_deterministic)

    def forward(self, verts, faces, verts_color, faces_opacity, mv_mats, proj_mats, verts_depth, faces_intense):
        """
        verts [P,3] f32, faces [F,3] int, verts_color [P,3], faces_opacity [F]   (view independent)
        mv_mats, proj_mats [B,4,4] in maths convention (transposed here, as the reference does)
        verts_depth [B,P], faces_intense [B,F]                                   (per view)
        verts_depth=None (extension): use each vertex's own NDC z, computed and differentiated inside the
        renderer (what DMesh's callers otherwise compute upstream in PyTorch, a [B,P] tensor per step)
        returns color [B,3,H,W], depth [B,1,H,W]
        """
        return render_tri(verts, faces.to(dtype=th.int32), verts_color, faces_opacity, mv_mats.transpose(1, 2),
                          proj_mats.transpose(1, 2), verts_depth, faces_intense, self.render_settings,
                          self.deterministic)


# =============================================================================
# TetRenderer: faces of a tetrahedral complex, exact depth order by ray marching
# through tet adjacency; gradients to vertex colours and face opacities only.
# =============================================================================
class TetRenderSettings(NamedTuple):      # reference __init__.py:237-241
    image_height: int
    image_width: int
    bg: th.Tensor
    ray_random_seed: int


def render_tet(verts, faces, verts_color, faces_opacity, mv_mats, proj_mats, verts_depth, faces_intense, tets,
               face_tets, tet_faces, render_settings: TetRenderSettings, deterministic=None):
    """reference __init__.py:243-275 (`deterministic`: see set_deterministic)"""
    det = _deterministic_default if deterministic is None else bool(deterministic)
    return _RenderTet.apply(verts, faces, verts_color, faces_opacity, mv_mats, proj_mats, verts_depth, faces_intense,
                            tets, face_tets, tet_faces, render_settings, det)


class _RenderTet(th.autograd.Function):
    """reference __init__.py:277-424"""

    @staticmethod
    def forward(ctx, verts, faces, verts_color, faces_opacity, mv_mats, proj_mats, verts_depth, faces_intense, tets,
                face_tets, tet_faces, render_settings, deterministic=False):
        ctx.deterministic = deterministic
        inv = _C._Inverses(mv_mats, proj_mats)   # th.inverse x2 (reference :298-299): one launch, no device sync
        inv_mv_mats, inv_proj_mats = inv.inv_mv, inv.inv_proj
        mv_mats, proj_mats = inv.mv, inv.proj    # contiguous copies written by the same launch
        args = (render_settings.bg, verts, faces, verts_color, faces_opacity, mv_mats, proj_mats, inv_mv_mats,
                inv_proj_mats, verts_depth, faces_intense, tets, face_tets, tet_faces, render_settings.image_height,
                render_settings.image_width, render_settings.ray_random_seed, inv)
        try:
            color, depth, active, pointBuffer, faceBuffer, binningBuffer, imgBuffer = _C.render_tets(*args)
        except Exception as ex:
            print("\nAn error occured in forward.")
            raise ex
        active = (active > 0.5)
        ctx.render_settings = render_settings
        ctx.tet_records = getattr(faceBuffer, "tet_records", None)   # cached adjacency records (no gradient, not an input)
        ctx.save_for_backward(verts, faces, verts_color, faces_opacity, mv_mats, proj_mats, inv_mv_mats, inv_proj_mats,
                              verts_depth, faces_intense, tets, face_tets, tet_faces, pointBuffer, faceBuffer,
                              binningBuffer, imgBuffer)
        ctx.mark_non_differentiable(active)
        return color, depth, active

    @staticmethod
    def backward(ctx, grad_out_color, grad_out_depth, grad_out_active):
        render_settings = ctx.render_settings
        (verts, faces, verts_color, faces_opacity, mv_mats, proj_mats, inv_mv_mats, inv_proj_mats, verts_depth,
         faces_intense, tets, face_tets, tet_faces, pointBuffer, faceBuffer, binningBuffer, imgBuffer) = ctx.saved_tensors
        args = (render_settings.bg, verts, faces, verts_color, faces_opacity, mv_mats, proj_mats, inv_mv_mats,
                inv_proj_mats, verts_depth, faces_intense, tets, face_tets, tet_faces, grad_out_color, grad_out_depth,
                pointBuffer, faceBuffer, binningBuffer, imgBuffer, render_settings.ray_random_seed)
        try:
            grad_verts_color, grad_faces_opacity = _C.render_tets_backward(*args, deterministic=ctx.deterministic,
                                                                           tet_records=ctx.tet_records)
        except Exception as ex:
            print("\nAn error occured in backward.\n")
            raise ex
        # gradient positions: reference __init__.py:407-422
        return (None, None, grad_verts_color, grad_faces_opacity, None, None, None, None, None, None, None, None, None)


class TetRenderer(th.nn.Module):
    """reference __init__.py:426-488"""

    def __init__(self, render_settings: TetRenderSettings, deterministic=None):
        super().__init__()
        self.render_settings = render_settings
        self.deterministic = deterministic     # None: the process-wide default (set_deterministic)

    def forward(self, verts, faces, verts_color, faces_opacity, mv_mats, proj_mats, verts_depth, faces_intense, tets,
                face_tets, tet_faces):
        """
        Gradients only for verts_color and faces_opacity.  verts_depth is accepted
        but unused (depth comes from the re-projected hit point).
        tets [T,4], face_tets [F,2] (-1 = boundary), tet_faces [T,4]
        returns color [B,3,H,W], depth [B,1,H,W], active [B,H,W] bool
        """
        f32, i32 = th.float32, th.int32
        return render_tet(verts.to(dtype=f32), faces.to(dtype=i32), verts_color.to(dtype=f32),
                          faces_opacity.to(dtype=f32), mv_mats.to(dtype=f32).transpose(1, 2),
                          proj_mats.to(dtype=f32).transpose(1, 2), verts_depth.to(dtype=f32),
                          faces_intense.to(dtype=f32), tets.to(dtype=i32), face_tets.to(dtype=i32),
                          tet_faces.to(dtype=i32), self.render_settings, self.deterministic)
