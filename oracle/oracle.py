"""ctypes wrapper of the CPU oracle (oracle/oracle.cpp).

TEST INFRASTRUCTURE ONLY -- the checker, never the product.  Imported by
tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / reference legs;
nothing under dmesh_renderer_b200/ may import it.
"""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "oracle.cpp")
LIB = os.path.join(HERE, "liboracle.so")

_lib = None


def build(force=False):
    if force or not os.path.exists(LIB) or os.path.getmtime(SRC) > os.path.getmtime(LIB):
        cmd = ["g++", "-O2", "-std=c++17", "-fopenmp", "-ffp-contract=off", "-fPIC", "-shared", SRC, "-o", LIB]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("oracle build failed:\n" + r.stderr[-4000:])
    return LIB


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(LIB) or (os.path.exists(SRC) and os.path.getmtime(SRC) > os.path.getmtime(LIB)):
            build()
        L = ctypes.CDLL(LIB)
        L.oracle_tri_forward.restype = ctypes.c_void_p
        L.oracle_tet_forward.restype = ctypes.c_void_p
        L.oracle_tri_num_rendered.restype = ctypes.c_uint
        L.oracle_tet_num_rendered.restype = ctypes.c_uint
        L.oracle_tri_num_rendered.argtypes = [ctypes.c_void_p]
        L.oracle_tet_num_rendered.argtypes = [ctypes.c_void_p]
        L.oracle_tri_free.argtypes = [ctypes.c_void_p]
        L.oracle_tet_free.argtypes = [ctypes.c_void_p]
        L.oracle_num_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def _np(t, dtype):
    a = t.detach().cpu().numpy() if hasattr(t, "detach") else np.asarray(t)
    return np.ascontiguousarray(a, dtype=dtype)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def _mats(scene):
    """Column-major 16-float matrices + inverses, computed like the reference's
    Python wrapper does (transpose in the module, torch.inverse in the autograd fn)."""
    import torch
    mv = scene.mv_mats.detach().cpu().transpose(1, 2).contiguous()
    pj = scene.proj_mats.detach().cpu().transpose(1, 2).contiguous()
    return [_np(m, np.float32) for m in (mv, pj, torch.inverse(mv), torch.inverse(pj))]


class TriOracle:
    def __init__(self, scene, mats=None):
        L = lib()
        self.B, self.P, self.F = scene.mv_mats.shape[0], scene.verts.shape[0], scene.faces.shape[0]
        self.W, self.H = scene.W, scene.H
        mv, pj, imv, ipj = mats if mats is not None else _mats(scene)
        a = [_np(scene.verts, np.float32), _np(scene.faces, np.int32), _np(scene.verts_color, np.float32),
             _np(scene.faces_opacity, np.float32), mv, pj, imv, ipj, _np(scene.verts_depth, np.float32),
             _np(scene.faces_intense, np.float32), _np(scene.bg, np.float32)]
        self.h = ctypes.c_void_p(L.oracle_tri_forward(self.B, self.P, self.F, self.W, self.H, *[_p(x) for x in a]))
        self.R = int(L.oracle_tri_num_rendered(self.h))

    def outputs(self):
        B, P, F, W, H, R = self.B, self.P, self.F, self.W, self.H, self.R
        tiles = B * ((W + 15) // 16) * ((H + 15) // 16)
        o = dict(color=np.zeros((B, 3, H, W), np.float32), depth=np.zeros((B, 1, H, W), np.float32),
                 verts_image=np.zeros((B * P, 2), np.float32), ndc_z=np.zeros(B * P, np.float32),
                 tiles_touched=np.zeros(B * F, np.uint32), offsets=np.zeros(B * F, np.uint32),
                 depth_keys=np.zeros(B * F, np.uint32), keys_unsorted=np.zeros(R, np.uint64),
                 values_unsorted=np.zeros(R, np.uint32), keys_sorted=np.zeros(R, np.uint64),
                 values_sorted=np.zeros(R, np.uint32), ranges=np.zeros((tiles, 2), np.uint32),
                 n_contrib=np.zeros(B * H * W, np.uint32), final_T=np.zeros(B * H * W, np.float32))
        lib().oracle_tri_get(self.h, *[_p(v) for v in o.values()])
        o["R"] = R
        return o

    def backward(self, dL_dcolor, dL_ddepth):
        B, P, F = self.B, self.P, self.F
        gc, gd = _np(dL_dcolor, np.float32), _np(dL_ddepth, np.float32)
        g = [np.zeros((P, 3)), np.zeros((P, 3)), np.zeros(F), np.zeros((B, P)), np.zeros((B, F))]
        lib().oracle_tri_backward(self.h, _p(gc), _p(gd), *[_p(x) for x in g])
        return g   # float64: verts, verts_color, faces_opacity, verts_depth, faces_intense

    def close(self):
        if self.h:
            lib().oracle_tri_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class TetOracle:
    def __init__(self, scene, mats=None):
        L = lib()
        self.B, self.P, self.F, self.T = scene.mv_mats.shape[0], scene.verts.shape[0], scene.faces.shape[0], scene.tets.shape[0]
        self.W, self.H = scene.W, scene.H
        mv, pj, imv, ipj = mats if mats is not None else _mats(scene)
        a = [_np(scene.verts, np.float32), _np(scene.faces, np.int32), _np(scene.verts_color, np.float32),
             _np(scene.faces_opacity, np.float32), mv, pj, imv, ipj, _np(scene.faces_intense, np.float32),
             _np(scene.tets, np.int32), _np(scene.face_tets, np.int32), _np(scene.tet_faces, np.int32),
             _np(scene.bg, np.float32)]
        self.h = ctypes.c_void_p(L.oracle_tet_forward(self.B, self.P, self.F, self.T, self.W, self.H, *[_p(x) for x in a]))
        self.R = int(L.oracle_tet_num_rendered(self.h))

    def outputs(self):
        B, F, W, H, R = self.B, self.F, self.W, self.H, self.R
        tiles = B * ((W + 15) // 16) * ((H + 15) // 16)
        o = dict(color=np.zeros((B, 3, H, W), np.float32), depth=np.zeros((B, 1, H, W), np.float32),
                 active=np.zeros((B, H, W), np.float32), tiles_touched=np.zeros(B * F, np.uint32),
                 offsets=np.zeros(B * F, np.uint32), depth_keys=np.zeros(B * F, np.uint32),
                 keys_sorted=np.zeros(R, np.uint64), values_sorted=np.zeros(R, np.uint32),
                 ranges=np.zeros((tiles, 2), np.uint32), first_face=np.zeros(B * H * W, np.int32),
                 first_tet=np.zeros(B * H * W, np.int32), n_contrib=np.zeros(B * H * W, np.uint32))
        lib().oracle_tet_get(self.h, *[_p(v) for v in o.values()])
        o["R"] = R
        return o

    def backward(self, dL_dcolor, dL_ddepth):
        gc, gd = _np(dL_dcolor, np.float32), _np(dL_ddepth, np.float32)
        g = [np.zeros((self.P, 3)), np.zeros(self.F)]
        lib().oracle_tet_backward(self.h, _p(gc), _p(gd), *[_p(x) for x in g])
        return g   # float64: verts_color, faces_opacity

    def close(self):
        if self.h:
            lib().oracle_tet_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def num_threads():
    return int(lib().oracle_num_threads())
