"""Typed views into the opaque state buffers (dmr_debug_view) for parity checks."""
import ctypes

import numpy as np
import torch

from . import _lib

KINDS = {
    "verts_image": (0, np.float32, 4),
    "tiles_touched": (1, np.uint32, 1),
    "offsets": (2, np.uint32, 1),
    "depth_keys": (3, np.uint32, 1),
    "keys_unsorted": (4, np.uint64, 1),
    "values_unsorted": (5, np.uint32, 1),
    "keys_sorted": (6, np.uint64, 1),
    "values_sorted": (7, np.uint32, 1),
    "ranges": (8, np.uint32, 2),
    "n_contrib": (9, np.uint32, 1),
    "final_T": (10, np.float32, 1),
    "first_face": (11, np.int32, 1),
    "first_tet": (12, np.int32, 1),
}


def view(renderer, kind, buffer, B, P, F, W, H, R=0, T=0):
    """Return a numpy copy of one intermediate.  renderer: "tri" | "tet"."""
    lib = _lib.load()
    k, dtype, width = KINDS[kind]
    ptr = ctypes.c_void_p()
    cnt = ctypes.c_size_t()
    _lib.check(lib.dmr_debug_view(0 if renderer == "tri" else 1, k, B, P, F, T, W, H, R,
                                  ctypes.c_void_p(buffer.data_ptr()), ctypes.byref(ptr), ctypes.byref(cnt)))
    n = cnt.value
    if n == 0:
        return np.zeros((0, width) if width > 1 else (0,), dtype=dtype)
    off = ptr.value - buffer.data_ptr()
    nbytes = n * width * np.dtype(dtype).itemsize
    arr = buffer[off:off + nbytes].cpu().numpy().view(dtype)
    return arr.reshape(n, width) if width > 1 else arr
