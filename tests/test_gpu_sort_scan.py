"""GPU unit tests of the hand-written onesweep radix sort (dmr_sort_pairs) against a
stable CPU sort of the masked keys: bit-exact keys AND values (stability)."""
import ctypes

import numpy as np
import pytest
import torch

from dmesh_renderer_b200 import _lib

pytestmark = pytest.mark.gpu


def gpu_sort(keys, vals, end_bit):
    lib = _lib.load()
    n = keys.size
    dk = torch.from_numpy(keys.view(np.int64)).cuda()
    dv = torch.from_numpy(vals.view(np.int32)).cuda()
    ok = torch.empty_like(dk)
    ov = torch.empty_like(dv)
    temp = torch.empty(lib.dmr_sort_temp_bytes(n), dtype=torch.uint8, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr())
    _lib.check(lib.dmr_sort_pairs(p(dk), p(dv), p(ok), p(ov), n, end_bit, p(temp),
                                  ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    return ok.cpu().numpy().view(np.uint64), ov.cpu().numpy().view(np.uint32), dk.cpu().numpy().view(np.uint64)


def cpu_sort(keys, vals, end_bit):
    mask = np.uint64((1 << end_bit) - 1) if end_bit < 64 else np.uint64(0xFFFFFFFFFFFFFFFF)
    order = np.argsort(keys & mask, kind="stable")
    return keys[order], vals[order]


@pytest.mark.parametrize("n", [1, 2, 31, 257, 2048, 2049, 4096, 4097, 100_003, 1_000_000, 1_048_576, 1_048_577, 5_000_011])
@pytest.mark.parametrize("end_bit", [45, 64, 33])
def test_sort_random(n, end_bit):
    rng = np.random.default_rng(n * 131 + end_bit)
    keys = rng.integers(0, 2 ** 63, size=n, dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=n, dtype=np.uint64)
    if end_bit < 64:
        keys &= np.uint64((1 << end_bit) - 1)   # the renderer's keys have no bits above end_bit
    vals = np.arange(n, dtype=np.uint32)
    gk, gv, kin = gpu_sort(keys, vals, end_bit)
    ck, cv = cpu_sort(keys, vals, end_bit)
    np.testing.assert_array_equal(kin, keys)      # input untouched
    np.testing.assert_array_equal(gk, ck)
    np.testing.assert_array_equal(gv, cv)


def test_sort_renderer_like_keys_skips_constant_digits():
    # tile ids in the high word, depth floats in [0.5,1) in the low word: the top
    # depth byte (0x3f) is constant -> that pass is skipped; many duplicates test stability
    rng = np.random.default_rng(7)
    n = 2_000_003
    tiles = rng.integers(0, 4096, size=n, dtype=np.uint64)
    depth = (rng.random(n, dtype=np.float32) * 0.45 + 0.5).astype(np.float32)
    depth[::7] = depth[0]   # ties
    keys = (tiles << np.uint64(32)) | depth.view(np.uint32).astype(np.uint64)
    vals = rng.integers(0, 200_000, size=n, dtype=np.uint32)
    gk, gv, _ = gpu_sort(keys, vals, 32 + 13)
    ck, cv = cpu_sort(keys, vals, 45)
    np.testing.assert_array_equal(gk, ck)
    np.testing.assert_array_equal(gv, cv)


def test_sort_all_equal_and_sorted_inputs():
    n = 300_000
    vals = np.arange(n, dtype=np.uint32)
    for keys in (np.full(n, 0x3f000000_00000123, dtype=np.uint64) & np.uint64((1 << 45) - 1),
                 np.arange(n, dtype=np.uint64), np.arange(n, dtype=np.uint64)[::-1].copy()):
        gk, gv, _ = gpu_sort(keys, vals, 45)
        ck, cv = cpu_sort(keys, vals, 45)
        np.testing.assert_array_equal(gk, ck)
        np.testing.assert_array_equal(gv, cv)


def gpu_sort_u32(keys, vals, end_bit):
    """dmr_sort_pairs_u32: the (u32 key, u32 value) form the renderers use; vals=None stands for the identity."""
    lib = _lib.load()
    n = keys.size
    dk = torch.from_numpy(keys.view(np.int32)).cuda()
    dv = torch.from_numpy(vals.view(np.int32)).cuda() if vals is not None else None
    ok = torch.empty_like(dk)
    ov = torch.empty_like(dk)
    temp = torch.empty(lib.dmr_sort_temp_bytes(n), dtype=torch.uint8, device="cuda")
    p = lambda t: ctypes.c_void_p(t.data_ptr() if t is not None else 0)
    _lib.check(lib.dmr_sort_pairs_u32(p(dk), p(dv), p(ok), p(ov), n, end_bit, p(temp),
                                      ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    torch.cuda.synchronize()
    return ok.cpu().numpy().view(np.uint32), ov.cpu().numpy().view(np.uint32)


# 2^20 keys is where the onesweep passes switch from 2048-key to 4096-key tiles (csrc/radix_sort.cu: RS_SMALL_N);
# 2047 / 2048 / 2049 and 4095 / 4096 / 4097 straddle one tile of either shape
@pytest.mark.parametrize("n", [2047, 2048, 2049, 4095, 4097, 1_048_575, 1_048_576, 1_048_577, 3_000_001])
@pytest.mark.parametrize("end_bit,identity", [(32, True), (13, False), (17, False)])
def test_sort_u32_both_tile_shapes(n, end_bit, identity):
    rng = np.random.default_rng(n * 7 + end_bit)
    keys = rng.integers(0, 2 ** 32, size=n, dtype=np.uint64).astype(np.uint32)
    if end_bit < 32:
        keys &= np.uint32((1 << end_bit) - 1)
    keys[::5] = keys[0]   # ties: stability
    vals = None if identity else rng.integers(0, 2 ** 31, size=n, dtype=np.uint64).astype(np.uint32)
    gk, gv = gpu_sort_u32(keys, vals, end_bit)
    order = np.argsort(keys, kind="stable")
    np.testing.assert_array_equal(gk, keys[order])
    np.testing.assert_array_equal(gv, order.astype(np.uint32) if identity else vals[order])
