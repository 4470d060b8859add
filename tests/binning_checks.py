"""Checks shared by the GPU parity tests for the two-level binning (csrc/common.cuh: bin_faces /
bin_instances).  The native path emits instances in DEPTH order of the faces instead of the reference's
face-major order, so

  * `offsets` is the inclusive scan of tiles_touched in that order (same total R),
  * the unsorted (key, value) list is a permutation of the reference's,
  * the SORTED lists must be bit-identical to the reference's (checked by the callers).
"""
import numpy as np


def check_face_order_and_offsets(order, depth_keys, tiles_touched, offsets, ref_offsets):
    n = order.size
    assert n == tiles_touched.size == offsets.size == ref_offsets.size
    if n == 0:
        return
    np.testing.assert_array_equal(np.sort(order), np.arange(n, dtype=np.uint32))          # a permutation
    d = depth_keys[order]
    assert np.all(d[1:] >= d[:-1])                                                          # sorted by depth key
    same = d[1:] == d[:-1]
    assert np.all(order[1:][same] > order[:-1][same])                                       # stable
    np.testing.assert_array_equal(np.cumsum(tiles_touched[order], dtype=np.uint64).astype(np.uint32), offsets)
    assert int(offsets[-1]) == int(ref_offsets[-1])                                         # R


def check_same_pairs(keys, values, ref_keys, ref_values):
    """(key, value) multisets agree."""
    assert keys.size == ref_keys.size
    a = np.lexsort((values, keys))
    b = np.lexsort((ref_values, ref_keys))
    np.testing.assert_array_equal(keys[a], ref_keys[b])
    np.testing.assert_array_equal(values[a], ref_values[b])
