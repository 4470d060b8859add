"""CPU: the C-ABI library builds, loads, and exports every symbol include/dmesh_b200.h declares
(no compute calls -- there is no GPU here)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    src = open(os.path.join(ROOT, "include", "dmesh_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(dmr_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for s in ["dmr_tri_state_bytes", "dmr_binning_bytes", "dmr_tri_forward_bin", "dmr_tri_forward_render",
              "dmr_tri_backward", "dmr_tet_state_bytes", "dmr_tet_forward_bin", "dmr_tet_forward_render",
              "dmr_tet_backward", "dmr_debug_view", "dmr_sort_pairs", "dmr_sort_temp_bytes", "dmr_last_error",
              "dmr_camera_inverses", "dmr_tri_backward_deterministic", "dmr_tet_backward_deterministic",
              "dmr_tri_depth_chain", "dmr_nvls_allreduce_sum_f32_fused"]:
        assert s in syms


def test_library_builds_and_exports_every_declared_symbol():
    from dmesh_renderer_b200 import build
    lib_path = build.build()
    assert os.path.exists(lib_path)
    lib = ctypes.CDLL(lib_path)
    for s in declared_symbols():
        assert hasattr(lib, s), "libdmesh_b200.so does not export " + s


def test_python_binding_covers_the_header():
    from dmesh_renderer_b200 import _lib
    assert sorted(_lib.SIGNATURES) == declared_symbols()
    lib = _lib.load()
    assert lib.dmr_abi_version() == 2


def test_size_queries_need_no_gpu():
    from dmesh_renderer_b200 import _lib
    lib = _lib.load()
    out = (ctypes.c_size_t * 3)()
    assert lib.dmr_tri_state_bytes(2, 1000, 500, 128, 64, out) == 0
    assert out[0] >= 2 * 1000 * 16 and out[1] >= 2 * 500 * 144 and out[2] >= 2 * 128 * 64 * 12
    assert lib.dmr_tet_state_bytes(1, 100, 300, 120, 64, 64, out) == 0
    assert out[1] >= 120 * 224 + 300 * 64
    assert lib.dmr_binning_bytes(1000) >= 1000 * 36
    assert lib.dmr_sort_temp_bytes(10_000) >= 10_000 * 12
    # negative sizes are rejected with a message, not a crash
    assert lib.dmr_tri_state_bytes(-1, 1, 1, 16, 16, out) != 0
    assert b"size" in lib.dmr_last_error()
    # deterministic-mode workspaces: 24 statistics per (view, face) + per-vertex / per-face accumulators, 8 bytes each
    assert lib.dmr_tri_backward_deterministic_bytes(2, 1000, 500) >= 8 * (24 * 2 * 500 + 8 * 1000 + 2 * 1000 + 500)
    assert lib.dmr_tet_backward_deterministic_bytes(100, 300) >= 8 * (4 * 100 + 300)
    # argument validation happens before any CUDA call
    assert lib.dmr_camera_inverses(0, None, 0, 0, 0, None, 0, 0, 0, None, None, None) == 0          # empty batch
    assert lib.dmr_camera_inverses(-1, None, 0, 0, 0, None, 0, 0, 0, None, None, None) == 1         # DMR_EINVAL
    assert lib.dmr_camera_inverses(2, None, 16, 4, 1, None, 16, 4, 1, None, None, None) == 1
    assert b"null" in lib.dmr_last_error()
    assert lib.dmr_profile_stage_count() >= 16
    names = [lib.dmr_profile_stage_name(i).decode() for i in range(lib.dmr_profile_stage_count())]
    assert "tri_render_backward" in names and "sort_pass0" in names


def test_missing_library_fails_loudly(monkeypatch):
    from dmesh_renderer_b200 import _lib
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", "/nonexistent/libdmesh_b200.so")
    with pytest.raises(ImportError, match="no fallback"):
        _lib.load()
