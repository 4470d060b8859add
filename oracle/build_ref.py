"""Build the UNMODIFIED reference CUDA extension into oracle/_ref/ (test infrastructure).

TEST INFRASTRUCTURE ONLY.  Nothing under dmesh_renderer_b200/ may import or
link anything produced here.  Only tests/, __graft_entry__.smoke() and
bench.py (reference arm / cpu_baseline leg) may load it.

The reference (SonSang/dmesh_renderer, /root/reference) ships only a CUDA
implementation.  This recipe compiles its eight sources *where they lie*
(nothing is copied into the repo) with plain nvcc invocations -- the
reference's own setup.py is not run -- and writes a single pybind module
``oracle/_ref/dmesh_ref_C<EXT_SUFFIX>`` that exposes the four entry points of
/root/reference/ext.cpp:4-12 (render_tris, render_tris_backward, render_tets,
render_tets_backward).

Two workarounds, neither touching reference sources:
  * /root/reference/setup.py:25 points at an un-vendored third_party/glm/;
    no glm:: symbol is used anywhere, so a 3-line stub header is generated
    under oracle/_ref/stub/glm/glm.hpp.
  * */rasterizer_impl.h, renderer_impl.h use uint32_t/std::uintptr_t without
    <cstdint> (fails on gcc 13) -> ``-include cstdint``.

oracle/_ref/ is git-ignored but NOT gpurun-ignored, so the .so travels to the
GPU box; /root/reference itself does not exist there.
"""
import os
import subprocess
import sys
import sysconfig
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "_ref")
REF = os.environ.get("DMESH_REFERENCE_DIR", "/root/reference")
MODNAME = "dmesh_ref_C"

SOURCES = [
    "cuda_rasterizer/rasterizer_impl.cu",
    "cuda_rasterizer/forward.cu",
    "cuda_rasterizer/backward.cu",
    "cuda_renderer/renderer_impl.cu",
    "cuda_renderer/forward.cu",
    "cuda_renderer/backward.cu",
    "render.cu",
    "ext.cpp",
]


def so_path():
    return os.path.join(OUT, MODNAME + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force=False, verbose=True, relink=False):
    """Compile the reference extension.  Returns the .so path, or None when
    /root/reference is absent (GPU box: the prebuilt file is used)."""
    target = so_path()
    if not os.path.isdir(REF):
        return target if os.path.exists(target) else None
    if os.path.exists(target) and not force and not relink:
        return target
    import torch  # noqa: F401  (header locations only)
    from torch.utils import cpp_extension as ce

    os.makedirs(os.path.join(OUT, "stub", "glm"), exist_ok=True)
    os.makedirs(os.path.join(OUT, "obj"), exist_ok=True)
    with open(os.path.join(OUT, "stub", "glm", "glm.hpp"), "w") as f:
        f.write("#pragma once\n#include <cstdint>\n#include <cstddef>\n")

    inc = ["-I" + p for p in ce.include_paths("cuda")]
    inc += ["-I" + sysconfig.get_paths()["include"], "-I" + os.path.join(OUT, "stub"), "-I" + REF]
    common = [
        "-DTORCH_EXTENSION_NAME=" + MODNAME,
        "-DTORCH_API_INCLUDE_EXTENSION_H",
        "-D_GLIBCXX_USE_CXX11_ABI=%d" % int(torch._C._GLIBCXX_USE_CXX11_ABI),
        "-std=c++17",
        # nvcc defaults, exactly like the reference's own setup.py build (SURVEY
        # App. A.10): device code -O3 / fmad=true / IEEE div+sqrt / no fast-math,
        # HOST code unoptimised.  The latter matters: Renderer::forward
        # (cuda_renderer/renderer_impl.cu:193,410) is declared int and has no
        # return statement; with host -O2/-O3 gcc 13 treats the end of the
        # function as unreachable and falls through into the CHECK_CUDA throw
        # ("RuntimeError: no error").  At nvcc's default host -O0, gcc 13 turns the
        # same spot into a trap (-funreachable-traps is on by default at -O0) and
        # the process dies with SIGILL; -fno-unreachable-traps restores the
        # benign fall-off-the-end behaviour older compilers gave the reference.
        "-Xcompiler", "-fno-unreachable-traps",
        "-gencode", "arch=compute_100,code=sm_100",
        "-include", "cstdint",
        "--expt-relaxed-constexpr",
        "-Xcompiler", "-fPIC",
        "-w",
    ]
    objs = []
    cmds = []
    for s in SOURCES:
        o = os.path.join(OUT, "obj", s.replace("/", "_") + ".o")
        objs.append(o)
        src = os.path.join(REF, s)
        extra = ["-x", "cu"] if s.endswith(".cpp") else []
        if force or not os.path.exists(o):
            cmds.append(["nvcc", "-c", *extra, src, "-o", o, *common, *inc])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("reference build failed: %s\n%s" % (" ".join(cmd), r.stderr[-4000:]))
        if verbose:
            print("[build_ref] compiled", [c for c in cmd if c.endswith((".cu", ".cpp"))][0])

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        list(ex.map(run, cmds))

    libdirs = ce.library_paths("cuda")
    link = ["nvcc", "-shared", *objs, "-o", target]
    for d in libdirs:
        link += ["-L" + d, "-Xlinker", "-rpath=" + d]
    link += ["-lc10", "-ltorch_cpu", "-ltorch", "-ltorch_python", "-lc10_cuda", "-ltorch_cuda", "-lcudart"]
    r = subprocess.run(link, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("reference link failed:\n" + r.stderr[-4000:])
    if verbose:
        print("[build_ref] linked", target)
    return target


def load():
    """Import the built reference module (needs torch imported first).
    Returns the module or None when it has not been built."""
    p = so_path()
    if not os.path.exists(p):
        return None
    import importlib.util
    import torch  # noqa: F401
    spec = importlib.util.spec_from_file_location(MODNAME, p)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == "__main__":
    p = build(force="--force" in sys.argv, relink="--relink" in sys.argv)
    print(p)
