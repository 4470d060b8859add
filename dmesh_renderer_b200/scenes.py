"""Seeded synthetic scenes of BASELINE.json's configs.

Everything is generated on the CPU with a seeded torch.Generator so that the
reference extension, the CPU oracle and the CUDA path see identical bits.
Matrices are returned in maths convention ([B,4,4], row-major, column vectors);
`TriRenderer` / `TetRenderer` transpose them like the reference does.
"""
import math
from typing import NamedTuple, Optional

import torch


class Scene(NamedTuple):
    name: str
    kind: str                    # "tri" | "tet"
    H: int
    W: int
    verts: torch.Tensor          # [P,3] f32
    faces: torch.Tensor          # [F,3] i32
    verts_color: torch.Tensor    # [P,3]
    faces_opacity: torch.Tensor  # [F]
    mv_mats: torch.Tensor        # [B,4,4]
    proj_mats: torch.Tensor      # [B,4,4]
    verts_depth: torch.Tensor    # [B,P]
    faces_intense: torch.Tensor  # [B,F]
    bg: torch.Tensor             # [3]
    tets: Optional[torch.Tensor] = None       # [T,4]
    face_tets: Optional[torch.Tensor] = None  # [F,2]
    tet_faces: Optional[torch.Tensor] = None  # [T,4]


def perspective(fov_y_deg, aspect, near, far):
    f = 1.0 / math.tan(math.radians(fov_y_deg) / 2)
    m = torch.zeros(4, 4, dtype=torch.float64)
    m[0, 0] = f / aspect
    m[1, 1] = f
    m[2, 2] = (far + near) / (near - far)
    m[2, 3] = 2 * far * near / (near - far)
    m[3, 2] = -1.0
    return m


def look_at(eye, target=(0.0, 0.0, 0.0), up=(0.0, 1.0, 0.0)):
    eye = torch.tensor(eye, dtype=torch.float64)
    target = torch.tensor(target, dtype=torch.float64)
    up = torch.tensor(up, dtype=torch.float64)
    fwd = target - eye
    fwd = fwd / fwd.norm()
    right = torch.linalg.cross(fwd, up)
    right = right / right.norm()
    u = torch.linalg.cross(right, fwd)
    m = torch.eye(4, dtype=torch.float64)
    m[0, :3], m[1, :3], m[2, :3] = right, u, -fwd
    m[0, 3], m[1, 3], m[2, 3] = -right.dot(eye), -u.dot(eye), fwd.dot(eye)
    return m


def fibonacci_dirs(n, gen):
    """n quasi-uniform directions on the sphere with a seeded rotation offset."""
    off = torch.rand(1, generator=gen, dtype=torch.float64).item()
    i = torch.arange(n, dtype=torch.float64) + 0.5
    z = 1 - 2 * i / n
    r = torch.sqrt(torch.clamp(1 - z * z, min=0))
    phi = (i * math.pi * (3 - math.sqrt(5)) + 2 * math.pi * off)
    d = torch.stack([r * torch.cos(phi), z * 0.8 + 0.0, r * torch.sin(phi)], dim=1)   # keep |y| < 1 for a stable up vector
    return d / d.norm(dim=1, keepdim=True)


def ndc_depth(verts, mv, proj):
    """verts_depth[b,p] = NDC z of vertex p in view b (torch, f32): the natural
    upstream definition used by DMesh (App. E)."""
    P = verts.shape[0]
    vh = torch.cat([verts, torch.ones(P, 1, dtype=verts.dtype)], dim=1)          # [P,4]
    clip = torch.einsum("bij,bjk,pk->bpi", proj, mv, vh)                          # [B,P,4]
    return (clip[..., 2] / clip[..., 3]).contiguous()


def cameras(dirs, dist, W, H, near, far):
    mv = torch.stack([look_at((dist * d).tolist()) for d in dirs]).to(torch.float32)
    pj = perspective(45.0, W / H, near, far).to(torch.float32).unsqueeze(0).repeat(len(dirs), 1, 1)
    return mv.contiguous(), pj.contiguous()


DEFAULT_DIR = torch.tensor([[0.3, 0.2, 1.0]], dtype=torch.float64) / math.sqrt(0.3 ** 2 + 0.2 ** 2 + 1.0)


def random_tri_scene(name, seed, F, sigma, H, W, B=1, opacity=(0.1, 0.7), dist=3.0, near=0.5, far=6.0, view_slice=None):
    """`view_slice`: keep only these views of the B-view scene (same random stream, so view b is the same camera and
    the same intensities whether or not the other views are kept; the per-view tensors of the dropped views --
    verts_depth is [B,P] -- are never built)."""
    g = torch.Generator().manual_seed(seed)
    centres = torch.rand(F, 1, 3, generator=g) * 1.8 - 0.9
    verts = (centres + sigma * torch.randn(F, 3, 3, generator=g)).reshape(3 * F, 3).contiguous()
    faces = torch.arange(3 * F, dtype=torch.int32).reshape(F, 3)
    verts_color = torch.rand(3 * F, 3, generator=g)
    faces_opacity = torch.rand(F, generator=g) * (opacity[1] - opacity[0]) + opacity[0]
    faces_intense = torch.rand(B, F, generator=g) * 0.5 + 0.5
    dirs = DEFAULT_DIR if B == 1 else fibonacci_dirs(B, g)
    if view_slice is not None:
        dirs, faces_intense = dirs[view_slice], faces_intense[view_slice].contiguous()
    mv, pj = cameras(dirs, dist, W, H, near, far)
    verts_depth = ndc_depth(verts, mv, pj)
    bg = torch.ones(3)
    return Scene(name, "tri", H, W, verts, faces, verts_color, faces_opacity, mv, pj, verts_depth, faces_intense, bg)


def kuhn_tet_grid(n, seed, jitter=0.1):
    """n^3 cells on [-1,1]^3, 6 Kuhn tets per cell along the (0,0,0)-(1,1,1)
    diagonal (faces conform across cells), seeded jitter on interior vertices.
    Returns verts [P,3], tets [T,4], faces [F,3], face_tets [F,2], tet_faces [T,4]."""
    g = torch.Generator().manual_seed(seed)
    m = n + 1
    idx = torch.arange(m)
    I, J, K = torch.meshgrid(idx, idx, idx, indexing="ij")
    grid = torch.stack([I, J, K], dim=-1).reshape(-1, 3)
    verts = grid.to(torch.float32) * (2.0 / n) - 1.0
    interior = ((grid > 0) & (grid < n)).all(dim=1)
    jit = (torch.rand(verts.shape, generator=g) * 2 - 1) * jitter * (2.0 / n)
    verts = verts + jit * interior.unsqueeze(1)

    def vid(i, j, k):
        return (i * m + j) * m + k

    c = torch.arange(n)
    CI, CJ, CK = torch.meshgrid(c, c, c, indexing="ij")
    CI, CJ, CK = CI.reshape(-1), CJ.reshape(-1), CK.reshape(-1)
    perms = [(0, 1, 2), (0, 2, 1), (1, 0, 2), (1, 2, 0), (2, 0, 1), (2, 1, 0)]
    tets = []
    for perm in perms:
        pts = [torch.stack([CI, CJ, CK], dim=1)]
        cur = pts[0].clone()
        for ax in perm:
            cur = cur.clone()
            cur[:, ax] += 1
            pts.append(cur)
        tets.append(torch.stack([vid(p[:, 0], p[:, 1], p[:, 2]) for p in pts], dim=1))
    tets = torch.stack(tets, dim=1).reshape(-1, 4)          # [T,4], cell-major
    T = tets.shape[0]
    # faces: the 4 vertex triples of each tet, made unique
    combos = torch.tensor([[0, 1, 2], [0, 1, 3], [0, 2, 3], [1, 2, 3]])
    tri = tets[:, combos]                                     # [T,4,3]
    tri_sorted, _ = tri.sort(dim=2)
    flat = tri_sorted.reshape(-1, 3)
    Pn = verts.shape[0]
    key = (flat[:, 0] * Pn + flat[:, 1]) * Pn + flat[:, 2]
    ukey, inv = torch.unique(key, return_inverse=True)
    F = ukey.shape[0]
    faces = torch.stack([ukey // (Pn * Pn), (ukey // Pn) % Pn, ukey % Pn], dim=1)
    tet_faces = inv.reshape(T, 4)
    face_tets = torch.full((F, 2), -1, dtype=torch.long)
    tet_id = torch.arange(T).repeat_interleave(4)
    order = torch.argsort(inv, stable=True)
    inv_s, tet_s = inv[order], tet_id[order]
    first = torch.ones_like(inv_s, dtype=torch.bool)
    first[1:] = inv_s[1:] != inv_s[:-1]
    face_tets[inv_s[first], 0] = tet_s[first]
    face_tets[inv_s[~first], 1] = tet_s[~first]
    return (verts.contiguous(), tets.to(torch.int32).contiguous(), faces.to(torch.int32).contiguous(),
            face_tets.to(torch.int32).contiguous(), tet_faces.to(torch.int32).contiguous())


def tet_grid_scene(name, seed, n, H, W, B=1, opacity=(0.0, 0.1), dist=4.0, near=0.5, far=8.0):
    g = torch.Generator().manual_seed(seed + 1000)
    verts, tets, faces, face_tets, tet_faces = kuhn_tet_grid(n, seed)
    P, F = verts.shape[0], faces.shape[0]
    verts_color = torch.rand(P, 3, generator=g)
    faces_opacity = torch.rand(F, generator=g) * (opacity[1] - opacity[0]) + opacity[0]
    faces_intense = torch.ones(B, F)
    dirs = DEFAULT_DIR if B == 1 else fibonacci_dirs(B, g)
    mv, pj = cameras(dirs, dist, W, H, near, far)
    verts_depth = ndc_depth(verts, mv, pj)
    bg = torch.ones(3)
    return Scene(name, "tet", H, W, verts, faces, verts_color, faces_opacity, mv, pj, verts_depth, faces_intense, bg,
                 tets, face_tets, tet_faces)


# ---------------------------------------------------------------------------
# BASELINE.json configs
# ---------------------------------------------------------------------------
def config(name, views=None, view_slice=None):
    """C1..C5 of SURVEY.md section 8.  `views` overrides the number of cameras of C4; `view_slice` keeps a rank's
    share of them (a rank renders 64/G views of the 64-view scene)."""
    if name == "C1":
        return random_tri_scene("C1", 0, 10_000, 0.08, 256, 256)
    if name == "C2":
        return random_tri_scene("C2", 1, 200_000, 0.02, 1024, 1024)
    if name == "C3":
        return tet_grid_scene("C3", 2, 64, 512, 512)
    if name == "C4":
        return random_tri_scene("C4", 3, 1_000_000, 0.01, 1024, 1024, B=views or 64, view_slice=view_slice)
    if name == "C5":
        return random_tri_scene("C5", 4, 4_000_000, 0.02, 2048, 2048, opacity=(0.3, 0.9))
    # small variants for tests
    if name == "tiny_tri":
        return random_tri_scene("tiny_tri", 10, 300, 0.15, 64, 64, B=2)
    if name == "small_tri":
        return random_tri_scene("small_tri", 11, 3000, 0.08, 128, 160, B=2)
    if name == "tiny_tet":
        return tet_grid_scene("tiny_tet", 12, 4, 64, 64, B=2, opacity=(0.0, 0.5))
    if name == "small_tet":
        return tet_grid_scene("small_tet", 13, 12, 128, 128, B=1, opacity=(0.0, 0.2))
    raise KeyError(name)


def to_device(scene: Scene, device):
    return Scene(*[v.to(device) if isinstance(v, torch.Tensor) else v for v in scene])


def cotangents(scene: Scene, seed=1234):
    """dL_dcolor, dL_ddepth ~ U(-1,1), seeded (not a scalar loss)."""
    g = torch.Generator().manual_seed(seed)
    B = scene.mv_mats.shape[0]
    return (torch.rand(B, 3, scene.H, scene.W, generator=g) * 2 - 1, torch.rand(B, 1, scene.H, scene.W, generator=g) * 2 - 1)
