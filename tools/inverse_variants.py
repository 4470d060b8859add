"""Dev tool: find the fp32 operation order that reproduces torch.inverse on 4x4 stacks bit for bit (tools/inverse_variants.cu)."""
import ctypes, math, os, sys
import torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dmesh_renderer_b200 import scenes
lib = ctypes.CDLL(os.path.join(ROOT, "tools", "_bin", "libinv.so"))
dev = "cuda"
g = torch.Generator().manual_seed(1)

def cameras(n):
    # random look-at model-view matrices and perspective projections, row-vector convention (transposed) as the API passes them
    eye = torch.randn(n, 3, generator=g) * 2.5
    f = -eye / eye.norm(dim=1, keepdim=True)
    up = torch.tensor([0.0, 1.0, 0.0]).expand(n, 3) + 0.1 * torch.randn(n, 3, generator=g)
    s = torch.cross(f, up, dim=1); s = s / s.norm(dim=1, keepdim=True)
    u = torch.cross(s, f, dim=1)
    mv = torch.zeros(n, 4, 4); mv[:, 0, :3] = s; mv[:, 1, :3] = u; mv[:, 2, :3] = -f; mv[:, 3, 3] = 1
    mv[:, 0, 3] = -(s * eye).sum(1); mv[:, 1, 3] = -(u * eye).sum(1); mv[:, 2, 3] = (f * eye).sum(1)
    fov = 0.4 + torch.rand(n, generator=g); near = 0.01 + torch.rand(n, generator=g) * 0.5; far = 10 + 100 * torch.rand(n, generator=g)
    t = 1.0 / torch.tan(fov / 2)
    pj = torch.zeros(n, 4, 4); pj[:, 0, 0] = t; pj[:, 1, 1] = t; pj[:, 2, 2] = -(far + near) / (far - near)
    pj[:, 2, 3] = -2 * far * near / (far - near); pj[:, 3, 2] = -1
    return mv.transpose(1, 2).contiguous(), pj.transpose(1, 2).contiguous(), mv.contiguous(), pj.contiguous()

sets = {"randn": torch.randn(20000, 4, 4, generator=g)}
mvT, pjT, mv, pj = cameras(20000)
sets.update({"mvT": mvT, "pjT": pjT, "mv": mv, "pj": pj})
for name in ("C1", "C2", "C4"):
    s = scenes.config(name)
    sets[name + "_mvT"] = s.mv_mats.transpose(1, 2).contiguous()
    sets[name + "_pjT"] = s.proj_mats.transpose(1, 2).contiguous()

def ours(x, variant):
    out = torch.empty_like(x); info = torch.empty(x.shape[0], dtype=torch.int32, device=dev)
    rc = lib.inv4_run(ctypes.c_void_p(x.data_ptr()), ctypes.c_void_p(out.data_ptr()), ctypes.c_void_p(info.data_ptr()),
                      x.shape[0], variant, ctypes.c_void_p(torch.cuda.current_stream().cuda_stream))
    assert rc == 0
    return out

def same(a, b):
    return (a.view(torch.int32) == b.view(torch.int32)).flatten(1).all(1)

best = {}
for name, x in sets.items():
    x = x.to(dev)
    ref_b = torch.linalg.inv_ex(x)[0]
    k = min(200, x.shape[0])
    ref_1 = torch.stack([torch.linalg.inv_ex(x[i:i + 1])[0][0] for i in range(k)])
    ref_2 = torch.cat([torch.linalg.inv_ex(x[i:i + 2])[0] for i in range(0, k - k % 2, 2)]) if k >= 2 else ref_1
    print("%-8s n=%d  torch batched==single: %.4f  pairs==single: %.4f" % (name, x.shape[0], same(ref_b[:k], ref_1).float().mean().item(),
          same(ref_2, ref_1[:ref_2.shape[0]]).float().mean().item()))
    rows = []
    for v in range(64):
        o = ours(x, v)
        rows.append((same(o, ref_b).float().mean().item(), same(o[:k], ref_1).float().mean().item(), v))
        best.setdefault(v, []).append((rows[-1][0], rows[-1][1]))
    rows.sort(reverse=True)
    print("   top vs batched:", ["v%d:%.4f" % (v, a) for a, b, v in rows[:5]])
    rows.sort(key=lambda r: -r[1])
    print("   top vs single: ", ["v%d:%.4f" % (v, b) for a, b, v in rows[:5]])
    o = ours(x, rows[0][2])
    err = ((o[:k] - ref_1).abs().max() / ref_1.abs().max()).item()
    print("   max rel err of best-vs-single variant: %.3g" % err)
print("overall (min over sets) batched:", sorted(((min(a for a, b in r), v) for v, r in best.items()), reverse=True)[:5])
print("overall (min over sets) single: ", sorted(((min(b for a, b in r), v) for v, r in best.items()), reverse=True)[:5])
