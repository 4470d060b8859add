"""Dev tool: time dmr_sort_pairs (stage events) on renderer-like keys.  usage: bench_sort.py N [end_bit]"""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from dmesh_renderer_b200 import _lib  # noqa: E402

n = int(float(sys.argv[1])) if len(sys.argv) > 1 else 32_000_000
tiles = 16384 if n > 4_000_000 else 4096
end_bit = int(sys.argv[2]) if len(sys.argv) > 2 else 32 + tiles.bit_length()
lib = _lib.load()
g = torch.Generator(device="cuda").manual_seed(1)
tile = torch.randint(0, tiles, (n,), device="cuda", generator=g, dtype=torch.int64)
depth = (torch.rand(n, device="cuda", generator=g) * 0.24 + 0.73).view(torch.int32).to(torch.int64)
keys = (tile << 32) | depth
vals = torch.randint(0, 1 << 22, (n,), device="cuda", generator=g, dtype=torch.int32)
ok, ov = torch.empty_like(keys), torch.empty_like(vals)
temp = torch.empty(lib.dmr_sort_temp_bytes(n), dtype=torch.uint8, device="cuda")
P = lambda t: ctypes.c_void_p(t.data_ptr())
st = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
nst = lib.dmr_profile_stage_count()
names = [lib.dmr_profile_stage_name(i).decode() for i in range(nst)]
buf = (ctypes.c_float * nst)()
lib.dmr_profile_enable(1)
acc = {}
for it in range(8):
    _lib.check(lib.dmr_sort_pairs(P(keys), P(vals), P(ok), P(ov), n, end_bit, P(temp), st))
    torch.cuda.synchronize()
    lib.dmr_profile_read(buf)
    if it >= 3:
        for i in range(nst):
            if buf[i] >= 0:
                acc.setdefault(names[i], []).append(buf[i])
assert bool((ok[1:] >= ok[:-1]).all())
tot = sum(sum(v) / len(v) for v in acc.values())
passes = [sum(v) / len(v) for k, v in acc.items() if k.startswith("sort_pass")]
execd = [p for p in passes if p > 0.02 * max(passes)]
gbps = 24 * n / (sum(execd) / len(execd) * 1e-3) / 1e9
print("cfg=%s n=%d end_bit=%d total %.3f ms | hist %.3f | passes %s | per executed pass %.3f ms = %.0f GB/s (%.1f%% of 6554)" %
      (os.environ.get("DMR_SORT_CFG", "0"), n, end_bit, tot, sum(acc["sort_histogram"]) / len(acc["sort_histogram"]),
       " ".join("%.3f" % p for p in passes), sum(execd) / len(execd), gbps, 100 * gbps / 6553.9))
