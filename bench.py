#!/usr/bin/env python
"""bench.py -- headline benchmark of the DMesh tile rasterizer hot path on B200.

    python bench.py --gpus N --steps K --warmup W [--impl reference] [--workload C2|C1|C4|C5|C3]

Metric (BASELINE.json): fwd+bwd throughput in views/s (and ms per view) of the
tri renderer.  A "step" = one optimisation step of the hot path: forward +
backward of this rank's view(s) of the seeded synthetic scene through the PUBLIC
API (TriRenderer -> autograd -> _C shim -> C ABI -> sm_100a kernels), followed,
when N > 1, by the single all-reduce of the packed scene gradients
(dmesh_renderer_b200/multiview.py).

Default workload = BASELINE.json configs[1] ("C2": 200k triangles, 1024x1024,
one view per rank and step; weak scaling: rank r renders camera r of the same
replicated scene).  `value` is measured with the step inputs resident in HBM,
`e2e` with the per-view inputs (cameras, verts_depth, faces_intense, target
images) in pinned HOST memory, copied H2D inside the timed region, and the scalar
loss read back D2H.

The reference arm (--impl reference) runs the UNMODIFIED reference CUDA extension
(oracle/_ref, built from /root/reference by oracle/build_ref.py) on the same
workload with the same protocol.  The reference has no CPU implementation; the
CPU oracle (oracle/oracle.cpp) is timed as `cpu_baseline` on our arm's line and is
the fallback of the reference arm when oracle/_ref cannot be loaded.
"""
import argparse
import ctypes
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

METRIC = "tri_fwd_bwd_views_per_sec_1024x1024"
UNIT = "views/s"


# --------------------------------------------------------------------------- helpers
def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([x.strip() for x in line.split(",")])

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(sm)}


class L2Flush:
    """Write a buffer larger than the 126 MB L2 between timed steps."""

    def __init__(self, device, mbytes=256):
        self.buf = torch.empty(mbytes * 1024 * 1024, dtype=torch.uint8, device=device)
        self.v = 0

    def __call__(self):
        self.v = (self.v + 1) & 0xff
        self.buf.fill_(self.v)


CPU_BINDING = None   # cores this rank was bound to (N > 1 only)


def dist_setup(gpus):
    ws = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if ws > 1:
        torch.cuda.set_device(local)
        if os.environ.get("DMR_BENCH_NO_CPU_BINDING") != "1":
            from dmesh_renderer_b200.multiview import bind_to_gpu_cpus
            cores = bind_to_gpu_cpus(local)     # pinned staging buffers on the GPU's own NUMA node
            global CPU_BINDING
            CPU_BINDING = len(cores) if cores else None
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    else:
        torch.cuda.set_device(0)
    return ws, rank, local


def barrier(ws):
    if ws > 1:
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, ws, dev):
    if ws == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


# --------------------------------------------------------------------------- workload
def make_workload(name, rank, ws, dev):
    """Scene replicated on every rank; rank r owns its camera(s).  Returns device scene + pinned host copies of the
    per-step inputs."""
    from dmesh_renderer_b200 import scenes
    if name in ("C1", "C2", "C5"):
        base = scenes.config(name)
        views_per_rank = 1
        if ws > 1:   # weak scaling: camera r of a seeded ring around the same scene
            g = torch.Generator().manual_seed(99)
            dirs = scenes.fibonacci_dirs(max(ws, 2), g)
            mv, pj = scenes.cameras(dirs[rank:rank + 1], 3.0, base.W, base.H, 0.5, 6.0)
            base = base._replace(mv_mats=mv, proj_mats=pj, verts_depth=scenes.ndc_depth(base.verts, mv, pj))
        total_views = ws
    elif name == "C4":   # 64 views sharded over the ranks (strong scaling)
        from dmesh_renderer_b200.multiview import shard_views
        full = scenes.config("C4", views=64)
        mine = shard_views(64, rank, ws)
        sl = slice(mine.start, mine.stop)
        base = full._replace(mv_mats=full.mv_mats[sl].contiguous(), proj_mats=full.proj_mats[sl].contiguous(),
                             verts_depth=full.verts_depth[sl].contiguous(), faces_intense=full.faces_intense[sl].contiguous())
        views_per_rank = len(mine)
        total_views = 64
    else:
        raise SystemExit("unknown workload " + name)
    s = scenes.to_device(base, dev)
    gen = torch.Generator().manual_seed(1234 + rank)
    B = base.mv_mats.shape[0]
    host = dict(mv=base.mv_mats, proj=base.proj_mats, verts_depth=base.verts_depth, faces_intense=base.faces_intense,
                target_color=torch.rand(B, 3, base.H, base.W, generator=gen),
                target_depth=torch.rand(B, 1, base.H, base.W, generator=gen))
    host = {k: v.contiguous().pin_memory() for k, v in host.items()}
    return s, host, views_per_rank, total_views


def describe(name, s, ws, views_per_rank, total_views):
    return {
        "workload": {"C1": "configs[0]", "C2": "configs[1]", "C4": "configs[3]", "C5": "configs[4]"}[name] +
        ": tri renderer fwd+bwd, %d triangles, %dx%d, %d view(s) per rank and step" % (s.faces.shape[0], s.W, s.H, views_per_rank),
        "triangles": int(s.faces.shape[0]), "vertices": int(s.verts.shape[0]), "image": [s.H, s.W],
        "views_per_step_total": total_views, "parallelism": "camera-sharded x%d, 1 all-reduce of (6P+F) fp32 (NVLS multimem kernel over symmetric memory, NCCL fallback)" % ws if ws > 1 else "single GPU",
        "l2_flush": "256 MB device fill between timed steps, outside the timed events",
    }


# --------------------------------------------------------------------------- our arm
def run_ours(args, ws, rank, local):
    from dmesh_renderer_b200 import TriRenderer, TriRenderSettings, _lib
    from dmesh_renderer_b200.multiview import PackedSceneGrads
    dev = torch.device("cuda", local)
    s, host, vpr, total_views = make_workload(args.workload, rank, ws, dev)
    lib = _lib.load()
    renderer = TriRenderer(TriRenderSettings(s.H, s.W, s.bg))
    leaves = PackedSceneGrads(s.verts.clone(), s.verts_color.clone(), s.faces_opacity.clone())
    verts, vcol, fopa = leaves.leaves
    vdep = s.verts_depth.clone().requires_grad_()
    fint = s.faces_intense.clone().requires_grad_()
    tgt_c, tgt_d = host["target_color"].to(dev), host["target_depth"].to(dev)
    flush = L2Flush(dev)

    def step_device():
        leaves.zero_()
        vdep.grad = None
        fint.grad = None
        with leaves.direct():   # backward kernels accumulate straight into the packed gradient buffer
            color, depth = renderer(verts, s.faces, vcol, fopa, s.mv_mats, s.proj_mats, vdep, fint)
            # image loss gradient as cotangent (device resident targets)
            torch.autograd.backward([color, depth], [color.detach() - tgt_c, depth.detach() - tgt_d])
        leaves.all_reduce()

    # the per-view render inputs (cameras, verts_depth, faces_intense) live in ONE pinned staging buffer and travel
    # as one copy; the device tensors handed to the renderer are views of its device twin
    rkeys = ("mv", "proj", "verts_depth", "faces_intense")
    n_in = sum(host[k].numel() for k in rkeys)
    stage_host = torch.empty(n_in, dtype=torch.float32).pin_memory()
    stage_dev = torch.empty(n_in, dtype=torch.float32, device=dev)
    dbuf, o = {}, 0
    for k in rkeys:
        n = host[k].numel()
        stage_host[o:o + n].copy_(host[k].reshape(-1))
        dbuf[k] = stage_dev[o:o + n].view(host[k].shape)
        o += n
    for k in ("target_color", "target_depth"):
        dbuf[k] = torch.empty_like(host[k], device=dev)
    n_c, n_d = host["target_color"].numel(), host["target_depth"].numel()
    diff = torch.empty(n_c + n_d, dtype=torch.float32, device=dev)     # [color - target | depth - target]
    diff_c, diff_d = diff[:n_c].view(host["target_color"].shape), diff[n_c:].view(host["target_depth"].shape)
    loss_host = torch.zeros(1).pin_memory()
    h2d_bytes = sum(v.numel() * v.element_size() for v in host.values())

    copy_stream = torch.cuda.Stream(device=dev)

    def step_e2e():
        # H2D of this step's inputs from pinned host memory, inside the timed region.  The per-view render inputs
        # (cameras, verts_depth, faces_intense: 3.2 MB) are submitted first, on the compute stream; the target
        # images (16.8 MB, ~0.31 ms of PCIe time), which only the loss needs, follow on a copy stream and travel
        # while the forward pass runs (the H2D engine serves copies in submission order).
        main = torch.cuda.current_stream()
        stage_dev.copy_(stage_host, non_blocking=True)
        copy_stream.wait_stream(main)           # the previous step's readers of the target buffers are done
        with torch.cuda.stream(copy_stream):
            for k in ("target_color", "target_depth"):
                dbuf[k].copy_(host[k], non_blocking=True)
            targets_ready = torch.cuda.Event()
            targets_ready.record(copy_stream)
        leaves.zero_()
        vd = dbuf["verts_depth"].requires_grad_()
        fi = dbuf["faces_intense"].requires_grad_()
        with leaves.direct():
            color, depth = renderer(verts, s.faces, vcol, fopa, dbuf["mv"], dbuf["proj"], vd, fi)
            main.wait_event(targets_ready)
            # image loss 0.5 * ||render - target||^2; its gradient is the cotangent
            torch.sub(color.detach(), dbuf["target_color"], out=diff_c)
            torch.sub(depth.detach(), dbuf["target_depth"], out=diff_d)
            torch.autograd.backward([color, depth], [diff_c, diff_d])
            loss = torch.linalg.vector_norm(diff).square() * 0.5
        leaves.all_reduce()
        loss_host.copy_(loss.reshape(1), non_blocking=True)
        torch.cuda.current_stream().synchronize()
        dbuf["verts_depth"].requires_grad_(False)
        dbuf["faces_intense"].requires_grad_(False)
        return float(loss_host[0])

    # ---- warm-up
    for _ in range(max(args.warmup, 3)):
        step_device()
    barrier(ws)

    # ---- timed region 1: device-resident inputs, CUDA events per step, L2 flushed between steps
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    lib.dmr_launch_count.restype = ctypes.c_ulonglong
    launches0 = lib.dmr_launch_count()
    total_ms = 0.0
    for _ in range(args.steps):
        flush()
        barrier(ws)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step_device()
        b.record()
        barrier(ws)
        total_ms += a.elapsed_time(b)
    launches = int(lib.dmr_launch_count() - launches0)
    total_ms = max_over_ranks(total_ms, ws, dev)
    clocks = sampler.stop() if rank == 0 else None
    ms_per_step = total_ms / args.steps
    value = total_views / (ms_per_step / 1e3)

    # ---- timed region 2: per-kernel durations (stage events inside the library) for the roofline
    nst = lib.dmr_profile_stage_count()
    names = [lib.dmr_profile_stage_name(i).decode() for i in range(nst)]
    acc = [0.0] * nst
    cnt = [0] * nst
    lib.dmr_profile_enable(1)
    buf = (ctypes.c_float * nst)()
    for _ in range(args.steps):
        flush()
        torch.cuda.synchronize()
        # forward and backward are read separately: a stage's events are overwritten by the next call
        leaves.zero_()
        color, depth = renderer(verts, s.faces, vcol, fopa, s.mv_mats, s.proj_mats, vdep, fint)
        lib.dmr_profile_read(buf)
        for i in range(nst):
            if buf[i] >= 0:
                acc[i] += buf[i]; cnt[i] += 1
        torch.autograd.backward([color, depth], [color.detach() - tgt_c, depth.detach() - tgt_d])
        lib.dmr_profile_read(buf)
        for i in range(nst):
            if buf[i] >= 0:
                acc[i] += buf[i]; cnt[i] += 1
    lib.dmr_profile_enable(0)
    stage_ms = {names[i]: acc[i] / cnt[i] for i in range(nst) if cnt[i]}
    R = int(renderer_last_R(s, renderer))

    # ---- timed region 3: end to end (host inputs, H2D + D2H inside), wall clock
    for _ in range(2):
        step_e2e()
    e2e_s = 0.0
    for _ in range(args.steps):
        flush()
        barrier(ws)
        t0 = time.perf_counter()
        step_e2e()
        torch.cuda.synchronize()
        e2e_s += time.perf_counter() - t0
    e2e_s = max_over_ranks(e2e_s, ws, dev)
    e2e_value = total_views / (e2e_s / args.steps)

    if rank != 0:
        return
    # ---- roofline of the dominant kernel
    P, F, px = int(s.verts.shape[0]), int(s.faces.shape[0]), s.H * s.W
    tiles = vpr * ((s.W + 15) // 16) * ((s.H + 15) // 16)
    npass = (bit_length(tiles) + 7) // 8          # instance sort: tile bits only (two-level binning)
    BF = F * vpr
    alg = {   # algorithmic bytes per launch (SURVEY.md 8d / DESIGN.md 3), per rank (vpr views)
        "preprocess_points": 32 * P * vpr, "preprocess_faces": (12 + 48 + 72 + 8 + 16 + 144) * BF,
        "face_depth_sort": (4 + 4 * 16) * BF,      # histogram read + 4 eight-bit passes over (u32 key, u32 index) pairs
        "scan": 12 * BF,                           # order + gathered tiles_touched read, offsets written
        "duplicate_with_keys": 16 * BF + 8 * R,    # order, offsets, rect read; (u32 tile, u32 face) written
        "sort_histogram": 4 * R,
        "tile_ranges": 4 * R + 8 * tiles,
        "tri_render_forward": 132 * R + 28 * px * vpr,
        "tri_render_backward": 132 * R + 28 * px * vpr + 4 * (6 * P + F) + 4 * (P + F) * vpr,
        "tri_grad_finish": (96 + 144) * BF + 4 * (6 * P + F) + 4 * (P + F) * vpr,
    }
    for i in range(8):
        alg["sort_pass%d" % i] = 16 * R
    dom = max(stage_ms, key=stage_ms.get)
    peak, peak_src = peaks()
    achieved = alg.get(dom, 0) / (stage_ms[dom] * 1e-3) / 1e9
    traffic, winst = None, None
    tp = os.path.join(ROOT, "profiles", "traffic.json")
    if os.path.exists(tp):
        with open(tp) as f:
            tj = json.load(f)
        traffic = tj.get(args.workload, {}).get(dom)
        winst = tj.get(args.workload + "_warp_instructions", {}).get(dom)
    roofline = {"kernel": dom, "bound": "hbm", "achieved": round(achieved, 2), "peak": peak, "unit": "GB/s",
                "frac": round(achieved / peak, 4), "traffic": traffic, "peak_source": peak_src,
                "kernel_ms": round(stage_ms[dom], 4), "algorithmic_bytes": alg.get(dom, 0),
                "note": "render kernels are issue-bound (coverage tests + shading), not HBM-bound; see DESIGN.md",
                "stage_ms": {k: round(v, 4) for k, v in stage_ms.items()}, "instances_R": R, "sort_passes": npass}
    if winst:
        # the bound that actually limits the dominant kernel: warp-instruction issue (148 SMs x 4 schedulers x SM clock);
        # instruction count from the committed ncu capture, duration measured live above
        sm_mhz = (clocks or {}).get("sm_mhz") or 1965.0
        peak_issue = 148 * 4 * sm_mhz * 1e6
        roofline["issue_bound"] = {"warp_instructions": winst, "achieved_ginst_per_s": round(winst / (stage_ms[dom] * 1e-3) / 1e9, 1),
                                   "peak_ginst_per_s": round(peak_issue / 1e9, 1),
                                   "frac": round(winst / (stage_ms[dom] * 1e-3) / peak_issue, 3)}

    # ---- CPU baseline: the oracle port on the host cores, bounded sample
    cpu = cpu_baseline(args.workload) if ws == 1 else None

    out = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": ws, "steps": args.steps,
           "warmup": max(args.warmup, 3), "ms_per_step": round(ms_per_step, 4), "higher_is_better": True,
           "scaling": "strong" if args.workload == "C4" else "weak", "vs_baseline": None, "dtype": "f32",
           "data": "synthetic (seeded, SURVEY.md App. E)", "config": describe(args.workload, s, ws, vpr, total_views),
           "ms_per_view": round(ms_per_step * ws / total_views if args.workload != "C4" else ms_per_step / vpr, 4),
           "e2e": {"value": round(e2e_value, 3), "unit": UNIT, "h2d_bytes_per_step": int(h2d_bytes),
                   "d2h_bytes_per_step": 4, "ms_per_step": round(e2e_s / args.steps * 1e3, 4)},
           "gpu_launches": launches, "clocks": clocks, "roofline": roofline, "impl": "dmesh_renderer_b200"}
    if ws > 1:
        out["config"]["host_cores_bound_to_gpu_numa_node"] = CPU_BINDING
    if cpu:
        out["cpu_baseline"] = cpu
    print(json.dumps(out), flush=True)


def bit_length(n):
    return max(int(n).bit_length(), 1)


def renderer_last_R(s, renderer):
    from dmesh_renderer_b200 import _C
    mv, pj = s.mv_mats.transpose(1, 2), s.proj_mats.transpose(1, 2)
    return _C.render_tris(s.bg, s.verts, s.faces, s.verts_color, s.faces_opacity, mv, pj, torch.inverse(mv),
                          torch.inverse(pj), s.verts_depth, s.faces_intense, s.H, s.W)[0]


def cpu_baseline(workload):
    """The CPU oracle (oracle/oracle.cpp, OpenMP) on the host cores: fwd+bwd of one view of the workload."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    try:
        import oracle
        from dmesh_renderer_b200 import scenes
        name = workload if workload in ("C1", "C2") else "C2"
        s = scenes.config(name)
        gc, gd = scenes.cotangents(s)
        t0 = time.perf_counter()
        o = oracle.TriOracle(s)
        o.backward(gc, gd)
        dt = time.perf_counter() - t0
        o.close()
        return {"value": round(1.0 / dt, 4), "unit": UNIT, "cores": oracle.num_threads(), "kind": "port",
                "sample": "1 full fwd+bwd view of %s (%d triangles, %dx%d), %.2f s" % (name, s.faces.shape[0], s.W, s.H, dt)}
    except Exception as ex:   # the oracle is a reported baseline only
        return {"value": None, "unit": UNIT, "cores": 0, "kind": "port", "sample": "unavailable: %r" % (ex,)}


# --------------------------------------------------------------------------- reference arm
def run_reference(args, ws, rank, local):
    if rank != 0:
        return   # rank 0 alone runs the reference arm
    dev = torch.device("cuda", 0)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_ref
    s, host, vpr, total_views = make_workload(args.workload, 0, 1, dev)
    ref = None
    try:
        ref = build_ref.load()
    except Exception as ex:
        sys.stderr.write("reference extension failed to load: %r\n" % (ex,))
    cfg = describe(args.workload, s, 1, vpr, total_views)
    if ref is None:
        cpu = cpu_baseline(args.workload)
        out = {"metric": METRIC, "value": cpu["value"], "unit": UNIT, "n_gpus": 1, "steps": 1, "warmup": 0,
               "ms_per_step": round(1e3 / cpu["value"], 3) if cpu["value"] else None, "higher_is_better": True,
               "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
               "impl": "reference", "cpu_baseline": cpu,
               "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
               "note": "oracle/_ref (the reference CUDA extension) could not be loaded; CPU oracle port timed instead"}
        print(json.dumps(out), flush=True)
        return
    mv, pj = s.mv_mats.transpose(1, 2).contiguous(), s.proj_mats.transpose(1, 2).contiguous()
    tgt_c, tgt_d = host["target_color"].to(dev), host["target_depth"].to(dev)
    flush = L2Flush(dev)

    def step():
        # exactly what the reference's Python wrapper does (reference __init__.py:62-88, 124-149)
        imv, ipj = torch.inverse(mv), torch.inverse(pj)
        a = (s.bg, s.verts, s.faces, s.verts_color, s.faces_opacity, mv, pj, imv, ipj, s.verts_depth, s.faces_intense)
        R, color, depth, pb, fb, bb, ib = ref.render_tris(*a, s.H, s.W)
        ref.render_tris_backward(*a, color - tgt_c, depth - tgt_d, R, pb, fb, bb, ib)

    for _ in range(max(args.warmup, 3)):
        step()
    torch.cuda.synchronize()
    sampler = ClockSampler(0)
    sampler.start()
    total_ms = 0.0
    for _ in range(args.steps):
        flush()
        torch.cuda.synchronize()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        step()
        b.record()
        torch.cuda.synchronize()
        total_ms += a.elapsed_time(b)
    clocks = sampler.stop()
    ms = total_ms / args.steps
    value = vpr / (ms / 1e3)
    out = {"metric": METRIC, "value": round(value, 3), "unit": UNIT, "n_gpus": ws, "ranks_used": 1, "steps": args.steps,
           "warmup": max(args.warmup, 3), "ms_per_step": round(ms, 4), "higher_is_better": True, "scaling": "weak",
           "vs_baseline": None, "dtype": "f32", "data": "synthetic (seeded, SURVEY.md App. E)", "config": cfg,
           "impl": "reference", "clocks": clocks,
           "note": "the reference has no multi-GPU path: rank 0 alone runs it on one B200 (launched with %d rank(s))" % ws,
           "cpu_baseline": {"value": round(value, 3), "unit": UNIT, "cores": 1, "kind": "reference",
                            "sample": "the reference has no CPU path: its unmodified CUDA extension (oracle/_ref) ran the full "
                                      "workload on one B200, driven by 1 host thread"},
           "e2e": {"value": round(value, 3), "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(out), flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="C2", choices=["C1", "C2", "C4", "C5"])
    args = ap.parse_args()
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (B200); there is no CPU fallback")
    ws, rank, local = dist_setup(args.gpus)
    try:
        if args.impl == "reference":
            run_reference(args, ws, rank, local)
        else:
            run_ours(args, ws, rank, local)
    finally:
        if ws > 1 and dist.is_initialized():
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
