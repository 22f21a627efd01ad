#!/usr/bin/env python3
"""Benchmark of the extract_mesh hot path (BASELINE.json metric / configs[1]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

A *step* is one TSR.extract_mesh over one synthetic scene code per GPU: 256^3 lattice,
3x40x64x64 triplane with a baked analytic field (SURVEY 8d family B), random-init decoder,
threshold = median density (random-init density never reaches the default 25.0).  With N
GPUs every rank meshes its own scene code (independent objects, no data-path collective):
weak scaling.  `--mode sharded` instead splits ONE 512^3 lattice into x-slabs over the ranks
(BASELINE configs[2], strong scaling; NCCL only to gather the slab meshes).

One JSON line on rank 0.  value = lattice points / s through the whole path (query + MLP +
marching cubes) with the triplane and decoder already in HBM; e2e = the same through the
C-ABI host-buffer call (host triplane in, host mesh out, copies inside the timed region).
"""
from __future__ import annotations

import argparse
import ctypes
import json
import os
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

RADIUS = 0.87
FLOP_PER_POINT = 2 * (120 * 64 + 8 * 64 * 64 + 64 * 4)  # 81408: the reference's NeRFMLP (all 4 outputs), SURVEY 8d
KERNELS_PER_STEP = 5  # project_planes, lattice_tc_ta_kernel (also ballots the MC sign masks), mc_count, mc_totals, mc_emit


def baked_triplane(seed: int, H: int = 64, W: int = 64, noise: float = 0.05) -> torch.Tensor:
    g = torch.Generator().manual_seed(seed)
    A = torch.randn(3, 40, 1, 1, generator=g)
    u = ((torch.arange(W) + 0.5) / W * 2 - 1).view(1, 1, 1, W)
    v = ((torch.arange(H) + 0.5) / H * 2 - 1).view(1, 1, H, 1)
    return (A * (u * u + v * v) / 2 + noise * torch.randn(3, 40, H, W, generator=g)).float()


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            d = json.load(f)
        return d, "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0}, "fallback"


class ClockSampler:
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index: int):
        self.proc = None
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-lms", "100", "-i", str(index)],
                stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True,
            )
        except Exception:  # noqa: BLE001
            self.proc = None

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            out, _ = self.proc.communicate(timeout=5)
        except Exception:  # noqa: BLE001
            self.proc.kill()
            out = ""
        sm, mx, pw, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            parts = [x.strip() for x in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0]))
                mx.append(float(parts[1]))
                pw.append(float(parts[2]))
            except ValueError:
                continue
            for n, v in zip(names, parts[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        # "under load" = samples in the upper half of the power range seen
        if sm:
            thr = (max(pw) + min(pw)) / 2
            load = [s for s, p in zip(sm, pw) if p >= thr] or sm
            return {"sm_mhz": float(np.median(load)), "sm_max_mhz": max(mx), "power_w_max": max(pw), "samples": len(sm), "reasons": sorted(reasons)}
        return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}


def gpu_index(local_rank: int) -> int:
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        try:
            return int(vis.split(",")[local_rank])
        except (ValueError, IndexError):
            pass
    return local_rank


# ------------------------------------------------------------------ CPU arm
def decoder_numpy(seed: int = 0):
    from sculptmate_b200.tsr import NeRFMLP

    torch.manual_seed(seed)
    dec = NeRFMLP(dict(in_channels=120, n_neurons=64, n_hidden_layers=9, activation="silu"))
    sd = dec.state_dict()
    ws = [sd[f"layers.{i}.weight"].numpy().copy() for i in range(0, 20, 2)]
    bs = [sd[f"layers.{i}.bias"].numpy().copy() for i in range(0, 20, 2)]
    return dec, ws, bs


def cpu_reference_sample(R: int, target_s: float, threshold=None):
    """The reference's CPU extract_mesh path (PORT, oracle/cpu_reference_port.py) timed on a
    bounded x-slab of the same R^3 job, all host threads.  Returns (pts/s, description, dict)."""
    from oracle import cpu_reference_port as port

    torch.set_num_threads(os.cpu_count() or 1)
    _, ws, bs = decoder_numpy(0)
    layers = port.make_layers(ws, bs)
    tp = baked_triplane(100)
    # calibrate on 2 planes, then size the sample for ~target_s (at least 3 planes, at most R)
    t0 = time.perf_counter()
    _, _, tm = port.extract_mesh_slab(layers, tp, R, 0.5, R // 2, 2)
    rate = tm["points"] / (time.perf_counter() - t0)
    nx = int(max(3, min(R, round(rate * target_s / (R * R)))))
    x0 = (R - nx) // 2
    if threshold is None:
        q = port.query_triplane(layers, port._scale(torch.rand(20000, 3) * 2 - 1, (-1, 1), (-RADIUS, RADIUS)), tp)
        threshold = float(q["density_act"].median())
    t0 = time.perf_counter()
    v, f, tm = port.extract_mesh_slab(layers, tp, R, threshold, x0, nx)
    dt = time.perf_counter() - t0
    desc = (f"{nx} of {R} x-planes of the {R}^3 lattice ({int(tm['points'])} points): aten grid_sample + Linear/SiLU chain, "
            f"chunk 8192, fp32 ({tm['query_s']:.2f} s) + oracle MC, not skimage ({tm['mc_s']:.2f} s); V={len(v)} F={len(f)}")
    return tm["points"] / dt, desc, {"seconds": dt, "points": tm["points"], "threads": torch.get_num_threads()}


def run_reference_arm(args, rank: int):
    if rank != 0:
        return 0
    R = args.resolution
    for _ in range(args.warmup):
        cpu_reference_sample(R, 0.5)
    vals, secs, desc, info = [], 0.0, "", {}
    per_step = max(1.0, min(20.0, 60.0 / max(1, args.steps)))
    for _ in range(args.steps):
        v, desc, info = cpu_reference_sample(R, per_step)
        vals.append(v)
        secs += info["seconds"]
    value = float(np.mean(vals))
    line = {
        "impl": "reference", "metric": "extract_mesh_lattice_points_per_s", "value": value, "unit": "pts/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * secs / max(1, args.steps),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(args, 1),
        "cpu_baseline": {"value": value, "unit": "pts/s", "cores": info.get("threads", 1), "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": "pts/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))
    return 0


def workload_config(args, world: int):
    if args.mode == "sharded":
        return {"workload": f"TripoSR extract_mesh {args.resolution}^3, one lattice x-slab sharded over {world} GPU(s), synthetic 3x40x64x64 baked triplane, random-init NeRFMLP, threshold=median",
                "resolution": args.resolution, "parallelism": f"xslab{world}", "l2": "inputs rotated + 256 MiB L2 flush between steps"}
    return {"workload": f"TripoSR extract_mesh {args.resolution}^3 per GPU (BASELINE configs[1]), synthetic 3x40x64x64 baked triplane, random-init NeRFMLP, threshold=median",
            "resolution": args.resolution, "scene_codes_per_gpu_per_step": args.batch, "parallelism": f"dp{world}",
            "l2": "inputs rotated + 256 MiB L2 flush between steps"}


# ------------------------------------------------------------------ SF3D (configs[4])
def run_sf3d(args) -> int:
    """One step = SF3D.triplane_to_meshes on one synthetic 3x40x384x384 triplane per GPU (independent
    objects: weak scaling).  Tet grid: Kuhn grid of tet_n^3 cubes (the reference's blob is missing)."""
    import tempfile

    import torch.distributed as dist

    from sculptmate_b200 import _capi, runtime
    from sculptmate_b200.sf3d import SF3D, save_tet_grid

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _capi.check(_capi.load().smb_device_check(), "smb_device_check")
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    n = args.tet_n
    path = save_tet_grid(os.path.join(tempfile.mkdtemp(), f"tets{n}.npz"), n)
    torch.manual_seed(0)
    m = SF3D(dict(isosurface_resolution=n, radius=RADIUS, tets_path=path)).to(dev)
    scenes = [baked_triplane(200 + rank * 2 + s, 384, 384).to(dev) for s in range(2)]
    h = m.isosurface_helper
    h.topology(dev)  # static index arrays, built once (the reference caches all_edges the same way)
    pos = m._positions(dev)
    heads = runtime.get_sf3d_heads(m.decoder, dev)
    thr = []
    for tp in scenes:
        d = runtime.sf3d_query(runtime.prepare_planes_cl(tp), heads, -1.0, RADIUS, positions=pos, want=("density_act",))["density_act"]
        thr.append(float(d.median()))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731

    def step(i):
        m.cfg.isosurface_threshold = thr[i % 2]
        return m.triplane_to_meshes(scenes[i % 2][None])[0]

    for i in range(max(args.warmup, 3)):
        step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    rec = []
    for i in range(args.steps):
        flush.fill_(i & 0xFF)
        a, b = ev(), ev()
        a.record()
        mesh = step(i)
        b.record()
        rec.append((a, b))
    torch.cuda.synchronize()
    total_ms = torch.tensor([sum(a.elapsed_time(b) for a, b in rec)], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(total_ms, op=dist.ReduceOp.MAX)
    if rank == 0:
        nv = int(h.grid_vertices.shape[0])
        ms = float(total_ms.item()) / args.steps
        print(json.dumps({
            "metric": "sf3d_triplane_to_meshes_grid_vertices_per_s", "value": nv * world / (ms * 1e-3), "unit": "pts/s", "n_gpus": world,
            "steps": args.steps, "warmup": max(args.warmup, 3), "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f16 operands, f32 accumulate (query + heads); f32 / int32 marching tets", "data": "synthetic",
            "config": {"workload": f"SF3D triplane_to_meshes (BASELINE configs[4]): 3x40x384x384 triplane, MaterialMLP density+vertex_offset heads, "
                                   f"marching tets on a Kuhn grid n={n} (Nv={nv}, Nt={int(h.indices.shape[0])}); reference blob 160_tets.npz missing",
                       "parallelism": f"dp{world}", "l2": "inputs rotated + 256 MiB L2 flush between steps"},
            "mesh": {"verts": int(mesh.v_pos.shape[0]), "tris": int(mesh.t_pos_idx.shape[0])},
            "gpu_launches": 8 * args.steps * world,
        }))
    if world > 1:
        dist.destroy_process_group()
    return 0


# ------------------------------------------------------------------ GPU arm
def main() -> int:
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--resolution", type=int, default=None)
    ap.add_argument("--mode", default="dp", choices=["dp", "sharded", "sf3d"])
    ap.add_argument("--batch", type=int, default=1, help="scene codes per GPU per step in dp mode (configs[3]: 8 per GPU on 8 GPUs)")
    ap.add_argument("--tet-n", type=int, default=160, help="sf3d mode: Kuhn tet grid of n^3 cubes (the reference's 160_tets.npz blob is missing)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    args = ap.parse_args()
    if args.mode == "sf3d" and args.impl == "ours":
        return run_sf3d(args)
    if args.resolution is None:
        args.resolution = 512 if args.mode == "sharded" else 256
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        return run_reference_arm(args, rank)

    import torch.distributed as dist

    from sculptmate_b200 import _capi, runtime
    from sculptmate_b200.dist import extract_mesh_sharded
    from sculptmate_b200.tsr import TSR

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the B200 path has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    _capi.check(_capi.load().smb_device_check(), "smb_device_check")
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)

    R = args.resolution
    torch.manual_seed(0)
    model = TSR().to(dev)  # random-init NeRFMLP, seed 0 on every rank
    n_rot = 4
    if args.mode == "sharded":
        seeds = [100 + s for s in range(n_rot)]  # same scene on every rank
    else:
        n_rot = max(n_rot, args.batch)
        seeds = [100 + rank * n_rot + s for s in range(n_rot)]
    scenes = [baked_triplane(s).to(dev) for s in seeds]
    thresholds = []
    for tp in scenes:  # setup, untimed: data-dependent threshold (median)
        d = model.renderer.query_lattice(model.decoder, tp, min(R, 128))
        thresholds.append(float(d.median()))
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    ev = lambda: torch.cuda.Event(enable_timing=True)  # noqa: E731
    pack = runtime.get_decoder_pack(model.decoder, dev)
    model.set_marching_cubes_resolution(R)
    axis = model._axis(R, dev)

    def step(i, record=None):
        tp, thr = scenes[i % n_rot], thresholds[i % n_rot]
        if args.mode == "sharded":
            e0, e1 = ev(), ev()
            e0.record()
            v, f = extract_mesh_sharded(model, tp, R, thr, broadcast=False)
            e1.record()
            if record is not None:
                record.append((e0, e1, None, None, None))
            return v, f
        e0, k0, k1, k2, e1 = ev(), ev(), ev(), ev(), ev()
        e0.record()
        for b in range(args.batch):  # serial over the batch like the reference (system.py:173)
            tp, thr = scenes[(i + b) % n_rot], thresholds[(i + b) % n_rot]
            scene = runtime.prepare_scene(tp, pack, want_cl=False, want_q=True)
            if b == 0:
                k0.record()
            dens = runtime.query_lattice(scene, pack, axis, R, RADIUS, -1.0, mc_signs=(thr, 1.0))  # case bits balloted in-kernel
            if b == 0:
                k1.record()
            v, f, _ = runtime.mc_extract(dens, sub=thr, sign=1.0, flags=7, vdiv=float(R - 1.0), vmul=float(RADIUS - (-RADIUS)), vadd=float(-RADIUS),
                                         presigned=True, on_launched=k2.record if b == 0 else None)
        e1.record()
        if record is not None:
            record.append((e0, e1, k0, k1, k2))
        return v, f

    # nvidia-smi is started BEFORE the warm-up: its start-up takes driver locks for tens of ms and must not land in the
    # timed region (it keeps sampling through it).  The warm-up visits every rotated scene once, so that no per-shape
    # capacity (speculative emit buffers) is learnt inside the timed region.
    sampler = ClockSampler(gpu_index(local_rank)) if rank == 0 else None
    args.warmup = max(args.warmup, n_rot)
    for i in range(args.warmup):
        step(i)
    torch.cuda.synchronize()
    rec = []
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    nV = nF = 0
    for i in range(args.steps):
        flush.fill_(i & 0xFF)  # L2 flush between timed steps (outside the event pairs)
        v, f = step(i, rec)
        if v is not None:
            nV, nF = int(v.shape[0]), int(f.shape[0])
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    step_ms = [r[0].elapsed_time(r[1]) for r in rec]
    total_ms = float(sum(step_ms))
    kern_ms = [k0.elapsed_time(k1) for _, _, k0, k1, _ in rec if k0 is not None]
    # marching cubes on the device (count + totals + emit, launch gaps included) / the same up to the host knowing the sizes
    mc_ms = [k1.elapsed_time(k2) for _, _, k0, k1, k2 in rec if k0 is not None] if args.batch == 1 else []
    mc_host_ms = [k1.elapsed_time(e1) for _, e1, k0, k1, _ in rec if k0 is not None] if args.batch == 1 else []

    # ---- e2e: C-ABI host-buffer call (H2D triplane, D2H mesh inside the timed region)
    e2e_s, h2d, d2h = None, 0, 0
    if args.mode == "dp" and args.batch == 1:
        lib = _capi.load()
        _, ws, bs = decoder_numpy(0)
        fpp = ctypes.POINTER(ctypes.c_float)
        W = (fpp * 10)(*[w.ctypes.data_as(fpp) for w in ws])
        B = (fpp * 10)(*[b.ctypes.data_as(fpp) for b in bs])
        ex = ctypes.c_void_p()
        _capi.check(lib.smb_extractor_create(W, B, 9, RADIUS, -1.0, 64, 64, ctypes.byref(ex)), "smb_extractor_create")
        host_tp = [np.ascontiguousarray(baked_triplane(s).numpy()) for s in seeds]
        vp, fp_ = fpp(), ctypes.POINTER(ctypes.c_int64)()
        nv, nt = ctypes.c_int64(), ctypes.c_int64()
        # the step's input lives in PINNED host memory (the handle's staging buffer, written before the clock starts);
        # the timed call does the H2D copy from there
        pin = fpp()
        _capi.check(lib.smb_extractor_pinned_input(ex, ctypes.byref(pin)), "smb_extractor_pinned_input")
        pin_np = np.ctypeslib.as_array(pin, shape=host_tp[0].shape)

        def e2e_stage(i):
            np.copyto(pin_np, host_tp[i % n_rot])

        def e2e_step(i):
            rc = lib.smb_extract_mesh_host(ex, pin, R, thresholds[i % n_rot],
                                           ctypes.byref(vp), ctypes.byref(fp_), ctypes.byref(nv), ctypes.byref(nt))
            _capi.check(rc, "smb_extract_mesh_host")

        for i in range(args.warmup):
            e2e_stage(i)
            e2e_step(i)
        if world > 1:
            dist.barrier()
        e2e_s = 0.0
        for i in range(args.steps):
            e2e_stage(i)
            flush.fill_(i & 0xFF)
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            e2e_step(i)
            e2e_s += time.perf_counter() - t0
            d2h = int(nv.value) * 12 + int(nt.value) * 24
        h2d = host_tp[0].nbytes
        lib.smb_extractor_destroy(ex)
    clocks = sampler.stop() if sampler is not None else None

    # ---- max over ranks
    stats = torch.tensor([total_ms, (e2e_s or 0.0) * 1e3, float(np.mean(kern_ms)) if kern_ms else 0.0,
                          float(np.mean(mc_ms)) if mc_ms else 0.0, float(np.mean(mc_host_ms)) if mc_host_ms else 0.0], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
    total_ms, e2e_ms, kern_ms_avg, mc_ms_avg, mc_host_ms_avg = [float(x) for x in stats.tolist()]

    if rank == 0:
        units_per_step = float(R) ** 3 * (1 if args.mode == "sharded" else world * args.batch)
        value = units_per_step * args.steps / (total_ms * 1e-3)
        peaks, peak_kind = measured_peaks()
        line = {
            "metric": "extract_mesh_lattice_points_per_s", "value": value, "unit": "pts/s", "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong" if args.mode == "sharded" else "weak", "vs_baseline": None,
            "dtype": "f16 operands, f32 accumulate (layer 0 and head bias/exp in f32)", "data": "synthetic",
            "config": workload_config(args, world), "clocks": clocks,
            "extract_mesh_ms": total_ms / args.steps, "ms_per_step_median": float(np.median(step_ms)), "ms_per_step_max": float(np.max(step_ms)),
            "mesh": {"verts": nV, "tris": nF},
            "gpu_launches": KERNELS_PER_STEP * args.steps * world * (args.batch if args.mode == "dp" else 1),
        }
        if e2e_s is not None:
            line["e2e"] = {"value": units_per_step * args.steps / (e2e_ms * 1e-3), "unit": "pts/s", "ms_per_step": e2e_ms / args.steps,
                           "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "api": "smb_extract_mesh_host (C ABI; triplane in pinned host memory, mesh returned in pinned host memory)"}
        if kern_ms_avg > 0:
            flops = FLOP_PER_POINT * float(R) ** 3
            ach = flops / (kern_ms_avg * 1e-3) / 1e12
            traffic = None
            tpath = os.path.join(ROOT, "profiles", "traffic.json")
            if os.path.exists(tpath):
                with open(tpath) as fh:
                    traffic = json.load(fh).get(f"lattice_tc_kernel_R{R}")
            line["roofline"] = {
                "kernel": "lattice_tc_ta_kernel", "bound": "tensor", "achieved": ach, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": ach / peaks["bf16_tflops"], "traffic": traffic, "peak_source": f"{peak_kind} (burst)",
                "kernel_ms": kern_ms_avg, "query_points_per_s": float(R) ** 3 / (kern_ms_avg * 1e-3),
                "flop_per_point": FLOP_PER_POINT,
                "note": "algorithmic FLOPs = the reference NeRFMLP count (SURVEY 8d); layer 0 runs as a projected-plane interpolation in fp32, not as an MMA",
            }
        if mc_ms_avg > 0 and nV:
            # marching cubes is HBM-bound: algorithmic bytes = 4 R^3 (density read) + 12 V + 24 F (mesh write),
            # SURVEY 8d; the time spans count + totals + emit on the device (CUDA events on the launching stream)
            mc_bytes = 4.0 * float(R) ** 3 + 12.0 * nV + 24.0 * nF
            line["roofline_mc"] = {
                "kernels": "mc_count+mc_totals+mc_emit (sign masks come from the lattice kernel; emit launched behind count; the sizes are read back after it)", "bound": "hbm",
                "achieved": mc_bytes / (mc_ms_avg * 1e-3) / 1e9, "peak": peaks["hbm_gbs"], "unit": "GB/s",
                "frac": mc_bytes / (mc_ms_avg * 1e-3) / 1e9 / peaks["hbm_gbs"], "ms": mc_ms_avg, "ms_until_host_has_sizes": mc_host_ms_avg,
                "algorithmic_bytes": mc_bytes,
                "peak_source": peak_kind,
            }
        if not args.no_cpu_baseline and world == 1:
            v, desc, info = cpu_reference_sample(R, args.cpu_seconds, thresholds[0])
            line["cpu_baseline"] = {"value": v, "unit": "pts/s", "cores": info["threads"], "kind": "port", "sample": desc}
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
