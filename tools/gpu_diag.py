#!/usr/bin/env python3
"""One-shot GPU diagnostics: numbers, not pass/fail.  Run under gpurun:
    python tools/gpu_diag.py [--quick]
Prints parity statistics of every kernel against the oracle and rough timings."""
from __future__ import annotations

import os
import sys
import time
import traceback

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from conftest import golden_decoder, volume  # noqa: E402
from oracle import field_oracle as fo  # noqa: E402
from oracle import mc_oracle  # noqa: E402
from sculptmate_b200 import _capi, runtime  # noqa: E402
from sculptmate_b200.tsr import TSR  # noqa: E402

dev = torch.device("cuda:0")
GOLD = os.path.join(ROOT, "tests", "golden")


def section(name):
    print(f"\n===== {name}", flush=True)


def timed(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    evs = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        evs.append((a, b))
    torch.cuda.synchronize()
    t = sorted(a.elapsed_time(b) for a, b in evs)
    return t[len(t) // 2], t[0]


def diag_mc():
    section("marching cubes vs oracle (bit-exact expected)")
    for kind, shape in [("sphere", (16, 16, 16)), ("noise", (20, 20, 20)), ("gyroid", (33, 33, 33)), ("smooth", (40, 17, 70)),
                        ("noise", (9, 33, 65)), ("sphere", (64, 64, 64)), ("gyroid", (128, 128, 128))]:
        R = max(shape)
        g = volume(kind, R, seed=1)[: shape[0], : shape[1], : shape[2]].copy()
        gd = torch.from_numpy(g).to(dev)
        for emit_last in (True, False):
            v_ref, f_ref, c = mc_oracle.marching_cubes_slab(g, sub=0.02, sign=-1.0, x_origin=3, emit_last_plane=emit_last,
                                                            flags=7, vdiv=float(R - 1), vmul=1.74, vadd=-0.87)
            pend = runtime.mc_count(gd, sub=0.02, sign=-1.0, emit_last_plane=emit_last)
            v, f = runtime.mc_emit(pend, x_origin=3, flags=7, vdiv=float(R - 1), vmul=1.74, vadd=-0.87, vertex_id_offset=0)
            v, f = v.cpu().numpy(), f.cpu().numpy()
            ok_counts = (pend.nverts, pend.ntris, pend.nverts_numbered) == (c.nverts, c.ntris, c.nverts_numbered)
            ok_v = v.shape == v_ref.shape and np.array_equal(v.view(np.uint32), v_ref.view(np.uint32))
            ok_f = f.shape == f_ref.shape and np.array_equal(f, f_ref)
            print(f"{kind:7s} {str(shape):16s} last={int(emit_last)} V={pend.nverts} F={pend.ntris} counts_ok={ok_counts} verts_bitexact={ok_v} faces_exact={ok_f}")
            if not ok_v and v.shape == v_ref.shape:
                bad = np.nonzero((v != v_ref).any(1))[0]
                print("   first bad verts", bad[:5], v[bad[:3]], v_ref[bad[:3]])
            if not ok_f and f.shape == f_ref.shape:
                bad = np.nonzero((f != f_ref).any(1))[0]
                print("   first bad faces", bad[:5], f[bad[:3]], f_ref[bad[:3]])
        cs = runtime.mc_cases(gd, 0.02, -1.0).cpu().numpy()
        print("   cases exact:", np.array_equal(cs, mc_oracle.cube_cases(g, 0.02, -1.0)))


def load_decoder(g):
    ws, bs = golden_decoder(g)
    blob, lay = runtime.pack_decoder_host([torch.from_numpy(w) for w in ws], [torch.from_numpy(b) for b in bs])
    return ws, bs, runtime.DecoderPack(blob.to(dev), lay, ("diag",))


def diag_field():
    section("fp32 query kernel vs reference goldens")
    for name in ("field_small.npz", "field_64.npz"):
        g = np.load(os.path.join(GOLD, name))
        ws, bs, pack = load_decoder(g)
        if "triplane" in g:
            tp = torch.from_numpy(g["triplane"])
        else:
            torch.manual_seed(int(g["triplane_seed"]))
            tp = torch.randn(3, 40, 64, 64)
        scene = runtime.prepare_scene(tp.to(dev), pack)
        out = runtime.query_points(scene, pack, torch.from_numpy(g["positions"]).to(dev), 0.87, -1.0)
        for k in ("density", "features", "density_act", "color"):
            d = np.abs(out[k].cpu().numpy() - g[k])
            print(f"{name:16s} {k:12s} max_abs={d.max():.3e}")

    section("tensor-core lattice kernel vs oracle / fp32 kernel")
    g = np.load(os.path.join(GOLD, "field_64.npz"))
    ws, bs, pack = load_decoder(g)
    torch.manual_seed(int(g["triplane_seed"]))
    tp = torch.randn(3, 40, 64, 64)
    scene = runtime.prepare_scene(tp.to(dev), pack)
    for R in (8, 32, 64, 100, 128):
        ax = runtime.lattice_axis(R, 0.87, device=dev)
        act32, raw32 = runtime.query_lattice(scene, pack, ax, R, 0.87, -1.0, precision="fp32", want_raw=True)
        try:
            act, raw = runtime.query_lattice(scene, pack, ax, R, 0.87, -1.0, precision="tc", want_raw=True)
            torch.cuda.synchronize()
        except Exception as e:  # noqa: BLE001
            print(f"R={R}: tc kernel failed: {e}")
            raise
        d = (raw - raw32).abs()
        rel = (act / act32 - 1).abs()
        print(f"R={R:4d} tc-vs-fp32 logit max_abs={d.max().item():.3e} mean_abs={d.mean().item():.3e} density_act max_rel={rel.max().item():.3e}  logit range [{raw32.min().item():.3f},{raw32.max().item():.3f}] nan={int(torch.isnan(raw).sum())}")
        if R <= 32:
            ref = fo.grid_density(R, tp.numpy(), ws, bs)
            print(f"        fp32-kernel vs oracle density_act max_rel={np.abs(act32.cpu().numpy() / ref - 1).max():.3e}; tc vs oracle max_rel={np.abs(act.cpu().numpy() / ref - 1).max():.3e}")
        # slab consistency
        a2 = runtime.query_lattice(scene, pack, ax, R, 0.87, -1.0, x_begin=R // 3, nx=R // 2, precision="tc")
        print("        slab == full:", torch.equal(a2, act[R // 3 : R // 3 + R // 2]))


def diag_e2e(quick):
    section("extract_mesh end to end")
    g = np.load(os.path.join(GOLD, "extract_mesh.npz"))
    ws, bs = golden_decoder(g)
    m = TSR().to(dev)
    sd = {}
    for i in range(10):
        sd[f"layers.{2 * i}.weight"] = torch.from_numpy(ws[i])
        sd[f"layers.{2 * i}.bias"] = torch.from_numpy(bs[i])
    m.decoder.load_state_dict(sd)
    m.to(dev)
    tp = torch.from_numpy(g["triplane"]).to(dev)
    R, thr = int(g["resolution"]), float(g["threshold"])
    for prec in ("fp32", "tc"):
        v, f = m.extract_mesh_tensors(tp, R, thr, precision=prec)
        print(f"golden extract_mesh R={R} {prec}: V={len(v)} F={len(f)} (reference+oracle MC: V={len(g['verts'])} F={len(g['faces'])})",
              "verts_equal=", v.shape == g["verts"].shape and np.array_equal(v.cpu().numpy(), g["verts"]),
              "faces_equal=", f.shape == g["faces"].shape and np.array_equal(f.cpu().numpy(), g["faces"]))
    # bigger: baked field at 64^2 planes
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from make_golden import baked_triplane

    tp = baked_triplane(5, 64, 64).to(dev)
    for R in ((128,) if quick else (128, 256)):
        dens = m.renderer.query_lattice(m.decoder, tp, R)
        thr = float(dens.median())
        t_q = timed(lambda: m.renderer.query_lattice(m.decoder, tp, R))
        pend = runtime.mc_count(dens, sub=thr)
        t_c = timed(lambda: runtime.mc_count(dens, sub=thr))
        t_e = timed(lambda: runtime.mc_emit(pend, flags=7, vdiv=R - 1.0, vmul=1.74, vadd=-0.87))
        t_all = timed(lambda: m.extract_mesh_tensors(tp, R, thr))
        print(f"R={R}: query_lattice(tc) {t_q[0]:.3f} ms (min {t_q[1]:.3f}) = {R**3 / t_q[1] / 1e6:.1f} Gpts/s | mc_count {t_c[0]:.3f} | mc_emit {t_e[0]:.3f} | extract_mesh_tensors {t_all[0]:.3f} ms  V={pend.nverts} F={pend.ntris}")
        t32 = timed(lambda: m.renderer.query_lattice(m.decoder, tp, R, precision="fp32"), iters=2, warm=1)
        print(f"       query_lattice(fp32 cuda cores) {t32[0]:.2f} ms")
        # MC parity at this size vs oracle on the GPU's own density grid
        dh = dens.cpu().numpy()
        t0 = time.perf_counter()
        v_ref, f_ref, c = mc_oracle.marching_cubes_slab(dh, sub=thr, flags=7, vdiv=R - 1.0, vmul=1.74, vadd=-0.87)
        t1 = time.perf_counter()
        v, f = runtime.mc_emit(pend, flags=7, vdiv=R - 1.0, vmul=1.74, vadd=-0.87)
        print(f"       MC vs oracle at R={R}: verts_bitexact={np.array_equal(v.cpu().numpy(), v_ref)} faces_exact={np.array_equal(f.cpu().numpy(), f_ref)} (oracle MC {t1 - t0:.2f} s)")


def diag_host_extractor():
    section("C-ABI host-buffer extractor")
    import ctypes

    lib = _capi.load()
    g = np.load(os.path.join(GOLD, "extract_mesh.npz"))
    ws, bs = golden_decoder(g)
    fpp = ctypes.POINTER(ctypes.c_float)
    ws_c = [np.ascontiguousarray(w) for w in ws]
    bs_c = [np.ascontiguousarray(b) for b in bs]
    W = (fpp * 10)(*[w.ctypes.data_as(fpp) for w in ws_c])
    B = (fpp * 10)(*[b.ctypes.data_as(fpp) for b in bs_c])
    ex = ctypes.c_void_p()
    rc = lib.smb_extractor_create(W, B, 9, 0.87, -1.0, 16, 16, ctypes.byref(ex))
    print("create rc", rc)
    tp = np.ascontiguousarray(g["triplane"])
    vp, fp_ = fpp(), ctypes.POINTER(ctypes.c_int64)()
    nv, nt = ctypes.c_int64(), ctypes.c_int64()
    R, thr = int(g["resolution"]), float(g["threshold"])
    rc = lib.smb_extract_mesh_host(ex, tp.ctypes.data_as(fpp), R, thr, ctypes.byref(vp), ctypes.byref(fp_), ctypes.byref(nv), ctypes.byref(nt))
    print("extract rc", rc, "V", nv.value, "F", nt.value, "golden", g["verts"].shape, g["faces"].shape)
    rc = lib.smb_extract_mesh_host(ex, tp.ctypes.data_as(fpp), R, 1e9, ctypes.byref(vp), ctypes.byref(fp_), ctypes.byref(nv), ctypes.byref(nt))
    print("threshold above range -> rc", rc, _capi.status_string(rc))
    lib.smb_extractor_destroy(ex)


def main():
    quick = "--quick" in sys.argv
    print("device:", torch.cuda.get_device_name(0), "check:", _capi.load().smb_device_check())
    for fn in (diag_mc, diag_field, lambda: diag_e2e(quick), diag_host_extractor):
        try:
            fn()
        except Exception:  # noqa: BLE001
            traceback.print_exc()
            try:
                torch.cuda.synchronize()
            except Exception as e:  # noqa: BLE001
                print("CUDA context is dead:", e)
                return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
