#!/usr/bin/env python
"""Secondary measurements (not the headline bench): configs 1, 3, 4 of BASELINE.json, the Gaussian
FIR on a cube slab, and the latency of small drop-in calls.  One JSON object per line.

    python tools/bench_configs.py [--reps 10]
"""
import argparse
import json
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from shgyield_b200 import _lib, grid, shg  # noqa: E402
from shgyield_b200.synthetic import config5  # noqa: E402
from tests import cases  # noqa: E402


def timed(fn, reps, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(reps)]
    for a, b in ev:
        a.record()
        fn()
        b.record()
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    return ms[len(ms) // 2], ms[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--only-k1", action="store_true")
    ap.add_argument("--only-k3", action="store_true")
    ap.add_argument("--only-configs", action="store_true", help="configs 3 and 4 only")
    args = ap.parse_args()
    dev = torch.device("cuda", 0)
    put = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    consts = shg.physical_constants()
    ctx = _lib.context(0)
    m1, m2, m3, chi2 = cases.si111_material()

    # ---- configs 3 and 4: Si(111) material, device resident; prep (K1, spline, rotate) timed apart -----
    for tag, axes in (() if (args.only_k1 or args.only_k3) else (("config3", cases.config3_axes()), ("config4", cases.config4_axes()))):
        t0 = time.perf_counter()
        eps1w, eps2w, chi_rot, thick = shg._prepare(axes["energy"], m1, m2, m3, chi2, axes["thick"], axes["gamma"], 0.0,
                                                    0.0, "3-layer")
        prep_ms = 1e3 * (time.perf_counter() - t0)
        d = [put(axes["energy"]), put(axes["theta"]), put(axes["phi"]), put(np.atleast_1d(thick).astype(float)),
             put(eps1w), put(eps2w), put(chi_rot)]
        T, D, P, E = axes["theta"].size, axes["thick"].size, axes["phi"].size, axes["energy"].size
        out = torch.empty((T, D, 4, P, E), dtype=torch.float64, device=dev)
        for nt in (64, 256):
            with ctx.options(k2_nt=nt):
                med, best = timed(lambda: grid.yield_grid_device(*d, out=out, consts=consts), args.reps)
            print(json.dumps({"case": tag, "k2_threads": nt, "k2_ms_median": med, "k2_ms_best": best}), flush=True)
        med, best = timed(lambda: grid.yield_grid_device(*d, out=out, consts=consts), args.reps)
        pts = out.numel()
        flops = 28.0 * pts + 1000.0 * E * T * D
        print(json.dumps({"case": tag, "shape_TDPE": [T, D, P, E], "points": pts, "k2_ms_median": med, "k2_ms_best": best,
                          "points_per_s": pts / (med * 1e-3), "store_gbs": 8e-9 * pts / (med * 1e-3),
                          "alg_fp64_tflops": flops / (med * 1e-3) / 1e12, "host_prep_ms": prep_ms}), flush=True)
        for sigma in (5.0, 12.0):                   # the API's default output broadening (K3 for config 3, K2s -> K1 for config 4)
            med, best = timed(lambda: grid.yield_grid_device(*d, out=out, consts=consts, sigma_out=sigma), args.reps)
            print(json.dumps({"case": tag, "sigma_out": sigma, "ms_median": med, "ms_best": best,
                              "points_per_s": pts / (med * 1e-3), "store_gbs": 8e-9 * pts / (med * 1e-3)}), flush=True)
        del out

    if args.only_configs:
        return

    # ---- config-5 theta shards (what one of 8 / 4 ranks computes), both CTA sizes of K2 ----------------
    cfg = config5()
    dd = {k: put(v) for k, v in cfg.items()}
    for nth in (() if (args.only_k1 or args.only_k3) else (23, 46)):
        out = torch.empty((nth, 1, 4, 721, 20000), dtype=torch.float64, device=dev)
        for nt in (64, 256, None):
            with ctx.options(k2_nt=nt):
                med, best = timed(lambda: grid.yield_grid_device(dd["energy"], dd["theta"], dd["phi"], dd["thick"],
                                                                dd["eps1w"], dd["eps2w"], dd["chi2"], out=out, consts=consts,
                                                                t_range=(0, nth)), args.reps)
            print(json.dumps({"case": "config5_shard_%dtheta" % nth, "k2_threads": nt or "auto", "ms_median": med,
                              "ms_best": best, "store_gbs": 8e-9 * out.numel() / (med * 1e-3)}), flush=True)
        del out

    # ---- K3: config-5 theta shards with the output broadening (sigma_out samples along E) -------------
    # fused (one pass, FIR on the 13 azimuthal harmonics of a column) against two-pass (K2 -> K1 by slabs)
    for nth in (() if args.only_k1 else (23, 46)):
        out = torch.empty((nth, 1, 4, 721, 20000), dtype=torch.float64, device=dev)
        for sigma in (2.0, 5.0, 12.0):
            for k3 in ("1", "0"):
                with ctx.options(k3=int(k3)):
                    med, best = timed(lambda: grid.yield_grid_device(dd["energy"], dd["theta"], dd["phi"], dd["thick"],
                                                                    dd["eps1w"], dd["eps2w"], dd["chi2"], out=out,
                                                                    consts=consts, t_range=(0, nth), sigma_out=sigma),
                                      args.reps)
                print(json.dumps({"case": "config5_shard_%dtheta_sigma_out%g" % (nth, sigma),
                                  "path": "K3 fused" if k3 == "1" else "K2 -> K1 two-pass", "ms_median": med, "ms_best": best,
                                  "points_per_s": out.numel() / (med * 1e-3),
                                  "store_gbs": 8e-9 * out.numel() / (med * 1e-3)}), flush=True)
        del out
    if args.only_k3:
        return

    # ---- K1 on a cube slab: 8 thetas of config 5, sigma = 5 samples along E ---------------------------
    slab = grid.yield_grid_device(dd["energy"], dd["theta"], dd["phi"], dd["thick"], dd["eps1w"], dd["eps2w"], dd["chi2"],
                                  consts=consts, t_range=(60, 68))
    for sigma in (2.0, 5.0, 12.0):
        med, best = timed(lambda: grid.broaden_energy_device(slab, sigma), args.reps)
        n = slab.numel()
        taps = 2 * int(4 * sigma + 0.5) + 1
        print(json.dumps({"case": "k1_cube_sigma%g" % sigma, "samples": n, "taps": taps, "ms_median": med, "ms_best": best,
                          "samples_per_s": n / (med * 1e-3), "rw_gbs": 16e-9 * n / (med * 1e-3),
                          "fp64_tflops": 2.0 * taps * n / (med * 1e-3) / 1e12}), flush=True)
    del slab
    if args.only_k1:
        return

    # ---- latency of the reference's own call shapes through the drop-in API (host in, host out) -------
    def wall(fn, reps=20):
        fn()
        fn()
        t0 = time.perf_counter()
        for _ in range(reps):
            fn()
        return 1e3 * (time.perf_counter() - t0) / reps

    kw = dict(eps_m1=m1, eps_m2=m2, eps_m3=m3, chi2=chi2, theta=65, thick=10, gamma=0)
    ms = wall(lambda: shg.shgyield(energy=cases.E_C1, phi=30, sigma_out=5, **kw))
    print(json.dumps({"case": "config1_spectrum_call", "shape": "E=126, 1 theta, 1 phi, sigma_out=5", "ms_per_call": ms,
                      "points_per_s": 504 / (ms * 1e-3)}), flush=True)
    ms = wall(lambda: shg.shgyield(energy=1.7, phi=np.arange(0, 361), sigma_out=0.0, **kw))
    print(json.dumps({"case": "gui_polar_call", "shape": "E scalar, 361 phi", "ms_per_call": ms,
                      "points_per_s": 1444 / (ms * 1e-3)}), flush=True)
    ms = wall(lambda: shg.shgyield(energy=cases.E_C1, phi=30, sigma_eps=3.0, sigma_chi=3.0, sigma_out=5, **kw))
    print(json.dumps({"case": "config1_all_sigmas_call", "shape": "E=126, sigma_eps=sigma_chi=3, sigma_out=5",
                      "ms_per_call": ms}), flush=True)
    print(json.dumps({"case": "launch_count", "launches": _lib.context(0).launches}), flush=True)


if __name__ == "__main__":
    main()
