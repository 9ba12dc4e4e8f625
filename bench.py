#!/usr/bin/env python
"""
bench.py -- SSHG yield points/s on the "synthetic scale" sweep of BASELINE.json
(20k energies x 181 theta x 721 phi x 4 polarisations, random 27-component chi2; configs[4]).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one evaluation of the whole cube by the fused yield kernel (K2), the spectra already
resident in HBM; with N ranks the theta axis is cut into N contiguous shards (no data-path
collective, strong scaling: the cube is fixed).  Rank 0 prints ONE JSON line.

  value      whole-job points/s, device-resident (CUDA events, max over ranks)
  e2e        the same metric through the host-buffer C-ABI call (H2D of the spectra, kernel, D2H of the
             yields into pinned host memory inside the timed region) on a theta sub-sample of each shard
  roofline   HBM bound: algorithmic bytes (8 B per point, written once) / average K2 launch time
  cpu_baseline  the reference's own CPU code on the box's host cores, bounded sample: the UNMODIFIED reference
             (rad_pp .. rad_ss + the prefactor line of shg.py:428-436) when it is installed in baseline/_ref
             (oracle/install_reference.py, run by __graft_entry__.build(); kind "reference"), else the NumPy
             oracle port (oracle/shg_oracle.py; kind "port")
  configs    N = 1 only, outside the timed region: BASELINE.json configs[2] and [3] (Si(111) angular map, thin-film
             sweep) without and with the API's default output broadening

`--impl reference` times that same CPU code on all host cores (rank 0 only).
"""
import argparse
import json
import os
import sys
import time

import numpy as np

REPO = os.path.dirname(os.path.abspath(__file__))
if REPO not in sys.path:
    sys.path.insert(0, REPO)

N_ENERGY, N_THETA, N_PHI = 20000, 181, 721
WORKLOAD = "synthetic scale (BASELINE.json configs[4]): 20000 E x 181 theta x 721 phi x 4 pol, d = 10 nm, seed 20250"
METRIC = "SSHG yield points/sec (E x theta x phi x pol)"
# ncu DRAM traffic of yield_grid_kernel: parsed from the committed summary of a --set full capture (a 37-theta launch of
# the same kernel on the same sweep; the full 83.5 GB launch itself is not captured: ncu replays each kernel ~40 times
# with save/restore of everything it writes)
K2_PROFILE = os.path.join("profiles", "r02_k2_ncu_full.txt")
E2E_THETAS = 32           # thetas of each rank's shard that the end-to-end leg evaluates per step (clamped to the shard)
CPU_PHI_CHUNK = 8         # phi values per CPU task (keeps NumPy temporaries ~100 MB)


# ------------------------------------------------------------------------------------------------
# CPU leg: the reference's own code (or, when it is not installed, the oracle port) fanned over the host cores
# ------------------------------------------------------------------------------------------------
_CPU = {}


def workload_config():
    """`config` of the JSON line: the workload only, identical in both arms."""
    return {"workload": WORKLOAD,
            "shape": {"energy": N_ENERGY, "theta": N_THETA, "phi": N_PHI, "thick": 1, "pol": 4},
            "layout": "[theta][d][pol][phi][E] f64",
            "parallelism": "theta shards, one per GPU, no data-path collective",
            "l2": "every step writes >= 10 GB per GPU (>> 126 MB L2), never re-read; inputs 7.7 MB"}


def find_reference():
    """The UNMODIFIED reference module shgyield/shg.py, loaded by path under a private name (the repository's own
    `shgyield` package is the drop-in shim): $SSHG_REFERENCE_DIR, else baseline/_ref (pip-installed there from the
    read-only checkout by oracle/install_reference.py; git-ignored, travels to the GPU box).  /root/reference itself
    is never read at run time."""
    import importlib.util
    for d in (os.environ.get("SSHG_REFERENCE_DIR"), os.path.join(REPO, "baseline", "_ref")):
        path = os.path.join(d, "shgyield", "shg.py") if d else None
        if path and os.path.isfile(path):
            import warnings
            spec = importlib.util.spec_from_file_location("sshg_reference_shg", path)
            mod = importlib.util.module_from_spec(spec)
            with warnings.catch_warnings():
                warnings.simplefilter("ignore")
                spec.loader.exec_module(mod)
            return mod, d
    return None, None


def _cpu_init():
    from shgyield_b200.synthetic import config5
    cfg = config5(N_ENERGY, N_THETA, N_PHI)
    keys = ("xxx", "xyy", "xzz", "xyz", "xxz", "xxy", "yxx", "yyy", "yzz", "yyz", "yxz", "yxy",
            "zxx", "zyy", "zzz", "zyz", "zxz", "zxy")                                  # reference shg.py:66-133
    _CPU["cfg"] = cfg
    _CPU["e1"] = {"M%d" % (m + 1): cfg["eps1w"][m] for m in range(3)}
    _CPU["e2"] = {"M%d" % (m + 1): cfg["eps2w"][m] for m in range(3)}
    _CPU["chi"] = {k: cfg["chi2"][i] for i, k in enumerate(keys)}
    ref, where = find_reference()
    _CPU["ref"] = ref
    if ref is None:
        from oracle import shg_oracle as orc
        _CPU["orc"] = orc


def cpu_kind():
    ref, where = find_reference()
    if ref is not None:
        return "reference", "the UNMODIFIED reference (shgyield/shg.py in %s): rad_pp .. rad_ss and the prefactor line of " \
                            "shg.py:428-436, NumPy broadcast over a phi chunk" % os.path.relpath(where, REPO)
    return "port", "NumPy oracle port of the reference's shg.py (oracle/shg_oracle.py; the reference is not installed in baseline/_ref)"


def _cpu_task(task):
    """One theta x CPU_PHI_CHUNK phi x all energies x 4 polarisations; returns the point count."""
    if not _CPU:
        _cpu_init()
    it, ip = task
    cfg = _CPU["cfg"]
    phi = cfg["phi"][ip:ip + CPU_PHI_CHUNK]
    ref = _CPU["ref"]
    if ref is None:
        orc = _CPU["orc"]
        y = orc.yield_points(cfg["energy"], _CPU["e1"], _CPU["e2"], _CPU["chi"], cfg["theta"][it], phi[:, None], 10.0,
                             True, orc.SCIPY_1_18_CONSTANTS)
        return sum(int(np.asarray(y[p]).size) for p in ("pp", "ps", "sp", "ss"))
    # the reference's own functions, called as its shgyield() calls them (shg.py:428-436, sigma_out = 0); the spectra
    # are fed at this level because config 5 is defined on the arrays rad_* take
    import warnings
    from scipy import constants
    energy, e1, e2, chi = cfg["energy"], _CPU["e1"], _CPU["e2"], _CPU["chi"]
    th, ph = np.radians(cfg["theta"][it]), np.radians(phi[:, None])
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")             # scipy's ConstantWarning for the legacy CODATA key the reference uses
        hbar = constants.value("Planck constant over 2 pi in eV s")
    pref = 1e4 * ((energy / hbar) ** 2) / (2 * constants.epsilon_0 * constants.c ** 3 * np.cos(th) ** 2)
    n = 0
    for rad in (ref.rad_pp, ref.rad_ps, ref.rad_sp, ref.rad_ss):
        y = pref * np.absolute(1 / np.sqrt(e1["M2"]) * rad(energy, e1, e2, chi, th, ph, 10.0, True)) ** 2
        n += int(y.size)
    return n


def _cpu_tasks(n, offset=0):
    tasks = []
    for k in range(n):
        j = offset + k
        tasks.append(((7 * j) % (N_THETA - 1), (CPU_PHI_CHUNK * j) % (N_PHI - CPU_PHI_CHUNK)))
    return tasks


def cpu_cores():
    try:
        return len(os.sched_getaffinity(0))
    except AttributeError:
        return os.cpu_count() or 1


class CpuPool:
    """Process pool over the host cores (fork; created before CUDA is initialised)."""

    def __init__(self, workers):
        import multiprocessing as mp
        self.workers = workers
        self.pool = mp.get_context("fork").Pool(workers, initializer=_cpu_init) if workers > 1 else None
        if self.pool is None:
            _cpu_init()

    def run(self, tasks):
        t0 = time.perf_counter()
        if self.pool is None:
            pts = sum(_cpu_task(t) for t in tasks)
        else:
            pts = sum(self.pool.map(_cpu_task, tasks, chunksize=1))
        return pts, time.perf_counter() - t0

    def close(self):
        if self.pool is not None:
            self.pool.close()
            self.pool.join()


def cpu_baseline(target_s=12.0):
    """Bounded sample on all cores: calibrate with one task per worker, then ~target_s of work."""
    cores = cpu_cores()
    kind, what = cpu_kind()
    pool = CpuPool(cores)
    try:
        pool.run(_cpu_tasks(cores))                                  # warm-up (imports, page faults)
        pts, dt = pool.run(_cpu_tasks(cores, cores))
        rounds = max(1, min(120, int(target_s / max(dt, 1e-3))))
        pts, dt = pool.run(_cpu_tasks(cores * rounds, 2 * cores))
    finally:
        pool.close()
    # the reference itself is single-threaded: the same tasks on ONE core (SURVEY 8d asks for both figures)
    _cpu_task(_cpu_tasks(1)[0])
    t0 = time.perf_counter()
    pts1 = sum(_cpu_task(t) for t in _cpu_tasks(4, 1))
    dt1 = time.perf_counter() - t0
    return {"value": pts / dt, "unit": "points/s", "cores": cores, "kind": kind,
            "sample": "%d tasks of 1 theta x %d phi x %d E x 4 pol (%.3g points) of the same workload, %s, one process "
                      "per core" % (cores * rounds, CPU_PHI_CHUNK, N_ENERGY, pts, what),
            "seconds": dt, "single_core_value": pts1 / dt1}


def run_reference(args):
    """--impl reference: the reference's own CPU code (else the oracle port) on all host cores, rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = cpu_cores()
    kind, what = cpu_kind()
    pool = CpuPool(cores)
    tasks_per_step = 4 * cores
    try:
        for w in range(args.warmup):
            pool.run(_cpu_tasks(tasks_per_step, w * tasks_per_step))
        t0 = time.perf_counter()
        pts = 0
        for s in range(args.steps):
            p, _ = pool.run(_cpu_tasks(tasks_per_step, (args.warmup + s) * tasks_per_step))
            pts += p
        dt = time.perf_counter() - t0
    finally:
        pool.close()
    value = pts / dt
    sample = "each step = %d tasks of 1 theta x %d phi x %d E x 4 pol (%.3g points) of the workload; %s" % (
        tasks_per_step, CPU_PHI_CHUNK, N_ENERGY, pts / max(args.steps, 1), what)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / max(args.steps, 1),
            "higher_is_better": True, "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(),
            "cpu_baseline": {"value": value, "unit": "points/s", "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": "points/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------
# GPU leg
# ------------------------------------------------------------------------------------------------
class ClockSampler:
    """Samples SM clock, power and clock-event reasons of one GPU DURING the timed region: an NVML
    polling thread (2 ms period; the timed region of this bench is ~0.1 s, too short for
    `nvidia-smi -lms`)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap",
               0x80: "hw_power_brake_slowdown"}

    def __init__(self, torch_index):
        self.samples = []
        self.thread = None
        self.stop_flag = False
        self.h = None
        try:
            import pynvml
            import torch
            pynvml.nvmlInit()
            self.nv = pynvml
            try:
                uuid = "GPU-" + str(torch.cuda.get_device_properties(torch_index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(uuid.encode() if hasattr(uuid, "encode") else uuid)
            except Exception:
                vis = os.environ.get("CUDA_VISIBLE_DEVICES")
                idx = int(vis.split(",")[torch_index]) if vis and vis.split(",")[torch_index].isdigit() else torch_index
                self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
        except Exception as e:                      # NVML missing: the record says so
            self.err = repr(e)

    def _poll(self):
        nv, h = self.nv, self.h
        while not self.stop_flag:
            try:
                clk = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                pw = nv.nvmlDeviceGetPowerUsage(h) / 1000.0
                try:
                    rs = nv.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    rs = nv.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                self.samples.append((clk, pw, rs))
            except Exception:
                pass
            time.sleep(0.002)

    def start(self):
        if self.h is None:
            return
        import threading
        self.thread = threading.Thread(target=self._poll, daemon=True)
        self.thread.start()

    def stop(self):
        if self.thread is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable: " + getattr(self, "err", "?")]}
        self.stop_flag = True
        self.thread.join(timeout=2)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["no samples"]}
        clk = [s[0] for s in self.samples]
        pw = [s[1] for s in self.samples]
        bits = 0
        for s in self.samples:
            bits |= s[2]
        smax = self.nv.nvmlDeviceGetMaxClockInfo(self.h, self.nv.NVML_CLOCK_SM)
        busy = [c for c, p in zip(clk, pw) if p >= 0.5 * max(pw)] or clk
        return {"sm_mhz": float(np.median(busy)), "sm_max_mhz": float(smax),
                "reasons": sorted(n for b, n in self.REASONS.items() if bits & b),
                "samples": len(clk), "sm_mhz_min": float(min(clk)), "power_w_max": float(max(pw))}


def measured_peaks():
    path = os.path.join(REPO, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            d = json.load(f)
        return float(d["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (copy, read+write, burst)"
    except Exception:
        return 6650.0, "fallback of B200_PROFILING.md (MEASURED_PEAKS.json absent)"


def k2_traffic():
    """(dram read + write bytes, algorithmic bytes, source) of one captured yield_grid_kernel launch, parsed from the
    committed ncu summary (tools/ncu_summary.py full ...); None when no summary is present."""
    import re
    unit = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12}
    for rel in (K2_PROFILE,):
        path = os.path.join(REPO, rel)
        if not os.path.isfile(path):
            continue
        txt = open(path).read()
        if "yield_grid_kernel" not in txt:
            continue
        got = {}
        for key in ("dram__bytes_read.sum", "dram__bytes_write.sum"):
            m = re.search(r"^%s\s+(\w+)\s+([0-9.eE+-]+)" % re.escape(key), txt, re.M)
            if m and m.group(1) in unit:
                got[key] = float(m.group(2)) * unit[m.group(1)]
        m = re.search(r"^algorithmic_bytes\s+byte\s+([0-9.eE+]+)", txt, re.M)      # written by tools/ncu_summary.py alg_bytes=N
        if len(got) == 2 and m:
            return got["dram__bytes_read.sum"] + got["dram__bytes_write.sum"], float(m.group(1)), rel
    return None


def run_configs(ctx, dev, consts, peak, fp64_peak):
    """BASELINE.json configs[2] (Si(111) angular map) and configs[3] (thin-film sweep) on one GPU, device resident,
    without and with the API's default output broadening: median of 9 launches (CUDA events), outside the timed region
    of the headline.  Bounds as SURVEY.md 8d names them: HBM stores for config 3, FP64 for config 4."""
    import torch
    from shgyield_b200 import grid, shg
    from shgyield_b200.synthetic import config3_axes, config4_axes, si111_material
    m1, m2, m3, chi2 = si111_material()
    put = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    res = {}
    for tag, axes in (("config3", config3_axes()), ("config4", config4_axes())):
        eps1w, eps2w, chi_rot, thick = shg._prepare(axes["energy"], m1, m2, m3, chi2, axes["thick"], axes["gamma"], 0.0, 0.0,
                                                    "3-layer")
        d = [put(axes["energy"]), put(axes["theta"]), put(axes["phi"]), put(np.atleast_1d(thick).astype(float)),
             put(eps1w), put(eps2w), put(chi_rot)]
        T, D, P, E = axes["theta"].size, axes["thick"].size, axes["phi"].size, axes["energy"].size
        out = torch.empty((T, D, 4, P, E), dtype=torch.float64, device=dev)
        pts = out.numel()
        flops = 28.0 * pts + 1000.0 * E * T * D
        entry = {"workload": "BASELINE.json configs[%d]: %d E x %d theta x %d phi x %d d x 4 pol, Si(111)(1x1):H"
                             % (2 if tag == "config3" else 3, E, T, P, D), "points": pts}
        for sigma in (0.0, 5.0):
            run = lambda: grid.yield_grid_device(*d, out=out, consts=consts, sigma_out=sigma)
            for _ in range(3):
                run()
            torch.cuda.synchronize()
            ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(9)]
            for a, b in ev:
                a.record()
                run()
                b.record()
            torch.cuda.synchronize()
            ms = float(np.median([a.elapsed_time(b) for a, b in ev]))
            r = {"ms": ms, "points_per_s": pts / (ms * 1e-3), "store_gbs": 8e-9 * pts / (ms * 1e-3),
                 "alg_fp64_tflops": flops / (ms * 1e-3) / 1e12}
            if tag == "config3":
                r["bound"], r["frac"] = "hbm", r["store_gbs"] / peak
            else:
                r["bound"], r["frac"] = "fp64", (r["alg_fp64_tflops"] / fp64_peak) if fp64_peak else None
            entry["sigma_out_%g" % sigma] = r
        if tag == "config3":
            # the same map through the DROP-IN call, written the reference's way -- shg.shgyield(energy[E], theta[T, 1],
            # phi[P, 1, 1]), host arrays in, four fresh NumPy arrays out, sigma_out over every axis like the reference --
            # wall clock, best of 3 after one warm-up call (sshg_yield_broadcast + threaded staged copies)
            kw = dict(eps_m1=m1, eps_m2=m2, eps_m3=m3, chi2=chi2, thick=float(np.ravel(axes["thick"])[0]),
                      gamma=axes["gamma"])
            for sigma in (0.0, 5.0):
                call = lambda: shg.shgyield(energy=axes["energy"], theta=axes["theta"][:, None],
                                            phi=axes["phi"][:, None, None], sigma_out=sigma, **kw)
                call()
                best = None
                for _ in range(3):
                    t0 = time.perf_counter()
                    r = call()
                    dt = time.perf_counter() - t0
                    best = dt if best is None else min(best, dt)
                    del r
                entry["through_shgyield_sigma_out_%g" % sigma] = {
                    "s_per_call": best, "points_per_s": pts / best, "result_gbs": 8e-9 * pts / best,
                    "how": "host arrays in, four fresh NumPy arrays [P, T, E] out"}
        res[tag] = entry
        del out
    torch.cuda.empty_cache()
    return res


def run_ours(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("launch with torch.distributed.run --nproc-per-node %d for --gpus %d" % (args.gpus, args.gpus))
        raise SystemExit("WORLD_SIZE %d != --gpus %d" % (world, args.gpus))

    # CPU baseline first: the fork pool must exist before CUDA is initialised in this process
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cpu = cpu_baseline(args.cpu_seconds)

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (there is no CPU fallback)")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    from shgyield_b200 import _lib, grid
    from shgyield_b200.shg import physical_constants
    from shgyield_b200.synthetic import config5

    cfg = config5(N_ENERGY, N_THETA, N_PHI)
    consts = physical_constants()
    ctx = _lib.context(local)
    t_lo, t_hi = grid.shard_bounds(N_THETA, world, rank)
    put = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
    d = {k: put(v) for k, v in cfg.items()}
    out = torch.empty((t_hi - t_lo, 1, 4, N_PHI, N_ENERGY), dtype=torch.float64, device=dev)
    my_pts = out.numel()
    total_pts = N_THETA * 4 * N_PHI * N_ENERGY

    def step():
        grid.yield_grid_device(d["energy"], d["theta"], d["phi"], d["thick"], d["eps1w"], d["eps2w"], d["chi2"],
                               out=out, consts=consts, t_range=(t_lo, t_hi))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    # roofline denominators measured in the same run (rank 0 only; before the timed region)
    peaks = {}
    if rank == 0:
        import ctypes as C
        v = C.c_double()
        _lib.check(ctx.lib.sshg_peak_store(ctx.handle, 8 << 30, 5, C.byref(v)), ctx.lib)
        peaks["store_gbs"] = v.value
        _lib.check(ctx.lib.sshg_peak_copy(ctx.handle, 4 << 30, 5, C.byref(v)), ctx.lib)
        peaks["copy_bulk_gbs"] = v.value
        _lib.check(ctx.lib.sshg_peak_dfma(ctx.handle, 5, C.byref(v)), ctx.lib)
        peaks["fp64_tflops"] = v.value

    # ---- configs 3 and 4 (rank 0 of a single-GPU run only; outside -- and before -- the timed region: a few short
    # launches, so that they are not measured on a GPU that the 83.5 GB launches have just driven into its power cap)
    configs = None
    if rank == 0 and world == 1 and not args.no_configs:
        peak_hbm, _ = measured_peaks()
        configs = run_configs(ctx, dev, consts, peak_hbm, peaks.get("fp64_tflops"))

    for _ in range(max(args.warmup, 0)):
        step()
    barrier()
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    launches0 = ctx.launches
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t_start, t_end = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    t_start.record()
    for a, b in ev:
        a.record()
        step()
        b.record()
    t_end.record()
    barrier()
    clocks = sampler.stop() if rank == 0 else None
    launches = ctx.launches - launches0
    total_ms = t_start.elapsed_time(t_end)
    step_ms = [a.elapsed_time(b) for a, b in ev]

    # ---- end to end: host buffers through the C ABI (H2D spectra + kernel + D2H yields), per step ----
    n_e2e = min(E2E_THETAS, t_hi - t_lo)
    e2e_range = (t_lo, t_lo + n_e2e)
    host_out = _lib.pinned_empty((n_e2e, 1, 4, N_PHI, N_ENERGY))
    host_in = {k: np.ascontiguousarray(v) for k, v in cfg.items()}

    def e2e_step():
        grid.yield_grid_host(host_in["energy"], host_in["theta"], host_in["phi"], host_in["thick"], host_in["eps1w"],
                             host_in["eps2w"], host_in["chi2"], consts=consts, t_range=e2e_range, out=host_out)

    for _ in range(2):
        e2e_step()
    barrier()
    e2e_steps = max(3, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        e2e_step()
    barrier()
    e2e_s = time.perf_counter() - t0
    e2e_check = float(host_out[0, 0, :, 0, :].sum())
    # the ceiling of the end-to-end leg: a plain device-to-pinned-host copy of the same bytes (PCIe), measured with
    # the protocol of the leg itself -- every rank copies e2e_steps times between two barriers, wall clock, slowest
    # rank -- plus the best single copy of this rank
    probe = torch.empty(host_out.size, dtype=torch.float64, device=dev)
    pin_t = torch.from_numpy(host_out.reshape(-1))
    pin_t.copy_(probe, non_blocking=True)
    barrier()
    pcie_ms = []
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        pin_t.copy_(probe, non_blocking=True)
        c1.record()
        torch.cuda.synchronize()
        pcie_ms.append(c0.elapsed_time(c1))
    barrier()
    pcie_s = time.perf_counter() - t0
    pcie_gbs_best = host_out.nbytes / (min(pcie_ms) * 1e-3) / 1e9
    del probe
    h2d = sum(int(v.nbytes) for v in host_in.values())
    d2h = int(host_out.nbytes)                    # bytes DELIVERED into the host buffer per step
    # bytes that cross PCIe: the sS rows of a (phi, phi + 180 deg) pair are the same bits, the library copies them once
    # and duplicates them on the host (option pair_copy; 360 of the 2884 rows of a config-5 column)
    phi_h = host_in["phi"]
    h_pairs = phi_h.size // 2 if np.array_equal(phi_h[phi_h.size // 2:2 * (phi_h.size // 2)], phi_h[:phi_h.size // 2] + 180.0) else 0
    d2h_pcie = d2h * (4 * N_PHI - h_pairs) // (4 * N_PHI)
    e2e_pts = n_e2e * 4 * N_PHI * N_ENERGY
    # the same call with an ordinary (pageable, never touched) NumPy array per step as the result -- what a caller gets
    # who does not ask for page-locked memory: the library copies through page-locked staging buffers with several host
    # threads (sshg_copy_out).  Single-GPU runs only, 8 thetas per step.
    e2e_pageable = None
    if rank == 0 and world == 1:
        n_pg = min(8, n_e2e)
        pg_times = []
        for _ in range(3):
            fresh = np.empty((n_pg, 1, 4, N_PHI, N_ENERGY))
            t0 = time.perf_counter()
            grid.yield_grid_host(host_in["energy"], host_in["theta"], host_in["phi"], host_in["thick"], host_in["eps1w"],
                                 host_in["eps2w"], host_in["chi2"], consts=consts, t_range=(t_lo, t_lo + n_pg), out=fresh)
            pg_times.append(time.perf_counter() - t0)
            del fresh
        pg_pts = n_pg * 4 * N_PHI * N_ENERGY
        e2e_pageable = {"value": pg_pts / min(pg_times[1:]), "unit": "points/s", "d2h_gbs": 8e-9 * pg_pts / min(pg_times[1:]),
                        "sample": "%d thetas per call into a fresh np.empty array (first-touch page faults included), "
                                  "best of 2 after one warm-up call" % n_pg}

    # ---- the one collective of the path: gather of the slabs onto rank 0 (reported separately) --------
    gather = None
    if world > 1 and not args.no_gather:
        full = torch.empty((N_THETA, 1, 4, N_PHI, N_ENERGY), dtype=torch.float64, device=dev) if rank == 0 else None
        g_ms = []
        for it in range(3):
            barrier()
            g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            g0.record()
            grid.gather_slabs(out, N_THETA, 1, (rank, world, "theta"), dst=0, out=full)
            g1.record()
            barrier()
            if it:
                g_ms.append(g0.elapsed_time(g1))
        gt = torch.tensor([min(g_ms)], dtype=torch.float64, device=dev)
        dist.all_reduce(gt, op=dist.ReduceOp.MAX)
        recv_bytes = 8.0 * (total_pts - my_pts) if rank == 0 else 0.0
        if rank == 0:
            ok = bool(torch.equal(full[t_lo:t_hi], out))
            gather = {"ms": float(gt.item()), "bytes_into_rank0": recv_bytes, "gbs": recv_bytes / (gt.item() * 1e-3) / 1e9,
                      "how": "grouped ncclSend/ncclRecv of the true shard sizes into the cube on rank 0 (NVLink)",
                      "own_slab_intact": ok}
        del full
        torch.cuda.empty_cache()
        # the same gather WITHOUT a collective: every rank's K2 stores its slab straight into rank 0's cube
        # over NVLink (CUDA IPC mapping, grid.PeerCube) -- compute and gather are one kernel per rank
        try:
            if args.no_peer_gather:
                raise RuntimeError("skipped (--no-peer-gather)")
            pc = grid.PeerCube((N_THETA, 1, 4, N_PHI, N_ENERGY), dst=0)      # raises on EVERY rank or on none
            p_ms = []
            for it in range(3):
                barrier()
                g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                g0.record()
                grid.yield_grid_device(d["energy"], d["theta"], d["phi"], d["thick"], d["eps1w"], d["eps2w"], d["chi2"],
                                       consts=consts, t_range=(t_lo, t_hi), out_ptr=pc.slab_ptr(t_lo))
                g1.record()
                barrier()
                if it:
                    p_ms.append(g0.elapsed_time(g1))
            pt = torch.tensor([min(p_ms)], dtype=torch.float64, device=dev)
            dist.all_reduce(pt, op=dist.ReduceOp.MAX)
            if rank == 0:
                cube = pc.tensor()
                # re-compute the end of the last rank's slab here (at most 12 thetas: rank 0 already holds its slab and
                # the 83.5 GB cube) and compare bit for bit; rank 0's own slab against the one of the timed region
                lo_l, hi_l = grid.shard_bounds(N_THETA, world, world - 1)
                lo_l = max(lo_l, hi_l - 12)
                chk = grid.yield_grid_device(d["energy"], d["theta"], d["phi"], d["thick"], d["eps1w"], d["eps2w"],
                                             d["chi2"], consts=consts, t_range=(lo_l, hi_l))
                same = bool(torch.equal(cube[lo_l:hi_l], chk)) and bool(torch.equal(cube[t_lo:t_hi], out))
                gather["peer_store"] = {
                    "ms_compute_and_gather": float(pt.item()), "gbs_into_rank0": recv_bytes / (pt.item() * 1e-3) / 1e9,
                    "vs_compute_then_nccl_gather_ms": total_ms / args.steps + gather["ms"],
                    "how": "every rank's yield_grid_kernel stores its theta slab directly into rank 0's cube (cudaIpc "
                           "mapping over NVLink/NVSwitch): no slab buffer, no send/recv; max over ranks of the kernel time",
                    "bitwise_equal_to_local": same}
                del cube, chk
            pc.close()
            del pc
        except Exception as e:                                   # IPC not permitted on this box: report, keep the line
            if rank == 0 and gather is not None:
                gather["peer_store"] = {"unavailable": repr(e)[:200]}
        torch.cuda.empty_cache()

    # ---- secondary: the same shard with the reference's default output broadening (sigma_out = 5), K3 ---
    def blur_step():
        grid.yield_grid_device(d["energy"], d["theta"], d["phi"], d["thick"], d["eps1w"], d["eps2w"], d["chi2"],
                               out=out, consts=consts, t_range=(t_lo, t_hi), sigma_out=5.0)

    # the GPUs idled through the host-bound legs above: warm up for ~0.2 s so that the clocks are back up
    t_w = time.perf_counter()
    while time.perf_counter() - t_w < 0.2:
        for _ in range(3):
            blur_step()
        torch.cuda.synchronize()
    b_ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(5)]
    barrier()
    for a, b in b_ev:
        a.record()
        blur_step()
        b.record()
    barrier()
    blur_ms = float(np.mean([a.elapsed_time(b) for a, b in b_ev]))

    # ---- reduce over ranks (max time; sums of points) ----------------------------------------------
    stats = torch.tensor([total_ms, e2e_s, float(np.mean(step_ms)), blur_ms, pcie_s], dtype=torch.float64, device=dev)
    counts = torch.tensor([float(my_pts), float(e2e_pts), float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.MAX)
        dist.all_reduce(counts, op=dist.ReduceOp.SUM)
    total_ms, e2e_s, mean_step_ms, blur_ms_max, pcie_s = (float(x) for x in stats.cpu())
    sum_pts, sum_e2e_pts, sum_launches = (float(x) for x in counts.cpu())
    assert int(sum_pts) == total_pts

    if rank == 0:
        peak, peak_src = measured_peaks()
        traffic = k2_traffic()
        value = total_pts * args.steps / (total_ms * 1e-3)
        # the same metric when the cube has to end up on ONE GPU (north_star's gather over NVLink): the faster of
        # "every rank's kernel stores into rank 0's cube" and "compute, then grouped ncclSend/ncclRecv"
        value_with_gather = None
        if gather is not None:
            ways = {"compute_then_nccl_gather": total_ms / args.steps + gather["ms"]}
            ps = gather.get("peer_store") or {}
            if "ms_compute_and_gather" in ps:
                ways["peer_stores_from_the_kernel"] = ps["ms_compute_and_gather"]
            how = min(ways, key=ways.get)
            value_with_gather = {"value": total_pts / (ways[how] * 1e-3), "unit": "points/s", "ms_per_step": ways[how],
                                 "how": how, "all_ms": ways,
                                 "bound": "rank 0's NVLink ingress: %.1f GB at %.0f GB/s"
                                          % (gather["bytes_into_rank0"] / 1e9,
                                             gather["bytes_into_rank0"] / (ways[how] * 1e-3) / 1e9)}
        k2_ms = float(np.mean(step_ms))                      # this rank's K2 launch (+ the 1-CTA phi-table kernel)
        achieved = 8.0 * my_pts / (k2_ms * 1e-3) / 1e9
        flops = 28.0 * my_pts + 1000.0 * (t_hi - t_lo) * N_ENERGY
        line = {
            "metric": METRIC, "value": value, "unit": "points/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": workload_config(),
            "run": {"shards": "theta shards x%d" % world, "bytes_written_per_gpu_per_step": 8.0 * my_pts},
            "e2e": {"pageable_result": e2e_pageable, "value": sum_e2e_pts * e2e_steps / e2e_s, "unit": "points/s",
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h_pcie, "result_bytes_per_step": d2h,
                    "d2h_note": "the plain copy moves result_bytes_per_step; the call moves d2h_bytes_per_step over PCIe -- "
                                "the second sS row of every (phi, phi + 180 deg) pair is the same bits as the first and is "
                                "duplicated by host threads while the next slab is on the wire -- so d2h_gbs_achieved "
                                "(delivered bytes per second) can exceed the plain copy",
                    "sample": "sshg_yield with host buffers on the first %d thetas of each rank's shard "
                              "(%.3g points/step/GPU), pinned result buffer, %d steps" % (n_e2e, e2e_pts, e2e_steps),
                    "d2h_gbs_achieved": d2h * e2e_steps / e2e_s / 1e9,
                    "d2h_gbs_plain_copy": d2h * e2e_steps / pcie_s / 1e9,
                    "d2h_gbs_plain_copy_best_single": pcie_gbs_best,
                    "d2h_protocol": "per GPU; both figures: every rank moves the same bytes the same number of times "
                                    "between two barriers, wall clock of the slowest rank",
                    "bound": "PCIe device-to-host copy of the FP64 yields (rank 0's link; all ranks copy concurrently)",
                    "checksum": e2e_check},
            "gpu_launches": int(sum_launches),
            "clocks": clocks,
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         "traffic": (traffic[0] / traffic[1]) * 8.0 * my_pts if traffic else None,
                         "traffic_source": ("ncu dram__bytes_read.sum + dram__bytes_write.sum = %.4f GB for %.4f GB algorithmic "
                                            "(ratio %.4f) on the launch captured in %s, scaled to this launch's algorithmic "
                                            "bytes (the 83.5 GB launch itself is not captured under ncu)"
                                            % (traffic[0] / 1e9, traffic[1] / 1e9, traffic[0] / traffic[1], traffic[2]))
                         if traffic else "no ncu summary found under profiles/",
                         "peak_source": peak_src,
                         "kernel": "yield_grid_kernel (K2); algorithmic bytes = 8 B x %.4g points per launch; "
                                   "duration = mean CUDA-event time of the %d timed launches on rank 0 (%.3f ms)"
                                   % (my_pts, args.steps, k2_ms),
                         "store_peak_gbs_measured": peaks.get("store_gbs"),
                         "copy_peak_bulk_gbs_measured": peaks.get("copy_bulk_gbs"),
                         "fp64_peak_tflops_measured": peaks.get("fp64_tflops"),
                         "alg_fp64_tflops": flops / (k2_ms * 1e-3) / 1e12},
            "cpu_baseline": cpu,
            "configs": configs,
            "gather": gather,
            "value_with_gather": value_with_gather,
            "broadened": {"sigma_out": 5.0, "ms_per_step": blur_ms_max, "value": total_pts / (blur_ms_max * 1e-3),
                          "unit": "points/s", "hbm_frac_rank0": 8.0 * my_pts / (blur_ms * 1e-3) / 1e9 / peak,
                          "how": "the same cube with shgyield()'s default output broadening (sigma_out = 5 samples "
                                 "along E) in one pass (K3: FIR on the 13 azimuthal harmonics of each column, "
                                 "yield_blur_kernel, which also re-evaluates the rows next to an azimuthal node "
                                 "row-wise), device resident, mean of 5 launches after 0.2 s of warm-up, max over ranks"},
            "ms_per_step_device_max": mean_step_ms,
        }
        print(json.dumps(line), flush=True)
    _lib.pinned_free(host_out)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", choices=["ours", "reference"], default="ours")
    ap.add_argument("--cpu-seconds", type=float, default=12.0)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[2] / configs[3] extra (N = 1)")
    ap.add_argument("--no-gather", action="store_true", help="skip timing the NCCL gather of the slabs (N > 1)")
    ap.add_argument("--no-peer-gather", action="store_true", help="skip the gather by peer stores (N > 1)")
    args = ap.parse_args()
    if args.steps < 1:
        raise SystemExit("--steps must be >= 1")
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
