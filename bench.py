#!/usr/bin/env python
"""bench.py — FAVA grid-statistics hot path on B200 (contract: task statement; numbers explained in DESIGN.md §5).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME] [--no-extras]

One "step" = one pass of the hot path over one synthetic snapshot resident in HBM: Reynolds + Favre plane
profiles along x, y and z, plus the kinetic-energy spectrum.  Under torchrun (N > 1) the SAME global grid is
split into z-slabs, one per rank ("strong" scaling).  Rank 0 prints ONE JSON line.  After the timed region every
run (any N) pushes two closed-form snapshots through the same code path and checks the results (`parity_check`);
a failed check exits non-zero.
"""

from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "Gcells/s & %HBM roofline: Reynolds/Favre profiles + KE spectrum @1024^3, 1-8 GPU"
UNIT = "Gcells/s"

WORKLOADS = {
    "full1024": dict(n=1024, spectrum=True, cpu_n=256,
                     desc="1024^3 uniform fp64: Reynolds + Favre stress profiles along x/y/z + kinetic_energy_spectra "
                          "(BASELINE configs[3]; z-slab decomposed for N>1)"),
    "full512": dict(n=512, spectrum=True, cpu_n=192, desc="512^3 uniform fp64: profiles x/y/z + kinetic_energy_spectra"),
    "full256": dict(n=256, spectrum=True, cpu_n=128, desc="256^3 uniform fp64: profiles x/y/z + kinetic_energy_spectra (debug)"),
    "profiles512": dict(n=512, spectrum=False, cpu_n=256,
                        desc="512^3 uniform fp64 Reynolds + Favre stress profiles along x/y/z (BASELINE configs[2])"),
    "profiles1024": dict(n=1024, spectrum=False, cpu_n=256, desc="1024^3 uniform fp64 Reynolds + Favre profiles x/y/z"),
}
DEFAULT_WORKLOAD = "full1024"
AXES = (0, 1, 2)
PARITY_TOL = 1e-12  # BASELINE.json north_star: fp64 profiles and spectra within 1e-12 (max-norm per array)

# algorithmic bytes per cell, fp64 input (SURVEY §8d / DESIGN.md §4)
B_PROFILE = 32.0    # rho, ux, uy, uz read once
B_SPECTRUM = 200.0  # 3 separable line passes in Hermitian storage, weighting fused, binning incl. transposed operand
B_STEP = 232.0      # profiles (one read) + spectrum
B_XPASS = 32.0 + 24.0  # read 4 fields, write 3 half spectra (8 B/cell each)
B_COLPASS = 16.0    # one component: half spectrum read + written (8 + 8 B/cell)


def bin_algorithmic_bytes(n: int) -> float:
    """K6: every element inside the spectral sphere |k| <= n/2 - 1.5 of the stored half space read ONCE (3 components x
    16 B); the transposed operand of a point is another point's direct operand."""
    r = n / 2.0 - 1.5
    return 48.0 * 0.5 * (4.0 / 3.0) * math.pi * r**3


def measured_peak_gbs() -> tuple[float, str]:
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region."""

    QUERY = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
             "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines: list[str] = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.QUERY}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.proc = None
            return

        def pump():
            for line in self.proc.stdout:
                self.lines.append(line.strip())

        threading.Thread(target=pump, daemon=True).start()
        time.sleep(0.25)  # first sample lands before the timed region starts

    def stop(self) -> dict:
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in self.lines:
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                smax.append(float(f[2]))
                power.append(float(f[3]))
            except ValueError:
                continue
            for name, val in zip(names, f[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        busy = [s for s, p in zip(sm, power) if p > 0.5 * max(power)] if power else sm
        return {
            "sm_mhz": float(np.median(busy)) if busy else None,
            "sm_max_mhz": float(max(smax)) if smax else None,
            "power_w_max": float(max(power)) if power else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def synth_slab_device(n: int, z0: int, nz: int, dev, seed: int = 1234):
    """Synthetic snapshot slab generated directly in HBM (dens > 0, sheared velocities + noise)."""
    import torch

    g = torch.Generator(device=dev)
    g.manual_seed(seed + z0)
    shape = (nz, n, n)
    rho = 1.0 + 0.5 * torch.rand(shape, generator=g, device=dev, dtype=torch.float64)
    z = (torch.arange(z0, z0 + nz, device=dev, dtype=torch.float64) + 0.5).view(nz, 1, 1) / n
    y = (torch.arange(n, device=dev, dtype=torch.float64) + 0.5).view(1, n, 1) / n
    x = (torch.arange(n, device=dev, dtype=torch.float64) + 0.5).view(1, 1, n) / n
    two_pi = 2.0 * np.pi
    ux = torch.randn(shape, generator=g, device=dev, dtype=torch.float64).mul_(0.25).add_(torch.sin(two_pi * y))
    uy = torch.randn(shape, generator=g, device=dev, dtype=torch.float64).mul_(0.25).add_(torch.sin(2 * two_pi * z))
    uz = torch.randn(shape, generator=g, device=dev, dtype=torch.float64).mul_(0.25).add_(torch.sin(3 * two_pi * x))
    return rho, ux, uy, uz


class StageTimer:
    """CUDA-event brackets per pipeline stage, on the stream the kernels are launched on (torch's current)."""

    def __init__(self):
        self.pairs: dict[str, list] = {}

    def bracket(self, name: str):
        import torch

        timer = self

        class _Ctx:
            def __enter__(self_inner):
                self_inner.e0 = torch.cuda.Event(enable_timing=True)
                self_inner.e1 = torch.cuda.Event(enable_timing=True)
                self_inner.e0.record()

            def __exit__(self_inner, *exc):
                self_inner.e1.record()
                timer.pairs.setdefault(name, []).append((self_inner.e0, self_inner.e1))
                return False

        return _Ctx()

    def mean_ms(self) -> dict[str, float]:
        return {k: float(np.mean([a.elapsed_time(b) for a, b in v])) for k, v in self.pairs.items()}


KERNEL_NAMES = {
    "plane_moments_xyz": "k_moments_xyz (fava_plane_moments_xyz: x, y AND z profiles from one 32 B/cell read, TMA tensor tiles)",
    "plane_moments_xz": "k_moments_cols + k_partials_to_planes (fava_plane_moments_xz: x AND z profiles from one read)",
    "plane_moments_axis1": "k_moments_rows (fava_plane_moments, axis y)",
    "transform_x": "k_fft_x_row (fava_ke_transform_x: sqrt(rho) u weighting fused with the x transform; one row per CTA as a "
                   "half-length complex line in registers, rows fed by TMA bulk copies)",
    "transform_y": "k_fft_cols<.,1> (fava_ke_transform_y: in-place y transform, TMA tensor tiles, pruned outputs)",
    "transform_z": "k_fft_cols<.,2> (fava_ke_transform_z: in-place z transform, pruned to the spectral sphere)",
    "spectrum_bin": "k_spectrum_bin (fava_spectrum_bin)",
    "transform_y_exchange": "k_fft_cols<.,1,scatter> (fava_fft_y_scatter: y transform FUSED with the slab -> pencil exchange; "
                            "output rows stored straight into the owners' peer-mapped buffers over NVLink)",
}


def run_ours(args) -> dict:
    import torch

    from fava_b200 import device, dist, spectrum, stats
    from fava_b200.build import build_library

    build_library()  # no-op when the in-tree .so is current
    rank, world, local = dist.init_from_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    numa_bound = dist.bind_to_gpu_numa(local)  # pinned staging buffers next to the GPU (e2e, time series)
    dev = torch.device("cuda", local)
    wl = WORKLOADS[args.workload]
    n = wl["n"]
    if n % (2 * world):
        raise SystemExit(f"grid {n} not divisible by 2 x {world} ranks")
    if not device.fft_native_supported(n):
        raise SystemExit(f"bench workloads use power-of-two grids (hand-written transform path); got {n}")
    nz = n // world
    z0 = rank * nz
    fields = synth_slab_device(n, z0, nz, dev)
    cell_volume = 1.0 / float(n) ** 3
    layer_volume = 1.0 / float(n)
    ncells = float(n) ** 3
    peak, peak_src = measured_peak_gbs()

    def step(f):
        """The public per-step path: slab profiles (x,y,z) + slab spectrum.  Returns host-readable results."""
        return stats.slab_step(*f, n, cell_volume, layer_volume, axes=AXES, spectrum=wl["spectrum"], favre=True)

    def step_instrumented(timer: StageTimer):
        """Same kernels, stage by stage, with event brackets (used after the timed region)."""
        rho, ux, uy, uz = fields
        if device.plane_moments_xyz_supported(rho.shape):
            with timer.bracket("plane_moments_xyz"):  # x, y and z bins from ONE pass (fava_plane_moments_xyz)
                (mx, px), (my, py), (mz, pz) = device.plane_moments_xyz(rho, ux, uy, uz)
        else:
            with timer.bracket("plane_moments_xz"):  # x-bins and z-bins from one pass (fava_plane_moments_xz)
                (mx, px), (mz, pz) = device.plane_moments_xz(rho, ux, uy, uz)
            with timer.bracket("plane_moments_axis1"):
                my, py = device.plane_moments(rho, ux, uy, uz, 1)
        for ax, mom, piv in ((0, mx, px), (1, my, py), (2, mz, pz)):
            stats.slab_profiles_finish(mom, piv, ax, cell_volume, layer_volume, gather=False)
        if not wl["spectrum"]:
            return
        if world == 1:
            w = spectrum.spectral_buffers(n, n, dev)
            sums = torch.zeros((3, n // 2 - 1), dtype=torch.float64, device=dev)
            with timer.bracket("transform_x"):
                device.ke_transform_x(rho, ux, uy, uz, *w)
            for c in range(3):
                with timer.bracket("transform_y"):
                    device.ke_transform_y(w[c], n, n, dev)
            for c in range(3):
                with timer.bracket("transform_z"):
                    device.ke_transform_z(w[c], n, n, None, dev)
            with timer.bracket("spectrum_bin"):
                device.spectrum_bin(w[0], w[1], w[2], n, n, None, sums)
        else:
            p = spectrum._plan(n, rank, world, dev)
            with timer.bracket("transform_x"):
                device.ke_transform_x(rho, ux, uy, uz, *p.send)
            for c in range(3):
                with timer.bracket("transform_y_exchange"):  # ONE kernel: y pass whose rows go to their owners over NVLink
                    spectrum.exchange(p, c)
            dist.allreduce_sum_(p.tokens[0])
            for c in range(3):
                with timer.bracket("transform_z"):
                    device.ke_transform_z(p.recv[c], n, p.nyl, p.ky_of_local, dev)
            with timer.bracket("spectrum_bin"):
                device.spectrum_bin(p.recv[0], p.recv[1], p.recv[2], n, p.nyl, p.ky_of_local, p.sums)
            dist.allreduce_sum_(p.sums)

    for _ in range(args.warmup):
        step(fields)
    torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    if sampler:
        sampler.start()
    launches0 = device.launch_count()
    dist.barrier()
    torch.cuda.synchronize()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        step(fields)
    t1.record()
    torch.cuda.synchronize()
    dist.barrier()
    launches = device.launch_count() - launches0
    clocks = sampler.stop() if sampler else None
    elapsed_ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    dist.allreduce_max_(elapsed_ms)
    ms_per_step = float(elapsed_ms.item()) / args.steps
    value = ncells / (ms_per_step * 1e-3) / 1e9

    # ---- per-stage device times (same kernels, event-bracketed, after the timed region) -------------
    timer = StageTimer()
    for _ in range(2):
        step_instrumented(timer)
    torch.cuda.synchronize()
    stage_ms = timer.mean_ms()
    local_cells = ncells / world
    algo = {"plane_moments_xyz": B_PROFILE * local_cells, "plane_moments_xz": B_PROFILE * local_cells, "plane_moments_axis1": B_PROFILE * local_cells,
            "transform_x": B_XPASS * local_cells, "transform_y": B_COLPASS * local_cells,
            "transform_z": B_COLPASS * local_cells, "spectrum_bin": bin_algorithmic_bytes(n) / world,
            "transform_y_exchange": B_COLPASS * local_cells}
    stages = {}
    for name, ms in stage_ms.items():
        st = {"ms": ms, "launches_per_step": len(timer.pairs[name]) // 2, "algorithmic_bytes": algo[name]}
        st["achieved_gbs"] = algo[name] / (ms * 1e-3) / 1e9
        st["frac_of_hbm_peak"] = st["achieved_gbs"] / peak
        stages[name] = st
    if "transform_y_exchange" in stages:  # NVLink view: nominal = every column, on-wire = inside the spectral disc
        a = stages["transform_y_exchange"]
        nominal = 8.0 * local_cells * (world - 1) / world
        a["nvlink_nominal_gbs_per_gpu"] = nominal / (a["ms"] * 1e-3) / 1e9
        a["nvlink_onwire_gbs_per_gpu"] = 0.7854 * a["nvlink_nominal_gbs_per_gpu"]
        a["frac_of_nvlink_770_onwire"] = a["nvlink_onwire_gbs_per_gpu"] / 770.0
        a["ctas"] = spectrum.exchange_ctas(world)
        a["note"] = ("NVLink-bound for N >= 4: (N-1)/N of the output rows leave the GPU, pi/4 of the columns lie inside the "
                     "spectral disc and are sent; in the real step it runs beside the other kernels on `ctas` SMs")

    own = {k: v for k, v in stages.items() if k != "transform_y_exchange"}
    dom = max(own, key=lambda k: own[k]["ms"] * own[k]["launches_per_step"])
    traffic, traffic_note = load_profile_traffic(dom, world)
    roofline = {
        "kernel": KERNEL_NAMES.get(dom, dom),
        "bound": "hbm",
        "achieved": own[dom]["achieved_gbs"],
        "peak": peak,
        "unit": "GB/s",
        "frac": own[dom]["achieved_gbs"] / peak,
        "traffic": traffic,
        "traffic_note": traffic_note,
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": own[dom]["algorithmic_bytes"],
        "kernel_ms": own[dom]["ms"],
        "note": "dominant kernel by time per step (ms x launches); every stage is listed under roofline_stages; all "
                "of them are hand-written (no library call on the power-of-two path)",
    }
    prof_ms = sum(v["ms"] for k, v in stages.items() if k.startswith("plane_moments"))
    npass = sum(1 for k in stages if k.startswith("plane_moments"))
    summary = {"profiles_xyz": {"ms": prof_ms, "passes_over_the_fields": npass}}
    for label, bpc in (("bytes_read_%dB" % int(npass * B_PROFILE), npass * B_PROFILE), ("one_read_minimum_32B", B_PROFILE)):
        g = bpc * local_cells / (prof_ms * 1e-3) / 1e9
        summary["profiles_xyz"][label] = {"model_bytes_per_cell": bpc, "achieved_gbs": g, "frac_of_hbm_peak": g / peak}
    if wl["spectrum"]:
        spec_ms = sum(v["ms"] * v["launches_per_step"] for k, v in stages.items() if not k.startswith("plane_moments"))
        g = B_SPECTRUM * local_cells / (spec_ms * 1e-3) / 1e9
        summary["ke_spectrum"] = {"ms": spec_ms, "model_bytes_per_cell": B_SPECTRUM, "achieved_gbs": g,
                                  "frac_of_hbm_peak": g / peak,
                                  "note": "sum of the stage times (exchange included for N>1, where it overlaps in the real step)"}
        g = B_STEP * local_cells / (ms_per_step * 1e-3) / 1e9
        summary["step"] = {"ms": ms_per_step, "model_bytes_per_cell": B_STEP, "achieved_gbs": g, "frac_of_hbm_peak": g / peak,
                           "note": "timed region: profiles x/y/z + spectrum, per rank"}

    timeline = None
    if world > 1 and wl["spectrum"]:  # where one public step spends its time (events on all streams, rank 0)
        p = spectrum._plan(n, rank, world, dev)
        s0 = torch.cuda.Event(enable_timing=True)
        s1 = torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        s0.record()
        step(fields)
        s1.record()
        torch.cuda.synchronize()
        timeline = {"x_done_ms": s0.elapsed_time(p.ev_xy[0]), "y_exchange_kernel_done_ms": [s0.elapsed_time(e) for e in p.ev_packed],
                    "exchange_done_ms": [s0.elapsed_time(e) for e in p.ev_done],
                    "moments_done_ms": s0.elapsed_time(p.ev_mark["overlap"]), "fft_z_done_ms": s0.elapsed_time(p.ev_mark["fft_z"]),
                    "bin_done_ms": s0.elapsed_time(p.ev_mark["bin"]), "step_done_ms": s0.elapsed_time(s1)}
        wait_z = timeline["exchange_done_ms"][2] - timeline["moments_done_ms"]
        timeline["limiter"] = ("exchange (the z transforms wait %.2f ms for the last component's rows)" % wait_z if wait_z > 0.2
                               else "HBM-bound kernels (the exchange is hidden behind them)")

    parity = parity_check(n, rank, world, dev, fields, z0, nz, cell_volume, layer_volume, wl["spectrum"])

    def host_step(host, stage):
        return stats.host_step(host, n, cell_volume, layer_volume, axes=AXES, spectrum=wl["spectrum"], favre=True, stage=stage)

    fields = None  # the parity check overwrote the snapshot
    torch.cuda.empty_cache()
    fields = synth_slab_device(n, z0, nz, dev)
    e2e = run_e2e(args, wl, dev, rank, world, fields, host_step, parity, z0)
    e2e["numa_bound"] = bool(numa_bound)

    out = {
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": world,
        "steps": args.steps,
        "warmup": args.warmup,
        "ms_per_step": ms_per_step,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {
            "workload": wl["desc"],
            "grid": [n, n, n],
            "parallelism": f"z-slabs x{world}",
            "l2": "inputs (4 fields x %.2f GiB per GPU) exceed the 126 MB L2; no flush needed" % (8.0 * nz * n * n / 2**30),
        },
        "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": roofline,
        "roofline_stages": stages,
        "roofline_summary": summary,
        "timeline": timeline,
        "parity_check": parity,
        "clocks": clocks,
    }
    if not args.no_extras:
        del fields
        torch.cuda.empty_cache()
        out["extras"] = run_extras(rank, world, dev, peak)
    if rank == 0 and world == 1:
        out["cpu_baseline"] = cpu_baseline(wl)
    if not parity["passed"]:
        out["error"] = "parity_check failed"
    return out if rank == 0 else {"_failed": not parity["passed"]}


def parity_check(n, rank, world, dev, fields, z0, nz, cell_volume, layer_volume, with_spectrum) -> dict:
    """Known answers through the SAME calls the timed region makes (stats.slab_step on this rank's slab, the same plan,
    buffers and streams): the separable profile snapshot and the three-mode spectrum snapshot of fava_b200/knownanswer.py
    (closed forms pinned on the oracle at 32^3, tests/test_knownanswer_cpu.py), plus bitwise equality of two consecutive
    steps.  Collective: every rank checks its share; the verdict is the max over ranks."""
    import torch

    from fava_b200 import dist, knownanswer as ka, stats

    errs: dict[str, float] = {}
    cases = []

    def run():
        return stats.slab_step(*fields, n, cell_volume, layer_volume, axes=AXES, spectrum=with_spectrum, favre=True)

    ka.fill_profile_case(fields, n, z0)
    res = run()
    errs.update(ka.profile_errors(res, n, AXES, z0, nz))
    cases.append("profiles x/y/z of rho=1+b(y)/4, ux=3+a(x)+b(y), uy=a(z), uz=-2 (means, Reynolds, Favre: closed forms)")
    bitwise = True
    if with_spectrum and n >= 28:
        ka.fill_spectrum_case(fields, n, z0)
        res1 = run()
        errs.update(ka.spectrum_errors(res1["spectrum"], n))
        cases.append("spectrum of three Fourier modes |k| = 7, 5, 12 (total per shell from lattice-point counts; "
                     "transverse = total - longitudinal)")
        res2 = run()
        bitwise = all(np.array_equal(res1["spectrum"][k], res2["spectrum"][k], equal_nan=True) for k in res1["spectrum"])
        for ax in AXES:
            bitwise = bitwise and all(torch.equal(res1[ax][k], res2[ax][k]) for k in res1[ax])
        cases.append("two consecutive steps bitwise equal")
    worst = max(errs.values()) if errs else 0.0
    t = torch.tensor([worst, 0.0 if bitwise else 1.0], dtype=torch.float64, device=dev)
    dist.allreduce_max_(t)
    worst, bitwise = float(t[0].item()), bool(t[1].item() == 0.0)
    where = max(errs, key=errs.get) if errs else None
    return {"checked": True, "passed": bool(worst <= PARITY_TOL and bitwise), "max_rel_err": worst, "tolerance": PARITY_TOL,
            "worst_on_rank0": where, "bitwise_repeatable": bitwise, "ranks": world, "grid": n, "through": "stats.slab_step",
            "cases": cases}


def run_e2e(args, wl, dev, rank, world, dev_fields, host_step, parity, z0) -> dict:
    """Same step through HOST buffers: pinned host fields -> H2D (in the timed region) -> public step ->
    profiles and spectrum read back to the host.  Also: the known-answer snapshots through the same host path, and
    the H2D ceiling of this box (every rank copying its slab alone, nothing else running)."""
    import torch

    from fava_b200 import dist, knownanswer as ka

    n = wl["n"]
    nz = int(dev_fields[0].shape[0])
    host = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in dev_fields]
    stage = [torch.empty_like(t) for t in dev_fields]

    # ---- known answers through the host path (adds to parity_check) ----------------------------------------------
    herrs = {}
    ka.fill_profile_case(stage, n, z0)
    for h, d in zip(host, stage):
        h.copy_(d)
    torch.cuda.synchronize()
    res = host_step(host, stage)
    herrs.update(ka.profile_errors(res, n, AXES, z0, nz))
    if wl["spectrum"] and n >= 28:
        ka.fill_spectrum_case(stage, n, z0)
        for h, d in zip(host, stage):
            h.copy_(d)
        torch.cuda.synchronize()
        herrs.update(ka.spectrum_errors(host_step(host, stage)["spectrum"], n))
    t = torch.tensor([max(herrs.values())], dtype=torch.float64, device=dev)
    dist.allreduce_max_(t)
    parity["host_step_max_rel_err"] = float(t.item())
    parity["passed"] = bool(parity["passed"] and parity["host_step_max_rel_err"] <= PARITY_TOL)
    parity["through"] = "stats.slab_step and stats.host_step"

    for h, d in zip(host, dev_fields):
        h.copy_(d)
    torch.cuda.synchronize()
    h2d = sum(h.numel() * h.element_size() for h in host)
    d2h = {"bytes": 0}

    def e2e_step():
        # the user-facing call for host-resident snapshots: chunked H2D on a side stream, each chunk consumed as it lands
        res = host_step(host, stage)
        nbytes = 0
        for key, val in res.items():
            if key == "spectrum":
                nbytes += sum(v.nbytes for v in val.values())  # already on the host (fava_spectrum_finalize)
            else:
                hostv = {k: v.cpu() for k, v in val.items()}
                nbytes += sum(v.numel() * v.element_size() for v in hostv.values())
        d2h["bytes"] = nbytes

    e2e_step()
    steps = max(1, min(args.steps, 3))
    dist.barrier()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True)
    e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        e2e_step()
    e1.record()
    torch.cuda.synchronize()
    dist.barrier()
    dt = torch.tensor([e0.elapsed_time(e1)], dtype=torch.float64, device=dev)
    dist.allreduce_max_(dt)
    ms = float(dt.item()) / steps

    # ---- H2D ceiling: the same pinned buffers, the same chunking, copies only, all ranks at once --------------------
    chunk = 64
    dist.barrier()
    torch.cuda.synchronize()
    c0 = torch.cuda.Event(enable_timing=True)
    c1 = torch.cuda.Event(enable_timing=True)
    reps = 2
    c0.record()
    for _ in range(reps):
        for a in range(0, nz, chunk):
            for h, s in zip(host, stage):
                s[a:a + chunk].copy_(h[a:a + chunk], non_blocking=True)
    c1.record()
    torch.cuda.synchronize()
    dist.barrier()
    ct = torch.tensor([c0.elapsed_time(c1)], dtype=torch.float64, device=dev)
    dist.allreduce_max_(ct)
    ceiling = h2d / (float(ct.item()) / reps * 1e-3) / 1e9
    achieved = h2d / (ms * 1e-3) / 1e9
    return {
        "value": float(n) ** 3 / (ms * 1e-3) / 1e9,
        "unit": UNIT,
        "h2d_bytes_per_step": int(h2d) * world,
        "d2h_bytes_per_step": int(d2h["bytes"]),
        "ms_per_step": ms,
        "steps": steps,
        "h2d_gbs_per_gpu": achieved,
        "h2d_ceiling_gbs": ceiling,
        "frac_of_h2d_ceiling": achieved / ceiling,
        "note": "pinned host fp64 fields -> chunked cudaMemcpyAsync on a side stream, each chunk's moment passes and x/y "
                "transforms run as it lands (stats.host_step) -> z transforms, binning -> results on the host; PCIe-bound. "
                "h2d bytes are the whole-job total over all ranks; h2d_ceiling_gbs = the same pinned buffers copied in the "
                "same chunks with nothing else running, all ranks at once, slowest rank (per GPU)",
    }


def load_profile_traffic(stage: str, world: int):
    """dram bytes per launch of a kernel from the committed ncu capture of the N=1 bench (profiles/traffic.json), scaled
    by 1/N for a slab (ncu never wraps a multi-rank command)."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            d = json.loads(p.read_text())
            if d.get(stage) is not None:
                return float(d[stage]) / world, ("ncu --set full capture at N=1 (%s)%s" % (
                    d.get("source", "profiles/"), "" if world == 1 else f", divided by {world} ranks"))
        except Exception:
            pass
    return None, "no capture for this kernel"


# ------------------------------------------------------------------------------------------------
# secondary measurements (BASELINE configs the default workload does not run), short, after everything else
# ------------------------------------------------------------------------------------------------
def _timeit(fn, reps=5, warm=2):
    import torch

    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def run_extras(rank, world, dev, peak) -> dict:
    import torch

    from fava_b200 import device, dist, series, synth

    out = {}
    g = torch.Generator(device=dev)
    g.manual_seed(17 + rank)
    if world == 1:
        # ---- C3: 512^3 fp64 profiles on one GPU (inputs 4 GiB >> L2) ---------------------------------------------
        n = 512
        f = [torch.rand((n, n, n), generator=g, device=dev, dtype=torch.float64) + 0.5 for _ in range(4)]
        cells = float(n) ** 3
        t_xz = _timeit(lambda: device.plane_moments_xz(*f))
        t_y = _timeit(lambda: device.plane_moments(*f, 1))
        t_xyz = _timeit(lambda: device.plane_moments_xyz(*f))
        c3 = {"workload": "BASELINE configs[2]: 512^3 fp64 Reynolds + Favre profiles x/y/z, 1 GPU",
              "plane_moments_xyz_ms": t_xyz, "plane_moments_xyz_gcells_per_s": cells / (t_xyz * 1e-3) / 1e9,
              "plane_moments_xyz_frac_of_hbm_peak": B_PROFILE * cells / (t_xyz * 1e-3) / 1e9 / peak,
              "plane_moments_xz_ms": t_xz, "plane_moments_axis1_ms": t_y,
              "plane_moments_xz_frac_of_hbm_peak": B_PROFILE * cells / (t_xz * 1e-3) / 1e9 / peak,
              "plane_moments_axis1_frac_of_hbm_peak": B_PROFILE * cells / (t_y * 1e-3) / 1e9 / peak,
              "profiles_xyz_ms": t_xz + t_y, "gcells_per_s": cells / ((t_xz + t_y) * 1e-3) / 1e9,
              "frac_of_hbm_peak_two_reads_64B": 2 * B_PROFILE * cells / ((t_xz + t_y) * 1e-3) / 1e9 / peak,
              "frac_of_hbm_peak_one_read_32B": B_PROFILE * cells / ((t_xz + t_y) * 1e-3) / 1e9 / peak}
        out["c3_profiles512"] = c3
        del f
        # ---- C2: from_amr gather, 8^3 blocks over 4 levels -> 256^3 -----------------------------------------------
        mesh = synth.octree_mesh((4, 4, 4), (8, 8, 8), 4, seed=11, p_refine=0.5)
        leaves = np.flatnonzero(mesh.node_type == 1)
        blk = torch.rand((mesh.nblocks, 8, 8, 8), generator=g, device=dev, dtype=torch.float32)
        scale = (2 ** (mesh.lmax - mesh.level[leaves])).astype(np.int64)
        table = device.prolong_table(leaves, mesh.origin[leaves] * 8 * scale[:, None], scale)
        dst = torch.empty((256, 256, 256), dtype=torch.float64, device=dev)
        ms = _timeit(lambda: device.prolong(blk, table, (256, 256, 256), out=dst))
        nbytes = 4.0 * leaves.size * 512 + 8.0 * 256**3
        out["c2_from_amr_256"] = {"workload": f"BASELINE configs[1]: {leaves.size} leaves of 8^3 f32 over 4 levels -> 256^3 fp64 "
                                              "(fava_prolong, tables cached)", "ms": ms, "algorithmic_bytes": nbytes,
                                  "achieved_gbs": nbytes / (ms * 1e-3) / 1e9, "frac_of_hbm_peak": nbytes / (ms * 1e-3) / 1e9 / peak,
                                  "note": "134 MB written: the output fits the 126 MB L2 only partly; see prolong_512 for a size that does not"}
        del blk, dst
        # ---- block-list moments (AMR / multi-block files): 512^3 cells f32 in 8^3 and 16^3 blocks, 16 B/cell ------
        from fava_b200 import uniform_analysis as ua

        n = 512
        bm = {"workload": "reynolds_stress moments of a block file: 512^3 cells f32 as single-level 8^3 / 16^3 blocks, axis x "
                          "(fava_plane_moments_blocks_uid; tables cached), 16 B/cell"}
        for nb in (8, 16):
            nblk, per = (n // nb) ** 3, n // nb
            f = [torch.rand((nblk, nb, nb, nb), generator=g, device=dev, dtype=torch.float32) + (1.0 if i == 0 else -0.5)
                 for i in range(4)]
            b = np.arange(nblk)
            table = device.leaf_table(b, (b % per) * nb, np.ones(nblk, dtype=np.int64), np.full(nblk, 1.0 / n**3))
            ms = _timeit(lambda: device.plane_moments_blocks(*f, 0, table, n))
            bm[f"blocks{nb}_ms"] = ms
            bm[f"blocks{nb}_frac_of_hbm_peak"] = 16.0 * n**3 / (ms * 1e-3) / 1e9 / peak
            del f
        out["block_moments_512"] = bm
        # ---- box counting (fractal_dimension): 1024^3 f32, a wrinkled iso-surface, s B/cell -------------------------
        n = 1024
        ar = torch.arange(n, device=dev, dtype=torch.float64)
        sheet = (ar[None, None, :] - 0.5 * n - 0.25 - 20.0 * torch.sin(2 * np.pi * ar / n)[None, :, None]
                 * torch.cos(4 * np.pi * ar / n)[:, None, None]).to(torch.float32)
        counts = torch.zeros(32, dtype=torch.int64, device=dev)
        coarse = torch.zeros([n // 32] * 3, dtype=torch.uint8, device=dev)

        def count_boxes():
            counts.zero_()
            device.fractal_tiles(sheet, 0.0, n, 0, 0, n, counts, coarse)
            device.fractal_coarse(coarse, (n, n, n), ua.box_levels((n, n, n)), counts)

        ms = _timeit(count_boxes)
        out["box_counting_1024"] = {"workload": "fractal_dimension box counts of a wrinkled sheet, 1024^3 f32 (fava_fractal_tiles + "
                                                "fava_fractal_coarse), 4 B/cell", "ms": ms,
                                    "frac_of_hbm_peak": 4.0 * n**3 / (ms * 1e-3) / 1e9 / peak,
                                    "boxes_per_level": [int(v) for v in counts[: ua.box_levels((n, n, n))].tolist()]}
        del sheet, counts, coarse
    # ---- prolongation sharded over ranks: 16^3 blocks, 4 levels -> 512^3, every rank fills its z-slab --------------------
    mesh = synth.octree_mesh((4, 4, 4), (16, 16, 16), 4, seed=11, p_refine=0.5)
    leaves = np.flatnonzero(mesh.node_type == 1)
    scale = (2 ** (mesh.lmax - mesh.level[leaves])).astype(np.int64)
    off = mesh.origin[leaves] * 16 * scale[:, None]
    za, zb = dist.parallel_range(512)
    keep = (off[:, 2] < zb) & (off[:, 2] + 16 * scale > za)
    blk = torch.rand((mesh.nblocks, 16, 16, 16), generator=g, device=dev, dtype=torch.float32)
    mine = off[keep] - np.array([0, 0, za])[None, :]
    table = device.prolong_table(leaves[keep], mine, scale[keep])
    dst = torch.empty((zb - za, 512, 512), dtype=torch.float64, device=dev)
    dist.barrier()
    ms = _timeit(lambda: device.prolong(blk, table, (zb - za, 512, 512), out=dst))
    tms = torch.tensor([ms], dtype=torch.float64, device=dev)
    dist.allreduce_max_(tms)
    ms = float(tms.item())
    nbytes_rank = 4.0 * int(keep.sum()) * 4096 + 8.0 * (zb - za) * 512 * 512
    out["prolong_512"] = {"workload": f"{leaves.size} leaves of 16^3 f32 over 4 levels -> 512^3 fp64, output z-slabs over {world} "
                                      "rank(s), no collective (SURVEY 8e3)", "ms_slowest_rank": ms,
                          "gcells_per_s": 512.0**3 / (ms * 1e-3) / 1e9, "achieved_gbs_per_gpu": nbytes_rank / (ms * 1e-3) / 1e9,
                          "frac_of_hbm_peak": nbytes_rank / (ms * 1e-3) / 1e9 / peak}
    del blk, dst
    torch.cuda.empty_cache()
    # ---- C5: Reynolds-stress time series streamed from pinned host memory -----------------------------------------------
    out["c5_series"] = series.streamed_series_benchmark(rank, world, dev)
    return out


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm
# ------------------------------------------------------------------------------------------------
def _sample_fields(n_sample: int, seed: int = 1234) -> dict:
    rng = np.random.default_rng(seed)
    shape = (n_sample,) * 3
    return {
        "dens": 1.0 + 0.5 * rng.random(shape),
        "velx": 0.25 * rng.standard_normal(shape),
        "vely": 0.25 * rng.standard_normal(shape),
        "velz": 0.25 * rng.standard_normal(shape),
    }


def _cpu_sample_port(n_sample: int, spectrum: bool) -> float:
    """Time the oracle port of the reference algorithm (bit-identical to the reference on the golden vectors) on an
    n_sample^3 fp64 single-block snapshot: reynolds_stress along x/y/z (+ kinetic_energy_spectra)."""
    from oracle import fava_oracle as orc

    file_fields = _sample_fields(n_sample)
    shape = (n_sample,) * 3
    data3 = {k: orc.load_like_reference(v) for k, v in file_fields.items()}  # loader not timed (fields preloaded)
    del file_fields
    geom = orc.uniform_geom(shape, bbox_dtype=np.float64)
    data4 = {k: v[None, ...] for k, v in data3.items()}
    t0 = time.perf_counter()
    for ax in AXES:
        orc.reynolds_stress(geom, data4, axis=ax)
    if spectrum:
        orc.kinetic_energy_spectra(data3, shape)
    return time.perf_counter() - t0


def _cpu_sample_reference(n_sample: int, spectrum: bool) -> float:
    """Time the UNMODIFIED reference (oracle/ref_harness.py: /root/reference under import shims) on the same sample:
    FLASH.reynolds_stress three times (raxis = 0, 1, 2: the same arithmetic each time, _flash.py:1506-1611) and
    FlashUniform.kinetic_energy_spectra (FlashUniform.py:229-304); files written first, fields preloaded."""
    import tempfile

    from fava_b200 import synth
    from oracle import ref_harness

    RefAMR, RefUniform, _ = ref_harness.ref_modules()
    fields = _sample_fields(n_sample)
    shape = (n_sample,) * 3
    names = ["dens", "velx", "vely", "velz"]
    with tempfile.TemporaryDirectory(prefix="fava_refarm_") as tmp:
        mesh = synth.single_block_mesh(shape)
        blk = Path(tmp) / "sample_hdf5_chk_0000"
        synth.write_flash_file(blk, mesh, {k: v[None, ...] for k, v in fields.items()}, checkpoint=True)
        uni = Path(tmp) / "sample_hdf5_chk_uniform_0000"
        synth.write_flash_file(uni, mesh, fields, checkpoint=True, uniform3d=True)
        del fields
        m = RefAMR(str(blk))
        m.load()
        m.load_data(names)
        u = None
        if spectrum:
            u = RefUniform(str(uni))
            u.load()
            u.load_data(names)
        t0 = time.perf_counter()
        for ax in AXES:
            m.reynolds_stress(raxis=ax)
        if spectrum:
            u.kinetic_energy_spectra()
        return time.perf_counter() - t0


def _cpu_sample(n_sample: int, spectrum: bool) -> tuple[float, str]:
    from oracle import ref_harness

    if ref_harness.reference_available():
        try:
            import contextlib

            with contextlib.redirect_stdout(sys.stderr):  # the reference prints its own timing lines
                return _cpu_sample_reference(n_sample, spectrum), "reference"
        except Exception as exc:  # never lose the baseline to a harness problem: fall back to the port and say so
            print(f"reference harness failed ({exc!r}); timing the port", file=sys.stderr)
    return _cpu_sample_port(n_sample, spectrum), "port"


def _cpu_sample_text(n_s: int, wl, sec: float, kind: str) -> str:
    what = ("the unmodified reference (/root/reference under the import shims of oracle/ref_harness.py)" if kind == "reference"
            else "NumPy port of the reference algorithm (oracle/fava_oracle.py; /root/reference is absent on this box; BASELINE.md "
                 "holds both timed side by side)")
    return (f"{n_s}^3 fp64 single-block sample of the workload (reynolds_stress x/y/z"
            f"{' + kinetic_energy_spectra' if wl['spectrum'] else ''}), {what}, fields preloaded, {sec:.2f} s per "
            f"step; 1 process = 1 core: the reference parallelises only over MPI ranks (none here) and needs "
            f"~230 B/cell for the spectrum, so 1024^3 cannot run on a host; host has {os.cpu_count()} cpus")


def cpu_baseline(wl) -> dict:
    n_s = wl["cpu_n"]
    sec, kind = _cpu_sample(n_s, wl["spectrum"])
    return {"value": float(n_s) ** 3 / sec / 1e9, "unit": UNIT, "cores": 1, "kind": kind,
            "sample": _cpu_sample_text(n_s, wl, sec, kind)}


def run_reference(args) -> dict:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return {}
    wl = WORKLOADS[args.workload]
    n_s = wl["cpu_n"]
    for _ in range(min(args.warmup, 1)):
        _cpu_sample(64, wl["spectrum"])
    steps = max(1, min(args.steps, 3))
    runs = [_cpu_sample(n_s, wl["spectrum"]) for _ in range(steps)]
    sec = float(np.mean([r[0] for r in runs]))
    kind = runs[0][1]
    value = float(n_s) ** 3 / sec / 1e9
    sample = _cpu_sample_text(n_s, wl, sec, kind)
    return {
        "impl": "reference",
        "metric": METRIC,
        "value": value,
        "unit": UNIT,
        "n_gpus": int(os.environ.get("WORLD_SIZE", "1")),
        "steps": steps,
        "warmup": min(args.warmup, 1),
        "ms_per_step": sec * 1e3,
        "higher_is_better": True,
        "scaling": "strong",
        "vs_baseline": None,
        "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": wl["desc"], "grid": [wl["n"]] * 3, "sample": sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": kind, "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD, choices=sorted(WORKLOADS))
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary measurements (C2, C3, C5, sharded prolongation)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    # stdout carries exactly ONE line, the JSON: everything imported code prints there (NCCL's version banner, the
    # reference's timing lines and its FAVA_MPI.__del__ message at exit) is sent to stderr at the descriptor level
    sys.stdout.flush()
    json_fd = os.dup(1)
    os.dup2(2, 1)
    out = run_reference(args) if args.impl == "reference" else run_ours(args)
    failed = bool(out.pop("_failed", False)) or "error" in out
    if out:
        os.write(json_fd, (json.dumps(out) + "\n").encode())
    os.close(json_fd)
    try:
        import torch.distributed as td

        if td.is_available() and td.is_initialized():
            td.destroy_process_group()
    except Exception:
        pass
    if failed:
        sys.exit(3)


if __name__ == "__main__":
    main()
