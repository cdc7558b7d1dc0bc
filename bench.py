#!/usr/bin/env python
"""bench.py — FAVA grid-statistics hot path on B200 (contract: task statement; numbers explained in DESIGN.md §5).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one pass of the hot path over one synthetic snapshot resident in HBM: Reynolds + Favre plane
profiles along x, y and z, plus the kinetic-energy spectrum.  Under torchrun (N > 1) the SAME global grid is
split into z-slabs, one per rank ("strong" scaling).  Rank 0 prints ONE JSON line.
"""

from __future__ import annotations

import argparse
import json
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

# algorithmic bytes per cell (SURVEY §8d / DESIGN.md §5)
B_PROFILE = 32.0  # rho, ux, uy, uz read once, fp64
B_SPECTRUM = 200.0  # 3 separable line passes in Hermitian storage, weighting fused, binning incl. transposed operand
B_WEIGHT = 32.0 + 24.0  # read 4 fields, write 3 real fp64 arrays
B_BIN = 48.0  # 3 components x 16 B x 1/2 (r2c) x 2 (point + transposed operand)


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


def run_ours(args) -> dict:
    import torch

    from fava_b200 import device, dist, spectrum, stats
    from fava_b200.build import build_library

    build_library()  # no-op when the in-tree .so is current
    rank, world, local = dist.init_from_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    wl = WORKLOADS[args.workload]
    n = wl["n"]
    if n % (2 * world):
        raise SystemExit(f"grid {n} not divisible by 2 x {world} ranks")
    nz = n // world
    z0 = rank * nz
    fields = synth_slab_device(n, z0, nz, dev)
    cell_volume = 1.0 / float(n) ** 3
    layer_volume = 1.0 / float(n)
    ncells = float(n) ** 3
    peak, peak_src = measured_peak_gbs()

    def step(f, timer: StageTimer | None = None):
        """The public per-step path: slab profiles (x,y,z) + slab spectrum.  Returns host-readable results."""
        return stats.slab_step(*f, n, cell_volume, layer_volume, axes=AXES, spectrum=wl["spectrum"], favre=True)

    def step_instrumented(timer: StageTimer):
        """Same work, stage by stage, with event brackets (used once after the timed region)."""
        rho, ux, uy, uz = fields
        with timer.bracket("plane_moments_xz"):  # x-bins and z-bins from one pass (fava_plane_moments_xz)
            (mx, px), (mz, pz) = device.plane_moments_xz(rho, ux, uy, uz)
        with timer.bracket("plane_moments_axis1"):
            my, py = device.plane_moments(rho, ux, uy, uz, 1)
        for ax, mom, piv in ((0, mx, px), (1, my, py), (2, mz, pz)):
            stats.slab_profiles_finish(mom, piv, ax, cell_volume, layer_volume, gather=False)
        if not wl["spectrum"]:
            return
        nxh = n // 2 + 1
        if world == 1:
            w = [device.workspace(3 + 1 + c, 16 * n * n * nxh, dev) for c in range(3)]  # WS_FFT1..3 of fava_ke_spectrum
            sums = torch.zeros((3, n // 2 - 1), dtype=torch.float64, device=dev)
            with timer.bracket("ke_weight3"):
                device.ke_weight3(rho, ux, uy, uz, *w)
            for c in range(3):
                with timer.bracket("cufft_xy"):
                    device.fft_xy(w[c], n, n, n, dev)
                with timer.bracket("cufft_z"):
                    device.fft_z(w[c], n, n * nxh, dev)
            with timer.bracket("spectrum_bin"):
                device.spectrum_bin(w[0], w[1], w[2], n, n, None, None, sums)
        else:
            p = spectrum._plan(n, rank, world, dev)
            with timer.bracket("ke_weight3"):
                device.ke_weight3(rho, ux, uy, uz, *p.send)
            for c in range(3):
                with timer.bracket("cufft_xy"):
                    device.fft_xy(p.send[c], p.nzl, n, n, dev)
                with timer.bracket("a2a_pack"):
                    spectrum.exchange(p, c)
            dist.allreduce_sum_(p.tokens[0])
            for c in range(3):
                with timer.bracket("cufft_z"):
                    device.fft_z(p.recv[c], n, p.nyl * p.nxh, dev)
            with timer.bracket("spectrum_bin"):
                device.spectrum_bin(p.recv[0], p.recv[1], p.recv[2], n, p.nyl, p.ky_of_local, p.local_of_ky, p.sums)
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
    stages = {}
    for name, ms in stage_ms.items():
        if name.startswith("plane_moments"):  # the fused x+z pass also reads each field once: 32 B/cell
            algo = B_PROFILE * local_cells
        elif name == "ke_weight3":
            algo = B_WEIGHT * local_cells
        elif name == "spectrum_bin":
            algo = B_BIN * local_cells
        elif name in ("cufft_xy", "cufft_z"):
            algo = None  # library call (cuFFT), not our kernel
        elif name == "a2a_pack":
            algo = 8.0 * local_cells  # one complex r2c row set read once; (world-1)/world of it crosses NVLink
        else:
            algo = None
        stages[name] = {"ms": ms, "launches_per_step": len(timer.pairs[name]) // 2}
        if algo is not None:
            stages[name]["algorithmic_bytes"] = algo
            stages[name]["achieved_gbs"] = algo / (ms * 1e-3) / 1e9
            stages[name]["frac_of_hbm_peak"] = stages[name]["achieved_gbs"] / peak
    if "a2a_pack" in stages:
        stages["a2a_pack"]["nvlink_gbs_per_gpu"] = 8.0 * local_cells * (world - 1) / world / (stages["a2a_pack"]["ms"] * 1e-3) / 1e9
        stages["a2a_pack"]["frac_of_nvlink_770"] = stages["a2a_pack"]["nvlink_gbs_per_gpu"] / 770.0

    # dominant kernel of OUR code (time per step = ms x launches per step)
    own = {k: v for k, v in stages.items() if "algorithmic_bytes" in v and k != "a2a_pack"}
    dom = max(own, key=lambda k: own[k]["ms"] * own[k]["launches_per_step"])
    kernel_names = {"spectrum_bin": "k_spectrum_bin (fava_spectrum_bin)", "ke_weight3": "k_ke_weight3 (fava_ke_weight3)",
                    "plane_moments_xz": "k_moments_cols + k_partials_to_planes (fava_plane_moments_xz: x AND z profiles "
                                        "from one 32 B/cell read)",
                    "plane_moments_axis1": "k_moments_rows (fava_plane_moments, axis y)"}
    traffic = load_profile_traffic(dom)
    roofline = {
        "kernel": kernel_names.get(dom, dom),
        "bound": "hbm",
        "achieved": own[dom]["achieved_gbs"],
        "peak": peak,
        "unit": "GB/s",
        "frac": own[dom]["achieved_gbs"] / peak,
        "traffic": traffic,
        "peak_source": peak_src,
        "algorithmic_bytes_per_launch": own[dom]["algorithmic_bytes"],
        "kernel_ms": own[dom]["ms"],
        "note": "dominant hand-written kernel by time per step; every stage is listed under roofline_stages "
                "(cuFFT passes are library calls and carry no roofline claim)",
    }
    prof_ms = sum(v["ms"] for k, v in stages.items() if k.startswith("plane_moments"))
    # x, y and z profiles now take two passes over the fields (x+z fused, y): 64 B/cell of algorithmic traffic
    summary = {"profiles_xyz": {"ms": prof_ms, "model_bytes_per_cell": 2 * B_PROFILE,
                                "achieved_gbs": 2 * B_PROFILE * local_cells / (prof_ms * 1e-3) / 1e9}}
    summary["profiles_xyz"]["frac_of_hbm_peak"] = summary["profiles_xyz"]["achieved_gbs"] / peak
    if wl["spectrum"]:
        spec_ms = sum(v["ms"] * v["launches_per_step"] for k, v in stages.items() if not k.startswith("plane_moments"))
        summary["ke_spectrum"] = {"ms": spec_ms, "model_bytes_per_cell": B_SPECTRUM,
                                  "achieved_gbs": B_SPECTRUM * local_cells / (spec_ms * 1e-3) / 1e9}
        summary["ke_spectrum"]["frac_of_hbm_peak"] = summary["ke_spectrum"]["achieved_gbs"] / peak

    timeline = None
    if world > 1 and wl["spectrum"]:  # where one public step spends its time (events on both streams, rank 0)
        p = spectrum._plan(n, rank, world, dev)
        s0 = torch.cuda.Event(enable_timing=True)
        s1 = torch.cuda.Event(enable_timing=True)
        dist.barrier()
        torch.cuda.synchronize()
        s0.record()
        step(fields)
        s1.record()
        torch.cuda.synchronize()
        timeline = {"xy_done_ms": [s0.elapsed_time(e) for e in p.ev_xy], "exchange_done_ms": [s0.elapsed_time(e) for e in p.ev_done],
                    "moments_done_ms": s0.elapsed_time(p.ev_mark["overlap"]), "fft_z_done_ms": s0.elapsed_time(p.ev_mark["fft_z"]),
                    "bin_done_ms": s0.elapsed_time(p.ev_mark["bin"]), "step_done_ms": s0.elapsed_time(s1)}

    def host_step(host, stage):
        return stats.host_step(host, n, cell_volume, layer_volume, axes=AXES, spectrum=wl["spectrum"], favre=True, stage=stage)

    e2e = run_e2e(args, wl, dev, rank, world, fields, step, host_step)

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
        "clocks": clocks,
    }
    if rank == 0 and world == 1:
        out["cpu_baseline"] = cpu_baseline(wl)
    return out if rank == 0 else {}


def run_e2e(args, wl, dev, rank, world, dev_fields, step, host_step) -> dict:
    """Same step through HOST buffers: pinned host fields -> H2D (in the timed region) -> public step ->
    profiles and spectrum read back to the host."""
    import torch

    from fava_b200 import dist

    host = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in dev_fields]
    for h, d in zip(host, dev_fields):
        h.copy_(d)
    stage = [torch.empty_like(t) for t in dev_fields]
    torch.cuda.synchronize()
    h2d = sum(h.numel() * h.element_size() for h in host)
    d2h = {"bytes": 0}

    def e2e_step():
        # the user-facing call for host-resident snapshots: chunked H2D on a side stream, each chunk consumed as it
        # lands (stats.host_step); FAVA_E2E=serial copies the whole slab first and then runs the resident step
        if os.environ.get("FAVA_E2E") == "serial":
            for h, s in zip(host, stage):
                s.copy_(h, non_blocking=True)
            res = step(stage)
        else:
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
    n = wl["n"]
    return {
        "value": float(n) ** 3 / (ms * 1e-3) / 1e9,
        "unit": UNIT,
        "h2d_bytes_per_step": int(h2d) * world,
        "d2h_bytes_per_step": int(d2h["bytes"]),
        "ms_per_step": ms,
        "steps": steps,
        "note": "pinned host fp64 fields -> chunked cudaMemcpyAsync on a side stream, each chunk's moment passes and "
                "weighting + 2-D transforms run as it lands (stats.host_step) -> z transforms, binning -> results on the "
                "host; PCIe-bound (h2d bytes are the whole-job total over all ranks)",
    }


def load_profile_traffic(stage: str):
    """dram bytes per launch of a kernel from the committed ncu capture (profiles/traffic.json), if any."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get(stage)
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port of the reference's NumPy algorithm
# ------------------------------------------------------------------------------------------------
def _cpu_sample(n_sample: int, spectrum: bool, seed: int = 1234) -> float:
    """Time the reference algorithm (oracle port, bit-identical to the reference on the golden vectors) on an
    n_sample^3 fp64 single-block snapshot: reynolds_stress along x/y/z (+ kinetic_energy_spectra)."""
    from oracle import fava_oracle as orc

    rng = np.random.default_rng(seed)
    shape = (n_sample,) * 3
    file_fields = {
        "dens": 1.0 + 0.5 * rng.random(shape),
        "velx": 0.25 * rng.standard_normal(shape),
        "vely": 0.25 * rng.standard_normal(shape),
        "velz": 0.25 * rng.standard_normal(shape),
    }
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


def _cpu_sample_text(n_s: int, wl, sec: float) -> str:
    return (f"{n_s}^3 fp64 single-block sample of the workload (reynolds_stress x/y/z"
            f"{' + kinetic_energy_spectra' if wl['spectrum'] else ''}), NumPy port of the reference algorithm "
            f"(oracle/fava_oracle.py; the reference itself is pure NumPy/SciPy), fields preloaded, {sec:.2f} s per "
            f"step; 1 process = 1 core: the reference parallelises only over MPI ranks (none here) and needs "
            f"~230 B/cell for the spectrum, so 1024^3 cannot run on a host; host has {os.cpu_count()} cpus")


def cpu_baseline(wl) -> dict:
    n_s = wl["cpu_n"]
    sec = _cpu_sample(n_s, wl["spectrum"])
    return {"value": float(n_s) ** 3 / sec / 1e9, "unit": UNIT, "cores": 1, "kind": "port",
            "sample": _cpu_sample_text(n_s, wl, sec)}


def run_reference(args) -> dict:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return {}
    wl = WORKLOADS[args.workload]
    n_s = wl["cpu_n"]
    for _ in range(min(args.warmup, 1)):
        _cpu_sample(64, wl["spectrum"])
    steps = max(1, min(args.steps, 3))
    secs = [_cpu_sample(n_s, wl["spectrum"]) for _ in range(steps)]
    sec = float(np.mean(secs))
    value = float(n_s) ** 3 / sec / 1e9
    sample = _cpu_sample_text(n_s, wl, sec)
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
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": 1, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default=os.environ.get("FAVA_BENCH_WORKLOAD", DEFAULT_WORKLOAD), choices=sorted(WORKLOADS))
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    out = run_reference(args) if args.impl == "reference" else run_ours(args)
    if out:
        print(json.dumps(out), flush=True)
    try:
        import torch.distributed as td

        if td.is_available() and td.is_initialized():
            td.destroy_process_group()
    except Exception:
        pass


if __name__ == "__main__":
    main()
