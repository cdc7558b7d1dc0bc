#!/usr/bin/env python
"""bench.py — FAVA grid-statistics hot path on B200 (contract: see the task statement / DESIGN.md §Measurement).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload NAME]

One "step" = one pass of the hot path over one synthetic snapshot resident in HBM.  Under torchrun
(N>1) the snapshot is split into z-slabs, one per rank ("strong" scaling: the global grid is fixed).
Rank 0 prints ONE JSON line.
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
    # name: (N, do_profiles, do_spectrum)
    "profiles512": dict(n=512, axes=(0, 1, 2), spectrum=False,
                        desc="512^3 uniform fp64 Reynolds + Favre stress profiles along x/y/z (BASELINE configs[2])"),
    "profiles1024": dict(n=1024, axes=(0, 1, 2), spectrum=False,
                         desc="1024^3 uniform fp64 Reynolds + Favre stress profiles along x/y/z"),
    "full1024": dict(n=1024, axes=(0, 1, 2), spectrum=True,
                     desc="1024^3 uniform fp64: Reynolds + Favre profiles along x/y/z + kinetic_energy_spectra "
                          "(BASELINE configs[3])"),
    "full512": dict(n=512, axes=(0, 1, 2), spectrum=True,
                    desc="512^3 uniform fp64: profiles x/y/z + kinetic_energy_spectra"),
    "full256": dict(n=256, axes=(0, 1, 2), spectrum=True, desc="256^3 uniform fp64: profiles + spectrum (debug)"),
}
DEFAULT_WORKLOAD = "profiles512"


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
        self.thread = None

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

        self.thread = threading.Thread(target=pump, daemon=True)
        self.thread.start()

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
        return {
            "sm_mhz": float(np.median(sm)) if sm else None,
            "sm_max_mhz": float(max(smax)) if smax else None,
            "power_w_max": float(max(power)) if power else None,
            "samples": len(sm),
            "reasons": sorted(reasons),
        }


# ------------------------------------------------------------------------------------------------
# our arm
# ------------------------------------------------------------------------------------------------
def synth_slab_device(n: int, z0: int, nz: int, dev, seed: int = 1234):
    """Synthetic snapshot slab generated directly in HBM (dens>0, sheared velocities + noise)."""
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


def run_ours(args) -> dict:
    import torch

    from fava_b200 import device, dist, stats
    from fava_b200.build import build_library

    build_library()  # no-op when the in-tree .so is current
    rank, world, local = dist.init_from_env()
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device; the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    wl = WORKLOADS[args.workload]
    n = wl["n"]
    if n % world:
        raise SystemExit(f"grid {n} not divisible by {world} ranks")
    nz = n // world
    z0 = rank * nz
    rho, ux, uy, uz = synth_slab_device(n, z0, nz, dev)
    cell_volume = 1.0 / float(n) ** 3
    layer_volume = 1.0 / float(n)
    ncells_global = float(n) ** 3
    peak, peak_src = measured_peak_gbs()

    spectrum_fn = None
    if wl["spectrum"]:
        from fava_b200 import spectrum as spec_mod

        spectrum_fn = lambda: spec_mod.slab_ke_spectrum(rho, ux, uy, uz, n)  # noqa: E731

    ev_pairs = []  # (start, stop) events bracketing the dominant kernel's launches

    def step(record: bool):
        res = {}
        for ax in wl["axes"]:
            if record:
                e0 = torch.cuda.Event(enable_timing=True)
                e1 = torch.cuda.Event(enable_timing=True)
                piv = device.plane_pivots(ux, uy, uz, ax)
                if ax in (0, 1):
                    dist.broadcast_(piv, 0)
                e0.record()
                mom, _ = device.plane_moments(rho, ux, uy, uz, ax, pivots=piv)
                e1.record()
                ev_pairs.append((e0, e1))
                if ax in (0, 1):
                    dist.allreduce_sum_(mom)
                res[ax] = device.moments_finalize(mom, piv, cell_volume, layer_volume)
            else:
                res[ax] = stats.slab_profiles(rho, ux, uy, uz, ax, cell_volume, layer_volume, gather=False)
        if spectrum_fn is not None:
            res["spectrum"] = spectrum_fn()
        return res

    for _ in range(args.warmup):
        step(False)
    torch.cuda.synchronize()

    sampler = ClockSampler(local) if rank == 0 else None
    launches0 = device.launch_count()
    dist.barrier()
    torch.cuda.synchronize()
    if sampler:
        sampler.start()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record()
    for _ in range(args.steps):
        step(True)
    t1.record()
    torch.cuda.synchronize()
    dist.barrier()
    clocks = sampler.stop() if sampler else None
    launches = device.launch_count() - launches0
    elapsed_ms = torch.tensor([t0.elapsed_time(t1)], dtype=torch.float64, device=dev)
    dist.allreduce_max_(elapsed_ms)
    ms_per_step = float(elapsed_ms.item()) / args.steps
    value = ncells_global / (ms_per_step * 1e-3) / 1e9

    kern_ms = float(np.mean([a.elapsed_time(b) for a, b in ev_pairs]))
    algo_bytes = 32.0 * float(nz) * n * n  # 4 fp64 fields read once (4*s bytes per cell per call)
    achieved = algo_bytes / (kern_ms * 1e-3) / 1e9

    # ---- e2e: host buffers, H2D inside the timed region, result read back ---------------------
    e2e = run_e2e(args, wl, dev, rank, world, (rho, ux, uy, uz), cell_volume, layer_volume, spectrum_fn is not None)

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
            "l2": "inputs (4 fields x %.1f GiB per GPU) exceed the 126 MB L2; no flush needed" % (8.0 * nz * n * n / 2**30),
        },
        "e2e": e2e,
        "gpu_launches": int(launches),
        "roofline": {
            "kernel": "k_moments_{cols,rows} (fava_plane_moments, one launch per axis; mean over x,y,z launches)",
            "bound": "hbm",
            "achieved": achieved,
            "peak": peak,
            "unit": "GB/s",
            "frac": achieved / peak,
            "traffic": load_profile_traffic(),
            "peak_source": peak_src,
            "algorithmic_bytes_per_launch": algo_bytes,
            "kernel_ms": kern_ms,
        },
        "clocks": clocks,
    }
    if rank == 0:
        out["cpu_baseline"] = cpu_baseline(args, wl)
    return out if rank == 0 else {}


def run_e2e(args, wl, dev, rank, world, dev_fields, cell_volume, layer_volume, with_spectrum) -> dict:
    """Same step through host buffers: pinned host -> H2D -> kernels -> D2H of the profiles."""
    import torch

    from fava_b200 import dist, stats

    n = wl["n"]
    nz = n // world
    host = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in dev_fields]
    for h, d in zip(host, dev_fields):
        h.copy_(d)
    stage = [torch.empty_like(t) for t in dev_fields]
    torch.cuda.synchronize()
    h2d = sum(h.numel() * h.element_size() for h in host)
    d2h_holder = {}

    def e2e_step():
        for h, s in zip(host, stage):
            s.copy_(h, non_blocking=True)
        res = []
        for ax in wl["axes"]:
            out = stats.slab_profiles(*stage, ax, cell_volume, layer_volume, gather=False)
            res.extend(v.cpu() for v in out.values())
        if with_spectrum:
            from fava_b200 import spectrum as spec_mod

            sp = spec_mod.slab_ke_spectrum(*stage, n)
            res.extend(np.asarray(v) for v in sp.values())
        d2h_holder["bytes"] = sum(int(getattr(r, "nbytes", 0)) if isinstance(r, np.ndarray)
                                  else r.numel() * r.element_size() for r in res)
        return res

    e2e_step()
    dist.barrier()
    torch.cuda.synchronize()
    steps = max(1, min(args.steps, 5))
    t0 = time.perf_counter()
    for _ in range(steps):
        e2e_step()
    torch.cuda.synchronize()
    dist.barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    dist.allreduce_max_(dt)
    sec = float(dt.item()) / steps
    return {
        "value": float(n) ** 3 / sec / 1e9,
        "unit": UNIT,
        "h2d_bytes_per_step": int(h2d),
        "d2h_bytes_per_step": int(d2h_holder.get("bytes", 0)),
        "ms_per_step": sec * 1e3,
        "steps": steps,
        "note": "pinned host fp64 fields -> cudaMemcpyAsync -> kernels -> profiles copied back; per-rank bytes",
    }


def load_profile_traffic():
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if any."""
    p = ROOT / "profiles" / "traffic.json"
    if p.exists():
        try:
            return json.loads(p.read_text()).get("plane_moments_bytes_per_launch")
        except Exception:
            return None
    return None


# ------------------------------------------------------------------------------------------------
# CPU baseline / reference arm: the oracle port of the reference's NumPy algorithm
# ------------------------------------------------------------------------------------------------
def _cpu_sample(n_sample: int, axes, spectrum: bool, seed: int = 1234):
    """Time the reference algorithm (oracle port) on an n_sample^3 fp64 single-block snapshot."""
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
    for ax in axes:
        orc.reynolds_stress(geom, data4, axis=ax)
    if spectrum:
        orc.kinetic_energy_spectra(data3, shape)
    return time.perf_counter() - t0


def cpu_baseline(args, wl) -> dict:
    n_s = 256 if not wl["spectrum"] else 192
    sec = _cpu_sample(n_s, wl["axes"], wl["spectrum"])
    return {
        "value": float(n_s) ** 3 / sec / 1e9,
        "unit": UNIT,
        "cores": 1,
        "kind": "port",
        "sample": f"{n_s}^3 fp64 single-block snapshot, same step (reynolds_stress x/y/z"
                  f"{' + kinetic_energy_spectra' if wl['spectrum'] else ''}), NumPy port of the reference "
                  f"algorithm (oracle/fava_oracle.py), fields preloaded, {sec:.2f} s; host has {os.cpu_count()} cpus, "
                  "the reference parallelises only over MPI ranks/blocks so one block = one core",
    }


def run_reference(args) -> dict:
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return {}
    wl = WORKLOADS[args.workload]
    n_s = 256 if not wl["spectrum"] else 192
    for _ in range(min(args.warmup, 1)):
        _cpu_sample(64, wl["axes"], wl["spectrum"])
    steps = max(1, min(args.steps, 3))
    secs = [_cpu_sample(n_s, wl["axes"], wl["spectrum"]) for _ in range(steps)]
    sec = float(np.mean(secs))
    value = float(n_s) ** 3 / sec / 1e9
    sample = (f"{n_s}^3 fp64 single-block sample of the workload per step (the reference needs ~230 B/cell and hours at "
              f"1024^3), NumPy port of the reference algorithm, 1 process")
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
        "config": {"workload": wl["desc"], "sample": sample},
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
