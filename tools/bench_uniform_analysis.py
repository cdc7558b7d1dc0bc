#!/usr/bin/env python
"""Kernel timings of the uniform-grid analyses (1 x B200): box counting (csrc/fractal.cu) against the HBM roofline
(algorithmic bytes = s per cell, the field read once) and the structure-function gather / moment kernels
(csrc/structure.cu; random 4-byte gathers, reported as pairs/s).  CUDA-event times; prints one JSON line."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

from tools.bench_kernels import timeit  # noqa: E402


def main():
    import numpy as np
    import torch

    from fava_b200 import device
    from fava_b200 import uniform_analysis as ua

    dev = torch.device("cuda", 0)
    peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]) if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
    out = {"peak_gbs": peak, "kernels": {}}
    g = torch.Generator(device=dev)
    g.manual_seed(5)
    quick = "quick" in sys.argv[1:]  # one size only (used under ncu)
    for n, dt in ((1024, torch.float32),) if quick else ((1024, torch.float32), (512, torch.float64), (768, torch.float64)):
        if n % 2 ** (ua.box_levels((n, n, n)) - 1):
            tiles_only = True  # 768 is not a multiple of 512: time the tile kernel alone
        else:
            tiles_only = False
        f = torch.rand((n, n, n), generator=g, device=dev, dtype=dt)
        counts = torch.zeros(32, dtype=torch.int64, device=dev)
        coarse = torch.zeros([(n + 31) // 32] * 3, dtype=torch.uint8, device=dev)
        ar = torch.arange(n, device=dev, dtype=torch.float64)
        sheet = (ar[None, None, :] - 0.5 * n - 0.25 - 20.0 * torch.sin(2 * np.pi * ar / n)[None, :, None]
                 * torch.cos(4 * np.pi * ar / n)[:, None, None]).to(dt)  # a wrinkled surface: the realistic case
        for shape in ("default",):
            for contour, tag in ((0.5, "noise"), (2.0, "empty"), (0.0, "sheet")):
                src = sheet if tag == "sheet" else f

                def run():
                    counts.zero_()
                    device.fractal_tiles(src, contour, n, 0, 0, n, counts, coarse)
                    if not tiles_only:
                        device.fractal_coarse(coarse, (n, n, n), ua.box_levels((n, n, n)), counts)
                ms = timeit(run)
                nbytes = float(f.element_size()) * n**3
                key = f"fractal_{n}cubed_{str(dt).split('.')[-1]}_{tag}" + ("" if shape == "default" else f"_cta{shape}")
                out["kernels"][key] = {
                    "ms": ms, "algorithmic_bytes": nbytes, "achieved_gbs": nbytes / ms / 1e6,
                    "frac_of_hbm_peak": nbytes / ms / 1e6 / peak, "gcells_per_s": n**3 / ms / 1e6,
                    "note": "counts.zero_ + fava_fractal_tiles" + ("" if tiles_only else " + fava_fractal_coarse")}
        del f, sheet
    if quick:
        print(json.dumps(out))
        return
    n = 512
    vel = [torch.rand((n, n, n), generator=g, device=dev, dtype=torch.float32) for _ in range(3)]
    nsep, npts = 100, 10000
    p1 = torch.rand((nsep, npts, 3), generator=g, device=dev, dtype=torch.float64)
    p2 = torch.rand((nsep, npts, 3), generator=g, device=dev, dtype=torch.float64)
    err = torch.zeros(1, dtype=torch.int32, device=dev)
    lo, cell = np.zeros(3), np.full(3, 1.0 / n)
    ms_g = timeit(lambda: device.sf_gather(p1.view(-1, 3), *vel, n, 0, lo, cell, err))
    v1 = device.sf_gather(p1.view(-1, 3), *vel, n, 0, lo, cell, err)
    v2 = device.sf_gather(p2.view(-1, 3), *vel, n, 0, lo, cell, err)
    ms_m = timeit(lambda: device.sf_moments(p1, p2, v1, v2, nsep, npts, 7, False))
    out["kernels"]["sf_gather_1e6_points_512cubed_f32"] = {"ms": ms_g, "points_per_s": nsep * npts / ms_g * 1e3}
    out["kernels"]["sf_moments_100x10000_order7"] = {"ms": ms_m, "pairs_per_s": nsep * npts / ms_m * 1e3}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
