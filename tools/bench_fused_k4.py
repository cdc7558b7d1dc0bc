#!/usr/bin/env python
"""Fused weighting (FAVA_FUSE_K4=1: the x/z moment pass also writes sqrt(rho) u_n for the spectrum) against the
separate K4 pass on one GPU: bitwise parity of moments and weighted fields, then CUDA-event times of
moments_xz + K4 vs the fused pass, and of the whole resident step both ways.  Usage: bench_fused_k4.py [N=768]"""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fava_b200 import device, stats  # noqa: E402
from tools.fft_bench import timeit  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 768
    dev = torch.device("cuda", 0)
    nxh = n // 2 + 1
    g = torch.Generator(device=dev)
    g.manual_seed(2)
    f = [torch.rand((n, n, n), generator=g, device=dev, dtype=torch.float64) + 0.5 for _ in range(4)]
    pitch = 2 * nxh
    wt = [torch.zeros((n * n, pitch), dtype=torch.float64, device=dev) for _ in range(3)]
    w = [t.data_ptr() for t in wt]

    device.ke_weight3(*f, *w)
    ref_w = [t[:, :n].clone() for t in wt]
    (mx0, px0), (mz0, pz0) = device.plane_moments_xz(*f)
    for t in wt:
        t.zero_()
    (mx1, px1), (mz1, pz1) = device.plane_moments_xz(*f, weighted_out=w)
    assert torch.equal(mx0, mx1) and torch.equal(mz0, mz1) and torch.equal(px0, px1), "moments differ"
    for a, t in zip(ref_w, wt):
        assert torch.equal(a, t[:, :n]) and bool((t[:, n:] == 0).all()), "weighted fields differ"
    del ref_w
    print("parity ok (moments and weighted fields bitwise equal)")

    t_sep = timeit(lambda: (device.plane_moments_xz(*f), device.ke_weight3(*f, *w)), reps=5)
    t_fused = timeit(lambda: device.plane_moments_xz(*f, weighted_out=w), reps=5)
    print(f"n={n}: moments_xz + K4 {t_sep:.3f} ms, fused {t_fused:.3f} ms")
    cv, lv = 1.0 / float(n) ** 3, 1.0 / float(n)
    os.environ.pop("FAVA_FUSE_K4", None)
    a = stats.slab_step(*f, n, cv, lv)
    t0 = timeit(lambda: stats.slab_step(*f, n, cv, lv), reps=3)
    os.environ["FAVA_FUSE_K4"] = "1"
    b = stats.slab_step(*f, n, cv, lv)
    t1 = timeit(lambda: stats.slab_step(*f, n, cv, lv), reps=3)
    for k in a["spectrum"]:
        assert (a["spectrum"][k] == b["spectrum"][k]).all(), k
    for ax in (0, 1, 2):
        for k in a[ax]:
            assert torch.equal(a[ax][k], b[ax][k]), (ax, k)
    print(f"n={n}: resident step {t0:.3f} ms, with FAVA_FUSE_K4=1 {t1:.3f} ms (results bitwise equal)")


if __name__ == "__main__":
    main()
