#!/usr/bin/env python
"""K4 (fava_ke_weight3) variants at N^3 fp64 on one GPU: the pair-indexed grid-stride kernel (default) against the
row-walking kernel (FAVA_K4=rows); CUDA-event times.  Parity of the variant: tests/test_api_gpu.py under FAVA_K4=rows."""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fava_b200 import device  # noqa: E402
from tools.fft_bench import timeit  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    dev = torch.device("cuda", 0)
    nxh = n // 2 + 1
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    f = [torch.rand((n, n, n), generator=g, device=dev, dtype=torch.float64) + 0.5 for _ in range(4)]
    w = [device.workspace(4 + c, 16 * n * n * nxh, dev) for c in range(3)]
    for mode in ("pairs", "rows"):
        if mode == "rows":
            os.environ["FAVA_K4"] = "rows"
        else:
            os.environ.pop("FAVA_K4", None)
        t = timeit(lambda: device.ke_weight3(*f, *w), reps=5)
        print(f"K4 [{mode}] n={n}: {t:.3f} ms  {56.0 * n**3 / t / 1e6:.0f} GB/s")


if __name__ == "__main__":
    main()
