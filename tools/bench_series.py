#!/usr/bin/env python
"""BASELINE configs[4] in miniature: Reynolds-stress time series over synthetic multi-block plt files, sharded
over the ranks, staged file -> pinned ring -> HBM.  Prints one JSON line (not the driver's bench contract).

    python tools/bench_series.py [--grid 512] [--block 16] [--files 4]        (torchrun for N > 1)
"""
import argparse
import json
import sys
import tempfile
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--grid", dest="n", type=int, default=512)
    ap.add_argument("--block", type=int, default=16)
    ap.add_argument("--files", type=int, default=4)
    ap.add_argument("--dir", default=None)
    args = ap.parse_args()
    import torch

    import fava_b200
    from fava_b200 import dist, series, synth

    rank, world, local = dist.init_from_env()
    torch.cuda.set_device(local)
    n, nb = args.n, args.block
    root = Path(args.dir or tempfile.gettempdir()) / f"fava_series_{n}_{nb}"
    if rank == 0:
        root.mkdir(parents=True, exist_ok=True)
        mesh = synth.multiblock_mesh((n // nb,) * 3, (nb,) * 3)
        rng = np.random.default_rng(7)
        for i in range(args.files):
            path = root / f"series_hdf5_plt_cnt_{i:04d}"
            if path.exists():
                continue
            fields = {}
            for k in ("dens", "velx", "vely", "velz"):
                a = rng.random((mesh.nblocks, nb, nb, nb), dtype=np.float32)
                fields[k] = a + 1.0 if k == "dens" else a - 0.5
            synth.write_flash_file(path, mesh, fields, time=0.1 * i)
    dist.barrier()
    model = fava_b200.flash(root)
    series.reynolds_series(model, axis=0, indices=[0])  # warm-up: library load, pinned ring, page cache
    dist.barrier()
    t0 = time.perf_counter()
    res, timing = series.reynolds_series(model, axis=0)
    torch.cuda.synchronize()
    dist.barrier()
    dt = time.perf_counter() - t0
    if rank == 0:
        cells = float(n) ** 3 * args.files
        print(json.dumps({"workload": f"{args.files} plt files, {n}^3 cells in {nb}^3-cell blocks, f32, reynolds_stress(axis=0)",
                          "n_gpus": world, "seconds": dt, "gcells_per_s": cells / dt / 1e9,
                          "file_bytes_per_s_gb": 16.0 * cells / dt / 1e9, "timing_rank0": timing,
                          "profile_head": [float(v) for v in res[-1][2]["Rxx"][:3]]}))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
