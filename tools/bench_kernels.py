#!/usr/bin/env python
"""Rooflines of the kernels that bench.py's main workload does not exercise (1 x B200): the block-list moment
kernels (K1b), the prolongation gather (K3), the single-field plane sums.  Achieved GB/s = algorithmic bytes /
CUDA-event time; peak = MEASURED_PEAKS.json.  Prints one JSON line."""
import json
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def timeit(fn, reps=5):
    import torch

    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def main():
    import torch

    from fava_b200 import device, synth

    dev = torch.device("cuda", 0)
    peak = float(json.loads((ROOT / "MEASURED_PEAKS.json").read_text())["hbm_gbs"]) if (ROOT / "MEASURED_PEAKS.json").exists() else 6650.0
    out = {"peak_gbs": peak, "kernels": {}}

    def record(name, ms, nbytes, note):
        out["kernels"][name] = {"ms": ms, "algorithmic_bytes": nbytes, "achieved_gbs": nbytes / ms / 1e6,
                                "frac_of_hbm_peak": nbytes / ms / 1e6 / peak, "note": note}

    g = torch.Generator(device=dev)
    g.manual_seed(3)
    # ---- K1b on a single-level multi-block file layout: 512^3 cells in 16^3 blocks, f32 (config 5 per-GPU share)
    blocks_only = "blocks" in sys.argv[1:]  # the block-list kernels alone (used under ncu)
    for nb, n in ((16, 512), (8, 512)):
        nblk = (n // nb) ** 3
        f = [torch.rand((nblk, nb, nb, nb), generator=g, device=dev, dtype=torch.float32) + (1.0 if i == 0 else -0.5) for i in range(4)]
        per = n // nb
        b = np.arange(nblk)
        for axis in (0, 1, 2):
            idx = (b % per, (b // per) % per, b // (per * per))[axis]
            table = device.leaf_table(b, idx * nb, np.ones(nblk, dtype=np.int64), np.full(nblk, 1.0 / n**3))
            ms = timeit(lambda: device.plane_moments_blocks(*f, axis, table, n))
            record(f"block_moments_{nb}cubed_f32_axis{axis}", ms, 16.0 * n**3,
                   f"fava_plane_moments_blocks incl. host CSR build + table upload; {nblk} leaves of {nb}^3, 16 B/cell")
        del f
    if blocks_only:
        print(json.dumps(out))
        return
    # ---- K1b + K3 on the C2 mesh: 8^3 blocks, 4 levels -> 256^3
    mesh = synth.octree_mesh((4, 4, 4), (8, 8, 8), 4, seed=11, p_refine=0.5)
    leaves = np.flatnonzero(mesh.node_type == 1)
    blk = torch.rand((mesh.nblocks, 8, 8, 8), generator=g, device=dev, dtype=torch.float32)
    scale = (2 ** (mesh.lmax - mesh.level[leaves])).astype(np.int64)
    off = mesh.origin[leaves] * 8 * scale[:, None]
    table = device.prolong_table(leaves, off, scale)
    ms = timeit(lambda: device.prolong(blk, table, (256, 256, 256)))
    record("prolong_c2_256cubed_f32", ms, 4.0 * leaves.size * 512 + 8.0 * 256**3,
           f"fava_prolong incl. host lattice table + upload; {leaves.size} leaves of 8^3 over 4 levels -> 256^3 fp64")
    # ---- K3 at a size that fills the GPU: the same octree with 16^3 blocks -> 512^3
    mesh16 = synth.octree_mesh((4, 4, 4), (16, 16, 16), 4, seed=11, p_refine=0.5)
    leaves = np.flatnonzero(mesh16.node_type == 1)
    blk = torch.rand((mesh16.nblocks, 16, 16, 16), generator=g, device=dev, dtype=torch.float32)
    scale = (2 ** (mesh16.lmax - mesh16.level[leaves])).astype(np.int64)
    table = device.prolong_table(leaves, mesh16.origin[leaves] * 16 * scale[:, None], scale)
    out512 = torch.empty((512, 512, 512), dtype=torch.float64, device=dev)
    ms = timeit(lambda: device.prolong(blk, table, (512, 512, 512), out=out512))
    record("prolong_512cubed_f32", ms, 4.0 * leaves.size * 4096 + 8.0 * 512**3,
           f"fava_prolong (tables cached); {leaves.size} leaves of 16^3 over 4 levels -> 512^3 fp64")
    del blk, out512
    # ---- plane sums (slice_integral) 512^3 fp64
    fld = torch.rand((512, 512, 512), generator=g, device=dev, dtype=torch.float64)
    for axis in (0, 1, 2):
        ms = timeit(lambda: device.plane_sum(fld, axis))
        record(f"plane_sum_512cubed_f64_axis{axis}", ms, 8.0 * 512**3, "fava_plane_sum, 8 B/cell")
    print(json.dumps(out))


if __name__ == "__main__":
    main()
