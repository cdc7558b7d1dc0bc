"""Time series of plane statistics over the plt files of a FLASH run — the hot loop of the reference's
pipeline stage `Pipeline.reynolds_stress` (fava/__main__.py:76-97: for every plt file, load the mesh and call
`reynolds_stress`), without its result caching / flame-window fit (host-only driver logic, out of scope).

Every file is sharded over the ranks as contiguous block ranges (z-slabs for single-block files), staged
through the pinned ring straight into HBM and reduced by the block-list kernels; per file the ranks exchange
one packed [14][N] all-reduce.  The loop is staging-bound (PCIe / page cache), the kernels take a few per cent.
"""

from __future__ import annotations

import time

from fava_b200 import dist


def reynolds_series(model, axis: int = 0, file_type: str = "plt", indices=None, favre: bool = False, fields=None):
    """[(time, radius, stress, means), ...] for the selected files of `model` (a fava_b200.FLASH model).

    Also returns a timing dict: bytes staged on this rank, seconds spent staging / in the statistics call."""
    n = model.nfiles(file_type=file_type)
    idx = list(range(n)) if indices is None else list(indices)
    names = list(fields) if fields is not None else ["dens", "velx", "vely", "velz"]
    out = []
    t_stage = t_stat = 0.0
    nbytes = 0
    for i in idx:
        model.load(file_index=i, file_type=file_type)
        mesh = model.mesh
        t0 = time.perf_counter()
        mesh.load_data(names)
        import torch

        torch.cuda.synchronize()
        t1 = time.perf_counter()
        res = mesh.favre_stress(axis=axis) if favre else mesh.reynolds_stress(axis=axis)
        t2 = time.perf_counter()
        nbytes += sum(t.numel() * t.element_size() for t in mesh._dev.values())
        t_stage += t1 - t0
        t_stat += t2 - t1
        out.append((float(mesh.time),) + tuple(res))
    timing = {"files": len(idx), "staged_bytes_this_rank": int(nbytes), "stage_s": t_stage, "stats_s": t_stat,
              "stage_gbs_this_rank": nbytes / max(t_stage, 1e-12) / 1e9, "ranks": dist.world_size()}
    return out, timing
