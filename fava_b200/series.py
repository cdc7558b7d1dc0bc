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


def streamed_reynolds_series(chunk_source, nfiles: int, planes: int, n: int, axis: int, cell_volume: float,
                             layer_volume: float, dev, chunk_planes: int = 16, favre: bool = False):
    """The same time series for snapshots that arrive from PINNED HOST memory: per file, this rank's z-slab
    [planes][n][n] of rho, ux, uy, uz is copied to HBM in chunks of `chunk_planes` planes on a side stream (two device
    buffers, so the copy of chunk i+1 runs under the moment kernel of chunk i) and accumulated about the pivots of the
    file's first chunk; one packed [14][n] all-reduce and the finalize per file (stats.slab_profiles_finish).
    `chunk_source(file, first_plane, nplanes)` returns the four pinned host tensors [nplanes][n][n] of a chunk.
    Returns the per-file profile dicts (device tensors)."""
    import torch

    from fava_b200 import device, stats

    if axis not in (0, 1):
        raise ValueError("streamed series: planes normal to x or y cross every chunk (axis 0 or 1)")
    cur = torch.cuda.current_stream(dev)
    copy_stream = stats._copy_stream(dev)
    first = chunk_source(0, 0, min(chunk_planes, planes))
    bufs = [[torch.empty((chunk_planes, n, n), dtype=h.dtype, device=dev) for h in first] for _ in range(2)]
    landed = [torch.cuda.Event() for _ in range(2)]
    consumed = [torch.cuda.Event() for _ in range(2)]
    for ev in consumed:
        ev.record(cur)
    results = []
    i = 0
    for f in range(nfiles):
        mom = piv = None
        for a in range(0, planes, chunk_planes):
            b = min(a + chunk_planes, planes)
            host = first if (f == 0 and a == 0) else chunk_source(f, a, b - a)
            slot = i % 2
            with torch.cuda.stream(copy_stream):
                copy_stream.wait_event(consumed[slot])
                for h, d in zip(host, bufs[slot]):
                    d[: b - a].copy_(h, non_blocking=True)
                landed[slot].record(copy_stream)
            cur.wait_event(landed[slot])
            part = [d[: b - a] for d in bufs[slot]]
            if mom is None:
                mom, piv = device.plane_moments(*part, axis)
            else:
                device.plane_moments(*part, axis, pivots=piv, out=mom, accumulate=True)
            consumed[slot].record(cur)
            i += 1
        results.append(stats.slab_profiles_finish(mom, piv, axis, cell_volume, layer_volume, favre=favre))
    return results


def streamed_series_benchmark(rank: int, world: int, dev, n: int = 2048, planes_per_rank: int = 256, files: int = 20,
                              chunk_planes: int = 16) -> dict:
    """BASELINE configs[4] (2048^3 Reynolds-stress time series over 20 plt files on 8 GPUs) as a streamed run: every rank
    owns `planes_per_rank` z-planes of n x n f32 cells per file (256 planes x 8 ranks = 2048^3; with fewer ranks the same
    per-GPU share, i.e. a shorter grid), streamed from pinned host memory through stats/plane-moment kernels with one
    all-reduce per file.  The host cannot hold 20 x 128 GiB, so ONE pinned chunk (4 fields x chunk_planes planes) is
    re-sent for every chunk of every file: the PCIe bytes, kernels and collectives are those of the real series, the
    page-cache / disk side of fava_stage_h2d is not part of this number (profiles/r01_series_staging.json has it)."""
    import torch

    from fava_b200 import device, dist as d

    g = torch.Generator()
    g.manual_seed(5)
    host = [(torch.rand((chunk_planes, n, n), generator=g, dtype=torch.float32) + (1.0 if i == 0 else -0.5)).pin_memory()
            for i in range(4)]

    def source(f, a, m):
        return [h[:m] for h in host]

    cv, lv = 1.0 / (float(n) * n * planes_per_rank * world), 1.0 / n
    streamed_reynolds_series(source, 1, 2 * chunk_planes, n, 0, cv, lv, dev, chunk_planes)  # warm-up
    d.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    res = streamed_reynolds_series(source, files, planes_per_rank, n, 0, cv, lv, dev, chunk_planes)
    host_out = [{k: v.cpu() for k, v in r.items()} for r in res]  # results of every file on the host
    torch.cuda.synchronize()
    d.barrier()
    dt = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
    d.allreduce_max_(dt)
    sec = float(dt.item())
    # every chunk holds the same cells, so the streamed profile must equal the profile of ONE chunk
    one = device.plane_profiles(*[h.to(dev) for h in host], 0, 1.0 / (float(n) * n * chunk_planes), lv, favre=False)
    err = 0.0
    for k in ("means", "reynolds"):
        a, b = host_out[-1][k].numpy(), one[k].cpu().numpy()
        err = max(err, float(abs(a - b).max() / abs(b).max()))
    bytes_rank = 4.0 * 4 * n * n * planes_per_rank * files
    cells = float(n) * n * planes_per_rank * world * files
    return {"workload": f"BASELINE configs[4]: {files} snapshots of {n} x {n} x {planes_per_rank * world} f32 cells "
                        f"({planes_per_rank} z-planes per GPU; 8 GPUs = 2048^3), reynolds_stress(axis=0) per snapshot, streamed "
                        f"from pinned host memory in {chunk_planes}-plane chunks (one pinned chunk re-sent, see docstring)",
            "n_gpus": world, "seconds": sec, "files": files, "gcells_per_s": cells / sec / 1e9,
            "h2d_gbs_per_gpu": bytes_rank / sec / 1e9, "h2d_gbs_total": bytes_rank * world / sec / 1e9,
            "streamed_vs_single_chunk_max_rel_err": err}
