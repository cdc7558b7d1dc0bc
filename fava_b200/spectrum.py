"""Slab-decomposed kinetic-energy spectrum (BASELINE configs[3]: 1024^3 at 1/2/4/8 B200).

Rank r holds z-planes [r*nzl, (r+1)*nzl) of rho, ux, uy, uz.
  stage 1: weighting fused with the x transform of the local planes, all three components in one pass;
  per component: stage 2 (y transform, in place) -> slab->pencil exchange (K5: ONE kernel gathers the ky rows each
  destination owns and stores them straight into that rank's receive buffer over NVLink peer mappings) ->
  stage 3 (z transform, pruned to the spectral sphere) ; then shell binning (K6) and one all-reduce of the
  [3][N/2-1] shell sums.  (Grid sizes that are not a power of two in [256, 2048] run the same schedule on cuFFT.)
Spectral space is distributed over ky cyclically in |ky| (every rank bins the same share of the spectral sphere); the
reference's `.T` projection (FlashUniform.py:281) is evaluated point by point without a transposed operand
(csrc/spectrum.cu), so no rank needs another rank's rows.
"""

from __future__ import annotations

import numpy as np
import torch

from fava_b200 import device, dist


def ky_ownership(n: int, nranks: int) -> np.ndarray:
    """[nranks][nyl] global ky indices owned by each rank (-1 = padding).  Rank r owns the wavenumbers
    |ky| = r, r + P, r + 2P, ... < N/2 (cyclic: every rank gets the same share of the spectral SPHERE, so the
    binning work is balanced) as index pairs (j, N-j), the non-negative rows first; the Nyquist row j = N/2 is
    owned by nobody (it lies beyond the last bin edge).  nyl = N/P."""
    if n % (2 * nranks):
        raise ValueError(f"grid size {n} must be divisible by 2 x {nranks} ranks")
    h = n // (2 * nranks)
    own = -np.ones((nranks, 2 * h), dtype=np.int32)
    for r in range(nranks):
        pos = np.arange(r, n // 2, nranks)
        neg = [(n - j) % n for j in pos if j != 0]
        rows = list(pos) + neg
        own[r, : len(rows)] = rows
    return own


def exchange_ctas(world: int, num_sms: int = 148) -> int:
    """Persistent CTAs (one per SM) of the y pass fused with the exchange.  The kernel turns out ~14.5 GB/s of pruned
    output per SM (measured: 1024^3 y pass in 3.08 ms on 148 SMs); (N-1)/N of it leaves the GPU and NVLink carries
    ~770 GB/s per direction, so more than 770 / (14.5 (N-1)/N) SMs only queue stores behind the link.  The other SMs
    run the plane-profile kernels, the completion tokens (NCCL needs an SM of its own) and the z passes of earlier
    components at the same time; the column kernels take their tiles from a counter, so sharing the GPU costs them no
    tail.  Measured at 2 GPUs with every SM given to the exchange: tokens and profile kernels queued behind it and the
    step was serial (26.4 ms against 23.3 ms for half the single-GPU time); at 8 GPUs 64 / 80 / 100 CTAs all need 1.14 ms
    per component (645 GB/s on the wire: the link is the bound) and the step takes 6.52 / 6.71 / 6.76 ms."""
    want = int(np.ceil(1.05 * 770.0 / (14.5 * (world - 1) / world)))
    return min(want, num_sms - 16)


WS_SEND = 8  # FAVA_WS_USER0 + {0,1,2}: per-component slab buffers (weight -> in-place 2-D FFT)
WS_RECV = 11  # + {0,1,2}: per-component ky-pencil receive buffers (peer-mapped on the other ranks)


class SlabPlan:
    """Buffers, ky ownership tables and peer mappings of one (N, world) configuration."""

    def __init__(self, n: int, rank: int, world: int, dev: torch.device):
        if n % world or n % (2 * world):
            raise ValueError(f"grid size {n} must be divisible by 2 x {world} ranks")
        self.n, self.rank, self.world, self.dev = n, rank, world, dev
        self.pitch = device.spectral_pitch(n)  # complex elements per kx row
        self.nzl = n // world
        own = ky_ownership(n, world)
        self.nyl = own.shape[1]
        self.ky_of_dest = torch.from_numpy(own).to(dev)
        self.ky_of_local = torch.from_numpy(own[rank].copy()).to(dev)
        owner = -np.ones(n, dtype=np.int32)  # rank that owns global ky row k, and k's row inside that rank's set
        row = np.zeros(n, dtype=np.int32)
        for r in range(world):
            held = np.flatnonzero(own[r] >= 0)
            owner[own[r][held]] = r
            row[own[r][held]] = held
        self.owner_of_ky = torch.from_numpy(owner).to(dev)
        self.row_of_ky = torch.from_numpy(row).to(dev)
        self.native = device.fft_native_supported(n)  # the y pass itself scatters its rows to their owners
        send_bytes = 16 * self.nzl * n * self.pitch
        recv_bytes = 16 * n * self.nyl * self.pitch
        self.send = [device.workspace(WS_SEND + c, send_bytes, dev) for c in range(3)]
        self.recv = [device.workspace(WS_RECV + c, recv_bytes, dev) for c in range(3)]
        # peer mappings of every rank's receive buffers (CUDA IPC; NVLink P2P under NVSwitch)
        handles = [device.ipc_export(p) for p in self.recv]
        gathered = [None] * world
        torch.distributed.all_gather_object(gathered, handles)
        self.peer_tables = []
        self._opened = []
        for c in range(3):
            ptrs = []
            for r in range(world):
                if r == rank:
                    ptrs.append(self.recv[c])
                else:
                    p = device.ipc_open(gathered[r][c])
                    self._opened.append(p)
                    ptrs.append(p)
            self.peer_tables.append(torch.tensor(ptrs, dtype=torch.int64, device=dev))
        self.sums = torch.zeros((3, n // 2 - 1), dtype=torch.float64, device=dev)
        self.tokens = [torch.zeros(1, dtype=torch.float32, device=dev) for _ in range(3)]
        self.comm_stream = torch.cuda.Stream(device=dev, priority=-1)  # NVLink exchange runs beside the HBM-bound kernels
        self.token_stream = torch.cuda.Stream(device=dev, priority=-1)  # completion tokens: off the pack kernels' stream
        self.ev_xy = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        self.ev_packed = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        self.ev_done = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        self.ev_mark = {k: torch.cuda.Event(enable_timing=True) for k in ("pieces", "overlap", "fft_z", "bin")}
        dist.barrier()


    def close(self) -> None:
        """Unmap the peers' buffers (collective: every rank drops the plan at the same point)."""
        torch.cuda.synchronize(self.dev)
        dist.barrier()
        lib = device._lib.load()
        for p in self._opened:
            lib.fava_ipc_close(device.C.c_void_p(p))
        self._opened = []
        dist.barrier()


_plans: dict = {}


def _plan(n: int, rank: int, world: int, dev) -> SlabPlan:
    """One live plan at a time: the exchange buffers are grow-only context workspaces, so a plan for another
    grid size may re-allocate them and would leave the peers' mappings of the old buffers dangling."""
    key = (n, rank, world, str(dev))
    if key not in _plans:
        for old in _plans.values():
            old.close()
        _plans.clear()
        _plans[key] = SlabPlan(n, rank, world, dev)
    return _plans[key]


def exchange(p: SlabPlan, c: int, z_offset: int = 0, nz_chunk: int | None = None, y_done: bool = False) -> None:
    """Stage 2 + slab -> ky-pencil exchange of component c on the current stream.
    Hand-written path: ONE kernel - the y pass of planes [z_offset, z_offset + nz_chunk) stores every output row straight
    into its owner's peer-mapped receive buffer over NVLink (fava_fft_y_scatter), so the transfer overlaps the transform
    tile by tile and neither a send-side copy of the y-transformed slab nor a pack pass exists.
    cuFFT path (grid not a power of two): in-place 2-D transform, then the pack kernel K5 streams the ky rows global ->
    shared -> the owner's buffer with TMA bulk copies (whole slab only)."""
    nz_chunk = p.nzl if nz_chunk is None else nz_chunk
    if p.native:
        src = p.send[c] + z_offset * device.spectral_bytes(p.n, 1)
        device.fft_y_scatter(src, p.n, nz_chunk, p.peer_tables[c], p.owner_of_ky, p.row_of_ky, p.rank, p.nzl, p.nyl,
                             z_offset=z_offset, max_ctas=exchange_ctas(p.world))
    else:
        if z_offset or nz_chunk != p.nzl:
            raise ValueError("the cuFFT path exchanges whole slabs")
        if not y_done:  # (stats.host_step has transformed the slab chunk by chunk already)
            device.ke_transform_y(p.send[c], p.nzl, p.n, p.dev)
        device.a2a_pack(p.send[c], p.peer_tables[c], p.ky_of_dest, p.rank, p.world, p.nzl, p.n, p.nyl)


def spectral_buffers(n: int, nz_local: int, dev) -> list[int]:
    """Device addresses of the three per-component buffers stages 1 + 2 of a z-slab fill (complex
    [nz_local][n][pitch] afterwards): the slab plan's send buffers on several ranks, the workspaces of
    fava_ke_spectrum on one."""
    world = dist.world_size()
    if world == 1:
        return [device.workspace(4 + c, device.spectral_bytes(n, n), dev) for c in range(3)]  # WS_FFT1 + c
    return _plan(n, dist.rank(), world, dev).send


def spectrum_from_transformed_slabs(n: int, dev, epilogue=None) -> dict[str, np.ndarray]:
    """Rest of the spectrum once `spectral_buffers` hold the 2-D transforms of this rank's planes (filled chunk
    by chunk while the slab was still arriving from the host, stats.host_step): exchange, z transforms, binning."""
    if dist.world_size() > 1:
        return slab_ke_spectrum(None, None, None, None, n, epilogue=epilogue, xy_done=True, dev=dev)
    w = spectral_buffers(n, n, dev)
    sums = torch.zeros((3, n // 2 - 1), dtype=torch.float64, device=dev)
    for c in range(3):
        device.ke_transform_z(w[c], n, n, None, dev)
    device.spectrum_bin(w[0], w[1], w[2], n, n, None, sums)
    if epilogue is not None:
        epilogue()
    return device.spectrum_finalize(sums, n)


def slab_ke_spectrum(rho, ux, uy, uz, n: int, overlap=None, epilogue=None, xy_done: bool = False,
                     dev=None) -> dict[str, np.ndarray]:
    """Spectrum of the global N^3 grid formed by the ranks' z-slabs; every rank returns the full dict.

    Schedule: the y pass + exchange of the three components (NVLink-bound, on a side stream with a share of the SMs)
    overlaps whatever `overlap()` enqueues on the calling stream (bench.py and
    stats.slab_step put the HBM-bound plane-profile kernels there); the z-transform of component c
    starts as soon as ITS exchange has completed on every rank.  `epilogue()` (the caller's collectives: they need
    the results of `overlap` only) is enqueued on the token stream, behind the last completion token, and joins the
    calling stream before the host synchronises on the shell sums."""
    world, rank = dist.world_size(), dist.rank()
    if world == 1:
        hooks = list(overlap) if isinstance(overlap, (list, tuple)) else [overlap]
        for hook in hooks + [epilogue]:
            if hook is not None:
                hook()
        return device.ke_spectrum(rho, ux, uy, uz)
    if not xy_done:
        nzl = int(rho.shape[0])
        if nzl * world != n or tuple(rho.shape[1:]) != (n, n):
            raise ValueError(f"rank {rank}: slab shape {tuple(rho.shape)} is not [{n // world}][{n}][{n}]")
        dev = rho.device
    p = _plan(n, rank, world, dev)
    cur = torch.cuda.current_stream(dev)
    if not xy_done:
        device.ke_transform_x(rho, ux, uy, uz, *p.send)
    for c in range(3):
        p.ev_xy[c].record(cur)  # (hand-written path: marks the end of the x pass; the y pass is part of the exchange)
        with torch.cuda.stream(p.comm_stream):
            p.comm_stream.wait_event(p.ev_xy[c])
            if not xy_done:
                exchange(p, c)
            elif not p.native:  # stats.host_step on the cuFFT path: 2-D transforms done chunk by chunk, rows still here
                exchange(p, c, y_done=True)
            # (hand-written path: stats.host_step scattered the rows chunk by chunk as the slab arrived)
            p.ev_packed[c].record(p.comm_stream)
        with torch.cuda.stream(p.token_stream):
            # every rank's stores of component c into my receive buffer are complete once all ranks have
            # passed this stream-ordered collective (each enqueues it after its own exchange kernel); it runs on its own
            # stream so that the exchange of component c+1 starts without waiting for it
            p.token_stream.wait_event(p.ev_packed[c])
            dist.allreduce_sum_(p.tokens[c])
            p.ev_done[c].record(p.token_stream)
    # `overlap` may be one callable or a list of them: the z-transform of component c is enqueued after the
    # c-th piece, so that it starts as soon as its exchange is done instead of queueing behind all the pieces
    pieces = list(overlap) if isinstance(overlap, (list, tuple)) else ([overlap] if overlap is not None else [])
    for c in range(3):
        if c < len(pieces):
            pieces[c]()
        if c == min(2, max(len(pieces) - 1, 0)):
            for extra in pieces[3:]:
                extra()
            p.ev_mark["pieces"].record(cur)
            if epilogue is not None:
                # the caller's collectives (profile all-reduces) depend on the pieces only: they go to the token stream,
                # where they run behind the last completion token and BESIDE the z passes and the binning instead of
                # after them (0.5 ms of the 7 ms step at 8 GPUs)
                with torch.cuda.stream(p.token_stream):
                    p.token_stream.wait_event(p.ev_mark["pieces"])
                    epilogue()
        if c == 2:
            p.ev_mark["overlap"].record(cur)
        cur.wait_event(p.ev_done[c])
        device.ke_transform_z(p.recv[c], n, p.nyl, p.ky_of_local, dev)
    p.ev_mark["fft_z"].record(cur)
    device.spectrum_bin(p.recv[0], p.recv[1], p.recv[2], n, p.nyl, p.ky_of_local, p.sums)
    # shell sums and counts add across ranks; this collective also fences the receive buffers against
    # the next call's remote stores (a rank's next exchange is stream-ordered after it)
    p.ev_mark["bin"].record(cur)
    dist.allreduce_sum_(p.sums)
    cur.wait_stream(p.token_stream)  # the epilogue's results are complete when the caller synchronises on this stream
    return device.spectrum_finalize(p.sums, n)
