"""Process-group plumbing: the B200 replacement for the reference's FAVA_MPI singleton
(fava/util/_mpi.py:7-80).  One process per GPU, torch.distributed over NCCL (gloo on CPU for the
host-logic tests); with no process group everything degenerates to a single rank.

The grid is partitioned into z-slabs (contiguous byte ranges of the [z][y][x] file layout); block
datasets are partitioned into contiguous block ranges exactly like the reference
(_mpi_assign_blocks / _mpi_get_global_block_id, _flash.py:166-208).
"""

from __future__ import annotations

import os

import torch
import torch.distributed as dist


def initialized() -> bool:
    return dist.is_available() and dist.is_initialized()


def world_size() -> int:
    return dist.get_world_size() if initialized() else 1


def rank() -> int:
    return dist.get_rank() if initialized() else 0


def is_root() -> bool:
    return rank() == 0


def init_from_env(backend: str | None = None) -> tuple[int, int, int]:
    """Join the torchrun rendezvous if WORLD_SIZE > 1.  Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29500")
        if backend == "nccl":
            torch.cuda.set_device(local)
            dist.init_process_group(backend=backend, device_id=torch.device("cuda", local))
        else:
            dist.init_process_group(backend=backend)
    return rank(), world_size(), local


def bind_to_gpu_numa(local_gpu: int) -> bool:
    """Pin this process to the CPUs next to its GPU (NVML's ideal affinity), so that pinned staging buffers allocated
    afterwards live on the GPU's NUMA node and the H2D copies of several ranks do not cross the socket link.
    Best effort: returns False when NVML is not importable or refuses."""
    try:
        import pynvml

        pynvml.nvmlInit()
        handle = pynvml.nvmlDeviceGetHandleByIndex(int(local_gpu))
        pynvml.nvmlDeviceSetCpuAffinity(handle)
        return True
    except Exception:
        return False


def parallel_range(iterations: int, r: int | None = None, p: int | None = None) -> tuple[int, int]:
    """Contiguous block split with the remainder on the low ranks (fava/util/_mpi.py:68-77 and
    _flash.py:166-187 use the same rule)."""
    r = rank() if r is None else r
    p = world_size() if p is None else p
    extra = iterations % p
    local = iterations // p
    if r < extra:
        local += 1
        start = local * r
    else:
        start = extra * (local + 1) + (r - extra) * local
    return start, start + local


def allreduce_sum_(t: torch.Tensor) -> torch.Tensor:
    if world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t


def allreduce_max_(t: torch.Tensor) -> torch.Tensor:
    if world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t


def broadcast_(t: torch.Tensor, src: int = 0) -> torch.Tensor:
    if world_size() > 1:
        dist.broadcast(t, src=src)
    return t


def all_gather_cat(t: torch.Tensor, dim: int = 0) -> torch.Tensor:
    """Concatenate the per-rank tensors along `dim` in rank order.  The parts may differ in length along `dim`
    (z-slabs of a grid whose height is not a multiple of the rank count, `parallel_range`): lengths are exchanged
    first and the parts padded to the longest for the collective."""
    if world_size() == 1:
        return t
    n = torch.tensor([t.shape[dim]], dtype=torch.int64, device=t.device)
    counts = [torch.zeros_like(n) for _ in range(world_size())]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    longest = max(counts)
    src = t.contiguous()
    if t.shape[dim] < longest:
        pad_shape = list(t.shape)
        pad_shape[dim] = longest - t.shape[dim]
        src = torch.cat([src, torch.zeros(pad_shape, dtype=t.dtype, device=t.device)], dim=dim).contiguous()
    parts = [torch.empty_like(src) for _ in range(world_size())]
    dist.all_gather(parts, src)
    return torch.cat([p.narrow(dim, 0, c) for p, c in zip(parts, counts)], dim=dim)


def barrier() -> None:
    if world_size() > 1:
        dist.barrier()


def gather_cat_to_root(t: torch.Tensor) -> torch.Tensor | None:
    """Concatenate per-rank tensors along dim 0 on rank 0 (parts may differ in length); None elsewhere."""
    if world_size() == 1:
        return t
    n = torch.tensor([t.shape[0]], dtype=torch.int64, device=t.device)
    counts = [torch.zeros_like(n) for _ in range(world_size())]
    dist.all_gather(counts, n)
    counts = [int(c.item()) for c in counts]
    if is_root():
        parts = [torch.empty((c,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device) for c in counts]
        parts[0] = t
        for r in range(1, world_size()):
            dist.recv(parts[r], src=r)
        return torch.cat(parts, dim=0)
    dist.send(t.contiguous(), dst=0)
    return None


def all_gather_rows(t: torch.Tensor) -> list[torch.Tensor]:
    """All ranks' equally-shaped tensors, in rank order."""
    if world_size() == 1:
        return [t]
    parts = [torch.empty_like(t) for _ in range(world_size())]
    dist.all_gather(parts, t.contiguous())
    return parts
