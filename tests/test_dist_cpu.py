"""CPU (gloo, world_size 2): host-side logic of the N>1 path — block / slab partitioning, the collectives
wrappers, ky ownership of the distributed FFT.  The kernels themselves are covered on GPUs by
tests/test_multigpu.py."""
import os
import socket

import pytest
import torch
import torch.multiprocessing as mp

from fava_b200 import dist, spectrum


def _free_port() -> int:
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank: int, world: int, port: int, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world),
                      LOCAL_RANK=str(rank))
    try:
        r, w, _ = dist.init_from_env("gloo")
        assert (r, w) == (rank, world) and dist.is_root() == (rank == 0)
        # contiguous ranges, remainder on the low ranks (reference _mpi.py:68-77)
        assert dist.parallel_range(7) == ((0, 4) if rank == 0 else (4, 7))
        assert dist.parallel_range(1) == ((0, 1) if rank == 0 else (1, 1))
        t = torch.full((3, 4), float(rank + 1), dtype=torch.float64)
        assert torch.equal(dist.allreduce_sum_(t.clone()), torch.full((3, 4), 3.0, dtype=torch.float64))
        assert torch.equal(dist.allreduce_max_(t.clone()), torch.full((3, 4), 2.0, dtype=torch.float64))
        b = dist.broadcast_(t.clone(), src=0)
        assert torch.equal(b, torch.full((3, 4), 1.0, dtype=torch.float64))
        cat = dist.all_gather_cat(t, dim=1)
        assert cat.shape == (3, 8) and torch.equal(cat[:, :4], torch.ones(3, 4, dtype=torch.float64))
        # uneven z-slabs (7 planes over 2 ranks: 4 + 3): axis-z profiles [rows][local planes] concatenate ragged
        a, b2 = dist.parallel_range(7)
        prof = torch.arange(a, b2, dtype=torch.float64).repeat(2, 1)
        whole = dist.all_gather_cat(prof, dim=1)
        assert whole.shape == (2, 7) and torch.equal(whole[0], torch.arange(7, dtype=torch.float64))
        assert torch.equal(dist.all_gather_cat(torch.arange(a, b2)), torch.arange(7))
        rows = dist.all_gather_rows(t)
        assert len(rows) == 2 and float(rows[1][0, 0]) == 2.0
        # ragged gather to the root (z-slabs / block ranges of different length)
        mine = torch.arange((rank + 2) * 3, dtype=torch.float32).reshape(rank + 2, 3) + 100 * rank
        full = dist.gather_cat_to_root(mine)
        if rank == 0:
            assert full.shape == (5, 3) and float(full[2, 0]) == 100.0
        else:
            assert full is None
        dist.barrier()
        torch.distributed.destroy_process_group()
        q.put((rank, "ok"))
    except Exception as exc:  # pragma: no cover - reported to the parent
        q.put((rank, repr(exc)))


def test_gloo_world2_collectives_and_partitioning():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    results = dict(q.get(timeout=120) for _ in procs)
    for p in procs:
        p.join(timeout=60)
    assert results == {0: "ok", 1: "ok"}, results


def test_single_process_degenerates_to_one_rank():
    assert dist.world_size() == 1 and dist.rank() == 0 and dist.parallel_range(10) == (0, 10)
    t = torch.ones(2)
    assert dist.allreduce_sum_(t) is t and dist.all_gather_cat(t) is t and dist.gather_cat_to_root(t) is t


@pytest.mark.parametrize("n,world", [(16, 2), (64, 4), (64, 8), (1024, 8), (96, 2)])
def test_ky_ownership_is_symmetric_and_complete(n, world):
    own = spectrum.ky_ownership(n, world)
    assert own.shape == (world, n // world)
    held = own[own >= 0]
    assert sorted(held) == [j for j in range(n) if j != n // 2]  # every ky row once, Nyquist dropped
    for r in range(world):
        rows = set(own[r][own[r] >= 0].tolist())
        assert all(((n - j) % n) in rows for j in rows)  # +-ky on the same rank
        h = n // (2 * world)
        assert all(0 <= j < n // 2 for j in own[r][:h])  # the non-negative wavenumbers come first
        assert sorted(own[r][:h] % world) == [r] * h  # cyclic in |ky|: balanced share of the spectral sphere
    with pytest.raises(ValueError):
        spectrum.ky_ownership(20, 8)
