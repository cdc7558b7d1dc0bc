"""Micro-benchmark of the slab exchange kernel alone (run under torchrun on N GPUs): GB/s per GPU per direction."""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fava_b200 import device, dist, spectrum  # noqa: E402


def main():
    rank, world, local = dist.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    n = int(os.environ.get("FAVA_A2A_N", "1024"))
    p = spectrum._plan(n, rank, world, dev)
    nominal = 16.0 * p.nzl * (n - 1) * p.pitch * (world - 1) / world  # bytes leaving this GPU, unpruned
    for _ in range(2):
        spectrum.exchange(p, 0)
    dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        spectrum.exchange(p, 0)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1) / 5], device=dev)
    dist.allreduce_max_(ms)
    if rank == 0:
        print(f"A2A world={world} n={n}: {ms.item():.3f} ms  nominal {nominal / ms.item() / 1e6:.0f} GB/s  "
              f"on-wire ~{0.79 * nominal / ms.item() / 1e6:.0f} GB/s", flush=True)
    dist.barrier()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
