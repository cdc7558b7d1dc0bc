#!/usr/bin/env python
"""End-to-end step from pinned host buffers at the bench workload (1 x B200): the whole slab copied first and the
resident step afterwards ("serial") against the streamed step (stats.host_step, chunked H2D overlapped with the
kernels).  CUDA-event times; prints one JSON line."""
import json
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))


def main():
    import torch

    import bench
    from fava_b200 import stats

    n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    dev = torch.device("cuda", 0)
    torch.cuda.set_device(0)
    fields = bench.synth_slab_device(n, 0, n, dev)
    host = [torch.empty(t.shape, dtype=t.dtype, pin_memory=True) for t in fields]
    for h, d in zip(host, fields):
        h.copy_(d)
    cv, lv = 1.0 / float(n) ** 3, 1.0 / float(n)

    def serial():
        for h, s in zip(host, fields):
            s.copy_(h, non_blocking=True)
        return stats.slab_step(*fields, n, cv, lv)

    out = {"n": n}
    variants = {"serial": serial}
    for chunk in (32, 64, 128):
        variants[f"streamed_chunk{chunk}"] = (lambda c: lambda: stats.host_step(host, n, cv, lv, chunk_planes=c, stage=fields))(chunk)
    for name, fn in variants.items():
        res = fn()
        {k: v.cpu() for k, v in res[0].items()}
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(3):
            res = fn()
            for ax in (0, 1, 2):
                {k: v.cpu() for k, v in res[ax].items()}
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 3
        out[name] = {"ms": ms, "gcells_per_s": float(n) ** 3 / ms / 1e6, "h2d_gbs": sum(h.numel() * h.element_size() for h in host) / ms / 1e6}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
