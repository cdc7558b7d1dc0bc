"""Per-kernel timing of the native FFT path against the cuFFT path at N^3 (1 GPU)."""
import os
import sys
from pathlib import Path

import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))
from fava_b200 import device  # noqa: E402


def timeit(fn, reps=3):
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
    os.environ["FAVA_FFT"] = "native"
    n = int(os.environ.get("FAVA_FFT_N", "1024"))
    dev = torch.device("cuda", 0)
    nxh = n // 2 + 1
    g = torch.Generator(device=dev)
    g.manual_seed(1)
    f = [torch.rand((n, n, n), generator=g, device=dev, dtype=torch.float64) + 0.5 for _ in range(4)]
    w = [device.workspace(4 + c, 16 * n * n * nxh, dev) for c in range(3)]
    gb = 1e-9
    for threads in ("plain", "256", "384", "512"):
        if threads == "plain":
            os.environ["FAVA_FFT_X"] = "plain"
        else:
            os.environ.pop("FAVA_FFT_X", None)
            os.environ["FAVA_FFT_X_THREADS"] = threads
        t = timeit(lambda: device.fft_x_weight3(*f, *w))
        print(f"FFT n={n} x_weight3 [{threads}] (3 comps): {t:.3f} ms  {(32 + 24.05) * n**3 * gb / t * 1e3:.0f} GB/s")
    os.environ.pop("FAVA_FFT_X_THREADS", None)
    t = timeit(lambda: device.fft_cols(w[0], n, nxh, n, dev))
    print(f"FFT n={n} cols y (1 comp): {t:.3f} ms  {2 * 16 * n * n * nxh * gb / t * 1e3:.0f} GB/s")
    t = timeit(lambda: device.fft_cols(w[0], n, n * nxh, 1, dev, prune_grid_n=n))
    print(f"FFT n={n} cols z pruned (1 comp): {t:.3f} ms  {2 * 16 * n * n * nxh * gb / t * 1e3:.0f} GB/s (unpruned bytes)")
    t = timeit(lambda: device.fft_cols(w[0], n, n * nxh, 1, dev))
    print(f"FFT n={n} cols z full (1 comp): {t:.3f} ms  {2 * 16 * n * n * nxh * gb / t * 1e3:.0f} GB/s")
    t = timeit(lambda: device.ke_weight3(*f, *w))
    print(f"cuFFT path: ke_weight3 {t:.3f} ms")
    t = timeit(lambda: device.fft_xy(w[0], n, n, n, dev))
    print(f"cuFFT path: fft_xy (1 comp) {t:.3f} ms")
    t = timeit(lambda: device.fft_z(w[0], n, n * nxh, dev))
    print(f"cuFFT path: fft_z (1 comp) {t:.3f} ms")


if __name__ == "__main__":
    main()
