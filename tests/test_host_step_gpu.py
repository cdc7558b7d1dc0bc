"""GPU: the streamed host -> HBM step (stats.host_step: chunked H2D on a side stream, every chunk consumed as it
lands) returns what the resident step (stats.slab_step) returns for the same snapshot."""
import numpy as np
import pytest
import torch

from fava_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("n,chunk,dtype", [(64, 16, np.float64), (64, 24, np.float32), (32, 64, np.float64)])
def test_host_step_equals_resident_step(cuda_device, n, chunk, dtype):
    from fava_b200 import stats

    f = synth.uniform_fields((n, n, n), names=("dens", "velx", "vely", "velz"), dtype=dtype, seed=31, u0=3.0)
    host = [torch.from_numpy(f[k].copy()).pin_memory() for k in ("dens", "velx", "vely", "velz")]
    dev = [h.to(cuda_device) for h in host]
    cv, lv = 1.0 / n**3, 1.0 / n
    ref = stats.slab_step(*dev, n, cv, lv)
    for rep in range(2):  # the second call re-uses plans and buffers
        got = stats.host_step(host, n, cv, lv, chunk_planes=chunk)
        assert set(got) == set(ref)
        for ax in (0, 1, 2):
            for k in ref[ax]:
                a, b = got[ax][k].cpu().numpy(), ref[ax][k].cpu().numpy()
                assert np.max(np.abs(a - b)) <= 1e-13 * np.max(np.abs(b)) + 1e-300, (ax, k, rep)
        for k in ref["spectrum"]:
            a, b = got["spectrum"][k], ref["spectrum"][k]
            assert np.max(np.abs(a - b)) <= 1e-13 * np.max(np.abs(b)), (k, rep)
    only = stats.host_step(host, n, cv, lv, axes=(1,), spectrum=False, chunk_planes=chunk)
    assert list(only) == [1]
    assert np.array_equal(only[1]["reynolds"].cpu().numpy(), got[1]["reynolds"].cpu().numpy())
