"""The closed-form known-answer snapshots (fava_b200/knownanswer.py) that bench.py pushes through the GPU path at
every N agree with the oracle (itself pinned on the reference's goldens) at a size the oracle handles in seconds."""
import numpy as np
import torch

from fava_b200 import knownanswer as ka
from oracle import fava_oracle as orc
from tests._util import STRESS, maxnorm_close


def _fields(fill, n):
    f = [torch.empty((n, n, n), dtype=torch.float64) for _ in range(4)]
    fill(f, n, 0)
    return {k: t.numpy() for k, t in zip(("dens", "velx", "vely", "velz"), f)}


def test_profile_case_closed_forms_equal_the_oracle():
    n = 32
    file_fields = _fields(ka.fill_profile_case, n)
    data4 = {k: orc.load_like_reference(v)[None, ...] for k, v in file_fields.items()}
    geom = orc.uniform_geom((n, n, n), bbox_dtype=np.float64)
    for axis in (0, 1, 2):
        want = ka.expected_profiles(n, axis)
        _, stress, means = orc.reynolds_stress(geom, data4, axis=axis)
        fmeans, favre = orc.favre_stress(geom, data4, axis=axis)
        for row, key in enumerate(("dens", "velx", "vely", "velz")):
            assert np.max(np.abs(means[key] - want["means"][row])) < 1e-13, (axis, key)
        for row, key in enumerate(STRESS):
            assert np.max(np.abs(stress[key] - want["reynolds"][row])) < 1e-13, (axis, key)
            assert np.max(np.abs(favre[key] - want["favre"][row])) < 1e-13, (axis, key)
        for row, key in enumerate(("velx", "vely", "velz")):
            assert np.max(np.abs(fmeans[key] - want["favre_means"][row])) < 1e-13, (axis, key)


def test_profile_case_slab_fill_is_the_global_field():
    n = 16
    whole = [torch.empty((n, n, n), dtype=torch.float64) for _ in range(4)]
    ka.fill_profile_case(whole, n, 0)
    part = [torch.empty((4, n, n), dtype=torch.float64) for _ in range(4)]
    ka.fill_profile_case(part, n, 8)
    assert all(torch.equal(p, w[8:12]) for p, w in zip(part, whole))
    ka.fill_spectrum_case(whole, 32 if False else n, 0)
    ka.fill_spectrum_case(part, n, 8)
    assert all(torch.equal(p, w[8:12]) for p, w in zip(part, whole))


def test_spectrum_case_closed_form_equals_the_oracle():
    n = 32
    file_fields = _fields(ka.fill_spectrum_case, n)
    data3 = {k: orc.load_like_reference(v) for k, v in file_fields.items()}
    sp = orc.kinetic_energy_spectra(data3, (n, n, n))
    maxnorm_close(sp["total"], ka.expected_spectrum_total(n), 1e-12, "total")
    errs = ka.spectrum_errors(sp, n)
    assert max(errs.values()) < 1e-12, errs


def test_profile_errors_reports_slab_local_axis2():
    n = 16
    res = {}
    for ax in (0, 1, 2):
        want = ka.expected_profiles(n, ax)
        res[ax] = {k: torch.from_numpy(v[:, 4:8].copy() if ax == 2 else v.copy()) for k, v in want.items()}
    errs = ka.profile_errors(res, n, (0, 1, 2), 4, 4)
    assert max(errs.values()) == 0.0
    res[1]["reynolds"][0, 3] += 1e-6
    assert ka.profile_errors(res, n, (0, 1, 2), 4, 4)["axis1.reynolds[0]"] > 1e-7
