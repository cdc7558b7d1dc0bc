"""Multi-rank GPU parity worker (launched by tests/test_multigpu.py under torchrun, one rank per GPU).

Every rank computes the single-GPU answer on its own device from the full synthetic arrays and compares it
with the answer of the slab / block-range decomposed path (<= 1e-13, counts identical).
"""
import os
import sys
import tempfile
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT))

import fava_b200  # noqa: E402
from fava_b200 import device, dist, spectrum, stats, synth  # noqa: E402


def close(a, b, rtol, what):
    a, b = np.asarray(a), np.asarray(b)
    err = np.max(np.abs(a - b))
    scale = np.max(np.abs(b))
    assert err <= rtol * scale + 1e-300, f"{what}: {err:.3e} > {rtol:g} * {scale:.3e}"


def main():
    rank, world, local = dist.init_from_env("nccl")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    n = int(os.environ.get("FAVA_MGPU_N", "64"))
    shape = (n, n, n)
    full = synth.uniform_fields(shape, names=("dens", "velx", "vely", "velz"), seed=2024, u0=2.0)
    tf = [torch.from_numpy(full[k].copy()).to(dev) for k in ("dens", "velx", "vely", "velz")]
    z0, z1 = dist.parallel_range(n)
    ts = [t[z0:z1].contiguous() for t in tf]

    # --- Reynolds / Favre profiles: z-slabs + one packed all-reduce (axis x, y) / all-gather (axis z)
    cv, lv = 1.0 / n**3, 1.0 / n
    for axis in (0, 1, 2):
        one = device.plane_profiles(*tf, axis, cv, lv)
        many = stats.slab_profiles(*ts, axis, cv, lv)
        for k in one:
            close(many[k].cpu().numpy(), one[k].cpu().numpy(), 1e-13, f"slab profiles {k} axis {axis}")

    # --- kinetic-energy spectrum: slab FFT with the fused NVLink exchange
    one = device.ke_spectrum(*tf)
    for rep in range(2):  # second call re-uses plan, buffers and peer mappings
        many = spectrum.slab_ke_spectrum(*ts, n)
        for k in one:
            close(many[k], one[k], 1e-13, f"slab spectrum {k} (call {rep})")

    # a different grid size in the same process: the plan (buffers, peer mappings) is replaced collectively
    n2 = n // 2
    full2 = synth.uniform_fields((n2, n2, n2), names=("dens", "velx", "vely", "velz"), seed=77)
    tf2 = [torch.from_numpy(full2[k].copy()).to(dev) for k in ("dens", "velx", "vely", "velz")]
    a2, b2 = dist.parallel_range(n2)
    one2 = device.ke_spectrum(*tf2)
    many2 = spectrum.slab_ke_spectrum(*[t[a2:b2].contiguous() for t in tf2], n2)
    for k in one2:
        close(many2[k], one2[k], 1e-13, f"slab spectrum {k} (second grid size)")
    # ... and streamed from the host on that grid size too (cuFFT path: the rows are exchanged after the last chunk)
    host2 = [t[a2:b2].contiguous().cpu().pin_memory() for t in tf2]
    hs2 = stats.host_step(host2, n2, 1.0 / n2**3, 1.0 / n2, chunk_planes=max(1, (b2 - a2) // 2))
    for k in one2:
        close(hs2["spectrum"][k], one2[k], 1e-13, f"host_step spectrum {k} (second grid size)")
    step = stats.slab_step(*ts, n, cv, lv)
    for k in one:
        close(step["spectrum"][k], one[k], 1e-13, f"slab_step spectrum {k}")
    for axis in (0, 1):
        ref_ax = device.plane_profiles(*tf, axis, cv, lv)
        for k in ref_ax:
            close(step[axis][k].cpu().numpy(), ref_ax[k].cpu().numpy(), 1e-13, f"slab_step profiles {k} axis {axis}")
    # the same step streamed from pinned host memory in chunks (stats.host_step)
    host = [t.cpu().pin_memory() for t in ts]
    hs = stats.host_step(host, n, cv, lv, chunk_planes=max(1, (z1 - z0) // 3))
    for k in one:
        close(hs["spectrum"][k], one[k], 1e-13, f"host_step spectrum {k}")
    for axis in (0, 1, 2):
        for k in step[axis]:
            close(hs[axis][k].cpu().numpy(), step[axis][k].cpu().numpy(), 1e-13, f"host_step profiles {k} axis {axis}")

    # --- block datasets: contiguous block ranges per rank, file -> staging -> kernels
    tmp = Path(tempfile.gettempdir()) / f"fava_mgpu_{os.environ.get('MASTER_PORT', '0')}"
    if rank == 0:
        tmp.mkdir(exist_ok=True)
        mesh = synth.octree_mesh((2, 2, 2), (8, 8, 8), 3, seed=5, p_refine=0.35)
        fields = synth.block_fields(mesh, names=("dens", "velx", "vely", "velz"), seed=7)
        synth.write_flash_file(tmp / "amr_hdf5_plt_cnt_0000", mesh, fields)
        np.save(tmp / "nblocks.npy", np.array([mesh.nblocks]))
        synth.write_flash_file(tmp / "uni_hdf5_uniform_0000", synth.single_block_mesh(shape),
                               {k: full[k].astype(np.float32) for k in full}, uniform3d=True)
    dist.barrier()
    m = fava_b200.mesh.FLASH(tmp / "amr_hdf5_plt_cnt_0000")
    m.load()
    assert m.nblocks_local == dist.parallel_range(int(m.nblocks))[1] - dist.parallel_range(int(m.nblocks))[0]
    res = {}
    for axis in (0, 1, 2):
        res[axis] = m.reynolds_stress(axis=axis)
    # single-rank answer: every rank redoes the whole file alone (no process group involvement)
    saved = (dist.world_size, dist.rank)
    dist.world_size, dist.rank = (lambda: 1), (lambda: 0)
    try:
        s = fava_b200.mesh.FLASH(tmp / "amr_hdf5_plt_cnt_0000")
        s.load()
        ref = {axis: s.reynolds_stress(axis=axis) for axis in (0, 1, 2)}
        s2 = fava_b200.mesh.FLASH(tmp / "amr_hdf5_plt_cnt_0000")
        s2.load()
        s2.from_amr(np.array([[0.0, 1.0], [0.0, 1.0], [0.0, 1.0]]), fields=["dens"], filename=tmp / f"one{rank}_hdf5_uniform_0000")
        uni_one = s2.data("dens")
        u1 = fava_b200.mesh.FlashUniform(tmp / "uni_hdf5_uniform_0000")
        u1.load()
        fd_one = u1.fractal_dimension("velx", 0.1)
        np.random.seed(3)
        sf_one = u1.structure_functions(num_seps=4, num_points=2000, sep_bounds=[0.01, 0.4])
    finally:
        dist.world_size, dist.rank = saved
    # --- uniform-grid analyses: tile-aligned plane ranges + halo (box counting), slab gathers (structure functions)
    um = fava_b200.mesh.FlashUniform(tmp / "uni_hdf5_uniform_0000")
    um.load()
    fd_many = um.fractal_dimension("velx", 0.1)
    assert fd_many == fd_one, f"box counting over {world} ranks differs: {fd_many} != {fd_one}"
    np.random.seed(3)
    sf_many = um.structure_functions(num_seps=4, num_points=2000, sep_bounds=[0.01, 0.4])
    for kind in ("longitudinal", "transverse"):
        for o, v in sf_one[kind].items():
            assert np.array_equal(sf_many[kind][o], v), f"structure functions {kind} {o} differ over {world} ranks"
    for axis in (0, 1, 2):
        for a, b in zip(res[axis][1].values(), ref[axis][1].values()):
            close(a, b, 1e-13, f"block-range stress axis {axis}")
        for a, b in zip(res[axis][2].values(), ref[axis][2].values()):
            close(a, b, 1e-13, f"block-range means axis {axis}")

    # --- from_amr: blocks sharded by file order, each rank fills its z-slab of the uniform array
    m2 = fava_b200.mesh.FLASH(tmp / "amr_hdf5_plt_cnt_0000")
    m2.load()
    m2.from_amr(np.array([[0.0, 1.0], [0.0, 1.0], [0.0, 1.0]]), fields=["dens"], filename=tmp / "many_hdf5_uniform_0000")
    nz = uni_one.shape[2]
    a, b = dist.parallel_range(nz)
    mine = m2.data("dens")  # [NX][NY][nz_local]
    assert np.array_equal(mine, uni_one[:, :, a:b]), "sharded from_amr slab differs"
    dist.barrier()
    if rank == 0:
        from fava_b200 import h5lite

        with h5lite.File(tmp / "many_hdf5_uniform_0000") as f, h5lite.File(tmp / "one0_hdf5_uniform_0000") as g:
            assert np.array_equal(f["dens"][()], g["dens"][()]), "uniform file written by N ranks differs"
        print(f"MGPU_OK world={world} n={n}", flush=True)
    dist.barrier()
    torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
