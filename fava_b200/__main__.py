"""`python -m fava_b200` — the batch pipeline of the reference (fava/__main__.py:22-290) around the B200 hot path:

    per plt file   : reynolds_stress  -> cached in <stem>_analysis_NNNN; flame window from slice_average + fit
    all plt files  : linear fit of the window trajectory
    per plt file   : from_amr of the moving window -> <stem>_uniform_NNNN
    per uniform    : fractal_dimension, structure_functions, kinetic_energy_spectra

Settings come from ./pipeline_settings.json (same keys as the reference's fava/pipeline_settings.json); progress
is checkpointed to ./fava.checkpoint (JSON, next index per stage) and written on SIGINT / SIGTERM as well.
Run under torchrun for several GPUs (one process per GPU)."""

from __future__ import annotations

import copy
import json
import logging
import signal
import sys
from pathlib import Path

import numpy as np

from fava_b200 import dist, h5lite
from fava_b200.model import FLASH
from fava_b200.util import timer

LOGGER = logging.getLogger(__name__)
CHECKPOINT_NAME = "fava.checkpoint"
SETTINGS_NAME = "pipeline_settings.json"


class InterruptHandler:
    """Context manager: on SIGINT / SIGTERM call `external_handler` once (reference fava/util/_mpi.py:83-136)."""

    signals_caught = (signal.SIGINT, signal.SIGTERM)

    def __init__(self, external_handler=None):
        self.external_handler = external_handler
        self.interrupted = False
        self.released = False
        self.signal = None

    def __enter__(self):
        self.original = {sig: signal.getsignal(sig) for sig in self.signals_caught}

        def handler(signum, frame):
            if dist.is_root():
                print("Caught SIGINT..." if signum == signal.SIGINT else "Caught SIGTERM...", flush=True)
            self.signal = signum
            self.release()
            self.interrupted = True

        for sig in self.signals_caught:
            signal.signal(sig, handler)
        return self

    def __exit__(self, *exc):
        self.release()
        for sig, old in self.original.items():
            signal.signal(sig, old)

    def release(self) -> bool:
        if self.released:
            return False
        if self.external_handler is not None:
            if dist.is_root():
                print("Calling external handler", flush=True)
            self.external_handler()
        self.released = True
        return True


class Pipeline:
    def __init__(self, workdir=None) -> None:
        self.workdir = Path(workdir) if workdir is not None else Path.cwd()
        self.checkpoint_data: dict = {}

    # ---- settings / checkpoint (reference __main__.py:27-74) ----------------------------------------
    def load_settings(self, settings_path) -> None:
        with Path(settings_path).open("r") as f:
            self.settings = json.load(f)
        self.checkpoint_data["settings"] = copy.deepcopy(self.settings)
        for key, kind in (("basename", str), ("dimension", int), ("model", str), ("data folder", str), ("output folder", str)):
            assert key in self.settings and isinstance(self.settings[key], kind), f"pipeline setting {key!r}"
        self.data_dir = Path(self.settings["data folder"])
        self.output_dir = Path(self.settings["output folder"])
        win = self.settings.get("flame window") or {}
        self.half_width = float(win.get("half width", 16e5))  # the reference hard-codes 16e5 / 32e5 (cm)
        self.length = float(win.get("length", 2.0 * self.half_width))
        self.model = FLASH(self.data_dir)

    def checkpoint(self) -> None:
        if dist.is_root():
            with (self.workdir / CHECKPOINT_NAME).open("w") as f:
                json.dump(self.checkpoint_data, f, ensure_ascii=True, indent=4, default=float)

    def restart(self) -> None:
        ck = self.workdir / CHECKPOINT_NAME
        if ck.is_file():
            with ck.open("r") as f:
                self.checkpoint_data = json.load(f)
        self.load_settings(self.workdir / SETTINGS_NAME)

    def refresh_model(self) -> None:
        self.model = FLASH(self.data_dir)

    def _flam_or_rpv1(self) -> bool:
        mesh = self.model.mesh
        for name in ("rpv1", "flam"):
            if name in mesh.fields:
                self.flam = name
                return True
        return False

    # ---- stage 1: Reynolds stresses + flame window per plt file (reference :76-144) --------------------
    def reynolds_stress(self, index: int) -> None:
        self.model.load(file_index=index, file_type="plt")
        fn = self.output_dir / self.model.convert_filename_type("plt", "anl").stem
        if dist.is_root():
            print("REYNOLDS STRESS: ", fn, flush=True)
        pkey, skey = "reynolds stresses", "scalars"
        try:  # cached result of an earlier run
            with h5lite.File(fn, "r") as f:
                x = f[pkey]["radius"][()]
                s = {k: f[pkey]["tensor"][k][()] for k in f[pkey]["tensor"].keys()}
        except Exception:
            x, s, m = self.model.reynolds_stress()
            if dist.is_root():
                self.model.save_to_hdf5({pkey: {"tensor": s, "radius": x, "means": m}}, fn)
        if not self._flam_or_rpv1():
            return
        span, alp = self.model.slice_average(self.flam, axis=0)
        ccx = 0.5 * (x[1:] + x[:-1])
        mask = np.argwhere((0.0 < alp) & (alp < 1.0)).flatten()
        centroid = self.model.mesh.flame_window(ccx, s, mask)
        mesh = self.model.mesh
        left, right = mesh.domain_bounds[:, 0].copy(), mesh.domain_bounds[:, 1].copy()
        dx = float((self.settings.get("flame window") or {}).get("dx", 0.0))
        left[0], right[0] = centroid - self.half_width + dx, centroid + self.half_width + dx
        dims = ((right - left) / mesh.get_minimum_deltas(axis=1)).astype(int)
        if dist.is_root():
            print("Flame Window: ", right, dims, flush=True)
            self.model.save_to_hdf5({skey: {"time": mesh.time, "window left": left, "window right": right,
                                            "window dimensions": dims}}, fn)
        dist.barrier()

    # ---- stage 2: smooth the window trajectory (reference :146-165) --------------------------------------
    def smooth_window_trajectory(self) -> None:
        n = self.model.nfiles(file_type="plt")
        self.xmax, self.time = np.zeros(n), np.zeros(n)
        for i, p in enumerate(sorted(self.model.plt_files["by index"].keys())):
            self.model.load(file_index=p, file_type="plt")
            fn = self.output_dir / self.model.convert_filename_type("plt", "anl").stem
            with h5lite.File(fn, "r") as f:
                self.xmax[i] = f["scalars"]["window right"][()][0]
            self.time[i] = self.model.mesh.time
        coef = np.polyfit(self.time, self.xmax, 1) if n > 1 else np.array([0.0, self.xmax[0]])
        self.t0, self.x0 = self.time[0], self.xmax[0]
        self.func = np.poly1d(coef)

    # ---- stage 3: AMR -> uniform window per plt file (reference :167-186) ----------------------------------
    def extract_windows(self, index: int) -> None:
        self.model.load(file_index=index, file_type="plt")
        if not self._flam_or_rpv1():
            return
        mesh = self.model.mesh
        xmax = self.x0 + (self.func(mesh.time) - self.func(self.t0))
        hw = self.half_width
        sub = np.array([[xmax - self.length, xmax], [-hw, hw], [-hw, hw]])
        wanted = [self.flam, "dens", "pres", "temp", "velx", "vely", "velz", "divv", "igtm", "vort"]
        fields = [k for k in wanted if k in mesh.fields]  # the reference assumes all ten exist
        fn = self.output_dir / self.model.convert_filename_type("plt", "uni").stem
        if dist.is_root():
            print("EXTRACT: ", fn, flush=True)
        if fn.is_file():
            return
        mesh.from_amr(subdomain_coords=sub, fields=fields, filename=fn)
        dist.barrier()

    # ---- stage 4: analyses of the uniform windows (reference :188-224) ---------------------------------------
    def analyze_uniform_data(self, index: int) -> None:
        pkey = "analyze uniform data"
        self.model.load(file_index=index, file_type="uni")
        fn = self.output_dir / self.model.convert_filename_type("uni", "anl").stem
        if dist.is_root():
            print("ANALYSIS: ", fn, flush=True)
        analyses = {"fractal dimension": self.model.fractal_dimension,
                    "structure functions": self.model.structure_functions,
                    "kinetic energy spectra": self.model.kinetic_energy_spectra}
        keys = list(analyses)
        begin_key = self.checkpoint_data.setdefault(pkey, {}).get("analysis")
        begin = keys.index(begin_key) if begin_key in keys else 0
        for akey in keys[begin:]:
            self.checkpoint_data[pkey]["analysis"] = akey
            if (self.settings.get(akey) or {}).get("skip", False):
                continue
            retval = analyses[akey](**(self.settings[akey].get("settings", {}) if akey in self.settings else {}))
            dist.barrier()
            if dist.is_root():
                self.model.save_to_hdf5({akey: retval}, fn)
            dist.barrier()
        self.checkpoint_data[pkey]["analysis"] = None


@timer
def main(workdir=None) -> int:
    dist.init_from_env()
    pipe = Pipeline(workdir)
    pipe.restart()
    if dist.is_root():
        print("\n-------------\n", pipe.checkpoint_data, "\n-------------\n", flush=True)
    with InterruptHandler(external_handler=pipe.checkpoint):
        pkey = "reynolds stress"
        if not (pipe.settings.get(pkey) or {}).get("skip", False):
            begin = pipe.checkpoint_data.get(pkey, {}).get("index", 0)
            for i in sorted(pipe.model.plt_files["by index"].keys())[begin:]:
                pipe.reynolds_stress(index=i)
                pipe.checkpoint_data[pkey] = {"index": i + 1}
        dist.barrier()
        pipe.smooth_window_trajectory()
        dist.barrier()
        pkey = "extract windows"
        if not (pipe.settings.get(pkey) or {}).get("skip", False):
            begin = pipe.checkpoint_data.get(pkey, {}).get("index", 0)
            for i in sorted(pipe.model.plt_files["by index"].keys())[begin:]:
                pipe.extract_windows(index=i)
                pipe.checkpoint_data[pkey] = {"index": i + 1}
        dist.barrier()
        pipe.refresh_model()
        pkey = "analyze uniform data"
        begin = pipe.checkpoint_data.get(pkey, {}).get("index", 0)
        pipe.checkpoint_data.setdefault(pkey, {})
        for i in sorted(pipe.model.uni_files["by index"].keys())[begin:]:
            pipe.analyze_uniform_data(i)
            pipe.checkpoint_data[pkey]["index"] = i + 1
        dist.barrier()
        if dist.is_root():
            print("DONE!")
    return 0


if __name__ == "__main__":
    try:
        sys.exit(main())
    except Exception as exc:  # reference: log, then abort the MPI job (__main__.py:285-290)
        LOGGER.exception("", exc_info=exc)
        sys.exit(1)
