"""In-tree build of libfava_b200.so (sm_100a only) with nvcc.

`python -m fava_b200.build` or `fava_b200.build.build_library()`.  The shared object is written to
`fava_b200/lib/libfava_b200.so`; it is git-ignored but travels with the tree to the GPU box.
"""

from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG_DIR = Path(__file__).resolve().parent
REPO_ROOT = PKG_DIR.parent
CSRC = PKG_DIR / "csrc"
BUILD_DIR = PKG_DIR / "lib" / "obj"
LIB_PATH = PKG_DIR / "lib" / "libfava_b200.so"
CUDA_HOME = Path(os.environ.get("CUDA_HOME", "/usr/local/cuda"))

NVCC_FLAGS = [
    "-gencode",
    "arch=compute_100a,code=sm_100a",
    "-O3",
    "-lineinfo",
    "-std=c++17",
    "-Xcompiler",
    "-fPIC,-O3,-Wall,-Wno-unused-function",
    "--expt-relaxed-constexpr",
    "-I",
    str(REPO_ROOT / "include"),
    "-I",
    str(CSRC),
]


def _nvcc() -> str:
    cand = CUDA_HOME / "bin" / "nvcc"
    if cand.exists():
        return str(cand)
    found = shutil.which("nvcc")
    if not found:
        raise RuntimeError("nvcc not found; libfava_b200 cannot be built")
    return found


def _sources() -> list[Path]:
    return sorted(CSRC.glob("*.cu"))


def _stale(target: Path, deps: list[Path]) -> bool:
    if not target.exists():
        return True
    t = target.stat().st_mtime
    return any(d.stat().st_mtime > t for d in deps)


def build_library(force: bool = False, verbose: bool = False, extra_flags: list[str] | None = None) -> Path:
    """Compile every .cu under csrc/ for sm_100a and link libfava_b200.so."""
    nvcc = _nvcc()
    BUILD_DIR.mkdir(parents=True, exist_ok=True)
    headers = sorted(CSRC.glob("*.cuh")) + sorted((REPO_ROOT / "include").glob("*.h")) + [Path(__file__)]
    flags = NVCC_FLAGS + (extra_flags or [])

    jobs = []
    for src in _sources():
        obj = BUILD_DIR / (src.stem + ".o")
        if force or _stale(obj, [src] + headers):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc, *flags, "-c", str(src), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd), flush=True)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{res.stdout}\n{res.stderr}")
        if verbose and res.stderr.strip():
            print(res.stderr, file=sys.stderr)
        return obj

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as pool:
            list(pool.map(compile_one, jobs))

    objs = [BUILD_DIR / (s.stem + ".o") for s in _sources()]
    if force or jobs or _stale(LIB_PATH, objs):
        cmd = [
            nvcc,
            "-shared",
            "-gencode",
            "arch=compute_100a,code=sm_100a",
            "-o",
            str(LIB_PATH),
            *map(str, objs),
            "-lcufft",
            "-lpthread",
            "-Xlinker",
            f"-rpath,{CUDA_HOME / 'lib64'}",
        ]
        if verbose:
            print(" ".join(cmd), flush=True)
        res = subprocess.run(cmd, capture_output=True, text=True)
        if res.returncode != 0:
            raise RuntimeError(f"link failed:\n{res.stdout}\n{res.stderr}")
    return LIB_PATH


if __name__ == "__main__":
    p = build_library(force="--force" in sys.argv, verbose=True)
    print(p)
