"""Builds libpoyb200.so (the CUDA kernels + C ABI) in-tree with nvcc for sm_100a.

One translation unit per kernel family (csrc/k_*.cu) plus the host side (csrc/api.cu and the other csrc/*.cu), compiled
in parallel into build/obj/ and linked into poyd_b200/libpoyb200.so."""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(HERE, "libpoyb200.so")
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "..", "build", "obj")
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _deps():
    deps = [os.path.join(CSRC, f) for f in os.listdir(CSRC)]
    deps.append(os.path.join(HERE, "..", "include", "poyb200.h"))
    return deps


def _stale() -> bool:
    if not os.path.exists(SO):
        return True
    t = os.path.getmtime(SO)
    return any(os.path.getmtime(d) > t for d in _deps())


def build(force: bool = False, verbose: bool = False, defines=(), out: str | None = None) -> str:
    """Compile every CUDA source of the package; returns the path of the shared library.  `defines` / `out` build an
    experimental variant next to the product library (tools/build_variant.sh)."""
    target = out or SO
    if not (force or out or _stale()):
        return target
    nvcc = os.environ.get("NVCC", "nvcc")
    objdir = OBJ if not out else os.path.join(OBJ, os.path.basename(out) + ".d")
    os.makedirs(objdir, exist_ok=True)
    newest_header = max(os.path.getmtime(d) for d in _deps() if not d.endswith(".cu"))
    jobs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + ".o")
        if not force and not out and os.path.exists(obj) and os.path.getmtime(obj) > max(os.path.getmtime(src), newest_header):
            jobs.append((None, obj))
            continue
        cmd = [nvcc] + NVCC_FLAGS + list(defines) + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", obj, src]
        jobs.append((cmd, obj))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        for rc, (cmd, _) in zip(ex.map(lambda j: subprocess.call(j[0]) if j[0] else 0, jobs), jobs):
            if rc:
                raise subprocess.CalledProcessError(rc, cmd)
    subprocess.check_call([nvcc, "-shared", "-Wno-deprecated-gpu-targets", "-o", target] + [o for _, o in jobs] + ["-lcudart"])
    return target


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
