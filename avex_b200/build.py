"""Build libavexk.so in-tree with nvcc for sm_100a (no torch dependency: plain C ABI + cudart).

    python -m avex_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libavexk.so")
ARCH = ["-gencode", "arch=compute_100a,code=sm_100a"]
FLAGS = ["-O3", "-std=c++17", "-lineinfo", "-Xcompiler", "-fPIC", "--use_fast_math=false"]


def _nvcc() -> str:
    for c in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if c and os.path.exists(c):
            return c
    raise RuntimeError("nvcc not found; libavexk.so cannot be built")


def sources() -> list[str]:
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def tracked_files() -> list[str]:
    """Every file the binary is built from: csrc/*.cu, csrc/*.cuh, include/avexk.h."""
    files = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh", ".h"))]
    files.append(os.path.join(os.path.dirname(HERE), "include", "avexk.h"))
    return sorted(files)


def source_id() -> str:
    """sha256 over the tracked sources (name + contents).  Compiled into libavexk.so (`avexk_build_id()`), so that a stale
    binary -- the .so is git-ignored and ships with the gpurun snapshot -- can never be what the tests load."""
    h = hashlib.sha256()
    for f in tracked_files():
        h.update(os.path.basename(f).encode() + b"\0")
        with open(f, "rb") as fh:
            h.update(fh.read())
        h.update(b"\0")
    return h.hexdigest()[:32]


_MARK = b"AVEXK_BUILD_ID="


def binary_id(path: str = LIB) -> str | None:
    """The id compiled into an existing libavexk.so, read from its bytes (no dlopen), or None."""
    if not os.path.exists(path):
        return None
    with open(path, "rb") as fh:
        blob = fh.read()
    i = blob.find(_MARK)
    if i < 0:
        return None
    j = i + len(_MARK)
    return blob[j : j + 32].decode("ascii", "replace")


def _stale(target: str, deps: list[str]) -> bool:
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(os.path.dirname(HERE), "include", "avexk.h"))
    flags = [f for f in FLAGS if f != "--use_fast_math=false"]
    if verbose:
        flags += ["-Xptxas", "-v"]
    sid = source_id()
    if binary_id() != sid:
        force = True  # sources differ from what the shipped binary was built from (mtimes cannot be trusted after a snapshot)
    jobs = []
    objs = []
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or _stale(obj, [src] + headers):
            extra = [f'-DAVEXK_BUILD_ID_STR="{sid}"'] if os.path.basename(src) == "api.cu" else []
            jobs.append([nvcc, *ARCH, *flags, *extra, "-c", src, "-o", obj])

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("nvcc failed:\n" + " ".join(cmd) + "\n" + r.stdout + r.stderr)
        return r.stdout + r.stderr

    if jobs:
        with ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for log in ex.map(run, jobs):
                if verbose and log.strip():
                    print(log)
    if jobs or force or _stale(LIB, objs):
        run([nvcc, *ARCH, "-shared", "-o", LIB, *objs, "-cudart", "static"])
    got = binary_id()
    if got != sid:
        raise RuntimeError(f"libavexk.so carries build id {got}, sources hash to {sid}")
    return LIB


if __name__ == "__main__":
    path = build(force="--force" in sys.argv, verbose="--verbose" in sys.argv)
    print(path)
