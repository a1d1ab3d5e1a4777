"""Build libgm3d_sm100.so in-tree with nvcc (sm_100a only; cross-compiles without a GPU).

    python -m gm3d_b200.build [--force] [--verbose]

One `nvcc -c` per .cu (parallel), then one shared link.  No torch, no pybind: the library exports the
plain C ABI of include/gm3d.h and is loaded with ctypes (gm3d_b200/_lib.py).
"""
from __future__ import annotations

import argparse
import concurrent.futures
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
SO = os.path.join(PKG, "libgm3d_sm100.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC,-fvisibility=hidden",
    "-Xptxas", "-v",
    "-I", os.path.join(ROOT, "include"), "-I", CSRC,
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC=/path/to/nvcc)")


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _stale(out: str, deps) -> bool:
    if not os.path.exists(out):
        return True
    t = os.path.getmtime(out)
    return any(os.path.getmtime(d) > t for d in deps)


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = _nvcc()
    extra = os.environ.get("GM3D_NVCC_FLAGS", "").split()  # e.g. -DGM3D_CS_DEBUG (tuning aids of the fused kernel)
    flags = NVCC_FLAGS + extra
    force = force or bool(extra)
    os.makedirs(OBJ, exist_ok=True)
    headers = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    headers.append(os.path.join(ROOT, "include", "gm3d.h"))
    headers.append(os.path.abspath(__file__))
    jobs = []
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        if force or _stale(obj, [src] + headers):
            jobs.append((src, obj))

    def compile_one(job):
        src, obj = job
        cmd = [nvcc, *flags, "-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    with concurrent.futures.ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for src, r in ex.map(compile_one, jobs):
            if r.returncode != 0:
                sys.stderr.write(r.stdout + r.stderr)
                raise RuntimeError(f"nvcc failed on {src}")
            with open(os.path.join(OBJ, os.path.basename(src)[:-3] + ".ptxas.log"), "w") as f:
                f.write(r.stderr)
            if verbose:
                sys.stderr.write(r.stderr)

    objs = [os.path.join(OBJ, os.path.basename(s)[:-3] + ".o") for s in sources()]
    if force or jobs or _stale(SO, objs):
        cmd = [nvcc, "-shared", "-o", SO, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return SO


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--force", action="store_true")
    ap.add_argument("--verbose", action="store_true")
    a = ap.parse_args()
    print(build(force=a.force, verbose=a.verbose))
