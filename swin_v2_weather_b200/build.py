"""Builds libswinb200.so (the C-ABI kernel library) in-tree with nvcc for sm_100a.

`python -m swin_v2_weather_b200.build` or `__graft_entry__.build()`.  Objects are cached by a hash of
the source + headers + flags, so rebuilding after a one-file change recompiles one file.
"""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import shutil
import subprocess
import sys

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
INCLUDE = os.path.join(os.path.dirname(PKG), "include")
OBJ = os.path.join(CSRC, "build")
LIB = os.path.join(PKG, "libswinb200.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _digest(path: str, headers) -> str:
    h = hashlib.sha256()
    h.update(" ".join(NVCC_FLAGS).encode())
    for p in [path] + list(headers):
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()[:16]


def build(verbose: bool = False, force: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    sources = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    headers = sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cuh"))
    headers.append(os.path.join(INCLUDE, "swinb200.h"))
    headers.append(os.path.join(INCLUDE, "swinb200_debug.h"))
    jobs = []
    objs = []
    for src in sources:
        path = os.path.join(CSRC, src)
        obj = os.path.join(OBJ, f"{src[:-3]}.{_digest(path, headers)}.o")
        objs.append(obj)
        if force or not os.path.exists(obj):
            for old in os.listdir(OBJ):
                if old.startswith(src[:-3] + ".") and old.endswith(".o"):
                    os.remove(os.path.join(OBJ, old))
            jobs.append((src, [nvcc, *NVCC_FLAGS, "-I", INCLUDE, "-c", path, "-o", obj]))

    def run(job):
        src, cmd = job
        r = subprocess.run(cmd, capture_output=True, text=True)
        return src, r

    if jobs:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(jobs))) as ex:
            for src, r in ex.map(run, jobs):
                log = os.path.join(OBJ, src[:-3] + ".ptxas.log")
                with open(log, "w") as f:
                    f.write(r.stderr)
                if r.returncode != 0:
                    sys.stderr.write(r.stdout + r.stderr)
                    raise RuntimeError(f"nvcc failed on {src}")
                if verbose:
                    sys.stderr.write(r.stderr)
    if jobs or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB, *objs, "-lcudart_static", "-lpthread", "-ldl", "-lrt"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(verbose="-v" in sys.argv, force="-f" in sys.argv))
