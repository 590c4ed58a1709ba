"""Builds libgpmpc.so (sm_100a) in-tree with nvcc.  Used by __graft_entry__.build() and by hand:

    python gaussian-process-mpc_b200/build.py [--force]

The shared library is git-ignored but travels to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import concurrent.futures as cf
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "build")
LIB = os.path.join(HERE, "libgpmpc.so")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "-Xcompiler", "-O2"] + os.environ.get("GPMPC_EXTRA_NVCC_FLAGS", "").split()
PAIR_DIMS = range(2, 9)


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def _jobs():
    jobs = []
    for name in ("api", "fit", "gemm", "rollout", "fullcov", "split"):
        jobs.append((os.path.join(CSRC, name + ".cu"), os.path.join(OBJ, name + ".o"), []))
    for d in PAIR_DIMS:
        jobs.append((os.path.join(CSRC, "mm_pairs_inst.cu"), os.path.join(OBJ, f"mm_pairs_D{d}.o"),
                     [f"-DGPMPC_INST_D={d}"]))
        jobs.append((os.path.join(CSRC, "mm_full_inst.cu"), os.path.join(OBJ, f"mm_full_D{d}.o"),
                     [f"-DGPMPC_INST_D={d}"]))
    return jobs


def _newest_source():
    t = 0.0
    for root, _, files in os.walk(CSRC):
        for f in files:
            t = max(t, os.path.getmtime(os.path.join(root, f)))
    t = max(t, os.path.getmtime(os.path.join(os.path.dirname(HERE), "include", "gpmpc.h")))
    return t


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= _newest_source():
        return LIB
    os.makedirs(OBJ, exist_ok=True)
    nvcc = _nvcc()
    extra = ["-Xptxas", "-v"] if verbose else []

    def compile_one(job):
        src, obj, defs = job
        cmd = [nvcc] + NVCC_FLAGS + extra + defs + ["-c", src, "-o", obj]
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    with cf.ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 4)) as ex:
        for cmd, r in ex.map(compile_one, _jobs()):
            if verbose and r.stderr:
                sys.stderr.write(r.stderr)
            if r.returncode != 0:
                raise RuntimeError("nvcc failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    objs = [j[1] for j in _jobs()]
    cmd = [nvcc, "-shared", "-o", LIB] + objs + ["-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError("link failed: " + " ".join(cmd) + "\n" + r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
