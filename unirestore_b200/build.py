"""In-tree build of the C-ABI shared library (nvcc, sm_100a only).

``python -m unirestore_b200.build`` compiles every ``csrc/*.cu`` with
``-gencode arch=compute_100a,code=sm_100a -lineinfo`` into
``unirestore_b200/libunirestore_b200.so``.  The .so is git-ignored but travels with the
repo snapshot to the GPU box.  Object files are cached under ``build/`` keyed by mtime.
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libunirestore_b200.so")
OBJ_DIR = os.path.join(ROOT, "build", "obj")

NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "--compiler-options", "-fPIC", "-diag-suppress", "177"]


def _newer(path, than):
    return os.path.exists(path) and os.path.getmtime(path) >= than


def build(force: bool = False, verbose: bool = False) -> str:
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    srcs = sorted(f for f in os.listdir(CSRC) if f.endswith(".cu"))
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh"))]
    hdrs.append(os.path.join(ROOT, "include", "unirestore_b200.h"))
    hdr_m = max(os.path.getmtime(h) for h in hdrs)
    os.makedirs(OBJ_DIR, exist_ok=True)
    jobs, objs = [], []
    for s in srcs:
        src = os.path.join(CSRC, s)
        obj = os.path.join(OBJ_DIR, s[:-3] + ".o")
        objs.append(obj)
        if force or not _newer(obj, max(os.path.getmtime(src), hdr_m)):
            jobs.append([nvcc, *NVCC_FLAGS, "-c", src, "-o", obj] + (["-Xptxas", "-v"] if verbose else []))

    def run(cmd):
        r = subprocess.run(cmd, capture_output=True, text=True)
        return cmd, r

    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as ex:
        for cmd, r in ex.map(run, jobs):
            if verbose or r.returncode:
                sys.stderr.write(" ".join(cmd) + "\n" + r.stdout + r.stderr)
            if r.returncode:
                raise RuntimeError("nvcc failed for %s" % cmd[-3])
    if jobs or not os.path.exists(LIB):
        cmd = [nvcc, "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
