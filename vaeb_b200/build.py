"""Builds vaeb_b200/libvaeb_b200.so (the C-ABI library of include/vaeb_b200.h) in-tree with nvcc
for sm_100a.  `python -m vaeb_b200.build [--force]`."""
from __future__ import annotations

import concurrent.futures as cf
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "csrc", "_obj")
LIB = os.path.join(HERE, "libvaeb_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
         "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr", "-Xptxas", "-v"]


def _sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith(".cu"))


def _header_hash():
    h = hashlib.sha256()
    for root in (CSRC, os.path.join(HERE, "..", "include")):
        for f in sorted(os.listdir(root)):
            p = os.path.join(root, f)
            if os.path.isfile(p) and f.endswith((".cuh", ".h")):
                h.update(f.encode())
                h.update(open(p, "rb").read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _src_stamp(src, hh):
    return hashlib.sha256(open(src, "rb").read() + hh.encode()).hexdigest()


def _compile(job):
    src, hh, force = job
    obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
    stamp = _src_stamp(src, hh)
    if not force and os.path.exists(obj) and os.path.exists(obj + ".stamp") and open(obj + ".stamp").read() == stamp:
        return obj, False                      # this translation unit and every header are unchanged
    r = subprocess.run([NVCC, *FLAGS, "-c", src, "-o", obj], capture_output=True, text=True)
    log = r.stdout + r.stderr
    with open(obj + ".log", "w") as f:
        f.write(log)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed for %s:\n%s" % (src, log))
    with open(obj + ".stamp", "w") as f:
        f.write(stamp)
    return obj, True


def build(force=False, verbose=False):
    os.makedirs(OBJ, exist_ok=True)
    srcs = _sources()
    hh = _header_hash()
    with cf.ThreadPoolExecutor(max_workers=min(8, len(srcs))) as ex:
        res = list(ex.map(_compile, [(s, hh, force) for s in srcs]))
    objs = [o for o, _ in res]
    if any(c for _, c in res) or not os.path.exists(LIB):
        r = subprocess.run([NVCC, "-shared", "-o", LIB, *objs, "-ldl"], capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError("link failed:\n" + r.stdout + r.stderr)
    if verbose:
        for o in objs:
            sys.stdout.write(open(o + ".log").read())
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
