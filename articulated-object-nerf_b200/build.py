"""Builds libaon_b200.so (the C-ABI library of include/aon.h) in-tree with nvcc for sm_100a.

    python articulated-object-nerf_b200/build.py [--force] [--verbose]

nvcc cross-compiles without a GPU.  The .so is git-ignored but travels to the GPU box with the
gpurun snapshot.  A content hash of the sources is stored next to it so rebuilds are skipped when
nothing changed.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "libaon_b200.so")
SOURCES = ["aon_api.cu", "render_simt.cu", "render_tc.cu", "train_ops.cu", "gemm_tc.cu"]
CFLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
          "-Xcompiler", "-fPIC", "-Xptxas", "-v"]
FLAGS = CFLAGS + ["-shared"]


def _digest() -> str:
    h = hashlib.sha256()
    names = sorted(os.listdir(CSRC)) + ["../../include/aon.h"]
    for n in names:
        p = os.path.join(CSRC, n)
        if os.path.isfile(p):
            h.update(n.encode())
            with open(p, "rb") as f:
                h.update(f.read())
    h.update(" ".join(FLAGS).encode())
    return h.hexdigest()


def _compile_one(args):
    nvcc, src, obj = args
    cmd = [nvcc] + CFLAGS + ["-c", "-o", obj, src]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    return src, " ".join(cmd), proc.returncode, proc.stdout


def build(force: bool = False, verbose: bool = False) -> str:
    """One nvcc -c per source file, in parallel (objects cached under csrc/_obj, keyed by a hash of the file, the headers
    and the flags), then one link into libaon_b200.so."""
    from concurrent.futures import ThreadPoolExecutor
    stamp = LIB + ".sha256"
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(stamp) and open(stamp).read().strip() == dig:
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    objdir = os.path.join(CSRC, "_obj")
    os.makedirs(objdir, exist_ok=True)
    hdr = hashlib.sha256()
    for n in sorted(os.listdir(CSRC)) + ["../../include/aon.h"]:
        p = os.path.join(CSRC, n)
        if os.path.isfile(p) and not n.endswith(".cu"):
            hdr.update(n.encode() + open(p, "rb").read())
    hdr.update(" ".join(CFLAGS).encode())
    jobs, objs = [], []
    for s in SOURCES:
        src = os.path.join(CSRC, s)
        h = hashlib.sha256(hdr.digest() + open(src, "rb").read()).hexdigest()[:16]
        obj = os.path.join(objdir, "%s.%s.o" % (s[:-3], h))
        objs.append(obj)
        if force or not os.path.exists(obj):
            for old in os.listdir(objdir):
                if old.startswith(s[:-3] + "."):
                    os.remove(os.path.join(objdir, old))
            jobs.append((nvcc, src, obj))
    log = []
    with ThreadPoolExecutor(max_workers=max(1, min(len(jobs), os.cpu_count() or 1))) as ex:
        results = list(ex.map(_compile_one, jobs))
    failed = False
    for src, cmd, rc, out in results:
        log.append(cmd + "\n" + out)
        failed = failed or rc != 0
    if not failed:
        cmd = [nvcc, "-shared", "-o", LIB] + objs
        proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
        log.append(" ".join(cmd) + "\n" + proc.stdout)
        failed = proc.returncode != 0
    text = "\n".join(log)
    with open(os.path.join(HERE, "build.log"), "a" if jobs and not force else "w") as f:
        f.write(text)
    if verbose or failed:
        sys.stderr.write(text)
    if failed:
        raise RuntimeError("nvcc failed building libaon_b200.so (see build.log)")
    with open(stamp, "w") as f:
        f.write(dig)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
