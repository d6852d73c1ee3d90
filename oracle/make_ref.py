"""Recipe for ``oracle/_ref/``: a copy of the reference's OWN Python sources of the hot path, so that the CPU arm of
``bench.py`` can time the reference itself (not the port) on the GPU box, where /root/reference does not exist.

    python -m oracle.make_ref            # build container only; also run by __graft_entry__.build()

Copies ``models/**/*.py``, ``datasets/*.py``, ``utils/*.py`` and ``opt.py`` from /root/reference into ``oracle/_ref/``
unchanged (byte for byte; a MANIFEST with sha256 per file is written next to them).  ``oracle/_ref/`` is listed in
.gitignore -- reference sources never enter this repository's history -- but not in .gpurunignore, so the copy
travels to the GPU box with the snapshot exactly like the built libaon_b200.so.  The files are imported through
``oracle/ref_import.py`` (stubs for the six absent third-party modules).
"""
from __future__ import annotations

import glob
import hashlib
import os
import shutil
import sys

SRC = "/root/reference"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")
PATTERNS = ("models/*.py", "models/vanilla_nerf/*.py", "datasets/*.py", "utils/*.py", "opt.py")


def make(verbose: bool = True) -> bool:
    if not os.path.isdir(SRC):
        if verbose:
            print("oracle/make_ref: %s absent (not the build container); keeping the existing oracle/_ref" % SRC)
        return os.path.isdir(DST)
    if os.path.isdir(DST):
        shutil.rmtree(DST)
    lines = []
    for pat in PATTERNS:
        for src in sorted(glob.glob(os.path.join(SRC, pat))):
            rel = os.path.relpath(src, SRC)
            dst = os.path.join(DST, rel)
            os.makedirs(os.path.dirname(dst), exist_ok=True)
            shutil.copyfile(src, dst)
            lines.append("%s  %s" % (hashlib.sha256(open(dst, "rb").read()).hexdigest(), rel))
    with open(os.path.join(DST, "MANIFEST.sha256"), "w") as f:
        f.write("# copied unchanged from %s by oracle/make_ref.py\n" % SRC + "\n".join(lines) + "\n")
    if verbose:
        print("oracle/make_ref: %d files -> %s" % (len(lines), DST))
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
