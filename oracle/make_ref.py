#!/usr/bin/env python
"""Recipe for oracle/_ref/: the reference's own three hot-path files, so that the GPU box
(where /root/reference does not exist) can time the REAL numba path beside the GPU.

    python oracle/make_ref.py            # needs /root/reference (build container)

Copies, UNMODIFIED,
    picturedrocks/markers/mutualinformation/infoset.py
    picturedrocks/markers/mutualinformation/iterative.py
    picturedrocks/markers/_iterative.py
from the reference tree into oracle/_ref/picturedrocks/... (git-ignored, not gpurun-ignored: it
travels with the snapshot like the built .so files; the sources never enter the history).
oracle/ref_shim.py loads them from there exactly as it loads them from /root/reference: stub
parent packages, and the one `yshift` assignment hoisted in memory at load time so that numba
0.65 can compile `_sparse_entropy_wrt` (no arithmetic change).
TEST / BASELINE INFRASTRUCTURE: used only by bench.py's CPU legs and tests/.
"""
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC_ROOT = os.environ.get("PRB_REFERENCE_ROOT", "/root/reference")
FILES = ["picturedrocks/markers/mutualinformation/infoset.py",
         "picturedrocks/markers/mutualinformation/iterative.py",
         "picturedrocks/markers/_iterative.py"]


def make(verbose=True):
    if not os.path.isfile(os.path.join(SRC_ROOT, FILES[0])):
        if verbose:
            print("reference tree not present at %s: oracle/_ref left as it is" % SRC_ROOT)
        return False
    dst_root = os.path.join(HERE, "_ref")
    lines = []
    for rel in FILES:
        dst = os.path.join(dst_root, rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(os.path.join(SRC_ROOT, rel), dst)
        lines.append("%s  %s" % (hashlib.sha256(open(dst, "rb").read()).hexdigest(), rel))
    with open(os.path.join(dst_root, "SHA256SUMS"), "w") as f:
        f.write("\n".join(lines) + "\n")
    if verbose:
        print("\n".join(lines))
    return True


if __name__ == "__main__":
    sys.exit(0 if make() else 1)
