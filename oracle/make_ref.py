"""TEST / BENCH INFRASTRUCTURE ONLY -- stages the UNMODIFIED reference for the GPU box.

    python -m oracle.make_ref          (build container only: needs /root/reference)

The reference is Python, so there is nothing to compile: its own source files are copied, byte for
byte, from where they lie under /root/reference into ``oracle/_ref/`` (same relative layout).  That
directory is git-ignored -- no reference source ever enters the history -- but it is NOT
gpurun-ignored, so it travels to the GPU box like the built ``.so`` files do.  There it lets
``bench.py --impl reference`` and the ``cpu_baseline`` leg time the reference's OWN classification
loop (src/vilgod/zero_shot_detector.py:389-415 through src/utils/mv_utils.py, src/utils/clip_utils.py
and third_party/CLIP/clip/*) on the box's host cores: ``cpu_baseline.kind = "reference"``.
``oracle/ref_harness.py`` finds the copy when /root/reference itself is absent.

Only the files the hot path imports are staged: the ``src`` package (212 KB of .py files) and the
``clip`` package including its BPE vocabulary.  Nothing in vilgod_b200/ reads this directory.
"""
from __future__ import annotations

import filecmp
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SOURCE_ROOT = "/root/reference"
DEST = os.path.join(HERE, "_ref")
TREES = [("src", (".py",)), (os.path.join("third_party", "CLIP", "clip"), (".py", ".gz"))]


def stage(verbose=False):
    """-> number of files staged (0 when the reference tree is not mounted)."""
    if not os.path.isfile(os.path.join(SOURCE_ROOT, "src", "utils", "mv_utils.py")):
        return 0
    n = 0
    for tree, exts in TREES:
        for dirpath, _, files in os.walk(os.path.join(SOURCE_ROOT, tree)):
            rel = os.path.relpath(dirpath, SOURCE_ROOT)
            for f in files:
                if not f.endswith(exts):
                    continue
                src, dst = os.path.join(dirpath, f), os.path.join(DEST, rel, f)
                os.makedirs(os.path.dirname(dst), exist_ok=True)
                if not (os.path.exists(dst) and filecmp.cmp(src, dst, shallow=False)):
                    shutil.copyfile(src, dst)
                n += 1
                if verbose:
                    print("  staged", os.path.join(rel, f))
    with open(os.path.join(DEST, "PROVENANCE.txt"), "w") as fh:
        fh.write("Unmodified copies of files from chreisinger/ViLGOD (/root/reference), staged by\n"
                 "oracle/make_ref.py for timing the reference on the GPU box.  Not part of the repo.\n")
    return n


if __name__ == "__main__":
    k = stage(verbose="-v" in sys.argv)
    print(f"staged {k} reference files into {DEST}" if k else "reference tree not mounted: nothing staged")
