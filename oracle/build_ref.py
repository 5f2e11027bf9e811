#!/usr/bin/env python
"""
Recipe for oracle/_ref: the UNMODIFIED reference package, for use as the CPU arm / checker.

    python oracle/build_ref.py          # needs /root/reference (the build container); writes oracle/_ref/pygmu2/

rdpoor/pygmu2 is pure Python (SURVEY.md headline fact 1): there is nothing to compile, so "building" the
reference is copying its package directory, as it lies under /root/reference/src/pygmu2, into the git-ignored
oracle/_ref/ (never into history; it travels to the GPU box with the snapshot like a built .so would).  Its
imports of soundfile / sounddevice / mido / rtmidi / numba -- none of which exist in this image -- are met by
the stub modules in oracle/stubs/ (empty modules; numba.jit as identity; soundfile.read over stdlib wave), which
the hot path (numpy.fft in ConvolvePE, scipy.signal in SpatialHRTF) never touches.

Test infrastructure only: nothing under pygmu2_b200/ imports it.  Users: ``bench.py --impl reference`` and
``bench.py``'s ``cpu_baseline`` leg (kind "reference"), tests/test_oracle_ref.py (the oracle port against the
real reference wherever _ref is present).
"""
from __future__ import annotations

import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF_SRC = os.path.join(os.environ.get("PYGMU2_REFERENCE", "/root/reference"), "src", "pygmu2")
DST = os.path.join(HERE, "_ref")


def build_ref(force: bool = False) -> str | None:
    """Copy the reference package into oracle/_ref/pygmu2.  Returns the path, or None when the reference tree
    is not present (the GPU box: the prebuilt copy that travelled with the snapshot is used as is)."""
    dst_pkg = os.path.join(DST, "pygmu2")
    if not os.path.isdir(REF_SRC):
        return dst_pkg if os.path.isdir(dst_pkg) else None
    if os.path.isdir(dst_pkg) and not force:
        src_m = max(os.path.getmtime(os.path.join(r, f)) for r, _, fs in os.walk(REF_SRC) for f in fs)
        if os.path.getmtime(dst_pkg) >= src_m:
            return dst_pkg
    if os.path.isdir(dst_pkg):
        shutil.rmtree(dst_pkg)
    os.makedirs(DST, exist_ok=True)
    shutil.copytree(REF_SRC, dst_pkg, ignore=shutil.ignore_patterns("__pycache__", "*.pyc"))
    os.utime(dst_pkg, None)
    return dst_pkg


def import_ref():
    """-> the reference package (module ``pygmu2``) from oracle/_ref, or None when it is not there."""
    if not os.path.isdir(os.path.join(DST, "pygmu2")):
        return None
    for p in (DST, os.path.join(HERE, "stubs")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import pygmu2
    return pygmu2


if __name__ == "__main__":
    print(build_ref(force="--force" in sys.argv))
