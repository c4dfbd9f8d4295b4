"""TEST INFRASTRUCTURE ONLY.

Recipe that compiles the UNMODIFIED reference extension from the sources where
they lie under /root/reference into oracle/_ref/ (git-ignored, but it travels
to the GPU box with the gpurun snapshot).  No reference source is copied into
this repository: the only input is /root/reference/SuchTree/MuchTree.c (the
Cython 3.1.4 output of MuchTree.pyx that upstream checks in), compiled with
the flags the reference's own setup.py would get from distutils
(sysconfig CFLAGS: -O2 ...), no cmake / build system involved.

The resulting module still does `from dendropy import Tree` and
`from SuchTree.exceptions import ...` at import time; oracle/ref_loader.py
satisfies both without any reference file (see there).

Usage:  python oracle/build_ref.py [--force]
"""
import os
import shlex
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
REF_C = "/root/reference/SuchTree/MuchTree.c"
OUT_DIR = os.path.join(HERE, "_ref")


def ref_so_path():
    return os.path.join(OUT_DIR, "MuchTree" + sysconfig.get_config_var("EXT_SUFFIX"))


def build(force=False, verbose=True):
    """Returns the .so path, or None when /root/reference is not reachable
    (e.g. on the GPU box, where the prebuilt oracle/_ref/ is used as is)."""
    out = ref_so_path()
    if os.path.exists(out) and not force:
        if not os.path.exists(REF_C) or os.path.getmtime(out) >= os.path.getmtime(REF_C):
            return out
    if not os.path.exists(REF_C):
        return out if os.path.exists(out) else None
    import numpy

    os.makedirs(OUT_DIR, exist_ok=True)
    cc = shlex.split(sysconfig.get_config_var("CC") or "gcc")
    cflags = shlex.split(sysconfig.get_config_var("CFLAGS") or "-O2")
    cmd = (
        cc
        + cflags
        + ["-fPIC", "-shared", "-w", "-fwrapv"]
        + ["-I" + sysconfig.get_paths()["include"], "-I" + numpy.get_include()]
        + [REF_C, "-o", out]
    )
    if verbose:
        print("[oracle/build_ref]", " ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return out


if __name__ == "__main__":
    p = build(force="--force" in sys.argv)
    print(p if p else "reference sources not reachable and no prebuilt oracle/_ref")
