"""In-tree nvcc build of libsuchtree_b200.so (sm_100a only).

    python -m suchtree_b200.build [--force] [--verbose]

The .so is git-ignored but travels to the GPU box with the gpurun snapshot.
"""
import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libsuchtree_b200.so")
# the C header: <repo>/include in a checkout, <package>/include when pip-installed
INCLUDE = os.path.join(os.path.dirname(HERE), "include")
if not os.path.exists(os.path.join(INCLUDE, "suchtree_b200.h")):
    INCLUDE = os.path.join(HERE, "include")
OBJ = os.path.join(HERE, "_obj")

NVCC_FLAGS = [
    "-O3",
    "-std=c++17",
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo",
    "-Xcompiler", "-fPIC",
    "-Xcompiler", "-fvisibility=hidden",
    "--expt-relaxed-constexpr",
    "-shared",
    "-cudart", "static",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(INCLUDE, "*.h"))
    return any(os.path.getmtime(d) > t for d in deps)


def _compile_one(args):
    src, obj, flags, verbose = args
    cmd = [_nvcc()] + flags + ["-I", INCLUDE, "-c", "-o", obj, src]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return obj


def build(force=False, verbose=False, extra=()):
    """One object per translation unit (compiled in parallel, only the stale ones),
    then one device-link-free shared library."""
    if not force and not needs_build():
        return LIB
    from concurrent.futures import ThreadPoolExecutor

    os.makedirs(OBJ, exist_ok=True)
    flags = [f for f in NVCC_FLAGS if f != "-shared"] + list(extra)
    if verbose:
        flags = ["-Xptxas", "-v"] + flags
    hdr_t = max([os.path.getmtime(h) for h in glob.glob(os.path.join(CSRC, "*.cuh"))
                 + glob.glob(os.path.join(INCLUDE, "*.h"))] + [os.path.getmtime(__file__)])
    jobs, objs = [], []
    for src in sources():
        obj = os.path.join(OBJ, os.path.basename(src)[:-3] + ".o")
        objs.append(obj)
        if force or extra or not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t):
            jobs.append((src, obj, flags, verbose))
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as pool:
        list(pool.map(_compile_one, jobs))
    cmd = [_nvcc(), "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC", "-o", LIB] + objs
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
