"""In-tree nvcc build (sm_100a only) of

    libsuchtree_b200.so        the product: csrc/*.cu behind include/suchtree_b200.h
    libsuchtree_b200_bench.so  bench-only tooling: bench/*.cu behind include/suchtree_b200_bench.h
    libsuchtree_b200_py.so     CPython glue of the by-name entry points: pyglue/st_pynames.c (gcc)

    python -m suchtree_b200.build [--force] [--verbose]

The .so files are git-ignored but travel to the GPU box with the gpurun snapshot.
Each binary carries the hash of the sources it was compiled from (st_build_id());
the loader (suchtree_b200/_lib.py) rebuilds a library whose id differs from the
sources on disk, so a stale binary can never be measured by accident.
"""
import glob
import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
BENCH_SRC = os.path.join(HERE, "bench")
LIB = os.path.join(HERE, "libsuchtree_b200.so")
BENCH_LIB = os.path.join(HERE, "libsuchtree_b200_bench.so")
# CPython glue (name -> id walks of the by-name entry points): plain C against Python.h, loaded
# with ctypes.PyDLL; optional -- without it the shim runs the same walk in Python
PY_SRC = os.path.join(HERE, "pyglue", "st_pynames.c")
PY_LIB = os.path.join(HERE, "libsuchtree_b200_py.so")
# the C headers: <repo>/include in a checkout, <package>/include when pip-installed
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
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found")


def sources(which="product"):
    return sorted(glob.glob(os.path.join(CSRC if which == "product" else BENCH_SRC, "*.cu")))


def _deps(which):
    return (sources(which) + sorted(glob.glob(os.path.join(CSRC, "*.cuh")))
            + sorted(glob.glob(os.path.join(INCLUDE, "*.h"))))


def source_id(which="product"):
    """sha1 over names + contents of everything the library is compiled from."""
    h = hashlib.sha1()
    for path in _deps(which):
        h.update(os.path.basename(path).encode())
        with open(path, "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def pairs_kernel_id():
    """sha1 over the sources the pair kernel (k_pairs, the kernel bench.py's roofline is about)
    is compiled from: profiles/traffic.json is stamped with it, so the ncu DRAM-traffic figure
    is only reported for the kernel it was captured from."""
    h = hashlib.sha1()
    for name in ("st_query.cu", "st_device.cuh", "st_internal.cuh"):
        with open(os.path.join(CSRC, name), "rb") as f:
            h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()[:16]


def _id_file(lib):
    return lib + ".id"


def built_id(lib):
    try:
        with open(_id_file(lib)) as f:
            return f.read().strip()
    except OSError:
        return None


def needs_build(which="product"):
    lib = LIB if which == "product" else BENCH_LIB
    if not os.path.exists(lib):
        return True
    if not sources(which):  # installed without sources: nothing to compare with
        return False
    return built_id(lib) != source_id(which)


def _compile_one(args):
    src, obj, flags, verbose = args
    cmd = [_nvcc()] + flags + ["-I", INCLUDE, "-c", "-o", obj, src]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    return obj


def _build_one(which, force, verbose, extra):
    from concurrent.futures import ThreadPoolExecutor

    lib = LIB if which == "product" else BENCH_LIB
    sid = source_id(which)
    os.makedirs(OBJ, exist_ok=True)
    flags = list(NVCC_FLAGS) + list(extra)
    if verbose:
        flags = ["-Xptxas", "-v"] + flags
    hdr_t = max([os.path.getmtime(h) for h in glob.glob(os.path.join(CSRC, "*.cuh"))
                 + glob.glob(os.path.join(INCLUDE, "*.h"))] + [os.path.getmtime(__file__)])
    jobs, objs = [], []
    for src in sources(which):
        base = os.path.basename(src)[:-3]
        obj = os.path.join(OBJ, ("" if which == "product" else "bench_") + base + ".o")
        objs.append(obj)
        f = flags
        carries_id = base in ("st_index", "st_bench_gather")
        if carries_id:  # the TU that answers st_build_id(): recompiled whenever anything changed
            f = flags + ['-DST_BUILD_ID="%s"' % sid]
        if base == "st_query":
            f = flags + ['-DST_PAIRS_KERNEL_ID="%s"' % pairs_kernel_id()]
        stale = not os.path.exists(obj) or os.path.getmtime(obj) < max(os.path.getmtime(src), hdr_t)
        if force or extra or stale or (carries_id and built_id(lib) != sid):
            jobs.append((src, obj, f, verbose))
    with ThreadPoolExecutor(max_workers=min(8, max(1, len(jobs)))) as pool:
        list(pool.map(_compile_one, jobs))
    cmd = [_nvcc(), "-shared", "-cudart", "static", "-gencode", "arch=compute_100a,code=sm_100a",
           "-Xcompiler", "-fPIC", "-o", lib] + objs + ["-ldl"]
    if which != "product":
        cmd += ["-lcuda"]  # cuTensorMapEncodeTiled of the gather4 experiment
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    with open(_id_file(lib), "w") as f:
        f.write(sid + "\n")
    return lib


def py_glue_id():
    h = hashlib.sha1()
    with open(PY_SRC, "rb") as f:
        h.update(f.read())
    h.update(sys.version.encode())
    return h.hexdigest()[:16]


def py_glue_needs_build():
    if not os.path.exists(PY_SRC):
        return False
    return not os.path.exists(PY_LIB) or built_id(PY_LIB) != py_glue_id()


def build_py_glue(verbose=False):
    """gcc -shared against this interpreter's Python.h; returns the path, or None when there is
    no compiler / no header (the shim then keeps its Python walk)."""
    import sysconfig

    inc = sysconfig.get_paths().get("include") or ""
    cc = shutil.which("gcc") or shutil.which("cc")
    if not cc or not os.path.exists(os.path.join(inc, "Python.h")):
        return None
    cmd = [cc, "-O2", "-shared", "-fPIC", "-fvisibility=hidden", "-I", inc, PY_SRC, "-o", PY_LIB]
    if verbose:
        print(" ".join(cmd), flush=True)
    subprocess.check_call(cmd)
    with open(_id_file(PY_LIB), "w") as f:
        f.write(py_glue_id() + "\n")
    return PY_LIB


def build(force=False, verbose=False, extra=(), which=("product", "bench", "py")):
    """One object per translation unit (compiled in parallel, only the stale ones),
    then one device-link-free shared library per target."""
    for w in which:
        if w == "py":
            if force or py_glue_needs_build():
                build_py_glue(verbose)
        elif force or extra or needs_build(w):
            _build_one(w, force, verbose, extra)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
