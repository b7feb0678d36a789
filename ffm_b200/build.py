"""Builds libffm_b200.so (hand-written sm_100a kernels + C ABI) in-tree with nvcc.

``python -m ffm_b200.build`` or ``ffm_b200.build.build()``; the driver calls it through
``__graft_entry__.build()``.  nvcc cross-compiles without a GPU.

Every ``csrc/*.cu`` is one translation unit (one kernel family each), compiled to an object in
parallel and linked into the shared library.  An object is rebuilt when its source, ANY header under
``csrc/`` or ``include/``, or this file is newer than it -- so an edit to any kernel header rebuilds
what depends on it and a stale prebuilt library can never be loaded silently.
"""
import glob
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.environ.get("FFM_B200_LIB") or os.path.join(HERE, "libffm_b200.so")   # FFM_B200_LIB: load a prebuilt experiment build
CSRC = os.path.join(HERE, "csrc")
OBJDIR = os.environ.get("FFM_B200_OBJDIR") or os.path.join(HERE, "build")          # git-ignored
EXTRA_FLAGS = os.environ.get("FFM_B200_NVCC_FLAGS", "").split()     # experiment builds (e.g. -DFFM_PHASE_TIMING), with FFM_B200_LIB / FFM_B200_OBJDIR
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17",
    "-fmad=false",              # never contract a*b+c: the reference rounds the product first
    "-Xcompiler", "-fPIC",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def headers():
    return sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) +
                  glob.glob(os.path.join(CSRC, "*.inl")) +
                  glob.glob(os.path.join(ROOT, "include", "*.h"))) + [os.path.abspath(__file__)]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return "nvcc"


def _obj(src):
    return os.path.join(OBJDIR, os.path.splitext(os.path.basename(src))[0] + ".o")


def _newest(paths):
    return max(os.path.getmtime(p) for p in paths)


def stale_objects():
    hdr_t = _newest(headers())
    return [s for s in sources()
            if not os.path.exists(_obj(s)) or os.path.getmtime(_obj(s)) < max(hdr_t, os.path.getmtime(s))]


def is_stale():
    if not os.path.exists(LIB):
        return True
    if not glob.glob(os.path.join(OBJDIR, "*.o")):
        # a prebuilt library shipped without its objects (a gpurun snapshot carries the .so, not the build directory):
        # stale only if a source or header is newer than the library
        return os.path.getmtime(LIB) < max(_newest(headers()), _newest(sources()))
    if stale_objects():
        return True
    objs = [_obj(s) for s in sources()]
    stray = set(glob.glob(os.path.join(OBJDIR, "*.o"))) - set(objs)     # a deleted source leaves its object behind
    return bool(stray) or os.path.getmtime(LIB) < _newest(objs)


def build(force=False, verbose=False, jobs=None):
    if not force and not is_stale():
        return LIB
    os.makedirs(OBJDIR, exist_ok=True)
    todo = sources() if force or not glob.glob(os.path.join(OBJDIR, "*.o")) else stale_objects()
    nvcc = _nvcc()

    def compile_one(src):
        cmd = [nvcc] + NVCC_FLAGS + EXTRA_FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", "-o", _obj(src), src]
        return src, subprocess.run(cmd, capture_output=True, text=True)

    with ThreadPoolExecutor(max_workers=jobs or min(len(todo) or 1, os.cpu_count() or 1)) as pool:
        results = list(pool.map(compile_one, todo))
    for src, r in results:
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError(f"nvcc failed on {os.path.basename(src)}")
        if verbose:
            sys.stderr.write(r.stderr)
    objs = [_obj(s) for s in sources()]
    for stray in set(glob.glob(os.path.join(OBJDIR, "*.o"))) - set(objs):
        os.remove(stray)
    r = subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", LIB] + objs,
                       capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed linking libffm_b200.so")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
