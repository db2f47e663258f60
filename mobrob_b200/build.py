"""Build libmobrob_b200.so in-tree with nvcc for sm_100a (no torch headers, plain C ABI).

Every csrc/*.cu is compiled to an object in parallel (no relocatable device code: device functions
live in the .cuh headers) and linked into one shared library.  Staleness is decided by CONTENT: the
hash of all sources and flags is stored beside the library, so a snapshot copied to another box
(new mtimes) is not rebuilt, and an edited source always is.
"""
from __future__ import annotations

import glob
import hashlib
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(ROOT, "csrc")
LIB_DIR = os.path.join(ROOT, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libmobrob_b200.so")
OBJ_DIR = os.path.join(LIB_DIR, "obj")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a",
    "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _headers():
    return sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(
        os.path.join(os.path.dirname(ROOT), "include", "*.h")))


def _digest(paths, extra=()):
    h = hashlib.sha256()
    for x in extra:
        h.update(str(x).encode())
    for p in paths:
        h.update(os.path.basename(p).encode())
        with open(p, "rb") as f:
            h.update(f.read())
    return h.hexdigest()


def source_hash(defines=()):
    return _digest(sources() + _headers(), [*NVCC_FLAGS, *defines])


def _stamp(lib_path):
    return lib_path + ".srchash"


def needs_build(lib_path: str = LIB_PATH, defines=()) -> bool:
    if not os.path.exists(lib_path) or not os.path.exists(_stamp(lib_path)):
        return True
    with open(_stamp(lib_path)) as f:
        return f.read().strip() != source_hash(defines)


def build(force: bool = False, verbose: bool = False, lib_path: str = LIB_PATH, defines=()) -> str:
    """defines: extra -D flags (debug variants, e.g. ("-DMR_TRACE",) -> tools/trace_epoch.py)."""
    if not force and not needs_build(lib_path, defines):
        return lib_path
    tag = hashlib.sha256(" ".join(defines).encode()).hexdigest()[:8] if defines else "product"
    obj_dir = os.path.join(OBJ_DIR, tag)
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = os.environ.get("NVCC", "nvcc")
    hdr = _headers()

    def compile_one(src):
        obj = os.path.join(obj_dir, os.path.basename(src)[:-3] + ".o")
        want = _digest([src] + hdr, [*NVCC_FLAGS, *defines])
        stamp = obj + ".srchash"
        if not force and os.path.exists(obj) and os.path.exists(stamp) and open(stamp).read().strip() == want:
            return obj
        cmd = [nvcc, *NVCC_FLAGS, *defines, "-c", src, "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
            print(" ".join(cmd), flush=True)
        subprocess.check_call(cmd, cwd=CSRC)
        with open(stamp, "w") as f:
            f.write(want)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as pool:
        objs = list(pool.map(compile_one, sources()))
    subprocess.check_call([nvcc, "-shared", "-gencode", "arch=compute_100a,code=sm_100a", "-o", lib_path, *objs])
    with open(_stamp(lib_path), "w") as f:
        f.write(source_hash(defines))
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
