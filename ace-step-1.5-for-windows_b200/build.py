"""Builds libacestep_b200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m acestep_b200.build [--force]

The shared library is a plain C-ABI object (include/acestep_b200.h): no torch, no pybind.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG_DIR = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG_DIR, "csrc")
LIB_PATH = os.path.join(PKG_DIR, "libacestep_b200.so")
OBJ_DIR = os.path.join(PKG_DIR, "_obj")
# probe build: same sources + -DACE_PROBE (A/B switches, scalar reference GEMM, alternative kernel variants,
# ace_debug_set_* hooks) for tools/ and the two-path equivalence tests; never loaded by the product path
PROBE_LIB_PATH = os.path.join(PKG_DIR, "libacestep_b200_probe.so")
PROBE_OBJ_DIR = os.path.join(PKG_DIR, "_obj_probe")

SOURCES = ["runtime.cu", "elementwise.cu", "attention.cu", "attention_tc.cu", "dit.cu", "vae.cu", "cond.cu"]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr",
    # function-local statics of templates would otherwise be STB_GNU_UNIQUE: ONE copy per process even across
    # RTLD_LOCAL libraries, i.e. shared between the release and the probe library when a test loads both
    "-Xcompiler", "-fno-gnu-unique", "-Xcompiler", "-fvisibility=hidden",
]


def _nvcc() -> str:
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", shutil.which("nvcc")):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found (set NVCC or install the CUDA toolkit)")


def _digest() -> str:
    h = hashlib.sha256()
    for name in sorted(os.listdir(CSRC)):
        if name.endswith((".cu", ".cuh", ".h")):
            with open(os.path.join(CSRC, name), "rb") as f:
                h.update(name.encode())
                h.update(f.read())
    with open(os.path.join(PKG_DIR, "..", "include", "acestep_b200.h"), "rb") as f:
        h.update(f.read())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False, probe: bool = True) -> str:
    """Compile every .cu for sm_100a and link the release library (returns its path); with `probe` also the
    probe library next to it."""
    path = _build_one(LIB_PATH, OBJ_DIR, [], force, verbose)
    if probe:
        _build_one(PROBE_LIB_PATH, PROBE_OBJ_DIR, ["-DACE_PROBE"], force, verbose)
    return path


def _build_one(lib_path: str, obj_dir: str, defines, force: bool, verbose: bool) -> str:
    stamp = os.path.join(obj_dir, "digest.txt")
    digest = _digest()
    if not force and os.path.exists(lib_path) and os.path.exists(stamp):
        with open(stamp) as f:
            if f.read().strip() == digest:
                return lib_path
    os.makedirs(obj_dir, exist_ok=True)
    nvcc = _nvcc()
    sources = [s for s in SOURCES if os.path.exists(os.path.join(CSRC, s))]

    def compile_one(src: str) -> str:
        obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
        cmd = [nvcc, *NVCC_FLAGS, *defines, "-c", os.path.join(CSRC, src), "-o", obj]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
        if verbose:
            print(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(sources))) as ex:
        objs = list(ex.map(compile_one, sources))
    cmd = [nvcc, "-shared", "-o", lib_path, *objs, "-gencode", "arch=compute_100a,code=sm_100a", "-lcudart"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(digest)
    return lib_path


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
