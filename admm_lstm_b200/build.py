"""Builds the C-ABI shared library (admm_lstm_b200/libadmm_lstm_b200.so) with nvcc for sm_100a.

In-tree build: the .so travels to the GPU box with the repo snapshot.  Also builds the host-only
formula harness used by the CPU tests (tests/test_point_math.py).
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OBJ = os.path.join(PKG, "build")
LIB = os.path.join(PKG, "libadmm_lstm_b200.so")
HOSTLIB = os.path.join(PKG, "libadmm_point_math_host.so")
SOURCES = ["capi.cu", "gate_gemm_simt.cu", "atr_simt.cu", "small_kernels.cu", "gate_gemm_tc.cu", "probe_eval.cu", "grad_from_z.cu", "admm_l.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-Xcompiler", "-fPIC", "--expt-relaxed-constexpr"]


def _nvcc():
    for cand in (os.environ.get("NVCC"), shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: the CUDA library cannot be built (there is no CPU fallback)")


def _deps():
    out = [os.path.join(PKG, "..", "include", "admm_lstm_b200.h")]
    for f in os.listdir(CSRC):
        out.append(os.path.join(CSRC, f))
    return out


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _source_hash(deps) -> str:
    """Content hash of every source the library is built from (+ the flags): what decides whether a shipped .so is
    current -- file times do not survive a snapshot copy."""
    import hashlib
    h = hashlib.sha256(" ".join(NVCC_FLAGS + SOURCES).encode())
    for d in sorted(os.path.abspath(x) for x in deps):
        h.update(os.path.basename(d).encode())
        h.update(open(d, "rb").read())
    return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
    deps = _deps()
    stamp, want = LIB + ".srchash", _source_hash(deps)
    have = open(stamp).read().strip() if os.path.exists(stamp) else None
    if not force and os.path.exists(LIB) and have == want:
        return LIB                      # built from exactly these sources
    if not force and have is None and not _stale(LIB, deps) and shutil.which("nvcc") is None:
        return LIB                      # no stamp, no compiler: the shipped library is all there is
    nvcc = _nvcc()
    os.makedirs(OBJ, exist_ok=True)

    def compile_one(src):
        obj = os.path.join(OBJ, src.replace(".cu", ".o"))
        if force or have != want or _stale(obj, deps):
            cmd = [nvcc, *NVCC_FLAGS, "-c", os.path.join(CSRC, src), "-o", obj]
            if verbose:
                cmd.insert(-4, "-Xptxas")
                cmd.insert(-4, "-v")
            r = subprocess.run(cmd, capture_output=True, text=True)
            if r.returncode != 0:
                raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
            if verbose:
                sys.stderr.write(r.stderr)
        return obj

    with ThreadPoolExecutor(max_workers=min(8, len(SOURCES))) as ex:
        objs = list(ex.map(compile_one, SOURCES))
    r = subprocess.run([nvcc, "-shared", "-o", LIB, *objs, "-cudart", "static"], capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    with open(stamp, "w") as f:
        f.write(want + "\n")
    return LIB


def build_host_math(force: bool = False) -> str:
    """Host-only build of csrc/admm_math.cuh (same source the kernels inline) for CPU formula tests."""
    src = os.path.join(CSRC, "point_math_host.cpp")
    deps = [src, os.path.join(CSRC, "admm_math.cuh")]
    if force or _stale(HOSTLIB, deps):
        cxx = shutil.which("g++") or "g++"
        r = subprocess.run([cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-o", HOSTLIB, src],
                           capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"g++ failed:\n{r.stdout}\n{r.stderr}")
    return HOSTLIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="-v" in sys.argv))
    print(build_host_math(force="--force" in sys.argv))
