"""oracle/make_ref.py -- stage the UNMODIFIED reference for the GPU box  (TEST / BENCH INFRASTRUCTURE, NOT PRODUCT CODE).

The reference is pure Python (no native code, nothing to compile) and /root/reference does not exist
on the GPU box.  This recipe copies the reference's own source files for the hot path and its two callers, byte for
byte, from where they lie under /root/reference into oracle/_ref/ -- a directory that is git-ignored (never enters the
history) but NOT gpurun-ignored, so it travels to the box next to the built .so files.  Nothing is edited; SHA256SUMS in the
target records what was staged.  `__graft_entry__.build()` runs this whenever /root/reference is present.

Users (never the product path, admm_lstm_b200/ imports nothing from here):
  * bench.py --impl reference   : oracle/ref_runner.py times the reference's ADMMBasedOptimizer.step() on the host cores;
  * bench.py gpu_baseline        : the same unmodified step() with device='cuda' (eager torch), the "existing GPU path" bar
                                   of SURVEY.md section 2a / 8(d);
  * tests/test_gpu_callers.py    : the reference's demo.admm_demo / comparison.py call sites run UNCHANGED against the
                                   repo-root admm.py (north_star: "demo.py and comparison.py call it unchanged").

    python oracle/make_ref.py [--src /root/reference] [--dst oracle/_ref]
"""
from __future__ import annotations

import argparse
import hashlib
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))

# the hot path, its model / parameter / runtime modules, and the two callers of north_star
FILES = [
    "admm.py", "admm.no_dual_y.py", "_global.py", "parameters.py", "blocks/lstm.py",
    "demo.py", "dataset.py", "data_plot.py",
    "comparison_experiment/comparison.py",
    "comparison_experiment/admm_l/admm_lstm.py", "comparison_experiment/admm_l/main.py",
    "comparison_experiment/grad_based/grad_based.py",
    "comparison_experiment/admm_s/results.py",
    "datasets/GoogleStock/GOOG.xls",          # data file (678 KB) read by demo.py's default dataset loader (dataset.py:392-401)
]
OPTIONAL = ["blocks/__init__.py", "comparison_experiment/__init__.py", "comparison_experiment/admm_l/__init__.py",
            "comparison_experiment/grad_based/__init__.py", "comparison_experiment/admm_s/__init__.py"]


def stage(src: str, dst: str) -> int:
    if not os.path.isdir(src):
        print(f"make_ref: {src} not present (GPU box): keeping whatever is staged in {dst}")
        return 0
    os.makedirs(dst, exist_ok=True)
    sums = []
    for rel in FILES + OPTIONAL:
        s = os.path.join(src, rel)
        if not os.path.exists(s):
            if rel in OPTIONAL:
                continue
            raise FileNotFoundError(s)
        d = os.path.join(dst, rel)
        os.makedirs(os.path.dirname(d), exist_ok=True)
        shutil.copyfile(s, d)
        sums.append(f"{hashlib.sha256(open(d, 'rb').read()).hexdigest()}  {rel}")
    with open(os.path.join(dst, "SHA256SUMS"), "w") as f:
        f.write("\n".join(sums) + "\n")
    print(f"make_ref: staged {len(sums)} unmodified reference files into {dst}")
    return len(sums)


if __name__ == "__main__":
    ap = argparse.ArgumentParser()
    ap.add_argument("--src", default="/root/reference")
    ap.add_argument("--dst", default=os.path.join(HERE, "_ref"))
    a = ap.parse_args()
    stage(a.src, a.dst)
    sys.exit(0)
