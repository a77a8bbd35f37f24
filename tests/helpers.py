"""Shared helpers for the parity tests (oracle <-> golden fixtures <-> CUDA path)."""
import os

import numpy as np

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
WKEYS = [f"{s}2{g}" for g in "ifgo" for s in "xh"] + ["out"]

GOOGLE = {"rho": {"i": 1., "f": 1., "g": 1., "o": 1., "c": 0.008, "h": 0.00045, "y": 0.0000562},
          "beta": {k: 8e-7 for k in ("wi", "vi", "wf", "vf", "wg", "vg", "wo", "vo", "wy")}}
HAR = {"rho": {"i": 1.5, "f": 1.5, "g": 1.5, "o": 1.5, "c": 0.005, "h": 8e-04, "y": 4e-04},
       "beta": {k: 8e-7 for k in ("wi", "vi", "wf", "vf", "wg", "vg", "wo", "vo", "wy")}}
GEFCOM = {"rho": {"i": 1, "f": 1, "g": 1, "o": 1, "c": 0.1, "h": 0.01, "y": 0.01},
          "beta": {k: 8e-7 for k in ("wi", "vi", "wf", "vf", "wg", "vg", "wo", "vo", "wy")}}
GEFCOM_FAST = {"rho": dict(GEFCOM["rho"], h=0.001, y=0.0001), "beta": GEFCOM["beta"]}


def load(name):
    return dict(np.load(os.path.join(GOLDEN, name)))


def weights_from(rec, prefix):
    return {k: rec[f"{prefix}w_{k}"] for k in WKEYS}


def state_from(rec, prefix):
    gates = {k: rec[f"{prefix}gate_{k}"] for k in ("i", "f", "g", "o", "c", "h", "a")}
    duals = {k: rec[f"{prefix}dual_{k}"] for k in ("i", "f", "g", "o", "c", "h", "y")}
    return {"gates": gates, "duals": duals}


def rel_err(a, b):
    """max |a-b| / max |b|  (the scale-relative error SURVEY 8(c) uses for weights/gates)."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / max(np.max(np.abs(b)), 1e-30))


def dual_atol(params, gate, gates_max):
    """duals are cancellation residues; atol = 1e-5 * rho_g * max|gate|."""
    return 1e-5 * float(params["rho"][gate]) * max(gates_max, 1.0)


def synthetic_problem(n, t, d, h, o, seed=0, classification=False):
    """Same generator for oracle and build: rand inputs, Xavier-normal weights."""
    rng = np.random.default_rng(seed)
    x = rng.random((n, t, d), dtype=np.float32)
    if classification:
        y = np.eye(o, dtype=np.float32)[rng.integers(0, o, size=n)]
    else:
        y = rng.random((n, o), dtype=np.float32)
    w = {}
    for g in "ifgo":
        w["x2" + g] = (rng.standard_normal((d, h)) * np.sqrt(2.0 / (d + h))).astype(np.float32)
        w["h2" + g] = (rng.standard_normal((h, h)) * np.sqrt(2.0 / (h + h))).astype(np.float32)
    w["out"] = (rng.standard_normal((h, o)) * np.sqrt(2.0 / (h + o))).astype(np.float32)
    return x, y, w
