"""Pins oracle/admm_oracle.py against the UNMODIFIED reference run LIVE (oracle/_ref, the byte-for-byte copy staged by
oracle/make_ref.py; oracle/ref_runner.py runs it in its own process on the CPU) on problems that are NOT among the committed
fixtures: fresh seeds, odd shapes, both variants, two hyper-parameter sets.  Complements tests/test_oracle_golden.py (committed
vectors, which also travel to boxes without the reference).  Skipped where oracle/_ref is not staged."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

from oracle.admm_oracle import OracleADMM
from helpers import GOOGLE, HAR, WKEYS, rel_err, synthetic_problem

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
RUNNER = os.path.join(ROOT, "oracle", "ref_runner.py")

CASES = [  # variant, params, (N, T, D, H, O), seed, classification targets
    ("admm", GOOGLE, (29, 4, 5, 7, 1), 101, False),
    ("no_dual_y", GOOGLE, (41, 3, 2, 9, 3), 102, False),
    ("admm", HAR, (23, 6, 4, 5, 6), 103, True),
    ("no_dual_y", HAR, (64, 2, 7, 8, 2), 104, False),
]


@pytest.mark.parametrize("variant,params,shape,seed,cls", CASES)
def test_oracle_equals_live_reference(variant, params, shape, seed, cls, tmp_path):
    if not os.path.exists(os.path.join(REF, "admm.py")):
        pytest.skip("oracle/_ref not staged (python oracle/make_ref.py needs /root/reference)")
    n, t, d, h, o = shape
    x, y, w = synthetic_problem(n, t, d, h, o, seed=seed, classification=cls)
    data, dump = str(tmp_path / "problem.npz"), str(tmp_path / "ref_state.npz")
    np.savez(data, x=x, y=y, params_json=json.dumps(params), **w)
    steps = 3
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="")          # the reference picks CUDA when it sees one (_global.py:217)
    r = subprocess.run([sys.executable, RUNNER, "--data", data, "--variant", variant, "--steps", str(steps), "--threads", "2",
                        "--dump", dump], capture_output=True, text=True, env=env, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info.get("device") == "cpu" and len(info["step_s"]) == steps, info
    ref = dict(np.load(dump))
    ora = OracleADMM(w, x, y, params, variant=variant)
    for _ in range(steps):
        ora.step()
    for k in WKEYS:
        assert rel_err(ora.w[k], ref[f"w_{k}"]) < 1e-4, k
    for k in ("i", "f", "g", "o", "c", "h", "a"):
        assert rel_err(ora.gates[k], ref[f"gate_{k}"]) < 1e-4, k
    for k in ("i", "f", "g", "o", "c"):
        scale = max(float(np.abs(ref[f"dual_{k}"]).max()), float(params["rho"][k]))
        assert float(np.abs(ora.duals[k] - ref[f"dual_{k}"]).max()) <= 1e-4 * scale, k
