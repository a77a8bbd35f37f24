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


@pytest.mark.parametrize("shape,hidden,iters,seed", [((31, 4, 2), 5, 3, 201), ((53, 2, 6), 12, 3, 202)])
def test_l_oracle_equals_live_reference_functions(shape, hidden, iters, seed):
    """ADMM-LSTM-L: oracle/admm_l_oracle.py against the reference's own update functions (oracle/_ref/comparison_experiment/
    admm_l/admm_lstm.py, pure torch) driven live in the order of main.py:139-191 by tests/golden/make_golden_l.py's loop, on
    problems that are not among the fixtures."""
    import importlib.util
    import torch
    from test_oracle_l_golden import check_state, make_oracle, rel
    from oracle.admm_l_oracle import GATES
    ref_file = os.path.join(REF, "comparison_experiment", "admm_l", "admm_lstm.py")
    if not os.path.exists(ref_file):
        pytest.skip("oracle/_ref not staged (python oracle/make_ref.py needs /root/reference)")
    spec = importlib.util.spec_from_file_location("make_golden_l_live", os.path.join(ROOT, "tests", "golden", "make_golden_l.py"))
    gen = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gen)
    gen.REF = ref_file
    ref = gen.load_ref()
    torch.manual_seed(seed)
    x, y = torch.rand(*shape), torch.rand(shape[0], 1)
    d = gen.run_case(ref, x, y, None, None, hidden, iters, seed=seed + 1000)
    o = make_oracle(d)
    for k in range(1, iters + 1):
        o.step()
        for g in GATES:
            assert rel(o.W[g], d[f"it{k}_W{g}"]) < 1e-5, (k, g)
            assert rel(o.U[g], d[f"it{k}_U{g}"]) < 1e-5, (k, g)
        assert rel(o.Wy, d[f"it{k}_Wy"]) < 1e-5
        check_state(o, d, k)
        assert abs(o.loss() - d["train_loss"][k]) < 1e-5 * max(1.0, d["train_loss"][k])
