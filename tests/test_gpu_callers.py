"""north_star: "demo.py and comparison_experiment/comparison.py call it unchanged".

The reference's own demo.py (training_demo -> init -> admm_demo, demo.py:311-409) and comparison.py (__main__, :141-210) are
executed UNMODIFIED -- from the byte-for-byte copy oracle/make_ref.py stages in oracle/_ref -- in a tree where only admm.py and
comparison_experiment/admm_l/main.py are this repo's drop-ins (tests/run_reference_callers.py).  The loss curves they return
must be the golden curves of the unmodified reference (BASELINE.md section 3 / tests/golden/googlestock_*.npz)."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import load

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "oracle", "_ref")
RUNNER = os.path.join(ROOT, "tests", "run_reference_callers.py")


def _run(caller, variant, epochs, tmp_path, self_check=False):
    if not os.path.exists(os.path.join(REF, "demo.py")):
        pytest.skip("oracle/_ref not staged (python oracle/make_ref.py needs /root/reference)")
    out = tmp_path / f"{caller}_{variant}.json"
    cmd = [sys.executable, RUNNER, caller, "--variant", variant, "--epochs", str(epochs), "--out", str(out)]
    if self_check:
        cmd.append("--self-check")
    env = dict(os.environ)
    env.pop("ADMM_LSTM_VARIANT", None)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=1500, env=env, cwd=str(tmp_path))
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    return json.load(open(out))


def test_harness_reproduces_golden_curve_with_the_reference_itself(tmp_path):
    """CPU check of the harness: with the reference's OWN admm.py in the overlay it must give the golden curve."""
    res = _run("demo", "admm", 3, tmp_path, self_check=True)
    rec = load("googlestock_admm.npz")
    np.testing.assert_allclose(res["train_loss"], rec["train_loss"][:4], rtol=2e-5)
    np.testing.assert_allclose(res["val_loss"], rec["val_loss"][:4], rtol=2e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("variant", ["admm", "no_dual_y"])
def test_reference_demo_runs_unchanged(variant, tmp_path):
    """demo.training_demo() (GoogleStock, hidden 10, seed 0; configs[0]/[1] of BASELINE.json) for 5 epochs: the curve it
    returns = the unmodified reference's curve to 1e-4."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    res = _run("demo", variant, 5, tmp_path)
    rec = load(f"googlestock_{variant}.npz")
    assert res["name"] == "Fast ADMM-LSTM" and len(res["train_loss"]) == 6
    np.testing.assert_allclose(res["train_loss"], rec["train_loss"][:6], rtol=1e-4)
    np.testing.assert_allclose(res["val_loss"], rec["val_loss"][:6], rtol=1e-4)


@pytest.mark.gpu
def test_reference_comparison_runs_unchanged(tmp_path):
    """comparison.py as __main__ (Fast ADMM-LSTM through demo.admm_demo, ADMM-LSTM-L through admm_l.main.admm_l_demo, then the
    reference's own SGD / Adam / Adagrad baselines), 3 epochs, --save: curves of both ADMM methods = the reference's own
    (same seed, same RNG draw order), and SAVED_MODELS holds the five pickles with the reference's class paths."""
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")
    res = _run("comparison", "no_dual_y", 3, tmp_path)
    rec = load("googlestock_no_dual_y.npz")
    np.testing.assert_allclose(res["fast"]["train_loss"], rec["train_loss"][:4], rtol=1e-4)
    np.testing.assert_allclose(res["fast"]["val_loss"], rec["val_loss"][:4], rtol=1e-4)
    # ADMM-LSTM-L inside comparison.py: values of the unmodified reference run through the same harness (--self-check)
    np.testing.assert_allclose(res["admm_l"]["train_loss"], [0.06116463989019394, 0.06116463989019394, 0.051507044583559036,
                                                             0.04065319150686264], rtol=1e-3)
    np.testing.assert_allclose(res["admm_l"]["val_loss"], [0.6165746450424194, 0.6165746450424194, 0.5193830728530884,
                                                           0.4101332426071167], rtol=1e-3)
    assert res["saved"] == ["ADMM-LSTM-L.pt", "Adagrad.pt", "Adam.pt", "Fast ADMM-LSTM.pt", "SGD.pt"]
    assert res["fast_pickle_class"] == "blocks.lstm.LSTM"
    assert res["fast_pickle_params"] == ["x2i", "h2i", "x2f", "h2f", "x2g", "h2g", "x2o", "h2o", "out"]
    assert res["l_pickle_class"] == "admm_l.main.LSTM_L"
