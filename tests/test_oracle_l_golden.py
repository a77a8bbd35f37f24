"""Pins oracle/admm_l_oracle.py (ADMM-LSTM-L) to fixtures produced by the reference's own update
functions (tests/golden/make_golden_l.py).  CPU only."""
import os

import numpy as np
import pytest

from oracle.admm_l_oracle import GATES, OracleADMML

GOLD = os.path.join(os.path.dirname(__file__), "golden")


def rel(a, b):
    return float(np.max(np.abs(np.asarray(a) - np.asarray(b))) / (np.max(np.abs(b)) + 1e-30))


def make_oracle(d, x=None, y=None):
    W = {g: d[f"init_W{g}"] for g in GATES}
    U = {g: d[f"init_U{g}"] for g in GATES}
    return OracleADMML(W, U, d["init_Wy"], d["x"] if x is None else x, d["y"] if y is None else y)


def check_state(o, d, k, rows=None, tol=2e-5, dual_atol=2e-6):
    sl = slice(None) if rows is None else slice(0, rows)
    for g in GATES:
        for key, arr in ((f"z{g}", o.z[g]), (g, o.g[g])):
            assert rel(arr.transpose(1, 0, 2)[sl], d[f"it{k}_{key}"]) < tol, key
        # the duals of ADMM-LSTM-L sit at the fp32 rounding level (the constraints hold to rounding): absolute bound
        for key, arr in ((f"lams_{g}", o.lam_s[g]), (f"lamp_{g}", o.lam_p[g])):
            assert np.max(np.abs(arr.transpose(1, 0, 2)[sl] - d[f"it{k}_{key}"])) < dual_atol, key
    for key, arr in (("c", o.c), ("h", o.h)):
        assert rel(arr.transpose(1, 0, 2)[sl], d[f"it{k}_{key}"]) < tol, key
    for key, arr in (("lam9", o.lam9), ("lam10", o.lam10)):
        assert np.max(np.abs(arr.transpose(1, 0, 2)[sl] - d[f"it{k}_{key}"])) < dual_atol, key
    assert rel(o.a[sl], d[f"it{k}_a"]) < tol
    assert np.max(np.abs(o.lam11[sl] - d[f"it{k}_lam11"])) < dual_atol


@pytest.mark.parametrize("name", ["l_traj_small", "l_traj_h64"])
def test_l_oracle_trajectory(name):
    d = np.load(os.path.join(GOLD, name + ".npz"))
    o = make_oracle(d)
    for k in range(1, int(d["iters"]) + 1):
        o.step()
        for g in GATES:
            assert rel(o.W[g], d[f"it{k}_W{g}"]) < 1e-5, (k, g)
            assert rel(o.U[g], d[f"it{k}_U{g}"]) < 1e-5, (k, g)
        assert rel(o.Wy, d[f"it{k}_Wy"]) < 1e-5
        if f"it{k}_c" in d.files:
            check_state(o, d, k)
        assert abs(o.loss() - d["train_loss"][k]) < 1e-5 * max(1.0, d["train_loss"][k])


def test_l_oracle_googlestock_20_iterations():
    d = np.load(os.path.join(GOLD, "l_googlestock.npz"))
    data = np.load(os.path.join(GOLD, "googlestock_data.npz"))
    o = make_oracle(d, data["train_x"], data["train_y"])
    for k in range(1, 21):
        o.step()
        for g in GATES:
            assert rel(o.W[g], d[f"it{k}_W{g}"]) < 1e-4, (k, g)
            assert rel(o.U[g], d[f"it{k}_U{g}"]) < 1e-4, (k, g)
        assert rel(o.Wy, d[f"it{k}_Wy"]) < 1e-4
        assert abs(o.loss() - d["train_loss"][k]) < 1e-4 * d["train_loss"][k] + 1e-9
        assert abs(o.loss(data["val_x"], data["val_y"]) - d["val_loss"][k]) < 1e-4 * d["val_loss"][k] + 1e-9
    check_state(o, d, 20, rows=96, tol=1e-4)
