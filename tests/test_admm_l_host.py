"""Host logic of ADMM-LSTM-L on CPU: the Gram / right-hand-side formulation of the weight phase and the closed-form
backtracking exits (admm_lstm_b200/admm_l.py: l_weight_phase, exit_theta) against the reference-faithful oracle, with
the sums split over a world_size-2 gloo group exactly as the GPU ranks split them (Comm.allreduce_sum_ / _max_)."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORDER = ("i", "f", "g", "o")


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _problem():
    rng = np.random.default_rng(3)
    n, t, d, h = 157, 4, 3, 6
    x, y = rng.random((n, t, d), dtype=np.float32), rng.random((n, 1), dtype=np.float32)
    W = {g: (rng.standard_normal((d, h)) * 0.1).astype(np.float32) for g in "fiog"}
    U = {g: (rng.standard_normal((h, h)) * 0.1).astype(np.float32) for g in "fiog"}
    Wy = (rng.standard_normal((h, 1)) * 0.1).astype(np.float32)
    return x, y, W, U, Wy


def _local_sums(o, lo, hi):
    """The sums admm_l_sums / admm_l_sums_last / admm_l_gram_xx produce for the samples [lo, hi) (fp64)."""
    T, D, H = o.T, o.D, o.H
    ax, ah = np.zeros((5, D, H)), np.zeros((5, H, H))
    sxx = np.zeros((D, D))
    for t in range(T):
        xt = o.x[lo:hi, t, :].astype(np.float64)
        hp = (o.h[t - 1][lo:hi] if t > 0 else np.zeros((hi - lo, H))).astype(np.float64)
        sxx += xt.T @ xt
        for q, g in enumerate(ORDER):
            v = (o.z[g][t][lo:hi] + o.lam_s[g][t][lo:hi] / o.rs).astype(np.float64)
            ax[q] += xt.T @ v
            ah[q] += hp.T @ v
        ax[4] += xt.T @ hp
        ah[4] += hp.T @ hp
    hT = o.h[T - 1][lo:hi].astype(np.float64)
    stt = hT.T @ hT
    pt = (hT.T @ (o.a[lo:hi] + o.lam11[lo:hi] / o.r11).astype(np.float64)).reshape(-1)
    return sxx, ax, ah, stt, pt


def _worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from admm_lstm_b200.admm_l import l_weight_phase
        from admm_lstm_b200.comm import Comm
        from oracle.admm_l_oracle import OracleADMML
        comm = Comm()
        x, y, W, U, Wy = _problem()
        o = OracleADMML(W, U, Wy, x, y, n_norm=x.shape[0])
        for _ in range(2):
            o.step()
        lo, hi = comm.shard_range(x.shape[0])
        sums = [torch.from_numpy(np.ascontiguousarray(v)) for v in _local_sums(o, lo, hi)]
        comm.allreduce_sum_(*sums)
        m = torch.tensor([float(np.max(np.abs(o.h[0][lo:hi])))], dtype=torch.float64)
        comm.allreduce_max_(m)
        assert abs(float(m) - float(np.max(np.abs(o.h[0])))) < 1e-12
        # the per-timestep pair of update_c (admm_lstm.py:225,230): one collective for the MAX and the SUM scalar
        mx = torch.tensor([float(np.max(np.abs(o.c[0][lo:hi])))], dtype=torch.float32)
        sm = torch.tensor([float(np.sum(np.square(o.h[0][lo:hi]), dtype=np.float64))], dtype=torch.float64)
        comm.allreduce_max_and_sum_(mx, sm)
        assert abs(float(mx) - float(np.max(np.abs(o.c[0])))) < 1e-6
        assert abs(float(sm) - float(np.sum(np.square(o.h[0]), dtype=np.float64))) < 1e-9 * float(sm)
        wx = torch.stack([torch.from_numpy(o.W[g].copy()) for g in ORDER])
        wh = torch.stack([torch.from_numpy(o.U[g].copy()) for g in ORDER])
        wy = torch.from_numpy(o.Wy.copy()).reshape(-1)
        thetas = {}
        l_weight_phase(*sums[:1], sums[1], sums[2], sums[3], sums[4], wx, wh, wy, rho_s=1.0, rho11=float(o.r11),
                       lam_w=1e-6, lam_u=1e-6, thetas=thetas)
        np.savez(os.path.join(out_dir, f"rank{rank}.npz"), wx=wx.numpy(), wh=wh.numpy(), wy=wy.numpy(),
                 **{"th_" + k: float(v) for k, v in thetas.items()})
    finally:
        dist.destroy_process_group()


def test_gram_weight_phase_two_ranks_matches_oracle(tmp_path):
    sys.path.insert(0, ROOT)
    from oracle.admm_l_oracle import OracleADMML
    mp.spawn(_worker, args=(2, _free_port(), str(tmp_path)), nprocs=2, join=True)
    r0, r1 = (dict(np.load(tmp_path / f"rank{r}.npz")) for r in range(2))
    x, y, W, U, Wy = _problem()
    o = OracleADMML(W, U, Wy, x, y, n_norm=x.shape[0])
    for _ in range(2):
        o.step()
    o.update_wy()                                   # the reference's own weight phase (per-sample residuals, fp32 loops)
    for g in ("g", "o", "i", "f"):
        o.update_weight(g, "x")
        o.update_weight(g, "h")

    def rel(a, b):
        return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30))
    for k in ("wx", "wh", "wy"):
        assert np.array_equal(r0[k], r1[k]), k       # replicas stay bit-identical
    for q, g in enumerate(ORDER):
        assert rel(r0["wx"][q], o.W[g]) < 1e-5, ("W", g, rel(r0["wx"][q], o.W[g]))
        assert rel(r0["wh"][q], o.U[g]) < 1e-5, ("U", g, rel(r0["wh"][q], o.U[g]))
    assert rel(r0["wy"], o.Wy.reshape(-1)) < 1e-5
    # The closed form gives the exact-arithmetic exit.  The reference compares two fp32 sums over all samples that share
    # the large term 0.5*Form11; when the gradient is at the rounding level (the constraints hold to rounding in early
    # iterations) its comparison is decided by that noise and may stop one doubling away -- with no effect on the weights
    # at the tolerance above (the step G/theta is then ~1e-6 of the weight).
    same = 0
    for k, v in o.theta.items():
        if k != "h":
            ratio = float(r0["th_" + k]) / v
            assert min(abs(ratio - c) for c in (0.5, 1.0, 2.0)) < 1e-6, (k, float(r0["th_" + k]), v)
            same += abs(ratio - 1.0) < 1e-6
    assert same >= 6, same


def test_exit_theta_is_the_first_power_of_two_not_below_the_ratio():
    from admm_lstm_b200.admm_l import exit_theta
    f = lambda r, t0: float(exit_theta(torch.tensor(r, dtype=torch.float64), t0))      # noqa: E731
    assert f(0.3, 1.0) == 1.0 and f(1.0, 1.0) == 1.0 and f(1.0001, 1.0) == 2.0 and f(64.0, 1.0) == 64.0
    assert f(65.0, 1.0) == 128.0 and f(float("nan"), 1.0) == 1.0 and f(0.0, 0.01) == 0.01
    assert abs(f(0.05, 0.01) - 0.08) < 1e-15
