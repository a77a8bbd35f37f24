"""GPU parity tests of ADMM-LSTM-L: the CUDA path (through the C ABI) against the fixtures produced
by the reference's own update functions and against the oracle."""
import os

import numpy as np
import pytest
import torch

from oracle.admm_l_oracle import GATES, OracleADMML

pytestmark = pytest.mark.gpu
GOLD = os.path.join(os.path.dirname(__file__), "golden")


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def rel(a, b):
    a, b = np.asarray(a), np.asarray(b)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-30))


def make_opt(d, x, y, **kw):
    from admm_lstm_b200.admm_l import ADMMLOptimizer
    w = {}
    for g in GATES:
        w["W" + g], w["U" + g] = torch.from_numpy(d[f"init_W{g}"]), torch.from_numpy(d[f"init_U{g}"])
    w["Wy"] = torch.from_numpy(d["init_Wy"])
    return ADMMLOptimizer(w, torch.from_numpy(np.asarray(x)), torch.from_numpy(np.asarray(y)), **kw)


def check_weights(opt, d, k, tol):
    w = {n: v.cpu().numpy() for n, v in opt.weights().items()}
    for g in GATES:
        assert rel(w["W" + g], d[f"it{k}_W{g}"]) < tol, (k, "W" + g, rel(w["W" + g], d[f"it{k}_W{g}"]))
        assert rel(w["U" + g], d[f"it{k}_U{g}"]) < tol, (k, "U" + g, rel(w["U" + g], d[f"it{k}_U{g}"]))
    assert rel(w["Wy"], d[f"it{k}_Wy"]) < tol, (k, "Wy")


def check_state(opt, d, k, rows=None, tol=1e-4, dual_atol=5e-6):
    st = {n: v.cpu().numpy() for n, v in opt.state().items()}
    sl = slice(None) if rows is None else slice(0, rows)
    for g in GATES:
        for key in (f"z{g}", g):
            assert rel(st[key][sl], d[f"it{k}_{key}"]) < tol, (key, rel(st[key][sl], d[f"it{k}_{key}"]))
        for key in (f"lams_{g}", f"lamp_{g}"):          # duals sit at the fp32 rounding level: absolute bound
            assert np.max(np.abs(st[key][sl] - d[f"it{k}_{key}"])) < dual_atol, key
    for key in ("c", "h", "a"):
        assert rel(st[key][sl], d[f"it{k}_{key}"]) < tol, (key, rel(st[key][sl], d[f"it{k}_{key}"]))
    for key in ("lam9", "lam10", "lam11"):
        assert np.max(np.abs(st[key][sl] - d[f"it{k}_{key}"])) < dual_atol, key


@pytest.mark.parametrize("name,tc", [("l_traj_small", False), ("l_traj_h64", False), ("l_traj_h64", True)])
def test_l_trajectory_vs_reference_fixture(name, tc):
    """Fixtures: the reference's OWN update functions (admm_l/admm_lstm.py) called by tests/golden/make_golden_l.py in the order
    of main.py:139-191 -- the driver loop is restated there (admm_l_demo returns only losses); the unmodified admm_l_demo itself
    is exercised end to end by tests/test_gpu_callers.py::test_reference_comparison_runs_unchanged."""
    _need_gpu()
    d = np.load(os.path.join(GOLD, name + ".npz"))
    opt = make_opt(d, d["x"], d["y"], use_tensor_cores=tc)
    assert opt.uses_tensor_cores == tc
    for k in range(1, int(d["iters"]) + 1):
        opt.step()
        check_weights(opt, d, k, 1e-4)
        if f"it{k}_c" in d.files:
            check_state(opt, d, k)


def test_l_googlestock_20_iterations_vs_reference_curves():
    _need_gpu()
    d = np.load(os.path.join(GOLD, "l_googlestock.npz"))
    data = np.load(os.path.join(GOLD, "googlestock_data.npz"))
    opt = make_opt(d, data["train_x"], data["train_y"])
    tx, ty = torch.from_numpy(data["train_x"]).cuda(), torch.from_numpy(data["train_y"]).cuda()
    vx, vy = torch.from_numpy(data["val_x"]).cuda(), torch.from_numpy(data["val_y"]).cuda()
    for k in range(1, 21):
        opt.step()
        check_weights(opt, d, k, 1e-4)
        tl = float(torch.mean(torch.square(ty - opt.predict(tx))))
        vl = float(torch.mean(torch.square(vy - opt.predict(vx))))
        assert abs(tl - d["train_loss"][k]) < 1e-3 * d["train_loss"][k] + 1e-9, (k, tl, d["train_loss"][k])
        assert abs(vl - d["val_loss"][k]) < 1e-3 * d["val_loss"][k] + 1e-9, (k, vl, d["val_loss"][k])
    check_state(opt, d, 20, rows=96)


@pytest.mark.parametrize("shape,tc", [((333, 4, 5, 20), False), ((500, 3, 16, 128), True), ((1, 1, 1, 1), False)])
def test_l_random_shapes_vs_oracle(shape, tc):
    _need_gpu()
    n, t, dd, h = shape
    rng = np.random.default_rng(n + h)
    x, y = rng.random((n, t, dd), dtype=np.float32), rng.random((n, 1), dtype=np.float32)
    d = {"init_Wy": (rng.standard_normal((h, 1)) * 0.1).astype(np.float32)}
    for g in GATES:
        d[f"init_W{g}"] = (rng.standard_normal((dd, h)) * 0.1).astype(np.float32)
        d[f"init_U{g}"] = (rng.standard_normal((h, h)) * 0.1).astype(np.float32)
    opt = make_opt(d, x, y, use_tensor_cores=tc)
    ora = OracleADMML({g: d[f"init_W{g}"] for g in GATES}, {g: d[f"init_U{g}"] for g in GATES}, d["init_Wy"], x, y)
    for k in range(3):
        opt.step()
        ora.step()
        w = {nm: v.cpu().numpy() for nm, v in opt.weights().items()}
        for g in GATES:
            assert rel(w["W" + g], ora.W[g]) < 1e-4, (k, g)
            assert rel(w["U" + g], ora.U[g]) < 1e-4, (k, g)
        assert rel(w["Wy"], ora.Wy) < 1e-4
        st = opt.state()
        assert rel(st["h"].cpu().numpy(), ora.h.transpose(1, 0, 2)) < 1e-4
        assert rel(st["c"].cpu().numpy(), ora.c.transpose(1, 0, 2)) < 1e-4
        assert rel(st["a"].cpu().numpy(), ora.a) < 1e-4


def test_l_demo_surface(tmp_path, monkeypatch):
    """comparison.py:174-178 call shape; SAVED_MODELS/ADMM-LSTM-L.pt round-trips as an LSTM_L module."""
    _need_gpu()
    monkeypatch.chdir(tmp_path)
    from comparison_experiment.admm_l.main import admm_l_demo
    data = np.load(os.path.join(GOLD, "googlestock_data.npz"))
    tx, ty = torch.from_numpy(data["train_x"]), torch.from_numpy(data["train_y"])
    vx, vy = torch.from_numpy(data["val_x"]), torch.from_numpy(data["val_y"])
    torch.manual_seed(0)
    out = admm_l_demo(5, 10, tx, ty, vx, vy, save=True)
    ref = np.load(os.path.join(GOLD, "l_googlestock.npz"))        # generated with the same seed and draw order
    assert out["name"] == "ADMM-LSTM-L" and len(out["train_loss"]) == 6 and len(out["val_loss"]) == 6
    np.testing.assert_allclose(out["train_loss"], ref["train_loss"][:6], rtol=1e-3)
    np.testing.assert_allclose(out["val_loss"], ref["val_loss"][:6], rtol=1e-3)
    m = torch.load(tmp_path / "SAVED_MODELS" / "ADMM-LSTM-L.pt", weights_only=False)
    assert [n for n, _ in m.named_parameters()] == ["W_hi", "W_ii", "W_hf", "W_if", "W_ho", "W_io", "W_hg", "W_ig", "W_y"]


@pytest.mark.parametrize("shape,tc", [((301, 4, 5, 12), False), ((600, 3, 16, 128), True)])
def test_l_sharded_fake_world_equals_single(shape, tc):
    """Two sample shards stepped by two threads whose all-reduces (SUM of the Gram / RHS sums, MAX and SUM of the
    per-timestep scalars) are an in-process rendezvous must reproduce the unsharded run."""
    _need_gpu()
    import threading
    n, t, dd, h = shape
    rng = np.random.default_rng(7)
    x, y = rng.random((n, t, dd), dtype=np.float32), rng.random((n, 1), dtype=np.float32)
    d = {"init_Wy": (rng.standard_normal((h, 1)) * 0.1).astype(np.float32)}
    for g in GATES:
        d[f"init_W{g}"] = (rng.standard_normal((dd, h)) * 0.1).astype(np.float32)
        d[f"init_U{g}"] = (rng.standard_normal((h, h)) * 0.1).astype(np.float32)
    ref = make_opt(d, x, y, use_tensor_cores=tc, n_norm=float(n))

    class FakeComm:
        def __init__(self, rank, shared):
            self.active, self.world_size, self.rank, self.shared = True, 2, rank, shared

        def _reduce(self, t, op):
            torch.cuda.current_stream().synchronize()
            self.shared["buf"][self.rank] = t
            self.shared["bar"].wait()
            total = op(self.shared["buf"][0], self.shared["buf"][1])
            self.shared["bar"].wait()
            t.copy_(total)
            self.shared["bar"].wait()

        def allreduce_sum_(self, *tensors):
            for t in tensors:
                self._reduce(t, torch.add)

        def allreduce_max_(self, *tensors):
            for t in tensors:
                self._reduce(t, torch.maximum)

        def allreduce_max_and_sum_(self, max_t, sum_t):
            self._reduce(max_t, torch.maximum)
            self._reduce(sum_t, torch.add)

        def shard_range(self, n_total):
            half = (n_total + 1) // 2
            return (0, half) if self.rank == 0 else (half, n_total)

        def sum_int(self, v, device):
            return v

    shared = {"buf": [None, None], "bar": threading.Barrier(2)}
    opts, errs = [None, None], []

    def worker(rank):
        try:
            torch.cuda.set_device(0)
            o = make_opt(d, x, y, use_tensor_cores=tc, n_norm=float(n), comm=FakeComm(rank, shared), sharding="slice")
            opts[rank] = o
            for _ in range(3):
                o.step()
            torch.cuda.synchronize()
        except Exception as exc:   # pragma: no cover
            errs.append(exc)
            shared["bar"].abort()

    th = [threading.Thread(target=worker, args=(r,)) for r in (0, 1)]
    [t_.start() for t_ in th]
    [t_.join() for t_ in th]
    assert not errs, errs
    for _ in range(3):
        ref.step()
    wr = {k: v.cpu().numpy() for k, v in ref.weights().items()}
    w0 = {k: v.cpu().numpy() for k, v in opts[0].weights().items()}
    w1 = {k: v.cpu().numpy() for k, v in opts[1].weights().items()}
    for k in wr:
        assert np.array_equal(w0[k], w1[k]), k            # replicas stay bit-identical
        assert rel(w0[k], wr[k]) < 2e-5, (k, rel(w0[k], wr[k]))
    full = ref.state()["h"].cpu().numpy()
    half = opts[0].n_local
    assert rel(opts[0].state()["h"].cpu().numpy(), full[:half]) < 2e-5
    assert rel(opts[1].state()["h"].cpu().numpy(), full[half:]) < 2e-5
