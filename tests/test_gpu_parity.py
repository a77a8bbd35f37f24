"""Parity of the CUDA path (through the C ABI) with the oracle and with the golden fixtures made from
the unmodified reference.  Tolerances: weights / gates / a / MSE /
objective relative <= 1e-4 (scale-relative: max|diff| / max|ref|); duals absolute
1e-4 * max(max|dual|, rho) (they are cancellation residues); final train/val loss within 1 %."""
import os

import numpy as np
import pytest
import torch

from helpers import (GOOGLE, HAR, GEFCOM, GEFCOM_FAST, WKEYS, load, weights_from, state_from, rel_err,
                     synthetic_problem)

pytestmark = pytest.mark.gpu

REL = 1e-4


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _cmp_state(opt, ora, params, tag, rel=REL):
    from gpu_utils import np_state, weights_of
    gates, duals = np_state(opt)
    w = weights_of(opt)
    for k in WKEYS:
        assert rel_err(w[k], ora.w[k]) < rel, (tag, "weight", k, rel_err(w[k], ora.w[k]))
    for k in ("i", "f", "g", "o", "c", "h", "a"):
        assert rel_err(gates[k], ora.gates[k]) < rel, (tag, "gate", k, rel_err(gates[k], ora.gates[k]))
    for k in ("i", "f", "g", "o", "c", "h", "y"):
        scale = max(float(np.max(np.abs(ora.duals[k]))), float(params["rho"][k]))
        err = float(np.max(np.abs(duals[k] - ora.duals[k])))
        assert err < 1e-4 * scale, (tag, "dual", k, err, scale)


@pytest.mark.parametrize("variant", ["admm", "no_dual_y"])
def test_per_function_vs_reference_fixture(variant):
    """Each C-ABI phase from the same non-trivial state the reference's private methods were called on."""
    _need_gpu()
    from oracle.admm_oracle import OracleADMM
    from gpu_utils import make_opt, load_state, weights_of, np_state
    from admm_lstm_b200 import _lib
    rec = load(f"fn_{variant}.npz")
    base_w = {k: rec[f"base_w_{k}"] for k in WKEYS}
    st = state_from(rec, "base_")
    tt, T = int(rec["tt"]), rec["x"].shape[1]
    # invariant of every real run (admm.py:533-534): lambda_h is zero for t < T; the device layout
    # stores only the t = T slot, so the sweep comparison starts from a state that honours it.
    st["duals"]["h"] = st["duals"]["h"].copy()
    st["duals"]["h"][:, :T, :] = 0

    def fresh():
        model, opt = make_opt(base_w, rec["x"], rec["y"], GOOGLE, variant)
        load_state(opt, st)
        return opt

    # Wy  (admm.py:246-280)
    opt = fresh()
    opt._ADMMBasedOptimizer__update_wy(torch.cuda.current_stream().cuda_stream)
    assert rel_err(weights_of(opt)["out"], rec["fn_wy"]) < 2e-5
    # W: all four gates at once from the base state (admm.py:282-343)
    opt = fresh()
    opt._ADMMBasedOptimizer__update_weights(_lib.SRC_X, torch.cuda.current_stream().cuda_stream)
    w = weights_of(opt)
    for g in "igf":
        assert rel_err(w["x2" + g], rec[f"fn_w_x2{g}"]) < 2e-5, g
    opt = fresh()
    opt._ADMMBasedOptimizer__update_weights(_lib.SRC_H, torch.cuda.current_stream().cuda_stream)
    w = weights_of(opt)
    for g in "igo":
        assert rel_err(w["h2" + g], rec[f"fn_w_h2{g}"]) < 2e-5, g

    # fused sweep at an interior t and at t = T against the oracle's sequential updates
    # (the oracle itself is pinned per function to the reference in test_oracle_golden.py)
    for t in (tt, T):
        opt = fresh()
        ora = OracleADMM(base_w, rec["x"], rec["y"], GOOGLE, variant=variant, state=st)
        ora.update_gates(t)
        if t == T:
            ora.update_primal_a()
        ora.update_duals(t)
        s = torch.cuda.current_stream().cuda_stream
        opt._metrics.zero_()
        opt._call("admm_sweep_t", opt._pp, t, opt._metrics.data_ptr(), s)
        if t == T:
            opt._ADMMBasedOptimizer__update_last(s)
        gates, duals = np_state(opt)
        for k in ("i", "f", "g", "o", "c", "h"):
            np.testing.assert_allclose(gates[k][:, t], ora.gates[k][:, t], rtol=2e-5, atol=2e-6, err_msg=f"{k}@{t}")
            np.testing.assert_allclose(duals[k][:, t], ora.duals[k][:, t], rtol=2e-5, atol=2e-6, err_msg=f"lam_{k}@{t}")
        np.testing.assert_allclose(gates["a"], ora.gates["a"], rtol=2e-5, atol=2e-6)
        if t == T:
            assert abs(opt.theta_trace()["h_T"] - ora.trace["h_T"]) < 1e-6
    # direct per-function fixtures that are order independent (t < T h update, a update)
    opt = fresh()
    s = torch.cuda.current_stream().cuda_stream
    opt._call("admm_sweep_t", opt._pp, tt, 0, s)
    gates, _ = np_state(opt)
    # i is updated first and does not involve lambda_h: directly comparable with the reference's own output
    np.testing.assert_allclose(gates["i"][:, tt], rec["fn_primal_i"], rtol=2e-5, atol=2e-6)


@pytest.mark.parametrize("name,variant,params,dualy", [
    ("traj_admm.npz", "admm", GOOGLE, False),
    ("traj_no_dual_y.npz", "no_dual_y", GOOGLE, False),
    ("traj_har_admm.npz", "admm", HAR, False),
    ("traj_har_no_dual_y.npz", "no_dual_y", HAR, False),
    ("traj_admm_dualy.npz", "admm", GOOGLE, True),
])
def test_trajectory_vs_reference_fixture(name, variant, params, dualy):
    _need_gpu()
    from gpu_utils import make_opt, np_state, weights_of
    rec = load(name)
    model, opt = make_opt(weights_from(rec, "init_"), rec["x"], rec["y"], params, variant, with_dual_y=dualy)
    gates, _ = np_state(opt)
    for k in ("i", "f", "g", "o", "c", "h", "a"):
        np.testing.assert_allclose(gates[k], rec[f"s0_gate_{k}"], rtol=1e-5, atol=1e-6)
    x = torch.from_numpy(rec["x"]).cuda()
    y = torch.from_numpy(rec["y"]).cuda()
    for s in range(1, len(rec["losses"])):
        opt.step()
        w = weights_of(opt)
        gates, duals = np_state(opt)
        for k in WKEYS:
            assert rel_err(w[k], rec[f"s{s}_w_{k}"]) < REL, (s, k)
        for k in ("i", "f", "g", "o", "c", "h", "a"):
            assert rel_err(gates[k], rec[f"s{s}_gate_{k}"]) < REL, (s, k)
        for k in ("i", "f", "g", "o", "c", "h", "y"):
            scale = max(float(np.max(np.abs(rec[f"s{s}_dual_{k}"]))), float(params["rho"][k]))
            assert float(np.max(np.abs(duals[k] - rec[f"s{s}_dual_{k}"]))) < 1e-4 * scale, (s, k)
        with torch.no_grad():
            loss = float(torch.nn.functional.mse_loss(model(x), y))
        assert abs(loss - rec["losses"][s]) < REL * rec["losses"][s]


@pytest.mark.parametrize("dataset,variant,params", [
    ("googlestock", "admm", GOOGLE), ("googlestock", "no_dual_y", GOOGLE),
    ("gefcom_standin", "admm", GEFCOM), ("gefcom_standin", "no_dual_y", GEFCOM_FAST),
])
def test_real_data_50_iterations_vs_reference_curves(dataset, variant, params):
    """BASELINE.json configs 1-2 against the golden curves of the unmodified reference: 50 iterations,
    weights, loss curves and state snapshots.

    Tolerance note: on GoogleStock the x2g backtracking sits on a knife edge in EVERY
    iteration (D = 1 and sum x^2 / T is ~2^8, so f(beta)-est at the deciding theta is ~1e-8 against terms of
    1e-3); which side it falls on depends on the last bit of tanh.  A flipped theta is an equally valid ADMM
    iterate that moves the weights by ~1e-4..1e-3 for the rest of the run (measured by forcing one flip in
    the oracle).  So: weights/loss must match to 1e-4 until the first flip and to 2e-3 after it, the final
    losses to 1 %; the companion test below pins the arithmetic to 1e-4 with the decisions synchronised."""
    _need_gpu()
    from gpu_utils import make_opt, np_state, weights_of
    data = load(f"{dataset}_data.npz")
    rec = load(f"{dataset}_{variant}.npz")
    model, opt = make_opt(weights_from(rec, "init_"), data["train_x"], data["train_y"], params, variant)
    tx, ty = torch.from_numpy(data["train_x"]).cuda(), torch.from_numpy(data["train_y"]).cuda()
    vx, vy = torch.from_numpy(data["val_x"]).cuda(), torch.from_numpy(data["val_y"]).cuda()
    worst, first_flip = 0.0, None
    for it in range(1, 51):
        opt.step()
        w = weights_of(opt)
        e = max(rel_err(w[k], rec["wtraj_" + k][it]) for k in WKEYS)
        worst = max(worst, e)
        if e >= REL and first_flip is None:
            first_flip = it
        assert e < (REL if first_flip is None else 2e-3), (it, e, first_flip)
        with torch.no_grad():
            tr = float(torch.nn.functional.mse_loss(model(tx), ty))
            va = float(torch.nn.functional.mse_loss(model(vx), vy))
        tol = 1e-3 if first_flip is None else 3e-3
        assert abs(tr - rec["train_loss"][it]) < tol * rec["train_loss"][it], (it, tr, rec["train_loss"][it])
        assert abs(va - rec["val_loss"][it]) < tol * rec["val_loss"][it], (it, va, rec["val_loss"][it])
        if f"it{it}_gate_i" in rec and first_flip is None:
            gates, duals = np_state(opt)
            rows = rec[f"it{it}_gate_i"].shape[0]
            for k in ("i", "f", "g", "o", "c", "h"):
                assert rel_err(gates[k][:rows], rec[f"it{it}_gate_{k}"]) < REL, (it, k)
                scale = max(float(np.max(np.abs(rec[f"it{it}_dual_{k}"]))), float(params["rho"][k]))
                assert float(np.max(np.abs(duals[k][:rows] - rec[f"it{it}_dual_{k}"]))) < 1e-3 * scale, (it, k)
    assert abs(tr - rec["train_loss"][50]) < 0.01 * rec["train_loss"][50]
    assert abs(va - rec["val_loss"][50]) < 0.01 * rec["val_loss"][50]
    # With the default moment probes the decisions are the exact-arithmetic ones and no run flips: strict
    # 1e-4 on every weight for all 50 iterations on both datasets and both variants.  (The candidate-by-candidate probes of
    # probe='exact' flip one knife-edge decision on GoogleStock/admm; the looser branch above is for them.)
    if dataset != "googlestock" or opt.probe == "moments":
        assert first_flip is None, first_flip
    print(f"{dataset}/{variant}: worst weight rel err over 50 iterations = {worst:.2e}, first theta flip at {first_flip}")


@pytest.mark.parametrize("variant", ["admm", "no_dual_y"])
def test_googlestock_50_iterations_decision_synchronised(variant):
    """Arithmetic parity over 50 GoogleStock iterations with the backtracking decisions synchronised: the
    oracle (pinned to the reference in test_oracle_golden.py) applies the theta the GPU chose; every
    weight / gate / a must then agree to 1e-4 at every iteration.  Where the oracle's own loop would have
    chosen differently, the comparison it lost must be a genuine knife edge (|f(beta)-est| tiny against
    the terms compared), and such flips must be rare."""
    _need_gpu()
    from oracle.admm_oracle import OracleADMM
    from gpu_utils import make_opt, np_state, weights_of
    data = load("googlestock_data.npz")
    rec = load(f"googlestock_{variant}.npz")
    w0 = weights_from(rec, "init_")
    _, opt = make_opt(w0, data["train_x"], data["train_y"], GOOGLE, variant)
    ora = OracleADMM(w0, data["train_x"], data["train_y"], GOOGLE, variant=variant)
    flips, noise_flips = [], []
    for it in range(1, 51):
        opt.step()
        chosen = opt.theta_trace()
        ora.forced = dict(chosen)
        ora.step()
        n_elem = data["train_x"].shape[0] * data["train_x"].shape[1] * opt.hidden_size
        noise_floor = 100.0 * n_elem * (6e-8) ** 2
        for name, th in chosen.items():
            if name not in ora.trace or abs(ora.trace[name] - th) <= 1e-6 * th:
                continue
            assert name != "h_T", (it, ora.trace[name], th)      # the h_T loop has wide margins on these data
            # the comparison at the smaller of the two thetas (x2: the un-halved value) decided it
            th_dec = 2 * min(ora.trace[name], th)
            cmp = [m for m in ora.margins[name] if abs(m[0] - th_dec) < 1e-9 * th_dec]
            assert cmp, (it, name, ora.margins[name])
            _, fb, est = cmp[0]
            if max(fb, est) < noise_floor:
                # f itself is rounding noise: a sum over N*T*H residuals that are each a cancellation at the
                # fp32 ulp level (gate o early in the run; the "absorption exits" of SURVEY section 7).  Both
                # implementations decide on noise there and the resulting step is ~1e-9 of the weight.
                noise_flips.append((it, name, ora.trace[name], th))
                continue
            scale = max(abs(fb), abs(est), 1e-30)
            assert abs(fb - est) < 1e-4 * scale, ("not a knife edge", it, name, fb, est)
            flips.append((it, name, ora.trace[name], th, fb, est))
        w = weights_of(opt)
        gates, _ = np_state(opt)
        for k in WKEYS:
            assert rel_err(w[k], ora.w[k]) < REL, (it, k, rel_err(w[k], ora.w[k]))
        for k in ("i", "f", "g", "o", "c", "h", "a"):
            assert rel_err(gates[k], ora.gates[k]) < REL, (it, k)
    print(f"googlestock/{variant}: {len(flips)} knife-edge theta flips vs the oracle's own decisions: {flips}; "
          f"{len(noise_flips)} decisions taken on rounding noise by both sides")
    assert len(flips) <= 8, flips


@pytest.mark.parametrize("shape,params,variant,cls", [
    ((300, 6, 5, 40, 3), HAR, "admm", True),        # H not a multiple of 32, N not a multiple of 4
    ((517, 3, 7, 33, 1), GOOGLE, "no_dual_y", False),
    ((1, 2, 1, 1, 1), GOOGLE, "admm", False),        # degenerate sizes
    ((130, 1, 2, 3, 2), GOOGLE, "admm", False),      # T = 1: the only timestep is the last one
    ((1000, 4, 16, 64, 6), HAR, "no_dual_y", True),
])
def test_random_shapes_vs_oracle(shape, params, variant, cls):
    _need_gpu()
    from oracle.admm_oracle import OracleADMM
    from gpu_utils import make_opt
    n, t, d, h, o = shape
    x, y, w = synthetic_problem(n, t, d, h, o, seed=n + h, classification=cls)
    model, opt = make_opt(w, x, y, params, variant)
    ora = OracleADMM(w, x, y, params, variant=variant)
    _cmp_state(opt, ora, params, "init", rel=1e-5)
    for s in range(3):
        prev = ora.snapshot_primal()
        ora.step()
        opt.step()
        _cmp_state(opt, ora, params, f"step{s}")
        m_o, m_g = ora.metrics(prev), opt.metrics()
        # the residual norms are differences of O(1) state entries: a 1e-4-relative state agreement bounds them
        # absolutely, by ~1e-6 * sqrt(number of entries)
        atol = 1e-6 * np.sqrt(n * t * h)
        for key in ("objective", "primal_residual", "dual_residual", "loss_term"):
            assert abs(m_g[key] - m_o[key]) <= 2e-4 * abs(m_o[key]) + atol, (s, key, m_g[key], m_o[key])


def test_scratch_chunking_is_invisible():
    """The weight phase walks the timesteps in chunks sized by the scratch budget; results must not depend on it."""
    _need_gpu()
    from gpu_utils import make_opt, weights_of
    x, y, w = synthetic_problem(200, 7, 3, 12, 2, seed=4)
    _, a = make_opt(w, x, y, GOOGLE, "admm")
    _, b = make_opt(w, x, y, GOOGLE, "admm", scratch_bytes=1)     # one timestep per chunk
    assert a._tc_chunk == 7 and b._tc_chunk == 1
    for _ in range(2):
        a.step()
        b.step()
    wa, wb = weights_of(a), weights_of(b)
    for k in WKEYS:
        assert rel_err(wa[k], wb[k]) < 1e-6, k


def test_sharded_fake_world_equals_single():
    """SURVEY section 4 'distributed' row on one GPU: two shards stepped by two threads whose all-reduce is an
    in-process rendezvous must reproduce the unsharded run (reduction-order tolerance)."""
    _need_gpu()
    import threading
    from gpu_utils import make_opt, weights_of
    x, y, w = synthetic_problem(301, 5, 3, 10, 2, seed=9)
    _, ref = make_opt(w, x, y, GOOGLE, "admm")

    class FakeComm:
        def __init__(self, rank, shared):
            self.active, self.world_size, self.rank, self.shared = True, 2, rank, shared

        def allreduce_sum_(self, *tensors):
            for t in tensors:
                torch.cuda.current_stream().synchronize()
                self.shared["buf"][self.rank] = t
                self.shared["bar"].wait()
                total = self.shared["buf"][0] + self.shared["buf"][1]
                self.shared["bar"].wait()
                t.copy_(total)
                self.shared["bar"].wait()

        def allreduce_max_(self, *tensors):
            for t in tensors:
                torch.cuda.current_stream().synchronize()
                self.shared["buf"][self.rank] = t
                self.shared["bar"].wait()
                total = torch.maximum(self.shared["buf"][0], self.shared["buf"][1])
                self.shared["bar"].wait()
                t.copy_(total)
                self.shared["bar"].wait()

        def shard_range(self, n_total):
            half = (n_total + 1) // 2
            return (0, half) if self.rank == 0 else (half, n_total)

        def sum_int(self, v, device):
            return v

    shared = {"buf": [None, None], "bar": threading.Barrier(2)}
    opts = [None, None]
    errs = []

    def worker(rank):
        try:
            torch.cuda.set_device(0)
            _, o = make_opt(w, x, y, GOOGLE, "admm", comm=FakeComm(rank, shared), sharding="slice")
            opts[rank] = o
            for _ in range(3):
                o.step()
            torch.cuda.synchronize()
        except Exception as exc:   # pragma: no cover
            errs.append(exc)
            shared["bar"].abort()

    th = [threading.Thread(target=worker, args=(r,)) for r in (0, 1)]
    [t.start() for t in th]
    [t.join() for t in th]
    assert not errs, errs
    for _ in range(3):
        ref.step()
    wr, w0, w1 = weights_of(ref), weights_of(opts[0]), weights_of(opts[1])
    for k in WKEYS:
        assert np.array_equal(w0[k], w1[k]), k           # replicas stay bit-identical
        assert rel_err(w0[k], wr[k]) < 1e-5, (k, rel_err(w0[k], wr[k]))
    full = ref.gates["h"].cpu().numpy()
    half = opts[0].n_local
    assert rel_err(opts[0].gates["h"].cpu().numpy(), full[:half]) < 1e-5
    assert rel_err(opts[1].gates["h"].cpu().numpy(), full[half:]) < 1e-5


def test_model_binding_and_checkpoint_layout(tmp_path):
    """demo.py:302-308 / visualization.py:47-54: torch.save(model) after training round-trips with the
    parameter names and order of the reference's pickles; model(x) sees the updated weights."""
    _need_gpu()
    from gpu_utils import make_opt
    x, y, w = synthetic_problem(64, 4, 2, 6, 1, seed=1)
    model, opt = make_opt(w, x, y, GOOGLE, "no_dual_y")
    before = model.out.detach().clone()
    opt.step()
    assert not torch.equal(before, model.out.detach())
    path = tmp_path / "Fast ADMM-LSTM.pt"
    torch.save(model, path)
    loaded = torch.load(path, weights_only=False, map_location="cpu")
    assert [n for n, _ in loaded.named_parameters()] == ["x2i", "h2i", "x2f", "h2f", "x2g", "h2g", "x2o", "h2o", "out"]
    assert type(loaded).__module__ == "blocks.lstm" or type(loaded).__name__ == "LSTM"
    xt = torch.from_numpy(x)
    with torch.no_grad():
        np.testing.assert_allclose(loaded(xt).numpy(), model(xt.cuda()).cpu().numpy(), rtol=1e-5, atol=1e-6)


def test_full_size_properties():
    """At a bench-sized shard the oracle is too slow; check size-independent properties instead:
    (1) ghost rows never leak: growing N by ghost-only padding changes nothing; (2) permuting the
    samples permutes the state and leaves the weights unchanged up to reduction order; (3) slot t=0 stays
    zero and lambda_h stays zero for t < T (SURVEY 8(a) invariants)."""
    _need_gpu()
    from gpu_utils import make_opt, weights_of
    n, t, d, h, o = 20000, 8, 16, 64, 1
    x, y, w = synthetic_problem(n, t, d, h, o, seed=2)
    _, a = make_opt(w, x, y, GOOGLE, "admm")
    perm = np.random.default_rng(0).permutation(n)
    _, b = make_opt(w, x[perm], y[perm], GOOGLE, "admm")
    for _ in range(2):
        a.step()
        b.step()
    wa, wb = weights_of(a), weights_of(b)
    for k in WKEYS:
        assert rel_err(wa[k], wb[k]) < 1e-5, (k, rel_err(wa[k], wb[k]))
    ha = a.gates["h"].cpu().numpy()
    hb = b.gates["h"].cpu().numpy()
    assert rel_err(hb, ha[perm]) < 1e-5
    for k in ("i", "f", "g", "o", "c", "h"):
        assert float(a.gates[k][:, 0, :].abs().max()) == 0.0
    assert float(a.duals["h"][:, :t, :].abs().max()) == 0.0
    assert all(np.isfinite(v).all() for v in wa.values())


@pytest.mark.parametrize("shape", [(256, 3, 16, 64, 1), (300, 2, 9, 128, 2), (1000, 2, 64, 256, 1)])
def test_tensor_core_preactivations(shape):
    """tcgen05 3xTF32 gate GEMM vs the fp32 CUDA-core GEMM and vs float64: the split product must be
    fp32-accurate (plain TF32 would be ~1e-3 off)."""
    _need_gpu()
    from gpu_utils import make_opt
    n, t, d, h, o = shape
    x, y, w = synthetic_problem(n, t, d, h, o, seed=h)
    _, opt = make_opt(w, x, y, GOOGLE, "admm", use_tensor_cores=True)
    assert opt.uses_tensor_cores
    s = torch.cuda.current_stream().cuda_stream
    hs = opt.gates["h"].cpu().numpy().astype(np.float64)
    for tt in range(1, t + 1):
        z_tc = torch.zeros((4, h, opt.ldn), device="cuda")
        z_cc = torch.zeros((4, h, opt.ldn), device="cuda")
        opt._call("admm_debug_preact", opt._pp, tt, z_tc.data_ptr(), 1, s)
        opt._call("admm_debug_preact", opt._pp, tt, z_cc.data_ptr(), 0, s)
        torch.cuda.synchronize()
        ref = np.stack([x[:, tt - 1, :].astype(np.float64) @ w["x2" + g].astype(np.float64)
                        + hs[:, tt - 1, :] @ w["h2" + g].astype(np.float64) for g in "ifgo"])      # [4][n][h]
        ref = ref.transpose(0, 2, 1)
        e_tc = np.max(np.abs(z_tc[:, :, :n].cpu().numpy() - ref)) / np.max(np.abs(ref))
        e_cc = np.max(np.abs(z_cc[:, :, :n].cpu().numpy() - ref)) / np.max(np.abs(ref))
        print(f"shape {shape} t={tt}: rel err vs f64  tcgen05 3xTF32 {e_tc:.2e}   fp32 FMA {e_cc:.2e}")
        assert e_cc < 2e-6
        assert e_tc < 4e-6      # ~2^-22 per product from the dropped lo*lo term and the tf32 rounding of the lo parts


@pytest.mark.parametrize("variant,shape,params", [("admm", (384, 4, 16, 64, 1), GOOGLE), ("no_dual_y", (500, 3, 9, 128, 6), HAR)])
def test_tensor_core_step_vs_oracle(variant, shape, params):
    _need_gpu()
    from oracle.admm_oracle import OracleADMM
    from gpu_utils import make_opt
    n, t, d, h, o = shape
    x, y, w = synthetic_problem(n, t, d, h, o, seed=11, classification=(o > 1))
    _, opt = make_opt(w, x, y, params, variant, use_tensor_cores=True)
    ora = OracleADMM(w, x, y, params, variant=variant)
    _cmp_state(opt, ora, params, "init", rel=1e-5)
    for s in range(3):
        ora.step()
        opt.step()
        _cmp_state(opt, ora, params, f"tc step{s}")


def test_stored_preactivations_equal_recomputed():
    """The x-phase gradient from the z the previous sweep stored (admm_problem::z_valid, grad_from_z.cu) must give
    the same iterates as recomputing z with the gate GEMM: same kernel main loop, same inputs -> same bits."""
    _need_gpu()
    from gpu_utils import make_opt, weights_of
    n, t, d, h, o = 700, 5, 16, 128, 1
    x, y, w = synthetic_problem(n, t, d, h, o, seed=5)
    _, a = make_opt(w, x, y, GOOGLE, "admm", use_tensor_cores=True)
    _, b = make_opt(w, x, y, GOOGLE, "admm", use_tensor_cores=True)
    assert a.keeps_preactivations and a._p.z_valid == 1
    for _ in range(4):
        a.step()
        assert a._p.z_valid == 1
        b._set_z_valid(False)          # force the GEMM path
        b.step()
    wa, wb = weights_of(a), weights_of(b)
    for k in WKEYS:
        assert np.array_equal(wa[k], wb[k]), k
    for k in ("i", "f", "g", "o", "c", "h"):
        assert torch.equal(a.gates[k], b.gates[k]), k
    assert a.theta_trace() == b.theta_trace()


@pytest.mark.parametrize("shape,tc", [((150, 4, 3, 12, 2), False), ((300, 3, 16, 64, 1), True)])
def test_checkpoint_resume_is_bit_identical(tmp_path, shape, tc):
    """optimizer state saved next to the model file; a resumed run continues with the same iterates."""
    _need_gpu()
    from gpu_utils import make_opt, weights_of
    n, t, d, h, o = shape
    x, y, w = synthetic_problem(n, t, d, h, o, seed=9)
    model, a = make_opt(w, x, y, GOOGLE, "admm", use_tensor_cores=tc)
    for _ in range(3):
        a.step()
    torch.save(model, tmp_path / "model.pt")
    fn = a.save_state(str(tmp_path / "model.pt.admm"))
    assert os.path.exists(fn)
    for _ in range(3):
        a.step()
    model_b = torch.load(tmp_path / "model.pt", weights_only=False)
    from admm_lstm_b200.optimizer import ADMMBasedOptimizer
    b = ADMMBasedOptimizer(model_b, (torch.from_numpy(x), torch.from_numpy(y)), GOOGLE, verbose=False, variant="admm",
                           use_tensor_cores=tc)
    b.load_state(str(tmp_path / "model.pt.admm"))
    for _ in range(3):
        b.step()
    wa, wb = weights_of(a), weights_of(b)
    for k in WKEYS:
        assert np.array_equal(wa[k], wb[k]), k
    for k in ("i", "f", "g", "o", "c", "h", "a"):
        assert torch.equal(a.gates[k], b.gates[k]), k
    for k in ("i", "f", "g", "o", "c", "h", "y"):
        assert torch.equal(a.duals[k], b.duals[k]), k


@pytest.mark.parametrize("shape,tc", [((1500, 6, 16, 128, 1), True), ((777, 5, 3, 20, 2), False)])
def test_moment_probe_equals_exact_probe(shape, tc):
    """The one-pass moment evaluation of the backtracking candidates (admm_probe_plan::moments) against the
    candidate-by-candidate evaluation: same iterates to 1e-5 over 8 iterations.  The chosen thetas may differ only where the
    comparison f(beta) > est is below fp32 resolution on both sides (absorption exits): then the step G/theta is
    negligible either way, which is what the weight tolerance checks."""
    _need_gpu()
    from gpu_utils import make_opt, weights_of
    n, t, d, h, o = shape
    x, y, w = synthetic_problem(n, t, d, h, o, seed=13)
    _, a = make_opt(w, x, y, GOOGLE, "admm", use_tensor_cores=tc, probe="moments")
    _, b = make_opt(w, x, y, GOOGLE, "admm", use_tensor_cores=tc, probe="exact")
    same = total = 0
    for it in range(8):
        a.step()
        b.step()
        ta, tb = a.theta_trace(), b.theta_trace()
        for k in ta:
            total += 1
            same += abs(ta[k] - tb[k]) <= 1e-6 * tb[k]
        wa, wb = weights_of(a), weights_of(b)
        for k in WKEYS:
            assert rel_err(wa[k], wb[k]) < 1e-5, (it, k, rel_err(wa[k], wb[k]), ta, tb)
    for k in ("i", "f", "g", "o", "c", "h"):
        assert rel_err(a.gates[k].cpu().numpy(), b.gates[k].cpu().numpy()) < 1e-5, k
    print(f"moment vs exact probes: {same}/{total} identical thetas")


@pytest.mark.parametrize("shape,tc", [((600, 4, 16, 64, 1), True), ((333, 3, 5, 12, 2), False)])
def test_large_probe_steps_take_the_exact_passes(shape, tc):
    """A state far from consistency (random gates and duals) makes the gradient, hence Q = A G, large: at the first
    exponents the perturbation of the pre-activations is far outside the validity of the moment expansion.  The device
    must notice (diagnostic code 3 / failed bounds), fall back to the exact candidate-by-candidate passes and still
    reproduce the oracle's thetas and weights."""
    _need_gpu()
    from oracle.admm_oracle import OracleADMM
    from gpu_utils import make_opt, load_state, weights_of
    from admm_lstm_b200 import _lib
    n, t, d, h, o = shape
    x, y, w = synthetic_problem(n, t, d, h, o, seed=17)
    rng = np.random.default_rng(5)
    st = {"gates": {}, "duals": {}}
    for k in ("i", "f", "g", "o", "c", "h"):
        st["gates"][k] = rng.uniform(-1, 1, (n, t + 1, h)).astype(np.float32)
        st["gates"][k][:, 0] = 0
        st["duals"][k] = (0.5 * rng.standard_normal((n, t + 1, h))).astype(np.float32)
        st["duals"][k][:, 0] = 0
    st["duals"]["h"][:, :t] = 0
    st["gates"]["a"] = rng.random((n, o)).astype(np.float32)
    st["duals"]["y"] = np.zeros((n, o), np.float32)
    _, opt = make_opt(w, x, y, GOOGLE, "admm", use_tensor_cores=tc)
    load_state(opt, st)
    ora = OracleADMM(w, x, y, GOOGLE, variant="admm", state=st)
    s = torch.cuda.current_stream().cuda_stream
    fell_back = False
    for src, name in ((_lib.SRC_X, "x"), (_lib.SRC_H, "h")):
        opt._ADMMBasedOptimizer__update_weights(src, s)
        fell_back |= any(v != 0 for v in opt._done.cpu().tolist()[4:8])
        for g in "ifgo":
            ora.update_weights(name, g)
    assert fell_back, "the test state was meant to push the moment pass out of its validity range"
    th = opt.theta_trace()
    for k, v in ora.trace.items():
        if k in th:
            assert abs(th[k] - v) <= 1e-6 * v, (k, th[k], v)
    # with this state the new weights are ~ -G, a sum of N*T*H terms of both signs: the fp32 CUDA-core path is itself
    # 5e-5 from the oracle here (summation order), the split-precision tensor-core path 1.4e-4
    wg = weights_of(opt)
    for k in WKEYS:
        if k != "out":
            assert rel_err(wg[k], ora.w[k]) < 5e-4, (k, rel_err(wg[k], ora.w[k]))


def test_inplace_weight_write_is_noticed():
    """model.load_state_dict(sd) / p.data.mul_() write INTO the bound Parameters (same data_ptr): the fp16-pair
    weight operands and the stored pre-activations of the tensor-core path are derived data and must follow.  Reference
    behaviour: set_weight / setattr simply replace the tensor (blocks/lstm.py:31-41), no hazard there."""
    _need_gpu()
    from gpu_utils import make_opt, weights_of
    n, t, d, h, o = 300, 4, 16, 64, 1
    x, y, w = synthetic_problem(n, t, d, h, o, seed=23)
    ma, a = make_opt(w, x, y, GOOGLE, "admm", use_tensor_cores=True)
    mb, b = make_opt(w, x, y, GOOGLE, "admm", use_tensor_cores=True)
    assert a.keeps_preactivations
    a.step()
    b.step()
    sd = {k: v.detach().clone() * 0.9 for k, v in ma.state_dict().items()}
    ma.load_state_dict(sd)                       # nn.Module.load_state_dict -> param.copy_(): in place, data_ptr unchanged
    with torch.no_grad():
        for k, v in mb.state_dict().items():
            v.mul_(0.9)
    b.state_changed()                            # the explicit route (tests use it after writing buffers directly)
    for _ in range(2):
        a.step()
        b.step()
    wa, wb = weights_of(a), weights_of(b)
    for k in WKEYS:
        assert np.array_equal(wa[k], wb[k]), k
    # and against the oracle restarted from the same modified weights and state is covered by b's route (test_*_vs_oracle)


def test_forward_takes_the_library_kernel_with_grad_enabled():
    """SURVEY 8 f2 on the path the reference actually calls: demo.py:341-342 evaluates model(train_x) with autograd ENABLED.
    With the parameters owned by an ADMM optimizer (requires_grad False) that call must run admm_predict, not eager torch."""
    _need_gpu()
    from gpu_utils import make_opt
    from admm_lstm_b200 import _lib
    from admm_lstm_b200.optimizer import sharded_mse_loss
    n, t, d, h, o = 700, 6, 5, 24, 2
    x, y, w = synthetic_problem(n, t, d, h, o, seed=31)
    model, opt = make_opt(w, x, y, GOOGLE, "admm")
    opt.step()
    xc, yc = torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()
    lib = _lib.load()
    assert torch.is_grad_enabled()
    before = lib.admm_launch_count(0)
    pred = model(xc)                                             # grad enabled, as demo.py calls it
    assert lib.admm_launch_count(0) - before >= t, "model(x) did not go through the library's forward kernel"
    ref = model.init_gate_variables(xc)["a"]                     # eager torch (blocks/lstm.py:65-88)
    np.testing.assert_allclose(pred.cpu().numpy(), ref.cpu().numpy(), rtol=2e-5, atol=2e-6)
    loss = float(torch.nn.functional.mse_loss(pred, yc))
    assert abs(opt.training_loss() - loss) < 1e-5 * loss         # resident-data evaluation, same number
    assert abs(sharded_mse_loss(model, xc, yc) - loss) < 1e-5 * loss
    # chunked over samples: results independent of the chunk size
    import admm_lstm_b200.optimizer as om
    old = om._PREDICT_CHUNK
    try:
        om._PREDICT_CHUNK = 256
        np.testing.assert_array_equal(model(xc).cpu().numpy(), pred.cpu().numpy())
    finally:
        om._PREDICT_CHUNK = old
    # a model whose parameters DO require grad keeps the differentiable eager path
    from admm_lstm_b200.lstm import LSTM
    m2 = LSTM(d, h, o).cuda()
    out = m2(xc)
    assert out.requires_grad


@pytest.mark.parametrize("shape,tc", [((4224, 10, 1, 10, 1), False), ((384, 4, 16, 64, 1), True)])
def test_cuda_graph_replay_equals_eager(shape, tc):
    """Launch-bound shapes replay the step (admm.py:62-78: Wy, eight weights, T sweep launches, tail) as one CUDA graph; the
    iterates must be those of the eager launches, also across a change of the probe plan (re-capture) and after the state
    was written from outside (z_valid in the graph key)."""
    _need_gpu()
    from gpu_utils import make_opt, weights_of
    n, t, d, h, o = shape
    x, y, w = synthetic_problem(n, t, d, h, o, seed=41)
    _, a = make_opt(w, x, y, GOOGLE, "admm", use_tensor_cores=tc, use_cuda_graph=True)
    _, b = make_opt(w, x, y, GOOGLE, "admm", use_tensor_cores=tc, use_cuda_graph=False)
    for it in range(10):
        a.step()
        b.step()
        if it == 6:
            a.state_changed()
            b.state_changed()
    assert a.graph_replays >= 6 and b.graph_replays == 0
    assert a.graph_replayed_launches > 0
    wa, wb = weights_of(a), weights_of(b)
    for k in WKEYS:
        assert rel_err(wa[k], wb[k]) < 1e-6, (k, rel_err(wa[k], wb[k]))
    for k in ("i", "f", "g", "o", "c", "h"):
        assert rel_err(a.gates[k].cpu().numpy(), b.gates[k].cpu().numpy()) < 1e-6, k
    assert a.theta_trace() == b.theta_trace()
    ma, mb = a.metrics(), b.metrics()
    for k in ma:
        assert abs(ma[k] - mb[k]) <= 1e-6 * abs(mb[k]) + 1e-12, k


def test_h_operand_overflow_is_flagged():
    """The fp16-pair operand of h on the tensor-core path holds |h| < 32.  h_t = o tanh(c) - lambda_h/rho_h with o an
    unconstrained ADMM primal (admm.py:373-386, 455-457): a larger value must raise the sticky device flag instead of
    silently clamping."""
    _need_gpu()
    from gpu_utils import make_opt
    n, t, d, h, o = 256, 3, 16, 64, 1
    x, y, w = synthetic_problem(n, t, d, h, o, seed=3)
    _, opt = make_opt(w, x, y, GOOGLE, "admm", use_tensor_cores=True)
    opt.step()
    assert not opt.operand_overflow()
    # (1) the sweep writes it: lambda_o = -100 makes o ~ 100 (gate_prox), h = o tanh(c) far beyond 32
    opt._dual["o"][1, :, : n] = -100.0
    opt.state_changed()
    assert not opt.operand_overflow()
    opt.step()
    assert float(opt.gates["h"].abs().max()) > 32.0
    assert opt.operand_overflow(reset=True)
    assert not opt.operand_overflow()
    # (2) the state refresh sees it: h written from outside
    opt._state["h"][1, :, : n] = 50.0
    opt.state_changed()
    assert opt.operand_overflow()
    # the CUDA-core path has no such limit
    _, ref = make_opt(w, x, y, GOOGLE, "admm", use_tensor_cores=False)
    assert not ref.operand_overflow()


def test_prefetched_inputs_equal_direct_refresh():
    """opt.prefetch_inputs() (side-stream upload into a staging buffer) + refresh_inputs() installs the same inputs as
    refresh_inputs(x, y): the e2e path of bench.py double-buffers its uploads this way."""
    _need_gpu()
    from gpu_utils import make_opt, weights_of
    n, t, d, h, o = 300, 4, 16, 64, 1
    x, y, w = synthetic_problem(n, t, d, h, o, seed=51)
    x2, y2, _ = synthetic_problem(n, t, d, h, o, seed=52)
    _, a = make_opt(w, x, y, GOOGLE, "admm", use_tensor_cores=True)
    _, b = make_opt(w, x, y, GOOGLE, "admm", use_tensor_cores=True)
    xp, yp = torch.from_numpy(x2).pin_memory(), torch.from_numpy(y2).pin_memory()
    a.step(); b.step()
    a.prefetch_inputs(xp, yp)
    a.step(); b.step()
    a.refresh_inputs()
    b.refresh_inputs(xp, yp)
    for _ in range(2):
        a.step(); b.step()
    wa, wb = weights_of(a), weights_of(b)
    for k in WKEYS:
        assert np.array_equal(wa[k], wb[k]), k


def test_reinstalled_inputs_keep_or_drop_the_stored_preactivations():
    """admm_load_inputs decides ON THE DEVICE whether the stored pre-activations survive new inputs: identical values ->
    the run continues bit-identically to one that never re-installed them (streaming x-phase gradient); different values ->
    the x-phase falls back to its GEMM pass and the iterates equal those of an optimizer that keeps no z store at all."""
    _need_gpu()
    from gpu_utils import make_opt, weights_of
    n, t, d, h, o = 300, 4, 16, 64, 1
    x, y, w = synthetic_problem(n, t, d, h, o, seed=61)
    x2, y2, _ = synthetic_problem(n, t, d, h, o, seed=62)
    xt, yt = torch.from_numpy(x), torch.from_numpy(y)
    _, a = make_opt(w, x, y, GOOGLE, "admm", use_tensor_cores=True)      # re-installs the same inputs every step
    _, b = make_opt(w, x, y, GOOGLE, "admm", use_tensor_cores=True)      # never touches them
    for _ in range(3):
        a.refresh_inputs(xt, yt)
        assert a._p.z_valid == 1
        a.step()
        b.step()
    wa, wb = weights_of(a), weights_of(b)
    for k in WKEYS:
        assert np.array_equal(wa[k], wb[k]), k
    assert int(a._done.cpu()[0]) == 1
    # different inputs: against an optimizer without z store (every pass recomputes its pre-activations)
    _, c = make_opt(w, x, y, GOOGLE, "admm", use_tensor_cores=True, keep_preactivations=False)
    for _ in range(3):
        c.step()
    x2t, y2t = torch.from_numpy(x2), torch.from_numpy(y2)
    a.refresh_inputs(x2t, y2t)
    c.refresh_inputs(x2t, y2t)
    # the transposing install itself: device layout [T][D][ldn] / [O][ldn], ghost columns untouched (zero)
    assert torch.equal(a._x[:, :, :n].cpu(), x2t.permute(1, 2, 0)) and torch.equal(a._y[:, :n].cpu(), y2t.t())
    assert float(a._x[:, :, n:].abs().max()) == 0.0 and float(a._y[:, n:].abs().max()) == 0.0
    for _ in range(2):
        a.step()
        c.step()
    wa, wc = weights_of(a), weights_of(c)
    # (c runs the unfused probe passes and full-GEMM gradient passes: other summation orders, hence 2e-5 and not bit equality)
    for k in WKEYS:
        assert rel_err(wa[k], wc[k]) < 2e-5, (k, rel_err(wa[k], wc[k]))
    assert rel_err(a.gates["h"].cpu().numpy(), c.gates["h"].cpu().numpy()) < 2e-5
    assert a.theta_trace() == c.theta_trace()
