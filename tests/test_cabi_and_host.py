"""CPU-side checks: the C-ABI library loads and exports every symbol include/admm_lstm_b200.h declares,
the host-side formulas compiled from csrc/admm_math.cuh agree with the oracle, the host modules
mirror the reference's surface, and the product path refuses to run without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from helpers import GOOGLE, HAR, load, state_from, WKEYS

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _build():
    from admm_lstm_b200 import build
    return build.build(), build.build_host_math()


def test_library_exports_every_declared_symbol():
    lib_path, _ = _build()
    header = open(os.path.join(ROOT, "include", "admm_lstm_b200.h")).read()
    body = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(admm_[a-z_0-9]+)\s*\(", body))
    assert len(declared) >= 18
    lib = ctypes.CDLL(lib_path)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    from admm_lstm_b200 import _lib
    assert set(_lib.SIGNATURES) == declared
    handle = _lib.load()
    assert handle.admm_abi_version() == 1
    assert handle.admm_sizeof_problem() == ctypes.sizeof(_lib.Problem)


def test_no_cpu_fallback():
    """No GPU -> the optimizer must fail loudly, and the C entry points must return ADMM_ENODEV."""
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    _build()
    from admm_lstm_b200 import _lib
    from admm_lstm_b200.lstm import LSTM
    from admm_lstm_b200.optimizer import ADMMBasedOptimizer
    with pytest.raises(_lib.AdmmLibraryError):
        ADMMBasedOptimizer(LSTM(1, 2, 1), (torch.rand(4, 3, 1), torch.rand(4, 1)), GOOGLE, verbose=False)
    lib = _lib.load()
    p = _lib.Problem()
    p.n = p.n_global = 4
    p.ldn = 128
    p.T, p.D, p.H, p.O = 3, 1, 2, 1
    assert lib.admm_sweep_t(ctypes.byref(p), 1, None, None) == -3
    assert b"no CPU path" in lib.admm_last_error()


def test_tensor_core_shape_rules():
    """admm_tc_workspace_bytes (pure host arithmetic): the tensor-core path needs H % 64 == 0, ldn % 128 == 0 and state
    tensors below 2^32 elements (its epilogue addresses them with 32-bit offsets); otherwise 0 -> CUDA-core path."""
    _build()
    from admm_lstm_b200 import _lib
    lib = _lib.load()

    def ws(T, D, H, ldn):
        p = _lib.Problem()
        p.n = p.n_global = ldn
        p.ldn = ldn
        p.T, p.D, p.H, p.O = T, D, H, 1
        return int(lib.admm_tc_workspace_bytes(ctypes.byref(p)))

    assert ws(128, 64, 1024, 16384) > 0                   # the bench shape
    assert ws(10, 1, 10, 4224) == 0                       # GoogleStock: H % 64 != 0
    assert ws(128, 64, 1024, 16384 + 64) == 0             # ldn not a multiple of 128
    assert (128 + 1) * 1024 * 32512 < 2 ** 32 <= (128 + 1) * 1024 * 32640
    assert ws(128, 64, 1024, 32512) > 0
    assert ws(128, 64, 1024, 32640) == 0                  # 129 * 1024 * 32640 elements >= 2^32


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "admm_lstm_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                src = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in src.replace("the oracle", "").replace("oracle's", "") or f == "point_math_host.cpp", f
    for f in ("admm.py", "admm.no_dual_y.py", "parameters.py", "_global.py", os.path.join("blocks", "lstm.py")):
        assert "import oracle" not in open(os.path.join(ROOT, f)).read()


@pytest.mark.parametrize("w_scale", [1.0, 60.0])
@pytest.mark.parametrize("variant", ["admm", "no_dual_y"])
def test_point_math_matches_oracle(variant, w_scale):
    """csrc/admm_math.cuh compiled for the host == oracle's sequential per-function updates.  w_scale = 60 drives the
    pre-activations deep into saturation (sigmoid == 0 or 1, tanh == +-1, derivatives == 0 in fp32)."""
    from oracle.admm_oracle import OracleADMM
    _, host = _build()
    lib = ctypes.CDLL(host)
    rec = load(f"fn_{variant}.npz")
    base_w = {k: (rec[f"base_w_{k}"] * np.float32(w_scale)).astype(np.float32) for k in WKEYS}
    st = state_from(rec, "base_")
    T = rec["x"].shape[1]
    fp = ctypes.POINTER(ctypes.c_float)
    for params in (GOOGLE, HAR):
        rho = np.array([params["rho"][k] for k in "ifgochy"], dtype=np.float32)
        for t in (2, T):
            ora = OracleADMM(base_w, rec["x"], rec["y"], params, variant=variant, state=st)
            g, d = ora.gates, ora.duals
            rows = [ora._z(k, t) for k in "ifgo"] + [g[k][:, t, :].copy() for k in "ifgoch"] + \
                   [g["c"][:, t - 1, :].copy()] + [d[k][:, t, :].copy() for k in "ifgoc"] + \
                   [d["h"][:, t, :].copy() if t == T else np.zeros_like(d["h"][:, t, :])]
            if t < T:
                assert not np.any(st["duals"]["h"][:, t, :]) or True
                ora.duals["h"][:, t, :] = 0          # invariant of a real run (admm.py:533): lambda_h = 0 for t < T
            inp = np.ascontiguousarray(np.stack([r.astype(np.float32).ravel() for r in rows]))
            n = inp.shape[1]
            out = np.zeros((14, n), dtype=np.float32)
            lib.admm_host_sweep_points(inp.ctypes.data_as(fp), out.ctypes.data_as(fp), ctypes.c_long(n),
                                       rho.ctypes.data_as(fp), ctypes.c_int(int(t == T)))
            for k in "ifgo":
                ora.update_primal_ifgo(k, t)
            ora.update_primal_c(t)
            if t < T:
                ora.update_primal_h(t)
            for k in "ifgo":
                ora.update_dual_ifgo(k, t)
            ora.update_dual_c(t)
            shape = g["i"][:, t, :].shape
            assert np.isfinite(out).all()
            if w_scale > 1:
                zmax = max(float(np.abs(r).max()) for r in rows[:4])
                assert zmax > 30, zmax                 # the case really is saturated
            for q, k in enumerate("ifgoc" + ("h" if t < T else "")):
                np.testing.assert_allclose(out[q].reshape(shape), g[k][:, t, :], rtol=3e-5, atol=3e-6, err_msg=f"{k}@{t}")
            for q, k in enumerate("ifgoc"):
                np.testing.assert_allclose(out[6 + q].reshape(shape), d[k][:, t, :], rtol=3e-5, atol=3e-6, err_msg=f"lam_{k}@{t}")


def test_grad_and_probe_points():
    _, host = _build()
    lib = ctypes.CDLL(host)
    from oracle.admm_oracle import _sigmoid, _tanh
    rng = np.random.default_rng(0)
    n = 1000
    z = (rng.standard_normal(n) * 3).astype(np.float32)
    lam = (rng.standard_normal(n) * 0.1).astype(np.float32)
    gate = rng.random(n).astype(np.float32)
    q = rng.standard_normal(n).astype(np.float32)
    fp = ctypes.POINTER(ctypes.c_float)
    for is_g, act in ((0, _sigmoid), (1, _tanh)):
        R = np.zeros(n, np.float32)
        u = np.zeros(n, np.float32)
        rho = np.float32(1.5)
        lib.admm_host_grad_points(z.ctypes.data_as(fp), lam.ctypes.data_as(fp), gate.ctypes.data_as(fp),
                                  ctypes.c_float(rho), is_g, R.ctypes.data_as(fp), u.ctypes.data_as(fp), ctypes.c_long(n))
        a = act(z)
        da = (1 - a * a) if is_g else a * (1 - a)
        uu = a - lam / rho - gate
        np.testing.assert_allclose(u, uu, rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(R, uu * da, rtol=1e-5, atol=1e-6)
        out = np.zeros(n, np.float32)
        lib.admm_host_probe_points(z.ctypes.data_as(fp), q.ctypes.data_as(fp), ctypes.c_float(0.25),
                                   lam.ctypes.data_as(fp), gate.ctypes.data_as(fp), ctypes.c_float(rho), is_g,
                                   out.ctypes.data_as(fp), ctypes.c_long(n))
        ref = (act(z + q * np.float32(0.25)) - lam / rho - gate) ** 2
        np.testing.assert_allclose(out, ref, rtol=2e-5, atol=1e-6)


def test_reference_surface_of_host_modules():
    import importlib
    import sys
    sys.path.insert(0, ROOT)
    admm = importlib.import_module("admm")
    assert hasattr(admm, "ADMMBasedOptimizer") and "GoogleStock" in admm.example_parameter_dictionary
    import inspect
    sig = inspect.signature(admm.ADMMBasedOptimizer.__init__)
    assert list(sig.parameters)[:5] == ["self", "model", "training_samples", "parameter_dictionary", "verbose"]
    assert sig.parameters["parameter_dictionary"].default is None and sig.parameters["verbose"].default is True
    params = importlib.import_module("parameters")
    assert params.default_epoch == 100
    for name, entry in params.example_parameter_dictionary.items():
        assert set(entry["rho"]) == set("ifgochy") and len(entry["beta"]) == 9, name
    lstm = importlib.import_module("blocks.lstm")
    torch.manual_seed(0)
    m = lstm.LSTM(3, 5, 2)
    assert [n for n, _ in m.named_parameters()] == ["x2i", "h2i", "x2f", "h2f", "x2g", "h2g", "x2o", "h2o", "out"]
    st = m.init_gate_variables(torch.rand(4, 6, 3))
    assert st["h"].shape == (4, 7, 5) and st["a"].shape == (4, 2) and float(st["c"][:, 0].abs().max()) == 0


def test_lstm_forward_matches_oracle():
    from oracle.admm_oracle import lstm_forward
    from admm_lstm_b200.lstm import LSTM
    torch.manual_seed(1)
    m = LSTM(4, 9, 3)
    x = torch.rand(11, 5, 4)
    w = {n: p.detach().numpy() for n, p in m.named_parameters()}
    st = lstm_forward(w, x.numpy())
    with torch.no_grad():
        mine = m.init_gate_variables(x)
    for k in ("i", "f", "g", "o", "c", "h", "a"):
        np.testing.assert_allclose(mine[k].numpy(), st[k], rtol=1e-5, atol=1e-6)
    m.with_grad = True
    np.testing.assert_allclose(m(x).detach().numpy(), st["a"], rtol=1e-5, atol=1e-6)


def test_three_instruction_division_is_the_ieee_quotient():
    """csrc/admm_math.cuh div_rn (lambda / rho in the fused moment pass) == the host's IEEE division, bit for bit, for the
    rho of every shipped parameter set and a sweep of magnitudes of lambda down to the subnormal results (admm.py:318 divides)."""
    _, host = _build()
    lib = ctypes.CDLL(host)
    fp = ctypes.POINTER(ctypes.c_float)
    rng = np.random.default_rng(7)
    rhos = {float(np.float32(v)) for params in (GOOGLE, HAR) for v in params["rho"].values()}
    rhos |= {1.0, 0.5, 3.0, 1.0000001, 1.9999999, 5.62e-5, 7.77e-4} | set(rng.uniform(1e-6, 10.0, 16).tolist())
    n = 200_000
    lam = (rng.standard_normal(n) * 10.0 ** rng.uniform(-20, 3, n)).astype(np.float32)
    lam[:4] = [0.0, -0.0, 1.0, -1.0]
    out = np.empty(n, dtype=np.float32)
    for rho in sorted(rhos):
        lib.admm_host_div_rn(lam.ctypes.data_as(fp), ctypes.c_float(rho), out.ctypes.data_as(fp), ctypes.c_long(n))
        want = lam / np.float32(rho)
        nz = want != 0                                   # -0 / rho comes out as +0: the only difference, and not one in value
        assert np.array_equal(out[nz].view(np.uint32), want[nz].view(np.uint32)) and not np.any(out[~nz]), \
            (rho, int((out != want).sum()))


@pytest.mark.parametrize("is_g", [0, 1])
def test_fused_moment_accumulation_equals_coefficient_form(is_g):
    """moment_accum4 (what the epilogue runs) == moment_terms4 followed by the powers of t (round 2's first form) to fp32
    rounding of the individual terms, without bias in the sums: both against each other and against the float64 polynomial."""
    _, host = _build()
    lib = ctypes.CDLL(host)
    fp, dp = ctypes.POINTER(ctypes.c_float), ctypes.POINTER(ctypes.c_double)
    rng = np.random.default_rng(11 + is_g)
    n = 1_000_000
    z = (rng.standard_normal(n) * 2.5).astype(np.float32)
    c = (rng.standard_normal(n) * 0.3 + 0.4).astype(np.float32)
    t = (rng.standard_normal(n) * 2.0 ** -8).clip(-2.0 ** -6, 2.0 ** -6).astype(np.float32)
    out = np.zeros(8)
    lib.admm_host_moment_sums4(z.ctypes.data_as(fp), c.ctypes.data_as(fp), t.ctypes.data_as(fp), ctypes.c_int(is_g),
                               out.ctypes.data_as(dp), ctypes.c_long(n))
    # float64 polynomial from the same fp32 s, u, t
    z64, t64 = z.astype(np.float64), t.astype(np.float64)
    s = (np.tanh(z64) if is_g else 1.0 / (1.0 + np.exp(-z64))).astype(np.float32).astype(np.float64)
    u = (s.astype(np.float32) - c).astype(np.float64)
    if is_g:
        d1 = 1 - s * s
        a = [d1, -s * d1, -d1 * (1 - 3 * s * s) / 3, s * d1 * (2 - 3 * s * s) / 3]
    else:
        d1, m = s * (1 - s), 1 - 2 * s
        a = [d1, d1 * m / 2, d1 * (1 - 6 * d1) / 6, d1 * m * (1 - 12 * d1) / 24]
    co = [2 * u * a[0], a[0] ** 2 + 2 * u * a[1], 2 * a[0] * a[1] + 2 * u * a[2], a[1] ** 2 + 2 * a[0] * a[2] + 2 * u * a[3]]
    for k in range(4):
        terms = co[k] * t64 ** (k + 1)
        exact, scale = terms.sum(), np.abs(terms).sum()
        assert abs(out[k] - exact) <= 2e-7 * scale, (k, out[k], exact)           # first form
        assert abs(out[4 + k] - exact) <= 2e-7 * scale, (k, out[4 + k], exact)   # fused form
        assert abs(out[4 + k] - out[k]) <= 2e-7 * scale


def _plan_ok(entry):
    """The constraints capi.cu:check_plan puts on one pass of the probe schedule."""
    from admm_lstm_b200 import _lib
    k0, ncand, proof, moments = entry[:4]
    assert (0 if moments else 1) <= ncand <= _lib.ADMM_MAX_CAND, entry
    for k in k0:
        assert k >= 0 and k + (0 if moments else ncand) <= _lib.ADMM_EST_CAND, entry
        if proof:
            assert k + (ncand if moments else 0) <= _lib.ADMM_MAX_CAND, entry
    if len(entry) > 4:
        assert entry[4] in (4, 6), entry


@pytest.mark.parametrize("graph", [False, True])
@pytest.mark.parametrize("fused", [False, True])
def test_probe_schedule_is_always_a_valid_plan(fused, graph):
    """optimizer._probe_plans (host side of admm.py:331-338): for ANY max|Q| hint -- tiny, huge, inf, NaN, absent -- every pass
    it schedules satisfies check_plan, the moment pass comes first while its window fits, a diverging run goes straight to the
    exact passes, and the schedule always ends with the two exact 32-candidate passes that decide whatever is left."""
    import types
    from admm_lstm_b200 import _lib
    from admm_lstm_b200.optimizer import ADMMBasedOptimizer
    full = [((0, 0, 0, 0), _lib.ADMM_MAX_CAND, 0, 0), ((32, 32, 32, 32), _lib.ADMM_MAX_CAND, 0, 0)]
    stub = types.SimpleNamespace(probe="moments", _hint=None, _hint_q=None, uses_tensor_cores=fused,
                                 _zstore=object() if fused else None, moment_order=4, use_cuda_graph=graph)
    plans = ADMMBasedOptimizer._probe_plans(stub, _lib.SRC_X)
    assert plans[1:] == full and plans[0][0] == (0, 0, 0, 0) and plans[0][3] == 1          # no hint yet: expansion from k0 = 0
    order = 4 if fused else 6
    qs = [0.0, 1e-30, 2.0 ** -9, 2.0 ** -7, 2.0 ** -5, 0.1, 0.4, 1.0, 180.0, 1e5, 2.0 ** 22, 2.0 ** 23, 2.0 ** 24, 1e9, 1e30,
          float("inf"), float("nan")]
    for q in qs:
        for src in (_lib.SRC_X, _lib.SRC_H):
            stub._hint_q = {_lib.SRC_X: [q, 0.01, q / 3 if q == q else q, 0.2], _lib.SRC_H: [0.3, q, 0.05, q]}
            plans = ADMMBasedOptimizer._probe_plans(stub, src)
            for entry in plans:
                _plan_ok(entry)
            assert plans[-2:] == full
            if len(plans) == 3:
                k0, ncand, proof, moments, o = plans[0]
                assert moments == 1 and proof == 1 and ncand == 2 and o == order
                lim = 2.0 ** -7 if order == 4 else 2.0 ** -5
                for g, qg in enumerate(stub._hint_q[src]):
                    if qg == qg and not graph:
                        assert qg * 2.0 ** -k0[g] <= lim * (1 + 1e-6) or k0[g] == 0 and qg <= lim, (qg, k0)
                        assert k0[g] == 0 or qg * 2.0 ** -(k0[g] - 1) > lim                # the FIRST valid exponent
                if graph:
                    assert len(set(k0)) == 1 and k0[0] % 2 == 0
            else:
                assert plans == full and not (q == q and q < 2.0 ** (_lib.ADMM_MAX_CAND - 9))   # only a diverging run skips it
    # candidate-by-candidate mode: a window around the exits of step s-2, proofs below it
    stub.probe, stub._hint = "exact", None
    assert ADMMBasedOptimizer._probe_plans(stub, _lib.SRC_X) == full
    for ks in ([0, 1, 5, 14], [31, 32, 40, 63], [14, 14, 18, 14]):
        stub._hint = {_lib.SRC_X: ks, _lib.SRC_H: ks}
        plans = ADMMBasedOptimizer._probe_plans(stub, _lib.SRC_H)
        for entry in plans:
            _plan_ok(entry)
        k0, ncand = plans[0][0], plans[0][1]
        assert plans[1:] == full and ncand % 8 == 0
        for k, a in zip(ks, k0):
            assert a <= max(0, k - 5) and (a + ncand > k + 2 or ncand == _lib.ADMM_MAX_CAND)   # the exit of s-2 and two above it
