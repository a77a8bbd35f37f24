"""Oracle parity of the CUDA path AT THE BENCHMARK SHAPES (BASELINE.json configs[2..4]: (T, D, H) = (64, 16, 256),
(128, 64, 1024), (128, 9, 512)), with N small enough for the oracle to finish in seconds.

What the small-shape tests of test_gpu_parity.py do not reach and the bench runs: H = 1024 (32 h-k-blocks, 16 unit tiles
per sample tile), several sample tiles with ghost rows, time-chunking of the weight phase ON THE TENSOR-CORE PATH
(`scratch_bytes` set so that tc_chunk < T, ragged last chunk), the pre-activation store with chunking (and switched off),
the fp16-pair A^T R GEMM at K = 1024, the moment probes at T = 128 (and the exact probes), ADMM-LSTM-L at H = 1024 / 512,
and two sample shards at H = 1024.  Tolerances as everywhere: weights / gates / a 1e-4 scale-relative, duals
absolute 1e-4 * max(max|dual|, rho).  Reference lines matched: admm.py:282-343 (weights), 345-351 (gates), 504-510 (duals).
"""
import numpy as np
import pytest
import torch

from helpers import GOOGLE, HAR, WKEYS, rel_err, synthetic_problem

pytestmark = pytest.mark.gpu


def _need_gpu():
    if not torch.cuda.is_available():
        pytest.skip("needs a CUDA device")


def _scaled(params, n, h):
    """bench.py:bench_params -- rho_y keeps rho_y*N*H at the GoogleStock value (the Wy step is a fixed-size gradient step)."""
    rho = dict(params["rho"])
    rho["y"] = min(rho["y"], 5.62e-5 * (4224 * 10) / (float(n) * h))
    return {"rho": rho, "beta": dict(params["beta"])}


def _cmp(opt, ora, params, tag, rel=1e-4):
    from gpu_utils import np_state, weights_of
    gates, duals = np_state(opt)
    w = weights_of(opt)
    worst = 0.0
    for k in WKEYS:
        e = rel_err(w[k], ora.w[k])
        worst = max(worst, e)
        assert e < rel, (tag, "weight", k, e)
    for k in ("i", "f", "g", "o", "c", "h", "a"):
        e = rel_err(gates[k], ora.gates[k])
        worst = max(worst, e)
        assert e < rel, (tag, "gate", k, e)
    for k in ("i", "f", "g", "o", "c", "h", "y"):
        scale = max(float(np.max(np.abs(ora.duals[k]))), float(params["rho"][k]))
        err = float(np.max(np.abs(duals[k] - ora.duals[k])))
        assert err < 1e-4 * scale, (tag, "dual", k, err, scale)
    return worst


# (name, shape, params, variant, classification, timesteps per chunk, keep z, probe)
CASES = [
    ("cfg3-shortT", (256, 8, 64, 1024, 1), GOOGLE, "admm", False, 3, True, "moments"),
    ("cfg3-shortT-noz-exact", (256, 8, 64, 1024, 1), GOOGLE, "admm", False, 3, False, "exact"),
    ("cfg3-fast", (200, 6, 64, 1024, 1), GOOGLE, "no_dual_y", False, 4, True, "moments"),
    ("cfg3-fullT", (130, 128, 64, 1024, 1), GOOGLE, "admm", False, 8, True, "moments"),
    ("cfg2-fullT", (300, 64, 16, 256, 1), GOOGLE, "admm", False, 24, True, "moments"),
    ("cfg2-fullT-exact", (300, 64, 16, 256, 1), GOOGLE, "no_dual_y", False, 24, True, "exact"),
    ("cfg4-fullT", (200, 128, 9, 512, 6), HAR, "no_dual_y", True, 40, True, "moments"),
    ("cfg4-shortT-admm-noz", (260, 12, 9, 512, 6), HAR, "admm", True, 5, False, "moments"),
]


@pytest.mark.parametrize("name,shape,params,variant,cls,chunk,keepz,probe", CASES, ids=[c[0] for c in CASES])
def test_bench_shape_steps_vs_oracle(name, shape, params, variant, cls, chunk, keepz, probe):
    _need_gpu()
    from oracle.admm_oracle import OracleADMM
    from gpu_utils import make_opt
    n, t, d, h, o = shape
    x, y, w = synthetic_problem(n, t, d, h, o, seed=h + t, classification=cls)
    params = _scaled(params, n, h)
    ldn = (n + 127) // 128 * 128
    _, opt = make_opt(w, x, y, params, variant, use_tensor_cores=True, scratch_bytes=chunk * 32 * h * ldn,
                      keep_preactivations=keepz, probe=probe)
    assert opt.uses_tensor_cores and opt._tc_chunk == chunk and chunk < t
    assert opt.keeps_preactivations == keepz
    ora = OracleADMM(w, x, y, params, variant=variant)
    worst = _cmp(opt, ora, params, f"{name} init", rel=1e-5)
    for s in range(3):
        ora.step()
        opt.step()
        worst = max(worst, _cmp(opt, ora, params, f"{name} step{s}"))
        th = opt.theta_trace()
        # every backtracking decision is the oracle's (no knife edges on these data: margins are checked by the weights)
        diff = {k: (th[k], v) for k, v in ora.trace.items() if k in th and abs(th[k] - v) > 1e-6 * v}
        print(f"{name} step {s}: worst rel err {worst:.2e}; theta differences vs oracle: {diff}")
    assert opt._p.z_valid == int(keepz)


def test_bench_shape_two_shards_h1024():
    """SURVEY 8(e) at H = 1024: two sample shards (in-process rendezvous as the all-reduce) = the unsharded run = oracle."""
    _need_gpu()
    import threading
    from oracle.admm_oracle import OracleADMM
    from gpu_utils import make_opt, weights_of
    n, t, d, h, o = 300, 5, 64, 1024, 1
    x, y, w = synthetic_problem(n, t, d, h, o, seed=21)
    params = _scaled(GOOGLE, n, h)

    class FakeComm:
        def __init__(self, rank, shared):
            self.active, self.world_size, self.rank, self.shared = True, 2, rank, shared

        def _reduce(self, t_, op):
            torch.cuda.current_stream().synchronize()
            self.shared["buf"][self.rank] = t_
            self.shared["bar"].wait()
            total = op(self.shared["buf"][0], self.shared["buf"][1])
            self.shared["bar"].wait()
            t_.copy_(total)
            self.shared["bar"].wait()

        def allreduce_sum_(self, *tensors):
            for t_ in tensors:
                self._reduce(t_, torch.add)

        def allreduce_max_(self, *tensors):
            for t_ in tensors:
                self._reduce(t_, torch.maximum)

        def shard_range(self, n_total):
            half = 170                       # unequal shards: 170 + 130, both with ghost rows
            return (0, half) if self.rank == 0 else (half, n_total)

        def sum_int(self, v, device):
            return v

    shared = {"buf": [None, None], "bar": threading.Barrier(2)}
    opts, errs = [None, None], []

    def worker(rank):
        try:
            torch.cuda.set_device(0)
            _, o_ = make_opt(w, x, y, params, "admm", comm=FakeComm(rank, shared), sharding="slice", use_tensor_cores=True,
                             scratch_bytes=2 * 32 * h * 256)
            opts[rank] = o_
            for _ in range(3):
                o_.step()
            torch.cuda.synchronize()
        except Exception as exc:   # pragma: no cover
            errs.append(exc)
            shared["bar"].abort()

    th = [threading.Thread(target=worker, args=(r,)) for r in (0, 1)]
    [t_.start() for t_ in th]
    [t_.join() for t_ in th]
    assert not errs, errs
    ora = OracleADMM(w, x, y, params, variant="admm")
    for _ in range(3):
        ora.step()
    w0, w1 = weights_of(opts[0]), weights_of(opts[1])
    for k in WKEYS:
        assert np.array_equal(w0[k], w1[k]), k                     # replicas stay bit-identical
        assert rel_err(w0[k], ora.w[k]) < 1e-4, (k, rel_err(w0[k], ora.w[k]))
    for k in ("i", "f", "g", "o", "c", "h"):
        got = np.concatenate([opts[0].gates[k].cpu().numpy(), opts[1].gates[k].cpu().numpy()])
        assert rel_err(got, ora.gates[k]) < 1e-4, k


L_CASES = [("cfg3-L", (192, 8, 64, 1024), 3), ("cfg4-L", (130, 32, 9, 512), 12)]


@pytest.mark.parametrize("name,shape,chunk", L_CASES, ids=[c[0] for c in L_CASES])
def test_bench_shape_admm_l_vs_oracle(name, shape, chunk):
    """ADMM-LSTM-L at the bench shapes: the fp16-pair Gram / right-hand-side GEMM at K = 1024 with a chunked
    packing pass, against the oracle (admm_l/admm_lstm.py:107-163, main.py:139-191)."""
    _need_gpu()
    from admm_lstm_b200.admm_l import ADMMLOptimizer
    from oracle.admm_l_oracle import GATES, OracleADMML
    n, t, dd, h = shape
    rng = np.random.default_rng(n + h)
    x, y = rng.random((n, t, dd), dtype=np.float32), rng.random((n, 1), dtype=np.float32)
    W = {g: (rng.standard_normal((dd, h)) * 0.1).astype(np.float32) for g in GATES}
    U = {g: (rng.standard_normal((h, h)) * 0.1).astype(np.float32) for g in GATES}
    Wy = (rng.standard_normal((h, 1)) * 0.1).astype(np.float32)
    wt = {"Wy": torch.from_numpy(Wy)}
    for g in GATES:
        wt["W" + g], wt["U" + g] = torch.from_numpy(W[g]), torch.from_numpy(U[g])
    ldn = (n + 127) // 128 * 128
    opt = ADMMLOptimizer(wt, torch.from_numpy(x), torch.from_numpy(y), n_norm=float(n), use_tensor_cores=True,
                         scratch_bytes=chunk * 40 * h * ldn)
    assert opt.uses_tensor_cores and opt._tc_chunk == chunk and chunk < t
    ora = OracleADMML(W, U, Wy, x, y, n_norm=n)
    for k in range(3):
        opt.step()
        ora.step()
        got = {nm: v.cpu().numpy() for nm, v in opt.weights().items()}
        for g in GATES:
            assert rel_err(got["W" + g], ora.W[g]) < 1e-4, (name, k, "W" + g, rel_err(got["W" + g], ora.W[g]))
            assert rel_err(got["U" + g], ora.U[g]) < 1e-4, (name, k, "U" + g, rel_err(got["U" + g], ora.U[g]))
        assert rel_err(got["Wy"], ora.Wy) < 1e-4, (name, k)
        st = opt.state()
        for key, ref in (("h", ora.h), ("c", ora.c)):
            assert rel_err(st[key].cpu().numpy(), ref.transpose(1, 0, 2)) < 1e-4, (name, k, key)
        assert rel_err(st["a"].cpu().numpy(), ora.a) < 1e-4, (name, k)
