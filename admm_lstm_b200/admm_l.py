"""ADMM-LSTM-L on B200 -- host side.

Mirrors /root/reference/comparison_experiment/admm_l/main.py: `admm_l_demo(num_epochs, n_hiddens, train_x, train_y,
test_x, test_y, save=False)` returns the same dictionary and `LSTM_L` has the reference's parameter names
(main.py:28-49), so comparison_experiment/comparison.py:174-178 and SAVED_MODELS/ADMM-LSTM-L.pt keep working.  The
iteration itself (main.py:139-188 over the update functions of admm_l/admm_lstm.py) is `ADMMLOptimizer.step()`:

  * every N*T-sized operation is a CUDA kernel behind include/admm_lstm_b200.h (admm_l_* entry points);
  * the eight W/U updates and W_y use Gram / right-hand-side sums (admm_l_sums, admm_l_sums_last): one all-reduce of
    small fp64 matrices per iteration when the samples are sharded; the O(K^2 H) algebra on those replicated sums --
    gradient, the closed-form exit of the reference's backtracking loop and the prox step -- runs on the device in
    fp64 (torch, no host synchronisation);
  * the sweep needs the reference's global max / sum scalars inside every timestep: two tiny all-reduces per timestep.

Backtracking exits.  The subproblems are quadratic, so `while Func2 > Func1: theta *= 2` (admm_lstm.py:110-125) exits
at the smallest theta = theta0 * 2^k with  rho * <G, S G> / theta <= ||G||^2  (S the Gram matrix of the operand, G the
gradient; derivation in DESIGN.md).  The reference evaluates both sides from fp32 sums over all samples; the two agree
except when the reference's own comparison is decided by rounding noise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Dict, List, Optional

import torch
from torch import nn
from torch.nn import Parameter

from . import _lib
from .comm import Comm
from .logging_utils import info
from .optimizer import _feature_major, _require_cuda, _round_up, _stream_ptr

# One all-gather of the (MAX, SUM) pair per timestep instead of two all-reduces: measured SLOWER (2 GPUs, cfg4 shape: 144.3
# vs 141.3 ms per iteration -- the local reduction's extra small launches cost more than the saved collective), so off.
_PAIR_COLLECTIVE = os.environ.get("ADMM_L_PAIR_COLLECTIVE", "0") != "0"
_ORDER = ("i", "f", "g", "o")          # storage order of the stacked weights / state (as the main path)
_UPDATE_ORDER = ("g", "o", "i", "f")   # main.py:141-148


class LSTM_L(nn.Module):
    """main.py:28-73 (same parameter names and forward)."""

    def __init__(self, input_size, hidden_size, Ui, Wi, Wo, Uf, Wf, Uo, Ug, Wg, Wy):
        super().__init__()
        self.input_size, self.hidden_size = input_size, hidden_size
        self.W_hi, self.W_ii = Parameter(Ui), Parameter(Wi)
        self.W_hf, self.W_if = Parameter(Uf), Parameter(Wf)
        self.W_ho, self.W_io = Parameter(Uo), Parameter(Wo)
        self.W_hg, self.W_ig = Parameter(Ug), Parameter(Wg)
        self.W_y = Parameter(Wy)

    def forward(self, x, init_states=None):
        n = x.size(0)
        if init_states is None:
            h_t = torch.zeros((n, self.hidden_size), device=x.device)
            c_t = torch.zeros((n, self.hidden_size), device=x.device)
        else:
            h_t, c_t = init_states
        for t in range(x.size(1)):
            x_t = x[:, t, :]
            i_t = torch.sigmoid(h_t @ self.W_hi + x_t @ self.W_ii)
            f_t = torch.sigmoid(h_t @ self.W_hf + x_t @ self.W_if)
            o_t = torch.sigmoid(h_t @ self.W_ho + x_t @ self.W_io)
            g_t = torch.tanh(h_t @ self.W_hg + x_t @ self.W_ig)
            c_t = f_t * c_t + i_t * g_t
            h_t = o_t * torch.tanh(c_t)
        return h_t @ self.W_y


def exit_theta(ratio: torch.Tensor, theta0: float) -> torch.Tensor:
    """Exit of `while Func2 > Func1: theta *= 2` for a quadratic subproblem: the smallest theta0 * 2^k (k >= 0) with
    theta >= ratio, ratio = rho <G, S G> / ||G||^2 (NaN for G = 0 -> theta0).  No host synchronisation."""
    ratio = torch.nan_to_num(ratio, nan=0.0, posinf=0.0)
    k = torch.clamp(torch.ceil(torch.log2(torch.clamp(ratio / theta0, min=1e-300))), min=0.0)
    return theta0 * torch.pow(torch.tensor(2.0, dtype=torch.float64, device=ratio.device), k)


def l_weight_phase(s_xx, ax, ah, stt, pt, wx, wh, wy, *, rho_s, rho11, lam_w, lam_u, thetas=None) -> None:
    """W_y, then (W, U) per gate in the order g, o, i, f (main.py:140-148) from the reduced Gram / right-hand-side sums,
    in place on the fp32 weights `wx [4,D,H]`, `wh [4,H,H]`, `wy [H]` (storage order i, f, g, o).  fp64 algebra on
    replicated data; device-agnostic (the CPU tests drive it with sums computed from the oracle's state).

      ax[g] = sum_t x_t^T V_g,t (g < 4), ax[4] = S_xh;  ah[g] = sum_t h_{t-1}^T V_g,t, ah[4] = S_hh;
      stt = h_T^T h_T;  pt = h_T^T (a + lambda11/rho11);  V_g = z_g + lambda_g/rho   (admm_lstm.py:76-163)."""
    f64 = torch.float64
    thetas = thetas if thetas is not None else {}
    s_xh, s_hh = ax[4], 0.5 * (ah[4] + ah[4].t())
    s_tt = 0.5 * (stt + stt.t())
    # update_Wy (:92-104): theta0 = 0.01, Wy <- Wy + g/theta (no regulariser)
    wy64 = wy.to(f64)
    g = rho11 * (pt - s_tt @ wy64)
    th = exit_theta(rho11 * (g @ (s_tt @ g)) / (g @ g), 0.01)
    thetas["wy"] = th
    wy.copy_(wy + g.float() / th.float())
    # update_W / update_U (:107-163): w <- (theta w - G)/(lambda0 + theta)
    for gname in _UPDATE_ORDER:
        q = _ORDER.index(gname)
        w, u = wx[q].to(f64), wh[q].to(f64)
        gw = rho_s * (s_xx @ w + s_xh @ u - ax[q])
        th = exit_theta(rho_s * (gw * (s_xx @ gw)).sum() / (gw * gw).sum(), 1.0)
        thetas["W" + gname] = th
        thf = th.float()
        wx[q].copy_((thf * wx[q] - gw.float()) / (lam_w + thf))
        w = wx[q].to(f64)
        gu = rho_s * (s_xh.t() @ w + s_hh @ u - ah[q])
        th = exit_theta(rho_s * (gu * (s_hh @ gu)).sum() / (gu * gu).sum(), 1.0)
        thetas["U" + gname] = th
        thf = th.float()
        wh[q].copy_((thf * wh[q] - gu.float()) / (lam_u + thf))


class ADMMLOptimizer(object):
    """State and iteration of ADMM-LSTM-L.

    weights: {'W'+g: [D,H], 'U'+g: [H,H] for g in f,i,o,g; 'Wy': [H,1]} (the reference draws them with
    randn * 0.1, main.py:75-83).  Hyper-parameters default to main.py:111-129; `n_norm` is the constant 4224 of
    update_a (admm_lstm.py:263-264)."""

    def __init__(self, weights: Dict[str, torch.Tensor], train_x: torch.Tensor, train_y: torch.Tensor, *,
                 n_norm: float = 4224.0, comm: Optional[Comm] = None, sharding: str = "slice",
                 use_tensor_cores: Optional[bool] = None, scratch_bytes: Optional[int] = None) -> None:
        self._lib = _lib.load()
        self.device = _require_cuda()
        if self._lib.admm_device_ok() <= 0:
            raise _lib.AdmmLibraryError("no sm_100 (B200) device visible to libadmm_lstm_b200.so")
        self.comm = comm if comm is not None else Comm()
        self.kernel_events = None
        assert train_x.dim() == 3 and train_y.dim() == 2 and train_y.size(1) == 1, "train_x [N,T,D], train_y [N,1]"
        self.batch_size, self.seq_len, self.input_size = (int(v) for v in train_x.shape)
        self.hidden_size = int(weights["Wy"].shape[0])
        if self.comm.active and sharding == "slice":
            lo, hi = self.comm.shard_range(self.batch_size)
            local_x, local_y = train_x[lo:hi], train_y[lo:hi]
            self.n_global = self.batch_size
        else:
            local_x, local_y = train_x, train_y
            self.n_global = self.comm.sum_int(self.batch_size, self.device) if self.comm.active else self.batch_size
        N, T, D, H = int(local_x.size(0)), self.seq_len, self.input_size, self.hidden_size
        self.n_local, self.ldn = N, _round_up(N, 128)
        dev, f32, f64 = self.device, torch.float32, torch.float64
        ldn = self.ldn

        self._x = _feature_major(local_x, ldn, dev)
        self._y = _feature_major(local_y, ldn, dev)
        zeros = lambda: torch.zeros((T + 1, H, ldn), dtype=f32, device=dev)      # noqa: E731
        self._gate = {k: zeros() for k in ("i", "f", "g", "o", "c", "h")}
        self._z = {k: zeros() for k in _ORDER}
        self._lam_s = {k: zeros() for k in _ORDER}
        self._lam_p = {k: zeros() for k in _ORDER}
        self._lam9, self._lam10 = zeros(), zeros()
        self._a = torch.zeros((1, ldn), dtype=f32, device=dev)
        self._lam11 = torch.zeros((1, ldn), dtype=f32, device=dev)
        self._wx = torch.stack([weights["W" + g].to(dev, f32) for g in _ORDER]).contiguous()
        self._wh = torch.stack([weights["U" + g].to(dev, f32) for g in _ORDER]).contiguous()
        self._wy = weights["Wy"].to(dev, f32).reshape(H).contiguous()

        budget = scratch_bytes if scratch_bytes is not None else int(os.environ.get("ADMM_LSTM_SCRATCH_BYTES", 4 << 30))
        self._tc_chunk = max(1, min(T, budget // (40 * H * ldn)))
        self._scratch = torch.empty(max(10 * H * self._tc_chunk * ldn, 4 * H * ldn), dtype=f32, device=dev)
        self._tmp = torch.empty(ldn, dtype=f32, device=dev)
        # per slot: [0..3] max |gate - lambda_p/rho| (measured when the values are written, one iteration ahead),
        # [4] the mid-timestep max of update_c; two buffers: this iteration's and the next one's
        self._red_max = [torch.zeros((T + 1, 8), dtype=f32, device=dev) for _ in range(2)]
        self._red_cur = 0
        self._red_sum = torch.zeros(T + 1, dtype=f64, device=dev)
        # [acc_x (5*D*H) | acc_h (5*H*H) | s_tt (H*H) | p_t (H)]: one all-reduce per iteration
        self._n_ax, self._n_ah = 5 * D * H, 5 * H * H
        self._sums = torch.zeros(self._n_ax + self._n_ah + H * H + H, dtype=f64, device=dev)
        self._sxx = torch.zeros((D, D), dtype=f64, device=dev)
        self._theta_h = torch.ones(1, dtype=f32, device=dev)
        self.thetas: Dict[str, torch.Tensor] = {}

        lp = _lib.LProblem()
        p = lp.base
        p.n, p.n_global, p.ldn = N, self.n_global, ldn
        p.T, p.D, p.H, p.O = T, D, H, 1
        p.variant, p.with_dual_y = 0, 0
        p.x, p.y = self._x.data_ptr(), self._y.data_ptr()
        for q, k in enumerate(("i", "f", "g", "o", "c", "h")):
            p.gate[q] = self._gate[k].data_ptr()
        p.a, p.dual_y = self._a.data_ptr(), self._lam11.data_ptr()
        p.wx, p.wh, p.wy = self._wx.data_ptr(), self._wh.data_ptr(), self._wy.data_ptr()
        for q, k in enumerate(_ORDER):
            lp.z[q], lp.lam_s[q], lp.lam_p[q] = self._z[k].data_ptr(), self._lam_s[k].data_ptr(), self._lam_p[k].data_ptr()
        lp.lam9, lp.lam10 = self._lam9.data_ptr(), self._lam10.data_ptr()
        hp = lp.hp
        hp.rho_s = hp.rho_p = hp.rho9 = hp.rho10 = 1.0                      # main.py:115,120,125,127
        hp.rho11 = 0.0001                                                    # main.py:129
        hp.lam_w = hp.lam_u = hp.lam_y = 1e-6                                # main.py:112-114
        hp.n_norm = float(n_norm)
        self._lp, self._lpp, self._bp = lp, C.byref(lp), C.byref(lp.base)
        self._tc_ws = None
        ws_bytes = int(self._lib.admm_tc_workspace_bytes(self._bp))
        want_tc = (ws_bytes > 0) if use_tensor_cores is None else bool(use_tensor_cores)
        if want_tc and ws_bytes <= 0:
            raise _lib.AdmmLibraryError(f"tensor-core path requested but shape (D={D}, H={H}) is not eligible")
        if want_tc:
            self._tc_ws = torch.zeros(ws_bytes, dtype=torch.uint8, device=dev)
            p.tc_ws, p.tc_ws_bytes = self._tc_ws.data_ptr(), ws_bytes
            self._call("admm_tc_refresh", self._bp, _lib.TC_WEIGHTS | _lib.TC_INPUTS, _stream_ptr())
        self.uses_tensor_cores = want_tc

        st = _stream_ptr()
        for s in range(1, T + 1):                                            # main.py:105 LSTM_Forward(train_x)
            self._call("admm_l_forward_t", self._lpp, s, self._scratch.data_ptr(), self._red_max[0][s].data_ptr(), st)
        self._call("admm_l_output", self._lpp, st)
        self._call("admm_l_gram_xx", self._lpp, self._sxx.data_ptr(), st)
        self.comm.allreduce_sum_(self._sxx)
        self._sxx = 0.5 * (self._sxx + self._sxx.t())

    # ------------------------------------------------------------------------------------------ plumbing
    def _call(self, name, *args) -> None:
        if self.kernel_events is not None:
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            rc = getattr(self._lib, name)(*args)
            e1.record()
            self.kernel_events.append((name, e0, e1))
        else:
            rc = getattr(self._lib, name)(*args)
        if rc != 0:
            _lib.check(rc, name)

    def enable_kernel_timing(self, enabled: bool = True) -> None:
        """Bracket every C-ABI call with CUDA events on the launching stream (bench.py roofline)."""
        self.kernel_events = [] if enabled else None

    def kernel_time_summary(self):
        torch.cuda.synchronize()
        out = {}
        for name, e0, e1 in self.kernel_events or []:
            n, ms = out.get(name, (0, 0.0))
            out[name] = (n + 1, ms + e0.elapsed_time(e1))
        return out

    def refresh_inputs(self, train_x: torch.Tensor, train_y: torch.Tensor) -> None:
        """Re-upload this rank's samples from (pinned) host memory -- the host->device leg of bench.py's end-to-end
        measurement; S_xx is recomputed from the new inputs."""
        n = self.n_local
        self._x[:, :, :n].copy_(train_x.to(self.device, non_blocking=True).permute(1, 2, 0))
        self._y[:, :n].copy_(train_y.to(self.device, non_blocking=True).t())
        st = _stream_ptr()
        if self._tc_ws is not None:
            self._call("admm_tc_refresh", self._bp, _lib.TC_INPUTS, st)
        sxx = torch.zeros_like(self._sxx)
        self._call("admm_l_gram_xx", self._lpp, sxx.data_ptr(), st)
        self.comm.allreduce_sum_(sxx)
        self._sxx = 0.5 * (sxx + sxx.t())

    def _time_chunks(self):
        T, tc = self.seq_len, self._tc_chunk
        return [(t0, min(tc, T - t0)) for t0 in range(0, T, tc)]

    # ------------------------------------------------------------------------------------------ iteration
    def step(self) -> None:
        """One iteration of main.py:139-188."""
        st = _stream_ptr()
        lpp = self._lpp
        D, H, T = self.input_size, self.hidden_size, self.seq_len
        hp = self._lp.hp
        f64 = torch.float64
        # ---- Gram / right-hand-side sums of the current state (h, z, lambda do not change during the weight phase)
        self._sums.zero_()
        ax, ah = self._sums[: self._n_ax], self._sums[self._n_ax: self._n_ax + self._n_ah]
        stt = self._sums[self._n_ax + self._n_ah: self._n_ax + self._n_ah + H * H]
        pt = self._sums[self._n_ax + self._n_ah + H * H:]
        for t0, tc in self._time_chunks():
            self._call("admm_l_sums", lpp, t0, tc, self._scratch.data_ptr(), ax.data_ptr(), ah.data_ptr(), st)
        self._call("admm_l_sums_last", lpp, stt.data_ptr(), pt.data_ptr(), st)
        self.comm.allreduce_sum_(self._sums)
        l_weight_phase(self._sxx, ax.view(5, D, H), ah.view(5, H, H), stt.view(H, H), pt, self._wx, self._wh, self._wy,
                       rho_s=float(hp.rho_s), rho11=float(hp.rho11), lam_w=float(hp.lam_w), lam_u=float(hp.lam_u),
                       thetas=self.thetas)
        if self._tc_ws is not None:
            self._call("admm_tc_refresh", self._bp, _lib.TC_WEIGHTS, st)

        # ---- theta of update_h at t = T-1 (:251-257): Func2 > Func1  <=>  theta < rho11 ||Wy||^2
        wy = self._wy.to(f64)
        th = exit_theta(float(hp.rho11) * (wy @ wy), 1.0)
        self.thetas["h"] = th
        self._theta_h.copy_(th.float().reshape(1))

        # ---- sweep (main.py:149-188)
        cur, nxt = self._red_max[self._red_cur], self._red_max[1 - self._red_cur]
        self.comm.allreduce_max_(cur)                      # the T x 4 gate maxima of this iteration in one collective
        nxt.zero_()
        self._red_sum.zero_()
        sp = self._scratch.data_ptr()
        for s in range(1, T + 1):
            rm, rs = cur[s], self._red_sum[s:s + 1]
            self._call("admm_l_sweep_gates", lpp, s, sp, rm.data_ptr(), rs.data_ptr(), st)
            if _PAIR_COLLECTIVE:
                self.comm.allreduce_max_and_sum_(rm[4:5], rs)      # one collective per timestep (MAX and SUM scalar together)
            else:
                self.comm.allreduce_max_(rm[4:5])
                self.comm.allreduce_sum_(rs)
            self._call("admm_l_sweep_cell", lpp, s, sp, rm.data_ptr(), rs.data_ptr(), nxt[s].data_ptr(), st)
        self._call("admm_l_last", lpp, self._theta_h.data_ptr(), self._tmp.data_ptr(), sp, nxt[T].data_ptr(), st)
        self._red_cur = 1 - self._red_cur

    # ------------------------------------------------------------------------------------------ observables
    def weights(self) -> Dict[str, torch.Tensor]:
        out = {}
        for q, g in enumerate(_ORDER):
            out["W" + g], out["U" + g] = self._wx[q].clone(), self._wh[q].clone()
        out["Wy"] = self._wy.reshape(-1, 1).clone()
        return out

    def state(self) -> Dict[str, torch.Tensor]:
        """The reference's dictionaries as [N, T, H] tensors (slot t+1 of the device layout = timestep t)."""
        n = self.n_local
        view = lambda t: t[1:].permute(2, 0, 1)[:n]                           # noqa: E731
        out = {"c": view(self._gate["c"]), "h": view(self._gate["h"]), "lam9": view(self._lam9), "lam10": view(self._lam10),
               "a": self._a.t()[:n], "lam11": self._lam11.t()[:n]}
        for g in _ORDER:
            out[g], out["z" + g] = view(self._gate[g]), view(self._z[g])
            out["lams_" + g], out["lamp_" + g] = view(self._lam_s[g]), view(self._lam_p[g])
        return out

    def model(self) -> LSTM_L:
        w = self.weights()
        return LSTM_L(self.input_size, self.hidden_size, w["Ui"], w["Wi"], w["Wo"], w["Uf"], w["Wf"], w["Uo"], w["Ug"],
                      w["Wg"], w["Wy"])

    def predict(self, x: torch.Tensor) -> torch.Tensor:
        """LSTM_Forward(x)[-1] (main.py:85-103) through the library's forward kernel."""
        from .optimizer import predict_cuda

        class _M:                                                            # the nine tensors under the main path's names
            pass
        m = _M()
        m.hidden_size, m.output_size = self.hidden_size, 1
        for q, g in enumerate(_ORDER):
            setattr(m, "x2" + g, self._wx[q])
            setattr(m, "h2" + g, self._wh[q])
        m.out = self._wy.reshape(-1, 1)
        return predict_cuda(m, x.to(self.device))


def admm_l_demo(num_epochs, n_hiddens, train_x, train_y, test_x, test_y, save=False) -> Dict[str, List[float] or str]:
    """main.py:75-208 with the iteration on the GPU.  Same return value; `save` stores SAVED_MODELS/ADMM-LSTM-L.pt
    (an LSTM_L module, as demo.save_model does, demo.py:302-308)."""
    dev = _require_cuda()
    n_feature = test_x.size(2)
    weights = {}
    for g in ("f", "i", "o", "g"):                                            # main.py:75-82, same draw order
        weights["W" + g] = torch.randn(n_feature, n_hiddens) * 0.1
        weights["U" + g] = torch.randn(n_hiddens, n_hiddens) * 0.1
    weights["Wy"] = torch.randn(n_hiddens, 1) * 0.1
    opt = ADMMLOptimizer(weights, train_x, train_y)
    train_y_d, test_y_d = train_y.to(dev), test_y.to(dev)
    mse = lambda pred, y: torch.mean(torch.square(y - pred)).item()          # noqa: E731
    loss_train, loss_test = [mse(opt.predict(train_x), train_y_d)], [mse(opt.predict(test_x), test_y_d)]
    info(f"Loss at the beginning: {loss_train[0]}")
    for k in range(num_epochs):
        opt.step()
        loss_train.append(mse(opt.predict(train_x), train_y_d))
        loss_test.append(mse(opt.predict(test_x), test_y_d))
        info(f"ADMM-LSTM-L: k = {k + 1}, loss train = {loss_train[-1]}, loss test = {loss_test[-1]}")
    if save and opt.comm.rank == 0:                                           # replicated weights: one writer
        os.makedirs("SAVED_MODELS", exist_ok=True)
        torch.save(opt.model().cpu(), os.path.join("SAVED_MODELS", "ADMM-LSTM-L.pt"))
    return {"name": "ADMM-LSTM-L", "train_loss": loss_train, "val_loss": loss_test}
