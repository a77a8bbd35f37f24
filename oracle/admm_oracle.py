"""CPU oracle for the ADMM-LSTM sweep  --  TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A plain numpy/fp32 restatement of the reference's per-iteration ADMM sweep
(`ADMMBasedOptimizer.step()` in /root/reference/admm.py and its "Fast" variant
/root/reference/admm.no_dual_y.py over the blocks/lstm.py cell).  Only `tests/`,
`__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference`
leg may import this module; the product path (admm_lstm_b200/) never does and
fails loudly when its CUDA library is missing.

Parity pin: the reference has no tests or golden vectors of its own (SURVEY.md
section 4), so this file is pinned against outputs of the reference itself, imported
unmodified in the build container by `tests/golden/make_golden.py`; the
resulting fixtures live in `tests/golden/*.npz` and `tests/test_oracle_golden.py`
checks every function below against them.

Every method cites the reference lines it restates.  The arithmetic is IEEE
fp32 in the reference's operation order (autograd calls are replaced by the
closed form autograd evaluates).  Differences from the reference that do NOT
change results: no `.clone()` traffic, no autograd graph.

`allreduce` hook: every sum that couples samples goes through
`self.allreduce(ndarray) -> ndarray`; the default is the identity.  The
world_size-2 gloo tests plug torch.distributed in here to check the sample
sharding scheme.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
GATES = ("i", "f", "g", "o")


def _sigmoid(x):
    with np.errstate(over="ignore"):
        return (F32(1.0) / (F32(1.0) + np.exp(-x, dtype=F32))).astype(F32)


def _tanh(x):
    return np.tanh(x, dtype=F32)


def _act(gate):
    return _tanh if gate == "g" else _sigmoid


def _d_act(gate, z):
    # admm.py:239-244
    if gate == "g":
        t = _tanh(z)
        return F32(1.0) - t * t
    s = _sigmoid(z)
    return s * (F32(1.0) - s)


def lstm_forward(weights, x):
    """blocks/lstm.py:65-88 (init_gate_variables): returns the state dict with
    i,f,g,o,c,h of shape [N,T+1,H] (slot 0 zero) and a = h_T @ out."""
    x = np.asarray(x, dtype=F32)
    n, t_len, _ = x.shape
    hsz = weights["h2i"].shape[0]
    st = {k: np.zeros((n, t_len + 1, hsz), dtype=F32) for k in ("i", "f", "g", "o", "c", "h")}
    for t in range(1, t_len + 1):
        x_t = x[:, t - 1, :]
        h_b = st["h"][:, t - 1, :]
        for gate in GATES:
            z = x_t @ weights["x2" + gate] + h_b @ weights["h2" + gate]
            st[gate][:, t, :] = _act(gate)(z)
        st["c"][:, t, :] = st["f"][:, t, :] * st["c"][:, t - 1, :] + st["i"][:, t, :] * st["g"][:, t, :]
        st["h"][:, t, :] = st["o"][:, t, :] * _tanh(st["c"][:, t, :])
    st["a"] = st["h"][:, t_len, :] @ weights["out"]
    return st


def mse(pred, target):
    """nn.MSELoss() as used by demo.py:316,341-342."""
    d = np.asarray(pred, dtype=F32) - np.asarray(target, dtype=F32)
    return float(np.mean(d * d, dtype=F32))


class OracleADMM:
    """Restates ADMMBasedOptimizer (admm.py:22-546 / admm.no_dual_y.py:12-493).

    variant: 'admm' -> admm.py (shipped with_dual_y=False); 'no_dual_y' -> the Fast file.
    weights: dict with keys x2{i,f,g,o} [D,H], h2{i,f,g,o} [H,H], out [H,O] (fp32).
    """

    def __init__(self, weights, train_x, train_y, params, variant="admm", with_dual_y=False,
                 allreduce=None, n_global=None, state=None):
        assert variant in ("admm", "no_dual_y")
        assert not (with_dual_y and variant == "no_dual_y")
        self.variant, self.with_dual_y = variant, with_dual_y
        self.w = {k: np.array(v, dtype=F32) for k, v in weights.items()}
        self.train_x = np.ascontiguousarray(train_x, dtype=F32)
        self.train_y = np.ascontiguousarray(train_y, dtype=F32)
        self.batch_size, self.seq_len, self.input_size = self.train_x.shape
        self.output_size = self.train_y.shape[1]
        self.hidden_size = self.w["h2i"].shape[0]
        # admm.py:489-502 uses self.batch_size; under sample sharding that is the GLOBAL N.
        self.n_global = int(n_global) if n_global is not None else self.batch_size
        self.allreduce = allreduce if allreduce is not None else (lambda v: v)
        # admm.py:131-160
        self.betas = {"wy": F32(params["beta"]["wy"])}
        for gate in GATES:
            self.betas["x2" + gate] = F32(params["beta"]["w" + gate])
            self.betas["h2" + gate] = F32(params["beta"]["v" + gate])
        self.rhos = {k: F32(params["rho"][k]) for k in ("i", "f", "g", "o", "c", "h", "y")}
        if state is None:
            # admm.py:164-173
            self.gates = lstm_forward(self.w, self.train_x)
            shp = (self.batch_size, self.seq_len + 1, self.hidden_size)
            self.duals = {k: np.zeros(shp, dtype=F32) for k in ("i", "f", "g", "o", "c", "h")}
            self.duals["y"] = np.zeros((self.batch_size, self.output_size), dtype=F32)
        else:
            self.gates = {k: np.array(v, dtype=F32) for k, v in state["gates"].items()}
            self.duals = {k: np.array(v, dtype=F32) for k, v in state["duals"].items()}
        self.trace = {}        # theta decisions of the last step, for diagnostics only
        self.f_evals = 0       # number of full-data original_func evaluations (cost model)
        # Test hooks for knife-edge backtracking decisions: `forced` maps a weight name
        # (or 'h_T') to the theta to apply instead of the one the loop finds; `margins` records, per name,
        # (theta, f(beta), est, f(w)) of every comparison the loop made.
        self.forced = {}
        self.margins = {}

    # ---- reductions that couple samples -------------------------------------------------
    def _sum(self, v):
        """torch.sum of an fp32 tensor -> fp32 scalar; coupled across shards."""
        return F32(self.allreduce(np.asarray(np.sum(v, dtype=F32), dtype=F32).reshape(1))[0])

    def _fro(self, v):
        return self._sum(v * v)       # admm.py:224-229

    # ---- step ---------------------------------------------------------------------------
    def step(self):
        """admm.py:62-78 / admm.no_dual_y.py:52-66."""
        self.update_wy()
        for gate in GATES:
            for src in ("x", "h"):
                self.update_weights(src, gate)
        for t in range(1, self.seq_len + 1):
            self.update_gates(t)
            if t == self.seq_len:
                self.update_primal_a()
            self.update_duals(t)
        if self.with_dual_y:
            self.update_dual_y()

    # ---- Wy -------------------------------------------------------------------------------
    def update_wy(self):
        """admm.py:246-280 ; admm.no_dual_y.py:226-249."""
        T = self.seq_len
        h = self.gates["h"][:, T, :]
        a = self.gates["a"]
        wy = self.w["out"]
        rho_y = self.rhos["y"]
        shift = (self.duals["y"] / rho_y) if self.with_dual_y else F32(0.0)

        def original_func(beta_t):
            return F32(0.5) * rho_y * self._fro(h @ beta_t - a - shift)

        r = h @ wy - a - shift
        if self.variant == "admm":
            theta = 1.0
            # autograd of 0.5*rho_y*||h@wy - a||^2 wrt wy: h^T @ (rho_y * r)   (admm.py:253-260)
            gradient = self.allreduce(h.T @ (rho_y * r))
        else:
            theta = 0.01
            gradient = rho_y * self.allreduce(h.T @ r)            # no_dual_y:232

        def estimated_func(beta_t, theta_t):
            d = beta_t - wy
            return original_func(beta_t) + np.sum(gradient * d, dtype=F32) + F32(0.5 * theta_t) * np.sum(d * d, dtype=F32)

        beta = wy + gradient / F32(theta)
        guard = 0
        while original_func(beta) > estimated_func(beta, theta):     # admm.py:272 (never true)
            theta *= 2
            beta = wy + gradient / F32(theta)
            guard += 1
            if guard > 64:
                break
        theta /= 2
        self.trace["wy"] = theta
        if self.variant == "admm":
            self.w["out"] = ((F32(theta) * wy - gradient) / (F32(theta) + self.betas["wy"])).astype(F32)
        else:
            self.w["out"] = ((F32(theta) * wy - gradient) / (F32(theta) + F32(2) * self.betas["wy"])).astype(F32)

    # ---- W, U per gate ----------------------------------------------------------------------
    def update_weights(self, src, gate):
        """admm.py:282-343 (identical in admm.no_dual_y.py:251-312)."""
        T = self.seq_len
        w = self.w[f"{src}2{gate}"]
        rho = self.rhos[gate]
        act = _act(gate)
        if src == "x":
            a_val, b_val, w_other = self.train_x, self.gates["h"], self.w["h2" + gate]
        else:
            a_val, b_val, w_other = self.gates["h"], self.train_x, self.w["x2" + gate]
        lam, gv = self.duals[gate], self.gates[gate]

        # compute_grad, admm.py:302-312
        grad = np.zeros_like(w)
        for t in range(1, T + 1):
            a_t, b_t = a_val[:, t - 1, :], b_val[:, t - 1, :]
            z = a_t @ w + b_t @ w_other
            grad += a_t.T @ ((act(z) - lam[:, t, :] / rho - gv[:, t, :]) * _d_act(gate, z))
        gradient = (self.allreduce(grad) * rho).astype(F32)

        def original_func(beta_t):          # admm.py:316-325
            self.f_evals += 1
            val = F32(0.0)
            for t in range(1, T + 1):
                a_t, b_t = a_val[:, t - 1, :], b_val[:, t - 1, :]
                val = F32(val + F32(0.5) * rho * self._fro(
                    act(a_t @ beta_t + b_t @ w_other) - lam[:, t, :] / rho - gv[:, t, :]))
            return val

        def estimated_func(beta_t, theta_t):  # admm.py:327-329 (re-evaluates f(w) each time)
            d = beta_t - w
            return (original_func(w) + np.sum(gradient * d, dtype=F32)
                    + F32(T * 0.5 * theta_t) * np.sum(d * d, dtype=F32))

        theta = 1
        beta = w + gradient / F32(theta)
        name = f"{src}2{gate}"
        self.margins[name] = []

        def keep_going(beta_t, theta_t):
            fb, est = original_func(beta_t), estimated_func(beta_t, theta_t)
            self.margins[name].append((float(theta_t), float(fb), float(est)))
            return fb > est

        while keep_going(beta, theta):                                # admm.py:334
            theta *= 2
            beta = w + gradient / F32(theta)
            if theta > 2.0 ** 60:         # the reference has no cap; a NaN exits its loop anyway
                break
        theta /= 2
        self.trace[name] = theta
        theta = self.forced.get(name, theta)
        tau = F32(0.5) * rho * F32(T) * F32(theta)                    # admm.py:341 left-to-right
        self.w[f"{src}2{gate}"] = ((tau * w - gradient) /
                                   (self.betas[f"{src}2{gate}"] + F32(0.5) * rho * F32(theta) * F32(T))).astype(F32)

    # ---- per-timestep primal updates ---------------------------------------------------------
    def update_gates(self, t):
        for gate in GATES:                      # admm.py:345-351
            self.update_primal_ifgo(gate, t)
        self.update_primal_c(t)
        self.update_primal_h(t)

    def _z(self, gate, t):
        return (self.train_x[:, t - 1, :] @ self.w["x2" + gate]
                + self.gates["h"][:, t - 1, :] @ self.w["h2" + gate])

    def update_primal_ifgo(self, gate, t):
        """admm.py:353-386."""
        g = self.gates
        z = self._z(gate, t)
        rho1, lam = self.rhos[gate], self.duals[gate][:, t, :]
        if gate == "i":
            p1, p2, p3 = g["g"][:, t, :], g["f"][:, t, :], g["c"][:, t - 1, :]
        elif gate == "f":
            p1, p2, p3 = g["c"][:, t - 1, :], g["g"][:, t, :], g["i"][:, t, :]
        elif gate == "g":
            p1, p2, p3 = g["i"][:, t, :], g["f"][:, t, :], g["c"][:, t - 1, :]
        else:
            p1, p2, p3 = _tanh(g["c"][:, t, :]), F32(0.0), F32(0.0)
        if gate == "o":
            var2, rho2, lam2 = g["h"][:, t, :], self.rhos["h"], self.duals["h"][:, t, :]
        else:
            var2, rho2, lam2 = g["c"][:, t, :], self.rhos["c"], self.duals["c"][:, t, :]
        new = -(lam - rho1 * _act(gate)(z) + (rho2 * (p2 * p3 - var2) - lam2) * p1) / (rho1 + rho2 * p1 * p1)
        g[gate][:, t, :] = new.astype(F32)

    def update_primal_c(self, t):
        """admm.py:388-436: the theta loop never iterates (f(c) > f(c) is false), theta ends 0.5."""
        g = self.gates
        c, o, h = g["c"][:, t, :], g["o"][:, t, :], g["h"][:, t, :]
        rho_h, rho_c = self.rhos["h"], self.rhos["c"]
        div_h = self.duals["h"][:, t, :] / rho_h
        div_c = self.duals["c"][:, t, :] / rho_c
        zed = h + div_h
        y = _tanh(c)
        u = y * o - zed
        gradient = (u * o) * (F32(1.0) - y * y)      # autograd of .5*||tanh(c)*o - z||^2  (admm.py:409-414)
        A = div_c - g["f"][:, t, :] * g["c"][:, t - 1, :] - g["i"][:, t, :] * g["g"][:, t, :]
        theta = F32(0.5)
        g["c"][:, t, :] = ((theta * c - gradient - rho_c * A) / (rho_c + theta)).astype(F32)

    def update_primal_h(self, t):
        """admm.py:439-487 ; admm.no_dual_y.py:414-449."""
        g = self.gates
        T = self.seq_len
        h = g["h"][:, t, :].copy()
        wy, rho_h, rho_y = self.w["out"], self.rhos["h"], self.rhos["y"]
        lam_h = self.duals["h"][:, t, :]
        a = g["a"]
        o = g["o"][:, t, :]
        tanh_c = _tanh(g["c"][:, t, :])
        if t < T:
            g["h"][:, t, :] = ((rho_h * o * tanh_c - lam_h) / rho_h).astype(F32)     # admm.py:455-457
            return
        shift = (self.duals["y"] / rho_y) if self.with_dual_y else F32(0.0)

        def original_func(beta_t):
            return F32(0.5) * rho_y * self._fro(beta_t @ wy - a - shift)

        r = h @ wy - a - shift
        if self.variant == "admm":
            gradient = (rho_y * r) @ wy.T           # autograd, admm.py:459-464
        else:
            gradient = rho_h * (r @ wy.T)           # no_dual_y:426

        def estimated_func(beta_t, theta_t):
            d = beta_t - h
            return original_func(h) + self._sum(gradient * d) + F32(0.5 * theta_t) * self._fro(d)

        def compute_h(theta_t):
            th = F32(theta_t)
            return ((th * h + rho_h * o * tanh_c - lam_h - gradient) / (th + rho_h)).astype(F32)

        theta, theta_max = 0.1, 1
        probe = compute_h if self.variant == "admm" else (lambda th: (gradient / F32(th)).astype(F32))
        beta = probe(theta)
        while original_func(beta) > estimated_func(beta, theta):      # admm.py:475-480
            theta *= 2
            beta = probe(theta)
            if theta >= theta_max:
                break
        theta /= 2
        self.trace["h_T"] = theta
        theta = self.forced.get("h_T", theta)
        g["h"][:, t, :] = compute_h(theta)                              # admm.py:482-487

    def update_primal_a(self):
        """admm.py:489-502 ; admm.no_dual_y.py:451-456 (batch_size is the global N)."""
        T = self.seq_len
        rho_y = self.rhos["y"]
        nb = F32(self.n_global)
        hw = self.gates["h"][:, T, :] @ self.w["out"]
        num = F32(2.0) * self.train_y + nb * rho_y * hw
        if self.with_dual_y:
            num = num - nb * self.duals["y"]
        self.gates["a"] = (num / (F32(2.0) + nb * rho_y)).astype(F32)

    # ---- duals -----------------------------------------------------------------------------------
    def update_duals(self, t):
        for gate in GATES:                      # admm.py:504-510
            self.update_dual_ifgo(gate, t)
        self.update_dual_c(t)
        self.update_dual_h(t)

    def update_dual_ifgo(self, gate, t):
        """admm.py:512-522."""
        self.duals[gate][:, t, :] = (self.duals[gate][:, t, :] + self.rhos[gate] * (
            self.gates[gate][:, t, :] - _act(gate)(self._z(gate, t)))).astype(F32)

    def update_dual_c(self, t):
        """admm.py:524-530."""
        g = self.gates
        self.duals["c"][:, t, :] = (self.duals["c"][:, t, :] + self.rhos["c"] * (g["c"][:, t, :] - (
            g["f"][:, t, :] * g["c"][:, t - 1, :] + g["i"][:, t, :] * g["g"][:, t, :]))).astype(F32)

    def update_dual_h(self, t):
        """admm.py:532-539 (only t == T)."""
        if t < self.seq_len:
            return
        g = self.gates
        self.duals["h"][:, t, :] = (self.duals["h"][:, t, :] + self.rhos["h"] * (
            g["h"][:, t, :] - g["o"][:, t, :] * _tanh(g["c"][:, t, :]))).astype(F32)

    def update_dual_y(self):
        """admm.py:541-546 (disabled as shipped: with_dual_y=False, admm.py:12)."""
        T = self.seq_len
        self.duals["y"] = (self.duals["y"] + self.rhos["y"] * (
            self.gates["a"] - self.gates["h"][:, T, :] @ self.w["out"])).astype(F32)

    # ---- observables the reference never computes; DEFINED by the build ------
    def snapshot_primal(self):
        snap = {k: self.gates[k].copy() for k in ("i", "f", "g", "o", "c", "h")}
        snap["a"] = self.gates["a"].copy()
        return snap

    def metrics(self, prev_primal=None):
        """Objective (augmented Lagrangian), primal and dual residual norms, in float64.

        r_g,t = gate_g,t - act(x_t W_g + h_{t-1} U_g)      g in {i,f,g,o}, t = 1..T
        r_c,t = c_t - f_t c_{t-1} - i_t g_t                  t = 1..T
        r_h   = h_T - o_T tanh(c_T)                          (lambda_h lives only at t = T)
        r_y   = a - h_T Wy
        primal^2 = sum ||r||^2 ; dual^2 = sum_v rho_v^2 ||v^{k+1} - v^k||^2 over v in i,f,g,o,c,h,a
        objective = ||a-y||^2/N + sum_w (beta_w/2)||w||^2 + sum_r (<lambda_r, r> + rho_r/2 ||r||^2)
        (Wy uses beta_wy/2 in 'admm' and beta_wy in 'no_dual_y', matching the two prox denominators.)
        All cross-sample sums go through allreduce.
        """
        T, g, d = self.seq_len, self.gates, self.duals
        f64 = np.float64
        prim = f64(0.0)
        pen = f64(0.0)
        for t in range(1, T + 1):
            for gate in GATES:
                r = (g[gate][:, t, :] - _act(gate)(self._z(gate, t))).astype(f64)
                prim += np.sum(r * r)
                pen += np.sum(d[gate][:, t, :].astype(f64) * r) + 0.5 * f64(self.rhos[gate]) * np.sum(r * r)
            r = (g["c"][:, t, :] - (g["f"][:, t, :] * g["c"][:, t - 1, :] + g["i"][:, t, :] * g["g"][:, t, :])).astype(f64)
            prim += np.sum(r * r)
            pen += np.sum(d["c"][:, t, :].astype(f64) * r) + 0.5 * f64(self.rhos["c"]) * np.sum(r * r)
        r = (g["h"][:, T, :] - g["o"][:, T, :] * _tanh(g["c"][:, T, :])).astype(f64)
        prim += np.sum(r * r)
        pen += np.sum(d["h"][:, T, :].astype(f64) * r) + 0.5 * f64(self.rhos["h"]) * np.sum(r * r)
        r = (g["a"] - g["h"][:, T, :] @ self.w["out"]).astype(f64)
        prim += np.sum(r * r)
        pen += 0.5 * f64(self.rhos["y"]) * np.sum(r * r)
        if self.with_dual_y:
            pen += np.sum(d["y"].astype(f64) * r)
        loss = np.sum((g["a"] - self.train_y).astype(f64) ** 2)
        dual = f64(0.0)
        if prev_primal is not None:
            for k in ("i", "f", "g", "o", "c", "h"):
                dv = (g[k] - prev_primal[k]).astype(f64)
                dual += f64(self.rhos[k]) ** 2 * np.sum(dv * dv)
            dv = (g["a"] - prev_primal["a"]).astype(f64)
            dual += f64(self.rhos["y"]) ** 2 * np.sum(dv * dv)
        red = self.allreduce(np.array([prim, pen, loss, dual], dtype=f64))
        prim, pen, loss, dual = (f64(v) for v in red)
        reg = f64(0.0)
        for gate in GATES:
            for src in ("x", "h"):
                wv = self.w[f"{src}2{gate}"].astype(f64)
                reg += 0.5 * f64(self.betas[f"{src}2{gate}"]) * np.sum(wv * wv)
        wv = self.w["out"].astype(f64)
        reg += (0.5 if self.variant == "admm" else 1.0) * f64(self.betas["wy"]) * np.sum(wv * wv)
        return {"objective": float(loss / self.n_global + reg + pen),
                "primal_residual": float(np.sqrt(prim)),
                "dual_residual": float(np.sqrt(dual)),
                "loss_term": float(loss / self.n_global)}

    def predict(self, x):
        """model(x) as demo.py:341-342 evaluates it: LSTM forward with the current weights."""
        return lstm_forward(self.w, x)["a"]
