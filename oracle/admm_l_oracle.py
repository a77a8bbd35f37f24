"""CPU oracle for ADMM-LSTM-L -- TEST INFRASTRUCTURE, not a product path.

numpy fp32 restatement of /root/reference/comparison_experiment/admm_l/admm_lstm.py (the update functions)
driven in the order of /root/reference/comparison_experiment/admm_l/main.py:139-191.  Only tests/,
__graft_entry__.smoke() and bench.py's cpu_baseline leg may import this file.

Pinned: tests/test_oracle_l_golden.py holds it to the fixtures tests/golden/l_*.npz, which were produced by the
reference's own functions (tests/golden/make_golden_l.py).

State layout mirrors the reference: per timestep t = 0..T-1 arrays [N, H]; h[-1] = c[-1] = 0 (main.py:87-88).
Gate order everywhere: f, i, o, g (the reference's lambda numbering: singular 1,3,5,7 / plural 2,4,6,8).
"""
from __future__ import annotations

import numpy as np

F = np.float32
GATES = ("f", "i", "o", "g")
ORDER_W = ("g", "o", "i", "f")            # main.py:141-148


def _sig(x):
    return (F(1.0) / (F(1.0) + np.exp(-x, dtype=F))).astype(F)


def _tanh(x):
    return np.tanh(x, dtype=F)


class OracleADMML:
    """hyper: lambda00/02/03 = 1e-6, RHO_singular = RHO_plural = rho9 = rho10 = 1, rho11 = 1e-4 (main.py:111-129);
    update_a's constant 4224 (admm_lstm.py:263-264) is kept as `n_norm`."""

    def __init__(self, W, U, Wy, x, y, n_norm=4224):
        self.x = np.asarray(x, F)
        self.y = np.asarray(y, F)
        self.N, self.T, self.D = self.x.shape
        self.W = {g: np.asarray(W[g], F).copy() for g in GATES}
        self.U = {g: np.asarray(U[g], F).copy() for g in GATES}
        self.Wy = np.asarray(Wy, F).copy()
        self.H = self.Wy.shape[0]
        self.l00 = self.l02 = self.l03 = 1e-6
        self.rs = self.rp = self.r9 = self.r10 = F(1.0)
        self.r11 = F(0.0001)
        self.n_norm = n_norm
        N, T, H = self.N, self.T, self.H
        self.z, self.g, self.c, self.h, self.a = self.forward(self.x)
        self.lam_s = {g: np.zeros((T, N, H), F) for g in GATES}
        self.lam_p = {g: np.zeros((T, N, H), F) for g in GATES}
        self.lam9 = np.zeros((T, N, H), F)
        self.lam10 = np.zeros((T, N, H), F)
        self.lam11 = np.zeros((N, 1), F)
        self.theta = {}

    # main.py:85-103 ----------------------------------------------------------------------------------
    def forward(self, x):
        N, T = x.shape[0], x.shape[1]
        H = self.H
        z = {g: np.zeros((T, N, H), F) for g in GATES}
        gt = {g: np.zeros((T, N, H), F) for g in GATES}
        c = np.zeros((T, N, H), F)
        h = np.zeros((T, N, H), F)
        hp, cp = np.zeros((N, H), F), np.zeros((N, H), F)
        for t in range(T):
            for g in GATES:
                z[g][t] = x[:, t, :] @ self.W[g] + hp @ self.U[g]
                gt[g][t] = _tanh(z[g][t]) if g == "g" else _sig(z[g][t])
            c[t] = gt["f"][t] * cp + gt["i"][t] * gt["g"][t]
            h[t] = gt["o"][t] * _tanh(c[t])
            hp, cp = h[t], c[t]
        return z, gt, c, h, (h[T - 1] @ self.Wy).astype(F)

    def predict(self, x):
        return self.forward(np.asarray(x, F))[-1]

    def _hprev(self, t):
        return self.h[t - 1] if t > 0 else np.zeros((self.N, self.H), F)

    def _cprev(self, t):
        return self.c[t - 1] if t > 0 else np.zeros((self.N, self.H), F)

    # admm_lstm.py:76-104 -----------------------------------------------------------------------------
    def update_wy(self):
        a, Wy, h, lam, rho = self.a, self.Wy, self.h[self.T - 1], self.lam11, self.r11

        def eq1(w):
            temp = a - h @ w + lam / rho
            return rho / F(2) * np.sum(temp * temp, dtype=F)

        def eq1_w(w):
            temp = a - h @ w + lam / rho
            return (rho * (h.T @ temp)).astype(F)

        grad = eq1_w(Wy)
        theta = F(0.01)
        zeta = Wy + grad / theta
        while True:
            temp = zeta - Wy
            p = eq1(Wy) - np.sum(eq1_w(Wy) * temp, dtype=F) + np.sum(theta * temp * temp, dtype=F) / F(2)
            if not (eq1(zeta) > p):
                break
            theta = theta * F(2)
            zeta = Wy + grad / theta
        self.theta["wy"] = float(theta)
        self.Wy = zeta.astype(F)

    # admm_lstm.py:107-163 ----------------------------------------------------------------------------
    def _residuals(self, g, W, U):
        """Form10 of update_W/update_U for every t: -z + x W + h_{t-1} U - lambda/rho."""
        out = []
        for t in range(self.T):
            out.append(-self.z[g][t] + self.x[:, t, :] @ W + self._hprev(t) @ U - self.lam_s[g][t] / self.rs)
        return out

    def update_weight(self, g, src):
        W, U, rho = self.W[g], self.U[g], self.rs
        lam0 = self.l00 if src == "x" else self.l02
        res = self._residuals(g, W, U)
        form11 = F(0)
        grad = np.zeros_like(W if src == "x" else U)
        for t in range(self.T):
            form11 = form11 + rho * np.sum(res[t] * res[t], dtype=F)
            a_t = self.x[:, t, :] if src == "x" else self._hprev(t)
            grad = grad + rho * (a_t.T @ res[t])
        cur = W if src == "x" else U
        theta = F(1)
        while True:
            new = cur - grad / theta
            func1 = F(0.5) * form11 + np.sum(grad * (new - cur), dtype=F) + F(0.5) * theta * np.sum((new - cur) * (new - cur), dtype=F)
            res2 = self._residuals(g, new, U) if src == "x" else self._residuals(g, W, new)
            form21 = F(0)
            for t in range(self.T):
                form21 = form21 + rho * np.sum(res2[t] * res2[t], dtype=F)
            func2 = F(0.5) * form21
            if not (func2 > func1):
                break
            theta = theta * F(2)
        self.theta[("W" if src == "x" else "U") + g] = float(theta)
        new = ((theta * cur - grad) / F(lam0 + float(theta))).astype(F)
        if src == "x":
            self.W[g] = new
        else:
            self.U[g] = new

    # admm_lstm.py:166-188 ----------------------------------------------------------------------------
    def update_z(self, g, t):
        z, out = self.z[g][t], self.g[g][t]
        lam1, lam2, rho1, rho2 = self.lam_s[g][t], self.lam_p[g][t], self.rs, self.rp
        temp = np.max(np.abs(out - lam2 / rho2))
        if g == "g":
            appro = F(2) * (F(1) + temp) + F(2)
            act = _tanh(z)
            der = F(1) - act ** 2
        else:
            appro = F(0.5) * (F(1) + temp) + F(0.125)
            act = _sig(z)
            der = act * (F(1) - act)
        form1 = self.x[:, t, :] @ self.W[g] + self._hprev(t) @ self.U[g] - lam1 / rho1
        form2 = rho2 * (act - out + lam2 / rho2) * der
        form3 = rho1 * form1 + F(0.5) * rho2 * appro * z - form2
        self.z[g][t] = (F(2) * form3 / (F(2) * rho1 + rho2 * appro)).astype(F)

    # admm_lstm.py:191-220 ----------------------------------------------------------------------------
    def update_f(self, t):
        ct, c_ = self.c[t], self._cprev(t)
        form1 = self.rp * (_sig(self.z["f"][t]) + self.lam_p["f"][t] / self.rp) + self.r9 * c_ * (ct - self.g["g"][t] * self.g["i"][t] + self.lam9[t] / self.r9)
        self.g["f"][t] = (form1 / (self.rp + self.r9 * c_ * c_)).astype(F)

    def update_i(self, t):
        ct, c_, gg = self.c[t], self._cprev(t), self.g["g"][t]
        form1 = self.rp * (_sig(self.z["i"][t]) + self.lam_p["i"][t] / self.rp) + self.r9 * gg * (ct - c_ * self.g["f"][t] + self.lam9[t] / self.r9)
        self.g["i"][t] = (form1 / (self.rp + self.r9 * gg * gg)).astype(F)

    def update_o(self, t):
        tc = _tanh(self.c[t])
        form1 = self.rp * (_sig(self.z["o"][t]) + self.lam_p["o"][t] / self.rp) + self.r10 * tc * (self.h[t] - self.lam10[t] / self.r10)
        self.g["o"][t] = (form1 / (self.rp + self.r10 * tc * tc)).astype(F)

    def update_g(self, t):
        ct, c_, i = self.c[t], self._cprev(t), self.g["i"][t]
        form1 = self.rp * (_tanh(self.z["g"][t]) + self.lam_p["g"][t] / self.rp) + self.r9 * i * (ct - c_ * self.g["f"][t] + self.lam9[t] / self.r9)
        self.g["g"][t] = (form1 / (self.rp + self.r9 * i * i)).astype(F)

    # admm_lstm.py:223-241 ----------------------------------------------------------------------------
    def update_c(self, t):
        f, i, o, g = (self.g[k][t] for k in ("f", "i", "o", "g"))
        ct, c_, h = self.c[t], self._cprev(t), self.h[t]
        temp = np.max(np.abs((h - self.lam10[t] / self.r10) * F(1) / o))
        appro_h = F(2) * (F(1) + temp) + F(2)
        form1 = self.r9 * (g * i + c_ * f - self.lam9[t] / self.r9)
        tc = _tanh(ct)
        form2 = self.r10 * (tc * o - h + self.lam10[t] / self.r10) * (F(1) - tc ** 2) * o
        qua_o = np.sum(o * o, dtype=F)
        form3 = F(0.5) * self.r10 * qua_o * ct * appro_h
        form4 = self.r9 + F(0.5) * self.r10 * qua_o * appro_h
        self.c[t] = ((form1 - form2 + form3) / form4).astype(F)

    # admm_lstm.py:244-259 ----------------------------------------------------------------------------
    def update_h(self, t):
        c, o, h = self.c[t], self.g["o"][t], self.h[t]
        form1 = self.r10 * (_tanh(c) * o + self.lam10[t] / self.r10)
        if t < self.T - 1:
            self.h[t] = (form1 / self.r10).astype(F)
            return
        theta = F(1)
        while True:
            form10 = -self.a + h @ self.Wy - self.lam11 / self.r11
            form11 = form10 @ self.Wy.T
            h1 = h - self.r11 * form11 / theta
            form12 = (h1 - h) * (h1 - h)
            func1 = F(0.5) * self.r11 * np.sum(form10 * form10, dtype=F) + self.r11 * np.sum(form11 * (h1 - h), dtype=F) + F(0.5) * theta * np.sum(form12, dtype=F)
            form20 = self.a - h1 @ self.Wy + self.lam11 / self.r11
            func2 = F(0.5) * self.r11 * np.sum(form20 * form20, dtype=F)
            if not (func2 > func1):
                break
            theta = theta * F(2)
        self.theta["h"] = float(theta)
        self.h[t] = ((form1 - self.r11 * form11 + theta * h) / (self.r10 + theta)).astype(F)

    # admm_lstm.py:262-311 ----------------------------------------------------------------------------
    def update_a(self):
        h = self.h[self.T - 1]
        temp1 = F(2) * self.y / F(self.n_norm) + self.r11 * (h @ self.Wy) - self.lam11
        self.a = (temp1 / F(2 / self.n_norm + float(self.r11))).astype(F)

    def update_duals(self, t):
        if t == self.T - 1:
            pass
        f, i, o, g = (self.g[k][t] for k in ("f", "i", "o", "g"))
        self.lam10[t] = self.lam10[t] + self.r10 * (_tanh(self.c[t]) * o - self.h[t])
        self.lam9[t] = self.lam9[t] + self.r9 * (self.c[t] - g * i - self._cprev(t) * f)
        for k in ("g", "o", "i", "f"):
            act = _tanh(self.z[k][t]) if k == "g" else _sig(self.z[k][t])
            self.lam_p[k][t] = self.lam_p[k][t] + self.rp * (act - self.g[k][t])
            lin = self.z[k][t] - self.x[:, t, :] @ self.W[k] - self._hprev(t) @ self.U[k]
            self.lam_s[k][t] = self.lam_s[k][t] + self.rs * lin

    # main.py:139-188 ---------------------------------------------------------------------------------
    def step(self):
        self.update_wy()
        for g in ORDER_W:
            self.update_weight(g, "x")
            self.update_weight(g, "h")
        for t in range(self.T):
            self.update_z("f", t)
            self.update_f(t)
            self.update_z("i", t)
            self.update_i(t)
            self.update_z("o", t)
            self.update_o(t)
            self.update_z("g", t)
            self.update_g(t)
            self.update_c(t)
            self.update_h(t)
            if t == self.T - 1:
                self.update_a()
                self.lam11 = (self.lam11 + self.r11 * (self.a - self.h[t] @ self.Wy)).astype(F)
            self.update_duals(t)

    def loss(self, x=None, y=None):
        x = self.x if x is None else np.asarray(x, F)
        y = self.y if y is None else np.asarray(y, F)
        return float(np.mean(np.square(y - self.predict(x))))
