"""LSTM-Linear cell with the reference's public surface (reference: blocks/lstm.py:11-88).

Same class name, constructor, attribute and parameter names (x2{i,f,g,o} [D,H], h2{i,f,g,o} [H,H],
out [H,O], no biases, Xavier-normal init, registered in that order) so that whole-module pickles
written by the reference (SAVED_MODELS/*.pt, global `blocks.lstm LSTM`) load into it and pickles
written here load back into the reference.  The arithmetic is written independently: one fused gate
projection per timestep instead of eight matmuls, and, for inference on a CUDA tensor, the
library's forward kernel (admm_predict).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
from torch import nn

_GATES = ("i", "f", "g", "o")


class LSTM(nn.Module):
    def __init__(self, input_size: int, hidden_size: int, output_size: int, with_grad: bool = False) -> None:
        super().__init__()
        self.input_size, self.hidden_size, self.output_size = input_size, hidden_size, output_size
        self.init_parameters()
        self.sigmoid, self.tanh = nn.Sigmoid(), nn.Tanh()
        self.with_grad = with_grad

    def init_parameters(self) -> None:
        # registration order x2i,h2i,x2f,h2f,x2g,h2g,x2o,h2o,out and one randn per tensor, then
        # Xavier-normal over all of them: reproduces the reference's RNG stream for a given seed.
        for gate in _GATES:
            self.register_parameter(f"x2{gate}", nn.Parameter(torch.randn(self.input_size, self.hidden_size)))
            self.register_parameter(f"h2{gate}", nn.Parameter(torch.randn(self.hidden_size, self.hidden_size)))
        self.register_parameter("out", nn.Parameter(torch.randn(self.hidden_size, self.output_size)))
        for param in self.parameters():
            nn.init.xavier_normal_(param)

    # -- weight accessors the optimizer's drop-in surface exposes (reference blocks/lstm.py:31-41) ----
    def _param(self, name: str) -> nn.Parameter:
        return self._parameters[name]

    def get_weight(self, map_from: str, map_to: str) -> torch.Tensor:
        return self._param(map_from + "2" + map_to).data.clone()

    def set_weight(self, map_from: str, map_to: str, value: torch.Tensor) -> None:
        self._parameters[map_from + "2" + map_to] = nn.Parameter(value.data.clone())

    def get_wy(self) -> torch.Tensor:
        return self._param("out").data.clone()

    def set_wy(self, value: torch.Tensor) -> None:
        self._parameters["out"] = nn.Parameter(value)      # no copy, like the reference

    # -- recurrence: ONE fused [D+H, 4H] projection per timestep ---------------------------------------
    def _stacked(self):
        wx = torch.cat([self._param("x2" + k) for k in _GATES], dim=1)      # [D, 4H]
        wh = torch.cat([self._param("h2" + k) for k in _GATES], dim=1)      # [H, 4H]
        return wx, wh

    def _unroll(self, x: torch.Tensor, c_prev: torch.Tensor, h_prev: torch.Tensor):
        """Yields (t, i, f, g, o, c, h) for t = 1..T starting from (c_prev, h_prev), each [N, H]."""
        if x.dim() != 3 or x.shape[2] != self.input_size:
            raise AssertionError(f"expected x of shape [N, T, {self.input_size}], got {tuple(x.shape)}")
        wx, wh = self._stacked()
        width = self.hidden_size
        for t in range(1, x.shape[1] + 1):
            z = x[:, t - 1] @ wx + h_prev @ wh
            gate_i = torch.sigmoid(z[:, :width])
            gate_f = torch.sigmoid(z[:, width:2 * width])
            gate_g = torch.tanh(z[:, 2 * width:3 * width])
            gate_o = torch.sigmoid(z[:, 3 * width:])
            c_prev = gate_f * c_prev + gate_i * gate_g
            h_prev = gate_o * torch.tanh(c_prev)
            yield t, gate_i, gate_f, gate_g, gate_o, c_prev, h_prev

    def forward(self, x: torch.Tensor, c: Optional[torch.Tensor] = None, h: Optional[torch.Tensor] = None):
        if self.with_grad:
            return self.grad_forward(x, c, h)
        # demo.py:341-342 evaluates `loss_func(model(train_x), train_y)` with autograd ENABLED on a model whose parameters an
        # ADMM optimizer owns (requires_grad False): nothing can need a graph then, so the library's forward kernel runs
        # (chunked over samples; no [N,T+1,H] temporaries).  Eager torch only when a gradient could actually flow.
        if x.is_cuda and c is None and h is None and not (
                torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters()))):
            from .optimizer import predict_cuda
            return predict_cuda(self, x)
        return self.init_gate_variables(x, c, h)["a"]

    def grad_forward(self, x: torch.Tensor, c: Optional[torch.Tensor], h: Optional[torch.Tensor]) -> torch.Tensor:
        """Differentiable prediction from [N, H] initial states (reference :48-63)."""
        zeros = x.new_zeros(x.shape[0], self.hidden_size)
        last = zeros if h is None else h
        for *_, last in self._unroll(x, zeros if c is None else c, last):
            pass
        return last @ self._param("out")

    def init_gate_variables(self, x: torch.Tensor, c: Optional[torch.Tensor] = None,
                            h: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """State tensors [N, T+1, H] with slot 0 zero, plus a = h_T @ out (reference :65-88).

        Caller-provided c / h ([N, T+1, H]) are filled in place and their slot 0 is the initial state.
        """
        full = (x.shape[0], x.shape[1] + 1, self.hidden_size)
        names = ("i", "f", "g", "o", "c", "h")
        st = {k: x.new_zeros(full) for k in names}
        if c is not None:
            st["c"] = c
        if h is not None:
            st["h"] = h
        for t, *values in self._unroll(x, st["c"][:, 0], st["h"][:, 0]):
            for k, v in zip(names, values):
                st[k][:, t] = v
        st["a"] = st["h"][:, -1] @ self._param("out")
        return st
