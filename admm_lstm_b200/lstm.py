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

    # -- accessors used by the optimizer (reference blocks/lstm.py:31-41) --------------------------
    def get_weight(self, map_from: str, map_to: str) -> torch.Tensor:
        return getattr(self, f"{map_from}2{map_to}").detach().clone()

    def set_weight(self, map_from: str, map_to: str, value: torch.Tensor) -> None:
        setattr(self, f"{map_from}2{map_to}", nn.Parameter(value.detach().clone()))

    def get_wy(self) -> torch.Tensor:
        return self.out.detach().clone()

    def set_wy(self, value: torch.Tensor) -> None:
        setattr(self, "out", nn.Parameter(value))

    # -- forward -----------------------------------------------------------------------------------
    def _stacked(self):
        wx = torch.cat([getattr(self, f"x2{g}") for g in _GATES], dim=1)      # [D, 4H]
        wh = torch.cat([getattr(self, f"h2{g}") for g in _GATES], dim=1)      # [H, 4H]
        return wx, wh

    def forward(self, x: torch.Tensor, c: Optional[torch.Tensor] = None, h: Optional[torch.Tensor] = None):
        if self.with_grad:
            return self.grad_forward(x, c, h)
        if x.is_cuda and c is None and h is None and not torch.is_grad_enabled():
            from .optimizer import predict_cuda
            return predict_cuda(self, x)
        return self.init_gate_variables(x, c, h)["a"]

    def grad_forward(self, x: torch.Tensor, c: Optional[torch.Tensor], h: Optional[torch.Tensor]) -> torch.Tensor:
        assert x.size(2) == self.input_size
        batch, seq_len, _ = x.size()
        hs = self.hidden_size
        c = x.new_zeros(batch, hs) if c is None else c
        h = x.new_zeros(batch, hs) if h is None else h
        wx, wh = self._stacked()
        for t in range(seq_len):
            zi, zf, zg, zo = (x[:, t, :] @ wx + h @ wh).split(hs, dim=1)
            c = torch.sigmoid(zf) * c + torch.sigmoid(zi) * torch.tanh(zg)
            h = torch.sigmoid(zo) * torch.tanh(c)
        return h @ self.out

    def init_gate_variables(self, x: torch.Tensor, c: Optional[torch.Tensor] = None,
                            h: Optional[torch.Tensor] = None) -> Dict[str, torch.Tensor]:
        """State tensors [N, T+1, H] with slot 0 zero, plus a = h_T @ out (reference :65-88)."""
        assert x.size(2) == self.input_size
        batch, seq_len, _ = x.size()
        hs = self.hidden_size
        shape = (batch, seq_len + 1, hs)
        st = {k: x.new_zeros(shape) for k in ("i", "f", "g", "o")}
        st["c"] = x.new_zeros(shape) if c is None else c
        st["h"] = x.new_zeros(shape) if h is None else h
        wx, wh = self._stacked()
        for t in range(1, seq_len + 1):
            zi, zf, zg, zo = (x[:, t - 1, :] @ wx + st["h"][:, t - 1, :] @ wh).split(hs, dim=1)
            st["i"][:, t, :] = torch.sigmoid(zi)
            st["f"][:, t, :] = torch.sigmoid(zf)
            st["g"][:, t, :] = torch.tanh(zg)
            st["o"][:, t, :] = torch.sigmoid(zo)
            st["c"][:, t, :] = st["f"][:, t, :] * st["c"][:, t - 1, :] + st["i"][:, t, :] * st["g"][:, t, :]
            st["h"][:, t, :] = st["o"][:, t, :] * torch.tanh(st["c"][:, t, :])
        st["a"] = st["h"][:, seq_len, :] @ self.out
        return st
