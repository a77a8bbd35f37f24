"""Helpers for the -m gpu parity tests: build the CUDA optimizer from numpy fixtures and move state
between the reference layout ([N,T+1,H]) and the device layout ([T+1][H][ldn])."""
import numpy as np
import torch

from admm_lstm_b200.lstm import LSTM
from admm_lstm_b200.optimizer import ADMMBasedOptimizer

from helpers import WKEYS


def make_model(weights):
    d, h = weights["x2i"].shape
    o = weights["out"].shape[1]
    model = LSTM(d, h, o)
    with torch.no_grad():
        for k in WKEYS:
            getattr(model, k).copy_(torch.from_numpy(np.asarray(weights[k])))
    return model


def make_opt(weights, x, y, params, variant="admm", **kw):
    model = make_model(weights)
    opt = ADMMBasedOptimizer(model, (torch.from_numpy(np.asarray(x)), torch.from_numpy(np.asarray(y))), params,
                             verbose=False, variant=variant, **kw)
    return model, opt


def load_state(opt, state):
    n, T = opt.n_local, opt.seq_len
    dev = opt.device
    for k in ("i", "f", "g", "o", "c", "h"):
        opt._state[k][:, :, :n] = torch.from_numpy(state["gates"][k]).to(dev).permute(1, 2, 0)
    for k in ("i", "f", "g", "o", "c"):
        opt._dual[k][:, :, :n] = torch.from_numpy(state["duals"][k]).to(dev).permute(1, 2, 0)
    opt._dual_h[:, :n] = torch.from_numpy(state["duals"]["h"][:, T, :]).to(dev).t()
    opt._a[:, :n] = torch.from_numpy(state["gates"]["a"]).to(dev).t()
    opt._dual_y[:, :n] = torch.from_numpy(state["duals"]["y"]).to(dev).t()
    opt.state_changed()


def weights_of(opt):
    out = {}
    for q, g in enumerate("ifgo"):
        out["x2" + g] = opt._wx[q].cpu().numpy()
        out["h2" + g] = opt._wh[q].cpu().numpy()
    out["out"] = opt._wy.cpu().numpy()
    return out


def np_state(opt):
    gates = {k: opt.gates[k].cpu().numpy() for k in ("i", "f", "g", "o", "c", "h", "a")}
    duals = {k: opt.duals[k].cpu().numpy() for k in ("i", "f", "g", "o", "c", "h", "y")}
    return gates, duals
