"""Hyper-parameter table with the reference's keys and values (reference: parameters.py:9-91).

Layout of one entry: {'rho': {i,f,g,o,c,h,y}, 'beta': {wi,vi,wf,vf,wg,vg,wo,vo,wy}} -- exactly what
ADMMBasedOptimizer(parameter_dictionary=...) consumes (reference admm.py:110-162).
"""
from typing import Dict

__all__ = ["example_parameter_dictionary", "default_epoch"]

default_epoch = 100

_RHO_KEYS = ("i", "f", "g", "o", "c", "h", "y")
_BETA_KEYS = ("wi", "vi", "wf", "vf", "wg", "vg", "wo", "vo", "wy")


def _entry(gate_rho, c, h, y, beta, beta_wy=None):
    rho = dict(zip(_RHO_KEYS, (gate_rho,) * 4 + (c, h, y)))
    b = {k: beta for k in _BETA_KEYS}
    if beta_wy is not None:
        b["wy"] = beta_wy
    return {"rho": rho, "beta": b}


# dataset -> (rho_{i,f,g,o}, rho_c, rho_h, rho_y, beta[, beta_wy]); values as shipped in parameters.py:11-91
_TABLE = {
    "GoogleStock": (1., 0.008, 0.00045, 0.0000562, 8e-7),
    "GEFCOM2012": (1, 0.1, 0.01, 0.01, 8e-7),
    "YahooFinance": (1, 0.1, 0.02, 0.01, 1e-8),
    "MNISTDataset": (1, 0.012, 0.0012, 0.00005, 1, 10),
    "UCF101": (.1, 0.008, 0.0001, 0.000001, 1e-9),
    "HAR": (1.5, 0.005, 8e-04, 4e-04, 8e-7),
    "PTB": (.8, 5e-4, 5e-4, 1e-5, 8e-7),
    "DNA1": (1., 0.001, 0.03, 0.002, 8e-9),
    "SMSSpam": (1.0, 0.01, 0.001, 4e-05, 8e-9),
}

example_parameter_dictionary: Dict[str, Dict[str, Dict[str, float]]] = {
    name: _entry(*row) for name, row in _TABLE.items()
}

# parameters.py:24 -- the "without dual y" rho set the Fast variant needs on GEFCOM-format data
# (the shipped GEFCOM2012 values diverge to NaN at iteration 3 with admm.no_dual_y.py).
gefcom2012_without_dual_y = _entry(1, 0.1, 0.001, 0.0001, 8e-7)
