"""admm_lstm_b200 -- B200-native implementation of the ADMM-LSTM per-iteration sweep.

Public surface (mirrors the reference's admm.py): ADMMBasedOptimizer, example_parameter_dictionary.
All N*T-sized work runs in hand-written sm_100a CUDA kernels behind the C ABI declared in
include/admm_lstm_b200.h; there is no CPU fallback.
"""
from .parameters import example_parameter_dictionary, default_epoch  # noqa: F401


def __getattr__(name):
    if name == "ADMMBasedOptimizer":
        from .optimizer import ADMMBasedOptimizer
        return ADMMBasedOptimizer
    if name == "LSTM":
        from .lstm import LSTM
        return LSTM
    raise AttributeError(name)
