"""admm.no_dual_y.py -- the "Fast ADMM-LSTM" variant (reference: admm.no_dual_y.py).

Like the reference's file of the same name it is not importable by name (the dot); use it the way
the reference is used: copy it over admm.py.  It differs from admm.py only in the default variant.
"""
from admm_lstm_b200.optimizer import ADMMBasedOptimizer as _Optimizer
from admm_lstm_b200.parameters import example_parameter_dictionary  # noqa: F401

variant = "no_dual_y"


class ADMMBasedOptimizer(_Optimizer):
    def __init__(self, model, training_samples, parameter_dictionary=None, verbose=True, **kwargs):
        kwargs.setdefault("variant", variant)
        super().__init__(model, training_samples, parameter_dictionary, verbose, **kwargs)
