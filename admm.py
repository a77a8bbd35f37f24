"""admm.py -- drop-in replacement for the reference's admm.py (same import surface:
`from admm import ADMMBasedOptimizer, example_parameter_dictionary`, demo.py:23).

The work is done by admm_lstm_b200 (hand-written sm_100a kernels behind a C ABI).  Variant
selection mirrors how the reference does it -- by which file is named admm.py: this file is the
shipped admm.py behaviour (with_dual_y = False); copy `admm.no_dual_y.py` over it, set
`admm.variant = 'no_dual_y'`, or export ADMM_LSTM_VARIANT=no_dual_y for the "Fast" variant.
"""
import os

from admm_lstm_b200.optimizer import ADMMBasedOptimizer as _Optimizer
from admm_lstm_b200.parameters import example_parameter_dictionary  # noqa: F401  (re-export, reference admm.py:9)

variant = os.environ.get("ADMM_LSTM_VARIANT", "admm")
with_dual_y = False          # reference admm.py:12


class ADMMBasedOptimizer(_Optimizer):
    def __init__(self, model, training_samples, parameter_dictionary=None, verbose=True, **kwargs):
        kwargs.setdefault("variant", variant)
        kwargs.setdefault("with_dual_y", with_dual_y)
        super().__init__(model, training_samples, parameter_dictionary, verbose, **kwargs)
