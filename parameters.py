"""parameters.py -- same import surface as the reference's parameters.py (demo.py:22, admm.py:9)."""
from admm_lstm_b200.parameters import example_parameter_dictionary, default_epoch  # noqa: F401

__all__ = ["example_parameter_dictionary", "default_epoch"]
