"""Drop-in for the reference's comparison_experiment/admm_l/main.py: same names (`admm_l_demo`, `LSTM_L`), the
iteration runs on the B200 (admm_lstm_b200/admm_l.py).  comparison.py:174-178 imports it as `admm_l.main`."""
from admm_lstm_b200.admm_l import LSTM_L, ADMMLOptimizer, admm_l_demo  # noqa: F401

if __name__ == "admm_l.main":
    # imported the way comparison.py:175 imports it: SAVED_MODELS/ADMM-LSTM-L.pt then carries the reference's pickle
    # global `admm_l.main LSTM_L` and loads in the reference tree (visualization.py:47-54)
    LSTM_L.__module__ = __name__
