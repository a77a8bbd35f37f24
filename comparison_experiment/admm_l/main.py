"""Drop-in for the reference's comparison_experiment/admm_l/main.py: same names (`admm_l_demo`, `LSTM_L`), the
iteration runs on the B200 (admm_lstm_b200/admm_l.py).  comparison.py:174-178 imports it as `admm_l.main`."""
from admm_lstm_b200.admm_l import LSTM_L, ADMMLOptimizer, admm_l_demo  # noqa: F401
