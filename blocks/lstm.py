"""blocks/lstm.py -- module path of the pickled model class (SAVED_MODELS/*.pt reference the global
`blocks.lstm LSTM`); the implementation lives in admm_lstm_b200/lstm.py."""
from admm_lstm_b200.lstm import LSTM

LSTM.__module__ = "blocks.lstm"      # pickles written here stay loadable by the reference tree

__all__ = ["LSTM"]
