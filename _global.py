"""_global.py -- the slice of the reference's _global.py that its demo.py / comparison.py import
(device, info, warning, error, log_assert, global_dict); see admm_lstm_b200/logging_utils.py."""
import torch

from admm_lstm_b200.logging_utils import error, info, log_assert, warning  # noqa: F401


class GlobalDict:
    def __init__(self):
        self.contents = dict()

    def set(self, key, value):
        self.contents[key] = value

    def get(self, key):
        return self.contents[key]

    def keys(self):
        return self.contents.keys()

    __setitem__ = set
    __getitem__ = get


global_dict = GlobalDict()
global_dict.set("logger_filename", "logs/ADMMRunningLogs.log")
device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
