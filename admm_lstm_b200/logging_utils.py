"""Console/file logging with the reference's error semantics (reference: _global.py:117-200):
info/warning print a tagged line and append it to logs/ADMMRunningLogs.log under the CWD; error()
prints, logs and exits the process with `code`; log_assert(cond, msg) is error() on failure."""
from __future__ import annotations

import logging
import os
import sys
from datetime import datetime
from typing import Any, NoReturn

_COLORS = {"INFO": "\033[32m", "WARNING": "\033[33m", "ERROR": "\033[31m", "ASSERTION FAILURE": "\033[31m"}
_logger = None


def _file_logger():
    global _logger
    if _logger is None:
        _logger = logging.getLogger("admm_lstm_b200")
        _logger.setLevel(logging.DEBUG)
        _logger.propagate = False
        if os.environ.get("ADMM_LSTM_NO_LOGFILE") != "1":
            try:
                os.makedirs("logs", exist_ok=True)
                fh = logging.FileHandler(os.path.join("logs", "ADMMRunningLogs.log"))
                fh.setFormatter(logging.Formatter("%(asctime)s - %(name)s - %(levelname)s - %(message)s"))
                _logger.addHandler(fh)
            except OSError:
                pass
        if not _logger.handlers:
            _logger.addHandler(logging.NullHandler())
    return _logger


def _emit(tag: str, msg: Any) -> None:
    stamp = datetime.now().strftime("%H:%M:%S")
    color = _COLORS.get(tag, "") if sys.stdout.isatty() else ""
    reset = "\033[0m" if color else ""
    print(f"[{stamp}] {color}{tag}{reset}: {msg}")


def info(msg: Any = "", use_logger: bool = True) -> None:
    if use_logger:
        _file_logger().info(str(msg))
    _emit("INFO", msg)


def warning(msg: Any = "", use_logger: bool = True) -> None:
    if use_logger:
        _file_logger().warning(str(msg))
    _emit("WARNING", msg)


def error(msg: Any = "", code: int = 1, use_logger: bool = True, assertion: bool = False) -> NoReturn:
    if use_logger:
        _file_logger().error(str(msg))
    _emit("ASSERTION FAILURE" if assertion else "ERROR", msg)
    sys.exit(code)


def log_assert(condition: bool, msg: Any = "", code: int = 1) -> None:
    if not condition:
        error(msg, code, assertion=True)
