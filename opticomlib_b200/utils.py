"""The two helpers of ``opticomlib/utils.py`` the hot path uses: the LIFO wall-clock timer
(``tic``/``toc``, reference utils.py:293-341) that fills ``output.execution_time``."""
from __future__ import annotations

import time as _tm

_stack: list[float] = []


def tic() -> None:
    _stack.append(_tm.time())


def toc() -> float:
    if not _stack:
        raise Exception("toc() called without a matching tic()")
    return _tm.time() - _stack.pop()
