"""Import shim for the UNMODIFIED reference (armando-palacio/opticomlib).

TEST INFRASTRUCTURE ONLY.  Nothing under ``opticomlib_b200/`` imports this.

The reference cannot be imported as shipped in this image because
``opticomlib/typing.py:13`` needs ``pympler`` and ``typing.py:21-22`` /
``utils.py:49`` / ``devices.py:33`` need ``matplotlib``; neither is installed
and neither is used by the FIBER / DBP / LPF / BPF arithmetic.  We inject empty
``MagicMock`` modules for those two packages and put the reference tree on
``sys.path``.  The reference sources are never copied or modified.

The reference tree only exists in the build container (``/root/reference``) or,
on a GPU box, where the driver may have installed it (``baseline/_ref``).  Code
that must run on the GPU box uses the committed fixtures in ``tests/golden``
instead of this module.
"""
from __future__ import annotations

import os
import sys
from unittest.mock import MagicMock

_HERE = os.path.dirname(os.path.abspath(__file__))
_CANDIDATES = (
    os.path.join(os.path.dirname(_HERE), "baseline", "_ref"),
    "/root/reference",
)

_MOCKED = (
    "matplotlib",
    "matplotlib.pyplot",
    "matplotlib.collections",
    "matplotlib.animation",
    "matplotlib.widgets",
    "pympler",
    "pympler.asizeof",
)


def reference_root() -> str | None:
    """Directory that contains the reference ``opticomlib`` package, or None."""
    for root in _CANDIDATES:
        if os.path.isfile(os.path.join(root, "opticomlib", "devices.py")):
            return root
    return None


def import_reference():
    """Return the reference ``opticomlib`` package (devices/typing imported).

    Raises ImportError when the reference tree is not present.
    """
    root = reference_root()
    if root is None:
        raise ImportError("reference opticomlib not found (looked in %s)" % (_CANDIDATES,))
    for name in _MOCKED:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = MagicMock(name=name)
    if root not in sys.path:
        sys.path.insert(0, root)
    import opticomlib  # noqa: F401
    import opticomlib.devices  # noqa: F401
    import opticomlib.typing  # noqa: F401

    return sys.modules["opticomlib"]
