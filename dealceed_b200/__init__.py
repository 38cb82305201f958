"""Importable alias of the package directory `deal-and-ceed-on-gpu_b200/` (whose
name is not a Python identifier).  `import dealceed_b200.bindings` resolves to
`deal-and-ceed-on-gpu_b200/bindings.py`."""
import os

__path__.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                "deal-and-ceed-on-gpu_b200"))
from .bindings import *  # noqa: F401,F403,E402
from . import bindings  # noqa: E402
