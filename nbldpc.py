"""Import shim: the package directory carries the repository's hyphenated name, which the ``import``
statement cannot spell.  ``import nbldpc`` gives the same module object."""
import importlib
import os
import sys

_root = os.path.dirname(os.path.abspath(__file__))
if _root not in sys.path:
    sys.path.insert(0, _root)
_pkg = importlib.import_module("ems-decoder-of-nb-ldpc-codes_b200")
sys.modules[__name__] = _pkg
