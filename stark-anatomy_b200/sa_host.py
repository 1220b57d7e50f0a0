"""sa_host -- resolves the host-side value types of the drop-in.

The reference's caller (code/fast_stark.py) and the drop-in must share ONE
``algebra`` / ``univariate`` / ``merkle`` / ``ip`` module each: FieldElement
identity, ``Field`` instances and the pickled module names all flow through
them.  So:

  * if those modules are importable (the reference's code/ directory is on
    sys.path, the normal deployment), they are used as they are;
  * otherwise (the GPU test box) the independent mirrors in ./hostmirror are put
    at the END of sys.path and imported under the same names.
"""
import importlib
import os
import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
_MIRROR = os.path.join(_HERE, "hostmirror")


def _import(name):
    try:
        return importlib.import_module(name)
    except ImportError:
        if _MIRROR not in sys.path:
            sys.path.append(_MIRROR)
        return importlib.import_module(name)


algebra = _import("algebra")
univariate = _import("univariate")
merkle = _import("merkle")
ip = _import("ip")
