"""Loads the CUDA build of the engine (aeonflux_b200/csrc/libaeonflux_b200.so).  There is no CPU fallback: if the
shared library is missing or cannot be loaded the import of every product entry point fails loudly."""
import ctypes
import os

from ._binding import Binding

_HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(_HERE, "csrc", "libaeonflux_b200.so")
_binding = None


def load() -> Binding:
    global _binding
    if _binding is None:
        if not os.path.exists(SO_PATH):
            raise ImportError("aeonflux_b200: %s not built -- run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(nvcc, sm_100a).  There is no CPU fallback." % SO_PATH)
        _binding = Binding(ctypes.CDLL(SO_PATH))
    return _binding
