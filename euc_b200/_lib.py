"""Loads the CUDA back end (csrc/libeuc_b200.so).  There is no CPU fallback: if the library is missing or
no CUDA device is present, using the back end raises."""
import ctypes as C
import os

from . import abi

_LIB = None
# EUC_B200_LIB selects another build of the same library (kernel A/B experiments, tools/ab.sh); never a fallback
LIB_PATH = os.environ.get("EUC_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "csrc", "libeuc_b200.so")


class EucError(RuntimeError):
    def __init__(self, code, msg=""):
        self.code = code
        super().__init__(f"{abi.STATUS_NAMES.get(code, code)}: {msg}")


def load():
    global _LIB
    if _LIB is not None:
        return _LIB
    if not os.path.exists(LIB_PATH):
        raise ImportError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
                          "(euc_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in abi.SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    if lib.euc_abi_version() != abi.ABI_VERSION:
        raise ImportError("libeuc_b200.so ABI version mismatch")
    _LIB = lib
    return lib
