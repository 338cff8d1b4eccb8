"""b200bo — host side of the B200-native engine (PyTorch for memory/streams, ctypes for the
C ABI in include/b200bo.h).  There is no CPU fallback: importing ``b200bo.ops`` without the
built ``lib/libb200bo.so`` raises, and every op requires CUDA tensors."""
from . import _lib  # noqa: F401

__all__ = ["_lib"]
