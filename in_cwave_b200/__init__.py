"""in_cwave_b200 -- the in_cwave signal chain (unpack -> Hilbert -> modulator graph -> 16/24-bit
render) as hand-written sm_100a kernels behind a C ABI (include/icw_b200.h).

Python here is plumbing around libicw_b200.so: spec dicts, ctypes structs, device buffers.
There is no CPU fallback: using the package without the built CUDA extension raises.
"""
from . import _abi, spec, synth  # noqa: F401
from ._abi import IcwError, lib  # noqa: F401
from .engine import Engine, Session  # noqa: F401
from .spec import config_c1, config_c2, config_c3, default_spec  # noqa: F401

__all__ = ["Engine", "Session", "IcwError", "lib", "spec", "synth", "default_spec",
           "config_c1", "config_c2", "config_c3"]
