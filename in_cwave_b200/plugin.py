"""ctypes veneer over libicw_plugin.so: the reference's transcode entry points (include/icw_plugin.h)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from . import _abi, spec as _spec

PLUGIN_PATH = Path(__file__).resolve().parent / "libicw_plugin.so"
EXPORTS = ["winampGetExtendedRead_open", "winampGetExtendedRead_getData", "winampGetExtendedRead_setTime",
           "winampGetExtendedRead_close", "icwp_configure", "icwp_reset", "icwp_stats", "icwp_probe",
           "icwp_load_config", "icwp_save_config", "icwp_check_cwave", "icwp_io_stats", "winampGetInModule2"]


class Options(C.Structure):
    _fields_ = [("sec_align", C.c_uint), ("fade_in_ms", C.c_uint), ("fade_out_ms", C.c_uint),
                ("clr_nframe_trk", C.c_int), ("clr_hilb_trk", C.c_int), ("device", C.c_int),
                ("readahead_frames", C.c_int64)]


class IoStats(C.Structure):
    _fields_ = [("read_s", C.c_double), ("wait_s", C.c_double), ("gpu_s", C.c_double), ("read_bytes", C.c_uint64),
                ("frames", C.c_uint64), ("blocks_prefetched", C.c_uint64), ("blocks_sync", C.c_uint64),
                ("resettles", C.c_uint64)]


class FileInfo(C.Structure):
    _fields_ = [("fmt", C.c_int), ("n_channels", C.c_int), ("sample_rate", C.c_uint),
                ("n_samples", C.c_int64), ("n_tail", C.c_int64), ("offset_data", C.c_int64),
                ("n_fade_in", C.c_int64), ("n_fade_out", C.c_int64)]


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        _abi.lib()                                   # the CUDA library first: the plugin links against it
        if not PLUGIN_PATH.exists():
            raise RuntimeError(f"{PLUGIN_PATH} is missing: run in_cwave_b200/build.py")
        L = C.CDLL(str(PLUGIN_PATH))
        ip = C.POINTER(C.c_int)
        L.winampGetExtendedRead_open.argtypes = [C.c_char_p, ip, ip, ip, ip]
        L.winampGetExtendedRead_open.restype = C.c_ssize_t
        L.winampGetExtendedRead_getData.argtypes = [C.c_ssize_t, C.c_void_p, C.c_int, ip]
        L.winampGetExtendedRead_getData.restype = C.c_ssize_t
        L.winampGetExtendedRead_setTime.argtypes = [C.c_ssize_t, C.c_int]
        L.winampGetExtendedRead_close.argtypes = [C.c_ssize_t]
        L.winampGetExtendedRead_close.restype = None
        L.icwp_configure.argtypes = [C.POINTER(_abi.ChainSpecC), C.POINTER(Options)]
        L.icwp_reset.restype = None
        L.icwp_stats.argtypes = [C.POINTER(_abi.Stats)]
        L.icwp_io_stats.argtypes = [C.POINTER(IoStats), C.c_int]
        L.icwp_probe.argtypes = [C.c_char_p, C.POINTER(Options), C.POINTER(FileInfo)]
        L.icwp_check_cwave.argtypes = [C.c_char_p, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), ip]
        L.icwp_load_config.argtypes = [C.c_char_p, C.POINTER(_abi.ChainSpecC), C.POINTER(Options)]
        L.icwp_save_config.argtypes = [C.c_char_p, C.POINTER(_abi.ChainSpecC), C.POINTER(Options)]
        _lib = L
    return _lib


def configure(spec: dict | None, **opt) -> None:
    o = Options(**opt)
    rc = lib().icwp_configure(_spec.to_c(spec) if spec is not None else None, C.byref(o))
    _abi.check(rc)


def _node_tuple(n):
    return (n.mode, n.inputs_mask, n.xch_mode, n.l_iq_invert, n.r_iq_invert, n.l_gain, n.r_gain,
            n.n_out if n.mode != 0 else 0, n.l_tout, n.r_tout, n.l_on, n.r_on, tuple(n.l_p), tuple(n.r_p))


def load_config(path: str):
    """The reference's config file -> (accepted, ChainSpecC, Options, nodes in execution order as tuples)."""
    sp, o = _abi.ChainSpecC(), Options()
    ok = lib().icwp_load_config(str(path).encode(), C.byref(sp), C.byref(o))
    return bool(ok), sp, o, [_node_tuple(sp.nodes[i]) for i in range(sp.n_nodes)]


def save_config(path: str, spec: dict, **opt) -> bool:
    o = Options(**opt)
    return bool(lib().icwp_save_config(str(path).encode(), _spec.to_c(spec), C.byref(o)))


def check_cwave(path: str):
    """(readable, crc_calculated, crc_in_header, header_has_crc) -- the reference's file-info CRC check."""
    calc, filec, has = C.c_uint32(0), C.c_uint32(0), C.c_int(0)
    ok = lib().icwp_check_cwave(str(path).encode(), C.byref(calc), C.byref(filec), C.byref(has))
    return bool(ok), int(calc.value), int(filec.value), bool(has.value)


def io_stats(reset: bool = False) -> dict:
    """How the overlapped reader did since the last reset (include/icw_plugin.h icwp_iostats)."""
    st = IoStats()
    lib().icwp_io_stats(C.byref(st), int(reset))
    return {k: getattr(st, k) for k, _ in IoStats._fields_}


def probe(path: str, **opt):
    o, fi = Options(**opt), FileInfo()
    ok = lib().icwp_probe(str(path).encode(), C.byref(o), C.byref(fi))
    return fi if ok else None


def transcode(path: str, chunk: int = 65536, seek_ms: int | None = None):
    """winampGetExtendedRead_open -> getData loop -> close.  Returns (pcm bytes, (size, bps, nch, srate)) or None."""
    L = lib()
    size, bps, nch, srate = C.c_int(), C.c_int(), C.c_int(), C.c_int()
    h = L.winampGetExtendedRead_open(str(path).encode(), C.byref(size), C.byref(bps), C.byref(nch), C.byref(srate))
    if not h:
        return None
    if seek_ms is not None and not L.winampGetExtendedRead_setTime(h, seek_ms):
        L.winampGetExtendedRead_close(h)
        raise RuntimeError("setTime failed")
    out = bytearray()
    buf = (C.c_char * chunk)()
    kill = C.c_int(0)
    while True:
        got = L.winampGetExtendedRead_getData(h, buf, chunk, C.byref(kill))
        if got <= 0:
            break
        out += buf.raw[:got]
    L.winampGetExtendedRead_close(h)
    return np.frombuffer(bytes(out), dtype=np.uint8), (size.value, bps.value, nch.value, srate.value)
