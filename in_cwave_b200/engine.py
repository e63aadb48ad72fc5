"""Engine / Session: thin object layer over the C ABI (include/icw_b200.h).

PyTorch is only plumbing here: device buffers are torch tensors whose data_ptr() crosses the C
boundary; every sample is computed by the kernels in libicw_b200.so.
"""
from __future__ import annotations

import ctypes as C
import weakref

import numpy as np

from . import _abi, spec as _spec


class Engine:
    """One GPU (reference has no equivalent: it is a single-threaded CPU plugin)."""

    def __init__(self, device: int = 0):
        self._h = C.c_void_p()
        _abi.check(_abi.lib().icw_engine_create(int(device), C.byref(self._h)))
        self.device = int(device)
        self._sessions = weakref.WeakSet()

    def close(self) -> None:
        for ses in list(self._sessions):
            ses.close()
        if self._h:
            _abi.lib().icw_engine_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def session(self, spec: dict, n_streams: int = 1) -> "Session":
        return Session(self, spec, n_streams)

    # ---- leaves -----------------------------------------------------------------------------
    def hilbert(self, x, filter_no=1, is_kahan=1, is_reject=1, mode="exact", states=None):
        """hq_rp_process over independent channels.  x: torch float64 CUDA tensor [n_chan, n]."""
        import torch
        assert x.is_cuda and x.dtype == torch.float64 and x.dim() == 2
        x = x.contiguous()
        n_chan, n = x.shape
        out = torch.empty((n_chan, n, 2), dtype=torch.float64, device=x.device)
        st = (_abi.StreamState * n_chan)()
        if states is not None:
            for i in range(n_chan):
                C.memmove(C.byref(st[i]), C.byref(states[i]), C.sizeof(_abi.StreamState))
        _abi.check(_abi.lib().icw_hilbert_device(self._h, filter_no, is_kahan, is_reject,
                                                 _abi.HILBERT[mode] if isinstance(mode, str) else mode,
                                                 n_chan, n, x.data_ptr(), out.data_ptr(), st))
        return out, st

    def mt_words(self, seed: int, skip: int, n: int):
        import torch
        out = torch.empty(n, dtype=torch.int32, device=f"cuda:{self.device}")
        _abi.check(_abi.lib().icw_mt_words_device(self._h, seed, skip, n, out.data_ptr()))
        return out.cpu().numpy().view(np.uint32)

    def crc32(self, data) -> int:
        """CRC-32 of a CUDA uint8 tensor (the reference's CWAVE data check, src/crc32.c:55-108)."""
        out = C.c_uint32(0)
        _abi.check(_abi.lib().icw_crc32_device(self._h, data.data_ptr(), data.numel() * data.element_size(), C.byref(out)))
        return int(out.value)

    def debug_sincos(self, x):
        """x: torch float64 CUDA tensor -> [n, 4] = (sincos_2pi sin, cos, libdevice sin, cos)."""
        import torch
        out = torch.empty((x.numel(), 4), dtype=torch.float64, device=x.device)
        _abi.check(_abi.lib().icw_debug_sincos_device(self._h, x.numel(), x.data_ptr(), out.data_ptr()))
        return out.cpu().numpy()

    def debug_phase(self, spec: dict, n0: int, n: int, freq_hz: float):
        import torch
        out = torch.empty((n, 2), dtype=torch.float64, device=f"cuda:{self.device}")
        _abi.check(_abi.lib().icw_debug_phase_device(self._h, _spec.to_c(spec), n0, n, float(freq_hz), out.data_ptr()))
        return out.cpu().numpy()


class Session:
    """K independent streams with one chain spec: K fresh MOD_CONTEXTs (reference src/in_cwave.c:46-80)."""

    def __init__(self, engine: Engine, spec: dict, n_streams: int = 1):
        self.engine = engine
        self.spec = dict(spec)
        self.n_streams = int(n_streams)
        self._c = _spec.to_c(spec)
        self._h = C.c_void_p()
        _abi.check(_abi.lib().icw_session_create(engine._h, self._c, self.n_streams, C.byref(self._h)))
        self.frame_bytes = _abi.lib().icw_frame_bytes(self._c)
        self.out_frame_bytes = _abi.lib().icw_out_frame_bytes(self._c)
        self._taps = None
        engine._sessions.add(self)

    def close(self) -> None:
        if self._h:
            _abi.lib().icw_session_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_spec(self, spec: dict) -> None:
        c = _spec.to_c(spec)
        _abi.check(_abi.lib().icw_session_set_spec(self._h, c))
        self.spec, self._c = dict(spec), c
        self.frame_bytes = _abi.lib().icw_frame_bytes(c)
        self.out_frame_bytes = _abi.lib().icw_out_frame_bytes(c)

    def reset(self, what: int = _abi.RESET_ALL) -> None:
        _abi.check(_abi.lib().icw_session_reset(self._h, what))

    def get_state(self, k: int = 0) -> _abi.StreamState:
        st = _abi.StreamState()
        _abi.check(_abi.lib().icw_session_get_state(self._h, k, C.byref(st)))
        return st

    def set_state(self, k: int, st: _abi.StreamState) -> None:
        _abi.check(_abi.lib().icw_session_set_state(self._h, k, C.byref(st)))

    def stats(self) -> dict:
        s = _abi.Stats()
        _abi.check(_abi.lib().icw_session_stats(self._h, C.byref(s)))
        return dict(clips=(int(s.clips[0]), int(s.clips[1])), peak_db=(s.peak_db[0], s.peak_db[1]),
                    hb_rejects=int(s.hb_rejects), mt_redraws=int(s.mt_redraws),
                    kernel_launches=int(s.kernel_launches))

    def fp_stats(self, k: int = 0):
        """FP exception counters of stream k: [hilbert L, hilbert R, render L, render R] x
        [total, snan, qnan, ninf, nden, pden, pinf] (reference fecs_getcnts)."""
        out = ((C.c_uint32 * 7) * 4)()
        _abi.check(_abi.lib().icw_session_fp_stats(self._h, int(k), out))
        return [list(r) for r in out]

    def profile(self, on: bool = True) -> None:
        _abi.check(_abi.lib().icw_session_profile(self._h, int(on)))

    def profile_read(self, reset: bool = True) -> dict:
        """Device time per kernel class from CUDA events on the launching stream (syncs)."""
        p = _abi.Profile()
        _abi.check(_abi.lib().icw_session_profile_read(self._h, C.byref(p), int(reset)))
        return {name: dict(ms=p.ms[i], launches=int(p.launches[i])) for i, name in enumerate(_abi.K_NAMES)}

    def sync(self) -> None:
        _abi.check(_abi.lib().icw_session_sync(self._h))

    # ---- the hot call -------------------------------------------------------------------------
    def process_host(self, raw: np.ndarray) -> np.ndarray:
        """raw: uint8 [n_streams, n*frame_bytes] (or 1-D for one stream) in host memory -> PCM bytes."""
        raw = np.ascontiguousarray(raw, dtype=np.uint8)
        if raw.ndim == 1:
            raw = raw.reshape(1, -1)
        assert raw.shape[0] == self.n_streams
        n = raw.shape[1] // self.frame_bytes
        out = np.empty((self.n_streams, n * self.out_frame_bytes), dtype=np.uint8)
        _abi.check(_abi.lib().icw_session_process_host(
            self._h, n, raw.ctypes.data, raw.strides[0], out.ctypes.data, out.strides[0]))
        return out

    def process_host_into(self, raw_ptr: int, in_stride: int, n: int, out_ptr: int, out_stride: int) -> None:
        _abi.check(_abi.lib().icw_session_process_host(self._h, n, raw_ptr, in_stride, out_ptr, out_stride))

    def process_device(self, d_in, n: int, d_out, in_stride: int | None = None,
                       out_stride: int | None = None, stream=None) -> None:
        """d_in / d_out: torch uint8 CUDA tensors (or raw device pointers)."""
        pin = d_in.data_ptr() if hasattr(d_in, "data_ptr") else int(d_in)
        pout = d_out.data_ptr() if hasattr(d_out, "data_ptr") else int(d_out)
        if in_stride is None:
            in_stride = d_in.stride(0) * d_in.element_size() if hasattr(d_in, "stride") and d_in.dim() > 1 else n * self.frame_bytes
        if out_stride is None:
            out_stride = d_out.stride(0) * d_out.element_size() if hasattr(d_out, "stride") and d_out.dim() > 1 else n * self.out_frame_bytes
        # torch's default stream has handle 0, which this ABI reads as "engine stream": name the
        # legacy default stream explicitly (cudaStreamLegacy == 0x1)
        cs = 0 if stream is None else (int(stream) or 1)
        _abi.check(_abi.lib().icw_session_process_device(self._h, n, pin, in_stride, pout, out_stride, cs))

    def enable_taps(self, n_frames: int):
        """Test aid: capture the whole bus and the master output of the next call (device tensors)."""
        import torch
        dev = f"cuda:{self.engine.device}"
        bus = torch.zeros((self.n_streams, n_frames, _abi.N_PLUGS, 4), dtype=torch.float64, device=dev)
        lr = torch.zeros((self.n_streams, n_frames, 2), dtype=torch.float64, device=dev)
        _abi.check(_abi.lib().icw_session_set_taps(self._h, bus.data_ptr(), lr.data_ptr()))
        self._taps = (bus, lr)
        return bus, lr

    def disable_taps(self) -> None:
        _abi.check(_abi.lib().icw_session_set_taps(self._h, None, None))
        self._taps = None
