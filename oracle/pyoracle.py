"""ctypes bindings for the two CPU oracles.  TEST INFRASTRUCTURE ONLY.

* ``port()``  -> oracle/libicw_oracle.so, our plain-C restatement (icw_oracle.c); always buildable.
* ``ref()``   -> oracle/_ref/libicw_ref.so, the reference's own C sources compiled in place from
                 /root/reference by oracle/Makefile; present when it was built in the build
                 container (the file travels to the GPU box, the reference tree does not).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs import
this module.  Nothing under in_cwave_b200/ does.
"""
from __future__ import annotations

import ctypes as C
import os
import struct
import subprocess
import tempfile
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
N_PLUGS = 27
MAX_NODES = 32
MAX_ORD = 20
MT_N = 624

FMT = {
    "wav_u8": 0, "wav_i16": 1, "wav_i24": 2, "wav_i32": 3, "wav_f32": 4,
    "cw_f64": 16, "cw_i16": 17, "cw_i16f32": 18, "cw_f32": 19,
}
CHAN_BYTES = {0: 1, 1: 2, 2: 3, 3: 4, 4: 4, 16: 16, 17: 4, 18: 6, 19: 8}
MODE = {"master": 0, "shift": 1, "pm": 2, "mix": 3}


class Node(C.Structure):
    """icwo_node == icwref_node (same layout on purpose): one DSP-list node, execution order."""
    _fields_ = [
        ("mode", C.c_int), ("inputs_mask", C.c_uint), ("xch_mode", C.c_int),
        ("l_iq_invert", C.c_int), ("r_iq_invert", C.c_int),
        ("l_gain", C.c_double), ("r_gain", C.c_double),
        ("n_out", C.c_int), ("l_tout", C.c_int), ("r_tout", C.c_int),
        ("l_on", C.c_int), ("r_on", C.c_int),
        ("l_p", C.c_double * 4), ("r_p", C.c_double * 4),
    ]


class Spec(C.Structure):
    _fields_ = [
        ("fmt", C.c_int), ("n_channels", C.c_int), ("sample_rate", C.c_uint),
        ("n_samples", C.c_int64), ("n_fade_in", C.c_int64), ("n_fade_out", C.c_int64),
        ("filter_no", C.c_int), ("is_kahan", C.c_int), ("is_subnorm_reject", C.c_int),
        ("is_frmod_scaled", C.c_int), ("need24bits", C.c_int), ("dth_bits", C.c_double),
        ("quantz_type", C.c_uint), ("render_type", C.c_uint), ("nshape_type", C.c_uint),
        ("sign_bits16", C.c_uint), ("sign_bits24", C.c_uint),
        ("bypass", C.c_int), ("n_nodes", C.c_int), ("nodes", Node * MAX_NODES),
        ("is_fp_check", C.c_int),
    ]


class Iir(C.Structure):
    _fields_ = [("z", C.c_double * MAX_ORD), ("ix", C.c_int), ("rejects", C.c_uint64)]


class Mt(C.Structure):
    _fields_ = [("w", C.c_uint32 * MT_N), ("pos", C.c_int), ("drawn", C.c_uint64)]


NS_MAX_TAPS = 20


class Ns(C.Structure):
    _fields_ = [("e", C.c_double * NS_MAX_TAPS), ("o", C.c_double * NS_MAX_TAPS), ("prev_err", C.c_double)]


class State(C.Structure):
    _fields_ = [
        ("n_frame", C.c_uint64), ("pos", C.c_int64),
        ("bus", (C.c_double * 4) * N_PLUGS),
        ("lpf", (Iir * 2) * 2), ("quad", C.c_uint * 2),
        ("mt", Mt * 2), ("prev_rnd", C.c_double * 2),
        ("clips", C.c_uint * 2), ("peak_db", C.c_double * 2),
        ("ns", Ns * 2),
        ("fp_cnt", (C.c_uint * 7) * 4),
    ]


class RefCfg(C.Structure):
    _fields_ = [
        ("filter_no", C.c_int), ("is_kahan", C.c_int), ("is_subnorm_reject", C.c_int),
        ("subnorm_thr", C.c_double), ("is_frmod_scaled", C.c_int), ("need24bits", C.c_int),
        ("is_fp_check", C.c_int), ("dth_bits", C.c_double),
        ("quantz_type", C.c_uint), ("render_type", C.c_uint), ("nshape_type", C.c_uint),
        ("sign_bits16", C.c_uint), ("sign_bits24", C.c_uint),
        ("sec_align", C.c_uint), ("fade_in", C.c_uint), ("fade_out", C.c_uint),
        ("clr_nframe_trk", C.c_int), ("clr_hilb_trk", C.c_int),
    ]


class RefStats(C.Structure):
    _fields_ = [
        ("l_clips", C.c_uint), ("r_clips", C.c_uint), ("l_peak", C.c_double), ("r_peak", C.c_double),
        ("subnorm_cnt", C.c_uint64), ("n_frame", C.c_uint64),
    ]


_p = C.POINTER
_dbl_p = _p(C.c_double)
_u8_p = _p(C.c_uint8)


def _arr(a, ctype):
    return a.ctypes.data_as(_p(ctype)) if a is not None else None


def build(ref: bool = True) -> None:
    """(Re)build the oracle libraries with oracle/Makefile.  Building is not using."""
    subprocess.run(["make", "-s", "-C", str(HERE), "port"], check=True)
    if ref:
        subprocess.run(["make", "-s", "-C", str(HERE), "ref"], check=True)
        subprocess.run(["make", "-s", "-C", str(HERE), "ref_gpu"], check=True)     # needs libicw_b200.so: a no-op without it


_port = None
_ref = None


def port():
    global _port
    if _port is None:
        so = HERE / "libicw_oracle.so"
        if not so.exists():
            build(ref=False)
        L = C.CDLL(str(so))
        L.icwo_default_spec.argtypes = [_p(Spec)]
        L.icwo_state_init.argtypes = [_p(State)]
        L.icwo_frame_bytes.argtypes = [_p(Spec)]
        L.icwo_out_frame_bytes.argtypes = [_p(Spec)]
        L.icwo_process.argtypes = [_p(Spec), _p(State), _u8_p, C.c_int64, _u8_p, _dbl_p, _dbl_p,
                                   _p(C.c_int), C.c_int, _dbl_p]
        L.icwo_process.restype = C.c_int
        L.icwo_mt_seed.argtypes = [_p(Mt), C.c_uint32]
        L.icwo_mt_seed_key.argtypes = [_p(Mt), _p(C.c_uint32), C.c_uint32]
        L.icwo_mt_u32.argtypes = [_p(Mt)]
        L.icwo_mt_u32.restype = C.c_uint32
        L.icwo_mt_dsopen.argtypes = [_p(Mt)]
        L.icwo_mt_dsopen.restype = C.c_double
        L.icwo_unpack.argtypes = [C.c_int, C.c_int, _u8_p, C.c_int64, _dbl_p]
        L.icwo_hilbert.argtypes = [C.c_int, C.c_int, C.c_int, _p(Iir), _p(C.c_uint), _dbl_p,
                                   C.c_int64, _dbl_p, _dbl_p]
        L.icwo_iir_run.argtypes = [C.c_int, C.c_int, C.c_int, _p(Iir), _dbl_p, C.c_int64, _dbl_p]
        L.icwo_hilbert_truth.argtypes = [C.c_int, C.c_int, C.c_uint, _dbl_p, C.c_int64, _dbl_p, _dbl_p]
        L.icwo_render.argtypes = [_p(Spec), _p(Mt), _dbl_p, _p(Ns), _dbl_p, C.c_int64, _u8_p,
                                  _p(C.c_uint), _dbl_p]
        L.icwo_render.restype = C.c_int64
        _port = L
    return _port


def hilbert_truth(x: np.ndarray, filter_no: int = 1, drop_direct: int = 1, quad0: int = 0):
    """binary128 evaluation of the converter from zero state -> (I, Q) as float64 arrays."""
    x = np.ascontiguousarray(x, dtype=np.float64)
    oi, oq = np.zeros(x.size), np.zeros(x.size)
    port().icwo_hilbert_truth(filter_no, drop_direct, quad0, _arr(x, C.c_double), x.size,
                              _arr(oi, C.c_double), _arr(oq, C.c_double))
    return oi, oq


def have_ref() -> bool:
    return (HERE / "_ref" / "libicw_ref.so").exists()


def _bind_ref(L):
    L.icwref_default_cfg.argtypes = [_p(RefCfg)]
    L.icwref_reset.argtypes = [_p(RefCfg)]
    L.icwref_set_graph.argtypes = [_p(Node), C.c_int, C.c_int]
    L.icwref_set_graph.restype = C.c_int
    L.icwref_process_file.argtypes = [C.c_char_p, C.c_uint, C.c_char_p, C.c_int64, _dbl_p,
                                      _p(C.c_int), C.c_int]
    L.icwref_process_file.restype = C.c_int64
    L.icwref_transcode_file.argtypes = [C.c_char_p, C.c_int, C.c_char_p, C.c_int64, _p(C.c_int)]
    L.icwref_transcode_file.restype = C.c_int64
    L.icwref_get_stats.argtypes = [_p(RefStats), C.c_int]
    L.icwref_hilbert.argtypes = [C.c_uint, C.c_int, C.c_int, _dbl_p, C.c_int64, _dbl_p, _dbl_p]
    L.icwref_hilbert.restype = C.c_uint64
    L.icwref_iir.argtypes = [C.c_uint, C.c_int, C.c_int, _dbl_p, C.c_int64, _dbl_p]
    L.icwref_render.argtypes = [_p(RefCfg), C.c_uint32, _dbl_p, C.c_int64, C.c_char_p,
                                _p(C.c_uint), _dbl_p]
    L.icwref_render.restype = C.c_int64
    L.icwref_mt_words.argtypes = [C.c_uint32, C.c_int64, C.c_int64, _p(C.c_uint32)]
    L.icwref_mt_words_key.argtypes = [_p(C.c_uint32), C.c_uint32, C.c_int64, _p(C.c_uint32)]
    L.icwref_mt_dsopen.argtypes = [C.c_uint32, C.c_int64, _dbl_p]
    L.icwref_reset_from_file.argtypes = [C.c_char_p]
    L.icwref_reset_from_file.restype = C.c_int
    L.icwref_save_config.argtypes = [C.c_char_p]
    L.icwref_save_config.restype = C.c_int
    L.icwref_get_cfg.argtypes = [_p(RefCfg)]
    L.icwref_get_graph.argtypes = [_p(Node), C.c_int]
    L.icwref_get_graph.restype = C.c_int
    L.icwref_set_render_live.argtypes = [_p(RefCfg)]
    L.icwref_reset_live.argtypes = [C.c_int, C.c_int]
    return L


def ref():
    global _ref
    if _ref is None:
        so = HERE / "_ref" / "libicw_ref.so"
        if not so.exists():
            raise FileNotFoundError(f"{so} not built (needs /root/reference; run make -C oracle ref)")
        _ref = _bind_ref(C.CDLL(str(so)))
    return _ref


_ref_gpu = None


def have_ref_gpu() -> bool:
    return (HERE / "_ref_gpu" / "libicw_ref_gpu.so").exists()


def ref_gpu():
    """The SAME reference sources with amod_process_samples' frame loop running on the GPU
    (in_cwave_b200/host/adv_modulator_gpu.c over libicw_b200.so; `make -C oracle ref_gpu`)."""
    global _ref_gpu
    if _ref_gpu is None:
        so = HERE / "_ref_gpu" / "libicw_ref_gpu.so"
        if not so.exists():
            raise FileNotFoundError(f"{so} not built (needs /root/reference and libicw_b200.so; run make -C oracle ref_gpu)")
        L = _bind_ref(C.CDLL(str(so)))
        L.amod_gpu_last_error.restype = C.c_char_p
        _ref_gpu = L
    return _ref_gpu


# ---------------------------------------------------------------------------------------------
# spec helpers: a chain description as a plain dict (see in_cwave_b200.spec.ChainSpec.as_dict)
# ---------------------------------------------------------------------------------------------
def fill_node(dst: Node, nd: dict) -> None:
    dst.mode = MODE[nd["mode"]] if isinstance(nd["mode"], str) else int(nd["mode"])
    mask = 0
    for k in nd.get("inputs", [0]):
        mask |= 1 << int(k)
    dst.inputs_mask = mask
    dst.xch_mode = int(nd.get("xch", 0))
    dst.l_iq_invert = int(nd.get("l_iq_invert", 0))
    dst.r_iq_invert = int(nd.get("r_iq_invert", 0))
    dst.l_gain = float(nd.get("l_gain", 1.0))
    dst.r_gain = float(nd.get("r_gain", 1.0))
    dst.n_out = int(nd.get("out", 26))
    dst.l_tout = int(nd.get("l_tout", 0))
    dst.r_tout = int(nd.get("r_tout", 0))
    dst.l_on = int(nd.get("l_on", 1))
    dst.r_on = int(nd.get("r_on", 1))
    lp = list(nd.get("l_p", [])) + [0.0] * 4
    rp = list(nd.get("r_p", [])) + [0.0] * 4
    for i in range(4):
        dst.l_p[i] = float(lp[i])
        dst.r_p[i] = float(rp[i])


def make_spec(d: dict) -> Spec:
    sp = Spec()
    port().icwo_default_spec(C.byref(sp))
    sp.fmt = FMT[d["fmt"]] if isinstance(d.get("fmt", "wav_f32"), str) else int(d["fmt"])
    sp.n_channels = int(d.get("n_channels", 2))
    sp.sample_rate = int(d.get("sample_rate", 48000))
    sp.n_samples = int(d.get("n_samples", 0))
    sp.n_fade_in = int(d.get("n_fade_in", 0))
    sp.n_fade_out = int(d.get("n_fade_out", 0))
    sp.filter_no = int(d.get("filter_no", 1))
    sp.is_kahan = int(d.get("is_kahan", 1))
    sp.is_subnorm_reject = int(d.get("is_subnorm_reject", 1))
    sp.is_frmod_scaled = int(d.get("is_frmod_scaled", 1))
    sp.need24bits = int(d.get("need24bits", 1))
    sp.dth_bits = float(d.get("dth_bits", 1.0))
    sp.quantz_type = int(d.get("quantz_type", 1))
    sp.render_type = int(d.get("render_type", 0))
    sp.nshape_type = int(d.get("nshape_type", 0))
    sp.sign_bits16 = int(d.get("sign_bits16", 16))
    sp.sign_bits24 = int(d.get("sign_bits24", 24))
    sp.bypass = int(d.get("bypass", 0))
    sp.is_fp_check = int(d.get("is_fp_check", 0))
    nodes = d.get("nodes")
    if nodes is not None:
        sp.n_nodes = len(nodes)
        for i, nd in enumerate(nodes):
            fill_node(sp.nodes[i], nd)
    return sp


def make_refcfg(d: dict) -> RefCfg:
    c = RefCfg()
    ref().icwref_default_cfg(C.byref(c))
    for k in ("filter_no", "is_kahan", "is_subnorm_reject", "is_frmod_scaled", "need24bits",
              "quantz_type", "render_type", "nshape_type", "sign_bits16", "sign_bits24",
              "sec_align", "fade_in", "fade_out", "clr_nframe_trk", "clr_hilb_trk", "is_fp_check"):
        if k in d:
            setattr(c, k, int(d[k]))
    if "dth_bits" in d:
        c.dth_bits = float(d["dth_bits"])
    return c


def new_state() -> State:
    st = State()
    port().icwo_state_init(C.byref(st))
    return st


def frame_bytes(d: dict) -> int:
    f = FMT[d["fmt"]] if isinstance(d.get("fmt", "wav_f32"), str) else int(d["fmt"])
    return CHAN_BYTES[f] * int(d.get("n_channels", 2))


def port_process(d: dict, raw: np.ndarray, state: State | None = None, taps: list[int] | None = None,
                 want_analytic: bool = False, want_lr: bool = False):
    """Run the port oracle over raw input bytes.  Returns dict(pcm, analytic, bus, lr, state)."""
    L = port()
    sp = make_spec(d)
    st = state if state is not None else new_state()
    raw = np.ascontiguousarray(raw, dtype=np.uint8)
    fb = L.icwo_frame_bytes(C.byref(sp))
    n = raw.size // fb
    ob = L.icwo_out_frame_bytes(C.byref(sp))
    pcm = np.zeros(n * ob, dtype=np.uint8)
    analytic = np.zeros((n, 4)) if want_analytic else None
    lr = np.zeros((n, 2)) if want_lr else None
    taps = list(taps or [])
    bus = np.zeros((n, len(taps), 4)) if taps else None
    tp = (C.c_int * max(1, len(taps)))(*taps)
    rc = L.icwo_process(C.byref(sp), C.byref(st), _arr(raw, C.c_uint8), n, _arr(pcm, C.c_uint8),
                        _arr(analytic, C.c_double), _arr(bus, C.c_double), tp, len(taps),
                        _arr(lr, C.c_double))
    if rc != 0:
        raise ValueError(f"icwo_process rejected the spec (rc={rc})")
    return dict(pcm=pcm, analytic=analytic, bus=bus, lr=lr, state=st, frames=n)


# ---------------------------------------------------------------------------------------------
# files for the compiled reference (it only reads files: src/xwave_reader.c:593)
# ---------------------------------------------------------------------------------------------
def wav_bytes(d: dict, raw: np.ndarray, extensible: bool = False) -> bytes:
    """Canonical RIFF/WAVE around raw little-endian sample bytes (real formats only)."""
    fmt = d["fmt"]
    nch = int(d.get("n_channels", 2))
    bits = {"wav_u8": 8, "wav_i16": 16, "wav_i24": 24, "wav_i32": 32, "wav_f32": 32}[fmt]
    tag = 3 if fmt == "wav_f32" else 1
    sr = int(d.get("sample_rate", 48000))
    ba = nch * bits // 8
    data = np.ascontiguousarray(raw, dtype=np.uint8).tobytes()
    if extensible:
        guid = struct.pack("<IHH8B", tag, 0x0000, 0x0010, 0x80, 0x00, 0x00, 0xAA, 0x00, 0x38, 0x9B, 0x71)
        fmtc = struct.pack("<HHIIHHHHI", 0xFFFE, nch, sr, sr * ba, ba, bits, 22, bits, (1 << nch) - 1) + guid
    else:
        fmtc = struct.pack("<HHIIHH", tag, nch, sr, sr * ba, ba, bits)
    body = b"WAVE" + b"fmt " + struct.pack("<I", len(fmtc)) + fmtc + b"data" + struct.pack("<I", len(data)) + data
    return b"RIFF" + struct.pack("<I", len(body)) + body


def cwave_bytes(d: dict, raw: np.ndarray, version: int = 2, crc: int = 0) -> bytes:
    """CWAVE V2 header (reference src/cwave.h:47-59, 48 bytes packed) + raw I/Q sample bytes."""
    fmt = {"cw_f64": 0, "cw_i16": 1, "cw_i16f32": 2, "cw_f32": 3}[d["fmt"]]
    nch = int(d.get("n_channels", 2))
    data = np.ascontiguousarray(raw, dtype=np.uint8).tobytes()
    n = len(data) // frame_bytes(d)
    hdr = b"cPLXwAVE" + struct.pack("<IIIIIIiId", 48, version, fmt, nch, n,
                                     int(d.get("sample_rate", 96000)), -1, crc & 0xFFFFFFFF, 0.0)
    assert len(hdr) == 48
    return hdr + data


def ref_process(d: dict, raw: np.ndarray, taps: list[int] | None = None, read_quant: int = 4096,
                reset: bool = True, tmpdir: str | None = None, lib=None):
    """Run the compiled reference over the same bytes (written to a temp file).  lib: ref() (default) or ref_gpu()."""
    L = lib if lib is not None else ref()
    if reset:
        cfg = make_refcfg(d)
        # ms-based fades in the reference config; specs carry frames -> convert when exact
        sr = int(d.get("sample_rate", 48000))
        for key, ms_key in (("n_fade_in", "fade_in"), ("n_fade_out", "fade_out")):
            if d.get(key):
                ms = d[key] * 1000 // sr
                assert ms * sr // 1000 == d[key], "fade length must be a whole number of ms"
                setattr(cfg, ms_key, int(ms))
        L.icwref_reset(C.byref(cfg))
        nodes = d.get("nodes") or [dict(mode="master", inputs=[0], l_gain=0.8, r_gain=0.8)]
        arr = (Node * len(nodes))()
        for i, nd in enumerate(nodes):
            fill_node(arr[i], nd)
        rc = L.icwref_set_graph(arr, len(nodes), int(d.get("bypass", 0)))
        if rc:
            raise ValueError(f"icwref_set_graph rc={rc}")
    is_cw = d["fmt"].startswith("cw_")
    blob = cwave_bytes(d, raw) if is_cw else wav_bytes(d, raw, extensible=bool(d.get("wav_extensible", 0)))
    fd, path = tempfile.mkstemp(suffix=".cwave" if is_cw else ".wav", dir=tmpdir)
    try:
        with os.fdopen(fd, "wb") as f:
            f.write(blob)
        n = np.asarray(raw).size // frame_bytes(d)
        ob = 6 if int(d.get("need24bits", 1)) else 4
        pcm = np.zeros(n * ob + 64, dtype=np.uint8)
        taps = list(taps or [])
        bus = np.zeros((n, len(taps), 4)) if taps else None
        tp = (C.c_int * max(1, len(taps)))(*taps)
        got = L.icwref_process_file(path.encode(), read_quant, pcm.ctypes.data_as(C.c_char_p),
                                    n * ob, _arr(bus, C.c_double), tp, len(taps))
    finally:
        os.unlink(path)
    if got < 0:
        raise RuntimeError(f"reference refused the file (rc={got})")
    st = RefStats()
    L.icwref_get_stats(C.byref(st), 0)
    fp = ((C.c_uint * 7) * 4)()
    L.icwref_fp_stats(fp)
    return dict(pcm=pcm[: got * ob], bus=bus, frames=int(got), stats=st, fp_cnt=[list(r) for r in fp])


# ---------------------------------------------------------------------------------------------
# the reference's config file, through the reference's own load_config / save_config
# ---------------------------------------------------------------------------------------------
def _node_tuple(n: Node):
    return (n.mode, n.inputs_mask, n.xch_mode, n.l_iq_invert, n.r_iq_invert, n.l_gain, n.r_gain,
            n.n_out if n.mode != 0 else 0, n.l_tout, n.r_tout, n.l_on, n.r_on, tuple(n.l_p), tuple(n.r_p))


def ref_load_config(path: str):
    """Fresh reference plugin configured from `path`.  Returns (accepted, RefCfg, nodes in execution order)."""
    L = ref()
    ok = L.icwref_reset_from_file(str(path).encode())
    cfg = RefCfg()
    L.icwref_get_cfg(C.byref(cfg))
    arr = (Node * 64)()
    n = L.icwref_get_graph(arr, 64)
    return bool(ok), cfg, [_node_tuple(arr[i]) for i in range(n)]


def ref_save_config(d: dict, path: str) -> bool:
    """Configure the reference from a spec dict and let ITS save_config() write the file."""
    L = ref()
    cfg = make_refcfg(d)
    L.icwref_reset(C.byref(cfg))
    nodes = d.get("nodes") or [dict(mode="master", inputs=[0])]
    arr = (Node * len(nodes))()
    for i, nd in enumerate(nodes):
        fill_node(arr[i], nd)
    rc = L.icwref_set_graph(arr, len(nodes), 0)
    if rc:
        raise ValueError(f"icwref_set_graph rc={rc}")
    ok = L.icwref_save_config(str(path).encode())
    L.icwref_reset(C.byref(cfg))            # the list was consumed by the writer: re-initialise
    return bool(ok)
