"""ctypes mirror of include/icw_b200.h.  Loading fails loudly: there is no fallback path."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "libicw_b200.so"

N_PLUGS = 27
MAX_NODES = 32
MAX_ORD = 20
NS_MAX_TAPS = 20

OK, E_ARG, E_CUDA, E_UNSUPPORTED, E_NOMEM, E_MT_REDRAW = 0, -1, -2, -3, -4, -5

FMT = {
    "wav_u8": 0, "wav_i16": 1, "wav_i24": 2, "wav_i32": 3, "wav_f32": 4,
    "cw_f64": 16, "cw_i16": 17, "cw_i16f32": 18, "cw_f32": 19,
}
CHAN_BYTES = {0: 1, 1: 2, 2: 3, 3: 4, 4: 4, 16: 16, 17: 4, 18: 6, 19: 8}
MODE = {"master": 0, "shift": 1, "pm": 2, "mix": 3}
HILBERT = {"exact": 0, "scan": 1}
RESET_HILBERT, RESET_FRAMECNT, RESET_COUNTERS, RESET_FILEPOS, RESET_RENDER, RESET_RENDER_MEMORY, RESET_ALL = 1, 2, 4, 8, 16, 32, 255


class Node(C.Structure):
    _fields_ = [
        ("mode", C.c_int32), ("inputs_mask", C.c_uint32), ("xch_mode", C.c_int32),
        ("l_iq_invert", C.c_int32), ("r_iq_invert", C.c_int32),
        ("l_gain", C.c_double), ("r_gain", C.c_double),
        ("n_out", C.c_int32), ("l_tout", C.c_int32), ("r_tout", C.c_int32),
        ("l_on", C.c_int32), ("r_on", C.c_int32),
        ("l_p", C.c_double * 4), ("r_p", C.c_double * 4),
    ]


class ChainSpecC(C.Structure):
    _fields_ = [
        ("fmt", C.c_int32), ("n_channels", C.c_int32), ("sample_rate", C.c_uint32),
        ("n_samples", C.c_int64), ("n_fade_in", C.c_int64), ("n_fade_out", C.c_int64),
        ("filter_no", C.c_int32), ("is_kahan", C.c_int32), ("is_subnorm_reject", C.c_int32),
        ("hilbert_mode", C.c_int32), ("is_frmod_scaled", C.c_int32), ("need24bits", C.c_int32),
        ("dth_bits", C.c_double),
        ("quantz_type", C.c_uint32), ("render_type", C.c_uint32), ("nshape_type", C.c_uint32),
        ("sign_bits16", C.c_uint32), ("sign_bits24", C.c_uint32),
        ("bypass", C.c_int32), ("n_nodes", C.c_int32), ("nodes", Node * MAX_NODES),
        ("is_fp_check", C.c_int32), ("reserved_", C.c_int32),
    ]


class StreamState(C.Structure):
    _fields_ = [
        ("n_frame", C.c_uint64), ("pos", C.c_int64),
        ("hb", ((C.c_double * MAX_ORD) * 2) * 2),
        ("hb_rejects", (C.c_uint64 * 2) * 2),
        ("quad", C.c_uint32 * 2), ("mt_seed", C.c_uint32 * 2), ("mt_drawn", C.c_uint64 * 2),
        ("prev_rnd", C.c_double * 2), ("clips", C.c_uint32 * 2), ("peak", C.c_double * 2),
        ("bus", (C.c_double * 4) * N_PLUGS),
        ("hb_basis", C.c_uint32), ("reserved", C.c_uint32),
        ("ns_e", (C.c_double * NS_MAX_TAPS) * 2), ("ns_o", (C.c_double * NS_MAX_TAPS) * 2),
        ("ns_prev_err", C.c_double * 2),
    ]


class Stats(C.Structure):
    _fields_ = [
        ("clips", C.c_uint64 * 2), ("peak_db", C.c_double * 2), ("hb_rejects", C.c_uint64),
        ("mt_redraws", C.c_uint64), ("kernel_launches", C.c_uint64),
    ]


K_NAMES = ["hilbert", "chain", "mt", "misc", "scan_local", "scan_apply", "scan_fused"]


class Profile(C.Structure):
    _fields_ = [("ms", C.c_double * len(K_NAMES)), ("launches", C.c_uint64 * len(K_NAMES))]


# every symbol include/icw_b200.h declares; tests/test_abi.py checks the library exports them all
EXPORTS = [
    "icw_last_error", "icw_abi_version", "icw_default_spec", "icw_default_state", "icw_frame_bytes",
    "icw_out_frame_bytes", "icw_peak_db", "icw_engine_create", "icw_engine_destroy",
    "icw_session_create", "icw_session_destroy", "icw_session_set_spec", "icw_session_get_state",
    "icw_session_set_state", "icw_session_reset", "icw_session_process_host",
    "icw_session_process_device", "icw_session_sync", "icw_session_stats", "icw_session_fp_stats", "icw_hilbert_device",
    "icw_mt_words_device", "icw_session_set_taps", "icw_debug_phase_device", "icw_debug_sincos_device", "icw_debug_note_redraw", "icw_debug_patch_mt_word", "icw_mt_host_charpoly",
    "icw_mt_host_seq_state", "icw_mt_host_jump_state", "icw_mt_host_jump_state_family", "icw_mt_host_jump_state_product", "icw_mt_host_unit_blocks", "icw_host_scan_chunk_len",
    "icw_session_profile", "icw_session_profile_read", "icw_kernel_class_name", "icw_crc32_device", "icw_crc32_host", "icw_crc32_combine",
    "icw_pinned_alloc", "icw_pinned_free",
    "icw_session_boundary_export", "icw_session_boundary_import", "icw_session_seek_closed_form", "icw_comm_unique_id", "icw_comm_init",
    "icw_comm_destroy", "icw_nccl_version", "icw_session_handoff", "icw_session_reduce_counters", "icw_host_hb_convert",
]
BOUNDARY_DOUBLES = 2 * 2 * MAX_ORD
COMM_ID_BYTES = 128

_lib = None


def lib() -> C.CDLL:
    """The CUDA extension.  Raises if it has not been built: nothing else can do its work."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build the CUDA extension first "
            "(python -c 'import __graft_entry__ as g; g.build()' or python in_cwave_b200/build.py). "
            "in_cwave_b200 has no CPU fallback.")
    L = C.CDLL(str(LIB_PATH))
    vp, i64, u64, sz = C.c_void_p, C.c_int64, C.c_uint64, C.c_size_t
    P = C.POINTER
    L.icw_last_error.restype = C.c_char_p
    L.icw_abi_version.restype = C.c_int
    L.icw_default_spec.argtypes = [P(ChainSpecC)]
    L.icw_default_spec.restype = None
    L.icw_default_state.argtypes = [P(StreamState)]
    L.icw_default_state.restype = None
    L.icw_frame_bytes.argtypes = [P(ChainSpecC)]
    L.icw_out_frame_bytes.argtypes = [P(ChainSpecC)]
    L.icw_peak_db.argtypes = [C.c_double]
    L.icw_peak_db.restype = C.c_double
    L.icw_engine_create.argtypes = [C.c_int, P(vp)]
    L.icw_engine_destroy.argtypes = [vp]
    L.icw_engine_destroy.restype = None
    L.icw_session_create.argtypes = [vp, P(ChainSpecC), C.c_int, P(vp)]
    L.icw_session_destroy.argtypes = [vp]
    L.icw_session_destroy.restype = None
    L.icw_session_set_spec.argtypes = [vp, P(ChainSpecC)]
    L.icw_session_get_state.argtypes = [vp, C.c_int, P(StreamState)]
    L.icw_session_set_state.argtypes = [vp, C.c_int, P(StreamState)]
    L.icw_session_reset.argtypes = [vp, C.c_uint]
    L.icw_session_process_host.argtypes = [vp, i64, vp, sz, vp, sz]
    L.icw_session_process_device.argtypes = [vp, i64, vp, sz, vp, sz, vp]
    L.icw_session_sync.argtypes = [vp]
    L.icw_session_stats.argtypes = [vp, P(Stats)]
    L.icw_session_fp_stats.argtypes = [vp, C.c_int, P((C.c_uint32 * 7) * 4)]
    L.icw_session_profile.argtypes = [vp, C.c_int]
    L.icw_session_profile_read.argtypes = [vp, P(Profile), C.c_int]
    L.icw_crc32_device.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32)]
    L.icw_crc32_host.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t, C.POINTER(C.c_uint32)]
    L.icw_crc32_combine.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64]
    L.icw_crc32_combine.restype = C.c_uint32
    L.icw_kernel_class_name.argtypes = [C.c_int]
    L.icw_kernel_class_name.restype = C.c_char_p
    L.icw_hilbert_device.argtypes = [vp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, i64, vp, vp, P(StreamState)]
    L.icw_mt_words_device.argtypes = [vp, C.c_uint32, u64, i64, vp]
    L.icw_session_set_taps.argtypes = [vp, vp, vp]
    L.icw_debug_phase_device.argtypes = [vp, P(ChainSpecC), u64, i64, C.c_double, vp]
    L.icw_debug_sincos_device.argtypes = [vp, i64, vp, vp]
    L.icw_debug_note_redraw.argtypes = [vp, C.c_int, C.c_uint64]
    L.icw_debug_patch_mt_word.argtypes = [vp, C.c_int, C.c_uint64, C.c_uint32]
    L.icw_mt_host_charpoly.argtypes = [P(C.c_int), P(C.c_int)]
    L.icw_mt_host_seq_state.argtypes = [C.c_uint32, u64, P(C.c_uint32)]
    L.icw_mt_host_seq_state.restype = None
    L.icw_mt_host_jump_state.argtypes = [C.c_uint32, u64, P(C.c_uint32)]
    L.icw_mt_host_jump_state_family.argtypes = [C.c_uint32, u64, P(C.c_uint32)]
    L.icw_mt_host_jump_state_product.argtypes = [C.c_uint32, u64, P(C.c_uint32)]
    L.icw_mt_host_unit_blocks.argtypes = [u64, C.c_int]
    L.icw_mt_host_unit_blocks.restype = u64
    L.icw_host_scan_chunk_len.argtypes = [C.c_int, i64, C.c_int]
    L.icw_pinned_alloc.argtypes = [sz, P(vp)]
    L.icw_pinned_free.argtypes = [vp]
    L.icw_pinned_free.restype = None
    L.icw_session_boundary_export.argtypes = [vp, C.c_int, vp, P(C.c_int), vp]
    L.icw_session_boundary_import.argtypes = [vp, C.c_int, vp, C.c_int, vp]
    L.icw_session_seek_closed_form.argtypes = [vp, C.c_int, i64, P(StreamState)]
    L.icw_comm_unique_id.argtypes = [C.c_char_p]
    L.icw_comm_init.argtypes = [C.c_char_p, C.c_int, C.c_int, P(vp)]
    L.icw_comm_destroy.argtypes = [vp]
    L.icw_session_handoff.argtypes = [vp, vp, vp, C.c_int, C.c_int, vp]
    L.icw_session_reduce_counters.argtypes = [vp, vp, vp, P(C.c_uint64 * 2), P(C.c_double * 2)]
    L.icw_host_hb_convert.argtypes = [C.c_int, C.c_int, P(C.c_double), P(C.c_double)]
    _lib = L
    return L


class IcwError(RuntimeError):
    def __init__(self, code: int, msg: str):
        super().__init__(f"icw error {code}: {msg}")
        self.code = code


def check(rc: int) -> None:
    if rc != OK:
        raise IcwError(rc, lib().icw_last_error().decode(errors="replace"))
