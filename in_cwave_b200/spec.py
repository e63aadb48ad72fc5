"""Chain description on the Python side: the reference's config + reader facts + DSP list.

A spec is a plain dict (JSON-able) so that tests can hand the same description to the product
(``to_c``) and to the oracles (oracle/pyoracle.make_spec) without either depending on the other.
Keys mirror icw_chain_spec in include/icw_b200.h; nodes are listed in EXECUTION order, master last.
"""
from __future__ import annotations

import copy

from . import _abi

DEFAULT_MASTER = dict(mode="master", inputs=[0], l_gain=0.8, r_gain=0.8, l_tout=0, r_tout=0)


def default_spec(**over) -> dict:
    """Reference defaults (src/config.c:118-207; lone master src/adv_modulator.c:112-118)."""
    d = dict(
        fmt="wav_f32", n_channels=2, sample_rate=48000, n_samples=0, n_fade_in=0, n_fade_out=0,
        filter_no=1, is_kahan=1, is_subnorm_reject=1, hilbert_mode="exact", is_frmod_scaled=1,
        need24bits=1, dth_bits=1.0, quantz_type=1, render_type=0, nshape_type=0,
        sign_bits16=16, sign_bits24=24, bypass=0, is_fp_check=0, nodes=[copy.deepcopy(DEFAULT_MASTER)],
    )
    d.update(over)
    return d


def _fill_node(dst: _abi.Node, nd: dict) -> None:
    m = nd["mode"]
    dst.mode = _abi.MODE[m] if isinstance(m, str) else int(m)
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


def to_c(d: dict) -> _abi.ChainSpecC:
    full = default_spec()
    full.update(d)
    sp = _abi.ChainSpecC()
    _abi.lib().icw_default_spec(sp)
    f = full["fmt"]
    sp.fmt = _abi.FMT[f] if isinstance(f, str) else int(f)
    sp.n_channels = int(full["n_channels"])
    sp.sample_rate = int(full["sample_rate"])
    sp.n_samples = int(full["n_samples"])
    sp.n_fade_in = int(full["n_fade_in"])
    sp.n_fade_out = int(full["n_fade_out"])
    sp.filter_no = int(full["filter_no"])
    sp.is_kahan = int(full["is_kahan"])
    sp.is_subnorm_reject = int(full["is_subnorm_reject"])
    hm = full["hilbert_mode"]
    sp.hilbert_mode = _abi.HILBERT[hm] if isinstance(hm, str) else int(hm)
    sp.is_frmod_scaled = int(full["is_frmod_scaled"])
    sp.need24bits = int(full["need24bits"])
    sp.dth_bits = float(full["dth_bits"])
    sp.quantz_type = int(full["quantz_type"])
    sp.render_type = int(full["render_type"])
    sp.nshape_type = int(full["nshape_type"])
    sp.sign_bits16 = int(full["sign_bits16"])
    sp.sign_bits24 = int(full["sign_bits24"])
    sp.bypass = int(full["bypass"])
    sp.is_fp_check = int(full.get("is_fp_check", 0))
    nodes = full["nodes"]
    if len(nodes) > _abi.MAX_NODES:
        raise ValueError("too many nodes")
    sp.n_nodes = len(nodes)
    for i, nd in enumerate(nodes):
        _fill_node(sp.nodes[i], nd)
    return sp


def frame_bytes(d: dict) -> int:
    f = d.get("fmt", "wav_f32")
    f = _abi.FMT[f] if isinstance(f, str) else int(f)
    return _abi.CHAN_BYTES[f] * int(d.get("n_channels", 2))


def out_frame_bytes(d: dict) -> int:
    return 6 if int(d.get("need24bits", 1)) else 4


# ---- the BASELINE.json configurations ---------------------------------
def _shift100():
    return [dict(mode="shift", inputs=[0], out=1, l_p=[100.0], r_p=[100.0]),
            dict(mode="master", inputs=[1], l_gain=0.8, r_gain=0.8)]


def config_c1(**over) -> dict:
    """48 kHz stereo float WAV: Hilbert + one +100 Hz spectrum shift + 24-bit render, no dither."""
    return default_spec(**{**dict(fmt="wav_f32", sample_rate=48000, nodes=_shift100()), **over})


def config_c2(**over) -> dict:
    """192 kHz 24-bit PCM stereo: Hilbert + shift, TPDF dither on (MT seeded like the plugin)."""
    return default_spec(**{**dict(fmt="wav_i24", sample_rate=192000, render_type=2, nodes=_shift100()), **over})


def config_c3(**over) -> dict:
    """CWAVE f32 I/Q 96 kHz stereo: multi-shift + harmonic phase modulation + channel mixing, 16-bit."""
    nodes = [
        dict(mode="shift", inputs=[0], out=26, l_p=[7.5], r_p=[-7.5]),
        dict(mode="shift", inputs=[26], out=1, l_p=[-3.25], r_p=[3.25], l_gain=0.7, r_gain=0.7),
        dict(mode="pm", inputs=[0], out=2, l_p=[4.0, 0.0, 0.5, 0.0], r_p=[4.0, 0.5, 0.5, 0.0],
             l_gain=0.5, r_gain=0.5),
        dict(mode="mix", inputs=[1, 2], out=3, xch=4, r_iq_invert=1),
        dict(mode="master", inputs=[1, 3], l_gain=0.6, r_gain=0.6),
    ]
    return default_spec(**{**dict(fmt="cw_f32", sample_rate=96000, need24bits=0, nodes=nodes), **over})
