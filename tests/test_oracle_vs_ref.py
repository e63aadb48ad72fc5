"""The port oracle against the compiled reference itself, on fresh random inputs and at the leaf
level.  Skipped where oracle/_ref was not built (it needs /root/reference, i.e. the build
container); the committed golden vectors in test_oracle_golden.py cover the GPU box."""
import ctypes as C

import numpy as np
import pytest

from in_cwave_b200 import spec as S
from in_cwave_b200 import synth
from oracle import pyoracle as po

pytestmark = pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref/libicw_ref.so not built")


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


@pytest.mark.parametrize("ft", range(6))
@pytest.mark.parametrize("kahan", [0, 1])
@pytest.mark.parametrize("reject", [0, 1])
def test_hilbert_leaf(ft, kahan, reject):
    rng = np.random.default_rng(100 + ft)
    n = 6000
    x = (rng.random(n) - 0.5) * 30000.0
    x[100:140] = 0.0
    ri, rq = np.zeros(n), np.zeros(n)
    cnt = po.ref().icwref_hilbert(ft, kahan, reject, _dp(x), n, _dp(ri), _dp(rq))
    pi, pq = np.zeros(n), np.zeros(n)
    lpf = (po.Iir * 2)()
    quad = C.c_uint(0)
    po.port().icwo_hilbert(ft, kahan, reject, lpf, C.byref(quad), _dp(x), n, _dp(pi), _dp(pq))
    assert np.array_equal(ri, pi) and np.array_equal(rq, pq)
    assert cnt == lpf[0].rejects + lpf[1].rejects
    assert quad.value == n % 4


def test_hilbert_silence_rejects():
    """Digital silence / decaying tail: the |w| < 1 zeroing makes the filter non-linear."""
    n = 30000
    x = np.zeros(n)
    x[:50] = 1000.0
    for kahan in (0, 1):
        ri, rq = np.zeros(n), np.zeros(n)
        cnt = po.ref().icwref_hilbert(1, kahan, 1, _dp(x), n, _dp(ri), _dp(rq))
        pi, pq = np.zeros(n), np.zeros(n)
        lpf = (po.Iir * 2)()
        quad = C.c_uint(0)
        po.port().icwo_hilbert(1, kahan, 1, lpf, C.byref(quad), _dp(x), n, _dp(pi), _dp(pq))
        assert np.array_equal(ri, pi) and np.array_equal(rq, pq)
        assert cnt == lpf[0].rejects + lpf[1].rejects


@pytest.mark.parametrize("rt", range(5))
@pytest.mark.parametrize("need24,bits", [(1, 24), (1, 17), (0, 16), (0, 9)])
@pytest.mark.parametrize("qt", [0, 1])
def test_render_leaf(rt, need24, bits, qt):
    rng = np.random.default_rng(rt * 10 + bits)
    n = 5000
    x = (rng.random(n) - 0.5) * 70000.0          # beyond full scale: clips on both sides
    d = S.default_spec(render_type=rt, need24bits=need24, quantz_type=qt, dth_bits=2.5,
                       sign_bits24=bits if need24 else 24, sign_bits16=bits if not need24 else 16)
    cfg = po.make_refcfg(d)
    buf = C.create_string_buffer(n * 3)
    clips, peak = C.c_uint(0), C.c_double(0)
    nb = po.ref().icwref_render(C.byref(cfg), 0x13579BDF, _dp(x), n, buf, C.byref(clips), C.byref(peak))
    sp = po.make_spec(d)
    mt = po.Mt()
    po.port().icwo_mt_seed(C.byref(mt), 0x13579BDF)
    out = np.zeros(n * 3, dtype=np.uint8)
    pclips, ppeak, prev = C.c_uint(0), C.c_double(-555.0), C.c_double(0)
    pb = po.port().icwo_render(C.byref(sp), C.byref(mt), C.byref(prev), None, _dp(x), n,
                               out.ctypes.data_as(C.POINTER(C.c_uint8)), C.byref(pclips), C.byref(ppeak))
    assert nb == pb
    assert bytes(out[:pb]) == buf.raw[:nb]
    assert clips.value == pclips.value and clips.value > 0
    assert peak.value == ppeak.value


@pytest.mark.parametrize("ns", range(1, 18))
@pytest.mark.parametrize("rt,need24,bits,qt", [(0, 1, 24, 1), (2, 0, 16, 1), (4, 1, 20, 0), (3, 0, 12, 0)])
def test_render_leaf_with_noise_shaping(ns, rt, need24, bits, qt):
    """All 15 FIR and 2 IIR shapers (src/sound_render.c:75-235, 403-489): the error feedback makes every
    byte depend on every earlier one, so byte equality over 20 000 samples pins coefficients and order."""
    rng = np.random.default_rng(ns * 7 + rt)
    n = 20000
    x = (rng.random(n) - 0.5) * 50000.0
    x[n // 2:] *= 1.5                               # second half clips now and then
    d = S.default_spec(render_type=rt, need24bits=need24, quantz_type=qt, dth_bits=1.0, nshape_type=ns,
                       sign_bits24=bits if need24 else 24, sign_bits16=bits if not need24 else 16)
    cfg = po.make_refcfg(d)
    buf = C.create_string_buffer(n * 3)
    clips, peak = C.c_uint(0), C.c_double(0)
    nb = po.ref().icwref_render(C.byref(cfg), 0x479B22AB, _dp(x), n, buf, C.byref(clips), C.byref(peak))
    sp = po.make_spec(d)
    mt = po.Mt()
    po.port().icwo_mt_seed(C.byref(mt), 0x479B22AB)
    out = np.zeros(n * 3, dtype=np.uint8)
    pclips, ppeak, prev, nss = C.c_uint(0), C.c_double(-555.0), C.c_double(0), po.Ns()
    # two calls: the shaper memory must carry over
    h = n // 3
    u8 = C.POINTER(C.c_uint8)
    pb = po.port().icwo_render(C.byref(sp), C.byref(mt), C.byref(prev), C.byref(nss), _dp(x), h,
                               out.ctypes.data_as(u8), C.byref(pclips), C.byref(ppeak))
    pb += po.port().icwo_render(C.byref(sp), C.byref(mt), C.byref(prev), C.byref(nss), _dp(x[h:]), n - h,
                                out[pb:].ctypes.data_as(u8), C.byref(pclips), C.byref(ppeak))
    assert nb == pb
    assert bytes(out[:pb]) == buf.raw[:nb]
    assert clips.value == pclips.value
    assert peak.value == ppeak.value


@pytest.mark.parametrize("fmt", ["wav_u8", "wav_i16", "wav_i24", "wav_i32", "wav_f32",
                                 "cw_f64", "cw_i16", "cw_i16f32", "cw_f32"])
@pytest.mark.parametrize("nch", [1, 2])
def test_every_format_end_to_end(fmt, nch):
    d = S.default_spec(fmt=fmt, n_channels=nch, sample_rate=32000)
    n = 2500
    if fmt in ("wav_u8", "wav_i16", "wav_i24", "wav_i32", "cw_i16"):
        raw = np.random.default_rng(5).integers(0, 256, size=n * S.frame_bytes(d), dtype=np.uint8)
    else:
        raw = synth.stream_bytes(d, n, stream_id=3)
    a = po.port_process(d, raw, taps=[0])
    b = po.ref_process(d, raw, taps=[0])
    assert np.array_equal(a["pcm"], b["pcm"])
    assert np.array_equal(a["bus"], b["bus"])


def test_extensible_wav_header_same_result():
    d = S.config_c1()
    raw = synth.stream_bytes(d, 1200, stream_id=9)
    a = po.ref_process(d, raw)
    b = po.ref_process(dict(d, wav_extensible=1), raw)
    assert np.array_equal(a["pcm"], b["pcm"])


def test_state_survives_files():
    """Frame counter, Hilbert delay lines and the dither stream continue into the next file
    (reference defaults CLR_NFRAME_PT = CLR_HILB_PT = 0, src/config.c:171,174)."""
    d = S.config_c2()
    fb = S.frame_bytes(d)
    raw = synth.stream_bytes(d, 3000, stream_id=11)
    r1 = po.ref_process(d, raw[: 1000 * fb])
    r2 = po.ref_process(d, raw[1000 * fb:], reset=False)
    st = po.new_state()
    p1 = po.port_process(d, raw[: 1000 * fb], state=st)
    st.pos = 0                                  # new file: fade position restarts, the rest persists
    p2 = po.port_process(d, raw[1000 * fb:], state=st)
    assert np.array_equal(r1["pcm"], p1["pcm"]) and np.array_equal(r2["pcm"], p2["pcm"])


@pytest.mark.parametrize("over", [dict(render_type=3, dth_bits=3.25), dict(render_type=2, nshape_type=6), dict(render_type=4, nshape_type=17)])
def test_every_file_open_clears_the_render_memory(over):
    """mod_context_fopen runs sound_render_set_outbits -> sound_render_recalc on both renderers for EVERY file
    (src/in_cwave.c:231-234, src/sound_render.c:509,556-580): the sloped-TPDF memory and the shaper's buffers start from zero in
    each file, while the generators (and everything else) carry on.  What include/icw_b200.h calls ICW_RESET_RENDER_MEMORY."""
    d = S.config_c2(**over)
    fb = S.frame_bytes(d)
    raw = synth.stream_bytes(d, 3000, stream_id=12)
    po.ref_process(d, raw[: 1000 * fb])
    r2 = po.ref_process(d, raw[1000 * fb:], reset=False)
    st = po.new_state()
    po.port_process(d, raw[: 1000 * fb], state=st)
    carried = po.port_process(d, raw[1000 * fb:], state=C.pointer(st).contents.__class__.from_buffer_copy(st))
    st.pos = 0
    for c in range(2):
        st.prev_rnd[c] = 0.0
        C.memset(C.byref(st.ns[c]), 0, C.sizeof(st.ns[c]))
    p2 = po.port_process(d, raw[1000 * fb:], state=st)
    assert np.array_equal(r2["pcm"], p2["pcm"])
    assert not np.array_equal(r2["pcm"], carried["pcm"])        # ... and carrying that memory over is NOT what the reference does


def test_frame_counter_wrap():
    """Scaled counter wraps at sample_rate*1000 frames (src/adv_modulator.c:614-617)."""
    d = S.default_spec(fmt="cw_i16", sample_rate=8, need24bits=0,
                       nodes=[dict(mode="shift", inputs=[0], out=1, l_p=[1.3], r_p=[-0.7]),
                              dict(mode="master", inputs=[1], l_gain=0.9, r_gain=0.9)])
    n = 8 * 1000 + 700
    raw = np.random.default_rng(2).integers(0, 256, size=n * S.frame_bytes(d), dtype=np.uint8)
    a = po.port_process(d, raw, taps=[1])
    b = po.ref_process(d, raw, taps=[1])
    assert a["state"].n_frame == b["stats"].n_frame == 700
    assert np.array_equal(a["bus"], b["bus"]) and np.array_equal(a["pcm"], b["pcm"])


def test_transcode_entry_points_agree():
    """winampGetExtendedRead_* (src/transcode.c:40-118) == the direct frame loop."""
    d = S.config_c1()
    n = 5000
    raw = synth.stream_bytes(d, n, stream_id=21)
    direct = po.ref_process(d, raw)
    import os, tempfile
    cfg = po.make_refcfg(d)
    po.ref().icwref_reset(C.byref(cfg))
    nodes = d["nodes"]
    arr = (po.Node * len(nodes))()
    for i, nd in enumerate(nodes):
        po.fill_node(arr[i], nd)
    assert po.ref().icwref_set_graph(arr, len(nodes), 0) == 0
    fd, path = tempfile.mkstemp(suffix=".wav")
    os.write(fd, po.wav_bytes(d, raw))
    os.close(fd)
    try:
        pcm = np.zeros(n * 6, dtype=np.uint8)
        info = (C.c_int * 4)()
        got = po.ref().icwref_transcode_file(path.encode(), 4096, pcm.ctypes.data_as(C.c_char_p), pcm.size, info)
    finally:
        os.unlink(path)
    assert got == n * 6
    assert list(info) == [n * 6, 24, 2, 48000]
    assert np.array_equal(pcm, direct["pcm"])


def _exceptional(spec, n, seed):
    """Seeded input with NaN, +-Inf and values that make denormal intermediates scattered through it."""
    from in_cwave_b200 import synth
    raw = synth.stream_bytes(spec, n, stream_id=seed).copy()
    f64 = spec["fmt"] == "cw_f64"
    f = raw.view(np.float64 if f64 else np.float32).copy()
    idx = np.random.default_rng(seed).choice(f.size, 60, replace=False)
    vals = [np.nan, np.inf, -np.inf, 3e-312 if f64 else 1e-42, -2e-311 if f64 else -1e-43, 1e300 if f64 else 3e38]
    for j, i in enumerate(idx):
        f[i] = vals[j % len(vals)]
    return f.view(np.uint8)


FP_CHECK_CASES = {
    "hilbert_kahan_tpdf": lambda S: S.config_c1(is_fp_check=1, render_type=2),
    "hilbert_plain_noreject": lambda S: S.config_c1(is_fp_check=1, is_kahan=0, is_subnorm_reject=0),
    "hilbert_type5_mono": lambda S: S.config_c1(is_fp_check=1, filter_no=5, n_channels=1),
    "render_denormals_fir_shaper": lambda S: S.default_spec(
        fmt="cw_f64", is_fp_check=1, render_type=2, nshape_type=6, sample_rate=44100,
        nodes=[dict(mode="master", inputs=[0], l_gain=1.0, r_gain=1.0, l_tout=2, r_tout=3)]),
    # (no trig in the graphs of the shaper cases: a 1-ulp sin/cos difference between libm and libdevice flips an
    # LSB now and then, and error feedback turns one flip into a different sequence)
    "render_iir_shaper_16bit": lambda S: S.default_spec(
        fmt="cw_f32", is_fp_check=1, nshape_type=16, sample_rate=44100, need24bits=0,
        nodes=[dict(mode="master", inputs=[0], l_gain=0.9, r_gain=0.7, l_tout=0, r_tout=1)]),
    "render_flat_gauss": lambda S: S.default_spec(
        fmt="cw_f64", is_fp_check=1, render_type=4,
        nodes=[dict(mode="master", inputs=[0], l_gain=1.0, r_gain=1.0, l_tout=2, r_tout=3)]),
}


@pytest.mark.parametrize("case", sorted(FP_CHECK_CASES))
def test_fp_checked_twins_port_equals_reference(oracle, case):
    """FP_CHECK=1: the checked twins of the half-band filters, the renderer and the shapers flush
    NaN / denormals to 0 and +-Inf to +-65535 and count every event -- same PCM and the same four counter
    blocks as the compiled reference, on inputs full of exceptional values."""
    from in_cwave_b200 import spec as S
    spec = FP_CHECK_CASES[case](S)
    raw = _exceptional(spec, 12000, 11)
    ref = oracle.ref_process(spec, raw)
    port = oracle.port_process(spec, raw)
    assert np.array_equal(ref["pcm"], port["pcm"])
    cnt = [list(r) for r in port["state"].fp_cnt]
    print(case, cnt)
    assert cnt == ref["fp_cnt"]
    assert sum(r[0] for r in cnt) > 0
    # and with the check off the same input is processed unchecked: no counters move
    off = oracle.port_process(dict(spec, is_fp_check=0), raw)
    assert all(v == 0 for r in off["state"].fp_cnt for v in r)


FEEDBACK_NODES = [
    dict(mode="mix", inputs=[0, 7], out=1, l_gain=0.6, r_gain=0.6),            # plug 7 is written by the NEXT node: last frame's value
    dict(mode="shift", inputs=[1], out=7, l_p=[3.5], r_p=[-2.25], l_gain=0.5, r_gain=0.5),
    dict(mode="pm", inputs=[7, 9], out=9, l_p=[2.0, 0.0, 0.25, 0.0], r_p=[2.0, 0.5, 0.25, 0.0], l_gain=0.4, r_gain=0.4),   # reads its own output
    dict(mode="master", inputs=[7, 9], l_gain=0.7, r_gain=0.7),
]


@pytest.mark.parametrize("fmt", ["wav_f32", "cw_f32"])
def test_feedback_graph_port_equals_reference(fmt):
    """A DSP list may read a plug that a later node (or the node itself) writes: the bus persists in MOD_CONTEXT, so the
    value read is the previous frame's (src/adv_modulator.c:634-751).  The port must carry the bus the same way, also
    across calls and files."""
    spec = S.default_spec(fmt=fmt, sample_rate=48000, render_type=2, nodes=FEEDBACK_NODES)
    n = 6000
    raw = synth.stream_bytes(spec, n, stream_id=31)
    plugs = [0, 1, 7, 9]
    r = po.ref_process(spec, raw, read_quant=1111, taps=plugs)
    p = po.port_process(spec, raw, taps=plugs)
    assert np.array_equal(r["pcm"], p["pcm"])
    assert np.array_equal(r["bus"], p["bus"])
    assert np.any(r["bus"][1:, 2, :] != 0.0)
