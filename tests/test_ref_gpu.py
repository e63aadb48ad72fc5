"""The seam as a linked artefact: the reference's OWN sources (reader, transcode entry points, config, DSP-list editor,
reset functions -- compiled in place from /root/reference) with the frame loop of amod_process_samples
(src/adv_modulator.c:587-763) taken over by in_cwave_b200/host/adv_modulator_gpu.c, which calls libicw_b200.so.
oracle/_ref_gpu/libicw_ref_gpu.so against oracle/_ref/libicw_ref.so (the pure reference): same files, same calls of
winampGetExtendedRead_open / _getData / _setTime / _close (src/transcode.c:40-118), same bytes."""
import ctypes as C

import numpy as np
import pytest

from in_cwave_b200 import spec as S, synth
from oracle import pyoracle as po
from test_plugin import _bind_transcode, _run_script, random_transcode_script
from util import pcm_report

pytestmark = [pytest.mark.gpu,
              pytest.mark.skipif(not (po.have_ref() and po.have_ref_gpu()), reason="oracle/_ref or oracle/_ref_gpu not built")]


def _configure(L, spec, over=None):
    d = dict(spec); d.update(over or {})
    cfg = po.make_refcfg(d)
    L.icwref_reset(C.byref(cfg))
    _set_graph(L, spec["nodes"])
    return cfg


def _set_graph(L, nodes):
    arr = (po.Node * len(nodes))()
    for i, nd in enumerate(nodes):
        po.fill_node(arr[i], nd)
    assert L.icwref_set_graph(arr, len(nodes), 0) == 0


def _stats(L):
    st = po.RefStats()
    L.icwref_get_stats(C.byref(st), 0)
    return (st.l_clips, st.r_clips, st.l_peak, st.r_peak, st.subnorm_cnt, st.n_frame)


def _write(tmp_path, name, spec, raw):
    is_cw = spec["fmt"].startswith("cw_")
    path = tmp_path / (name + (".cwave" if is_cw else ".wav"))
    path.write_bytes(po.cwave_bytes(spec, raw) if is_cw else po.wav_bytes(spec, raw))
    return path


@pytest.mark.parametrize("mode", ["exact_tpdf", "cwave_graph", "fade_tail_16bit", "shaper_stpdf", "baseline_type3_gauss"])
def test_reference_entry_points_over_the_gpu_arithmetic(tmp_path, mode):
    """open -> getData x k -> setTime -> getData ... -> close -> next file, through the reference's own transcode.c and
    xwave_reader.c on both builds.  The context survives the first file (src/config.c:171,174), so the second file's
    bytes depend on every bit of state the stub hands back to MOD_CONTEXT."""
    over = {}
    if mode == "exact_tpdf":
        spec = S.config_c2(sample_rate=48000)
    elif mode == "cwave_graph":
        spec = S.config_c3(render_type=2, need24bits=1)
    elif mode == "fade_tail_16bit":
        spec = S.config_c1(sample_rate=8000, render_type=1, need24bits=0)
        over = dict(sec_align=3, fade_in=500, fade_out=1000)
    elif mode == "shaper_stpdf":
        spec = S.config_c1(sample_rate=44100, render_type=3, nshape_type=7, fmt="wav_i16")
    else:
        spec = S.config_c1(sample_rate=48000, render_type=4, filter_no=3, is_kahan=0, fmt="wav_i24")
    n = 30000
    raw = synth.stream_bytes(spec, n, stream_id=21)
    fb = S.frame_bytes(spec)
    a = _write(tmp_path, "a", spec, raw[: 18000 * fb])
    b = _write(tmp_path, "b", spec, raw[18000 * fb:])
    ms = lambda fr: fr * 1000 // spec["sample_rate"]
    script = [("open", a), ("get", 4998, 5), ("seek", ms(12000)), ("get", 6000, 3), ("seek", ms(1000)), ("get", 600, 7),
              ("close",), ("open", b), ("get", 4096, 2), ("seek", ms(11000)), ("drain", 4092), ("close",)]
    ref, gpu = po.ref(), po.ref_gpu()
    _configure(ref, spec, over)
    want = _run_script(_bind_transcode(ref), script)
    _configure(gpu, spec, over)
    got = _run_script(_bind_transcode(gpu), script)
    assert got.size == want.size and got.size > 0, gpu.amod_gpu_last_error()
    rep = pcm_report(got, want, 3 if spec.get("need24bits", 1) else 2)
    print(f"[ref over gpu: {mode}] {rep}  stats ref {_stats(ref)} gpu {_stats(gpu)}")
    # bit-exact chain; the only licence is the <= 2 ulp between libdevice's and glibc's sincos
    assert rep["max_lsb"] <= 1 and rep["mismatches"] <= 2
    sr, sg = _stats(ref), _stats(gpu)
    assert sr[:2] == sg[:2] and sr[4:] == sg[4:]
    assert abs(sr[2] - sg[2]) < 1e-9 and abs(sr[3] - sg[3]) < 1e-9


def test_live_parameter_changes_take_effect_at_the_next_block(tmp_path):
    """amod_add_lastdsp (through the list editor), srenders_set_vcfg and the two reset functions between getData calls:
    the GUI thread's writes on a live plugin (src/amod_gui_control.c:259-312,1555-1602).  The stub reads the plugin's
    objects at every block, so both builds must change course at the same frame."""
    spec = S.config_c2(sample_rate=48000)
    raw = synth.stream_bytes(spec, 40000, stream_id=22)
    path = _write(tmp_path, "live", spec, raw)
    graph2 = S.config_c3()["nodes"]                      # multi-shift + PM + mix
    out = []
    for L in (po.ref(), po.ref_gpu()):
        cfg = _configure(L, spec)
        T = _bind_transcode(L)
        got = bytearray()
        h, kill = 0, C.c_int(0)
        info = (C.c_int * 4)()
        h = T.winampGetExtendedRead_open(str(path).encode(), *[C.byref(C.c_int.from_buffer(info, 4 * i)) for i in range(4)])
        assert h
        buf = (C.c_char * 6000)()

        def pull(k):
            for _ in range(k):
                g = T.winampGetExtendedRead_getData(h, buf, 6000, C.byref(kill))
                got.extend(buf.raw[:g])
        pull(4)
        _set_graph(L, graph2)                           # a new DSP list under a running transcode
        pull(4)
        cfg.render_type = 3; cfg.dth_bits = 1.5; cfg.sign_bits24 = 20; cfg.nshape_type = 4
        L.icwref_set_render_live(C.byref(cfg))          # sloped TPDF, 20 significant bits, a FIR shaper
        pull(4)
        L.icwref_reset_live(1, 0)                       # mod_context_reset_hilbert
        pull(3)
        L.icwref_reset_live(0, 1)                       # mod_context_reset_framecnt
        cfg.render_type = 0; cfg.nshape_type = 0; cfg.sign_bits24 = 24
        L.icwref_set_render_live(C.byref(cfg))
        pull(3)
        T.winampGetExtendedRead_close(h)
        out.append(np.frombuffer(bytes(got), dtype=np.uint8))
    want, got = out
    assert got.size == want.size and got.size == 18 * 6000
    rep = pcm_report(got, want, 3)
    print(f"[ref over gpu: live changes] {rep}")
    assert rep["max_lsb"] <= 1 and rep["mismatches"] <= 2


def test_clip_and_peak_counters_through_the_replaced_getter(tmp_path):
    """amod_get_clips_peaks (src/adv_modulator.c:445-465) is the second function the stub replaces: a signal driven into
    the rails must count the same clips and show the same peak, and isReset must clear both."""
    spec = S.config_c1(sample_rate=48000, render_type=0, nodes=[dict(mode="master", inputs=[0], l_gain=2.0, r_gain=0.5)])
    raw = synth.stream_bytes(spec, 20000, stream_id=23, level=0.9)      # hot input, master at the reference's MAX_GAIN
    stats = []
    for L in (po.ref(), po.ref_gpu()):
        r = po.ref_process(spec, raw, read_quant=3000, lib=L)
        st = r["stats"]
        stats.append((st.l_clips, st.r_clips, st.l_peak, st.r_peak))
        z = po.RefStats()
        L.icwref_get_stats(C.byref(z), 1)
        L.icwref_get_stats(C.byref(z), 0)
        assert (z.l_clips, z.r_clips) == (0, 0) and z.l_peak == -555.0
        stats.append(r["pcm"])
    (sr, pr, sg, pg) = stats
    print(f"[ref over gpu: counters] ref {sr} gpu {sg}")
    assert sr[0] > 0 and sr[:2] == sg[:2]
    assert abs(sr[2] - sg[2]) < 1e-9 and abs(sr[3] - sg[3]) < 1e-9
    assert np.array_equal(pr, pg)


@pytest.mark.parametrize("seed", range(100, 112))
def test_random_transcode_scripts_through_the_seam(tmp_path, seed):
    """The random walk of tests/test_plugin.py::test_random_transcode_scripts over the reference's own reader and entry
    points with the frame loop on the GPU (adv_modulator_gpu.c) against the pure reference."""
    spec, script, _, geo = random_transcode_script(tmp_path, seed)
    out = []
    for L in (po.ref(), po.ref_gpu()):
        cfg = po.make_refcfg(dict(spec, **geo))
        L.icwref_reset(C.byref(cfg))
        arr = (po.Node * len(spec["nodes"]))()
        for i, nd in enumerate(spec["nodes"]):
            po.fill_node(arr[i], nd)
        assert L.icwref_set_graph(arr, len(spec["nodes"]), int(spec.get("bypass", 0))) == 0
        out.append(_run_script(_bind_transcode(L), script))
    want, got = out
    assert got.size == want.size and got.size > 0
    rep = pcm_report(got, want, 3 if spec["need24bits"] else 2)
    trig = not spec["bypass"] and any(nd["mode"] in ("shift", "pm") for nd in spec["nodes"])
    if trig:
        step = 1 << ((24 - spec["sign_bits24"]) if spec["need24bits"] else (16 - spec["sign_bits16"]))
        assert rep["max_lsb"] <= step and rep["mismatches"] <= max(2, rep["samples"] // 50000), (seed, rep)
    else:
        assert rep["mismatches"] == 0, (seed, rep)
