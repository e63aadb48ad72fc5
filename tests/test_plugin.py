"""The host C layer (include/icw_plugin.h): the reference's transcode entry points over the B200 library.

CPU part: header parsing agrees with the reference's own reader on good and malformed files.
GPU part: the same file through winampGetExtendedRead_open/getData/close of libicw_plugin.so and of
the compiled reference gives the same bytes (exact mode)."""
import ctypes as C
import os
import struct

import numpy as np
import pytest

from in_cwave_b200 import plugin, spec as S, synth
from oracle import pyoracle as po


def _ref_open(path, cfg_over=None):
    """The reference's winampGetExtendedRead_open on a file: (size, bps, nch, srate) or None."""
    d = S.default_spec()
    d.update(cfg_over or {})
    cfg = po.make_refcfg(d)
    po.ref().icwref_reset(C.byref(cfg))
    info = (C.c_int * 4)()
    pcm = np.zeros(16, dtype=np.uint8)
    got = po.ref().icwref_transcode_file(str(path).encode(), 4096, pcm.ctypes.data_as(C.c_char_p), 0, info)
    return None if got < 0 else tuple(info)


def _files(tmp_path):
    rng = np.random.default_rng(3)
    out = {}

    def put(name, blob):
        p = tmp_path / name
        p.write_bytes(blob)
        out[name] = p

    for fmt, nch in (("wav_u8", 1), ("wav_i16", 2), ("wav_i24", 2), ("wav_i32", 1), ("wav_f32", 2)):
        d = dict(fmt=fmt, n_channels=nch, sample_rate=44100)
        raw = rng.integers(0, 256, size=777 * S.frame_bytes(d), dtype=np.uint8)
        put(f"{fmt}_{nch}.wav", po.wav_bytes(d, raw))
        put(f"{fmt}_{nch}_ext.wav", po.wav_bytes(d, raw, extensible=True))
    for fmt in ("cw_f64", "cw_i16", "cw_i16f32", "cw_f32"):
        d = dict(fmt=fmt, n_channels=2, sample_rate=96000)
        raw = rng.integers(0, 256, size=500 * S.frame_bytes(d), dtype=np.uint8)
        put(f"{fmt}.cwave", po.cwave_bytes(d, raw))
        put(f"{fmt}_v1.cwave", po.cwave_bytes(d, raw, version=1))
    good = po.wav_bytes(dict(fmt="wav_i16", n_channels=2, sample_rate=48000), rng.integers(0, 256, size=4000, dtype=np.uint8))
    put("rwave_ext.RWAVE", good)
    put("wrong_ext.mp3", good)
    put("truncated.wav", good[:60])
    put("bad_riff.wav", b"RIFX" + good[4:])
    put("data_before_fmt.wav", good[:12] + good[36:44] + good[44:60] + good[12:36] + good[60:])
    junk = b"LIST" + struct.pack("<I", 10) + b"0123456789"
    put("extra_chunk.wav", good[:12] + junk + good[12:])
    put("rate_too_high.wav", good[:24] + struct.pack("<I", 3_000_000) + good[28:])
    put("three_channels.wav", good[:22] + struct.pack("<H", 3) + good[24:])
    put("one_frame.wav", po.wav_bytes(dict(fmt="wav_i16", n_channels=2, sample_rate=48000), np.zeros(4, dtype=np.uint8)))
    cw = po.cwave_bytes(dict(fmt="cw_f32", n_channels=2, sample_rate=96000), rng.integers(0, 256, size=1600, dtype=np.uint8))
    put("bad_magic.cwave", b"cPLXwAVX" + cw[8:])
    put("bad_version.cwave", cw[:12] + struct.pack("<I", 7) + cw[16:])
    put("short_data.cwave", cw[:-100])
    return out


@pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref/libicw_ref.so not built")
def test_header_parsing_agrees_with_the_reference_reader(tmp_path):
    files = _files(tmp_path)
    accepted = 0
    for name, path in sorted(files.items()):
        ref = _ref_open(path)
        fi = plugin.probe(path)
        assert (ref is None) == (fi is None), f"{name}: reference {'rejects' if ref is None else 'accepts'}, we do not agree"
        if ref is not None:
            accepted += 1
            size, bps, nch, srate = ref
            assert fi.n_samples * 6 == size and fi.sample_rate == srate and nch == 2, name
    assert accepted >= 19


@pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref/libicw_ref.so not built")
def test_tail_and_fade_geometry(tmp_path):
    d = dict(fmt="wav_i16", n_channels=2, sample_rate=8000)
    p = tmp_path / "t.wav"
    p.write_bytes(po.wav_bytes(d, np.zeros(12345 * 4, dtype=np.uint8)))
    fi = plugin.probe(p, sec_align=2, fade_in_ms=100, fade_out_ms=250)
    assert (fi.n_samples, fi.n_tail, fi.n_fade_in, fi.n_fade_out) == (12345, 16000 - 12345, 800, 2000)
    ref = _ref_open(p, dict(sec_align=2))
    assert ref[0] == 16000 * 6
    short = tmp_path / "s.wav"
    short.write_bytes(po.wav_bytes(d, np.zeros(900 * 4, dtype=np.uint8)))
    fi = plugin.probe(short, fade_in_ms=100, fade_out_ms=100)
    assert (fi.n_fade_in, fi.n_fade_out) == (300, 300)          # capped at a third of a short track
    tiny = tmp_path / "y.wav"
    tiny.write_bytes(po.wav_bytes(d, np.zeros(200 * 4, dtype=np.uint8)))
    fi = plugin.probe(tiny, fade_in_ms=100, fade_out_ms=100)
    assert (fi.n_fade_in, fi.n_fade_out) == (0, 0)


def test_plugin_exports_the_reference_symbols():
    L = plugin.lib()
    for name in plugin.EXPORTS:
        assert hasattr(L, name)


def test_input_module_table(tmp_path):
    """winampGetInModule2 (reference src/in_cwave.c:551-572): the SDK's In_Module layout (Winamp/IN2.H:56-154) with the
    members this library serves -- extension list, Init/Quit, GetFileInfo (length incl. the sec_align tail, like
    src/playback.c:getfileinfo), GetLength -- and the playback half present but refusing."""
    L = plugin.lib()
    V = C.c_void_p
    fn = lambda res, *a: C.CFUNCTYPE(res, *a)

    class InModule(C.Structure):
        _fields_ = [("version", C.c_int), ("description", C.c_char_p), ("hMainWindow", V), ("hDllInstance", V),
                    ("FileExtensions", V), ("is_seekable", C.c_int), ("UsesOutputPlug", C.c_int),
                    ("Config", fn(None, V)), ("About", fn(None, V)), ("Init", fn(C.c_int)), ("Quit", fn(None)),
                    ("GetFileInfo", fn(None, C.c_char_p, C.c_char_p, C.POINTER(C.c_int))),
                    ("InfoBox", fn(C.c_int, C.c_char_p, V)), ("IsOurFile", fn(C.c_int, C.c_char_p)),
                    ("Play", fn(C.c_int, C.c_char_p)), ("Pause", fn(None)), ("UnPause", fn(None)), ("IsPaused", fn(C.c_int)),
                    ("Stop", fn(None)), ("GetLength", fn(C.c_int)), ("GetOutputTime", fn(C.c_int)),
                    ("SetOutputTime", fn(None, C.c_int)), ("SetVolume", fn(None, C.c_int)), ("SetPan", fn(None, C.c_int))]
    L.winampGetInModule2.restype = C.POINTER(InModule)
    m = L.winampGetInModule2().contents
    assert m.version == 0x101 and m.is_seekable == 1 and m.UsesOutputPlug == 1
    ext = C.string_at(m.FileExtensions, 80)
    assert ext.startswith(b"cwave\0") and b"\0wav\0" in ext
    assert m.Init() == 0
    d = dict(fmt="wav_i16", n_channels=2, sample_rate=8000)
    f = tmp_path / "len.wav"
    f.write_bytes(po.wav_bytes(d, np.zeros(12345 * 4, dtype=np.uint8)))
    title = C.create_string_buffer(2048)
    ms = C.c_int(0)
    m.GetFileInfo(str(f).encode(), title, C.byref(ms))
    assert title.value == b"len.wav" and ms.value == 12345 * 1000 // 8000
    assert m.GetLength() == ms.value
    m.GetFileInfo(str(tmp_path / "missing.wav").encode(), title, C.byref(ms))
    assert ms.value == -1000 and title.value == b"-CAN'T OPEN-"
    assert m.Play(str(f).encode()) == 2 and m.IsPaused() == 0
    m.Stop(); m.Pause(); m.SetVolume(10); m.Quit()


@pytest.mark.gpu
@pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref/libicw_ref.so not built")
def test_a_reference_config_file_drives_both_plugins_to_the_same_bytes(tmp_path):
    """one config file (written by the reference's save_config), read by the reference's
    load_config on one side and by icwp_load_config on the other; then the same four transcode calls."""
    from util import pcm_report
    spec = S.config_c3(render_type=2, need24bits=1)          # full graph, TPDF, 24 bit
    n = 20000
    raw = synth.stream_bytes(spec, n, stream_id=9)
    path = tmp_path / "x.cwave"
    path.write_bytes(po.cwave_bytes(spec, raw))
    cfg_file = tmp_path / "in_cwave.cfg"
    d = dict(spec, fade_in=200, fade_out=300)
    assert po.ref_save_config(d, cfg_file)

    ok, _, nodes_r = po.ref_load_config(cfg_file)            # fresh reference plugin configured from the file
    assert ok and len(nodes_r) == len(spec["nodes"])
    cap = (n + 1000) * 6
    want = np.zeros(cap, dtype=np.uint8)
    info = (C.c_int * 4)()
    got = po.ref().icwref_transcode_file(str(path).encode(), 4096, want.ctypes.data_as(C.c_char_p), cap, info)
    assert got > 0
    want = want[:got]

    ok, sp, o, nodes_o = plugin.load_config(cfg_file)
    assert ok and nodes_o == nodes_r and (o.fade_in_ms, o.fade_out_ms) == (200, 300)
    plugin.lib().icwp_reset()
    o.readahead_frames = 5000
    assert plugin.lib().icwp_configure(C.byref(sp), C.byref(o)) == 0
    pcm, meta = plugin.transcode(path, chunk=4096)
    assert meta == tuple(info) and pcm.size == want.size
    rep = pcm_report(pcm, want, 3)
    print(f"[plugin from config file] {rep}")
    assert rep["max_lsb"] <= 1 and rep["mismatches"] <= 1


@pytest.mark.gpu
@pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref/libicw_ref.so not built")
@pytest.mark.parametrize("case", ["c1", "c2_tpdf", "c3", "tail_fade"])
def test_transcode_entry_points_match_the_reference(tmp_path, case):
    """Same file, same four calls, same bytes (exact Hilbert mode; trig LSB flips counted)."""
    from util import pcm_report
    n = 30000
    opts, over = {}, {}
    if case == "c1":
        spec = S.config_c1()
    elif case == "c2_tpdf":
        spec = S.config_c2(sample_rate=96000)
    elif case == "c3":
        spec = S.config_c3()
    else:
        spec = S.config_c1(sample_rate=8000)
        opts = dict(sec_align=3, fade_in_ms=500, fade_out_ms=1000)
        over = dict(sec_align=3, fade_in=500, fade_out=1000)
    raw = synth.stream_bytes(spec, n, stream_id=5)
    is_cw = spec["fmt"].startswith("cw_")
    path = tmp_path / ("x.cwave" if is_cw else "x.wav")
    path.write_bytes(po.cwave_bytes(spec, raw) if is_cw else po.wav_bytes(spec, raw))

    # the reference: fresh plugin, its own transcode loop
    d = dict(spec); d.update(over)
    cfg = po.make_refcfg(d)
    po.ref().icwref_reset(C.byref(cfg))
    arr = (po.Node * len(spec["nodes"]))()
    for i, nd in enumerate(spec["nodes"]):
        po.fill_node(arr[i], nd)
    assert po.ref().icwref_set_graph(arr, len(spec["nodes"]), int(spec.get("bypass", 0))) == 0
    cap = (n + 30000) * 6
    want = np.zeros(cap, dtype=np.uint8)
    info = (C.c_int * 4)()
    got = po.ref().icwref_transcode_file(str(path).encode(), 5000, want.ctypes.data_as(C.c_char_p), cap, info)
    assert got > 0
    want = want[:got]

    plugin.lib().icwp_reset()
    plugin.configure(spec, readahead_frames=7001, **opts)
    pcm, meta = plugin.transcode(path, chunk=5000)
    assert meta == tuple(info)
    assert pcm.size == want.size
    rep = pcm_report(pcm, want, 3 if spec.get("need24bits", 1) else 2)
    print(f"[plugin {case}] {rep}")
    assert rep["max_lsb"] <= 1 and rep["mismatches"] <= 1


@pytest.mark.gpu
@pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref/libicw_ref.so not built")
def test_state_survives_into_the_next_file(tmp_path):
    """Two files back to back: counter, delay lines and dither stream continue (reference defaults)."""
    spec = S.config_c2(sample_rate=48000)
    fb = S.frame_bytes(spec)
    raw = synth.stream_bytes(spec, 9000, stream_id=8)
    a, b = tmp_path / "a.wav", tmp_path / "b.wav"
    a.write_bytes(po.wav_bytes(spec, raw[: 4000 * fb]))
    b.write_bytes(po.wav_bytes(spec, raw[4000 * fb:]))
    r1 = po.ref_process(spec, raw[: 4000 * fb])
    r2 = po.ref_process(spec, raw[4000 * fb:], reset=False)
    plugin.lib().icwp_reset()
    plugin.configure(spec)
    p1, _ = plugin.transcode(a)
    p2, _ = plugin.transcode(b)
    assert np.array_equal(p1, r1["pcm"]) and np.array_equal(p2, r2["pcm"])


@pytest.mark.gpu
def test_cwave_crc_check_like_the_reference_info_dialog(tmp_path):
    """check_cwave (src/gui_cwave.c:82-130) -- CRC-32 of the sample data against the V2 header."""
    import zlib
    spec = S.config_c3()
    raw = synth.stream_bytes(spec, 70001, stream_id=3)
    good = zlib.crc32(np.ascontiguousarray(raw).tobytes())
    f = tmp_path / "ok.cwave"
    f.write_bytes(po.cwave_bytes(spec, raw, crc=good) + b"trailing bytes are not sample data")
    assert plugin.check_cwave(f) == (True, good, good, True)
    bad = bytearray(po.cwave_bytes(spec, raw, crc=good))
    bad[48 + 12345] ^= 0x40
    g = tmp_path / "bad.cwave"
    g.write_bytes(bytes(bad))
    ok, calc, filec, has = plugin.check_cwave(g)
    assert ok and has and filec == good and calc != good
    v1 = tmp_path / "v1.cwave"
    v1.write_bytes(po.cwave_bytes(spec, raw, version=1))
    ok, calc, filec, has = plugin.check_cwave(v1)
    assert ok and not has and calc == good
    assert plugin.check_cwave(tmp_path / "missing.cwave")[0] is False


# ---- read-ahead must be invisible: seek and early close (reference src/transcode.c:82-118) --------------------
def _bind_transcode(L):
    ip = C.POINTER(C.c_int)
    L.winampGetExtendedRead_open.argtypes = [C.c_char_p, ip, ip, ip, ip]
    L.winampGetExtendedRead_open.restype = C.c_ssize_t
    L.winampGetExtendedRead_getData.argtypes = [C.c_ssize_t, C.c_void_p, C.c_int, ip]
    L.winampGetExtendedRead_getData.restype = C.c_ssize_t
    L.winampGetExtendedRead_setTime.argtypes = [C.c_ssize_t, C.c_int]
    L.winampGetExtendedRead_setTime.restype = C.c_int
    L.winampGetExtendedRead_close.argtypes = [C.c_ssize_t]
    L.winampGetExtendedRead_close.restype = None
    return L


def _run_script(L, script):
    """script: list of ("open", path) / ("get", bytes, times) / ("seek", ms) / ("drain", bytes) / ("close",).
    The same four exported symbols on either library; returns every byte served, in order."""
    out = bytearray()
    h, kill = 0, C.c_int(0)
    info = (C.c_int * 4)()
    for op in script:
        if op[0] == "open":
            h = L.winampGetExtendedRead_open(str(op[1]).encode(), *[C.byref(C.c_int.from_buffer(info, 4 * i)) for i in range(4)])
            assert h, f"open failed: {op[1]}"
        elif op[0] == "get":
            buf = (C.c_char * op[1])()
            for _ in range(op[2]):
                got = L.winampGetExtendedRead_getData(h, buf, op[1], C.byref(kill))
                out += buf.raw[:got]
        elif op[0] == "seek":
            assert L.winampGetExtendedRead_setTime(h, op[1])
        elif op[0] == "drain":
            buf = (C.c_char * op[1])()
            while True:
                got = L.winampGetExtendedRead_getData(h, buf, op[1], C.byref(kill))
                if got <= 0:
                    break
                out += buf.raw[:got]
        elif op[0] == "close":
            L.winampGetExtendedRead_close(h)
            h = 0
    return np.frombuffer(bytes(out), dtype=np.uint8)


def _both(spec, script, readahead, opts=None, over=None):
    from util import pcm_report
    d = dict(spec); d.update(over or {})
    cfg = po.make_refcfg(d)
    po.ref().icwref_reset(C.byref(cfg))
    arr = (po.Node * len(spec["nodes"]))()
    for i, nd in enumerate(spec["nodes"]):
        po.fill_node(arr[i], nd)
    assert po.ref().icwref_set_graph(arr, len(spec["nodes"]), int(spec.get("bypass", 0))) == 0
    want = _run_script(_bind_transcode(po.ref()), script)
    plugin.lib().icwp_reset()
    plugin.configure(spec, readahead_frames=readahead, **(opts or {}))
    got = _run_script(plugin.lib(), script)
    assert got.size == want.size and got.size > 0
    rep = pcm_report(got, want, 3 if spec.get("need24bits", 1) else 2)
    ref_stats = po.ref_stats() if hasattr(po, "ref_stats") else None
    return rep, ref_stats


@pytest.mark.gpu
@pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref/libicw_ref.so not built")
@pytest.mark.parametrize("readahead", [7001, 0])
@pytest.mark.parametrize("mode", ["exact_tpdf", "cwave_graph", "fade_tail"])
def test_seek_in_the_middle_of_a_read_ahead_block(tmp_path, readahead, mode):
    """getData x k -> setTime -> getData ...: the reference moves only the reader (src/xwave_reader.c:782-808), its
    oscillator counter, Hilbert memory and dither stream stand at the frames actually served.  Our read-ahead
    renders up to a block beyond that; the seek must undo it."""
    opts, over = {}, {}
    if mode == "exact_tpdf":
        spec = S.config_c2(sample_rate=48000)
    elif mode == "cwave_graph":
        spec = S.config_c3(render_type=2, need24bits=1)
    else:
        spec = S.config_c1(sample_rate=8000, render_type=1)
        opts = dict(sec_align=3, fade_in_ms=500, fade_out_ms=1000)
        over = dict(sec_align=3, fade_in=500, fade_out=1000)
    n = 30000
    raw = synth.stream_bytes(spec, n, stream_id=11)
    is_cw = spec["fmt"].startswith("cw_")
    path = tmp_path / ("x.cwave" if is_cw else "x.wav")
    path.write_bytes(po.cwave_bytes(spec, raw) if is_cw else po.wav_bytes(spec, raw))
    ms = lambda fr: fr * 1000 // spec["sample_rate"]
    script = [("open", path), ("get", 4998, 5), ("seek", ms(20000)), ("get", 6000, 3), ("seek", ms(1000)), ("get", 600, 7),
              ("seek", ms(29000)), ("drain", 4096), ("close",)]
    rep, _ = _both(spec, script, readahead, opts, over)
    print(f"[seek {mode} ra={readahead}] {rep}")
    assert rep["max_lsb"] <= 1 and rep["mismatches"] <= 2


@pytest.mark.gpu
@pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref/libicw_ref.so not built")
@pytest.mark.parametrize("readahead", [5000, 0])
def test_cancelled_transcode_then_the_next_file(tmp_path, readahead):
    """getData x k -> close -> open the next file: the context survives across files (src/config.c:171,174), so the
    second file's bytes depend on exactly how many frames of the first were rendered."""
    spec = S.config_c2(sample_rate=48000)
    fb = S.frame_bytes(spec)
    raw = synth.stream_bytes(spec, 24000, stream_id=12)
    a, b = tmp_path / "a.wav", tmp_path / "b.wav"
    a.write_bytes(po.wav_bytes(spec, raw[: 14000 * fb]))
    b.write_bytes(po.wav_bytes(spec, raw[14000 * fb:]))
    script = [("open", a), ("get", 4092, 4), ("close",), ("open", b), ("drain", 8190), ("close",)]
    rep, _ = _both(spec, script, readahead)
    print(f"[cancel ra={readahead}] {rep}")
    assert rep["mismatches"] == 0


@pytest.mark.gpu
def test_overlapped_reader_prefetches_blocks(tmp_path):
    """N1: block k+1 is read by the reader thread while block k is rendered; only the first block of a file
    (and the first after a seek) is read by the caller."""
    spec = S.config_c2(sample_rate=48000, hilbert_mode="scan")
    raw = synth.stream_bytes(spec, 200_000, stream_id=13)
    path = tmp_path / "long.wav"
    path.write_bytes(po.wav_bytes(spec, raw))
    plugin.lib().icwp_reset()
    plugin.configure(spec, readahead_frames=32768)
    plugin.io_stats(reset=True)
    pcm, _ = plugin.transcode(path, chunk=65536)
    st = plugin.io_stats()
    assert pcm.size == 200_000 * 6
    assert st["blocks_sync"] == 1 and st["blocks_prefetched"] == (200_000 + 32767) // 32768 - 1
    assert st["frames"] == 200_000 and st["read_bytes"] == raw.size and st["resettles"] == 0


def random_transcode_script(tmp_path, seed):
    """A random configuration, two files of it on disk and a random walk over the four entry points."""
    from util import random_spec
    rng = np.random.default_rng(7000 + seed)
    spec = random_spec(rng)
    fb = S.frame_bytes(spec)
    is_cw = spec["fmt"].startswith("cw_")
    n = [int(rng.integers(3000, 40000)) for _ in range(2)]
    raw = synth.stream_bytes(spec, n[0] + n[1], stream_id=int(rng.integers(1, 1 << 30)))
    paths = []
    for j, (a, b) in enumerate(((0, n[0]), (n[0], n[0] + n[1]))):
        p = tmp_path / (f"f{j}.cwave" if is_cw else f"f{j}.wav")
        p.write_bytes(po.cwave_bytes(spec, raw[a * fb:b * fb]) if is_cw else po.wav_bytes(spec, raw[a * fb:b * fb]))
        paths.append(p)
    script = []
    for j, p in enumerate(paths):
        script.append(("open", p))
        dur_ms = n[j] * 1000 // spec["sample_rate"]
        for _ in range(int(rng.integers(1, 7))):
            if rng.random() < 0.3 and dur_ms > 1:
                script.append(("seek", int(rng.integers(0, dur_ms))))
            else:
                script.append(("get", int(rng.integers(6, 9000)), int(rng.integers(1, 7))))
        if j == 1 or rng.random() < 0.5:
            script.append(("drain", int(rng.integers(600, 70000))))
        script.append(("close",))
    # four walks in ten also get the reader's options: a silence tail up to whole seconds and fades (src/xwave_reader.c:693-723)
    geo = {}
    if rng.random() < 0.4:
        geo = dict(sec_align=int(rng.integers(1, 4)), fade_in=int(rng.integers(0, 400)), fade_out=int(rng.integers(0, 400)))
    return spec, script, rng, geo


@pytest.mark.gpu
@pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref/libicw_ref.so not built")
@pytest.mark.parametrize("seed", range(int(os.environ.get("ICW_PLUGIN_FUZZ", "16"))))     # a longer walk: ICW_PLUGIN_FUZZ=300
def test_random_transcode_scripts(tmp_path, seed):
    """Random configuration (tests/util.py::random_spec), two files, a random walk over the four entry points -- requests of any
    byte count, seeks anywhere, a transcode cancelled half way and the next file opened on the surviving context -- through
    libicw_plugin.so and through the compiled reference: the same bytes in the same order (src/transcode.c:40-118)."""
    spec, script, rng, geo = random_transcode_script(tmp_path, seed)
    readahead = int(rng.choice([0, 1, 777, 7001, 1 << 20]))
    opts = dict(sec_align=geo["sec_align"], fade_in_ms=geo["fade_in"], fade_out_ms=geo["fade_out"]) if geo else None
    rep, _ = _both(spec, script, readahead, opts, geo or None)
    print(f"[random script {seed} ra={readahead}] {rep} {[op[:1] + op[2:] if op[0] == 'open' else op for op in script]}")
    trig = not spec["bypass"] and any(nd["mode"] in ("shift", "pm") for nd in spec["nodes"])
    if trig:
        step = 1 << ((24 - spec["sign_bits24"]) if spec["need24bits"] else (16 - spec["sign_bits16"]))
        assert rep["max_lsb"] <= step and rep["mismatches"] <= max(2, rep["samples"] // 50000), (seed, rep)
    else:
        assert rep["mismatches"] == 0, (seed, rep)


def _damage_header(rng, blob, is_cw):
    """One to three random edits inside (or at the end of) a file's header: a byte replaced, a 16/32-bit field set to an edge
    value, bytes dropped or inserted, the file cut short."""
    b = bytearray(blob)
    for _ in range(int(rng.integers(1, 4))):
        hdr = min(len(b), 48 if is_cw else 80)
        if hdr < 12:
            break
        kind = int(rng.integers(0, 6))
        if kind == 0:
            b[int(rng.integers(0, hdr))] = int(rng.integers(0, 256))
        elif kind == 1:
            at = int(rng.integers(0, hdr - 4))
            b[at:at + 4] = struct.pack("<I", int(rng.choice([0, 1, 2, 3, 7, 8, 16, 24, 32, 48, 0xFFFF, 0x10000, 0x7FFFFFFF, 0x80000000, 0xFFFFFFFF,
                                                             len(b), max(len(b) - 44, 0), max(len(b) - 8, 0), 2_000_000, 2_000_001])))
        elif kind == 2:
            at = int(rng.integers(0, hdr - 2))
            b[at:at + 2] = struct.pack("<H", int(rng.choice([0, 1, 2, 3, 4, 6, 8, 16, 24, 32, 64, 0xFFFE, 0xFFFF])))
        elif kind == 3:
            at = int(rng.integers(0, hdr))
            del b[at:at + int(rng.integers(1, 9))]
        elif kind == 4:
            at = int(rng.integers(0, hdr))
            b[at:at] = bytes(int(v) for v in rng.integers(0, 256, size=int(rng.integers(1, 9))))
        else:
            del b[int(rng.integers(8, len(b))):]
    return bytes(b)


@pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref/libicw_ref.so not built")
@pytest.mark.parametrize("seed", range(12))
def test_damaged_headers_get_the_reference_readers_verdict(tmp_path, seed):
    """Random damage to WAV (plain and extensible) and CWAVE (V1, V2) headers: icwp_probe accepts exactly the files the
    reference's reader accepts (src/xwave_reader.c:243-585), with the same length, rate and output size."""
    rng = np.random.default_rng(600 + seed)
    agree = accepted = 0
    for trial in range(60):
        is_cw = bool(rng.integers(0, 2))
        if is_cw:
            d = dict(fmt=str(rng.choice(["cw_f64", "cw_i16", "cw_i16f32", "cw_f32"])), n_channels=int(rng.integers(1, 3)), sample_rate=96000)
            raw = rng.integers(0, 256, size=int(rng.integers(2, 400)) * S.frame_bytes(d), dtype=np.uint8)
            blob = po.cwave_bytes(d, raw, version=int(rng.integers(1, 3)))
        else:
            d = dict(fmt=str(rng.choice(["wav_u8", "wav_i16", "wav_i24", "wav_i32", "wav_f32"])), n_channels=int(rng.integers(1, 3)), sample_rate=44100)
            raw = rng.integers(0, 256, size=int(rng.integers(2, 400)) * S.frame_bytes(d), dtype=np.uint8)
            blob = po.wav_bytes(d, raw, extensible=bool(rng.integers(0, 2)))
        p = tmp_path / ("d.cwave" if is_cw else "d.wav")
        p.write_bytes(_damage_header(rng, blob, is_cw))
        ref = _ref_open(p)
        fi = plugin.probe(p)
        assert (ref is None) == (fi is None), (seed, trial, "reference " + ("rejects" if ref is None else "accepts"), p.read_bytes()[:96].hex())
        agree += 1
        if ref is not None:
            accepted += 1
            size, bps, nch, srate = ref
            assert (fi.n_samples + fi.n_tail) * 6 == size and fi.sample_rate == srate, (seed, trial, p.read_bytes()[:96].hex())
    assert agree == 60 and accepted > 0


@pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref/libicw_ref.so not built")
def test_random_track_geometry_against_the_reference(tmp_path):
    """Silence tail up to a whole number of SEC_ALIGN seconds (src/xwave_reader.c:693-703): the output size the reference's
    open reports for random lengths, rates and alignments is the one icwp_probe computes."""
    rng = np.random.default_rng(77)
    for trial in range(120):
        sr = int(rng.choice([8000, 11025, 22050, 44100, 48000, 96000]))
        d = dict(fmt=str(rng.choice(["wav_u8", "wav_i16", "wav_i24"])), n_channels=int(rng.integers(1, 3)), sample_rate=sr)
        n = int(rng.choice([2, 3, sr - 1, sr, sr + 1, int(rng.integers(2, 4 * sr))]))
        align = int(rng.integers(0, 21))
        p = tmp_path / "g.wav"
        p.write_bytes(po.wav_bytes(d, np.zeros(n * S.frame_bytes(d), dtype=np.uint8)))
        ref = _ref_open(p, dict(sec_align=align))
        fi = plugin.probe(p, sec_align=align, fade_in_ms=int(rng.integers(0, 300)), fade_out_ms=int(rng.integers(0, 300)))
        assert ref is not None and fi is not None, (trial, d, n, align)
        assert (fi.n_samples, (fi.n_samples + fi.n_tail) * 6) == (n, ref[0]), (trial, d, n, align, fi.n_tail, ref)
        assert fi.n_fade_in + fi.n_fade_out < max(n, 1) or (fi.n_fade_in, fi.n_fade_out) == (0, 0)
