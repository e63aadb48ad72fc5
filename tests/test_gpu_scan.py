"""Scan mode (time-parallel modal evaluation of the half-band recurrences).

Bars: the analytic signal is held to 1e-12 (relative to its RMS / peak) against the binary128
evaluation of the reference's filter (oracle.hilbert_truth) -- the reference's own FP64 sequence is
1e-9 .. 1e-3 away from that truth, so it is NOT the yardstick here.  Everything
after the converter is the same code as exact mode: fed the GPU's own analytic signal, the oracle
must reproduce the rendered PCM byte for byte (LSB flips from 1-ulp trig counted as usual).
"""
import ctypes as C

import numpy as np
import pytest

from in_cwave_b200 import _abi, spec as S
from util import pcm_report, rand_bytes, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12


def truth_iq(oracle, x, ft, kahan, quad0=0):
    return oracle.hilbert_truth(x, ft, 1 if kahan else 0, quad0)


@pytest.mark.parametrize("ft", range(6))
@pytest.mark.parametrize("kahan", [1, 0])
def test_scan_leaf_matches_exact_arithmetic(engine, oracle, ft, kahan):
    import torch
    rng = np.random.default_rng(300 + ft)
    n = 70001                                       # several chunks, two tiles, odd length
    x = (rng.random((3, n)) - 0.5) * 30000.0
    x[1] += 9000.0 * np.sin(2 * np.pi * 0.01 * np.arange(n))
    out, st = engine.hilbert(torch.from_numpy(x).cuda(), ft, kahan, 0, "scan")
    out = out.cpu().numpy()
    for c in range(3):
        ti, tq = truth_iq(oracle, x[c], ft, kahan)
        scale = np.sqrt(np.mean(ti ** 2 + tq ** 2))
        err = max(np.max(np.abs(out[c, :, 0] - ti)), np.max(np.abs(out[c, :, 1] - tq))) / scale
        print(f"[scan leaf] type {ft} kahan {kahan} ch {c}: max err / rms = {err:.2e}")
        assert err <= TOL
        assert st[c].quad[0] == n % 4 and st[c].hb_basis == 1


@pytest.mark.parametrize("splits", [[1], [2, 255], [257, 1000, 4097], [32768, 32769], [65535, 3]])
def test_scan_state_carry_any_split(engine, oracle, splits):
    """Calls of odd and even lengths, across chunk and tile boundaries, continue the same stream."""
    import torch
    n = sum(splits) + 5000
    x = (np.random.default_rng(5).random((1, n)) - 0.5) * 20000.0
    xa = torch.from_numpy(x).cuda()
    ti, tq = truth_iq(oracle, x[0], 1, 1)
    scale = np.sqrt(np.mean(ti ** 2 + tq ** 2))
    pieces, st, pos = [], None, 0
    for ln in splits + [n - sum(splits)]:
        o, st = engine.hilbert(xa[:, pos:pos + ln].contiguous(), 1, 1, 0, "scan", states=st)
        pieces.append(o.cpu().numpy())
        pos += ln
    got = np.concatenate(pieces, axis=1)[0]
    err = max(np.max(np.abs(got[:, 0] - ti)), np.max(np.abs(got[:, 1] - tq))) / scale
    print(f"[scan carry] splits {splits}: max err / rms = {err:.2e}")
    assert err <= TOL


def test_scan_is_closer_to_truth_than_the_reference(engine, oracle):
    """The yardstick argument in numbers: |scan - truth| << |reference - truth| for the default design."""
    import torch
    n = 60000
    x = (np.random.default_rng(9).random(n) - 0.5) * 20000.0
    ti, tq = truth_iq(oracle, x, 1, 1)
    scale = np.sqrt(np.mean(ti ** 2 + tq ** 2))
    out, _ = engine.hilbert(torch.from_numpy(x[None]).cuda(), 1, 1, 0, "scan")
    out = out.cpu().numpy()[0]
    pi, pq = np.zeros(n), np.zeros(n)
    lpf = (oracle.Iir * 2)()
    quad = C.c_uint(0)
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    oracle.port().icwo_hilbert(1, 1, 0, lpf, C.byref(quad), dp(x), n, dp(pi), dp(pq))
    e_scan = np.sqrt(np.mean((out[:, 0] - ti) ** 2 + (out[:, 1] - tq) ** 2)) / scale
    e_ref = np.sqrt(np.mean((pi - ti) ** 2 + (pq - tq) ** 2)) / scale
    print(f"rms error vs binary128 truth: scan {e_scan:.2e}, reference FP64 sequence {e_ref:.2e}")
    assert e_scan < 1e-13 and e_ref > 1e-6


@pytest.mark.parametrize("cfg", ["c1", "c2"])
def test_scan_chain_end_to_end(engine, oracle, cfg):
    """Whole chain in scan mode: analytic vs truth, then PCM vs the oracle fed that analytic signal."""
    spec = (S.config_c1 if cfg == "c1" else S.config_c2)(hilbert_mode="scan")
    n = 50003
    raw = rand_bytes(spec, n, 41)
    ses = engine.session(spec, 1)
    bus, lr = ses.enable_taps(n)
    pcm = ses.process_host(raw)[0]
    ana = bus.cpu().numpy()[0, :, 0, :]
    # truth for both channels from the unpacked samples
    un = np.zeros((n, 4))
    oracle.port().icwo_unpack(oracle.FMT[spec["fmt"]], 2, raw.ctypes.data_as(C.POINTER(C.c_uint8)), n,
                              un.ctypes.data_as(C.POINTER(C.c_double)))
    for ch in range(2):
        ti, tq = truth_iq(oracle, np.ascontiguousarray(un[:, 2 * ch]), 1, 1)
        scale = np.sqrt(np.mean(ti ** 2 + tq ** 2))
        err = max(np.max(np.abs(ana[:, 2 * ch] - ti)), np.max(np.abs(ana[:, 2 * ch + 1] - tq))) / scale
        print(f"[{cfg} scan] channel {ch}: analytic max err / rms = {err:.2e}")
        assert err <= TOL
    # downstream of the converter nothing changed: the oracle on the GPU's analytic must give the GPU's PCM
    spec_cw = dict(spec, fmt="cw_f64")
    ref = oracle.port_process(spec_cw, np.ascontiguousarray(ana).view(np.uint8).ravel())
    rep = pcm_report(pcm, ref["pcm"], 3)
    print(f"[{cfg} scan] PCM vs oracle-on-GPU-analytic: {rep}")
    assert rep["max_lsb"] <= 1 and rep["mismatches"] <= 1
    st = ses.get_state(0)
    assert st.hb_basis == 1 and st.n_frame == n and st.quad[0] == n % 4


@pytest.mark.parametrize("ft", range(6))
def test_scan_pcm_distance_from_the_reference_is_the_reference_noise(engine, oracle, ft):
    """Scan mode against the REFERENCE's bytes: the distance is counted and held to the reference's
    own rounding noise for that design (tests/noise_floor.py, re-measured on CPU by tests/test_truth_pin.py): RMS of the
    PCM difference <= 3 x noise x RMS of the PCM (+ half an LSB of quantisation), largest single difference <= 30 x."""
    from noise_floor import REF_NOISE_RMS
    from util import pcm_to_int
    spec = S.config_c2(hilbert_mode="scan", filter_no=ft)
    n = 50003
    raw = rand_bytes(spec, n, 41)
    pcm = engine.session(spec, 1).process_host(raw)[0]
    want = oracle.port_process(dict(spec, hilbert_mode="exact"), raw)["pcm"]
    g, w = pcm_to_int(pcm, 3).astype(np.float64), pcm_to_int(want, 3).astype(np.float64)
    sig = float(np.sqrt(np.mean(w ** 2)))
    d = np.abs(g - w)
    rms = float(np.sqrt(np.mean(d ** 2)))
    rep = dict(samples=int(g.size), mismatches=int(np.count_nonzero(d)), max_lsb=int(d.max()), rms_lsb=rms, rms_vs_reference=rms / sig,
               reference_noise=REF_NOISE_RMS[ft])
    print(f"[scan vs reference, type {ft}] {rep}")
    # TPDF dither of +-1 LSB sits on top: a difference below one LSB flips a sample now and then
    assert rms <= 3.0 * REF_NOISE_RMS[ft] * sig + 1.0
    assert d.max() <= 30.0 * REF_NOISE_RMS[ft] * sig + 2.0


def test_scan_streaming_equals_one_call(engine, oracle):
    spec = S.config_c2(hilbert_mode="scan")
    fb = S.frame_bytes(spec)
    n = 40001
    raw = rand_bytes(spec, n, 43)
    whole = engine.session(spec, 1).process_host(raw)[0]
    ses = engine.session(spec, 1)
    parts = [ses.process_host(raw[a * fb:b * fb])[0] for a, b in ((0, 777), (777, 33000), (33000, n))]
    rep = pcm_report(np.concatenate(parts), whole, 3)
    print(f"[scan streaming] {rep}")
    assert rep["max_lsb"] <= 1 and rep["mismatches"] <= 2    # carries round differently at the 1e-16 level


def test_scan_many_streams(engine, oracle):
    spec = S.config_c1(hilbert_mode="scan")
    K, n = 9, 6000
    raws = np.stack([rand_bytes(spec, n, 500 + k) for k in range(K)])
    ses = engine.session(spec, K)
    bus, _ = ses.enable_taps(n)
    ses.process_host(raws)
    ana = bus.cpu().numpy()[:, :, 0, :]
    for k in range(K):
        un = np.zeros((n, 4))
        oracle.port().icwo_unpack(4, 2, np.ascontiguousarray(raws[k]).ctypes.data_as(C.POINTER(C.c_uint8)), n,
                                  un.ctypes.data_as(C.POINTER(C.c_double)))
        ti, tq = truth_iq(oracle, np.ascontiguousarray(un[:, 2]), 1, 1)
        scale = np.sqrt(np.mean(ti ** 2 + tq ** 2))
        assert max(np.max(np.abs(ana[k, :, 2] - ti)), np.max(np.abs(ana[k, :, 3] - tq))) / scale <= TOL


@pytest.mark.parametrize("ft", [0, 1, 4])
@pytest.mark.parametrize("first", ["exact", "scan"])
def test_mode_switch_on_live_stream(engine, oracle, ft, first):
    """hilbert_mode changes on a live stream, no reset: the filter memory crosses over to the other basis
    (icw_hbconv.cpp, binary128 on the host; reference: filter settings change on a live stream, src/in_cwave.c:135-191).
    The continuation must be the same filter: analytic signal after the switch within 3 x the reference's noise of
    that design from the truth (a delay line holds 1 ulp of |w| ~ 1e10 x output, no conversion can do better) -- and
    nowhere near the O(1) error of a dropped or mis-mapped state."""
    from noise_floor import REF_NOISE_RMS
    second = "scan" if first == "exact" else "exact"
    spec = S.config_c1(hilbert_mode=first, filter_no=ft)
    fb = S.frame_bytes(spec)
    n, cut = 60000, 33001
    raw = rand_bytes(spec, n, 5 + ft)
    ses = engine.session(spec, 1)
    ses.process_host(raw[: cut * fb])
    ses.set_spec(dict(spec, hilbert_mode=second))
    bus, _ = ses.enable_taps(n - cut)
    ses.process_host(raw[cut * fb:])
    ana = bus.cpu().numpy()[0, :, 0, :]
    st = ses.get_state(0)
    assert st.hb_basis == (1 if second == "scan" else 0) and st.n_frame == n
    un = np.zeros((n, 4))
    oracle.port().icwo_unpack(oracle.FMT[spec["fmt"]], 2, raw.ctypes.data_as(C.POINTER(C.c_uint8)), n,
                              un.ctypes.data_as(C.POINTER(C.c_double)))
    for ch in range(2):
        ti, tq = truth_iq(oracle, np.ascontiguousarray(un[:, 2 * ch]), ft, 1)
        scale = np.sqrt(np.mean(ti ** 2 + tq ** 2))
        e = np.sqrt(np.mean((ana[:, 2 * ch] - ti[cut:]) ** 2 + (ana[:, 2 * ch + 1] - tq[cut:]) ** 2)) / scale
        e0 = max(abs(ana[0, 2 * ch] - ti[cut]), abs(ana[0, 2 * ch + 1] - tq[cut])) / scale
        print(f"[switch {first}->{second}, type {ft}] ch {ch}: rms err after the switch {e:.2e}, first frame {e0:.2e} "
              f"(reference noise {REF_NOISE_RMS[ft]:.1e})")
        assert e <= 3.0 * REF_NOISE_RMS[ft] + 1e-12
        assert e0 <= 30.0 * REF_NOISE_RMS[ft] + 1e-12


@pytest.mark.parametrize("ft,warm,n", [(0, 40_000, 120_000), (5, 1 << 19, 2 * (1 << 19) + 30_000)])
def test_time_sharding_stitches_on_one_gpu(engine, oracle, ft, warm, n):
    """The multi-GPU time split, emulated on one GPU: shard 1 starts from closed-form scalars and a
    filter state obtained by a warm-up over the tail of shard 0 (what rank 0 would send over NCCL).  Type 0 forgets
    fastest (0.99832^40000 ~ 1e-29); type 5 is the slowest design (pole radius 0.99986) and gets the hand-off's own
    warm-up length, dist.WARMUP_FRAMES = 2^19 (0.99986^(2^19) ~ 1e-32)."""
    from in_cwave_b200 import dist as D
    assert warm <= D.WARMUP_FRAMES
    spec = S.config_c2(hilbert_mode="scan", filter_no=ft)
    fb = S.frame_bytes(spec)
    raw = rand_bytes(spec, n, 61)
    whole = engine.session(spec, 1).process_host(raw)[0]
    a1, _ = D.shard_time(n, 1, 2)
    be0, be1 = D.CudaBackend(engine, spec), D.CudaBackend(engine, spec)
    tail = raw[(a1 - warm) * fb: a1 * fb]
    state = be0.hilbert_state_after(tail, D.closed_form_state(spec, a1 - warm).quad)
    be0.start_at(D.closed_form_state(spec, 0), np.zeros(D.STATE_DOUBLES))
    p0 = be0.process(raw[: a1 * fb])
    be1.start_at(D.closed_form_state(spec, a1), state)
    p1 = be1.process(raw[a1 * fb:])
    rep = pcm_report(np.concatenate([p0, p1]), whole, 3)
    print(f"[time shards] {rep}")
    assert rep["max_lsb"] <= 1 and rep["mismatches"] <= 2


@pytest.mark.parametrize("L", [1024, 4096])
@pytest.mark.parametrize("ft", [0, 1, 5])
def test_scan_long_chunks_late_starting_modes(engine, oracle, monkeypatch, ft, L):
    """Long streams use chunks of 1024 / 4096 frames and pass 1 starts the fast modes late (a mode of
    pole radius r has forgotten an input after log(1e-18)/log(r) samples).  Forced here on a short
    input (ICW_SCAN_L): same bar against the binary128 truth, split calls included."""
    import torch
    monkeypatch.setenv("ICW_SCAN_L", str(L))
    rng = np.random.default_rng(900 + ft)
    n = 3 * 128 * L // 2 + 12345                     # two tiles and a ragged tail
    n = min(n, 800_000)
    x = (rng.random((2, n)) - 0.5) * 30000.0
    x[1] += 12000.0 * np.sin(2 * np.pi * 0.2499 * np.arange(n))
    xa = torch.from_numpy(x).cuda()
    out, st = engine.hilbert(xa, ft, 1, 0, "scan")
    out = out.cpu().numpy()
    cut = 5 * L + 3
    o1, s1 = engine.hilbert(xa[:, :cut].contiguous(), ft, 1, 0, "scan")
    o2, _ = engine.hilbert(xa[:, cut:].contiguous(), ft, 1, 0, "scan", states=s1)
    both = np.concatenate([o1.cpu().numpy(), o2.cpu().numpy()], axis=1)
    for c in range(2):
        ti, tq = truth_iq(oracle, x[c], ft, 1)
        scale = np.sqrt(np.mean(ti ** 2 + tq ** 2))
        for name, got in (("one call", out), ("two calls", both)):
            err = max(np.max(np.abs(got[c, :, 0] - ti)), np.max(np.abs(got[c, :, 1] - tq))) / scale
            print(f"[scan L={L}] type {ft} ch {c} {name}: max err / rms = {err:.2e}")
            assert err <= TOL


def test_scan_long_chunks_whole_chain(engine, oracle, monkeypatch):
    """The C2 chain with 4096-frame chunks gives the bytes of the 256-frame chunks up to the scan's
    own rounding (carries differ at the 1e-16 level)."""
    spec = S.config_c2(hilbert_mode="scan")
    n = 700_001
    raw = rand_bytes(spec, n, 71)
    base = engine.session(spec, 1).process_host(raw)[0]
    monkeypatch.setenv("ICW_SCAN_L", "4096")
    long_ = engine.session(spec, 1).process_host(raw)[0]
    rep = pcm_report(long_, base, 3)
    print(f"[scan L=4096 vs 256] {rep}")
    assert rep["max_lsb"] <= 1 and rep["mismatches"] <= 4


def test_scan_long_chunks_many_streams(engine, oracle, monkeypatch):
    """Several streams in one launch group with 1024-frame chunks: same bytes as with 256-frame chunks up to
    the scan's own rounding, stream by stream."""
    spec = S.config_c1(hilbert_mode="scan")
    K, n = 5, 300_001
    raws = np.stack([rand_bytes(spec, n, 700 + k) for k in range(K)])
    base = engine.session(spec, K).process_host(raws)
    monkeypatch.setenv("ICW_SCAN_L", "1024")
    long_ = engine.session(spec, K).process_host(raws)
    for k in range(K):
        rep = pcm_report(long_[k], base[k], 3)
        assert rep["max_lsb"] <= 1 and rep["mismatches"] <= 4, (k, rep)
