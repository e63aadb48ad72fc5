"""Differential tests on RANDOM configurations: the same seeded bytes through the compiled reference
(oracle/_ref: /root/reference/src/adv_modulator.c:587-763 and everything under it, driven through its own reader),
the CPU restatement (oracle/icw_oracle.c) and the CUDA path (C ABI).  The specs cover every knob of the path
(format, channels, rate, filter design, summation, reject, scaled frequencies, bit depth, significant bits, quantiser,
dither type and depth, noise shaper) and DSP lists of any shape -- up to five nodes on random plugs with random
exchange modes, I/Q inversions, gains, switched-off channels, several inputs, bypass, one-frame feedback loops.

Bars: reference against restatement -- equal bytes.  CUDA path in exact mode -- equal bytes unless the list holds a
sin/cos (libdevice against glibc, <= 2 ulp: one LSB on at most 1 sample in 10^5, counted); clip counters, reject
counters and generator positions equal."""
import os

import numpy as np
import pytest

from in_cwave_b200 import spec as S, synth
from oracle import pyoracle as po
from util import add_exceptional_samples, add_random_fades, pcm_report, random_spec

N_CPU, N_GPU = 48, 48


def _bps(spec):
    return 3 if spec["need24bits"] else 2


def _input(spec, rng, n):
    lvl = float(rng.choice([0.25, 0.9, 1.6]))                 # the last one clips
    raw = synth.stream_bytes(spec, n, stream_id=int(rng.integers(1, 1 << 30)), level=lvl)
    return add_exceptional_samples(rng, spec, np.frombuffer(bytes(raw), dtype=np.uint8), p=float(os.environ.get("ICW_FUZZ_EXC", "0.15")))


@pytest.mark.parametrize("seed", range(N_CPU))
def test_restatement_equals_the_compiled_reference_on_random_specs(seed):
    if not po.have_ref():
        pytest.skip("oracle/_ref is not built")
    rng = np.random.default_rng(1000 + seed)
    spec = random_spec(rng)
    n = int(rng.integers(2, 6000))                 # MIN_FILE_SAMPLES (src/in_cwave.h:315): the reader refuses a shorter file
    spec = add_random_fades(rng, spec, n)
    raw = _input(spec, rng, n)
    r = po.ref_process(spec, raw, read_quant=int(rng.choice([4096, 1111, 1, 64])) if n < 2000 else 4096)
    p = po.port_process(spec, raw)
    assert np.array_equal(r["pcm"], p["pcm"]), (seed, spec, pcm_report(p["pcm"], r["pcm"], _bps(spec)))


def _has_trig(spec):
    return not spec["bypass"] and any(nd["mode"] in ("shift", "pm") and (nd.get("l_on", 1) or nd.get("r_on", 1)) for nd in spec["nodes"])


def _explain_flips(eng, spec, raw, pcm, want, label):
    """More flipped LSBs than one in 10^5: a list can make its output a difference of nearly equal numbers (a plug mixed with its
    own I/Q-swapped copy, master I-Q: seed 3140), whose sign -- the quantiser's step at zero -- then hangs on the last bit of a
    sin/cos.  Then, and only then: the same bytes again with the master output tapped; that output within 1e-12 of the bus
    scale of the restatement's; bytes differ ONLY where the tapped doubles differ."""
    from util import pcm_to_int
    ses = eng.session(spec, 1)
    try:
        n = raw.size // S.frame_bytes(spec)
        _, lr = ses.enable_taps(n)
        again = ses.process_host(raw)[0]
        assert np.array_equal(again, pcm), label
        plugs = sorted({0} | {nd["out"] for nd in spec["nodes"] if nd["mode"] != "master"})
        ref = po.port_process(spec, raw, taps=plugs, want_lr=True)
        got = lr.cpu().numpy()[0]
        scale = max(float(np.max(np.abs(ref["bus"]))), 1e-300)
        assert float(np.max(np.abs(got - ref["lr"]))) <= 1e-12 * scale, label
        bps = _bps(spec)
        flipped = (pcm_to_int(pcm, bps) != pcm_to_int(want, bps)).reshape(-1, 2)
        same_in = got == ref["lr"]
        assert not np.any(flipped & same_in), label
    finally:
        ses.close()


def make_case(seed, n_max=30000, k_choices=(1, 1, 2, 5)):
    """The random configuration, stream count, inputs and call cuts of case `seed` (tools/fuzz_one.py replays one of them)."""
    rng = np.random.default_rng(5000 + seed)
    spec = random_spec(rng)
    K = int(rng.choice(k_choices))
    n = int(rng.integers(2, n_max))
    spec = add_random_fades(rng, spec, n)
    raws = [np.frombuffer(bytes(_input(spec, rng, n)), dtype=np.uint8) for _ in range(K)]
    cuts = sorted({0, n} | {int(c) for c in rng.integers(0, n + 1, size=int(rng.integers(0, 4)))})
    return rng, spec, K, n, raws, cuts


def run_cuda_case(seed, eng, n_max=30000, k_choices=(1, 1, 2, 5)):
    """One random configuration through the C ABI (K streams, the call cut at random places) against the reference."""
    rng, spec, K, n, raws, cuts = make_case(seed, n_max, k_choices)
    fb = S.frame_bytes(spec)
    raw = np.stack(raws)
    ses = eng.session(spec, K)
    try:
        parts = []
        for a, b in zip(cuts[:-1], cuts[1:]):
            if b == a:
                continue
            if parts and not spec["is_fp_check"] and rng.random() < 0.3:
                # checkpoint / resume: everything a stream carries is in icw_stream_state (include/icw_b200.h) -- a NEW session
                # given those states goes on as if nothing had happened (the FP_CHECK counters alone live outside it)
                states = [ses.get_state(k) for k in range(K)]
                ses.close()
                ses = eng.session(spec, K)
                for k in range(K):
                    ses.set_state(k, states[k])
            parts.append(ses.process_host(np.ascontiguousarray(raw[:, a * fb:b * fb])))
        pcm = np.concatenate(parts, axis=1)
        stats = ses.stats()
        clips = [0, 0]
        for k in range(K):
            port = po.port_process(spec, raws[k])
            ref = po.ref_process(spec, raws[k]) if (po.have_ref() and k == 0) else port
            rep = pcm_report(pcm[k], ref["pcm"], _bps(spec))
            if _has_trig(spec):
                # a quantiser with few significant bits turns one flipped rounding into 2^(shift) output steps
                step = 1 << ((24 - spec["sign_bits24"]) if spec["need24bits"] else (16 - spec["sign_bits16"]))
                assert rep["max_lsb"] <= step, (seed, k, rep, spec)
                if rep["mismatches"] > max(1, rep["samples"] // 100000):
                    _explain_flips(eng, spec, raws[k], pcm[k], ref["pcm"], (seed, k, rep))
            else:
                assert rep["mismatches"] == 0, (seed, k, rep, spec)
            p = port["state"]
            st = ses.get_state(k)
            assert st.n_frame == p.n_frame
            assert (st.mt_drawn[0], st.mt_drawn[1]) == (p.mt[0].drawn, p.mt[1].drawn), (seed, k)
            if spec["is_fp_check"] and not _has_trig(spec):
                # FP_EXCEPT_STATS of the four checked stages (src/fp_check.c:48-99): [stage][total, snan, qnan, -inf, -den, +den, +inf]
                assert ses.fp_stats(k) == [list(r) for r in p.fp_cnt], (seed, k, spec)
            clips[0] += p.clips[0]; clips[1] += p.clips[1]
        if not _has_trig(spec):
            assert tuple(stats["clips"]) == tuple(clips), (seed, stats["clips"], clips)
    finally:
        ses.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(N_GPU))
def test_cuda_path_equals_the_reference_on_random_specs(engine, seed):
    run_cuda_case(seed, engine)


REAL_FORMATS = ["wav_u8", "wav_i16", "wav_i24", "wav_i32", "wav_f32"]


def run_scan_case(seed, eng, n_max=40000):
    """Scan mode on a random configuration, one to three streams, the call cut at random places: the analytic signal against the
    binary128 evaluation of the reference's filter (1e-12 of its RMS, tests/test_gpu_scan.py), and everything behind the
    converter against the restatement FED that analytic signal (the same bar as exact mode)."""
    import ctypes as C
    rng = np.random.default_rng(9000 + seed)
    spec = random_spec(rng, hilbert_mode="scan")
    spec["fmt"] = str(rng.choice(REAL_FORMATS))
    spec["is_fp_check"] = 0
    spec["is_subnorm_reject"] = 0                   # scan mode has no state zeroing (counted, not modelled)
    n = int(rng.integers(64, n_max))
    K = int(rng.choice([1, 1, 3]))
    fb = S.frame_bytes(spec)
    raws = [np.frombuffer(bytes(synth.stream_bytes(spec, n, stream_id=int(rng.integers(1, 1 << 30)), level=0.25)), dtype=np.uint8) for _ in range(K)]
    raw = np.stack(raws)
    cuts = sorted({0, n} | {int(c) for c in rng.integers(0, n + 1, size=int(rng.integers(0, 3)))})
    ses = eng.session(spec, K)
    try:
        bus, lr = ses.enable_taps(n)
        parts, anas = [], []
        for a, b in zip(cuts[:-1], cuts[1:]):
            if b == a:
                continue
            parts.append(ses.process_host(np.ascontiguousarray(raw[:, a * fb:b * fb])))
            # the tap buffer is packed per call: [stream][frames of this call][plug][4]
            anas.append(bus.reshape(-1)[: K * (b - a) * bus.shape[2] * 4].reshape(K, b - a, bus.shape[2], 4)[:, :, 0, :].cpu().numpy().copy())
        pcm, ana = np.concatenate(parts, axis=1), np.ascontiguousarray(np.concatenate(anas, axis=1))
        nch = spec["n_channels"]
        for k in range(K):
            un = np.zeros((n, 4))
            po.port().icwo_unpack(po.FMT[spec["fmt"]], nch, raws[k].ctypes.data_as(C.POINTER(C.c_uint8)), n, un.ctypes.data_as(C.POINTER(C.c_double)))
            for ch in range(2):
                ti, tq = po.hilbert_truth(np.ascontiguousarray(un[:, 2 * ch]), spec["filter_no"], 1 if spec["is_kahan"] else 0, 0)
                scale = max(float(np.sqrt(np.mean(ti ** 2 + tq ** 2))), 1e-300)
                err = max(np.max(np.abs(ana[k, :, 2 * ch] - ti)), np.max(np.abs(ana[k, :, 2 * ch + 1] - tq))) / scale
                assert err <= 1e-12, (seed, k, ch, err, spec)
            ref = po.port_process(dict(spec, fmt="cw_f64", n_channels=2), np.ascontiguousarray(ana[k]).view(np.uint8).ravel())
            rep = pcm_report(pcm[k], ref["pcm"], _bps(spec))
            if _has_trig(spec):
                step = 1 << ((24 - spec["sign_bits24"]) if spec["need24bits"] else (16 - spec["sign_bits16"]))
                assert rep["max_lsb"] <= step and rep["mismatches"] <= max(1, rep["samples"] // 100000), (seed, k, rep, spec)
            else:
                assert rep["mismatches"] == 0, (seed, k, rep, spec)
            st = ses.get_state(k)
            assert st.n_frame == n and st.hb_basis == 1
    finally:
        ses.close()


@pytest.mark.gpu
@pytest.mark.parametrize("seed", range(24))
def test_scan_mode_on_random_specs(engine, seed):
    run_scan_case(seed, engine)
