"""Shared helpers of the parity tests: seeded inputs, the BASELINE graphs, comparison reports."""
from __future__ import annotations

import numpy as np

from in_cwave_b200 import spec as S
from in_cwave_b200 import synth


def rand_bytes(spec: dict, n: int, seed: int = 0, level: float = 0.25) -> np.ndarray:
    return synth.stream_bytes(spec, n, stream_id=seed, level=level)


def raw_random_bytes(spec: dict, n: int, seed: int = 0) -> np.ndarray:
    """Arbitrary bit patterns (integer formats only): exercises sign extension and clipping."""
    rng = np.random.default_rng(seed)
    return rng.integers(0, 256, size=n * S.frame_bytes(spec), dtype=np.uint8)


def pcm_to_int(pcm: np.ndarray, bytes_per_sample: int) -> np.ndarray:
    b = np.asarray(pcm, dtype=np.uint8).reshape(-1, bytes_per_sample).astype(np.int64)
    v = b[:, 0] | (b[:, 1] << 8)
    if bytes_per_sample == 3:
        v |= b[:, 2] << 16
        v = np.where(v >= 1 << 23, v - (1 << 24), v)
    else:
        v = np.where(v >= 1 << 15, v - (1 << 16), v)
    return v


def pcm_report(got: np.ndarray, want: np.ndarray, bytes_per_sample: int) -> dict:
    g, w = pcm_to_int(got, bytes_per_sample), pcm_to_int(want, bytes_per_sample)
    assert g.shape == w.shape, (g.shape, w.shape)
    diff = np.abs(g - w)
    return dict(samples=int(g.size), mismatches=int(np.count_nonzero(diff)), max_lsb=int(diff.max() if diff.size else 0))


def rel_err(got: np.ndarray, want: np.ndarray) -> float:
    want = np.asarray(want, dtype=np.float64)
    got = np.asarray(got, dtype=np.float64)
    scale = max(float(np.max(np.abs(want))), 1e-300)
    return float(np.max(np.abs(got - want))) / scale


def port_state_to_dict(st) -> dict:
    return dict(
        n_frame=int(st.n_frame), pos=int(st.pos), clips=(int(st.clips[0]), int(st.clips[1])),
        peak_db=(float(st.peak_db[0]), float(st.peak_db[1])),
        rejects=sum(int(st.lpf[c][f].rejects) for c in range(2) for f in range(2)),
        drawn=(int(st.mt[0].drawn), int(st.mt[1].drawn)),
    )


# ---------------------------------------------------------------------------------------------
# randomised specs for the differential tests (tests/test_fuzz.py): every knob of the path the reference
# exposes (src/config.c:118-207 bounds), DSP lists of any shape incl. one-frame feedback loops
# ---------------------------------------------------------------------------------------------
FUZZ_FORMATS = ["wav_u8", "wav_i16", "wav_i24", "wav_i32", "wav_f32", "cw_f64", "cw_i16", "cw_i16f32", "cw_f32"]


def random_nodes(rng, feedback: bool) -> list:
    n_mid = int(rng.integers(0, 6))
    outs = sorted(int(p) for p in rng.choice(np.arange(1, 27), size=max(n_mid, 1), replace=bool(rng.integers(0, 2))))[:n_mid]
    rng.shuffle(outs)
    nodes, written = [], [0]

    def pick_inputs(pool):
        k = int(rng.integers(1, min(3, len(pool)) + 1))
        return sorted(int(p) for p in rng.choice(pool, size=k, replace=False))

    # A feedback loop with gain above one is chaotic: the reference itself answers a one-ulp change of a gain with thousands of
    # different samples, so the <= 2 ulp between glibc's and libdevice's sin/cos cannot be held to an LSB there.  Feedback lists are
    # therefore drawn either without sin/cos (any gains, also unstable ones: both sides do the same IEEE operations -> equal bytes,
    # through Inf and NaN) or contractive (at most three inputs a node, every gain <= 0.3).
    trig_ok = not feedback or rng.random() < 0.6
    if feedback and trig_ok:
        gain = lambda: float(rng.choice([0.3, 0.25, 0.1, round(float(rng.uniform(0.02, 0.3)), 3)]))
    else:
        gain = lambda: float(rng.choice([1.0, 0.8, 0.5, round(float(rng.uniform(0.05, 1.2)), 3)]))
    for i, out in enumerate(outs):
        mode = str(rng.choice(["shift", "pm", "mix"])) if trig_ok else "mix"
        pool = sorted(set(written) | (set(outs) if feedback else set()))
        nd = dict(mode=mode, inputs=pick_inputs(pool), out=out, xch=int(rng.integers(0, 5)),
                  l_iq_invert=int(rng.integers(0, 2)), r_iq_invert=int(rng.integers(0, 2)), l_gain=gain(), r_gain=gain(),
                  l_on=int(rng.random() < 0.85), r_on=int(rng.random() < 0.85))
        if mode == "shift":
            nd["l_p"] = [round(float(rng.uniform(-60, 60)), 3)]
            nd["r_p"] = [round(float(rng.uniform(-60, 60)), 3)]
        elif mode == "pm":
            nd["l_p"] = [round(float(rng.uniform(0, 30)), 3), round(float(rng.uniform(-1, 1)), 3), round(float(rng.uniform(0, 2)), 3), round(float(rng.uniform(-1, 1)), 3)]
            nd["r_p"] = [round(float(rng.uniform(0, 30)), 3), round(float(rng.uniform(-1, 1)), 3), round(float(rng.uniform(0, 2)), 3), round(float(rng.uniform(-1, 1)), 3)]
        nodes.append(nd)
        written.append(out)
    pool = sorted(set(written))
    nodes.append(dict(mode="master", inputs=pick_inputs(pool), xch=int(rng.integers(0, 5)) if rng.random() < 0.3 else 0,
                      l_iq_invert=int(rng.random() < 0.2), r_iq_invert=int(rng.random() < 0.2),
                      l_gain=gain(), r_gain=gain(), l_tout=int(rng.integers(0, 4)), r_tout=int(rng.integers(0, 4))))
    return nodes


def random_spec(rng, hilbert_mode: str = "exact", allow_feedback: bool = True, allow_shaping: bool = True) -> dict:
    fmt = str(rng.choice(FUZZ_FORMATS))
    need24 = int(rng.integers(0, 2))
    feedback = bool(allow_feedback and rng.random() < 0.2)
    d = dict(
        fmt=fmt, n_channels=int(rng.integers(1, 3)), sample_rate=int(rng.choice([8000, 22050, 44100, 48000, 96000, 192000, 384000])),
        filter_no=int(rng.integers(0, 6)), is_kahan=int(rng.integers(0, 2)), is_subnorm_reject=int(rng.integers(0, 2)),
        hilbert_mode=hilbert_mode, is_frmod_scaled=int(rng.random() < 0.7), need24bits=need24,
        dth_bits=float(rng.choice([1.0, 0.5, 2.0, 3.25])), quantz_type=int(rng.integers(0, 2)), render_type=int(rng.integers(0, 5)),
        nshape_type=int(rng.integers(1, 18)) if (allow_shaping and rng.random() < 0.25) else 0,
        sign_bits16=int(rng.choice([16, 16, 12, 8, 2])), sign_bits24=int(rng.choice([24, 24, 20, 17, 9, 2])),
        bypass=int(rng.random() < 0.1), nodes=random_nodes(rng, feedback),
        is_fp_check=int(hilbert_mode == "exact" and rng.random() < 0.15),      # the FP-exception-checked twins (scan mode refuses them)
    )
    return S.default_spec(**d)


def add_random_fades(rng, spec: dict, n: int) -> dict:
    """Fade in / out (reference src/xwave_reader.c:707-723, 921-936) on a track of n frames: whole milliseconds (the reference's
    unit) at a rate that makes them whole frames, short enough that its one-third cap does not apply."""
    sr = spec["sample_rate"]
    if sr % 1000 or n < 64 or rng.random() < 0.6:
        return spec
    per_ms = sr // 1000
    most = (n - 1) // 2 // per_ms                       # n_fade_in + n_fade_out < n_samples
    if most < 1:
        return spec
    fi = int(rng.integers(0, min(most, 40) + 1)) * per_ms
    fo = int(rng.integers(0, min(most, 40) + 1)) * per_ms
    return dict(spec, n_samples=n, n_fade_in=fi, n_fade_out=fo)


def add_exceptional_samples(rng, spec: dict, raw: np.ndarray, p: float = 0.15) -> np.ndarray:
    """Floating-point formats, now and then: NaN, +-Inf, denormals and huge values scattered through the input (what the
    reference's FP_CHECK twins count, src/fp_check.c:48-99, and what must clip / propagate the same way without them)."""
    fmt = spec["fmt"]
    if fmt not in ("wav_f32", "cw_f32", "cw_f64") or rng.random() >= p:
        return raw
    f64 = fmt == "cw_f64"
    f = np.frombuffer(bytes(raw), dtype=np.float64 if f64 else np.float32).copy()
    k = int(min(f.size, rng.integers(1, 40)))
    vals = [np.nan, np.inf, -np.inf, 3e-312 if f64 else 1e-42, -2e-311 if f64 else -1e-43, 1e300 if f64 else 3e38, -0.0]
    for i in rng.choice(f.size, k, replace=False):
        f[i] = vals[int(rng.integers(0, len(vals)))]
    return f.view(np.uint8)
