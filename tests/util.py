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
