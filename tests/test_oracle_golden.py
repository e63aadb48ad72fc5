"""The port oracle (oracle/icw_oracle.c) against the committed golden vectors.

tests/golden/*.npz were produced by the reference's own C sources compiled in place
(tests/golden/make_golden.py); mt19937_kat.npz is the reference's one known-answer vector
(src/mersene_twister/test_mt_jrnd/mt19937ar_out.c).  Everything here must be bit-exact.
"""
import ctypes as C
import json
from pathlib import Path

import numpy as np
import pytest

from in_cwave_b200 import synth

GOLD = Path(__file__).resolve().parent / "golden"
INDEX = json.loads((GOLD / "index.json").read_text())


@pytest.mark.parametrize("name", sorted(INDEX))
def test_port_matches_golden(oracle, name):
    rec = INDEX[name]
    g = np.load(GOLD / f"{name}.npz")
    raw = synth.stream_bytes(rec["spec"], rec["n"], stream_id=rec["seed"], level=rec["level"])
    taps = [int(t) for t in g["taps"]]
    out = oracle.port_process(rec["spec"], raw, taps=taps)
    assert np.array_equal(out["pcm"], g["pcm"]), "rendered PCM differs from the reference"
    assert np.array_equal(out["bus"], g["bus"]), "bus taps differ from the reference"
    st = out["state"]
    assert (st.clips[0], st.clips[1]) == tuple(int(v) for v in g["clips"])
    assert (st.peak_db[0], st.peak_db[1]) == tuple(float(v) for v in g["peak"])
    rej = sum(int(st.lpf[c][f].rejects) for c in range(2) for f in range(2))
    assert rej == int(g["rejects"][0])
    assert int(st.n_frame) == int(g["n_frame"][0])


def test_mt_known_answer(oracle):
    """init_key({0x123,0x234,0x345,0x456}) then 1000 x gen_ui32 (test_mt_jrnd/main.c:30-44)."""
    g = np.load(GOLD / "mt19937_kat.npz")
    L = oracle.port()
    mt = oracle.Mt()
    key = (C.c_uint32 * 4)(*[int(k) for k in g["key"]])
    L.icwo_mt_seed_key(C.byref(mt), key, 4)
    got = np.array([L.icwo_mt_u32(C.byref(mt)) for _ in range(1000)], dtype=np.uint32)
    assert np.array_equal(got, g["words"])


def test_mt_derived_generators(oracle):
    """dsemi / dsopen re-derived from the same golden words (test_mt_jrnd/main.c:180-253)."""
    g = np.load(GOLD / "mt19937_kat.npz")
    w = g["words"].astype(np.uint64)
    L = oracle.port()
    mt = oracle.Mt()
    key = (C.c_uint32 * 4)(*[int(k) for k in g["key"]])
    L.icwo_mt_seed_key(C.byref(mt), key, 4)
    got = np.array([L.icwo_mt_dsopen(C.byref(mt)) for _ in range(500)])
    a, b = (w[0::2] >> np.uint64(5)).astype(np.float64), (w[1::2] >> np.uint64(6)).astype(np.float64)
    want = (a * 67108864.0 + b) * (1.0 / 9007199254740992.0) * 2.0 - 1.0
    assert np.array_equal(got, want)


def test_split_calls_equal_one_call(oracle):
    """State carried across calls: two half blocks == one block (the MOD_CONTEXT persists)."""
    from in_cwave_b200 import spec as S
    sp = S.config_c2()
    raw = synth.stream_bytes(sp, 4000, stream_id=7)
    one = oracle.port_process(sp, raw)
    fb = S.frame_bytes(sp)
    st = oracle.new_state()
    a = oracle.port_process(sp, raw[: 1777 * fb], state=st)
    b = oracle.port_process(sp, raw[1777 * fb:], state=st)
    assert np.array_equal(np.concatenate([a["pcm"], b["pcm"]]), one["pcm"])
