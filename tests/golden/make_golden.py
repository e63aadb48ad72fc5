#!/usr/bin/env python3
"""Generate tests/golden/*.npz from the COMPILED REFERENCE (oracle/_ref/libicw_ref.so).

Run in the build container only (needs /root/reference to have been compiled by
`make -C oracle ref`).  Each fixture stores the seeded input bytes' recipe (spec + seed + n), the
reference's rendered PCM, the bus taps of the plugs the graph uses, and the counters.  The inputs
themselves are re-created from the recipe by in_cwave_b200.synth (deterministic numpy), so the
files stay small.  Also extracts the reference's one known-answer vector (MT19937, 1000 words,
src/mersene_twister/test_mt_jrnd/mt19937ar_out.c:21) into mt19937_kat.npz.
"""
import json
import re
import sys
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent.parent
sys.path.insert(0, str(ROOT))

from in_cwave_b200 import spec as S  # noqa: E402
from in_cwave_b200 import synth  # noqa: E402
from oracle import pyoracle as po  # noqa: E402

HERE = Path(__file__).resolve().parent


def cases():
    n = 3000
    yield "c1_f32_shift_24bit", S.config_c1(), n, [0, 1]
    yield "c1_nokahan", S.config_c1(is_kahan=0), n, [0, 1]
    for ft in (0, 2, 3, 4, 5):
        yield f"c1_filter{ft}", S.config_c1(filter_no=ft), 1500, [0, 1]
    yield "c2_i24_tpdf", S.config_c2(), n, [0, 1]
    for rt, name in ((1, "rpdf"), (3, "stpdf"), (4, "gauss")):
        yield f"c2_i24_{name}", S.config_c2(render_type=rt), 1500, [1]
    yield "c3_cwave_graph_16bit", S.config_c3(), n, [0, 1, 2, 3, 26]
    yield "c3_cwave_tpdf_tread", S.config_c3(render_type=2, quantz_type=0), 1500, [1, 3]
    yield "mono_i16_fade", S.default_spec(fmt="wav_i16", n_channels=1, sample_rate=44100, n_samples=2000,
                                          n_fade_in=441, n_fade_out=882), 2000, [0]
    yield "u8_stereo_bits12", S.default_spec(fmt="wav_u8", sample_rate=22050, sign_bits24=12), 1500, [0]
    yield "i32_16bit_bits10", S.default_spec(fmt="wav_i32", need24bits=0, sign_bits16=10), 1500, [0]
    yield "cw_f64_mono_sub", S.default_spec(fmt="cw_f64", n_channels=1, sample_rate=96000,
                                            nodes=[dict(mode="master", inputs=[0], l_gain=1.0, r_gain=0.5, l_tout=1, r_tout=3)]), 1500, [0]
    yield "cw_i16_bypass", S.config_c3(fmt="cw_i16", bypass=1), 1500, [0]
    yield "cw_i16f32_unscaled", S.config_c3(fmt="cw_i16f32", is_frmod_scaled=0), 1500, [1, 3]
    yield "loud_clipping", S.config_c1(), 1500, [1]


def main():
    if not po.have_ref():
        sys.exit("oracle/_ref/libicw_ref.so missing: run `make -C oracle ref` in the build container")
    index = {}
    for name, spec, n, taps in cases():
        level = 2.2 if name == "loud_clipping" else 0.25
        raw = synth.stream_bytes(spec, n, stream_id=len(index) + 1, level=level)
        out = po.ref_process(spec, raw, taps=taps)
        assert out["frames"] == n, (name, out["frames"])
        st = out["stats"]
        np.savez_compressed(
            HERE / f"{name}.npz", pcm=out["pcm"], bus=out["bus"], taps=np.array(taps),
            clips=np.array([st.l_clips, st.r_clips]), peak=np.array([st.l_peak, st.r_peak]),
            rejects=np.array([st.subnorm_cnt]), n_frame=np.array([st.n_frame]))
        index[name] = dict(spec=spec, n=n, seed=len(index) + 1, level=level)
        print(f"{name}: {n} frames, pcm {out['pcm'].size} B, clips {st.l_clips}/{st.r_clips}")
    (HERE / "index.json").write_text(json.dumps(index, indent=1))

    src = Path("/root/reference/src/mersene_twister/test_mt_jrnd/mt19937ar_out.c").read_text()
    body = src[src.index("test_u32[1000]"):]
    body = body[body.index("{") + 1: body.index("}")]
    words = np.array([int(w) for w in re.findall(r"(\d+)U", body)], dtype=np.uint32)
    assert words.size == 1000
    key = np.array([0x123, 0x234, 0x345, 0x456], dtype=np.uint32)
    np.savez_compressed(HERE / "mt19937_kat.npz", key=key, words=words)
    print("mt19937_kat: 1000 words")


if __name__ == "__main__":
    main()
