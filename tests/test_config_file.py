"""the reference's configuration file drives the B200 path.

The oracle here is the reference's OWN load_config()/save_config() (src/config.c:815-975, compiled into
oracle/_ref): files written by the reference must parse to the same chain in icwp_load_config, files
written by icwp_save_config must be accepted by the reference and give the same chain there, and the
reference's rejection rules (bad line, unknown keyword, wrong version, list without a leading master,
lock fix-ups, clamping) must hold on hand-made files.  No GPU is touched.
"""
import sys
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parents[1]
sys.path.insert(0, str(ROOT))

from oracle import pyoracle as po          # noqa: E402
from in_cwave_b200 import plugin, spec as S  # noqa: E402

pytestmark = pytest.mark.skipif(not po.have_ref(), reason="oracle/_ref not built")


def _cfg_tuple_ref(c):
    return (c.filter_no, c.is_kahan, c.is_subnorm_reject, c.is_frmod_scaled, c.need24bits, c.dth_bits,
            c.quantz_type, c.render_type, c.nshape_type, c.sign_bits16, c.sign_bits24,
            c.sec_align, c.fade_in, c.fade_out, c.clr_nframe_trk, c.clr_hilb_trk, c.is_fp_check)


def _cfg_tuple_ours(sp, o):
    return (sp.filter_no, sp.is_kahan, sp.is_subnorm_reject, sp.is_frmod_scaled, sp.need24bits, sp.dth_bits,
            sp.quantz_type, sp.render_type, sp.nshape_type, sp.sign_bits16, sp.sign_bits24,
            o.sec_align, o.fade_in_ms, o.fade_out_ms, o.clr_nframe_trk, o.clr_hilb_trk, sp.is_fp_check)


CASES = {
    "c1": S.config_c1(),
    "c2": S.config_c2(dth_bits=1.5, sign_bits24=20),
    "c3": S.config_c3(),
    "plain": S.default_spec(filter_no=4, is_kahan=0, is_subnorm_reject=0, is_frmod_scaled=0, need24bits=0,
                            quantz_type=0, render_type=4, sign_bits16=12),
    "fp_check": S.config_c1(is_fp_check=1, render_type=2),
}


@pytest.mark.parametrize("name", sorted(CASES))
def test_files_written_by_the_reference_parse_to_the_same_chain(tmp_path, name):
    d = CASES[name]
    f = tmp_path / "ref.cfg"
    assert po.ref_save_config(d, f)
    ok_r, cfg_r, nodes_r = po.ref_load_config(f)
    ok_o, sp, o, nodes_o = plugin.load_config(f)
    assert ok_r and ok_o
    assert _cfg_tuple_ours(sp, o) == _cfg_tuple_ref(cfg_r)
    assert nodes_o == nodes_r                       # bit-equal doubles: the file stores bit patterns
    assert len(nodes_o) == len(d["nodes"])


@pytest.mark.parametrize("name", sorted(CASES))
def test_files_written_by_us_are_accepted_by_the_reference(tmp_path, name):
    d = CASES[name]
    f = tmp_path / "ours.cfg"
    assert plugin.save_config(f, d, sec_align=3, fade_in_ms=250, fade_out_ms=1000, clr_hilb_trk=1)
    ok_r, cfg_r, nodes_r = po.ref_load_config(f)
    ok_o, sp, o, nodes_o = plugin.load_config(f)
    assert ok_r and ok_o
    assert _cfg_tuple_ours(sp, o) == _cfg_tuple_ref(cfg_r)
    assert (cfg_r.sec_align, cfg_r.fade_in, cfg_r.fade_out, cfg_r.clr_hilb_trk) == (3, 250, 1000, 1)
    assert nodes_o == nodes_r


def _base_lines():
    return ["VER_CONFIG=10", "IIR_HBLPF_IX=3", "RENDER_TYPE=2", "NEED24BITS=1"]


def _inputs(*on):
    return " ".join("1" if k in on else "0" for k in range(27))


MASTER = "NODE_DSP=Master 0.8 0.8 0 " + _inputs(1) + " 0 0 0 0 0 0"


def _both(tmp_path, lines):
    f = tmp_path / "hand.cfg"
    f.write_text("\n".join(lines) + "\n")
    ok_r, cfg_r, nodes_r = po.ref_load_config(f)
    ok_o, sp, o, nodes_o = plugin.load_config(f)
    assert ok_o == ok_r
    assert _cfg_tuple_ours(sp, o) == _cfg_tuple_ref(cfg_r)
    assert nodes_o == nodes_r
    return ok_o, sp, o, nodes_o


def test_decimal_doubles_escapes_case_and_blank_lines(tmp_path):
    lines = _base_lines() + ["", "   ", "dither_bits=2.5", "  Fade_In=40",
                             "NODE_DSP=My%%% name% here%x 0.5 0x3FF8000000000000 0 " + _inputs(2) + " 0 0 0 0 2 3",
                             "NODE_DSP=Shift-1 1 1 0 " + _inputs(0) + " 1 0 1 1 12.25 1 -3.5 1 2 0 0"]
    ok, sp, o, nodes = _both(tmp_path, lines)
    assert ok and sp.dth_bits == 2.5 and o.fade_in_ms == 40
    assert len(nodes) == 2 and nodes[1][5:7] == (0.5, 1.5) and nodes[1][8:10] == (2, 3)
    assert nodes[0][12][0] == 12.25 and nodes[0][13][0] == -3.5


def test_values_are_clamped_like_the_reference(tmp_path):
    lines = _base_lines() + ["SEC_ALIGN=99", "FADE_OUT=123456", "DITHER_BITS=77", "SIGNBITS16=1", "SIGNBITS24=99",
                             "IIR_HBLPF_IX=42", "QUANTIZE_TYPE=9", "NOISE_SHAPING=99",
                             "NODE_DSP=Master 5 -1 0 " + _inputs(1) + " 9 0 0 0 7 -2",
                             "NODE_DSP=PM-1 1 1 0 " + _inputs(0) + " 0 0 0 2 99 -3 5 3 1 41 0.25 0.5 -0.5 1 40 0 0 0 0",
                             "NODE_DSP=Shift-2 1 1 0 " + _inputs(0) + " 0 0 0 1 -25 1 21 1 0 0 0"]
    ok, sp, o, nodes = _both(tmp_path, lines)
    assert ok and (o.sec_align, o.fade_out_ms, sp.dth_bits, sp.sign_bits16, sp.sign_bits24) == (20, 10000, 23.0, 2, 24)
    assert nodes[-1][5:7] == (2.0, 0.0) and nodes[-1][2] == 4 and nodes[-1][8:10] == (3, 0)
    assert nodes[0][12][0] == -20.0 and nodes[0][13][0] == 20.0 and nodes[0][7] == 1
    assert nodes[1][12] == (40.0, -1.0, 1.0, 1.0) and nodes[1][7] == 27


def test_lock_flags_copy_left_to_right(tmp_path):
    lines = _base_lines() + ["NODE_DSP=Master 0.7 0.1 1 " + _inputs(1, 2) + " 0 1 0 0 0 0",
                             "NODE_DSP=Shift-1 1 1 0 " + _inputs(0) + " 0 0 0 1 5.5 1 1.0 0 1 1 1",
                             "NODE_DSP=PM-2 1 1 0 " + _inputs(0) + " 0 0 0 2 4 0.1 0.2 0.3 1 9 0.9 0.8 0.7 0 2 1 1 0 1"]
    ok, sp, o, nodes = _both(tmp_path, lines)
    assert ok
    master, shift, pm = nodes[2], nodes[1], nodes[0]
    assert master[5:7] == (0.7, 0.7) and master[3:5] == (1, 1)
    assert shift[13][0] == -5.5 and shift[11] == 1
    assert pm[13] == (4.0, 0.1, 0.8, 0.3) and pm[11] == 1


@pytest.mark.parametrize("bad", [
    ["VER_CONFIG=9"],                                   # wrong version
    ["VER_CONFIG=10", "NO_SUCH_KEY=1", "RENDER_TYPE=2"],  # unknown keyword
    ["VER_CONFIG=10", "just some text", "RENDER_TYPE=2"],  # no '='
    ["VER_CONFIG=10", "RENDER_TYPE=2\x01"],             # control character
    ["RENDER_TYPE=2"],                                  # no version line at all
    ["VER_CONFIG=ten", "RENDER_TYPE=2"],                # version does not parse -> stays 0
])
def test_rejected_files_leave_the_defaults(tmp_path, bad):
    ok, sp, o, nodes = _both(tmp_path, bad)
    assert not ok
    assert sp.render_type == 0 and sp.filter_no == 1 and len(nodes) == 1 and nodes[0][0] == 0


def test_values_that_do_not_parse_are_skipped_not_fatal(tmp_path):
    """The reference overwrites a handler's verdict with the next line's read (src/config.c:846-903)."""
    lines = _base_lines() + ["RENDER_TYPE=two", "SIGNBITS16=x", "SIGNBITS16=12", "DITHER_BITS=", MASTER,
                             "NODE_DSP=Shift-1 1 1 0 1 0 0",                       # truncated node: dropped
                             "NODE_DSP=Mix-2 1 1 0 " + _inputs(0) + " 0 0 0 3 5",
                             "FADE_IN=oops"]                                        # even as the last line
    ok, sp, o, nodes = _both(tmp_path, lines)
    assert ok and sp.render_type == 2 and sp.sign_bits16 == 12 and sp.dth_bits == 1.0 and o.fade_in_ms == 0
    assert [n[0] for n in nodes] == [3, 0]


def test_list_without_a_leading_master_is_dropped_but_the_file_is_accepted(tmp_path):
    lines = _base_lines() + ["NODE_DSP=Shift-1 1 1 0 " + _inputs(0) + " 0 0 0 1 5 1 5 1 1 0 0", MASTER]
    ok, sp, o, nodes = _both(tmp_path, lines)
    assert ok and sp.filter_no == 3 and len(nodes) == 1 and nodes[0][0] == 0
    two = _base_lines() + [MASTER, MASTER]
    ok, sp, o, nodes = _both(tmp_path, two)
    assert ok and len(nodes) == 1


def test_missing_file(tmp_path):
    ok, sp, o, nodes = plugin.load_config(tmp_path / "nope.cfg")
    assert not ok and sp.n_nodes == 1 and sp.filter_no == 1


# ---- randomised: files of random specs, and random damage to them, through both loaders -----------------------------------
def _mutate(rng, lines):
    """One random edit of a configuration file: the kind of damage a hand-edited or truncated file carries."""
    import numpy as np
    lines = list(lines)
    k = int(rng.integers(0, len(lines)))
    kind = int(rng.integers(0, 10))
    toks = lines[k].split(" ")
    if kind == 0:
        del lines[k]
    elif kind == 1:
        lines.insert(k, lines[int(rng.integers(0, len(lines)))])                    # a line twice / out of place
    elif kind == 2 and len(toks) > 1:
        j = int(rng.integers(1, len(toks)))
        toks[j] = str(rng.choice(["-1", "0", "1", "2", "7", "99", "1e9", "-0.0", "0.5", "nan", "inf", "x", "", "0x3FF0000000000000",
                                  "0x7FF8000000000000", "4294967296", "-2147483649", "1.5e-320"]))
        lines[k] = " ".join(toks)
    elif kind == 3 and len(toks) > 2:
        lines[k] = " ".join(toks[: int(rng.integers(1, len(toks)))])                # truncated line
    elif kind == 4:
        lines[k] = lines[k] + " " + " ".join(str(int(v)) for v in rng.integers(0, 3, size=int(rng.integers(1, 5))))   # extra fields
    elif kind == 5 and "=" in lines[k]:
        key, val = lines[k].split("=", 1)
        lines[k] = str(rng.choice([key.lower(), key.title(), " " + key, key + " ", key[:-1], key + "X"])) + "=" + val
    elif kind == 6:
        lines[k] = lines[k].replace("=", str(rng.choice(["", "==", " = ", ":"])), 1)
    elif kind == 7:
        lines.insert(k, str(rng.choice(["", "   ", "\t", "# comment", "; comment", "[section]", "NODE_DSP=", "NODE_DSP=Master",
                                        "VER_CONFIG=10", "VER_CONFIG=11"])))
    elif kind == 8 and len(toks) > 3:
        a, b = (int(v) for v in rng.integers(1, len(toks), size=2))
        toks[a], toks[b] = toks[b], toks[a]
        lines[k] = " ".join(toks)
    else:
        lines[k] = lines[k].replace("Master", str(rng.choice(["master", "Mast%er", "M%%", "Shift-1", "%"])), 1)
    return lines


@pytest.mark.parametrize("seed", range(40))
def test_random_specs_and_random_damage_through_both_loaders(tmp_path, seed):
    """A random chain written by the reference's save_config parses to the same chain in both loaders -- and so does the
    same file after one to three random edits (lines dropped, doubled, truncated, fields replaced by junk, keys re-cased,
    separators broken): same accept / reject verdict, same values, same DSP list, bit for bit (src/config.c:562-975)."""
    import numpy as np
    sys.path.insert(0, str(ROOT / "tests"))
    from util import random_spec
    rng = np.random.default_rng(4000 + seed)
    d = random_spec(rng)
    f = tmp_path / "ref.cfg"
    assert po.ref_save_config(d, f)
    ok_r, cfg_r, nodes_r = po.ref_load_config(f)
    ok_o, sp, o, nodes_o = plugin.load_config(f)
    assert ok_r and ok_o and _cfg_tuple_ours(sp, o) == _cfg_tuple_ref(cfg_r) and nodes_o == nodes_r
    base = f.read_text().split("\n")
    while base and base[-1] == "":
        base.pop()
    for trial in range(25):
        lines = base
        for _ in range(int(rng.integers(1, 4))):
            lines = _mutate(rng, lines)
        g = tmp_path / "damaged.cfg"
        g.write_text("\n".join(lines) + "\n")
        ok_r, cfg_r, nodes_r = po.ref_load_config(g)
        ok_o, sp, o, nodes_o = plugin.load_config(g)
        ctx = (seed, trial, "\n".join(lines))
        assert ok_o == ok_r, ctx
        assert _cfg_tuple_ours(sp, o) == _cfg_tuple_ref(cfg_r), ctx
        assert _same_nodes(nodes_o, nodes_r), ctx


def _same_nodes(a, b):
    """Node tuples equal, NaN equal to NaN (a damaged file can carry one)."""
    import math

    def eq(x, y):
        if isinstance(x, tuple):
            return isinstance(y, tuple) and len(x) == len(y) and all(eq(p, q) for p, q in zip(x, y))
        if isinstance(x, float) and isinstance(y, float) and math.isnan(x) and math.isnan(y):
            return True
        return x == y
    return len(a) == len(b) and all(eq(p, q) for p, q in zip(a, b))
