"""Filter state between the two Hilbert bases (in_cwave_b200/csrc/icw_hbconv.cpp; host only, no GPU).

The delay line of the reference's DF-II recurrence (src/hblpf.c:894-926) and the modal states of scan mode describe
the same filter driven by the same input: s_k[n] = sum_j p_k^(n-j) u[j] on one side, w[n] = u[n] + sum_i fb_i w[n-1-i] on
the other.  The check evaluates both independently with mpmath (50 digits) from the reference's rounded coefficients
and holds the library's conversion to them."""
import ctypes as C
import re
import struct
from pathlib import Path

import numpy as np
import pytest

from in_cwave_b200 import _abi

mp = pytest.importorskip("mpmath")
ROOT = Path(__file__).resolve().parent.parent


def _tables():
    txt = (ROOT / "in_cwave_b200/csrc/icw_hb_tables.inc").read_text()
    orders = [int(v) for v in re.search(r"ICW_HB_ORDER\[ICW_HB_NTYPES\] = \{([^}]*)\}", txt).group(1).split(",")]

    def tab(name):
        m = re.search(r"ICW_HB_%s\[ICW_HB_NTYPES\]\[ICW_HB_MAXORD \+ 1\] = \{(.*?)\n\};" % name, txt, re.S)
        rows = re.findall(r"\{(.*?)\}", m.group(1), re.S)
        return [[struct.unpack("<d", struct.pack("<Q", int(w, 16)))[0] for w in re.findall(r"0x([0-9A-F]{16})ULL", r)] for r in rows]
    return orders, tab("A")


def _convert(ft, to_basis, v):
    a = (C.c_double * _abi.MAX_ORD)(*[float(x) for x in v])
    o = (C.c_double * _abi.MAX_ORD)()
    assert _abi.lib().icw_host_hb_convert(ft, to_basis, a, o) == 0
    return np.array(list(o))


@pytest.mark.parametrize("ft", [0, 1, 2, 5])
def test_delay_line_to_modal_against_an_independent_evaluation(ft):
    mp.mp.dps = 60
    orders, A = _tables()
    n = orders[ft]
    a = [mp.mpf(x) for x in A[ft][: n + 1]]
    poles = mp.polyroots(a, maxsteps=5000, extraprec=2000)
    upper = sorted([p for p in poles if mp.im(p) > mp.mpf("1e-40")], key=lambda p: float(mp.arg(p)))
    reals = [p for p in poles if abs(mp.im(p)) <= mp.mpf("1e-40")]
    modes = upper + reals                                   # the order of icw_hb_modal.inc
    rng = np.random.default_rng(ft)
    u = [mp.mpf(float(x)) for x in rng.standard_normal(400) * 1000.0]
    # DF-II state in exact arithmetic: w[k] = u[k] - sum_{i>=1} a_i w[k-i]   (a_0 == 1)
    w = []
    for k in range(len(u)):
        acc = u[k]
        for i in range(1, n + 1):
            if k - i >= 0:
                acc -= a[i] * w[k - i]
        w.append(acc)
    z = [w[len(u) - 1 - j] for j in range(n)] + [mp.mpf(0)] * (_abi.MAX_ORD - n)      # newest first
    # modal states in exact arithmetic
    want = np.zeros(_abi.MAX_ORD)
    for m, p in enumerate(modes):
        sacc = mp.mpc(0)
        for k in range(len(u)):
            sacc = sacc * p + u[k]
        want[2 * m] = float(mp.re(sacc))
        if m < len(upper):
            want[2 * m + 1] = float(mp.im(sacc))
    # the delay line as doubles carries 1 ulp of |w| ~ 1e10 x output: feed the conversion the exactly rounded state
    got = _convert(ft, 1, [float(x) for x in z])
    scale = np.max(np.abs(want))
    wmax = max(abs(float(x)) for x in z)
    # what 1/2 ulp of the delay line can do to a modal state: sum |q_i| * ulp(w) -- bounded here by 1e-6 of the state scale
    err = np.max(np.abs(got - want)) / scale
    print(f"type {ft}: |w| max {wmax:.3g}, modal scale {scale:.3g}, conversion error / scale {err:.2e}")
    assert err < 2e-5
    # and the way back lands on the same delay line to a few ulps of its largest entry
    back = _convert(ft, 0, got)
    zf = np.array([float(x) for x in z])
    assert np.max(np.abs(back - zf)) / np.max(np.abs(zf)) < 1e-13


def test_round_trip_and_bad_arguments():
    rng = np.random.default_rng(7)
    for ft in range(6):
        v = np.zeros(_abi.MAX_ORD)
        n = [15, 19, 18, 19, 20, 20][ft]
        v[:n] = rng.standard_normal(n) * 1e12
        m = _convert(ft, 1, v)
        b = _convert(ft, 0, m)
        # an arbitrary vector is not a filter state: its modal image is as large as itself and rounding THAT to double
        # comes back through a map of condition ~1e5..1e10 -- the bound is the conditioning, not the arithmetic
        assert np.max(np.abs(b - v)) / np.max(np.abs(v)) < 1e-4
        assert np.all(m[n:] == 0.0) and np.all(b[n:] == 0.0)
    z = (C.c_double * _abi.MAX_ORD)()
    assert _abi.lib().icw_host_hb_convert(6, 1, z, z) != 0
    assert _abi.lib().icw_host_hb_convert(1, 2, z, z) != 0
