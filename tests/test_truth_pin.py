"""Pins the yardstick of scan mode: oracle.hilbert_truth -- the converter's recurrences in IEEE
binary128 -- is checked (a) against an independent exact evaluation with mpmath and (b) through the distance it
measures between the reference's FP64 sequence and the truth, design by design (Type 0 ~ 1.4e-9, default Type 1 ~ 5e-5).
CPU only."""
import ctypes as C
import re
import struct
from pathlib import Path

import numpy as np
import pytest

from noise_floor import REF_NOISE_RMS

ROOT = Path(__file__).resolve().parent.parent


def _coef(ft):
    txt = (ROOT / "in_cwave_b200/csrc/icw_hb_tables.inc").read_text()
    orders = [int(v) for v in re.search(r"ICW_HB_ORDER\[ICW_HB_NTYPES\] = \{([^}]*)\}", txt).group(1).split(",")]

    def tab(name):
        m = re.search(r"ICW_HB_%s\[ICW_HB_NTYPES\]\[ICW_HB_MAXORD \+ 1\] = \{(.*?)\n\};" % name, txt, re.S)
        rows = re.findall(r"\{(.*?)\}", m.group(1), re.S)
        return [[struct.unpack("<d", struct.pack("<Q", int(w, 16)))[0] for w in re.findall(r"0x([0-9A-F]{16})ULL", r)] for r in rows]
    n = orders[ft]
    return n, tab("A")[ft][: n + 1], tab("B")[ft][: n + 1]


@pytest.mark.parametrize("ft,drop_direct", [(0, 1), (1, 1), (1, 0), (4, 1)])
def test_truth_evaluator_against_mpmath(oracle, ft, drop_direct):
    """hq_rp_process (src/lpf_hilbert_quad.c:129-156) around the DF-II recurrences (src/hblpf.c:894-926), evaluated with
    60-digit arithmetic from the rounded coefficients; drop_direct = the Kahan path's missing d0*x term (hblpf.c:1056)."""
    mp = pytest.importorskip("mpmath")
    mp.mp.dps = 60
    n, a, b = _coef(ft)
    a = [mp.mpf(v) for v in a]
    b = [mp.mpf(v) for v in b]
    rng = np.random.default_rng(ft)
    N = 600
    x = (rng.random(N) - 0.5) * 20000.0
    ti, tq = oracle.hilbert_truth(x, ft, drop_direct, 0)
    wI, wQ = [], []
    err = 0.0
    scale = float(np.sqrt(np.mean(ti ** 2 + tq ** 2)))
    for k in range(N):
        q = k & 3
        xi = [mp.mpf(x[k]), 0, -mp.mpf(x[k]), 0][q]
        xq = [0, -mp.mpf(x[k]), 0, mp.mpf(x[k])][q]
        outs = []
        for w, u in ((wI, xi), (wQ, xq)):
            acc = mp.mpf(u)
            for i in range(1, n + 1):
                if k - i >= 0:
                    acc -= a[i] * w[k - i]
            y = b[0] * (acc - mp.mpf(u)) if drop_direct else b[0] * acc     # Kahan: d0 * (w - x), i.e. everything but d0 * x
            for i in range(1, n + 1):
                if k - i >= 0:
                    y += b[i] * w[k - i]
            w.append(acc)
            outs.append(y)
        yI, yQ = outs
        re_, im_ = [(2 * yI, 2 * yQ), (-2 * yQ, 2 * yI), (-2 * yI, -2 * yQ), (2 * yQ, -2 * yI)][q]
        err = max(err, abs(float(re_ - mp.mpf(ti[k]))), abs(float(im_ - mp.mpf(tq[k]))))
    # the truth is returned as doubles: half an ulp of the output is all that may separate it from exact arithmetic
    print(f"type {ft}: truth vs mpmath, max abs err / rms = {err / scale:.2e}")
    assert err / scale < 4e-16 * 8


@pytest.mark.parametrize("ft", range(6))
def test_reference_noise_floor_per_design(oracle, ft):
    """The number the scan-mode bounds are built on.  Type 0's is the one DESIGN.md quotes as ~1.4e-9."""
    dp = lambda arr: arr.ctypes.data_as(C.POINTER(C.c_double))
    n = 60000
    x = (np.random.default_rng(9).random(n) - 0.5) * 20000.0
    for kahan in (1, 0):
        ti, tq = oracle.hilbert_truth(x, ft, 1 if kahan else 0, 0)
        scale = np.sqrt(np.mean(ti ** 2 + tq ** 2))
        pi, pq = np.zeros(n), np.zeros(n)
        lpf = (oracle.Iir * 2)()
        quad = C.c_uint(0)
        oracle.port().icwo_hilbert(ft, kahan, 0, lpf, C.byref(quad), dp(x), n, dp(pi), dp(pq))
        e = float(np.sqrt(np.mean((pi - ti) ** 2 + (pq - tq) ** 2)) / scale)
        print(f"type {ft} kahan {kahan}: reference vs truth rms = {e:.3e} (table {REF_NOISE_RMS[ft]:.1e})")
        assert 0.5 * REF_NOISE_RMS[ft] < e < 2.0 * REF_NOISE_RMS[ft]
    if ft == 0:
        assert 1.0e-9 < e < 1.8e-9
