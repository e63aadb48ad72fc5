"""MT19937 jump-ahead mathematics on the host (GF(2) polynomials; no GPU).

Reference generator: src/mersene_twister/mt_jrnd.c:28-134.  The oracle's sequential generator is
the checker; the product's jump path must land on the identical 624-word state."""
import ctypes as C

import numpy as np
import pytest

from in_cwave_b200 import _abi


def test_characteristic_polynomial():
    n, d = C.c_int(), C.c_int()
    assert _abi.lib().icw_mt_host_charpoly(C.byref(n), C.byref(d)) == 0
    assert d.value == 19937
    assert n.value == 135          # the known weight of MT19937's characteristic polynomial


@pytest.mark.parametrize("seed", [0x13579BDF, 0x479B22AB, 5489])
@pytest.mark.parametrize("blocks", [0, 1, 2, 3, 33, 34, 1000, 4097])
def test_jump_lands_on_the_sequential_state(oracle, seed, blocks):
    L = _abi.lib()
    seq, jmp, fam, prod = (C.c_uint32 * 624)(), (C.c_uint32 * 624)(), (C.c_uint32 * 624)(), (C.c_uint32 * 624)()
    L.icw_mt_host_seq_state(seed, blocks, seq)
    assert L.icw_mt_host_jump_state(seed, blocks, jmp) == 0
    assert L.icw_mt_host_jump_state_family(seed, blocks, fam) == 0
    assert L.icw_mt_host_jump_state_product(seed, blocks, prod) == 0
    assert list(seq) == list(jmp) == list(fam) == list(prod)
    # and the sequential state is the oracle's: draw 624*blocks words, then compare the next ones
    mt = oracle.Mt()
    P = oracle.port()
    P.icwo_mt_seed(C.byref(mt), seed)
    for _ in range(624 * blocks):
        P.icwo_mt_u32(C.byref(mt))
    nxt = np.array([P.icwo_mt_u32(C.byref(mt)) for _ in range(8)], dtype=np.uint32)
    st = np.array(list(seq), dtype=np.uint32)
    # regenerate once in numpy from the jumped state and temper
    u = st.copy()
    for k in range(8):
        a, b = u[k], u[k + 1]
        mix = (a & np.uint32(0x80000000)) | (b & np.uint32(0x7FFFFFFF))
        tw = (mix >> np.uint32(1)) ^ (np.uint32(0x9908B0DF) if b & np.uint32(1) else np.uint32(0))
        u[k] = u[k + 397] ^ tw
    y = u[:8].copy()
    y ^= y >> np.uint32(11)
    y ^= (y << np.uint32(7)) & np.uint32(0x9D2C5680)
    y ^= (y << np.uint32(15)) & np.uint32(0xEFC60000)
    y ^= y >> np.uint32(18)
    assert np.array_equal(y, nxt)


def test_far_jump_consistency():
    """2^33 blocks ahead (far beyond what can be generated sequentially here): the direct power and
    the squared-family composition must agree with each other."""
    L = _abi.lib()
    a, b, c = (C.c_uint32 * 624)(), (C.c_uint32 * 624)(), (C.c_uint32 * 624)()
    blocks = (1 << 33) + 12345
    assert L.icw_mt_host_jump_state(0x13579BDF, blocks, a) == 0
    assert L.icw_mt_host_jump_state_family(0x13579BDF, blocks, b) == 0
    assert L.icw_mt_host_jump_state_product(0x13579BDF, blocks, c) == 0
    assert list(a) == list(b) == list(c)


def test_unit_length_covers_the_range_with_one_wave():
    """Jump-ahead units are m * 2^k blocks with m < 16: never more units than CTAs, and the wave at least
    8/9 full once there is more than one block per CTA."""
    L = _abi.lib()
    for max_units in (1, 8, 592, 1184):
        for nb in (1, 2, 15, 16, 17, 591, 592, 593, 2565, 107_000, 4_430_770, 10 ** 9 + 7):
            u = int(L.icw_mt_host_unit_blocks(nb, max_units))
            m = u
            while m % 2 == 0 and m >= 16:
                m //= 2
            assert m < 16 and u >= 1
            units = -(-nb // u)
            assert units <= max_units
            if nb >= 16 * max_units:
                assert units * 9 >= max_units * 8 - 9, (nb, max_units, u, units)
    assert int(L.icw_mt_host_unit_blocks(4_430_770, 592)) == 7680        # the C2 call: 577 units of 15 * 512 blocks


def test_scan_chunk_length_policy(monkeypatch):
    L = _abi.lib()
    monkeypatch.delenv("ICW_SCAN_L", raising=False)
    assert L.icw_host_scan_chunk_len(1, 50_000, 148) == 256
    assert L.icw_host_scan_chunk_len(1, 138_240_000, 148) == 1024
    assert L.icw_host_scan_chunk_len(1, 691_200_000, 148) == 2048
    assert L.icw_host_scan_chunk_len(4096, 480_000, 148) == 2048          # many streams count like one long one
    monkeypatch.setenv("ICW_SCAN_L", "4096")
    assert L.icw_host_scan_chunk_len(1, 1000, 148) == 4096
    monkeypatch.setenv("ICW_SCAN_L", "300")                                # not a multiple of 256: ignored
    assert L.icw_host_scan_chunk_len(1, 1000, 148) == 256
