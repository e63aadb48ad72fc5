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
