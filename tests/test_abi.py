"""The C-ABI library loads and exports every symbol include/icw_b200.h declares (no GPU needed)."""
import ctypes as C
import re
from pathlib import Path

from in_cwave_b200 import _abi

ROOT = Path(__file__).resolve().parent.parent


def _declared():
    text = (ROOT / "include" / "icw_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(icw_[a-z0-9_]+)\s*\(", text)))


def test_header_and_binding_list_agree():
    assert _declared() == sorted(_abi.EXPORTS)


def test_library_exports_everything():
    L = _abi.lib()
    for name in _declared():
        assert hasattr(L, name), f"{name} is declared in include/icw_b200.h but not exported"
    assert L.icw_abi_version() == 1


def test_struct_sizes_match_header():
    """ctypes mirrors vs the C compiler's view of the same header."""
    import subprocess, tempfile, os
    src = '#include <stdio.h>\n#include "icw_b200.h"\nint main(void){printf("%zu %zu %zu %zu\\n",sizeof(icw_node),sizeof(icw_chain_spec),sizeof(icw_stream_state),sizeof(icw_stats));return 0;}\n'
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "s.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "s")
        subprocess.run(["gcc", "-std=c99", "-I", str(ROOT / "include"), c, "-o", exe], check=True)
        sizes = [int(v) for v in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert sizes == [C.sizeof(_abi.Node), C.sizeof(_abi.ChainSpecC), C.sizeof(_abi.StreamState), C.sizeof(_abi.Stats)]


def test_defaults_are_the_reference_defaults():
    L = _abi.lib()
    sp = _abi.ChainSpecC()
    L.icw_default_spec(sp)
    assert (sp.filter_no, sp.is_kahan, sp.is_subnorm_reject, sp.is_frmod_scaled) == (1, 1, 1, 1)
    assert (sp.need24bits, sp.quantz_type, sp.render_type, sp.nshape_type) == (1, 1, 0, 0)
    assert (sp.sign_bits16, sp.sign_bits24, sp.dth_bits) == (16, 24, 1.0)
    assert sp.n_nodes == 1 and sp.nodes[0].mode == 0 and sp.nodes[0].l_gain == 0.8 and sp.nodes[0].inputs_mask == 1
    st = _abi.StreamState()
    L.icw_default_state(st)
    assert (st.mt_seed[0], st.mt_seed[1]) == (0x13579BDF, 0x479B22AB)
    assert L.icw_frame_bytes(sp) == 8 and L.icw_out_frame_bytes(sp) == 6
    assert L.icw_peak_db(0.0) == -555.0 and abs(L.icw_peak_db(0.5) + 6.020599913279624) < 1e-15


def test_missing_gpu_fails_loudly():
    """No silent CPU path: without a usable B200 the engine refuses to exist."""
    import pytest
    import torch
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    h = C.c_void_p()
    rc = _abi.lib().icw_engine_create(0, C.byref(h))
    assert rc != 0 and not h.value
    assert _abi.lib().icw_last_error()


def test_enum_values_match_the_python_mirror():
    """The constants the ctypes layer carries are the header's (compiled, not parsed: a C program prints them)."""
    import os
    import subprocess
    import tempfile
    names = ["ICW_RESET_HILBERT", "ICW_RESET_FRAMECNT", "ICW_RESET_COUNTERS", "ICW_RESET_FILEPOS", "ICW_RESET_RENDER",
             "ICW_RESET_RENDER_MEMORY", "ICW_RESET_ALL", "ICW_HILBERT_EXACT", "ICW_HILBERT_SCAN", "ICW_MODE_MASTER", "ICW_MODE_SHIFT",
             "ICW_MODE_PM", "ICW_MODE_MIX", "ICW_N_PLUGS", "ICW_MAX_NODES"]
    src = '#include <stdio.h>\n#include "icw_b200.h"\nint main(void){' + "".join(f'printf("%d\\n",(int){n});' for n in names) + "return 0;}\n"
    with tempfile.TemporaryDirectory() as td:
        c = os.path.join(td, "e.c")
        open(c, "w").write(src)
        exe = os.path.join(td, "e")
        subprocess.run(["gcc", "-std=c99", "-I", str(ROOT / "include"), c, "-o", exe], check=True)
        vals = dict(zip(names, (int(v) for v in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split())))
    assert (vals["ICW_RESET_HILBERT"], vals["ICW_RESET_FRAMECNT"], vals["ICW_RESET_COUNTERS"], vals["ICW_RESET_FILEPOS"], vals["ICW_RESET_RENDER"],
            vals["ICW_RESET_RENDER_MEMORY"], vals["ICW_RESET_ALL"]) == (_abi.RESET_HILBERT, _abi.RESET_FRAMECNT, _abi.RESET_COUNTERS,
                                                                        _abi.RESET_FILEPOS, _abi.RESET_RENDER, _abi.RESET_RENDER_MEMORY, _abi.RESET_ALL)
    assert (vals["ICW_HILBERT_EXACT"], vals["ICW_HILBERT_SCAN"]) == (_abi.HILBERT["exact"], _abi.HILBERT["scan"])
    assert [vals[f"ICW_MODE_{m.upper()}"] for m in ("master", "shift", "pm", "mix")] == [_abi.MODE[m] for m in ("master", "shift", "pm", "mix")]
    assert vals["ICW_N_PLUGS"] == _abi.N_PLUGS
    assert vals["ICW_RESET_ALL"] & vals["ICW_RESET_RENDER_MEMORY"]          # a full reset includes what a file open clears
