#!/usr/bin/env python
"""tools/fuzz_one.py SEED: stream 0 of case SEED of tests/test_fuzz.py (make_case) in one call with taps: the differing samples, the
master output and the bus against the restatement."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import numpy as np
import in_cwave_b200 as icw
from in_cwave_b200 import spec as S, synth
from oracle import pyoracle as po
from util import pcm_to_int
from test_fuzz import make_case

seed = int(sys.argv[1])
_, spec, K, n, raws, cuts = make_case(seed, k_choices=(1, 1, 2, 5, 33))       # tools/fuzz_gpu.py's stream counts
raw = raws[0]
print({k: v for k, v in spec.items() if k != "nodes"}, "streams", K, "frames", n, "cuts", cuts)
for nd in spec["nodes"]:
    print("   ", nd)
eng = icw.Engine(0)
ses = eng.session(spec, 1)
bus, lrt = ses.enable_taps(n)
pcm = ses.process_host(raw)[0]
plugs = sorted({0} | {nd["out"] for nd in spec["nodes"] if nd["mode"] != "master"})
ref = po.port_process(spec, raw, want_lr=True, taps=plugs)
bps = 3 if spec["need24bits"] else 2
g, w = pcm_to_int(pcm, bps), pcm_to_int(ref["pcm"], bps)
bad = np.nonzero(g != w)[0]
lr = lrt.cpu().numpy()[0]
print("mismatches", bad.size)
for i in bad[:12]:
    f, c = divmod(int(i), 2)
    print(f"frame {f} ch {c}: gpu {g[i]} ref {w[i]}  lr gpu {lr[f, c]!r} ref {ref['lr'][f, c]!r}  diff {lr[f, c] - ref['lr'][f, c]:.3e}")
b = bus.cpu().numpy()[0]
for j, p in enumerate(plugs):
    d = np.abs(b[:, p, :] - ref["bus"][:, j, :])
    print("plug", p, "max abs diff", d.max(), "rel", d.max() / max(1e-300, np.abs(ref["bus"][:, j, :]).max()))
