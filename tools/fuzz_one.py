#!/usr/bin/env python
"""tools/fuzz_one.py SEED: one case of tests/test_fuzz.py::run_cuda_case with the differing samples printed."""
import sys
from pathlib import Path
ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]
import numpy as np
import in_cwave_b200 as icw
from in_cwave_b200 import spec as S, synth
from oracle import pyoracle as po
from util import random_spec, pcm_to_int

seed = int(sys.argv[1])
rng = np.random.default_rng(5000 + seed)
spec = random_spec(rng)
K = int(rng.choice((1, 1, 2, 5, 33))); n = int(rng.integers(1, 30000))
lvl = float(rng.choice([0.25, 0.9, 1.6]))
raw = np.frombuffer(bytes(synth.stream_bytes(spec, n, stream_id=int(rng.integers(1, 1 << 30)), level=lvl)), dtype=np.uint8)
eng = icw.Engine(0)
ses = eng.session(spec, 1)
bus, lrt = ses.enable_taps(n)
pcm = ses.process_host(raw)[0]
ref = po.port_process(spec, raw, want_lr=True, taps=[0, 4, 13, 22])
bps = 3 if spec["need24bits"] else 2
g, w = pcm_to_int(pcm, bps), pcm_to_int(ref["pcm"], bps)
bad = np.nonzero(g != w)[0]
lr = lrt.cpu().numpy()[0]
print("mismatches", bad.size)
for i in bad[:12]:
    f, c = divmod(int(i), 2)
    print(f"frame {f} ch {c}: gpu {g[i]} ref {w[i]}  lr gpu {lr[f, c]!r} ref {ref['lr'][f, c]!r}  diff {lr[f, c] - ref['lr'][f, c]:.3e}")
b = bus.cpu().numpy()[0]
for j, p in enumerate([0, 4, 13, 22]):
    d = np.abs(b[:, p, :] - ref["bus"][:, j, :])
    print("plug", p, "max abs diff", d.max(), "rel", d.max() / max(1e-300, np.abs(ref["bus"][:, j, :]).max()))
