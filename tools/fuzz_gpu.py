#!/usr/bin/env python
"""tools/fuzz_gpu.py FIRST COUNT [N_MAX] [--scan]: the randomised differential test of tests/test_fuzz.py over a long range of seeds on
the GPU box (one engine, random specs / stream counts / call cuts, compiled reference as the checker); prints every failing seed."""
import sys
import traceback
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path[:0] = [str(ROOT), str(ROOT / "tests")]

import in_cwave_b200 as icw          # noqa: E402
from test_fuzz import run_cuda_case, run_scan_case   # noqa: E402

first, count = int(sys.argv[1]), int(sys.argv[2])
scan = "--scan" in sys.argv
if scan:
    sys.argv.remove("--scan")
n_max = int(sys.argv[3]) if len(sys.argv) > 3 else 30000
eng = icw.Engine(0)
bad = 0
for seed in range(first, first + count):
    try:
        if scan:
            run_scan_case(seed, eng)
        else:
            run_cuda_case(seed, eng, n_max=n_max, k_choices=(1, 1, 2, 5, 33))
    except Exception as e:              # noqa: BLE001
        bad += 1
        print("FAIL seed", seed, type(e).__name__, str(e)[:1500])
        if not isinstance(e, AssertionError):
            traceback.print_exc()
print(f"seeds {first}..{first + count - 1}: {bad} failing")
eng.close()
