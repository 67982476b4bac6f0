#!/usr/bin/env python
"""TSD-bounded variant generation on a C2-sized batch (10 000 flanked candidates) for timing / profiling:
python tools/tsd_probe.py [n_candidates] [iterations]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import hite_b200

n = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000
rng = np.random.default_rng(11)
bases = np.frombuffer(b"ACGT", dtype=np.uint8)
seqs = []
for _ in range(n):
    L = int(rng.integers(300, 3100))
    a = bases[rng.integers(0, 4, L)]
    k = int(rng.choice([2, 3, 4, 5, 6, 8, 9, 10, 11]))
    a[50 - k:50] = a[L - 50:L - 50 + k]
    seqs.append(a.tobytes())
starts = [51] * n
ends = [len(s) - 50 for s in seqs]
ctx = hite_b200.Context(0)
for it in range(int(sys.argv[2]) if len(sys.argv) > 2 else 5):
    t0 = time.perf_counter()
    recs, cnt = ctx.tsd_candidates(seqs, starts, ends, 50, 1)
    wall = time.perf_counter() - t0
    tm = ctx.timing()
    print("tsd n=%d records=%d  h2d %.3f ms  kernel %.3f ms  d2h %.3f ms  wall %.1f ms  (%.2f M candidates/s on the device)" % (
        n, int(cnt.sum()), tm["h2d_ms"], tm["kernel_ms"], tm["d2h_ms"], 1e3 * wall, n / tm["kernel_ms"] / 1e3), flush=True)
