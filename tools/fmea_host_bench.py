#!/usr/bin/env python
"""Host-side timing of the FMEA half of the coarse stage (hb_fmea_chain_files, host C++ inside the C-ABI library) without a
GPU: the HSPs of a C2-like genome come from the oracle's CPU self-join (test infrastructure, used here only as the input
generator and cached under /tmp), then the PRODUCT's chaining is timed on 1..N host threads and compared with the oracle's.

    python tools/fmea_host_bench.py [--mbp 30] [--threads 1,4,8] [--reps 3]
"""
import argparse
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ap = argparse.ArgumentParser()
ap.add_argument("--mbp", type=int, default=30)
ap.add_argument("--threads", default="1,4,8")
ap.add_argument("--reps", type=int, default=3)
ap.add_argument("--check", action="store_true", help="compare the intervals with the oracle's FMEA (slow: Python loop per segment)")
a = ap.parse_args()

cache = "/tmp/fm/hsps_%d.npz" % a.mbp
if not os.path.exists(cache):
    from workloads import synth
    from oracle import oracle as O
    n_contigs = max(1, a.mbp // 10)
    wl = synth.make_mixed_workload(a.mbp * 1_000_000, n_contigs, 100 * a.mbp, mix=(("tir", 1.0),), copies=("uniform", 6, 100),
                                   sub=(0.0, 0.10), indel=(0.0, 0.02), cands_per_family=20, seed=1)
    t = time.perf_counter()
    G = O.Genome(wl.contig_seqs)
    lens = [len(c) for c in wl.contig_seqs]
    import ctypes as C
    L = G._L
    segs, goff = [], []
    p = 64
    for c, ln in enumerate(lens):
        for o in range(0, ln, 1_000_000):
            segs.append((c, o)); goff.append(p + o)
        p += ln + 64
    goff = np.asarray(goff, dtype=np.uint64)
    L.hbo_self_hsps.restype = C.c_uint64
    L.hbo_self_hsps.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p,
                                C.c_int, C.c_int, C.c_void_p, C.c_uint64]
    slen = np.asarray([min(1_000_000, lens[c] - o) for c, o in segs], dtype=np.uint64)
    n = L.hbo_self_hsps(G.g, G.mm_pos_order.ctypes.data, G.index.ctypes.data, len(G.index), G.total, G.k, goff.ctypes.data,
                        slen.ctypes.data, len(segs), 64, None, 0)
    recs = np.zeros((max(1, n), 6), dtype=np.int32)
    L.hbo_self_hsps(G.g, G.mm_pos_order.ctypes.data, G.index.ctypes.data, len(G.index), G.total, G.k, goff.ctypes.data,
                    slen.ctypes.data, len(segs), 64, recs.ctypes.data, n)
    print("oracle self-join: %d HSPs in %.1f s" % (n, time.perf_counter() - t), file=sys.stderr)
    np.savez(cache, recs=recs[:n], segs=np.asarray(segs, dtype=np.int64), slen=slen)
z = np.load(cache)
recs, segs_a, slen = z["recs"], z["segs"], z["slen"]

from hite_b200 import _lib
order = np.argsort(recs[:, 0], kind="stable")
r = recs[order]
n_seg = len(segs_a)
q_off = np.zeros(n_seg + 1, dtype=np.uint64)
q_off[1:] = np.cumsum(np.bincount(r[:, 0], minlength=n_seg))
hsps = np.zeros(len(r), dtype=_lib.HSP_DTYPE)
for i, f in enumerate(("subj", "qs", "qe", "ss", "se")):
    hsps[f] = r[:, 1 + i]
segs = np.zeros(n_seg, dtype=_lib.SEG_DTYPE)
segs["chr"] = segs_a[:, 0]; segs["start"] = segs_a[:, 1]
fos, f, count = np.zeros(n_seg, dtype=np.uint32), 0, 0
for g in range(n_seg):
    fos[g] = f
    count += int(slen[g])
    if count >= 1_000_000:
        f += 1; count = 0
print("%d Mbp, %d segments, %d HSPs" % (a.mbp, n_seg, len(hsps)))
ref = None
for th in [int(x) for x in a.threads.split(",")]:
    best = 1e9
    for _ in range(a.reps):
        t = time.perf_counter()
        iv = _lib.fmea_chain_files(hsps, q_off, segs, fos, 2000, 30000, th)
        best = min(best, time.perf_counter() - t)
    key = iv.tobytes()
    assert ref is None or key == ref, "result depends on the thread count"
    ref = key
    print("threads %2d: %8.1f ms  (%d intervals, %.1f M HSP/s)" % (th, best * 1e3, len(iv), len(hsps) / best / 1e6))
import hashlib
print("intervals sha1", hashlib.sha1(ref).hexdigest())
