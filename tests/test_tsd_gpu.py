"""GPU parity: TSD-bounded variant generation through the C-ABI (hb_tsd_candidates) vs the C oracle
(oracle/hbo_tsd.c, itself pinned on the reference goldens) and vs the reference goldens directly.
Bit-exact records (k, start, end, dist, idx, short_class) in the oracle's canonical tie order."""
import collections
import gzip
import json
import os
import zlib

import numpy as np
import pytest

import tsd_synth
from oracle import oracle as O

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIELDS = ("k", "start", "end", "dist", "idx", "short_class")


@pytest.fixture(scope="module")
def golden_tsd():
    with gzip.open(os.path.join(ROOT, "tests", "golden", "tsd_v4.json.gz")) as f:
        return json.loads(f.read())


def _run(ctx, cases):
    """group by (d, plant) -- one launch per group -- and return records per case in input order"""
    out = [None] * len(cases)
    groups = collections.defaultdict(list)
    for i, c in enumerate(cases):
        groups[(c["d"], c["plant"])].append(i)
    for (d, plant), idx in groups.items():
        recs, cnt = ctx.tsd_candidates([cases[i]["seq"] for i in idx], [cases[i]["raw_start"] for i in idx],
                                       [cases[i]["raw_end"] for i in idx], d, plant)
        for row, i in enumerate(idx):
            out[i] = [tuple(int(r[f]) for f in FIELDS) for r in recs[row][:int(cnt[row])]]
    return out


def _oracle(c):
    return [tuple(r[f] for f in FIELDS) for r in
            O.search_confident_tir_v4(c["seq"], c["raw_start"], c["raw_end"], c["d"], c["plant"])]


def test_tsd_matches_oracle_on_golden_inputs(ctx, golden_tsd):
    got = _run(ctx, golden_tsd)
    n = 0
    for c, g in zip(golden_tsd, got):
        assert g == _oracle(c), (c["seed"], c["scenario"])
        n += len(g)
    assert n > 10_000


def test_tsd_matches_reference_goldens_directly(ctx, golden_tsd):
    """Without the oracle in between: the multiset of (tsd, distance, length, crc32, short flag) the reference
    returned; only a distance tie straddling the 100-record cut may select different members."""
    got = _run(ctx, golden_tsd)
    for c, g in zip(golden_tsd, got):
        seq = c["seq"]
        mine = collections.Counter()
        for k, s, e, dist, idx, sc in g:
            sub = seq[s:e + 1]
            mine[(seq[e + 1:e + 1 + k], dist, len(sub), zlib.crc32(sub.encode()) & 0xFFFFFFFF, 1 if sc else 0)] += 1
        ref = collections.Counter((e[1], e[2], e[3], e[4], e[5]) for e in c["expect"])
        if mine != ref:
            assert len(g) == 100 and len(c["expect"]) == 100, (c["seed"], mine - ref, ref - mine)
            dmax = max(e[2] for e in c["expect"])
            assert all(t[1] == dmax for t in list((mine - ref).elements()) + list((ref - mine).elements())), c["seed"]


def test_tsd_matches_oracle_on_fresh_seeds(ctx):
    cases = [tsd_synth.synth_tsd_case(s) for s in range(10_000, 11_500)]
    got = _run(ctx, cases)
    for c, g in zip(cases, got):
        assert g == _oracle(c), (c["seed"], c["scenario"])


def test_tsd_edge_inputs(ctx):
    rng = np.random.default_rng(3)
    R = lambda n: "".join(rng.choice(list("ACGT"), int(n)))
    cases = []
    # shortest sequences, boundaries at either end, windows that overlap or cover the whole candidate, N runs, d extremes
    for n, rs, re_, d in [(1, 1, 1, 50), (5, 1, 5, 50), (99, 1, 99, 50), (100, 1, 100, 1), (130, 60, 70, 128), (260, 130, 131, 128),
                          (400, 51, 350, 50), (400, 350, 51, 50), (400, 1, 400, 128), (1200, 51, 1150, 128), (3000, 2, 2999, 3)]:
        cases.append(dict(seq=R(n), raw_start=rs, raw_end=re_, d=d, plant=1, seed=len(cases), scenario="edge"))
    cases.append(dict(seq="N" * 300, raw_start=51, raw_end=250, d=50, plant=0, seed=97, scenario="allN"))
    cases.append(dict(seq="TA" * 200, raw_start=51, raw_end=350, d=50, plant=1, seed=98, scenario="TAx"))
    cases.append(dict(seq="A" * 400, raw_start=51, raw_end=350, d=128, plant=1, seed=99, scenario="polyA"))
    got = _run(ctx, cases)
    for c, g in zip(cases, got):
        assert g == _oracle(c), (c["seed"], c["scenario"], len(c["seq"]))
    recs, cnt = ctx.tsd_candidates([], [], [], 50, 1)
    assert recs.shape[0] == 0 and cnt.shape[0] == 0
    with pytest.raises(Exception):
        ctx.tsd_candidates(["ACGT"], [1], [9], 50, 1)           # boundary outside the sequence
    with pytest.raises(Exception):
        ctx.tsd_candidates(["ACGT" * 50], [1], [100], 500, 1)   # search distance out of range


def test_tsd_full_size_properties(ctx):
    """C2-sized batch (10 000 flanked candidates of 300-3100 bp): sorted, capped, in range, and every record's
    two flanking k-mers are equal -- the property that defines a TSD-bounded variant."""
    rng = np.random.default_rng(11)
    seqs, starts, ends = [], [], []
    bases = np.frombuffer(b"ACGT", dtype=np.uint8)
    for _ in range(10_000):
        n = int(rng.integers(300, 3100))
        a = bases[rng.integers(0, 4, n)]
        k = int(rng.choice([2, 3, 4, 5, 6, 8, 9, 10, 11]))
        a[50 - k:50] = a[n - 50:n - 50 + k]                     # planted TSD around the raw boundaries
        seqs.append(a.tobytes())
        starts.append(51)
        ends.append(n - 50)
    recs, cnt = ctx.tsd_candidates(seqs, starts, ends, 50, 1)
    assert int(cnt.max()) <= 100 and int((cnt > 0).sum()) > 9_000
    t = ctx.timing()
    for i in range(0, 10_000, 7):
        s, r = seqs[i], recs[i][:int(cnt[i])]
        assert np.all(np.diff(r["dist"]) >= 0) and np.all(np.diff(r["idx"]) > 0)
        for x in r:
            k, st, en = int(x["k"]), int(x["start"]), int(x["end"])
            assert 0 <= st - k and en + 1 + k <= len(s) and en - st + 1 >= 100
            assert s[st - k:st] == s[en + 1:en + 1 + k]
            assert int(x["dist"]) == abs(st - 50) + abs(en - (len(s) - 51))
    assert t["kernel_ms"] > 0 and t["launches"] == 1
    print("tsd_kernel: 10 000 candidates in %.3f ms" % t["kernel_ms"])


def test_host_mirror_batch(ctx, golden_tsd):
    """hite_b200.tir_prefilter.search_confident_tir_batch: the reference's dict shapes, record names (module/Util.py:7836)
    and insertion order, with raw boundaries = flanking_len + 1 / len - flanking_len as its caller sets them (:7371-7372)."""
    from hite_b200 import tir_prefilter as TP
    cases = [c for c in golden_tsd if len(c["seq"]) > 110][:120]
    contigs = {"chr1:%d-%d" % (1000 * i, 1000 * i + len(c["seq"])): c["seq"] for i, c in enumerate(cases)}
    contigs["with_gap"] = "ACGT" * 30 + "N" * 10 + "ACGT" * 60      # skipped, Util.py:7368
    for plant in (0, 1):
        allc, short = TP.search_confident_tir_batch(contigs, 50, plant, ctx=ctx)
        assert not any(n.startswith("with_gap") for n in allc) and len(allc) > 3000
        exp_all, exp_short = {}, {}
        for qn, seq in contigs.items():
            if "NNNNNNNNNN" in seq:                                  # Util.py:7368
                continue
            for r in O.search_confident_tir_v4(seq, 51, len(seq) - 50, 50, plant):
                name = "%s-C_%d-tsd_%s-distance_%d" % (qn, r["idx"], seq[r["end"] + 1:r["end"] + 1 + r["k"]], r["dist"])
                exp_all[name] = seq[r["start"]:r["end"] + 1]
                if r["short_class"]:
                    exp_short[name] = exp_all[name]
        assert list(allc.items()) == list(exp_all.items())
        assert list(short.items()) == list(exp_short.items())
        assert TP.get_short_tir_contigs(allc, plant) == exp_short          # string rule == kernel classes
        qn, seq = next(iter(contigs.items()))
        one = TP.search_confident_tir_v4(seq, 51, len(seq) - 50, 50, qn, plant, ctx=ctx)
        assert one == {n: s for n, s in allc.items() if n.startswith(qn + "-C_")}
