"""GPU parity: the non-LTR candidate generator's scan through the C-ABI (hb_nonltr_candidates) vs the C oracle (bit-exact
records), vs the reference goldens directly (through the host mirror), and the mirror's file contract."""
import gzip
import json
import os
import zlib

import numpy as np
import pytest

from oracle import oracle as O
from polya_synth import polya_case
from test_polya_oracle import candidates_from

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FIELDS = ("found", "direct", "end_5", "end_3", "tsd_k", "seq_start", "seq_end")


@pytest.fixture(scope="module")
def golden_polya():
    with gzip.open(os.path.join(ROOT, "tests", "golden", "polya_tsd.json.gz")) as f:
        return json.loads(f.read())


def _check_vs_oracle(ctx, cases):
    by_flank = {}
    for c in cases:
        by_flank.setdefault(c["flank"], []).append(c)
    n = 0
    for flank, sub in by_flank.items():
        recs = ctx.nonltr_candidates([c["seq"] for c in sub], flank)
        for c, r in zip(sub, recs):
            exp = O.search_polyA_TSD(c["seq"], flank)[3]
            assert {f: int(r[f]) for f in FIELDS} == exp, (c["seed"], c["scenario"])
            n += int(r["found"])
    return n


def test_polya_matches_oracle(ctx, golden_polya):
    assert _check_vs_oracle(ctx, [polya_case(g["seed"]) for g in golden_polya["cases"]]) > 100
    assert _check_vs_oracle(ctx, [polya_case(s) for s in range(50_000, 52_000)]) > 500


def test_polya_odd_inputs(ctx):
    rng = np.random.default_rng(0)
    cases = []
    for it in range(3000):
        L = int(rng.integers(1, 260))
        alphabet = list("ACGT") if rng.random() < 0.5 else list("AAATTTCGN")
        s = "".join(rng.choice(alphabet, L))
        if rng.random() < 0.5 and L > 40:
            p = int(rng.integers(0, L - 20)); s = s[:p] + "A" * int(rng.integers(6, 20)) + s[p + 10:]
        if rng.random() < 0.5 and L > 40:
            p = int(rng.integers(0, L - 20)); s = s[:p] + "T" * int(rng.integers(6, 20)) + s[p + 10:]
        cases.append(dict(seed=it, scenario="odd", seq=s, flank=int(rng.choice([0, 5, 25, 30, 50, 60]))))
    _check_vs_oracle(ctx, cases)
    assert len(ctx.nonltr_candidates([], 50)) == 0
    with pytest.raises(Exception):
        ctx.nonltr_candidates(["ACGT"], 5000)


def test_mirror_matches_reference_goldens(ctx, golden_polya, tmp_path):
    from hite_b200 import nonltr_prefilter as NP
    cases = [polya_case(g["seed"]) for g in golden_polya["cases"]]
    for g, c in list(zip(golden_polya["cases"], cases))[::7]:
        found, tsd, body = NP.search_polyA_TSD(c["seq"], c["flank"], 25, list(range(8, 21)), ctx=ctx)
        assert [found, tsd, len(body), zlib.crc32(body.encode()) & 0xFFFFFFFF] == g["expect"], g["seed"]
    with pytest.raises(Exception):
        NP.search_polyA_TSD("ACGT" * 100, 50, 30, list(range(8, 21)), ctx=ctx)
    sub = [c for c in cases if c["flank"] == 50]
    path = str(tmp_path / "longest_repeats_0.flanked.fa")
    with open(path, "w") as f:
        for c in sub:
            f.write(">chr1:%d-%d\n%s\n" % (c["seed"] * 10000, c["seed"] * 10000 + len(c["seq"]), c["seq"]))
    sine, line = NP.get_candidate_non_LTR(path, 50, ctx=ctx)
    pack = lambda d: [[n, len(s), zlib.crc32(s.encode()) & 0xFFFFFFFF] for n, s in d.items()]
    assert pack(sine) == golden_polya["sine"] and pack(line) == golden_polya["line"]
    sp, lp = NP.get_candidate_non_ltr_parallel(path, str(tmp_path), 8, ctx=ctx)
    assert sp.endswith("/candidate_SINE.fa") and lp.endswith("/candidate_LINE.fa")
    assert open(sp).read() == "".join(">%s\n%s\n" % (n, s) for n, s in sine.items())
    assert open(lp).read() == "".join(">%s\n%s\n" % (n, s) for n, s in line.items())


def test_polya_full_size_properties(ctx):
    """C2-sized batch: 10 000 flanked candidates with a planted poly-A tail and 12 bp TSD -> found, boundaries exact."""
    rng = np.random.default_rng(5)
    B = np.frombuffer(b"ACGT", dtype=np.uint8)
    seqs, truth = [], []
    for _ in range(10_000):
        L = int(rng.integers(150, 3000))
        body = B[rng.integers(0, 4, L)].tobytes() + b"A" * int(rng.integers(10, 25))
        tsd = B[rng.integers(0, 4, 12)].tobytes()
        left, right = B[rng.integers(0, 4, 38)].tobytes(), B[rng.integers(0, 4, 38)].tobytes()
        seqs.append(left + tsd + body + tsd + right)
        truth.append((50, 50 + len(body)))
    recs = ctx.nonltr_candidates(seqs, 50)
    tm = ctx.timing()
    ok = 0
    for r, (a, b) in zip(recs, truth):
        ok += int(r["found"] == 1 and r["direct"] == 1 and r["tsd_k"] >= 8 and abs(int(r["seq_start"]) - a) <= 3 and abs(int(r["seq_end"]) - b) <= 6)
    assert ok > 8_000, ok
    assert tm["launches"] == 1 and tm["kernel_ms"] > 0
    print("polya_tsd_kernel: 10 000 candidates in %.3f ms" % tm["kernel_ms"])
