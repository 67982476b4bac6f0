"""Seeded flanked candidates for the non-LTR candidate generator tests (reference search_polyA_TSD, module/Util.py:10915):
poly-A / tandem tails at the 3' end or poly-T / tandem heads at the 5' end, TSDs of 8-20 bp with an optional mismatch or
indel, boundaries off by a few bases, N runs, tails missing."""
import numpy as np


def rc(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


def polya_case(seed):
    rng = np.random.default_rng(seed)
    R = lambda n: "".join(rng.choice(list("ACGT"), int(n)))
    flank = int(rng.choice([50, 50, 50, 30, 60]))
    L = int(rng.integers(100, 1500))
    scenario = ["polyA", "polyA", "tandem", "polyT", "tandemT", "none", "shortA", "nn", "edge", "both"][seed % 10]
    body = R(L)
    k = int(rng.integers(8, 21))
    tsd = R(k)
    tsd2 = tsd
    r = rng.random()
    if r < 0.25:                                          # one substitution
        p = int(rng.integers(0, k)); tsd2 = tsd[:p] + rc(tsd[p])[0] + tsd[p + 1:]
    elif r < 0.4:                                         # one deletion
        p = int(rng.integers(0, k)); tsd2 = tsd[:p] + tsd[p + 1:]
    elif r < 0.5:                                         # two substitutions: no match
        tsd2 = rc(tsd[0]) + tsd[1:-1] + rc(tsd[-1])
    motif = R(int(rng.integers(2, 7)))
    tailA = "A" * int(rng.integers(8, 30))
    if scenario == "tandem":
        tail = motif * int(rng.integers(4, 8))
    elif scenario == "shortA":
        tail = "A" * int(rng.integers(3, 8))
    elif scenario == "none":
        tail = ""
    else:
        tail = tailA
    if scenario in ("polyT", "tandemT", "both"):
        head = ("T" * int(rng.integers(8, 30))) if scenario != "tandemT" else motif * int(rng.integers(4, 8))
        elem = head + body + (tail if scenario == "both" else "")
    else:
        elem = body + tail
    left, right = R(flank + 20), R(flank + 20)
    if scenario == "nn":
        tsd = tsd[:3] + "NN" + tsd[5:]; tsd2 = tsd
    seq = left + tsd + elem + tsd2 + right
    off = int(rng.integers(-12, 13))
    # the caller's flank is what it believes; the real element starts at len(left)+k
    cut_l = len(left) + k - flank + off
    cut_r = len(right) + len(tsd2) - flank + int(rng.integers(-12, 13))
    seq = seq[max(0, cut_l):len(seq) - max(0, cut_r)]
    if scenario == "edge":
        seq = seq[:int(rng.integers(60, 140))]
    return dict(seed=seed, scenario=scenario, seq=seq, flank=flank)
