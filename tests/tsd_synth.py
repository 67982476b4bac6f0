"""Seeded synthetic flanked TIR candidates for the TSD-bounded variant generation tests
(shared by tests/golden/make_golden_tsd.py, tests/test_tsd_oracle.py and tests/test_tsd_gpu.py)."""
import numpy as np

KS = [2, 3, 4, 5, 6, 8, 9, 10, 11]


def rc(s):
    return s[::-1].translate(str.maketrans("ACGT", "TGCA"))


def synth_tsd_case(seed):
    rng = np.random.default_rng(seed)
    R = lambda n: "".join(rng.choice(list("ACGT"), int(n)))
    d = int(rng.choice([50, 50, 50, 20, 100]))
    flank = int(rng.choice([d, d, d + 7, max(3, d - 9)]))          # what the caller believes vs what is there
    L = int(rng.integers(90, 620))
    scenario = ["tsd", "tsd", "tsd", "lowcomplex", "nn", "tgca", "tata", "ccc", "short", "edge"][seed % 10]
    plant = int(rng.integers(0, 2))
    tir = R(rng.integers(5, 16))
    body = tir + R(max(1, L - 2 * len(tir))) + rc(tir)
    k = int(rng.choice(KS))
    tsd = R(k)
    if k == 2:
        tsd = "TA" if rng.random() < 0.7 else tsd
    if k == 4:
        tsd = "TTAA" if rng.random() < 0.7 else tsd
    left, right = R(flank), R(flank)
    if scenario == "lowcomplex":
        left = ("TA" * flank)[:flank]
        right = ("AT" * flank)[:flank] if rng.random() < 0.5 else ("TA" * flank)[:flank]
    if scenario == "nn":
        p = int(rng.integers(0, max(1, flank - 12)))
        left = left[:p] + "N" * 12 + left[p + 12:]
        right = right[:p] + "N" * 12 + right[p + 12:]
    if scenario == "tgca":
        body = "TG" + body[2:-2] + "CA"
    if scenario == "tata":
        body = ("TATATATA" if rng.random() < 0.5 else "ATATATAT") + body[8:]
    if scenario == "ccc":
        plant = 0
        body = "CCC" + body[3:-3] + "GGG"
        tsd = R(2)
    if scenario == "short":
        body = body[:int(rng.integers(40, 110))]
    off_l, off_r = int(rng.integers(-6, 7)), int(rng.integers(-6, 7))
    seq = left[:len(left) - len(tsd)] + tsd + body + tsd + right[len(tsd):]
    raw_start = len(left) + 1 + off_l
    raw_end = len(left) + len(body) + off_r
    if scenario == "edge":
        raw_start = int(rng.choice([1, 2, 3]))
        raw_end = len(seq) - int(rng.integers(0, 3))
    raw_start = max(1, min(len(seq), raw_start))
    raw_end = max(1, min(len(seq), raw_end))
    return dict(seed=seed, scenario=scenario, seq=seq, raw_start=raw_start, raw_end=raw_end, d=d, plant=plant)
