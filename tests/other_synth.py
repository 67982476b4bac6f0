"""Seeded library-vs-genome workload for the homology non-LTR module tests (module/judge_Other_transposons.py):
a library of LINE/SINE-like consensus sequences and a genome holding full-length, 5'-truncated, internally deleted and
reverse-strand copies of them at known places."""
import numpy as np

_B = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.full(256, ord("N"), dtype=np.uint8)
for _a, _b in zip(b"ACGT", b"TGCA"):
    _COMP[_a] = _b


def rc(a):
    return _COMP[a[::-1]]


def mutate(rng, a, sub):
    a = a.copy()
    m = rng.random(len(a)) < sub
    idx = np.searchsorted(_B, a[m])
    a[m] = _B[(idx + rng.integers(1, 4, size=int(m.sum()))) % 4]
    return a


def make(seed=5, n_lib=40, n_contig=3, contig_bp=400_000, copies=(2, 7), length=(300, 6000), sub=(0.0, 0.08)):
    """-> (contig_names, contig_seqs[bytes], lib_names, lib_seqs[bytes], truth) with truth = list of dicts
    (query, contig, start1, end1, strand, kind, qstart1) for every planted copy (1-based inclusive coordinates)."""
    rng = np.random.default_rng(seed)
    lib = [_B[rng.integers(0, 4, int(rng.integers(*length)))] for _ in range(n_lib)]
    for x in lib[: n_lib // 4]:                       # poly-A tails, as LINE/SINE consensi have
        x[-20:] = ord("A")
    contigs = [_B[rng.integers(0, 4, contig_bp)] for _ in range(n_contig)]
    truth, used = [], [np.zeros(contig_bp, dtype=bool) for _ in range(n_contig)]
    for qi, cons in enumerate(lib):
        for _ in range(int(rng.integers(*copies))):
            kind = rng.choice(["full", "full", "trunc5", "deletion"])
            body = mutate(rng, cons, float(rng.uniform(*sub)))
            qstart = 1
            if kind == "trunc5" and len(cons) > 600:
                cut = int(rng.integers(len(cons) // 10, len(cons) // 2))
                body, qstart = body[cut:], cut + 1
            elif kind == "deletion" and len(cons) > 1200:
                a = int(rng.integers(300, len(cons) - 700))
                body = np.concatenate([body[:a], body[a + 150:]])      # 150 bp missing: two HSPs, chained again
            else:
                kind = "full"
            strand = "+" if rng.random() < 0.5 else "-"
            ins = body if strand == "+" else rc(body)
            for _try in range(50):
                c = int(rng.integers(0, n_contig))
                p = int(rng.integers(200, contig_bp - len(ins) - 200))
                if not used[c][p - 100:p + len(ins) + 100].any():
                    break
            else:
                continue
            used[c][p:p + len(ins)] = True
            contigs[c][p:p + len(ins)] = ins
            truth.append(dict(query=qi, contig=c, start1=p + 1, end1=p + len(ins), strand=strand, kind=str(kind), qstart1=qstart,
                              qlen=len(cons)))
    names = ["chr%d" % (i + 1) for i in range(n_contig)]
    lib_names = ["fam%d#LINE/L1" % i if i % 2 == 0 else "fam%d#SINE/tRNA" % i for i in range(n_lib)]
    return names, [c.tobytes() for c in contigs], lib_names, [x.tobytes() for x in lib], truth
