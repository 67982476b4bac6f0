"""Seeded synthetic genomes with planted TIR families (SURVEY.md section 8d workloads C2..C5).

Uniform i.i.d. ACGT background; each family is a terminal-inverted-repeat element whose copies are
inserted with target-site duplications, per-base substitutions and short indels.  Candidates are
given the way HiTE's TIR module hands them to ``flank_region_align_v5`` (reference
module/judge_TIR_transposons.py:29-38): one record per chosen copy locus, name ``contig:start-end``,
boundaries exact for most and off by a few bp for the rest (what the dynamic adjustment must fix).
Pure numpy; the same file feeds the CPU and GPU arms.
"""
import numpy as np

_ASCII = np.frombuffer(b"ACGT", dtype=np.uint8)
_COMP = np.zeros(256, dtype=np.uint8)
_COMP[:] = ord("N")
for _a, _b in zip(b"ACGT", b"TGCA"):
    _COMP[_a] = _b


def revcomp_arr(a):
    return _COMP[a[::-1]]


class Workload:
    def __init__(self, contig_names, contig_seqs, cand_names, cand_seqs, truth):
        self.contig_names = contig_names      # list[str]
        self.contig_seqs = contig_seqs        # list[np.uint8 array, ASCII]
        self.cand_names = cand_names
        self.cand_seqs = cand_seqs            # list[bytes]
        self.truth = truth                    # list[dict(family, consensus(bytes), exact(bool))]

    @property
    def genome_bp(self):
        return int(sum(len(s) for s in self.contig_seqs))

    def write_fasta(self, genome_path, cand_path):
        with open(genome_path, "wb") as f:
            for n, s in zip(self.contig_names, self.contig_seqs):
                f.write(b">" + n.encode() + b"\n" + s.tobytes() + b"\n")
        with open(cand_path, "wb") as f:
            for n, s in zip(self.cand_names, self.cand_seqs):
                f.write(b">" + n.encode() + b"\n" + s + b"\n")


def _mutate(rng, cons, sub, indel):
    a = cons.copy()
    if sub > 0:
        m = rng.random(len(a)) < sub
        k = int(m.sum())
        if k:
            # substitute with a different base: add 1..3 in ACGT index space
            idx = np.searchsorted(_ASCII, a[m])  # A C G T are sorted in ASCII
            a[m] = _ASCII[(idx + rng.integers(1, 4, size=k)) % 4]
    n_indel = rng.binomial(len(a), indel) if indel > 0 else 0
    for _ in range(n_indel):
        p = int(rng.integers(25, max(26, len(a) - 25)))
        if rng.random() < 0.5:
            a = np.delete(a, p)
        else:
            a = np.insert(a, p, _ASCII[rng.integers(0, 4)])
    return a


def make_workload(genome_bp=2_000_000, n_contigs=2, n_families=20, cands_per_family=5, copies=(8, 40),
                  length=(200, 2000), sub=(0.0, 0.08), indel=0.004, tsd_choices=(2, 3, 4, 8, 9), seed=1,
                  inexact_frac=0.4, rc_frac=0.5, kind="tir"):
    """Returns a Workload.  ``n_families * cands_per_family`` candidates.  ``kind``: 'tir' (terminal inverted
    repeats + TSD), 'helitron' (5'-TC ... CTRR-3' inserted between A and T, no TSD, forward strand only so the
    candidate carries the motif) or 'non_ltr' (polyA tail, 8-15 bp TSD, forward strand only)."""
    rng = np.random.default_rng(seed)
    clen = genome_bp // n_contigs
    contigs = [_ASCII[rng.integers(0, 4, size=clen, dtype=np.uint8)] for _ in range(n_contigs)]
    names = ["chr%d" % (i + 1) for i in range(n_contigs)]
    slot = length[1] + 400
    slots = [(c, s) for c in range(n_contigs) for s in range(200, clen - slot - 200, slot)]
    order = rng.permutation(len(slots))
    next_slot = 0
    cand_names, cand_seqs, truth = [], [], []
    for fam in range(n_families):
        L = int(rng.integers(length[0], length[1] + 1))
        tir_len = int(rng.integers(10, 31))
        tir = _ASCII[rng.integers(0, 4, size=tir_len)]
        # avoid terminal TA/TTAA-like ambiguity about where the element starts being the common case
        body = _ASCII[rng.integers(0, 4, size=max(0, L - 2 * tir_len))]
        cons = np.concatenate([tir, body, revcomp_arr(tir)])
        if kind == "helitron":
            cons = _ASCII[rng.integers(0, 4, size=L)]
            cons[:2] = np.frombuffer(b"TC", dtype=np.uint8)
            cons[-4:] = np.frombuffer([b"CTAG", b"CTAA", b"CTGG", b"CTGA"][int(rng.integers(0, 4))], dtype=np.uint8)
        elif kind == "non_ltr":
            cons = _ASCII[rng.integers(0, 4, size=L)]
            cons[-int(rng.integers(8, 20)):] = ord("A")
        n_copies = int(rng.integers(copies[0], copies[1] + 1))
        fam_sub = float(rng.uniform(sub[0], sub[1]))
        tsd_len = int(rng.choice(tsd_choices))
        loci = []
        for _ in range(n_copies):
            if next_slot >= len(order):
                break
            c, s = slots[order[next_slot]]
            next_slot += 1
            cp = _mutate(rng, cons, fam_sub, indel)
            strand = "-" if (kind == "tir" and rng.random() < rc_frac) else "+"
            if strand == "-":
                cp = revcomp_arr(cp)
            pos = s + int(rng.integers(60, 200))
            g = contigs[c]
            if kind == "helitron":
                # keep the termini exact (the motif search is exact), insertion site A | T, no TSD
                cp[:2] = cons[:2]; cp[-4:] = cons[-4:]
                g[pos - 1] = ord("A")
                g[pos:pos + len(cp)] = cp
                g[pos + len(cp)] = ord("T")
                loci.append((c, pos, pos + len(cp), strand))
                continue
            if kind == "non_ltr":
                k = int(rng.integers(8, 16))
                ntail = int((cons[::-1] == ord("A")).argmin())
                cp[-ntail:] = ord("A")                   # the polyA test is exact
                g[pos:pos + len(cp)] = cp
                g[pos + len(cp):pos + len(cp) + k] = g[pos - k:pos]
                loci.append((c, pos, pos + len(cp), strand))
                continue
            if tsd_len == 2:
                g[pos - 2:pos] = np.frombuffer(b"TA", dtype=np.uint8)
            elif tsd_len == 3:
                g[pos - 3:pos] = np.frombuffer(b"TAA" if rng.random() < 0.5 else b"TTA", dtype=np.uint8)
            elif tsd_len == 4:
                g[pos - 4:pos] = np.frombuffer(b"TTAA", dtype=np.uint8)
            g[pos:pos + len(cp)] = cp
            g[pos + len(cp):pos + len(cp) + tsd_len] = g[pos - tsd_len:pos]
            loci.append((c, pos, pos + len(cp), strand))
        picks = rng.permutation(len(loci))[:cands_per_family]
        for pi in picks:
            c, a, b, strand = loci[int(pi)]
            exact = rng.random() >= inexact_frac
            if not exact:
                a2 = a - int(rng.integers(-8, 21))
                b2 = b + int(rng.integers(-8, 21))
            else:
                a2, b2 = a, b
            seq = contigs[c][a2:b2].tobytes()
            cand_names.append("%s:%d-%d" % (names[c], a2, b2))
            cand_seqs.append(seq)
            truth.append(dict(family=fam, consensus=cons.tobytes(), exact=exact, strand=strand, n_copies=len(loci)))
    return Workload(names, contigs, cand_names, cand_seqs, truth)
