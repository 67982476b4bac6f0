"""HB-cdhit host logic (hite_b200/cluster.py) with the all-vs-all alignments from the oracle: planted families cluster the way
`cd-hit-est -aS .95 -aL .95 -c .8 -G 0 -g 1 -A 80` is documented to cluster them, files follow the binary's contract, the
command line the reference writes is understood."""
import os

import numpy as np

from oracle import oracle as O


class OracleBackedCtx:
    device = 0

    def load_genome(self, seqs, names=None):
        self.G = O.Genome(seqs)

    def build_index(self, k, w):
        pass

    def query_hsps(self, seqs, occ_cap=512, max_loci=300):
        from hite_b200 import _lib
        rows = []
        for qi, q in enumerate(seqs):
            for h in self.G.query_hsps(q, occ_cap, max_loci):
                rows.append((qi, h["contig"], h["qs"], h["qe"], h["ss"], h["se"], h["n_match"], h["aln_len"], h["score"], h["rank"]))
        return np.array(rows, dtype=_lib.QHSP_DTYPE) if rows else np.zeros(0, dtype=_lib.QHSP_DTYPE)


def _families(seed=5):
    rng = np.random.default_rng(seed)
    B = np.frombuffer(b"ACGT", dtype=np.uint8)
    comp = bytes.maketrans(b"ACGT", b"TGCA")
    recs, fam_of = [], []
    for f in range(6):
        cons = B[rng.integers(0, 4, int(rng.integers(400, 1500)))].copy()
        for m in range(int(rng.integers(3, 7))):
            s = cons.copy()
            mask = rng.random(len(s)) < 0.06                      # ~6 % substitutions: well above -c 0.8
            s[mask] = B[rng.integers(0, 4, int(mask.sum()))]
            seq = s.tobytes()
            if rng.random() < 0.4:
                seq = seq.translate(comp)[::-1]                    # either strand clusters (-r 1)
            recs.append(("fam%d_m%d some description" % (f, m), seq.decode()))
            fam_of.append(f)
        recs.append(("fam%d_frag" % f, cons[:len(cons) // 2].tobytes().decode()))        # a half-length fragment: -aL .95 keeps it apart
        fam_of.append(100 + f)
    return recs, fam_of


def test_planted_families_cluster_and_files_follow_the_binary(tmp_path):
    from hite_b200 import cluster
    recs, fam_of = _families()
    inp, out = str(tmp_path / "in.fa"), str(tmp_path / "out.cons.fa")
    with open(inp, "w") as f:
        for n, s in recs:
            f.write(">%s\n" % n)
            for o in range(0, len(s), 70):
                f.write(s[o:o + 70].lower() + "\n")                # multi-line, lower case: read like any FASTA
    n = cluster.cd_hit_est(inp, out, c=0.8, aS=0.95, aL=0.95, A=80, g=1, ctx=OracleBackedCtx())
    assert n == 12                                                  # 6 families + their 6 fragments
    reps = cluster.read_records(out)
    assert len(reps) == 12 and all(h in dict(recs) for h, _ in reps)          # whole header lines kept
    assert [len(s) for _, s in reps] == sorted((len(s) for _, s in reps), reverse=True)      # longest first
    clstr = open(out + ".clstr").read().split(">Cluster ")[1:]
    assert len(clstr) == 12
    seen = set()
    for block in clstr:
        lines = block.strip().splitlines()[1:]
        names = [l.split(">")[1].split("...")[0] for l in lines]
        fams = {fam_of[[r[0].split()[0] for r in recs].index(nm)] for nm in names}
        assert len(fams) == 1                                       # never two families in one cluster
        assert sum(l.rstrip().endswith("*") for l in lines) == 1 and lines[0].rstrip().endswith("*")
        assert all(" at +/" in l or " at -/" in l for l in lines[1:])
        seen |= set(names)
    assert len(seen) == len(recs)


def test_greedy_rules():
    from hite_b200 import cluster
    dt = [("query", "<u4"), ("contig", "<i4"), ("qs", "<i4"), ("qe", "<i4"), ("ss", "<i4"), ("se", "<i4"), ("n_match", "<i4"),
          ("aln_len", "<i4"), ("score", "<i4"), ("rank", "<i4")]
    L = [1000, 990, 980, 500, 970]
    hs = np.array([(1, 0, 1, 990, 1, 1000, 900, 1000, 0, 0),        # 1 -> 0: identity .9, covers both
                   (2, 0, 1, 980, 5, 990, 800, 990, 0, 0),          # 2 -> 0: identity .808
                   (2, 1, 1, 980, 1, 985, 960, 985, 0, 1),          # 2 -> 1: better identity, but 1 is not a representative
                   (3, 0, 1, 500, 1, 500, 500, 500, 0, 0),          # 3 covers only half of 0: -aL fails
                   (4, 0, 1, 970, 1000, 20, 700, 990, 0, 0)],       # 4 -> 0 minus strand, identity .707 < -c
                  dtype=dt)
    order, rep, ident, strand = cluster.greedy_clusters(L, hs, 0.8, 0.95, 0.95, 80, 1)
    assert order == [0, 1, 2, 4, 3]
    assert rep == [0, 0, 0, 3, 4] and abs(ident[2] - 800 / 990) < 1e-12
    hs2 = hs.copy(); hs2["n_match"][4] = 900
    assert cluster.greedy_clusters(L, hs2, 0.8, 0.95, 0.95, 80, 1)[1][4] == 0 and cluster.greedy_clusters(L, hs2, 0.8, 0.95, 0.95, 80, 1)[3][4]


def test_command_line_of_the_reference(tmp_path, monkeypatch):
    from hite_b200 import cluster
    recs, _ = _families(seed=6)
    inp, out = str(tmp_path / "tir_tsd_0.fa"), str(tmp_path / "tir_tsd_0.cons.fa")
    with open(inp, "w") as f:
        for n, s in recs:
            f.write(">%s\n%s\n" % (n, s))
    ctx = OracleBackedCtx()
    monkeypatch.setattr(cluster.boundary, "get_context", lambda device=None, role="reference": ctx)
    rc = cluster.main("-aS 0.95 -aL 0.95 -c 0.8 -G 0 -g 1 -A 80 -i %s -o %s -T 0 -M 0" % (inp, out)).__int__() if False else \
        cluster.main(("-aS 0.95 -aL 0.95 -c 0.8 -G 0 -g 1 -A 80 -i %s -o %s -T 0 -M 0" % (inp, out)).split())
    assert rc == 0 and os.path.getsize(out) > 0 and os.path.exists(out + ".clstr")
    assert cluster.main(["-i", inp, "-o", out, "-G", "1"]) == 2                    # global identity: refused, not silently different
