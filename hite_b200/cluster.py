"""HB-cdhit: the stand-in for ``cd-hit-est -aS .95 -aL .95 -c .8 -G 0 -g 1 -A 80`` at the stages' merge points (reference
module/judge_TIR_transposons.py:87, judge_Helitron_transposons.py:63-76 / :108-121, get_nonRedundant_lib.py:34-49, :141).

cd-hit-est is a third-party binary (``environment.yml:27``, unpinned) absent from the reference repository and from this
image, so its raw clusters cannot be reproduced: PARITY UNPINNED.  What is kept is its published procedure and exactly the
options the reference passes:

* sequences are processed longest first; each either joins an existing cluster's representative or founds a new cluster;
* a sequence joins representative r only through an alignment with identity >= ``-c`` (``-G 0``: local identity = identical
  columns / alignment columns), covering >= ``-aS`` of the shorter (itself) and >= ``-aL`` of the longer (r) sequence, of
  at least ``-A`` aligned bases; both strands (``-r 1``, the default);
* ``-g 1``: among the qualifying representatives the most similar one, not the first.

The alignments come from the kernels of the boundary path: the sequence set itself is loaded as the "genome"
(``hb_genome_load`` + ``hb_index_build``) and every sequence is a query of ``hb_query_hsps`` (seeding + banded extension),
which is the all-vs-all; the greedy pass over the HSP list is host code.  Output contract of the binary: ``-o <file>`` =
the representatives as FASTA (the first header word is what HiTE reads), ``<file>.clstr`` in cd-hit's format.
``module/cd-hit-est`` of this repository is the executable of that name for ``$PATH``.  There is no CPU fallback.
"""
import numpy as np

from . import boundary

MAX_LEN = 32000


def read_records(path):
    """(header line without '>', upper-cased sequence) in file order -- cd-hit keeps the whole header."""
    recs, name, parts = [], None, []
    with open(path) as f:
        for line in f:
            line = line.rstrip("\n")
            if line.startswith(">"):
                if name is not None:
                    recs.append((name, "".join(parts).upper()))
                name, parts = line[1:], []
            elif name is not None:
                parts.append(line.strip())
    if name is not None:
        recs.append((name, "".join(parts).upper()))
    return recs


def greedy_clusters(lengths, hsps, c=0.8, aS=0.95, aL=0.95, A=80, g=1):
    """``hsps``: structured array (query, contig, qs, qe, ss, se, n_match, aln_len) of the all-vs-all.  Returns
    (order, rep_of, ident_of, strand_of): ``order`` = processing order (longest first, input order on ties),
    ``rep_of[i]`` = representative of sequence i (itself when it founds a cluster), identity and strand of the joining
    alignment."""
    n = len(lengths)
    order = sorted(range(n), key=lambda i: (-lengths[i], i))
    rank = np.empty(n, dtype=np.int64)
    rank[order] = np.arange(n)
    by_q = {}
    if len(hsps):
        q = hsps["query"].astype(np.int64)
        s = hsps["contig"].astype(np.int64)
        ok = (q != s) & (hsps["aln_len"] > 0)
        ident = hsps["n_match"] / np.maximum(hsps["aln_len"], 1)
        qcov = (hsps["qe"] - hsps["qs"] + 1) / np.asarray(lengths, dtype=np.float64)[q]
        span = np.abs(hsps["se"].astype(np.int64) - hsps["ss"]) + 1
        scov = span / np.asarray(lengths, dtype=np.float64)[s]
        ok &= (ident >= c) & (qcov >= aS) & (scov >= aL) & (hsps["qe"] - hsps["qs"] + 1 >= A) & (span >= A)
        ok &= rank[s] < rank[q]                                   # only longer (earlier) sequences can be representatives
        for k in np.flatnonzero(ok):
            by_q.setdefault(int(q[k]), []).append((float(ident[k]), int(s[k]), bool(hsps["ss"][k] > hsps["se"][k])))
    rep_of = list(range(n))
    ident_of = [1.0] * n
    strand_of = [False] * n
    is_rep = [False] * n
    for i in order:
        best = None
        for (idn, s, minus) in by_q.get(i, ()):
            if not is_rep[s]:
                continue
            if best is None or (g and idn > best[0]) or (not g and rank[s] < rank[best[1]]):
                best = (idn, s, minus)
        if best is None:
            is_rep[i] = True
        else:
            ident_of[i], rep_of[i], strand_of[i] = best
    return order, rep_of, ident_of, strand_of


def cd_hit_est(input_path, output_path, c=0.8, aS=0.95, aL=0.95, A=80, g=1, ctx=None, max_loci=300):
    """The command the reference runs, as a function; returns the number of clusters."""
    recs = read_records(input_path)
    recs = [r for r in recs if r[1]]
    usable = [i for i, (_, s) in enumerate(recs) if 3 <= len(s) <= MAX_LEN]
    lengths = [len(s) for _, s in recs]
    hs = np.zeros(0, dtype=[("query", "<u4"), ("contig", "<i4"), ("qs", "<i4"), ("qe", "<i4"), ("ss", "<i4"), ("se", "<i4"),
                            ("n_match", "<i4"), ("aln_len", "<i4"), ("score", "<i4"), ("rank", "<i4")])
    if len(usable) > 1:
        if ctx is None:
            ctx = boundary.get_context(role="members")           # never evicts the stage's reference genome
        seqs = [recs[i][1].encode("ascii") for i in usable]
        ctx.load_genome(seqs)
        ctx.build_index(15, 10)
        parts = []
        for b in range(0, len(seqs), 8192):
            h = ctx.query_hsps(seqs[b:b + 8192], max_loci=max_loci).copy()
            h["query"] += b
            parts.append(h)
        hs = np.concatenate(parts)
        u = np.asarray(usable, dtype=np.int64)                    # back to record numbers
        hs = hs.copy()
        hs["query"] = u[hs["query"].astype(np.int64)]
        hs["contig"] = u[hs["contig"].astype(np.int64)]
    order, rep_of, ident_of, strand_of = greedy_clusters(lengths, hs, c, aS, aL, A, g)
    reps = [i for i in order if rep_of[i] == i]
    members = {r: [] for r in reps}
    for i in order:
        members[rep_of[i]].append(i)
    with open(output_path, "w") as f:
        for r in reps:
            f.write(">%s\n%s\n" % (recs[r][0], recs[r][1]))
    with open(output_path + ".clstr", "w") as f:
        for k, r in enumerate(reps):
            f.write(">Cluster %d\n" % k)
            for m, i in enumerate(members[r]):
                name = recs[i][0].split()[0] if recs[i][0].split() else recs[i][0]
                if i == r:
                    f.write("%d\t%dnt, >%s... *\n" % (m, lengths[i], name))
                else:
                    f.write("%d\t%dnt, >%s... at %s/%.2f%%\n" % (m, lengths[i], name, "-" if strand_of[i] else "+", 100.0 * ident_of[i]))
    return len(reps)


def main(argv=None):
    """``cd-hit-est`` command line as the reference writes it (module/judge_TIR_transposons.py:87): -i -o -c -aS -aL -G -g -A
    (-T / -M / anything else is accepted and ignored; ``-G 1`` global identity is not implemented and refused)."""
    import sys
    argv = list(sys.argv[1:] if argv is None else argv)
    opt = {"-c": 0.9, "-aS": 0.0, "-aL": 0.0, "-A": 0, "-g": 0, "-G": 1}
    inp = out = None
    i = 0
    while i < len(argv):
        a = argv[i]
        if a in ("-i", "-o") and i + 1 < len(argv):
            if a == "-i":
                inp = argv[i + 1]
            else:
                out = argv[i + 1]
            i += 2
        elif a in opt and i + 1 < len(argv):
            opt[a] = float(argv[i + 1]); i += 2
        elif a.startswith("-") and i + 1 < len(argv) and not argv[i + 1].startswith("-"):
            i += 2
        else:
            i += 1
    if not inp or not out:
        sys.stderr.write("usage: cd-hit-est -i in.fa -o out.fa [-c .8 -aS .95 -aL .95 -G 0 -g 1 -A 80]\n")
        return 2
    if int(opt["-G"]) != 0:
        sys.stderr.write("hite_b200 cd-hit-est: only -G 0 (local identity, what HiTE passes) is implemented\n")
        return 2
    n = cd_hit_est(inp, out, c=opt["-c"], aS=opt["-aS"], aL=opt["-aL"], A=int(opt["-A"]), g=int(opt["-g"]))
    print("hite_b200 cd-hit-est: %s -> %d clusters" % (inp, n))
    return 0
