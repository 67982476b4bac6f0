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


# ------------------------------------------------------------------------------------------------
# SURVEY.md section 8(d) workloads at spec: mixed families, U / Zipf family sizes, per-family substitution and
# indel rates, block indels, nested insertions, repeat-content targets.  Copies are generated per family as
# one matrix (vectorised substitutions), laid out over the contigs in shuffled order with random gaps.

def _family_consensus(rng, kind, L):
    if kind == "tir":
        tir_len = int(rng.integers(10, 31))
        tir = _ASCII[rng.integers(0, 4, size=tir_len)]
        body = _ASCII[rng.integers(0, 4, size=max(0, L - 2 * tir_len))]
        return np.concatenate([tir, body, revcomp_arr(tir)])
    if kind == "ltr":                                       # LTR-like: direct repeats TG..CA around an internal region
        ltr_len = int(rng.integers(200, 601))
        ltr = _ASCII[rng.integers(0, 4, size=ltr_len)]
        ltr[:2] = np.frombuffer(b"TG", dtype=np.uint8)
        ltr[-2:] = np.frombuffer(b"CA", dtype=np.uint8)
        body = _ASCII[rng.integers(0, 4, size=max(0, L - 2 * ltr_len))]
        return np.concatenate([ltr, body, ltr])
    cons = _ASCII[rng.integers(0, 4, size=L)]
    if kind == "helitron":
        cons[:2] = np.frombuffer(b"TC", dtype=np.uint8)
        cons[-4:] = np.frombuffer([b"CTAG", b"CTAA", b"CTGG", b"CTGA"][int(rng.integers(0, 4))], dtype=np.uint8)
    elif kind == "non_ltr":
        cons[-int(rng.integers(8, 20)):] = ord("A")
    return cons


def _family_copies(rng, cons, n, sub, indel, block):
    """n mutated copies of ``cons``: substitutions as one n x L matrix, then per copy one vectorised deletion mask and one
    multi-position insertion, then ``block`` = (count, lo, hi) block indels of lo..hi bp (half deletions, half insertions)."""
    L = len(cons)
    m = np.tile(cons, (n, 1))
    if sub > 0:
        mask = rng.random((n, L)) < sub
        k = int(mask.sum())
        if k:
            idx = np.searchsorted(_ASCII, m[mask])
            m[mask] = _ASCII[(idx + rng.integers(1, 4, size=k)) % 4]
    out = []
    for r in range(n):
        a = m[r]
        if indel > 0:
            keep = rng.random(L) >= indel / 2
            keep[:25] = True; keep[-25:] = True
            a = a[keep]
            n_ins = int(rng.binomial(len(a), indel / 2))
            if n_ins and len(a) > 60:
                a = np.insert(a, rng.integers(25, len(a) - 25, size=n_ins), _ASCII[rng.integers(0, 4, size=n_ins)])
        if block is not None:
            for _ in range(int(block[0])):
                bl = int(rng.integers(block[1], block[2] + 1))
                if len(a) < 2 * bl + 120:
                    continue
                p = int(rng.integers(50, len(a) - bl - 50))
                if rng.random() < 0.5:
                    a = np.concatenate([a[:p], a[p + bl:]])
                else:
                    a = np.concatenate([a[:p], _ASCII[rng.integers(0, 4, size=bl)], a[p:]])
        out.append(a)
    return out


def _zipf_sizes(rng, s, cap, n):
    return np.minimum(cap, np.maximum(2, rng.zipf(s, size=n)))


def make_mixed_workload(genome_bp, n_contigs, n_candidates, mix=(("tir", 1.0),), lengths=None, copies=("uniform", 6, 100),
                        sub=(0.0, 0.10), indel=(0.0, 0.02), block=None, nested_frac=0.0, repeat_frac=None, cands_per_family=20,
                        tsd_choices=(2, 3, 4, 8, 9), inexact_frac=0.4, seed=1, contig_names=None, min_gap=120):
    """SURVEY 8(d)-style workload.  ``mix``: (kind, fraction of families) with kind in tir / ltr / helitron / non_ltr;
    ``lengths``: {kind: (lo, hi)}; ``copies``: ("uniform", lo, hi) or ("zipf", s, cap); ``sub`` / ``indel``: per-family rate
    ranges; ``block``: (count, lo, hi) block indels per copy; ``nested_frac``: fraction of copies spliced INTO another
    copy (the host copy is interrupted, the nested one intact); families are added until ``n_candidates`` candidates
    exist and, when ``repeat_frac`` is given, until the copies fill that fraction of the genome.  Returns a Workload whose
    ``truth`` has one dict per candidate (family, kind, consensus, exact, strand, n_copies) and whose ``copies_truth``
    lists every planted intact copy as (family, contig, start, end, strand)."""
    rng = np.random.default_rng(seed)
    lengths = dict({"tir": (200, 3000), "ltr": (4000, 10000), "helitron": (500, 3000), "non_ltr": (300, 3000)}, **(lengths or {}))
    kinds = [k for k, _ in mix]
    probs = np.array([p for _, p in mix], dtype=float)
    probs /= probs.sum()
    target_bp = int(genome_bp * repeat_frac) if repeat_frac else 0
    budget_bp = int(genome_bp * 0.95) - 1000 * n_contigs - 20000
    fams, items = [], []            # items: dict(fam, seq, strand, kind, tsd, nested=[...])
    n_cand_avail, total_bp = 0, 0
    while (n_cand_avail < n_candidates or total_bp < target_bp):
        kind = kinds[int(rng.choice(len(kinds), p=probs))]
        lo, hi = lengths[kind]
        L = int(rng.integers(lo, hi + 1))
        cons = _family_consensus(rng, kind, L)
        if copies[0] == "uniform":
            n = int(rng.integers(copies[1], copies[2] + 1))
        else:
            n = int(_zipf_sizes(rng, copies[1], copies[2], 1)[0])
        if total_bp + n * (int(L * 1.03) + min_gap + 50) > budget_bp:
            n = max(0, (budget_bp - total_bp) // (int(L * 1.03) + min_gap + 50))
            if n < 2:
                break
        f_sub = float(rng.uniform(sub[0], sub[1]))
        f_indel = float(rng.uniform(indel[0], indel[1]))
        tsd_len = {"tir": int(rng.choice(tsd_choices)), "ltr": 5, "helitron": 0, "non_ltr": int(rng.integers(8, 16))}[kind]
        fam = len(fams)
        fams.append(dict(kind=kind, consensus=cons.tobytes(), n_copies=n, sub=f_sub, indel=f_indel, tsd=tsd_len))
        for a in _family_copies(rng, cons, n, f_sub, f_indel, block):
            strand = "-" if (kind in ("tir", "ltr") and rng.random() < 0.5) else "+"
            if kind == "helitron":
                a[:2] = cons[:2]; a[-4:] = cons[-4:]
            elif kind == "non_ltr":
                ntail = int((cons[::-1] == ord("A")).argmin())
                a[-ntail:] = ord("A")
            if strand == "-":
                a = revcomp_arr(a)
            items.append(dict(fam=fam, seq=a, strand=strand, kind=kind, tsd=tsd_len, nested=None))
            total_bp += len(a) + 2 * tsd_len + 4 + min_gap
        n_cand_avail += min(n, cands_per_family)
    # nested insertions: splice a copy into the middle third of another (longer) copy; the host is interrupted
    top = list(range(len(items)))
    if nested_frac > 0 and len(items) > 4:
        n_nest = int(len(items) * nested_frac)
        perm = rng.permutation(len(items))
        guests = set(int(x) for x in perm[:n_nest])
        hosts = [int(x) for x in perm[n_nest:]]
        hi_ = 0
        for gidx in sorted(guests):
            host = hosts[hi_ % len(hosts)]; hi_ += 1
            if items[host]["nested"] is not None or len(items[host]["seq"]) < 300:
                continue
            items[host]["nested"] = gidx
        nested_ids = {it["nested"] for it in items if it["nested"] is not None}
        top = [i for i in range(len(items)) if i not in nested_ids]
    order = [top[int(i)] for i in rng.permutation(len(top))]
    # layout: items dealt to contigs round-robin by cumulative length; gaps share the free space of the contig
    clen = genome_bp // n_contigs
    names = list(contig_names) if contig_names is not None else ["chr%d" % (i + 1) for i in range(n_contigs)]
    per_contig = [[] for _ in range(n_contigs)]
    used = np.zeros(n_contigs, dtype=np.int64)

    def full_len(i):
        it = items[i]
        n = len(it["seq"]) + 2 * it["tsd"] + 4
        if it["nested"] is not None:
            g = items[it["nested"]]
            n += len(g["seq"]) + 2 * g["tsd"] + 4
        return n
    for i in order:
        c = int(np.argmin(used))
        per_contig[c].append(i)
        used[c] += full_len(i) + min_gap
    contigs, placed = [], {}
    for c in range(n_contigs):
        g = _ASCII[rng.integers(0, 4, size=clen, dtype=np.uint8)]
        free = clen - int(used[c]) - 400
        if free < 0:
            raise ValueError("workload does not fit: contig %d over by %d bp" % (c, -free))
        n_it = max(1, len(per_contig[c]))
        w = rng.random(n_it)
        extra = np.floor(w / w.sum() * free).astype(np.int64)
        cur = 200
        for x, i in enumerate(per_contig[c]):
            cur += min_gap + int(extra[x])
            cur = _place(rng, g, cur, items, i, placed, c)
        contigs.append(g)
    # candidates: up to cands_per_family intact copies of every family, boundaries exact or off by a few bp
    by_fam = {}
    for i, (c, a, b) in placed.items():
        by_fam.setdefault(items[i]["fam"], []).append((i, c, a, b))
    cand_names, cand_seqs, truth, copies_truth = [], [], [], []
    for fam in sorted(by_fam):
        lst = sorted(by_fam[fam])
        hosts_interrupted = {i for i, *_ in lst if items[i]["nested"] is not None}
        intact = [t for t in lst if t[0] not in hosts_interrupted]
        for (i, c, a, b) in intact:
            copies_truth.append((fam, c, a, b, items[i]["strand"]))
        picks = rng.permutation(len(intact))[:cands_per_family]
        for pi in picks:
            if len(cand_seqs) >= n_candidates:
                break
            i, c, a, b = intact[int(pi)]
            exact = rng.random() >= inexact_frac
            a2, b2 = (a, b) if exact else (a - int(rng.integers(-8, 21)), b + int(rng.integers(-8, 21)))
            a2 = max(0, a2); b2 = min(clen, b2)
            cand_names.append("%s:%d-%d" % (names[c], a2, b2))
            cand_seqs.append(contigs[c][a2:b2].tobytes())
            f = fams[fam]
            truth.append(dict(family=fam, kind=f["kind"], consensus=f["consensus"], exact=exact, strand=items[i]["strand"],
                              n_copies=len(intact), contig=c, start=a, end=b, sub=f["sub"], indel=f["indel"]))
    wl = Workload(names, contigs, cand_names, cand_seqs, truth)
    wl.copies_truth = copies_truth
    wl.families = fams
    wl.repeat_bp = int(sum(b - a for (_, _, a, b, _) in copies_truth))
    return wl


def _write_copy(rng, g, pos, it, seq):
    """Write one copy at ``pos`` with its target-site duplication / insertion context; returns the end of what was written."""
    kind, tsd = it["kind"], it["tsd"]
    n = len(seq)
    if kind == "helitron":
        g[pos - 1] = ord("A")
        g[pos:pos + n] = seq
        g[pos + n] = ord("T")
        return pos + n + 1
    if kind == "tir":
        if tsd == 2:
            g[pos - 2:pos] = np.frombuffer(b"TA", dtype=np.uint8)
        elif tsd == 3:
            g[pos - 3:pos] = np.frombuffer(b"TAA" if rng.random() < 0.5 else b"TTA", dtype=np.uint8)
        elif tsd == 4:
            g[pos - 4:pos] = np.frombuffer(b"TTAA", dtype=np.uint8)
    g[pos:pos + n] = seq
    g[pos + n:pos + n + tsd] = g[pos - tsd:pos]
    return pos + n + tsd


def _place(rng, g, cur, items, i, placed, c):
    it = items[i]
    pos = cur + it["tsd"] + 1
    if it["nested"] is None:
        end = _write_copy(rng, g, pos, it, it["seq"])
        placed[i] = (c, pos, pos + len(it["seq"]))
        return end + 1
    guest = items[it["nested"]]
    seq = it["seq"]
    cut = int(rng.integers(len(seq) // 3, 2 * len(seq) // 3))
    # host left part, then the guest with its own TSD copied from the host's bases, then the host right part
    g[pos:pos + cut] = seq[:cut]
    if it["kind"] == "tir" and it["tsd"] in (2, 3, 4):
        _write_copy(rng, g, pos, dict(kind="tir", tsd=it["tsd"]), seq[:cut])          # host's 5' TSD motif
    gpos = pos + cut
    gend = _write_copy(rng, g, gpos, guest, guest["seq"])
    placed[it["nested"]] = (c, gpos, gpos + len(guest["seq"]))
    rest = seq[cut:]
    g[gend:gend + len(rest)] = rest
    e = gend + len(rest)
    g[e:e + it["tsd"]] = g[pos - it["tsd"]:pos] if it["tsd"] else g[e:e]
    placed[i] = (c, pos, e)          # interrupted host: listed, but not an intact copy
    return e + it["tsd"] + 1


# ---- the named configurations of BASELINE.json / SURVEY.md section 8(d) ------------------------------------------------
def config_c2(scale=1, seed=1):
    """C2: 100 Mbp (10 contigs x 10 Mbp), 10 000 planted TIR candidates; families of length U[200, 3000], 10-30 bp TIR, TSD in
    {2,3,4,8,9}, copies / family U[6, 100], 0-10 % per-base divergence, indels <= 2 %.  ``scale`` multiplies genome and
    candidates (weak scaling)."""
    return make_mixed_workload(100_000_000 * scale, 10, 10_000 * scale, mix=(("tir", 1.0),), copies=("uniform", 6, 100),
                               sub=(0.0, 0.10), indel=(0.0, 0.02), cands_per_family=20, seed=seed)


def config_c3(seed=2):
    """C3: 135 Mbp (5 contigs), 60 % TIR / 25 % LTR-like (direct repeats, 4-10 kbp) / 15 % Helitron, ~20 % repeat content."""
    return make_mixed_workload(135_000_000, 5, 10_000, mix=(("tir", 0.60), ("ltr", 0.25), ("helitron", 0.15)),
                               copies=("uniform", 6, 24), sub=(0.0, 0.10), indel=(0.0, 0.02), repeat_frac=0.20,
                               cands_per_family=15, seed=seed)


def config_c4(seed=3):
    """C4: 400 Mbp (12 contigs), 50 000 candidates."""
    return make_mixed_workload(400_000_000, 12, 50_000, mix=(("tir", 1.0),), copies=("uniform", 6, 100), sub=(0.0, 0.10),
                               indel=(0.0, 0.02), cands_per_family=20, seed=seed)


def config_c5(seed=4, n_candidates=100_000):
    """C5: 2.3 Gbp (10 contigs), ~85 % repeat content, nested insertions, family sizes Zipf s = 1.2 capped at 5 000."""
    return make_mixed_workload(2_300_000_000, 10, n_candidates, mix=(("tir", 0.5), ("ltr", 0.3), ("helitron", 0.2)),
                               copies=("zipf", 1.2, 5000), sub=(0.0, 0.10), indel=(0.0, 0.02), nested_frac=0.15,
                               repeat_frac=0.85, cands_per_family=40, seed=seed)
