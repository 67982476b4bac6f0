"""Seeded synthetic MSA generator for the judge parity tests (test infrastructure).

Produces (candidate, rows) pairs shaped like what ``is_TE_from_align_file`` (reference
Util.py:10407) hands to ``remove_sparse_col_in_align_file`` + ``judge_boundary_v5/v6/v9``:
rows are *copies only* (flank + element + flank, aligned, '-' for gaps), candidate not included.
Uses ``random.Random`` only, so a (scenario, seed) pair is reproducible anywhere.
"""
import random

BASES = "ACGT"
COMP = {"A": "T", "C": "G", "G": "C", "T": "A"}


def revcomp(s):
    return "".join(COMP.get(b, "N") for b in reversed(s))


def rand_seq(rng, n):
    return "".join(rng.choice(BASES) for _ in range(n))


def make_tir_element(rng, length, tir_len, head=None):
    tir = rand_seq(rng, tir_len)
    if head:
        tir = head + tir[len(head):]
    body = rand_seq(rng, max(0, length - 2 * tir_len))
    return tir + body + revcomp(tir)


def make_tsd(rng, kind):
    if kind == 2:
        return "TA"
    if kind == 3:
        return rng.choice(["TAA", "TTA"])
    if kind == 4:
        return "TTAA"
    return rand_seq(rng, kind)


def mutate_row(rng, elem, sub, dele):
    out = []
    for b in elem:
        r = rng.random()
        if r < dele:
            out.append("-")
        elif r < dele + sub:
            out.append(rng.choice([x for x in BASES if x != b]))
        else:
            out.append(b)
    return "".join(out)


def synth_case(seed, scenario=None):
    """Returns dict(cand=str, rows=[str], plant=int, te_type='tir', scenario=str, truth=str)."""
    rng = random.Random(seed)
    scen = scenario or rng.choice(
        ["clean", "clean", "clean", "shift_out", "shift_in", "fp_homolog_flank", "fragments",
         "few_rows", "no_anchor", "many_rows", "ta_head", "n_bases", "short"])
    flank = 50
    nrows = rng.randint(6, 40)
    length = rng.randint(90, 420)
    sub = rng.choice([0.0, 0.02, 0.05, 0.1])
    dele = rng.choice([0.0, 0.005, 0.02])
    ins_cols = rng.randint(0, 6)
    tsd_kind = rng.choice([2, 3, 4, 5, 8, 9, 10])
    head = None
    if scen == "few_rows":
        nrows = rng.randint(1, 5)
    elif scen == "many_rows":
        nrows = rng.randint(101, 130)
        length = rng.randint(90, 160)
    elif scen == "ta_head":
        head = rng.choice(["TTAA", "TAA", "TTA", "TA", "AA", "A"])
    elif scen == "short":
        length = rng.randint(40, 85)
    tir_len = rng.randint(8, 20)
    elem = make_tir_element(rng, length, tir_len, head)
    if scen == "ta_head":
        tail = rng.choice(["TTAA", "TAA", "TTA", "TA", "TT", "T"])
        elem = elem[:-len(tail)] + tail
    shared_l = rand_seq(rng, flank)
    shared_r = rand_seq(rng, flank)
    rows = []
    for r in range(nrows):
        tsd = make_tsd(rng, tsd_kind)
        if rng.random() < 0.15:
            tsd_r = make_tsd(rng, tsd_kind) if tsd_kind > 4 else tsd      # broken TSD
        else:
            tsd_r = tsd
        lf = rand_seq(rng, flank - len(tsd)) + tsd
        rf = tsd_r + rand_seq(rng, flank - len(tsd_r))
        if scen == "fp_homolog_flank":
            lf = mutate_row(rng, shared_l, 0.03, 0.0).replace("-", "A")
            rf = mutate_row(rng, shared_r, 0.03, 0.0).replace("-", "A")
        body = mutate_row(rng, elem, sub, dele)
        if scen == "fragments" and r > 0 and rng.random() < 0.4:
            cut = rng.randint(15, len(body) // 2)
            if rng.random() < 0.5:
                body = "-" * cut + body[cut:]
                lf = "-" * flank
            else:
                body = body[:-cut] + "-" * cut
                rf = "-" * flank
        if scen == "n_bases" and rng.random() < 0.3:
            p = rng.randrange(len(body))
            body = body[:p] + "N" + body[p + 1:]
        rows.append(lf + body + rf)
    # insertion columns: a random row (or a few) carries a base, everyone else a gap
    for _ in range(ins_cols):
        p = rng.randint(1, len(rows[0]) - 1)
        who = set(rng.sample(range(nrows), k=min(nrows, rng.choice([1, 1, 2, nrows // 2 + 1]))))
        rows = [row[:p] + (rng.choice(BASES) if i in who else "-") + row[p:] for i, row in enumerate(rows)]
    cand = elem
    if scen == "shift_out":
        # candidate carries a few bp of copy-0's own flank on each side (coarse boundary too wide)
        a, b = rng.randint(0, 12), rng.randint(0, 12)
        r0 = rows[0].replace("-", "")
        s = r0.find(rows[0][flank:flank + 30].replace("-", "")[:12])
        core_start = len(rows[0][:flank].replace("-", ""))
        core_end = len(r0) - len(rows[0][-flank:].replace("-", ""))
        cand = r0[max(0, core_start - a):core_end + b]
    elif scen == "shift_in":
        a, b = rng.randint(0, 15), rng.randint(0, 15)
        cand = elem[a:len(elem) - b]
    elif scen == "no_anchor":
        cand = rand_seq(rng, length)
    return dict(cand=cand, rows=rows, plant=rng.choice([0, 1]), te_type="tir", scenario=scen, truth=elem)


def synth_case_helitron(seed, scenario=None):
    """Helitron-shaped cases for judge_boundary_v6 (reference Util.py:9821): element 5'-TC ... CTRR-3',
    inserted between an A and a T of the host (so the row reads ...A|TC......CTRR|T...)."""
    rng = random.Random(seed)
    scen = scenario or rng.choice(["clean", "clean", "clean", "shift_out", "shift_in", "fp_homolog_flank",
                                   "fragments", "few_rows", "no_anchor", "no_motif", "many_rows"])
    flank = 50
    nrows = rng.randint(3, 30)
    if scen == "few_rows":
        nrows = rng.randint(1, 5)
    elif scen == "many_rows":
        nrows = rng.randint(101, 120)
    length = rng.randint(100, 380)
    sub = rng.choice([0.0, 0.02, 0.05])
    dele = rng.choice([0.0, 0.005, 0.02])
    tail = rng.choice(["CTAG", "CTAA", "CTGG", "CTGA"]) if scen != "no_motif" else rand_seq(rng, 4)
    head = "TC" if scen != "no_motif" or rng.random() < 0.5 else rand_seq(rng, 2)
    elem = head + rand_seq(rng, length - 6) + tail
    shared_l, shared_r = rand_seq(rng, flank), rand_seq(rng, flank)
    rows = []
    for r in range(nrows):
        lf = rand_seq(rng, flank - 1) + "A"
        rf = "T" + rand_seq(rng, flank - 1)
        if scen == "fp_homolog_flank":
            lf = mutate_row(rng, shared_l, 0.03, 0.0).replace("-", "A")
            rf = mutate_row(rng, shared_r, 0.03, 0.0).replace("-", "A")
        body = mutate_row(rng, elem, sub, dele)
        if scen == "fragments" and r > 0 and rng.random() < 0.4:
            cut = rng.randint(10, len(body) // 2)
            if rng.random() < 0.5:
                body = "-" * cut + body[cut:]
                lf = "-" * flank
            else:
                body = body[:-cut] + "-" * cut
                rf = "-" * flank
        rows.append(lf + body + rf)
    for _ in range(rng.randint(0, 4)):
        p = rng.randint(1, len(rows[0]) - 1)
        who = set(rng.sample(range(nrows), k=min(nrows, rng.choice([1, 1, 2, nrows // 2 + 1]))))
        rows = [row[:p] + (rng.choice(BASES) if i in who else "-") + row[p:] for i, row in enumerate(rows)]
    cand = elem
    if scen == "shift_out":
        r0 = rows[0].replace("-", "")
        core_start = len(rows[0][:flank].replace("-", ""))
        core_end = len(r0) - len(rows[0][-flank:].replace("-", ""))
        cand = r0[max(0, core_start - rng.randint(0, 10)):core_end + rng.randint(0, 10)]
    elif scen == "shift_in":
        cand = elem[rng.randint(0, 12):len(elem) - rng.randint(0, 12)]
    elif scen == "no_anchor":
        cand = rand_seq(rng, length)
    return dict(cand=cand, rows=rows, plant=1, te_type="helitron", scenario=scen, truth=elem)


def synth_case_nonltr(seed, scenario=None, nrows=None):
    """non-LTR-shaped cases for judge_boundary_v9 (reference Util.py:9483): element with a polyA (or
    short tandem) 3' tail, flanked by 8-20 bp target-site duplications; some copies 5'-truncated."""
    rng = random.Random(seed)
    scen = scenario or rng.choice(["clean", "clean", "clean", "tandem_tail", "trunc5", "no_tsd", "few_rows",
                                   "no_anchor", "fp_homolog_flank", "shift_out", "short"])
    flank = 50
    nrows_in = nrows
    nrows = rng.randint(4, 30)
    if scen == "few_rows":
        nrows = rng.randint(1, 6)
    if nrows_in is not None:                      # same random stream as the golden cases, only more rows
        nrows = nrows_in
    length = rng.randint(120, 400) if scen != "short" else rng.randint(60, 95)
    sub = rng.choice([0.0, 0.02, 0.05])
    dele = rng.choice([0.0, 0.005, 0.01])
    if scen == "tandem_tail":
        unit = rand_seq(rng, rng.randint(2, 5))
        tail = unit * rng.randint(4, 7)
    else:
        tail = "A" * rng.randint(7, 20)
    body = rand_seq(rng, length - len(tail))
    elem = body + tail
    shared_l, shared_r = rand_seq(rng, flank), rand_seq(rng, flank)
    rows = []
    for r in range(nrows):
        k = rng.randint(8, 18)
        tsd = rand_seq(rng, k)
        tsd_r = tsd if scen != "no_tsd" else rand_seq(rng, k)
        if scen != "no_tsd" and rng.random() < 0.25:
            p = rng.randrange(k)
            tsd_r = tsd[:p] + rng.choice([x for x in BASES if x != tsd[p]]) + tsd[p + 1:]
        lf = rand_seq(rng, flank - k) + tsd
        rf = tsd_r + rand_seq(rng, flank - k)
        if scen == "fp_homolog_flank":
            lf = mutate_row(rng, shared_l, 0.03, 0.0).replace("-", "A")
            rf = mutate_row(rng, shared_r, 0.03, 0.0).replace("-", "A")
        # keep the tail intact (polyA detection is exact), mutate the body
        b = mutate_row(rng, body, sub, dele) + tail
        if scen == "trunc5" and r > 0 and rng.random() < 0.5:
            cut = rng.randint(10, len(body) // 2)
            b = "-" * cut + b[cut:]
            lf = "-" * (flank - k) + tsd if rng.random() < 0.5 else "-" * flank
        rows.append(lf + b + rf)
    for _ in range(rng.randint(0, 4)):
        p = rng.randint(1, len(rows[0]) - 1)
        who = set(rng.sample(range(nrows), k=min(nrows, rng.choice([1, 1, 2, nrows // 2 + 1]))))
        rows = [row[:p] + (rng.choice(BASES) if i in who else "-") + row[p:] for i, row in enumerate(rows)]
    cand = elem
    if scen == "shift_out":
        r0 = rows[0].replace("-", "")
        core_start = len(rows[0][:flank].replace("-", ""))
        core_end = len(r0) - len(rows[0][-flank:].replace("-", ""))
        cand = r0[max(0, core_start - rng.randint(0, 10)):core_end + rng.randint(0, 8)]
    elif scen == "no_anchor":
        cand = rand_seq(rng, length)
    return dict(cand=cand, rows=rows, plant=1, te_type="non_ltr", scenario=scen, truth=elem)


def synth_hsps(seed):
    """Synthetic BLAST-outfmt-6 style HSP list of ONE query segment against several subject segments for the
    FMEA chaining tests (reference get_longest_repeats_v4, module/Util.py:4122).  Returns dict(query=(chr,off),
    subjects=[(chr,off)], self_subj=index|-1, hsps=[(subj,qs,qe,ss,se)], skip_gap, max_len)."""
    rng = random.Random(seed)
    chrs = ["chr1", "chr2", "chr3"]
    query = (rng.choice(chrs), rng.choice([0, 1_000_000, 2_000_000]))
    subjects = []
    while len(subjects) < rng.randint(2, 5):
        s = (rng.choice(chrs), rng.choice([0, 1_000_000, 2_000_000, 3_000_000]))
        if s not in subjects:
            subjects.append(s)
    self_subj = -1
    if rng.random() < 0.7:
        if query in subjects:
            self_subj = subjects.index(query)
        else:
            subjects.append(query)
            self_subj = len(subjects) - 1
    skip_gap = rng.choice([200, 2000, 4000])
    hsps = []
    for fam in range(rng.randint(2, 7)):
        flen = rng.randint(100, 6000)
        qcopies = [rng.randint(1, 990_000 - flen) for _ in range(rng.randint(1, 3))]
        for si in range(len(subjects)):
            for _ in range(rng.randint(0, 3)):
                spos = rng.randint(1, 990_000 - flen)
                strand = rng.random() < 0.4
                for qpos in qcopies:
                    # one alignment, possibly broken into pieces separated by small/large gaps
                    pieces = rng.randint(1, 4)
                    cuts = sorted(rng.sample(range(20, flen - 20), pieces - 1)) if pieces > 1 and flen > 60 else []
                    bounds = [0] + cuts + [flen]
                    qshift = sshift = 0
                    for a, b in zip(bounds[:-1], bounds[1:]):
                        if b - a < 12:
                            continue
                        g1 = rng.choice([0, 0, 5, 50, 300, skip_gap + 10])
                        g2 = rng.choice([0, 0, 5, 50, 300, skip_gap + 10])
                        qshift += g1; sshift += g2
                        qs, qe = qpos + a + qshift, qpos + b - 1 + qshift
                        if not strand:
                            ss, se = spos + a + sshift, spos + b - 1 + sshift
                        else:
                            ss, se = spos + flen - a + 6000 - sshift, spos + flen - b + 1 + 6000 - sshift
                        if qe > 1_000_000 or max(ss, se) > 1_000_000 or min(ss, se) < 1:
                            continue
                        hsps.append((si, qs, qe, ss, se))
                        if rng.random() < 0.1:
                            hsps.append((si, qs, qe, ss, se))            # exact duplicate line
    if self_subj >= 0:
        for _ in range(rng.randint(1, 3)):                               # self hits that must be skipped
            a = rng.randint(1, 900_000); b = a + rng.randint(100, 5000)
            hsps.append((self_subj, a, b, a, b))
    rng.shuffle(hsps)
    return dict(query=query, subjects=subjects, self_subj=self_subj, hsps=hsps, skip_gap=skip_gap,
                max_len=rng.choice([30000, 3000]))


def synth_copy_hsps(seed):
    """HSPs of ONE query (a TE of length qlen) against several subject sequences for get_query_copies
    (reference module/Util.py:6828): full-length hits broken into pieces, partial hits, duplicates."""
    rng = random.Random(seed)
    qlen = rng.randint(300, 4000)
    n_subj = rng.randint(1, 4)
    subj_lens = [rng.randint(20_000, 200_000) for _ in range(n_subj)]
    hsps, idents = [], []
    for si in range(n_subj):
        for _ in range(rng.randint(1, 8)):
            spos = rng.randint(1, subj_lens[si] - qlen - 600)
            strand = rng.random() < 0.4
            lo = 1 if rng.random() < 0.7 else rng.randint(2, qlen // 3)
            hi = qlen if rng.random() < 0.7 else rng.randint(2 * qlen // 3, qlen - 1)
            pieces = rng.randint(1, 4)
            cuts = sorted(rng.sample(range(lo + 10, hi - 10), pieces - 1)) if pieces > 1 and hi - lo > 60 else []
            bounds = [lo] + cuts + [hi + 1]
            qshift = sshift = 0
            for a, b in zip(bounds[:-1], bounds[1:]):
                if b - a < 12:
                    continue
                qshift += rng.choice([0, 0, 3, 40, 150, 260]); sshift += rng.choice([0, 0, 3, 40, 150, 260])
                qs, qe = a + qshift, b - 1 + qshift
                if not strand:
                    ss, se = spos + a + sshift, spos + b - 1 + sshift
                else:
                    ss, se = spos + qlen + 600 - a - sshift, spos + qlen + 600 - b + 1 - sshift
                if min(ss, se) < 1 or max(ss, se) > subj_lens[si]:
                    continue
                hsps.append((si, qs, qe, ss, se)); idents.append(round(rng.uniform(80, 100), 3))
                if rng.random() < 0.08:
                    hsps.append((si, qs, qe, ss, se)); idents.append(idents[-1])
    order = list(range(len(hsps)))
    rng.shuffle(order)
    return dict(qlen=qlen, subj_lens=subj_lens, hsps=[hsps[i] for i in order], idents=[idents[i] for i in order],
                query_coverage=rng.choice([0.95, 0.8]), subject_coverage=rng.choice([0, 0, 0.001]),
                max_copy_num=rng.choice([100, 3]))
