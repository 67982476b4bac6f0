"""Host mirror of the TSD / short-TIR pre-filter that runs right before the boundary-adjustment path in HiTE's TIR
module (reference module/judge_TIR_transposons.py:71-77 -> Util.py:6630 ``multi_process_tsd_v1`` ->
``search_confident_tir_batch`` :7358): same function names, arguments and return shapes; the k-mer matching runs in
``hb_tsd_candidates`` (hite_b200/csrc/hb_tsd.cu).  The two ``itrsearch`` calls of the reference loop (a closed-source
binary, Util.py:216-224) are NOT part of this module: it returns exactly what the reference hands to them.

Record names: the reference numbers the variants of a candidate by their position in a distance-sorted Python *set*,
so among equal distances its ``-C_<i>-`` is hash-seed dependent; here ties are ordered by (start, end, k).
"""
from . import _lib

_CTX = None


def _ctx():
    global _CTX
    if _CTX is None:
        _CTX = _lib.Context(0)
    return _CTX


def _records_to_dict(query_name, seq, recs, n):
    out = {}
    for r in recs[:n]:
        s, e, k = int(r["start"]), int(r["end"]), int(r["k"])
        name = "%s-C_%d-tsd_%s-distance_%d" % (query_name, int(r["idx"]), seq[e + 1:e + 1 + k], int(r["dist"]))   # :7836
        out[name] = seq[s:e + 1]
    return out


def search_confident_tir_v4(orig_seq, raw_tir_start, raw_tir_end, tsd_search_distance, query_name, plant, ctx=None):
    """reference module/Util.py:7734 -- {new_query_name: tir_seq}, at most 100, closest to the raw boundaries first."""
    recs, cnt = (ctx or _ctx()).tsd_candidates([orig_seq], [raw_tir_start], [raw_tir_end], tsd_search_distance, plant)
    return _records_to_dict(query_name, orig_seq, recs[0], int(cnt[0]))


def get_short_tir_contigs(cur_itr_contigs, plant, classes=None):
    """reference module/Util.py:7297.  ``classes`` (name -> short_class from ``search_confident_tir_batch``) avoids
    re-deriving what the kernel already computed; without it the rule is evaluated here on the strings."""
    if classes is not None:
        return {n: s for n, s in cur_itr_contigs.items() if classes.get(n, 0)}
    comp = {"A": "T", "T": "A", "C": "G", "G": "C"}
    rc = lambda x: "".join(comp.get(b, "N") for b in reversed(x))
    out = {}
    for name, tir in cur_itr_contigs.items():
        parts = name.split("-tsd_")
        tsd_len = len(parts[1].split("-")[0]) if len(parts) > 1 else 0
        if tir[:5] == rc(tir[-5:]):
            if (tsd_len == 8 and len(tir) < 4000) or 9 <= tsd_len <= 11 or \
                    (plant == 1 and tsd_len == 3 and tir[:5] in ("CACTA", "CACTG")):
                out[name] = tir
        elif tir[:3] == rc(tir[-3:]) and tir[:3] == "CCC":
            out[name] = tir
    return out


def search_confident_tir_batch(cur_contigs, flanking_len, plant, ctx=None):
    """The per-partition loop of the reference's ``search_confident_tir_batch`` (module/Util.py:7358-7378) up to the
    first ``itrsearch`` call: (all_copies_itr_contigs, short_itr_contigs), insertion-ordered like the reference."""
    names, seqs = [], []
    for query_name, seq in cur_contigs.items():
        if "NNNNNNNNNN" in seq:                                  # :7368
            continue
        names.append(query_name)
        seqs.append(seq)
    starts = [flanking_len + 1] * len(seqs)                      # :7371
    ends = [len(s) - flanking_len for s in seqs]                 # :7372
    keep = [i for i in range(len(seqs)) if 1 <= ends[i] <= len(seqs[i]) and starts[i] <= len(seqs[i])]
    recs, cnt = (ctx or _ctx()).tsd_candidates([seqs[i] for i in keep], [starts[i] for i in keep], [ends[i] for i in keep],
                                              flanking_len, plant)
    all_contigs, classes = {}, {}
    for row, i in enumerate(keep):
        d = _records_to_dict(names[i], seqs[i], recs[row], int(cnt[row]))
        for (name, _), r in zip(d.items(), recs[row][:int(cnt[row])]):
            classes[name] = int(r["short_class"])
        all_contigs.update(d)
    return all_contigs, get_short_tir_contigs(all_contigs, plant, classes)


def filter_dup_itr_v3(cur_copies_out_contigs, TIR_len_dict):
    """reference module/Util.py:2791-2812: of one candidate's TSD-bounded variants keep the one closest to the raw boundaries
    (first among equals, in the dict's order), renamed ``<query>-tir_<len>-tsd_<tsd>``; nothing if it is 30 000 bp or longer.
    Host-side selection only (one string split per variant)."""
    best_name, best_dist = "", 100000
    for name in cur_copies_out_contigs:
        dist = int(name.split("-distance_")[1])
        if dist < best_dist:
            best_name, best_dist = name, dist
    query, rest = best_name.split("-C_")[0], best_name.split("-C_")[1]
    tsd = rest.split("-tsd_")[1].split("-")[0]
    seq = cur_copies_out_contigs[best_name]
    if len(seq) >= 30_000:
        return {}
    return {"%s-tir_%d-tsd_%s" % (query, TIR_len_dict.get(best_name, 0), tsd): seq}
