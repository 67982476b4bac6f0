"""TEST INFRASTRUCTURE (oracle) -- ctypes front-end of oracle/liboracle.so (plain-C restatement).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module; the product (hite_b200/) never does.
"""
import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

INFO = {0: "", 1: "nb", 2: "fl1"}


class HboCall(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("is_te", "info", "copy_num", "align_start", "align_end",
                                       "homo_start", "homo_end", "cons_len", "trim_lead", "trim_trail")]


def build(force=False):
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "-s", "-B"])
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
    return _LIB


def _mat(rows):
    R = len(rows)
    Cn = len(rows[0]) if R else 0
    assert all(len(r) == Cn for r in rows), "ragged matrix"
    return "".join(rows).encode(), R, Cn


def remove_sparse_cols(rows):
    m, R, Cn = _mat(rows)
    out = C.create_string_buffer(max(1, R * Cn))
    cout = C.c_int(0)
    lib().hbo_remove_sparse_cols(m, R, Cn, out, C.byref(cout))
    k = cout.value
    raw = out.raw
    return [raw[r * k:(r + 1) * k].decode() for r in range(R)]


def judge_v5(cand, rows, plant=1, result_type="cons"):
    """-> dict mirroring judge_boundary_v5's (is_TE, info, cons_seq, copy_num) + boundaries."""
    m, R, Cn = _mat(rows)
    call = HboCall()
    cons = C.create_string_buffer(Cn + 2)
    cb = cand.encode()
    lib().hbo_judge_v5(cb, len(cb), m, R, Cn, int(plant), 1 if result_type == "cons" else 0,
                       C.byref(call), cons)
    return dict(is_te=bool(call.is_te), info=INFO[call.info], cons=cons.value.decode(),
                copy_num=call.copy_num, align_start=call.align_start, align_end=call.align_end,
                homo_start=call.homo_start, homo_end=call.homo_end)


def judge_v6(cand, rows):
    """-> dict mirroring judge_boundary_v6's (is_TE, info, cons_seq, copy_num) (Helitron)."""
    m, R, Cn = _mat(rows)
    call = HboCall()
    cons = C.create_string_buffer(Cn + 8)
    cb = cand.encode()
    lib().hbo_judge_v6(cb, len(cb), m, R, Cn, C.byref(call), cons)
    return dict(is_te=bool(call.is_te), info=INFO[call.info], cons=cons.value.decode(),
                copy_num=call.copy_num, align_start=call.align_start, align_end=call.align_end,
                homo_start=call.homo_start, homo_end=call.homo_end)


def judge_v9(cand, rows):
    """-> dict mirroring judge_boundary_v9's (is_TE, info, cons_seq, copy_num) (non-LTR); ``cons`` is the
    model sequence the reference returns even when is_TE is False."""
    m, R, Cn = _mat(rows)
    call = HboCall()
    cons = C.create_string_buffer(Cn + 8)
    cb = cand.encode()
    lib().hbo_judge_v9(cb, len(cb), m, R, Cn, C.byref(call), cons)
    return dict(is_te=bool(call.is_te), info=INFO[call.info], cons=cons.value.decode(),
                copy_num=call.copy_num, align_start=call.align_start, align_end=call.align_end,
                homo_start=call.homo_start, homo_end=call.homo_end, tsd_count=call.trim_lead)


def truncated_rows(copies, rows):
    """Truncated-member matrix of the reference's first pass (Util.py:8116-8124): the kept copies whose
    extended length exceeds 1000 bp, each reduced to its first 500 + last 500 bases (same columns)."""
    kept = [c for c in copies if c["keep"]]
    out = []
    for c, row in zip(kept, rows):
        if c["ext_len"] <= 1000:
            continue
        idx = [i for i, ch in enumerate(row) if ch != "-"]
        keep = set(idx[:500]) | set(idx[-500:])
        out.append("".join(ch if i in keep else "-" for i, ch in enumerate(row)))
    return out


def judge_with_trunc_gate(cand, copies, rows, judge=None, plant=1):
    """run_find_members_v8 (Util.py:10439-10449): truncated members first, then the full members."""
    judge = judge or (lambda c, r: judge_v5(c, r, plant))
    t = truncated_rows(copies, rows)
    if t:
        rt = judge(cand, remove_sparse_cols(t))
        if not rt["is_te"]:
            return rt
    return judge(cand, remove_sparse_cols(rows))


class HboHsp(C.Structure):
    _fields_ = [("subj", C.c_int32), ("qs", C.c_int32), ("qe", C.c_int32), ("ss", C.c_int32), ("se", C.c_int32)]


class HboInterval(C.Structure):
    _fields_ = [("chr", C.c_int32), ("pad", C.c_int32), ("start", C.c_int64), ("end", C.c_int64)]


def fmea(query, subjects, self_subj, hsps, skip_gap, max_len, shared=None):
    """get_longest_repeats_v4 for one query segment -> ['chr:start-end', ...] in the reference's order.  ``shared``: a handle
    from ``fmea_file_state()`` carrying chr_pos_candidates from one query record of a BLAST result file to the next."""
    chr_ids = shared["chr_ids"] if shared is not None else {}
    for name in [s[0] for s in subjects] + [query[0]]:
        chr_ids.setdefault(name, len(chr_ids))
    n = len(hsps)
    arr = (HboHsp * max(1, n))(*[HboHsp(*h) for h in hsps])
    sc = (C.c_int32 * max(1, len(subjects)))(*[chr_ids[s[0]] for s in subjects])
    so = (C.c_int64 * max(1, len(subjects)))(*[s[1] for s in subjects])
    cap = n + 8
    out = (HboInterval * cap)()
    L = lib()
    L.hbo_fmea_shared.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int64, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_int,
                                  C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    k = L.hbo_fmea_shared(arr, n, chr_ids[query[0]], query[1], self_subj, sc, so, len(subjects), skip_gap, max_len,
                          shared["set"] if shared is not None else None, out, cap)
    return ["%s:%d-%d" % (query[0], out[i].start, out[i].end) for i in range(k)]


class fmea_file_state:
    """chr_pos_candidates of one BLAST result file (module/Util.py:4155): shared by the query records of that file."""

    def __init__(self):
        L = lib()
        L.hbo_fmea_set_new.restype = C.c_void_p
        L.hbo_fmea_set_free.argtypes = [C.c_void_p]
        self._L = L
        self.d = {"set": C.c_void_p(L.hbo_fmea_set_new()), "chr_ids": {}}

    def __getitem__(self, k):
        return self.d[k]

    def close(self):
        if self.d["set"]:
            self._L.hbo_fmea_set_free(self.d["set"])
            self.d["set"] = None


def split_file_groups(seg_lens, base_threshold=1_000_000):
    """split_and_store_sequences (module/Util.py:4987-5012): records are packed into one file until it holds at least
    ``base_threshold`` bases -> group id per record."""
    groups, g, count = [], 0, 0
    for ln in seg_lens:
        groups.append(g)
        count += ln
        if count >= base_threshold:
            g += 1
            count = 0
    return groups


class HboCopyOut(C.Structure):
    _fields_ = [("subj", C.c_int32), ("start", C.c_int32), ("end", C.c_int32), ("qlen", C.c_int32), ("strand", C.c_int32)]


def query_copies(hsps, idents, query_len, subj_lens, query_coverage=0.95, subject_coverage=0.0, qthr=200, sthr=200,
                 max_copy_num=100):
    """get_query_copies (Util.py:6828) for one query -> [(subj, start, end, len, '+'|'-')]."""
    n = len(hsps)
    arr = (HboHsp * max(1, n))(*[HboHsp(*h) for h in hsps])
    idn = (C.c_double * max(1, n))(*idents)
    sl = (C.c_int32 * max(1, len(subj_lens)))(*subj_lens)
    cap = n + 8
    out = (HboCopyOut * cap)()
    L = lib()
    L.hbo_query_copies.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_double, C.c_double,
                                   C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    k = L.hbo_query_copies(arr, idn, n, query_len, sl, len(subj_lens), query_coverage, subject_coverage, qthr, sthr,
                           max_copy_num, out, cap)
    return [(out[i].subj, out[i].start, out[i].end, out[i].qlen, "-" if out[i].strand else "+") for i in range(k)]


def full_length_copies_v1(hsps, query_len, full_length_threshold=0.95):
    """get_full_length_copies_from_blastn_v1 (Util.py:5907) for one query -> [(subj, start0, end_excl, '+'|'-')], unique."""
    n = len(hsps)
    arr = (HboHsp * max(1, n))(*[HboHsp(*h) for h in hsps])
    cap = n + 8
    out = (HboCopyOut * cap)()
    L = lib()
    L.hbo_full_length_copies_v1.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_double, C.c_void_p, C.c_int]
    k = L.hbo_full_length_copies_v1(arr, n, query_len, full_length_threshold, out, cap)
    return [(out[i].subj, out[i].start, out[i].end, "-" if out[i].strand else "+") for i in range(k)]


def find_near_matches_ends(p, s, d):
    fs, fe, ls, le = C.c_int(-1), C.c_int(-1), C.c_int(-1), C.c_int(-1)
    n = lib().hbo_find_near_matches(p.encode(), len(p), s.encode(), len(s), d,
                                    C.byref(fs), C.byref(fe), C.byref(ls), C.byref(le))
    return n, (fs.value, fe.value), (ls.value, le.value)


def levenshtein(a, b):
    return lib().hbo_levenshtein(a.encode(), len(a), b.encode(), len(b))


# ---------------------------------------------------------------------------------------------
# HB-align restatement (oracle/hbo_align.c)
class HboCopy(C.Structure):
    _fields_ = [("rank", C.c_int32), ("rel", C.c_int32), ("score", C.c_int32), ("i_s", C.c_int32),
                ("i_e", C.c_int32), ("contig", C.c_int32), ("j_s", C.c_int64), ("j_e", C.c_int64),
                ("start1", C.c_int64), ("end1", C.c_int64), ("keep", C.c_int32), ("ext_len", C.c_int32)]


class HboTsdRec(C.Structure):
    _fields_ = [("k", C.c_int32), ("start", C.c_int32), ("end", C.c_int32), ("dist", C.c_int32), ("idx", C.c_int32),
                ("idx_lo", C.c_int32), ("idx_hi", C.c_int32), ("short_class", C.c_int32)]


def search_confident_tir_v4(seq, raw_start1, raw_end1, d, plant):
    """C restatement of search_confident_tir_v4 (+ get_short_tir_contigs class) -> list of dicts in output order."""
    L = lib()
    sb = seq if isinstance(seq, (bytes, bytearray)) else seq.encode()
    out = (HboTsdRec * 100)()
    L.hbo_search_confident_tir_v4.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
    n = L.hbo_search_confident_tir_v4(sb, len(sb), int(raw_start1), int(raw_end1), int(d), int(plant), out, 100)
    return [{f: getattr(out[i], f) for f, _ in HboTsdRec._fields_} for i in range(n)]


class HboQhsp(C.Structure):
    _fields_ = [(f, C.c_int32) for f in ("rank", "contig", "qs", "qe", "ss", "se", "n_match", "aln_len", "score")]


class HboPolyaRec(C.Structure):
    _fields_ = [(f, C.c_int32) for f in ("found", "direct", "end_5", "end_3", "tsd_k", "seq_start", "seq_end")]


def search_polyA_TSD(seq, flanking_len):
    """C restatement of search_polyA_TSD (Util.py:10915) -> (found_TSD, TSD_seq, non_ltr_seq, record dict)."""
    L = lib()
    sb = seq if isinstance(seq, (bytes, bytearray)) else seq.encode()
    r = HboPolyaRec()
    L.hbo_search_polyA_TSD.argtypes = [C.c_char_p, C.c_int, C.c_int, C.c_void_p]
    L.hbo_search_polyA_TSD(sb, len(sb), int(flanking_len), C.byref(r))
    rec = {f: getattr(r, f) for f, _ in HboPolyaRec._fields_}
    s = sb.decode()
    tsd = ""
    if rec["found"]:
        tsd = s[rec["end_3"]:rec["end_3"] + rec["tsd_k"]] if rec["direct"] == 1 else s[rec["end_3"] - rec["tsd_k"]:rec["end_3"]]
    body = s[rec["seq_start"]:rec["seq_end"]] if rec["direct"] else ""
    if rec["direct"] == 2:
        body = body[::-1].translate(str.maketrans("ACGT", "TGCA", "")).translate({ord(c): "N" for c in set(body) - set("ACGT")})
    return bool(rec["found"]), tsd, body, rec


class HboLocus(C.Structure):
    _fields_ = [("cand", C.c_uint32), ("rel", C.c_uint32), ("nseeds", C.c_uint32), ("rank", C.c_uint32),
                ("dmin", C.c_int64), ("dmax", C.c_int64), ("j0", C.c_int64)]


class Genome:
    """CPU genome + minimizer index (restatement of S1/S2)."""

    def __init__(self, contig_seqs, k=15, w=10):
        import numpy as np
        L = lib()
        L.hbo_genome_create.restype = C.c_void_p
        L.hbo_genome_create.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
        L.hbo_genome_free.argtypes = [C.c_void_p]
        L.hbo_genome_total.restype = C.c_uint64
        L.hbo_genome_total.argtypes = [C.c_void_p]
        L.hbo_minimizers.restype = C.c_uint64
        L.hbo_minimizers.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_uint64]
        L.hbo_sort_u64.argtypes = [C.c_void_p, C.c_uint64]
        seqs = [np.frombuffer(s, dtype=np.uint8) if isinstance(s, (bytes, bytearray)) else np.asarray(s, dtype=np.uint8)
                for s in contig_seqs]
        off = np.zeros(len(seqs) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(s) for s in seqs])
        flat = np.concatenate(seqs) if seqs else np.zeros(0, np.uint8)
        self._L = L
        self.k, self.w = k, w
        self.g = L.hbo_genome_create(flat.ctypes.data, off.ctypes.data, len(seqs))
        self.total = L.hbo_genome_total(self.g)
        n = L.hbo_minimizers(self.g, k, w, None, 0)
        self.mm_pos_order = np.zeros(n, dtype=np.uint64)
        L.hbo_minimizers(self.g, k, w, self.mm_pos_order.ctypes.data, n)
        self.index = self.mm_pos_order.copy()
        L.hbo_sort_u64(self.index.ctypes.data, n)

    def set_tie_rule(self, tie_rule, contig_names=None):
        """S9 tie rule: 1 (default) = the reference's ready_for_MSA.sh rule (larger member name wins a length tie; what
        copy_select_kernel does), 0 = earlier-ranked copy.  ``contig_names`` give the name order the rule needs (without
        them: the contig index, as in the library before hb_genome_contig_rank)."""
        import numpy as np
        self._L.hbo_genome_set_tie_rule.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        rank = None
        if contig_names is not None:
            order = sorted(range(len(contig_names)), key=lambda i: (contig_names[i] + ":").encode())
            rank = np.zeros(len(contig_names), dtype=np.int32)
            rank[order] = np.arange(len(contig_names), dtype=np.int32)
        self._L.hbo_genome_set_tie_rule(self.g, int(tie_rule), rank.ctypes.data if rank is not None else None)

    def close(self):
        if self.g:
            self._L.hbo_genome_free(self.g)
            self.g = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def candidate_msa(self, cand, occ_cap=512, max_loci=300, flank=50, max_rows=100):
        """-> (copies[list of dict], rows[list of str]) for one candidate (S3..S10)."""
        import numpy as np
        L = self._L
        cb = cand if isinstance(cand, (bytes, bytearray)) else cand.encode()
        n = len(cb)
        copies = (HboCopy * (max_loci + 1))()
        ncp = C.c_int(0)
        cap = (max_rows + 1) * (n + 2 * flank) * 4 + 1024
        mat = np.zeros(cap, dtype=np.uint8)
        R, Cn = C.c_int(0), C.c_int(0)
        L.hbo_candidate_msa.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_char_p, C.c_int,
                                        C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                        C.c_uint64, C.c_void_p, C.c_void_p]
        rc = L.hbo_candidate_msa(self.g, self.index.ctypes.data, len(self.index), self.k, self.w, cb, n, occ_cap,
                                 max_loci, flank, max_rows, copies, C.byref(ncp), mat.ctypes.data, cap,
                                 C.byref(R), C.byref(Cn))
        if rc != 0:
            raise RuntimeError("hbo_candidate_msa: matrix capacity exceeded")
        cps = [{f: getattr(copies[i], f) for f, _ in HboCopy._fields_} for i in range(ncp.value)]
        raw = mat[:R.value * Cn.value].tobytes()
        rows = [raw[r * Cn.value:(r + 1) * Cn.value].decode() for r in range(R.value)]
        return cps, rows

    def query_hsps(self, query, occ_cap=512, max_loci=300):
        """Library-vs-genome HSPs of one query -> list of dicts (rank, contig, qs, qe, ss, se, n_match, aln_len, score)."""
        L = self._L
        qb = query if isinstance(query, (bytes, bytearray)) else query.encode()
        out = (HboQhsp * (max_loci + 1))()
        L.hbo_query_hsps.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_char_p, C.c_int, C.c_int,
                                     C.c_int, C.c_void_p, C.c_int]
        n = L.hbo_query_hsps(self.g, self.index.ctypes.data, len(self.index), self.k, self.w, qb, len(qb), occ_cap, max_loci,
                             out, max_loci + 1)
        return [{f: getattr(out[i], f) for f, _ in HboQhsp._fields_} for i in range(n)]

    def run_batch(self, cand_seqs, threads, occ_cap=512, max_loci=300, flank=50, max_rows=100, plant=1):
        """Whole per-candidate path (copy search + MSA + clean + judge) on ``threads`` host threads.
        -> int32 array [n, 4] = (is_te, info, copy_num, cons_len)."""
        import numpy as np
        cb = [c if isinstance(c, (bytes, bytearray)) else c.encode() for c in cand_seqs]
        off = np.zeros(len(cb) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(c) for c in cb])
        flat = np.frombuffer(b"".join(cb) + b"\0", dtype=np.uint8)
        out = np.zeros((len(cb), 4), dtype=np.int32)
        L = self._L
        L.hbo_run_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_uint32, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p]
        L.hbo_run_batch(self.g, self.index.ctypes.data, len(self.index), self.k, self.w, flat.ctypes.data,
                        off.ctypes.data, len(cb), occ_cap, max_loci, flank, max_rows, plant, threads, out.ctypes.data)
        return out

    def coarse_boundary(self, contig_lens, seg_len=1_000_000, skip_gap=2000, max_len=30000, occ_cap=64):
        """HB-hsp self-join (hbo_hsp.c) + FMEA (hbo_fmea.c) -> (hsps [n,6] int32 in sorted-key order,
        segments [(contig, start)], intervals [(contig, start, end)] in query-segment order)."""
        import numpy as np
        L = self._L
        segs, goff = [], []
        p = 64
        for c, ln in enumerate(contig_lens):
            for o in range(0, ln, seg_len):
                segs.append((c, o)); goff.append(p + o)
            p += ln + 64
        goff = np.asarray(goff, dtype=np.uint64)
        L.hbo_self_hsps.restype = C.c_uint64
        L.hbo_self_hsps.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.c_int, C.c_void_p, C.c_void_p,
                                    C.c_int, C.c_int, C.c_void_p, C.c_uint64]
        slen = np.asarray([min(seg_len, contig_lens[c] - o) for c, o in segs], dtype=np.uint64)
        n = L.hbo_self_hsps(self.g, self.mm_pos_order.ctypes.data, self.index.ctypes.data, len(self.index), self.total, self.k,
                            goff.ctypes.data, slen.ctypes.data, len(segs), occ_cap, None, 0)
        recs = np.zeros((max(1, n), 6), dtype=np.int32)
        L.hbo_self_hsps(self.g, self.mm_pos_order.ctypes.data, self.index.ctypes.data, len(self.index), self.total, self.k,
                        goff.ctypes.data, slen.ctypes.data, len(segs), occ_cap, recs.ctypes.data, n)
        recs = recs[:n]
        intervals = []
        # the query records of one split file share chr_pos_candidates and the file's longest_repeats dict (:4152-4155)
        seg_lens = [min(seg_len, contig_lens[c] - o) for c, o in segs]
        groups = split_file_groups(seg_lens)
        state, cur_group, file_names = None, -1, set()
        for g, (qc, qo) in enumerate(segs):
            if groups[g] != cur_group:
                if state is not None:
                    state.close()
                state, cur_group, file_names = fmea_file_state(), groups[g], set()
            mine = recs[recs[:, 0] == g]
            if len(mine) == 0:
                continue
            hs = [(int(r[1]), int(r[2]), int(r[3]), int(r[4]), int(r[5])) for r in mine]
            names = fmea(("c%d" % qc, qo), [("c%d" % c, o) for c, o in segs], g, hs, skip_gap, max_len, shared=state)
            for nm in names:
                if nm in file_names:
                    continue
                file_names.add(nm)
                a, b = nm.split(":")[1].split("-")
                intervals.append((qc, int(a), int(b)))
        if state is not None:
            state.close()
        return recs, segs, intervals


def itr_search(seq, min_ident=0.7, min_len=7, max_w=500):
    """HB-itr (hbo_tsd.c::hbo_itr_search) -> dict(found, len5, len3, score, n_match, n_mismatch, n_gap)."""
    import numpy as np
    L = lib()
    L.hbo_itr_search.argtypes = [C.c_char_p, C.c_int, C.c_double, C.c_int, C.c_int, C.c_void_p]
    L.hbo_itr_search.restype = None
    b = seq.encode() if isinstance(seq, str) else bytes(seq)
    out = np.zeros(7, dtype=np.int32)
    L.hbo_itr_search(b, len(b), float(min_ident), int(min_len), int(max_w), out.ctypes.data)
    return dict(zip(("found", "len5", "len3", "score", "n_match", "n_mismatch", "n_gap"), (int(x) for x in out)))
