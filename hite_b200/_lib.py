"""ctypes binding of libhite_b200.so (include/hite_b200.h).  Fails loudly when the extension is
missing or no GPU is present -- there is no CPU fallback on the product path."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

TE_TYPES = {"tir": 0, "helitron": 1, "non_ltr": 2}
INFO = {0: "", 1: "nb", 2: "fl1"}


class HiteB200Error(RuntimeError):
    pass


class HbCall(C.Structure):
    _fields_ = [("is_te", C.c_int32), ("info", C.c_int32), ("copy_num", C.c_int32),
                ("align_start", C.c_int32), ("align_end", C.c_int32),
                ("homo_start", C.c_int32), ("homo_end", C.c_int32), ("cons_len", C.c_int32),
                ("trim_lead", C.c_int32), ("trim_trail", C.c_int32), ("clean_cols", C.c_int32),
                ("flags", C.c_int32), ("cons_off", C.c_int64)]


CALL_DTYPE = np.dtype([("is_te", "<i4"), ("info", "<i4"), ("copy_num", "<i4"), ("align_start", "<i4"),
                       ("align_end", "<i4"), ("homo_start", "<i4"), ("homo_end", "<i4"), ("cons_len", "<i4"),
                       ("trim_lead", "<i4"), ("trim_trail", "<i4"), ("clean_cols", "<i4"), ("flags", "<i4"),
                       ("cons_off", "<i8")])
assert CALL_DTYPE.itemsize == C.sizeof(HbCall)


class HbParams(C.Structure):
    _fields_ = [("occ_cap", C.c_int32), ("max_loci", C.c_int32), ("flank", C.c_int32), ("max_rows", C.c_int32),
                ("te_type", C.c_int32), ("plant", C.c_int32), ("result_type", C.c_int32), ("trunc_gate", C.c_int32),
                ("tb_budget_bytes", C.c_uint64), ("overlap", C.c_int32), ("flags", C.c_int32)]


COPY_DTYPE = np.dtype([("cand", "<u4"), ("rank", "<i4"), ("rel", "<i4"), ("nseeds", "<i4"), ("score", "<i4"),
                       ("i_s", "<i4"), ("i_e", "<i4"), ("keep", "<i4"), ("row", "<i4"), ("contig", "<i4"),
                       ("j0", "<i8"), ("j_s", "<i8"), ("j_e", "<i8"), ("start1", "<i8"), ("end1", "<i8"),
                       ("ext_len", "<i4"), ("trow", "<i4")])
assert COPY_DTYPE.itemsize == 88


QHSP_DTYPE = np.dtype([("query", "<u4"), ("contig", "<i4"), ("qs", "<i4"), ("qe", "<i4"), ("ss", "<i4"), ("se", "<i4"),
                       ("n_match", "<i4"), ("aln_len", "<i4"), ("score", "<i4"), ("rank", "<i4")])
assert QHSP_DTYPE.itemsize == 40
POLYA_DTYPE = np.dtype([(f, "<i4") for f in ("found", "direct", "end_5", "end_3", "tsd_k", "seq_start", "seq_end")])
ITR_DTYPE = np.dtype([(f, "<i4") for f in ("found", "len5", "len3", "score", "n_match", "n_mismatch", "n_gap")])
TSD_DTYPE = np.dtype([("k", "<i4"), ("start", "<i4"), ("end", "<i4"), ("dist", "<i4"), ("idx", "<i4"), ("short_class", "<i4")])
HSP_DTYPE = np.dtype([("subj", "<i4"), ("qs", "<i4"), ("qe", "<i4"), ("ss", "<i4"), ("se", "<i4")])
SEG_DTYPE = np.dtype([("chr", "<i4"), ("pad", "<i4"), ("start", "<i8")])
INTERVAL_DTYPE = np.dtype([("chr", "<i4"), ("qseg", "<i4"), ("start", "<i8"), ("end", "<i8")])


def fmea_chain(hsps, q_off, segs, skip_gap, max_len, threads=1):
    """FMEA interval chaining (reference get_longest_repeats_v4 + process_all_seqs) on HSP arrays:
    ``hsps`` HSP_DTYPE grouped by query segment, ``q_off`` uint64 [n_seg+1], ``segs`` SEG_DTYPE."""
    lib = load()
    hsps = np.ascontiguousarray(hsps, dtype=HSP_DTYPE)
    q_off = np.ascontiguousarray(q_off, dtype=np.uint64)
    segs = np.ascontiguousarray(segs, dtype=SEG_DTYPE)
    n = C.c_uint64(0)
    rc = lib.hb_fmea_chain(_ptr(hsps), _ptr(q_off), len(segs), _ptr(segs), int(skip_gap), int(max_len), int(threads), None, 0,
                           C.byref(n))
    if rc != 0:
        raise HiteB200Error("hb_fmea_chain failed (%d)" % rc)
    out = np.zeros(max(1, n.value), dtype=INTERVAL_DTYPE)
    rc = lib.hb_fmea_chain(_ptr(hsps), _ptr(q_off), len(segs), _ptr(segs), int(skip_gap), int(max_len), int(threads), _ptr(out),
                           len(out), C.byref(n))
    if rc != 0:
        raise HiteB200Error("hb_fmea_chain failed (%d)" % rc)
    return out[:n.value]


def fmea_chain_files(hsps, q_off, segs, file_of_seg, skip_gap, max_len, threads=1):
    """``hb_fmea_chain_files``: FMEA with the reference's BLAST result files explicit (file id per query segment)."""
    lib = load()
    hsps = np.ascontiguousarray(hsps, dtype=HSP_DTYPE)
    q_off = np.ascontiguousarray(q_off, dtype=np.uint64)
    segs = np.ascontiguousarray(segs, dtype=SEG_DTYPE)
    fos = np.ascontiguousarray(file_of_seg, dtype=np.uint32)
    out = np.zeros(len(hsps) + 1, dtype=INTERVAL_DTYPE)
    n = C.c_uint64(0)
    rc = lib.hb_fmea_chain_files(_ptr(hsps), _ptr(q_off), len(segs), _ptr(segs), _ptr(fos), int(skip_gap), int(max_len), int(threads),
                                 _ptr(out), len(out), C.byref(n))
    if rc != 0:
        raise HiteB200Error("hb_fmea_chain_files failed (%d)" % rc)
    return out[:n.value]


def debug_fmea_blocks(hsps, q_off, segs, file_of_seg, skip_gap, max_len, threads, n_blocks):
    """Test hook (``hb_debug_fmea_blocks``): the chaining fed in ``n_blocks`` submissions of whole files."""
    lib = load()
    hsps = np.ascontiguousarray(hsps, dtype=HSP_DTYPE)
    q_off = np.ascontiguousarray(q_off, dtype=np.uint64)
    segs = np.ascontiguousarray(segs, dtype=SEG_DTYPE)
    fos = np.ascontiguousarray(file_of_seg, dtype=np.uint32)
    out = np.zeros(len(hsps) + 1, dtype=INTERVAL_DTYPE)
    n = C.c_uint64(0)
    rc = lib.hb_debug_fmea_blocks(_ptr(hsps), _ptr(q_off), len(segs), _ptr(segs), _ptr(fos), int(skip_gap), int(max_len), int(threads),
                                  int(n_blocks), _ptr(out), len(out), C.byref(n))
    if rc != 0:
        raise HiteB200Error("hb_debug_fmea_blocks failed (%d)" % rc)
    return out[:n.value]


QCOPY_DTYPE = np.dtype([("subj", "<i4"), ("start", "<i4"), ("end", "<i4"), ("qlen", "<i4"), ("strand", "<i4")])


def query_copies(hsps, idents, query_len, subj_lens, query_coverage=0.95, subject_coverage=0.0, qthr=200, sthr=200,
                 max_copy_num=100):
    """get_query_copies (reference module/Util.py:6828) for one query: ``hsps`` HSP_DTYPE, ``idents`` float64."""
    lib = load()
    hsps = np.ascontiguousarray(hsps, dtype=HSP_DTYPE)
    idents = np.ascontiguousarray(idents, dtype=np.float64)
    sl = np.ascontiguousarray(subj_lens, dtype=np.int32)
    out = np.zeros(len(hsps) + 8, dtype=QCOPY_DTYPE)
    n = C.c_uint32(0)
    rc = lib.hb_query_copies(_ptr(hsps), _ptr(idents), len(hsps), int(query_len), _ptr(sl), float(query_coverage),
                             float(subject_coverage), int(qthr), int(sthr), int(max_copy_num), _ptr(out), len(out), C.byref(n))
    if rc != 0:
        raise HiteB200Error("hb_query_copies failed (%d)" % rc)
    return out[:n.value]


def full_length_copies(hsps, query_len, full_length_threshold=0.95):
    """``hb_full_length_copies``: get_full_length_copies_from_blastn_v1 (reference module/Util.py:5907) for one query.
    ``hsps`` is an HSP_DTYPE array; returns QCOPY_DTYPE records (subj, start 0-based, end exclusive, qlen, strand)."""
    lib = load()
    hs = np.ascontiguousarray(hsps, dtype=HSP_DTYPE)
    out = np.zeros(len(hs) + 1, dtype=QCOPY_DTYPE)
    n = C.c_uint32(0)
    rc = lib.hb_full_length_copies(_ptr(hs), len(hs), int(query_len), float(full_length_threshold), _ptr(out), len(out), C.byref(n))
    if rc != 0:
        raise HiteB200Error("hb_full_length_copies failed (%d)" % rc)
    return out[:n.value]


def _chain_batch(call, hsps, q_off):
    out = np.zeros(len(hsps) + 1, dtype=QCOPY_DTYPE)
    oq = np.zeros(len(hsps) + 1, dtype=np.uint32)
    n = C.c_uint64(0)
    rc = call(out, oq, n)
    if rc != 0:
        raise HiteB200Error("batched chaining failed (%d)" % rc)
    return out[:n.value], oq[:n.value]


def full_length_copies_batch(hsps, q_off, query_lens, full_length_threshold=0.95, threads=8):
    """``hb_full_length_copies_batch``: query g owns hsps[q_off[g]:q_off[g+1]] -> (QCOPY_DTYPE records, query of each)."""
    lib = load()
    hs = np.ascontiguousarray(hsps, dtype=HSP_DTYPE)
    qo = np.ascontiguousarray(q_off, dtype=np.uint64)
    ql = np.ascontiguousarray(query_lens, dtype=np.int32)
    return _chain_batch(lambda out, oq, n: lib.hb_full_length_copies_batch(
        _ptr(hs), _ptr(qo), len(ql), _ptr(ql), float(full_length_threshold), int(threads), _ptr(out), _ptr(oq), len(out), C.byref(n)),
        hs, qo)


def query_copies_batch(hsps, idents, q_off, query_lens, query_coverage=0.95, qthr=200, sthr=200, max_copy_num=100, threads=8):
    """``hb_query_copies_batch`` (subject_coverage = 0) -> (QCOPY_DTYPE records, query of each)."""
    lib = load()
    hs = np.ascontiguousarray(hsps, dtype=HSP_DTYPE)
    idn = np.ascontiguousarray(idents, dtype=np.float64)
    qo = np.ascontiguousarray(q_off, dtype=np.uint64)
    ql = np.ascontiguousarray(query_lens, dtype=np.int32)
    return _chain_batch(lambda out, oq, n: lib.hb_query_copies_batch(
        _ptr(hs), _ptr(idn), _ptr(qo), len(ql), _ptr(ql), float(query_coverage), int(qthr), int(sthr), int(max_copy_num),
        int(threads), _ptr(out), _ptr(oq), len(out), C.byref(n)), hs, qo)


class HbTiming(C.Structure):
    _fields_ = [("h2d_ms", C.c_float), ("kernel_ms", C.c_float), ("d2h_ms", C.c_float),
                ("k_ms", C.c_float * 8), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
                ("launches", C.c_uint32), ("pad", C.c_uint32)]


def lib_path():
    return os.path.join(_HERE, "libhite_b200.so")


def load():
    """Load the shared library; raises HiteB200Error when it has not been built."""
    global _LIB
    if _LIB is not None:
        return _LIB
    path = lib_path()
    if not os.path.exists(path):
        raise HiteB200Error(
            "libhite_b200.so is not built (%s). Run `python -c 'import __graft_entry__ as g; g.build()'` "
            "or `make -C hite_b200/csrc`. There is no CPU fallback." % path)
    lib = C.CDLL(path)
    vp, u8p, u64p, i32p = C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_uint64), C.POINTER(C.c_int32)
    lib.hb_version.restype = C.c_int
    lib.hb_device_count.restype = C.c_int
    lib.hb_create.argtypes = [C.c_int, C.POINTER(vp)]
    lib.hb_destroy.argtypes = [vp]
    lib.hb_destroy.restype = None
    lib.hb_last_error.argtypes = [vp]
    lib.hb_last_error.restype = C.c_char_p
    lib.hb_get_timing.argtypes = [vp, C.POINTER(HbTiming)]
    lib.hb_vote_matrices.argtypes = [vp, C.c_uint32, vp, vp, vp, vp, vp, vp, C.c_int, C.c_int, C.c_int, vp, vp,
                                     C.c_uint64]
    lib.hb_genome_load.argtypes = [vp, vp, vp, C.c_uint32]
    lib.hb_genome_contig_rank.argtypes = [vp, vp, C.c_uint32]
    lib.hb_index_build.argtypes = [vp, C.c_int, C.c_int]
    lib.hb_genome_info.argtypes = [vp, u64p, u64p, vp, C.c_uint32]
    lib.hb_adjust.argtypes = [vp, C.c_uint32, vp, vp, C.POINTER(HbParams), vp, u64p]
    lib.hb_last_cons.argtypes = [vp, vp, C.c_uint64]
    lib.hb_last_cons_device.argtypes = [vp, C.POINTER(vp), u64p]
    lib.hb_last_copies.argtypes = [vp, vp, C.c_uint64, u64p]
    lib.hb_last_pair_counts.argtypes = [vp, vp, C.c_uint32]
    lib.hb_debug_index.argtypes = [vp, vp, C.c_uint64]
    lib.hb_debug_scan_total.argtypes = [vp, C.c_uint64, C.c_uint32, u64p, C.POINTER(C.c_uint32)]
    lib.hb_last_msa.argtypes = [vp, C.c_uint32, vp, C.c_uint64, i32p, i32p]
    lib.hb_last_msa_trunc.argtypes = [vp, C.c_uint32, vp, C.c_uint64, i32p, i32p]
    lib.hb_last_cons_view.argtypes = [vp, C.POINTER(C.c_void_p), u64p]
    lib.hb_host_register.argtypes = [vp, C.c_uint64]
    lib.hb_host_unregister.argtypes = [vp]
    lib.hb_fmea_chain.argtypes = [vp, vp, C.c_uint32, vp, C.c_int, C.c_int, C.c_int, vp, C.c_uint64, u64p]
    lib.hb_query_copies.argtypes = [vp, vp, C.c_uint32, C.c_int, vp, C.c_double, C.c_double, C.c_int, C.c_int, C.c_int, vp,
                                    C.c_uint32, C.POINTER(C.c_uint32)]
    lib.hb_coarse_boundary.argtypes = [vp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, vp, C.c_uint64, u64p]
    lib.hb_coarse_boundary_part.argtypes = [vp, C.c_int64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_uint32, C.c_uint32, vp, C.c_uint64, u64p]
    lib.hb_last_intervals.argtypes = [vp, vp, C.c_uint64, u64p]
    lib.hb_last_hsps.argtypes = [vp, vp, C.c_uint64, u64p, vp, vp, C.c_uint32, C.POINTER(C.c_uint32)]
    lib.hb_full_length_copies.argtypes = [vp, C.c_uint32, C.c_int, C.c_double, vp, C.c_uint32, C.POINTER(C.c_uint32)]
    lib.hb_fmea_chain_files.argtypes = [vp, vp, C.c_uint32, vp, vp, C.c_int, C.c_int, C.c_int, vp, C.c_uint64, u64p]
    lib.hb_debug_fmea_blocks.argtypes = [vp, vp, C.c_uint32, vp, vp, C.c_int, C.c_int, C.c_int, C.c_uint32, vp, C.c_uint64, u64p]
    lib.hb_full_length_copies_batch.argtypes = [vp, vp, C.c_uint32, vp, C.c_double, C.c_int, vp, vp, C.c_uint64, u64p]
    lib.hb_query_copies_batch.argtypes = [vp, vp, vp, C.c_uint32, vp, C.c_double, C.c_int, C.c_int, C.c_int, C.c_int, vp, vp,
                                          C.c_uint64, u64p]
    lib.hb_genome_mask.argtypes = [vp, C.c_uint32, vp, vp, vp, u64p]
    lib.hb_genome_export.argtypes = [vp, C.c_uint32, vp, C.c_uint64]
    lib.hb_query_hsps.argtypes = [vp, C.c_uint32, vp, vp, C.POINTER(HbParams), u64p]
    lib.hb_last_query_hsps.argtypes = [vp, vp, C.c_uint64, u64p]
    lib.hb_nonltr_candidates.argtypes = [vp, C.c_uint32, vp, vp, C.c_int32, vp]
    lib.hb_itr_search.argtypes = [vp, C.c_uint32, vp, vp, C.c_double, C.c_int32, C.c_int32, vp]
    lib.hb_tsd_candidates.argtypes = [vp, C.c_uint32, vp, vp, vp, vp, C.c_int32, C.c_int32, vp, vp]
    _LIB = lib
    return lib


def _ptr(a):
    return a.ctypes.data_as(C.c_void_p)


def contig_name_rank(contig_names):
    """rank[c] = position of contig c among the names sorted by (name + ':') bytewise: the order of the member names
    ``contig:start-end(strand)`` that tools/ready_for_MSA.sh's ``sort -nk 2 -r`` falls back on (C / POSIX locale)."""
    order = sorted(range(len(contig_names)), key=lambda i: (contig_names[i] + ":").encode())
    rank = np.zeros(len(contig_names), dtype=np.int32)
    rank[order] = np.arange(len(contig_names), dtype=np.int32)
    return rank


def _calls_to_dicts(calls, cons_b):
    """hb_call records -> list of dicts (field columns converted once; structured scalar access is slow)."""
    cols = {k: calls[k].tolist() for k in CALL_DTYPE.names}
    text = cons_b.decode("latin-1")       # 1:1 bytes -> str; slices are pure ASCII
    out = []
    for i in range(len(calls)):
        o = cols["cons_off"][i]
        out.append(dict(is_te=bool(cols["is_te"][i]), info=INFO[cols["info"][i]], cons=text[o:o + cols["cons_len"][i]],
                        copy_num=cols["copy_num"][i], align_start=cols["align_start"][i], align_end=cols["align_end"][i],
                        homo_start=cols["homo_start"][i], homo_end=cols["homo_end"][i], clean_cols=cols["clean_cols"][i],
                        no_copies=bool(cols["flags"][i] & 1), trunc_reject=bool(cols["flags"][i] & 2)))
    return out


class Context:
    """One context per process per GPU (include/hite_b200.h).  Not thread-safe."""

    def __init__(self, device=0):
        self._lib = load()
        h = C.c_void_p()
        rc = self._lib.hb_create(int(device), C.byref(h))
        if rc != 0:
            raise HiteB200Error("hb_create(device=%d) failed (%d): %s" %
                                (device, rc, (self._lib.hb_last_error(None) or b"").decode()))
        self._h = h
        self.device = device

    def close(self):
        if getattr(self, "_h", None):
            self._lib.hb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc, what):
        if rc != 0:
            raise HiteB200Error("%s failed (%d): %s" % (what, rc, (self._lib.hb_last_error(self._h) or b"").decode()))

    def coarse_timing(self):
        """timing of the last ``coarse_boundary`` call, summed over its passes"""
        return self._coarse_timing

    def timing(self):
        t = HbTiming()
        self._lib.hb_get_timing(self._h, C.byref(t))
        return dict(h2d_ms=t.h2d_ms, kernel_ms=t.kernel_ms, d2h_ms=t.d2h_ms, k_ms=list(t.k_ms),
                    h2d_bytes=t.h2d_bytes, d2h_bytes=t.d2h_bytes, launches=t.launches)

    # ------------------------------------------------------------------------------------
    def vote_matrices(self, cands, matrices, te_type="tir", plant=1, result_type="cons"):
        """Judge pre-aligned matrices (kernel iii).  ``matrices`` is a list of lists of equal-length
        row strings (copies only, '-' gaps) exactly as mafft would have written them; returns one dict
        per candidate mirroring ``judge_boundary_v5``'s ``(is_TE, info, cons_seq, copy_num)`` after
        ``remove_sparse_col_in_align_file`` (reference module/Util.py:10419-10428)."""
        n = len(cands)
        assert n == len(matrices)
        if n == 0:
            return []
        R = np.array([len(m) for m in matrices], dtype=np.int32)
        Cc = np.array([len(m[0]) if len(m) else 0 for m in matrices], dtype=np.int32)
        for m, c in zip(matrices, Cc):
            if any(len(r) != c for r in m):
                raise ValueError("ragged alignment matrix")
        mat_off = np.zeros(n + 1, dtype=np.uint64)
        mat_off[1:] = np.cumsum(R.astype(np.uint64) * Cc.astype(np.uint64))
        rows = np.frombuffer("".join("".join(m) for m in matrices).encode("ascii"), dtype=np.uint8)
        cb = [c.encode("ascii") for c in cands]
        cand_off = np.zeros(n + 1, dtype=np.uint64)
        cand_off[1:] = np.cumsum([len(c) for c in cb])
        cand = np.frombuffer(b"".join(cb) + b"\0", dtype=np.uint8)
        calls = np.zeros(n, dtype=CALL_DTYPE)
        cons_cap = int(Cc.astype(np.int64).sum()) + 2 * n
        cons = np.zeros(max(1, cons_cap), dtype=np.uint8)
        rc = self._lib.hb_vote_matrices(self._h, n, _ptr(rows), _ptr(mat_off), _ptr(R), _ptr(Cc), _ptr(cand),
                                        _ptr(cand_off), TE_TYPES[te_type], int(plant),
                                        1 if result_type == "cons" else 0, _ptr(calls), _ptr(cons), cons_cap)
        self._check(rc, "hb_vote_matrices")
        return _calls_to_dicts(calls, cons.tobytes())

    # ------------------------------------------------------------------------------------
    def load_genome(self, contig_seqs, contig_names=None):
        """Kernel (i) part 1.  ``contig_seqs``: list of bytes / uint8 arrays (ASCII, any case).  ``contig_names``: their
        FASTA names -- the "100 longest members" rule breaks length ties by member name (tools/ready_for_MSA.sh:63), so the
        library needs the names' bytewise order (``hb_genome_contig_rank``); without names the contig index is used."""
        seqs = [np.frombuffer(x, dtype=np.uint8) if isinstance(x, (bytes, bytearray)) else
                (np.frombuffer(x.encode("ascii"), dtype=np.uint8) if isinstance(x, str) else np.asarray(x, dtype=np.uint8))
                for x in contig_seqs]
        off = np.zeros(len(seqs) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(x) for x in seqs])
        padded = int(off[-1]) + 64 * (len(seqs) + 2)
        if padded >= (1 << 32):                               # checked before 4 GB are concatenated and shipped
            raise HiteB200Error("hb_genome_load: %d bp in %d contigs do not fit the index's 32-bit positions (limit 2^32 bp including the "
                                "64 N between contigs); run the stage per genome chunk (HiTE --chunk_size, default 400 MB) or on a subset "
                                "of the contigs" % (int(off[-1]), len(seqs)))
        flat = np.concatenate(seqs) if len(seqs) > 1 else np.ascontiguousarray(seqs[0])
        self._check(self._lib.hb_genome_load(self._h, _ptr(flat), _ptr(off), len(seqs)), "hb_genome_load")
        self.contig_lens = [len(x) for x in seqs]
        tm = self.timing()
        if contig_names is not None:
            if len(contig_names) != len(seqs):
                raise ValueError("contig_names and contig_seqs differ in length")
            self.contig_names = list(contig_names)
            rank = contig_name_rank(contig_names)
            self._check(self._lib.hb_genome_contig_rank(self._h, _ptr(rank), len(rank)), "hb_genome_contig_rank")
        return tm

    def build_index(self, k=15, w=10):
        self._check(self._lib.hb_index_build(self._h, int(k), int(w)), "hb_index_build")
        return self.timing()

    def genome_info(self):
        tot, nmm = C.c_uint64(0), C.c_uint64(0)
        goff = np.zeros(max(1, len(getattr(self, "contig_lens", [0]))), dtype=np.uint64)
        self._check(self._lib.hb_genome_info(self._h, C.byref(tot), C.byref(nmm), _ptr(goff), len(goff)), "hb_genome_info")
        return dict(total_padded=tot.value, n_minimizers=nmm.value, contig_goff=goff)

    def debug_scan_total(self, n, value):
        tot, last = C.c_uint64(0), C.c_uint32(0)
        self._check(self._lib.hb_debug_scan_total(self._h, int(n), int(value), C.byref(tot), C.byref(last)), "hb_debug_scan_total")
        return tot.value, last.value

    def debug_index(self):
        n = self.genome_info()["n_minimizers"]
        keys = np.zeros(max(1, n), dtype=np.uint64)
        self._check(self._lib.hb_debug_index(self._h, _ptr(keys), len(keys)), "hb_debug_index")
        return keys[:n]

    def pin_host(self, arr):
        """Page-lock a caller-owned numpy buffer that is passed to ``adjust_raw`` repeatedly (``hb_host_register``): the
        host->device copy then runs as one DMA.  Returns False when the range cannot be locked (the buffer stays usable);
        call ``unpin_host`` before the array is freed."""
        return self._lib.hb_host_register(_ptr(arr), arr.nbytes) == 0

    def unpin_host(self, arr):
        return self._lib.hb_host_unregister(_ptr(arr)) == 0

    def adjust_raw(self, cand_seqs, te_type="tir", plant=1, result_type="cons", occ_cap=512, max_loci=300, flank=50,
                   max_rows=100, tb_budget_bytes=0, trunc_gate=True, overlap=None, rows_first=False, no_judge=False,
                   cons_view=False):
        """One ``flank_region_align_v5``-equivalent round (kernels ii + iii) over the candidates.
        Returns ``(calls, cons)``: a structured array (CALL_DTYPE, one record per candidate) and the
        consensus bytes; candidate i's consensus is ``cons[calls['cons_off'][i] : +calls['cons_len'][i]]``.
        ``overlap`` >= 2: that many candidate sub-batches software-pipelined over two CUDA streams (same bytes
        back; measured on B200 it does not beat the serial batch -- every stage already fills the SMs, see
        tools/overlap_sweep.py -- so the default is one serial batch, which ``last_copies`` / ``last_msa`` need).
        ``cons_view``: the consensus bytes come back as a read-only view of the library's page-locked result buffer
        (``hb_last_cons_view``: no second pass over them), valid until the next call on this context."""
        if isinstance(cand_seqs, tuple):                      # pre-flattened (uint8 array, uint64 offsets)
            flat, off = cand_seqs
            n = len(off) - 1
        else:
            cb = [c if isinstance(c, (bytes, bytearray)) else c.encode("ascii") for c in cand_seqs]
            n = len(cb)
            off = np.zeros(n + 1, dtype=np.uint64)
            off[1:] = np.cumsum([len(c) for c in cb])
            flat = np.frombuffer(b"".join(cb) + b"\0", dtype=np.uint8)
        calls = np.zeros(n, dtype=CALL_DTYPE)
        if n == 0:
            return calls, np.zeros(0, dtype=np.uint8)
        prm = HbParams(occ_cap, max_loci, flank, max_rows, TE_TYPES[te_type], int(plant),
                       1 if result_type == "cons" else 0, 1 if trunc_gate else 0, int(tb_budget_bytes),
                       int(overlap) if overlap is not None else 0, (1 if rows_first else 0) | (2 if no_judge else 0))
        nbytes = C.c_uint64(0)
        self._check(self._lib.hb_adjust(self._h, n, _ptr(flat), _ptr(off), C.byref(prm), _ptr(calls), C.byref(nbytes)),
                    "hb_adjust")
        self.last_calls = calls
        if cons_view:
            p, nb = C.c_void_p(), C.c_uint64(0)
            self._check(self._lib.hb_last_cons_view(self._h, C.byref(p), C.byref(nb)), "hb_last_cons_view")
            if not nb.value:
                return calls, np.zeros(0, dtype=np.uint8)
            view = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(int(nb.value),))
            view.flags.writeable = False
            return calls, view
        cons = np.empty(max(1, nbytes.value), dtype=np.uint8)
        self._check(self._lib.hb_last_cons(self._h, _ptr(cons), len(cons)), "hb_last_cons")
        return calls, cons[:nbytes.value]

    def adjust(self, cand_seqs, *a, **kw):
        """Same as ``adjust_raw`` but returns a list of dicts mirroring ``is_TE_from_align_file``'s tuple
        (reference module/Util.py:10434-10437) per candidate."""
        calls, cons = self.adjust_raw(cand_seqs, *a, **kw)
        return _calls_to_dicts(calls, cons.tobytes())

    def last_cons_device(self):
        """The consensus bytes of the last adjust, still on the GPU, as an object torch can wrap without a
        copy (``torch.as_tensor(ctx.last_cons_device(), device='cuda')``); None when empty."""
        p, n = C.c_void_p(), C.c_uint64(0)
        self._check(self._lib.hb_last_cons_device(self._h, C.byref(p), C.byref(n)), "hb_last_cons_device")
        if not n.value:
            return None

        class _Dev:
            __cuda_array_interface__ = {"shape": (int(n.value),), "typestr": "|u1", "data": (int(p.value), False),
                                        "version": 2}
        return _Dev()

    def last_pair_counts(self, n_cand):
        """loci aligned per candidate in the last adjust (``hb_last_pair_counts``): x length = the cost of a candidate"""
        out = np.zeros(max(1, int(n_cand)), dtype=np.uint32)
        self._check(self._lib.hb_last_pair_counts(self._h, _ptr(out), len(out)), "hb_last_pair_counts")
        return out[:int(n_cand)]

    def last_copies(self):
        n = C.c_uint64(0)
        self._check(self._lib.hb_last_copies(self._h, None, 0, C.byref(n)), "hb_last_copies")
        arr = np.zeros(max(1, n.value), dtype=COPY_DTYPE)
        self._check(self._lib.hb_last_copies(self._h, _ptr(arr), len(arr), C.byref(n)), "hb_last_copies")
        return arr[:n.value]

    def last_msa(self, cand, trunc=False):
        """Rows of candidate ``cand``'s star MSA from the last adjust (``trunc``: its truncated-member matrix)."""
        fn = self._lib.hb_last_msa_trunc if trunc else self._lib.hb_last_msa
        R, Cn = C.c_int32(0), C.c_int32(0)
        self._check(fn(self._h, cand, None, 0, C.byref(R), C.byref(Cn)), "hb_last_msa")
        if R.value == 0:
            return []
        buf = np.zeros(R.value * Cn.value, dtype=np.uint8)
        self._check(fn(self._h, cand, _ptr(buf), len(buf), C.byref(R), C.byref(Cn)), "hb_last_msa")
        raw = buf.tobytes()
        return [raw[r * Cn.value:(r + 1) * Cn.value].decode() for r in range(R.value)]

    # ------------------------------------------------------------------------------------
    def coarse_boundary(self, seg_len=1_000_000, skip_gap=2000, max_len=30000, occ_cap=64, threads=8, part=0, n_parts=1):
        """The all-vs-all + FMEA step of ``determine_repeat_boundary_v5`` (reference module/Util.py:4657-4663) on
        the loaded genome -> INTERVAL_DTYPE array ('chr:start-end', 0-based half open).  ``part`` / ``n_parts``: this
        process's share of a stage sharded over several GPUs (``hb_coarse_boundary_part``): the parts' results
        concatenated in part order are the unsharded result.  A share whose seed pairs do not fit one pass (HB_ENOMEM:
        more than 2^32 pairs, repeat-rich Gbp genomes) is run as 2, 4, ... consecutive sub-parts; ``timing()`` afterwards
        is the sum over the passes."""
        split = 1
        while True:
            outs, tm_sum = [], None
            try:
                for sub in range(split):
                    n = C.c_uint64(0)
                    self._check(self._lib.hb_coarse_boundary_part(self._h, int(seg_len), int(skip_gap), int(max_len), int(occ_cap),
                                                                  int(threads), int(part) * split + sub, int(n_parts) * split, None, 0,
                                                                  C.byref(n)), "hb_coarse_boundary")
                    tm = self.timing()
                    out = np.zeros(max(1, n.value), dtype=INTERVAL_DTYPE)
                    self._check(self._lib.hb_last_intervals(self._h, _ptr(out), len(out), C.byref(n)), "hb_last_intervals")
                    outs.append(out[:n.value])
                    if tm_sum is None:
                        tm_sum = tm
                    else:
                        tm_sum["kernel_ms"] += tm["kernel_ms"]
                        tm_sum["k_ms"] = [a + b for a, b in zip(tm_sum["k_ms"], tm["k_ms"])]
                        tm_sum["launches"] += tm["launches"]
                        tm_sum["d2h_bytes"] += tm["d2h_bytes"]
                self._coarse_timing = tm_sum
                return np.concatenate(outs) if len(outs) > 1 else outs[0]
            except HiteB200Error as e:
                if "(-4)" not in str(e) or split >= 1024:
                    raise
                split *= 2

    def mask_genome(self, contig, start, end):
        """N-fill [start, end) (0-based) of the resident packed genome (``hb_genome_mask``; reference module/Util.py:6414-6420)
        -> bases newly masked.  The index must be rebuilt afterwards."""
        c = np.ascontiguousarray(contig, dtype=np.int32)
        a = np.ascontiguousarray(start, dtype=np.int64)
        b = np.ascontiguousarray(end, dtype=np.int64)
        n = C.c_uint64(0)
        self._check(self._lib.hb_genome_mask(self._h, len(c), _ptr(c), _ptr(a), _ptr(b), C.byref(n)), "hb_genome_mask")
        return int(n.value)

    def export_contig(self, contig, length):
        out = np.zeros(int(length), dtype=np.uint8)
        self._check(self._lib.hb_genome_export(self._h, int(contig), _ptr(out), len(out)), "hb_genome_export")
        return out.tobytes()

    def query_hsps(self, query_seqs, occ_cap=512, max_loci=300):
        """Library-vs-genome local alignments (``hb_query_hsps``; replaces the blastn of reference module/Util.py:7213-7225)
        -> QHSP_DTYPE records ordered by (query, locus rank)."""
        if isinstance(query_seqs, tuple):
            flat, off = query_seqs
        else:
            qb = [q if isinstance(q, (bytes, bytearray)) else q.encode("ascii") for q in query_seqs]
            off = np.zeros(len(qb) + 1, dtype=np.uint64)
            off[1:] = np.cumsum([len(q) for q in qb])
            flat = np.frombuffer(b"".join(qb) + b"\0", dtype=np.uint8)
        prm = HbParams(occ_cap=occ_cap, max_loci=max_loci, flank=50, max_rows=100, te_type=0, plant=1, result_type=1, trunc_gate=0)
        n = C.c_uint64(0)
        self._check(self._lib.hb_query_hsps(self._h, len(off) - 1, _ptr(flat), _ptr(off), C.byref(prm), C.byref(n)), "hb_query_hsps")
        out = np.zeros(n.value, dtype=QHSP_DTYPE)
        if n.value:
            self._check(self._lib.hb_last_query_hsps(self._h, _ptr(out), n.value, C.byref(n)), "hb_last_query_hsps")
        return out

    def nonltr_candidates(self, seqs, flanking_len=50):
        """``hb_nonltr_candidates``: search_polyA_TSD (reference module/Util.py:10915) per flanked candidate -> POLYA_DTYPE [n]."""
        cb = [c if isinstance(c, (bytes, bytearray)) else c.encode("ascii") for c in seqs]
        out = np.zeros(len(cb), dtype=POLYA_DTYPE)
        if not cb:
            return out
        off = np.zeros(len(cb) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(c) for c in cb])
        flat = np.frombuffer(b"".join(cb) + b"\0", dtype=np.uint8)
        self._check(self._lib.hb_nonltr_candidates(self._h, len(cb), _ptr(flat), _ptr(off), int(flanking_len), _ptr(out)),
                    "hb_nonltr_candidates")
        return out

    def itr_search(self, seqs, min_identity=0.7, min_len=7, max_window=500):
        """``hb_itr_search``: HB-itr, the stand-in for the closed ``tools/itrsearch -i 0.7 -l 7`` (reference
        module/Util.py:216-224) -> ITR_DTYPE [n] (found, len5, len3, score, n_match, n_mismatch, n_gap)."""
        cb = [c if isinstance(c, (bytes, bytearray)) else c.encode("ascii") for c in seqs]
        out = np.zeros(len(cb), dtype=ITR_DTYPE)
        if not cb:
            return out
        off = np.zeros(len(cb) + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(c) for c in cb])
        flat = np.frombuffer(b"".join(cb) + b"\0", dtype=np.uint8)
        self._check(self._lib.hb_itr_search(self._h, len(cb), _ptr(flat), _ptr(off), float(min_identity), int(min_len),
                                            int(max_window), _ptr(out)), "hb_itr_search")
        return out

    def tsd_candidates(self, seqs, raw_starts, raw_ends, tsd_search_distance, plant):
        """TSD-bounded variants of flanked candidates (reference ``search_confident_tir_v4`` + the classes of
        ``get_short_tir_contigs``; module/Util.py:7734, :7297) -> (records [n, 100] TSD_DTYPE, counts [n])."""
        cb = [c if isinstance(c, (bytes, bytearray)) else c.encode("ascii") for c in seqs]
        n = len(cb)
        out = np.zeros((n, 100), dtype=TSD_DTYPE)
        cnt = np.zeros(n, dtype=np.uint32)
        if n == 0:
            return out, cnt
        off = np.zeros(n + 1, dtype=np.uint64)
        off[1:] = np.cumsum([len(c) for c in cb])
        flat = np.frombuffer(b"".join(cb) + b"\0", dtype=np.uint8)
        rs = np.ascontiguousarray(raw_starts, dtype=np.int32)
        re_ = np.ascontiguousarray(raw_ends, dtype=np.int32)
        self._check(self._lib.hb_tsd_candidates(self._h, n, _ptr(flat), _ptr(off), _ptr(rs), _ptr(re_), int(tsd_search_distance),
                                                int(plant), _ptr(out), _ptr(cnt)), "hb_tsd_candidates")
        return out, cnt

    def last_hsps(self):
        n, ns = C.c_uint64(0), C.c_uint32(0)
        self._check(self._lib.hb_last_hsps(self._h, None, 0, C.byref(n), None, None, 0, C.byref(ns)), "hb_last_hsps")
        hsps = np.zeros(max(1, n.value), dtype=HSP_DTYPE)
        q_off = np.zeros(ns.value + 1, dtype=np.uint64)
        segs = np.zeros(max(1, ns.value), dtype=SEG_DTYPE)
        self._check(self._lib.hb_last_hsps(self._h, _ptr(hsps), len(hsps), C.byref(n), _ptr(q_off), _ptr(segs), len(segs),
                                           C.byref(ns)), "hb_last_hsps")
        return hsps[:n.value], q_off, segs[:ns.value]
