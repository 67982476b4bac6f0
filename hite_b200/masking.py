"""Host mirror of the known-TE genome masking (the step right before the all-vs-all search in the coarse-boundary stage and
before LTR detection):

* ``mask_genome_intactTE``  reference module/Util.py:6389-6431 -- same signature, writes ``genome_path + '.masked'``
* ``mask_genome_with_TE``   reference module/Util.py:4699-4705 -- same signature

The reference maps the library to the genome with minimap2 (:6318), rewrites the PAF as outfmt 6 (:6335-6374, which drops
the strand: subject spans always ascend), chains each query's records into full-length copies
(``get_full_length_copies_from_blastn_v1`` :5907, threshold 0.95) and overwrites every copy with N.  Here the records are
``hb_query_hsps`` local alignments (seeding + banded extension kernels), the chaining is ``hb_full_length_copies``
(bit-identical to the reference on the same record list) and the N-fill runs both on the file's text (so letters outside
ACGT survive exactly as in the reference) and on the packed genome resident in HBM (``hb_genome_mask``), which the next
stage reuses without another upload.  No CPU fallback.
"""
import os
import shutil

import numpy as np

from . import _lib, boundary, other
from .fasta import file_exist, read_fasta, store_fasta


def records_to_intervals(records, query_lens, subj_names, full_length_threshold=0.95):
    """``records``: {query_name: {subject_name: [(q_start, q_end, s_start, s_end, ...), ...]}} (outfmt-6 conventions) ->
    list of (query, subject index, start 0-based, end exclusive).  Query names are cut at '#' and subject spans made
    ascending, as the reference does before chaining (:5913-5917, :5924, :6351-6354)."""
    subj_index = {n: i for i, n in enumerate(subj_names)}
    merged, qlen = {}, {}
    for name, ln in query_lens.items():                       # later names win, like new_query_contigs (:5913-5916)
        qlen[name.split("#")[0]] = ln
    for qname, subject_dict in records.items():
        q = qname.split("#")[0]
        rows = merged.setdefault(q, [])
        for sname, hs in subject_dict.items():
            si = subj_index[sname]
            for h in hs:
                rows.append((si, h[0], h[1], min(h[2], h[3]), max(h[2], h[3])))
    out = []
    for q, rows in merged.items():
        if q not in qlen:                                     # :5944
            continue
        hs = np.array(rows, dtype=_lib.HSP_DTYPE)
        for c in _lib.full_length_copies(hs, qlen[q], full_length_threshold):
            out.append((q, int(c["subj"]), int(c["start"]), int(c["end"])))
    return out


def fill_text(contigs, names, intervals):
    """ref_seq_array[start:end] = 'N' (:6414-6420) on the FASTA text; returns {name: masked sequence}."""
    arrs = {n: np.frombuffer(contigs[n].encode("ascii"), dtype=np.uint8).copy() for n in names}
    for _, si, a, b in intervals:
        arr = arrs[names[si]]
        a, b = max(0, a), min(len(arr), b)
        if b > a:
            arr[a:b] = ord("N")
    return {n: arrs[n].tobytes().decode("ascii") for n in names}


def intervals_from_hsp_records(rec, names, contigs, full_length_threshold=0.95, threads=8):
    """The array form of ``records_to_intervals`` over ``other.query_hsp_records`` output: one
    ``hb_full_length_copies_batch`` call.  Queries whose names agree up to '#' are merged (:5913-5917, :5924)."""
    keys = [n.split("#")[0] for n in names]
    gid_of, glen = {}, []
    for n, k in zip(names, keys):
        if k not in gid_of:
            gid_of[k] = len(glen); glen.append(0)
        glen[gid_of[k]] = len(contigs[n])                     # later names win, like new_query_contigs
    gid = np.array([gid_of[k] for k in keys], dtype=np.int64)[rec["query"]] if len(rec) else np.zeros(0, dtype=np.int64)
    order = np.argsort(gid, kind="stable")
    g = gid[order]
    hs = np.zeros(len(rec), dtype=_lib.HSP_DTYPE)
    hs["subj"] = rec["subj"][order]; hs["qs"] = rec["qs"][order]; hs["qe"] = rec["qe"][order]
    hs["ss"] = np.minimum(rec["ss"], rec["se"])[order]        # the PAF -> outfmt-6 rewrite drops the strand (:6351-6354)
    hs["se"] = np.maximum(rec["ss"], rec["se"])[order]
    q_off = np.searchsorted(g, np.arange(len(glen) + 1)).astype(np.uint64)
    out, oq = _lib.full_length_copies_batch(hs, q_off, np.asarray(glen, dtype=np.int32), full_length_threshold, threads)
    gkeys = list(gid_of.keys())
    return [(gkeys[qg], subj, start, end) for qg, (subj, start, end, _, _) in zip(oq.tolist(), out.tolist())]


def mask_context(ctx, lib_names, lib_contigs, full_length_threshold=0.95, threads=8, record_len=None):
    """Mask the genome resident in ``ctx`` with the library: HSPs -> full-length copies -> ``hb_genome_mask`` + index
    rebuild.  Returns (intervals, newly masked bases).  ``record_len``: the resident contigs stand for FASTA records of this
    length (the 1 Mbp ``chr$start`` records of a genome chunk, split_genome_chunks.py:44-52): the reference aligns the
    library to those records, so a copy lying across a record border is two partial hits on two subjects, never a
    full-length copy, and stays unmasked -- such intervals are dropped here too."""
    rec = other.query_hsp_records(ctx, lib_names, lib_contigs)
    intervals = intervals_from_hsp_records(rec, lib_names, lib_contigs, full_length_threshold, threads)
    if record_len:
        intervals = [iv for iv in intervals if iv[2] // record_len == (iv[3] - 1) // record_len]
    fresh = 0
    if intervals:
        fresh = ctx.mask_genome([i[1] for i in intervals], [i[2] for i in intervals], [i[3] for i in intervals])
        ctx.build_index(15, 10)
    return intervals, fresh


def mask_genome_intactTE(TE_lib, genome_path, work_dir, thread, ref_index, debug=0):
    masked_genome_path = genome_path + ".masked"
    if TE_lib is not None and file_exist(TE_lib):
        lib_names, lib_contigs = read_fasta(TE_lib)
        ctx = boundary.ensure_genome(boundary.get_context(), genome_path)
        intervals, _ = mask_context(ctx, lib_names, lib_contigs, threads=max(1, int(thread)))
        boundary._GENOME.pop(ctx.device, None)                 # the resident genome is now the masked one, not the file
        ref_names, ref_contigs = read_fasta(genome_path)
        store_fasta(fill_text(ref_contigs, ref_names, intervals), masked_genome_path)
    else:
        shutil.copy(genome_path, masked_genome_path)


def mask_genome_with_TE(prev_TE, filter_tandem_file, tmp_output_dir, threads, ref_index, debug=0):
    masked_file_path = filter_tandem_file + ".masked"
    mask_genome_intactTE(prev_TE, filter_tandem_file, tmp_output_dir, threads, ref_index, debug=debug)
    return masked_file_path
