"""Host mirror of the homology non-LTR module's copy search (reference module/judge_Other_transposons.py:45):

* ``multi_process_align_and_get_copies``  reference module/Util.py:8382-8430 (same signature and return value)

The reference formats a BLAST database of the genome, splits the library into 50-query files, runs
``blastn -evalue 1e-20 -outfmt 6`` on each in a process pool (``multiple_alignment_blast_and_get_copies``, :7213-7225)
and chains every file's HSPs with ``get_copies_v1`` -> ``get_query_copies`` (:7032-7060, :6828-7030).  Here the HSPs
come from one ``hb_query_hsps`` call per query batch (the seeding + banded extension kernels of the boundary path, the
library sequences as queries) and the chaining is ``hb_query_copies`` (bit-identical to the reference on the same HSP
list).  There is no CPU fallback.

Parity note: BLAST's own HSP set is not reproducible here (rmblast is a third-party binary absent from the reference
repo); the HSPs are HB-align local alignments -- what is pinned is everything downstream of the
HSP list.
"""
import os
import shutil

import numpy as np

from . import boundary
from .fasta import read_fasta

MAX_QUERY = 32000          # hb_query_hsps limit per query; longer library entries are tiled
BATCH_BASES = 8_000_000    # query bases per hb_query_hsps call
BATCH_QUERIES = 8192


def _tiles(seq):
    """(offset, piece) tiles of a query longer than MAX_QUERY; an alignment that crosses a tile boundary comes back
    as two HSPs that get_query_copies chains again (gap 0 < 200)."""
    if len(seq) <= MAX_QUERY:
        return [(0, seq)]
    return [(o, seq[o:o + MAX_QUERY]) for o in range(0, len(seq), MAX_QUERY)]


def query_hsps(ctx, names, contigs, occ_cap=512, max_loci=128):
    """{query_name: {subject_name: [(q_start, q_end, s_start, s_end, identity), ...]}} -- the ``query_records`` dict of
    get_copies_v1 (module/Util.py:7033-7055), HSPs of a query in (locus rank) order."""
    subj_names = ctx.contig_names
    records = {}
    units = []                                   # (query index, tile offset, bytes)
    for qi, name in enumerate(names):
        for off, piece in _tiles(contigs[name]):
            if len(piece) >= 3:
                units.append((qi, off, piece.encode("ascii")))
    timing = dict(kernel_ms=0.0, pairs=0, hsps=0, calls=0)
    lo = 0
    while lo < len(units):
        hi, bases = lo, 0
        while hi < len(units) and hi - lo < BATCH_QUERIES and (hi == lo or bases + len(units[hi][2]) <= BATCH_BASES):
            bases += len(units[hi][2])
            hi += 1
        recs = ctx.query_hsps([u[2] for u in units[lo:hi]], occ_cap=occ_cap, max_loci=max_loci)
        tm = ctx.timing()
        timing["kernel_ms"] += tm["kernel_ms"]; timing["pairs"] += int(tm["k_ms"][6]); timing["hsps"] += len(recs)
        timing["calls"] += 1
        ident = np.round(100.0 * recs["n_match"] / np.maximum(recs["aln_len"], 1), 3)
        for r, idn in zip(recs, ident):
            qi, off, _ = units[lo + int(r["query"])]
            qname, sname = names[qi], subj_names[int(r["contig"])]
            if qname == sname:                                   # :7045
                continue
            records.setdefault(qname, {}).setdefault(sname, []).append(
                (int(r["qs"]) + off, int(r["qe"]) + off, int(r["ss"]), int(r["se"]), float(idn)))
        lo = hi
    query_hsps.last = timing
    return records


def multi_process_align_and_get_copies(query_path, subject_path, tmp_blast_dir, TE_type, threads, is_removed_dir=True,
                                       query_coverage=0.99, subject_coverage=0):
    """reference module/Util.py:8382.  Returns ``{query_name: [(subject_name, start, end, length, '+'|'-'), ...]}``.
    Like the reference, the chaining runs with get_copies_v1's own defaults (query_coverage 0.95, subject_coverage 0;
    :7221 passes neither argument on)."""
    if is_removed_dir and os.path.exists(tmp_blast_dir):
        shutil.rmtree(tmp_blast_dir, ignore_errors=True)
    os.makedirs(tmp_blast_dir, exist_ok=True)
    names, contigs = read_fasta(query_path)
    ctx = boundary.ensure_genome(boundary.get_context(), subject_path)
    records = query_hsps(ctx, names, contigs)
    return boundary.get_query_copies(list(records.items()), contigs, "", 0.95, 0)
