"""Host mirror of the non-LTR candidate generator that runs right before the boundary adjustment in HiTE's non-LTR module
(reference module/judge_Non_LTR_transposons.py:34): same function names, arguments, return shapes and output files as

* ``search_polyA_TSD``                 reference module/Util.py:10915-11006
* ``get_candidate_non_LTR``            reference module/Util.py:11008-11045
* ``get_candidate_non_ltr_parallel``   reference module/Util.py:11047-11081 (writes candidate_SINE.fa / candidate_LINE.fa)

The per-candidate scan (poly-A / poly-T / tandem tail near the raw ends, TSD of 8-20 bp with one edit within 25 bp of the
5' end) runs in ``hb_nonltr_candidates`` (hite_b200/csrc/hb_tsd.cu), one warp per candidate; the reference's process pool
over FASTA partitions becomes one kernel launch.  No CPU fallback.
"""
from . import _lib
from .fasta import read_fasta, store_fasta

_CTX = None
_COMP = {"A": "T", "T": "A", "C": "G", "G": "C"}


def _ctx():
    """The stage process's context (shared with hite_b200.boundary: one context per process per GPU, HITE_GPU_DEVICES honoured)."""
    from . import boundary
    return boundary.get_context()


def _reverse_sequence(s):                                     # getReverseSequence, module/Util.py:1635
    return "".join(_COMP.get(b, "N") for b in reversed(s))


def _unpack(seq, r):
    found, direct = bool(r["found"]), int(r["direct"])
    tsd = ""
    if found:
        e3, k = int(r["end_3"]), int(r["tsd_k"])
        tsd = seq[e3:e3 + k] if direct == 1 else seq[e3 - k:e3]
    body = seq[int(r["seq_start"]):int(r["seq_end"])] if direct else ""
    if direct == 2:
        body = _reverse_sequence(body)
    return found, tsd, body


def search_polyA_TSD(seq, flanking_len, end_5_window_size=25, TSD_sizes=None, ctx=None):
    """reference module/Util.py:10915 -> (found_TSD, TSD_seq, non_ltr_seq).  The kernel implements the one configuration the
    reference uses (:11009-11010): a 25 bp window and TSD sizes 8..20."""
    if end_5_window_size != 25 or (TSD_sizes is not None and list(TSD_sizes) != list(range(8, 21))):
        raise _lib.HiteB200Error("search_polyA_TSD: only end_5_window_size=25, TSD_sizes=8..20 (the reference's call) is built")
    recs = (ctx or _ctx()).nonltr_candidates([seq], flanking_len)
    return _unpack(seq, recs[0])


def get_candidate_non_LTR(longest_repeats_flanked_path, flanking_len=50, ctx=None):
    """reference module/Util.py:11008 -> (candidate_SINE_contigs, candidate_LINE_contigs), SINE 100-700 bp, LINE 700-8000 bp."""
    contig_names, contigs = read_fasta(longest_repeats_flanked_path)
    sine_names, line_names = [], []
    for name in contig_names:
        seq_len = len(contigs[name]) - 2 * flanking_len
        if 100 <= seq_len <= 700:
            sine_names.append(name)
        elif 700 < seq_len <= 8000:
            line_names.append(name)
    names = sine_names + line_names
    recs = (ctx or _ctx()).nonltr_candidates([contigs[n] for n in names], flanking_len)
    get_candidate_non_LTR.last_kernel_ms = (ctx or _ctx()).timing()["kernel_ms"] if names else 0.0
    sine, line = {}, {}
    for i, name in enumerate(names):
        found, tsd, body = _unpack(contigs[name], recs[i])
        if not found:
            continue
        if i < len(sine_names):
            if 100 <= len(body) <= 700:
                sine[name + "\tTSD:" + tsd] = body
        elif 700 < len(body) <= 8000:
            line[name + "\tTSD:" + tsd] = body
    return sine, line


def get_candidate_non_ltr_parallel(longest_repeats_flanked_path, work_dir, threads, ctx=None):
    """reference module/Util.py:11047: same output files; the partition / process-pool fan-out is one kernel launch."""
    sine, line = get_candidate_non_LTR(longest_repeats_flanked_path, ctx=ctx)
    candidate_SINE_path = work_dir + "/candidate_SINE.fa"
    candidate_LINE_path = work_dir + "/candidate_LINE.fa"
    store_fasta(sine, candidate_SINE_path)
    store_fasta(line, candidate_LINE_path)
    return candidate_SINE_path, candidate_LINE_path
