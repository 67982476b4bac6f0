"""Host-side mirror of the coarse-boundary stage (reference module/coarse_boundary.py).

* ``determine_repeat_boundary_v5``  reference module/Util.py:4637-4669 -- same signature.  The all-vs-all
  ``blastn`` of 1 Mbp segments (process_blast_alignments, :4724-4780) becomes one self-join of the minimizer
  index on the GPU (``hb_coarse_boundary``); FMEA chaining (get_longest_repeats_v4, :4122) runs in the C-ABI
  library; ``generate_final_result`` (:4783-4796) slices the intervals out of the reference genome.
  The known-TE masking of :4654 (``mask_genome_with_TE``) runs on the chunk already resident in HBM
  (``hite_b200.masking.mask_context``: ``prev_TE`` vs chunk HSPs, full-length chaining, N-fill kernel, index rebuild).
  The tandem-repeat filter of :4650-4651 is the reference's own ``filter_tandem_repeats`` (:4672-4697, ``trf`` per 100 kbp
  file under its process pool), called as it is when the HiTE checkout is importable and ``trf`` is on ``$PATH``; otherwise
  the stage says so loudly and takes the chunk unfiltered.
* ``flanking_seq``  reference module/Util.py:4614-4634 -- same signature and clamping rules.
"""
import os
import shutil
import sys

import numpy as np

from .boundary import _GENOME, _reference_util, get_context
from .fasta import file_exist, read_fasta, store_fasta


def _filter_tandem_repeats(repeats_path, tmp_output_dir, ref_index, threads):
    """module/Util.py:4650-4651 by the reference's own code: returns the path of ``filter_tandem_{ref_index}.fa`` (the chunk
    with its tandem repeats N-masked by ``trf``), or ``repeats_path`` itself -- with a loud message -- when that cannot run."""
    U = _reference_util()
    why = None
    if U is None or not hasattr(U, "filter_tandem_repeats"):
        why = "the HiTE checkout is not importable (set HITE_HOME)"
    elif shutil.which("trf") is None:
        why = "trf not found on this host"
    if why:
        sys.stderr.write("hite_b200: tandem-repeat filter (trf) SKIPPED for %s: %s; the chunk is searched unfiltered\n"
                         % (os.path.basename(repeats_path), why))
        return repeats_path
    repeat_names, repeat_contigs = U.read_fasta(repeats_path)
    U.filter_tandem_repeats(repeat_names, repeat_contigs, tmp_output_dir, ref_index, max(1, int(threads)))
    return os.path.join(tmp_output_dir, "filter_tandem_%s.fa" % ref_index)


def _chunk_to_chromosomes(repeats_path, name_order=None):
    """genome.cut{i}.fa holds records ``chr$start`` (split_genome_chunks.py:47-52).  Rebuild per-chromosome
    arrays (pieces missing from this chunk stay 'N') so segments sit at their real coordinates.  ``name_order``: record
    order to follow instead of the file's (the trf-filtered file is written in process-pool completion order)."""
    names, contigs = read_fasta(repeats_path)
    if name_order is not None:
        names = [n for n in name_order if n in contigs] + [n for n in names if n not in set(name_order)]
    pieces = {}
    order = []
    for n in names:
        if "$" in n:
            c, off = n.rsplit("$", 1)
            off = int(off)
        else:
            c, off = n, 0
        if c not in pieces:
            pieces[c] = []
            order.append(c)
        pieces[c].append((off, contigs[n]))
    chroms = []
    for c in order:
        end = max(o + len(s) for o, s in pieces[c])
        arr = np.full(end, ord("N"), dtype=np.uint8)
        for o, s in pieces[c]:
            arr[o:o + len(s)] = np.frombuffer(s.encode("ascii"), dtype=np.uint8)
        chroms.append(arr)
    return order, chroms


def determine_repeat_boundary_v5(repeats_path, longest_repeats_path, prev_TE, fixed_extend_base_threshold,
                                 max_single_repeat_len, tmp_output_dir, threads, ref_index, reference, debug,
                                 chrom_seg_length=1_000_000, occ_cap=64):
    if not os.path.exists(tmp_output_dir):
        os.makedirs(tmp_output_dir)
    filtered = _filter_tandem_repeats(repeats_path, tmp_output_dir, ref_index, threads)          # :4650-4651
    chr_names, chroms = _chunk_to_chromosomes(filtered, read_fasta(repeats_path)[0] if filtered != repeats_path else None)
    ctx = get_context()
    _GENOME.pop(ctx.device, None)                 # this context now holds the chunk, not the reference
    ctx.load_genome(chroms)
    ctx.build_index(15, 10)
    ctx.contig_names = chr_names
    if prev_TE is not None and file_exist(prev_TE):              # mask_genome_with_TE (module/Util.py:4654)
        from . import masking
        te_names, te_contigs = read_fasta(prev_TE)
        masking.mask_context(ctx, te_names, te_contigs, threads=max(1, int(threads)), record_len=int(chrom_seg_length))
    iv = ctx.coarse_boundary(seg_len=int(chrom_seg_length), skip_gap=int(fixed_extend_base_threshold),
                             max_len=int(max_single_repeat_len), occ_cap=int(occ_cap), threads=max(1, int(threads)))
    ref_names, ref_contigs = read_fasta(reference)
    final = {}
    for o in iv:                                  # generate_final_result (module/Util.py:4783-4796)
        c = chr_names[int(o["chr"])]
        a, b = int(o["start"]), int(o["end"])
        final["%s:%d-%d" % (c, a, b)] = ref_contigs[c][a:b]
    store_fasta(final, longest_repeats_path)
    return longest_repeats_path


def flanking_seq(longest_repeats_path, longest_repeats_flanked_path, reference, flanking_len):
    seq_names, _ = read_fasta(longest_repeats_path)
    _, ref_contigs = read_fasta(reference)
    flanked = {}
    for name in seq_names:
        ref_name, pos = name.split(":")[0], name.split(":")[1]
        a, b = pos.split("-")
        start1, end1 = int(a) + 1, int(b)                      # 1-based inclusive
        seq = ref_contigs[ref_name]
        if start1 - 1 - flanking_len < 0:                      # clamp so a full flank exists (:4625-4628)
            start1 = flanking_len + 1
        if end1 + flanking_len > len(seq):
            end1 = len(seq) - flanking_len
        flanked["%s:%d-%d" % (ref_name, start1 - flanking_len, end1 + flanking_len)] = \
            seq[start1 - 1 - flanking_len:end1 + flanking_len]
    store_fasta(flanked, longest_repeats_flanked_path)
