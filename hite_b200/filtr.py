"""Host-side mirror of the FiLTR fork of the flank-copy alignment loop -- how ``judge_LTR_transposons.py`` reaches the path
(reference module/judge_LTR_transposons.py:146 -> run_FiLTR, module/Util.py:13847 -> ``python main.py`` ->
``python LTR_filter.py`` -> bin/FiLTR-main/src/LTR_filter.py:66).

* ``generate_both_ends_frame_from_seq_minimap2``  bin/FiLTR-main/src/Util.py:956-1043, same signature.  Its two external
  steps -- the genome-wide ``minimap2 -ax map-ont -N 300`` copy search (:354-395, rules :6509) and one ``mafft`` per
  candidate (``generate_msa`` :1315-1339) -- become ONE ``hb_adjust`` call with ``HB_PRM_ROWS_FIRST | HB_PRM_NO_JUDGE``
  (FiLTR keeps the first 100 members of a file, ``extract_copies`` :1306, and stops at the alignment); what reads the
  alignment -- ``get_both_ends_frame`` (:1403-1500: anchors, sparse columns, the two flank frames) -- stays FiLTR's own code
  and is called on the aligned member file exactly as ``generate_msa`` calls it.
* ``run_FiLTR``  module/Util.py:13847-13857, same signature: the same ``python main.py`` command, started with
  ``HITE_B200_FILTR=1`` and this repository's import hook (module/filtr_hook) on ``PYTHONPATH`` so that the interpreter
  FiLTR spawns for ``LTR_filter.py`` rebinds the function above inside FiLTR's own ``Util``.

Known difference: when more than 100 copies are kept AND only some are longer than 1000 bp, FiLTR takes the first 100 of the
truncated-member file, this takes the long ones among the first 100 kept copies.
"""
import os
import shutil
import sys
import time

from . import boundary
from .fasta import read_fasta


def _filtr_util():
    U = sys.modules.get("Util")
    if U is None or not hasattr(U, "get_both_ends_frame"):
        raise RuntimeError("hite_b200.filtr: FiLTR's Util module (bin/FiLTR-main/src/Util.py) is not imported in this process")
    return U


def generate_both_ends_frame_from_seq_minimap2(candidate_sequence_path, reference, flanking_len, threads, temp_dir, output_dir,
                                               full_length_output_dir, max_copy_num, ctx=None, util=None):
    U = util or _filtr_util()
    debug = 0
    starttime = time.time()
    for d in (temp_dir, output_dir, full_length_output_dir):                      # :961-969
        if os.path.exists(d):
            shutil.rmtree(d)
        os.makedirs(d, exist_ok=True)
    names, contigs = read_fasta(candidate_sequence_path)
    ok = [n for n in names if 3 <= len(contigs[n]) <= 32000]
    if ctx is None:
        ctx = boundary.ensure_genome(boundary.get_context(), reference)
    contig_names = ctx.contig_names
    jobs = []
    for b in range(0, len(ok), 16384):
        batch = ok[b:b + 16384]
        ctx.adjust_raw([contigs[n] for n in batch], "tir", 1, flank=int(flanking_len), max_loci=min(512, 3 * int(max_copy_num)),
                       max_rows=100, rows_first=True, no_judge=True)
        copies = ctx.last_copies()
        kept = copies[copies["keep"] != 0]
        for ci, query_name in enumerate(batch):
            mine = kept[kept["cand"] == ci]
            if len(mine) == 0:
                continue
            trunc = mine[mine["trow"] >= 0]
            use_trunc = len(trunc) > 0                                          # generate_msa :1318-1321
            rows = ctx.last_msa(ci, trunc=use_trunc)
            members = trunc[trunc["trow"].argsort()] if use_trunc else mine[mine["row"].argsort()]
            member_file = temp_dir + "/" + query_name + (".trunc.blast.bed.fa" if use_trunc else ".blast.bed.fa")
            align_file = member_file + ".maf.fa"                                  # :1329
            with open(align_file, "w") as f:
                for m, r in zip(members, rows):
                    f.write(">%s:%d-%d(%s)\n%s\n" % (contig_names[int(m["contig"])], int(m["start1"]), int(m["end1"]),
                                                     "-" if int(m["rel"]) else "+", r))
            first = rows[0].replace("-", "")
            cur_seq = first[int(flanking_len):-int(flanking_len)]                 # :1334
            jobs.append((query_name, cur_seq, align_file))
    print("Running time of get copies: %.8s s" % (time.time() - starttime))
    starttime = time.time()
    from concurrent.futures import ProcessPoolExecutor
    with ProcessPoolExecutor(max(1, int(threads))) as ex:                          # :1021-1028: pure Python, FiLTR's own code
        futs = [ex.submit(U.get_both_ends_frame, q, s, a, output_dir, full_length_output_dir, int(flanking_len), debug)
                for q, s, a in jobs]
        for f in futs:
            f.result()
    if not debug:
        shutil.rmtree(temp_dir)
    print("Running time of MSA: %.8s s" % (time.time() - starttime))


def run_FiLTR(reference, tmp_output_dir, FiLTR_home, threads, miu, recover, debug, is_output_lib, log):
    starttime = time.time()
    log.logger.debug("start FiLTR detection...")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hook = os.path.join(root, "module", "filtr_hook")
    pp = os.pathsep.join([hook, root] + ([os.environ["PYTHONPATH"]] if os.environ.get("PYTHONPATH") else []))
    FiLTR_command = ("cd " + FiLTR_home + " && HITE_B200_FILTR=1 PYTHONPATH=" + pp + " python main.py --genome " + reference
                     + " --out_dir " + tmp_output_dir + " --thread " + str(threads) + " --miu " + str(miu)
                     + " --recover " + str(recover) + " --debug " + str(debug) + " --is_output_lib " + str(is_output_lib))
    log.logger.debug(FiLTR_command)
    os.system(FiLTR_command)
    log.logger.debug("FiLTR running time: %.8s s" % (time.time() - starttime))
