"""The Helitron module as main.py runs it, end to end on the CPU through THIS repository's drop-in script:
module/judge_Helitron_transposons.py -> the reference's own stage code (argparse, candidate collection, cd-hit-est calls, the
boundary-adjustment round with judge_boundary_v6 semantics, length filter, rename_fasta, update_prev_TE, copy-back) with
flank_region_align_v5 rebound to the mirror.  HelitronScanner (java, multi_process_helitronscanner) is replaced by a stand-in
that hands over the de-flanked candidates; the GPU context is stood in by the oracle, `cd-hit-est` by a copying stand-in, the
blastx domain search finds nothing.  Needs the reference checkout (skipped on the GPU box)."""
import os
import stat
import sys

import pytest

from workloads import synth
from oracle import oracle as O
from oracle import ref_import
from test_tir_stage_script_cpu import FAKE_CDHIT

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference checkout not mounted")


def test_helitron_stage_script_end_to_end(tmp_path, monkeypatch):
    from hite_b200 import boundary, fasta
    ref_import.load_util()
    wl = synth.make_workload(genome_bp=1_200_000, n_contigs=2, n_families=10, cands_per_family=2, copies=(8, 20),
                             length=(250, 1400), seed=95, kind="helitron")
    ref = str(tmp_path / "genome.fa")
    wl.write_fasta(ref, str(tmp_path / "cand_unused.fa"))
    _, ref_contigs = fasta.read_fasta(ref)
    flanked = {}
    for name in wl.cand_names:
        c, span = name.split(":")
        a, b = (int(x) for x in span.split("-"))
        if a - 50 >= 0 and b + 50 <= len(ref_contigs[c]):
            flanked["%s:%d-%d" % (c, a - 50, b + 50)] = ref_contigs[c][a - 50:b + 50]
    fpath = str(tmp_path / "longest_repeats_0.flanked.fa")
    fasta.store_fasta(flanked, fpath)
    G = O.Genome(wl.contig_seqs)

    class OracleBackedCtx:
        device = 0

        def adjust(self, seqs, te, plant, result_type, flank=50):
            assert te == "helitron"
            res = []
            for s in seqs:
                cps, rows = G.candidate_msa(s.encode())
                res.append(O.judge_with_trunc_gate(s, cps, rows, judge=lambda c, r: O.judge_v6(c, r)) if rows
                           else dict(is_te=False, info="", cons="", copy_num=0))
            return res
    ctx = OracleBackedCtx()
    monkeypatch.setattr(boundary, "get_context", lambda device=None, role="reference": ctx)
    monkeypatch.setattr(boundary, "ensure_genome", lambda c, r, k=15, w=10, _parsed=None: c)
    bindir = tmp_path / "bin"
    bindir.mkdir()
    p = bindir / "cd-hit-est"
    p.write_text(FAKE_CDHIT % sys.executable)
    p.chmod(p.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(bindir) + os.pathsep + os.environ["PATH"])
    monkeypatch.setenv("HITE_HOME", ref_import.REF_ROOT)
    monkeypatch.setenv("HITE_GPU_DEVICES", "0")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    monkeypatch.syspath_prepend(os.path.join(root, "module"))
    import _hite_b200_patch as patch
    outdir, work = tmp_path / "out", tmp_path / "work"
    outdir.mkdir(); work.mkdir()
    prev_te, low = str(tmp_path / "prev_TE.fa"), str(tmp_path / "all_low_copy_helitron.fa")
    open(prev_te, "w").close()
    monkeypatch.setattr(sys, "argv", ["judge_Helitron_transposons.py", "--seqs", fpath, "-t", "4", "--tmp_output_dir", str(outdir),
                                      "--flanking_len", "50", "--ref_index", "0", "--recover", "0", "--debug", "0", "-r", ref,
                                      "--split_ref_dir", str(tmp_path / "ref_chr"), "--prev_TE", prev_te,
                                      "--all_low_copy_helitron", low, "--min_TE_len", "80", "-w", str(work)])
    import Util

    def fake_helitronscanner(flanked_path, out_path, sh_dir, temp_dir, HSDIR, HSJAR, threads, debug):
        n, c = fasta.read_fasta(flanked_path)
        fasta.store_fasta({x: c[x][50:-50] for x in n}, out_path)          # what HelitronScanner would cut out of the flanked record
    names_patched = ("flank_region_align_v5", "run_find_members_v8", "is_TE_from_align_file", "multi_process_helitronscanner", "get_domain_info")
    saved = {n: getattr(Util, n) for n in names_patched}
    Util.multi_process_helitronscanner = fake_helitronscanner
    Util.get_domain_info = lambda cons, lib, output_table, threads, temp_dir: open(output_table, "w").write("#\n#\n")
    try:
        patch.run_stage("judge_Helitron_transposons.py")
    finally:
        for n, f in saved.items():
            setattr(Util, n, f)
        G.close()
    names, got = fasta.read_fasta(str(outdir / "confident_helitron_0.fa"))
    assert len(names) >= 5 and all("Helitron_0_" in n for n in names)
    truth = {t["consensus"].decode() for t in wl.truth}
    close = sum(1 for s in got.values() if any(s in t or t in s for t in truth))
    assert close >= 0.6 * len(names), (close, len(names))
    assert open(prev_te).read().count(">") == len(names)
    print("Helitron stage: %d flanked candidates -> %d confident records, %d matching a planted element" % (len(flanked), len(names), close))
