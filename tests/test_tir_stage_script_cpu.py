"""The TIR module as main.py runs it, end to end on the CPU through THIS repository's drop-in script:
module/judge_TIR_transposons.py -> the reference's own stage code (argparse, the TSD/TIR pre-filter call, cd-hit-est call, three
boundary-adjustment rounds, length filter, rename_fasta, lib_add_prefix, update_prev_TE under its FileLock, copy-back) with
multi_process_tsd_v1 and flank_region_align_v5 rebound to the mirrors.  The GPU context is stood in by the oracle, `itrsearch`
and `cd-hit-est` by small stand-ins.  A functional check of the whole stage boundary (the reference's own run needs minimap2 /
mafft / itrsearch): planted TIR families come out as confident_tir_{i}.fa records with their exact sequence, and prev_TE grows.
Needs the reference checkout (skipped on the GPU box)."""
import os
import stat
import sys

import numpy as np
import pytest

from workloads import synth
from oracle import oracle as O
from oracle import ref_import
from test_tir_stage_dropin_cpu import FAKE_ITRSEARCH

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference checkout not mounted")

FAKE_CDHIT = r'''#!%s
import shutil, sys
a = sys.argv
shutil.copy(a[a.index("-i") + 1], a[a.index("-o") + 1])          # no clustering: every candidate is its own cluster
'''


def test_tir_stage_script_end_to_end(tmp_path, monkeypatch):
    from hite_b200 import _lib, boundary, fasta, tir_prefilter
    ref_import.load_util()
    wl = synth.make_workload(genome_bp=1_200_000, n_contigs=2, n_families=10, cands_per_family=2, copies=(8, 20),
                             length=(250, 1400), seed=91)
    ref = str(tmp_path / "genome.fa")
    wl.write_fasta(ref, str(tmp_path / "cand_unused.fa"))
    _, ref_contigs = fasta.read_fasta(ref)
    # longest_repeats_{i}.flanked.fa as the coarse stage writes it: 'chr:start-end' with +-50 bp (module/Util.py:4614-4634)
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

        def tsd_candidates(self, seqs, raw_starts, raw_ends, d, plant):
            out = np.zeros((len(seqs), 100), dtype=_lib.TSD_DTYPE)
            cnt = np.zeros(len(seqs), dtype=np.uint32)
            for i, (s, a, b) in enumerate(zip(seqs, raw_starts, raw_ends)):
                r = O.search_confident_tir_v4(s, a, b, d, plant)
                cnt[i] = len(r)
                for j, x in enumerate(r):
                    for f in ("k", "start", "end", "dist", "idx", "short_class"):
                        out[i][j][f] = x[f]
            return out, cnt

        def itr_search(self, seqs, min_identity=0.7, min_len=7, max_window=500):     # HB-itr, from its CPU restatement
            out = np.zeros(len(seqs), dtype=_lib.ITR_DTYPE)
            for i, s in enumerate(seqs):
                for f, v in O.itr_search(s, min_identity, min_len, max_window).items():
                    out[i][f] = v
            return out

        def adjust(self, seqs, te, plant, result_type, flank=50):
            res = []
            for s in seqs:
                cps, rows = G.candidate_msa(s.encode())
                res.append(O.judge_with_trunc_gate(s, cps, rows, plant=int(plant)) if rows
                           else dict(is_te=False, info="", cons="", copy_num=0))
            return res
    ctx = OracleBackedCtx()
    monkeypatch.setattr(boundary, "get_context", lambda device=None, role="reference": ctx)
    monkeypatch.setattr(boundary, "ensure_genome", lambda c, r, k=15, w=10, _parsed=None: c)
    bindir = tmp_path / "bin"
    bindir.mkdir()
    for tool, body in (("itrsearch", FAKE_ITRSEARCH), ("cd-hit-est", FAKE_CDHIT)):
        p = bindir / tool
        p.write_text(body % sys.executable)
        p.chmod(p.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setattr(tir_prefilter, "_ctx", lambda: ctx)      # run_itrsearch (HB-itr) asks for the stage's context
    monkeypatch.setenv("PATH", str(bindir) + os.pathsep + os.environ["PATH"])
    monkeypatch.setenv("HITE_HOME", ref_import.REF_ROOT)
    monkeypatch.setenv("HITE_GPU_DEVICES", "0")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    monkeypatch.syspath_prepend(os.path.join(root, "module"))
    import _hite_b200_patch as patch
    outdir, work = tmp_path / "out", tmp_path / "work"
    outdir.mkdir(); work.mkdir()
    prev_te, low = str(tmp_path / "prev_TE.fa"), str(tmp_path / "all_low_copy_tir.fa")
    open(prev_te, "w").close()
    monkeypatch.setattr(sys, "argv", ["judge_TIR_transposons.py", "--seqs", fpath, "-t", "4", "--tmp_output_dir", str(outdir),
                                      "--tandem_region_cutoff", "0.5", "--ref_index", "0", "--plant", "1", "--flanking_len", "50",
                                      "--recover", "0", "--debug", "0", "-r", ref, "--split_ref_dir", str(tmp_path / "ref_chr"),
                                      "--prev_TE", prev_te, "--all_low_copy_tir", low, "--min_TE_len", "80", "-w", str(work)])
    import Util
    saved = {n: getattr(Util, n) for n in ("flank_region_align_v5", "run_find_members_v8", "is_TE_from_align_file", "multi_process_tsd_v1",
                                            "run_itrsearch")}
    try:
        patch.run_stage("judge_TIR_transposons.py", extra=(("multi_process_tsd_v1", "hite_b200.tir_prefilter"),
                                                           ("run_itrsearch", "hite_b200.tir_prefilter")))
    finally:
        for n, f in saved.items():
            setattr(Util, n, f)
        G.close()
    names, got = fasta.read_fasta(str(outdir / "confident_tir_0.fa"))
    assert len(names) >= 8 and all("TIR_0_" in n for n in names)                  # rename_fasta + lib_add_prefix
    comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
    truth = set()
    for t in wl.truth:
        s = t["consensus"].decode()
        truth.add(s)
        truth.add("".join(comp[b] for b in reversed(s)))
    exact = sum(1 for s in got.values() if s in truth)
    assert exact >= 0.7 * len(names), (exact, len(names))                         # planted elements recovered base for base
    assert os.path.getsize(prev_te) > 0 and open(prev_te).read().count(">") == len(names)     # update_prev_TE (:6378)
    assert os.path.exists(str(outdir / "HiTE_tir.log"))
    print("TIR stage: %d flanked candidates -> %d confident TIR records, %d equal to a planted element base for base" % (len(flanked), len(names), exact))
