"""The whole flank-copy alignment loop end to end on the CPU, reference against mirror, on the same copies and the same MSA.

The REFERENCE's own flank_region_align_v5 (module/Util.py:8032-8299, imported unmodified) runs with its process pool, its
run_find_members_v8 / is_TE_from_align_file / remove_sparse_col_in_align_file / judge_boundary_v5|v6|v9 and its post filter.
Only its externals are replaced: the minimap2 copy search returns the oracle's copy records, `mafft` on $PATH is a stand-in
that prints the oracle's star MSA of exactly the member file it is given (and refuses if the file's sequences are not the rows'
ungapped content -- so the reference's own copy extraction :8095-8137 is checked against ours), and the low-copy rescue
(itrsearch / blastx, out of scope here as in the mirror) rescues nothing.  The MIRROR hite_b200.boundary.flank_region_align_v5
runs with a stand-in context whose verdicts come from the oracle.  Both must write the same confident-TE and low-copy records:
this pins the glue around the kernels -- extension and flank rules, reverse strand, the truncated-member gate, 'TG..CA' and
copy-number thresholds, file contract -- on the reference's code itself.  Needs the reference checkout (skipped on the GPU box)."""
import json
import os
import re
import stat
import sys

import pytest

from hite_b200 import synth
from oracle import oracle as O
from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference checkout not mounted")

FAKE_MAFFT = r'''#!%s
import json, os, sys
path = sys.argv[-1]
db = json.load(open(os.environ["HB_FAKE_MAFFT_DB"]))
names, seqs = [], []
for line in open(path):
    line = line.strip()
    if line.startswith(">"):
        names.append(line[1:]); seqs.append("")
    elif line:
        seqs[-1] += line.upper()
rows = db[os.path.basename(path)]
if len(rows) != len(seqs) or any(r.replace("-", "") != s for r, s in zip(rows, seqs)):
    sys.stderr.write("fake mafft: member file does not hold the rows' sequences: " + path + "\n")
    sys.exit(3)
for n, r in zip(names, rows):
    sys.stdout.write(">" + n + "\n" + r + "\n")
'''


class _Log:
    class logger:
        @staticmethod
        def info(msg):
            pass


@pytest.mark.parametrize("te_type,seed,plant", [("tir", 23, 1), ("tir", 24, 0), ("helitron", 23, 1), ("non_ltr", 23, 1)])
def test_reference_loop_equals_mirror_on_the_same_copies_and_msa(tmp_path, monkeypatch, te_type, seed, plant):
    from hite_b200 import boundary, fasta
    U = ref_import.load_util()
    wl = synth.make_workload(genome_bp=1_500_000, n_contigs=2, n_families=14, cands_per_family=3, copies=(3, 30),
                             length=(150, 1900), seed=seed, kind=te_type)
    ref, cand = str(tmp_path / "genome.fa"), str(tmp_path / "cand.fa")
    wl.write_fasta(ref, cand)
    names, contigs = fasta.read_fasta(cand)
    G = O.Genome(wl.contig_seqs)
    judge = {"tir": (lambda c, r: O.judge_v5(c, r, plant)), "helitron": lambda c, r: O.judge_v6(c, r), "non_ltr": lambda c, r: O.judge_v9(c, r)}[te_type]
    all_copies, db, verdicts = {}, {}, {}
    n_trunc = 0
    for name in names:
        seq = contigs[name]
        cps, rows = G.candidate_msa(seq.encode())
        kept = [c for c in cps if c["keep"]]
        if rows:
            all_copies[name] = [(wl.contig_names[c["contig"]], c["start1"], c["end1"], c["end1"] - c["start1"] + 1,
                                 "-" if c["rel"] else "+") for c in kept]
            fname = re.sub(r'[<>:"/\\|?*]', "-", name)
            db[fname + ".blast.bed.fa"] = rows
            t = O.truncated_rows(cps, rows)
            if t:
                db[fname + ".trunc.blast.bed.fa"] = t
                n_trunc += 1
            verdicts[seq] = O.judge_with_trunc_gate(seq, cps, rows, judge=judge, plant=plant)
        else:
            verdicts[seq] = dict(is_te=False, info="", cons="", copy_num=0)
    G.close()
    assert n_trunc >= 5 and len(all_copies) >= 30
    # ---- the reference, externals replaced ----------------------------------------------------------------------
    bindir = tmp_path / "bin"
    bindir.mkdir()
    mafft = bindir / "mafft"
    mafft.write_text(FAKE_MAFFT % sys.executable)
    mafft.chmod(mafft.stat().st_mode | stat.S_IEXEC)
    dbpath = str(tmp_path / "msa_db.json")
    json.dump(db, open(dbpath, "w"))
    monkeypatch.setenv("PATH", str(bindir) + os.pathsep + os.environ["PATH"])
    monkeypatch.setenv("HB_FAKE_MAFFT_DB", dbpath)
    monkeypatch.setattr(U, "get_full_length_copies_minimap2", lambda q, r, t, m, th: dict(all_copies))

    def no_tirs(confident_tir_path, plant, TRsearch_dir, temp_dir):          # low-copy rescue: nothing rescued
        a, b = confident_tir_path + ".with", confident_tir_path + ".without"
        open(a, "w").close()
        open(b, "w").close()
        return a, b

    def no_domains(cons, lib, output_table, threads, temp_dir):
        open(output_table, "w").write("#\n#\n")
    monkeypatch.setattr(U, "remove_no_tirs", no_tirs)
    monkeypatch.setattr(U, "get_domain_info", no_domains)
    out_ref, low_ref = str(tmp_path / "ref_real.fa"), str(tmp_path / "ref_low.fa")
    U.flank_region_align_v5(cand, out_ref, 50, ref, None, te_type, str(tmp_path / "refwork"), 4, 0, _Log, "", plant, 0, 0, low_ref, "cons")

    # ---- the mirror, verdicts from the oracle -------------------------------------------------------------------
    class OracleBackedCtx:
        device = 0

        def adjust(self, seqs, te, plant, result_type, flank=50):
            assert te == te_type and flank == 50 and plant == globals_plant
            return [dict(verdicts[s]) for s in seqs]

    globals_plant = plant
    ctx = OracleBackedCtx()
    monkeypatch.setattr(boundary, "get_context", lambda device=None: ctx)
    monkeypatch.setattr(boundary, "ensure_genome", lambda c, r, k=15, w=10, _parsed=None: c)
    monkeypatch.setenv("HITE_GPU_DEVICES", "0")
    out_mine, low_mine = str(tmp_path / "mine_real.fa"), str(tmp_path / "mine_low.fa")
    boundary.flank_region_align_v5(cand, out_mine, 50, ref, None, te_type, str(tmp_path / "minework"), 4, 0, _Log, "", plant, 0, 0, low_mine, "cons")

    def records(p):
        lines = open(p).read().splitlines() if os.path.exists(p) else []
        return sorted(zip(lines[0::2], lines[1::2]))
    assert records(out_mine) == records(out_ref)
    assert records(low_mine) == records(low_ref)
    assert len(records(out_ref)) + len(records(low_ref)) >= 15
    print("%s: %d candidates, %d with copies, %d through the truncated-member gate -> %d confident, %d low-copy records (reference == mirror)"
          % (te_type, len(names), len(all_copies), n_trunc, len(records(out_ref)), len(records(low_ref))))
