"""The FiLTR fork of the flank-copy alignment loop (how judge_LTR_transposons.py reaches the path), reference against mirror.

FiLTR's own generate_both_ends_frame_from_seq_minimap2 (bin/FiLTR-main/src/Util.py:956-1043, imported unmodified with its
third-party imports stubbed) runs its member-file extraction, its process pool, generate_msa / extract_copies and
get_both_ends_frame; only its externals are replaced: the minimap2 copy search returns the oracle's copy records and `mafft`
is a stand-in that prints the oracle's star MSA of exactly the member file it is handed (refusing any other content).
The mirror hite_b200.filtr.generate_both_ends_frame_from_seq_minimap2 runs with a context whose copies and matrices come
from the oracle.  Both must leave byte-identical frame files (<query>.matrix) in output_dir and full_length_output_dir.
Also: the import hook rebinds the function inside FiLTR's Util in a fresh interpreter, and only there."""
import importlib
import json
import os
import stat
import subprocess
import sys
import types
from unittest.mock import MagicMock

import numpy as np
import pytest

from workloads import synth
from oracle import oracle as O
from oracle import ref_import
from test_flank_loop_dropin_cpu import FAKE_MAFFT

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference checkout not mounted")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
FILTR_SRC = os.path.join(ref_import.REF_ROOT, "bin", "FiLTR-main", "src")


def _load_filtr_util():
    """FiLTR's src/Util.py as module 'filtr_Util' (HiTE's own Util keeps the name 'Util' in this process)."""
    if "filtr_Util" in sys.modules:
        return sys.modules["filtr_Util"]
    from oracle import fnm
    for name in ("pysam", "intervaltree"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = MagicMock()
    fz = types.ModuleType("fuzzysearch"); fz.find_near_matches = fnm.find_near_matches
    lv = types.ModuleType("Levenshtein"); lv.distance = fnm.levenshtein
    saved = {k: sys.modules.get(k) for k in ("fuzzysearch", "Levenshtein")}
    sys.modules["fuzzysearch"], sys.modules["Levenshtein"] = fz, lv
    import warnings
    spec = importlib.util.spec_from_file_location("filtr_Util", os.path.join(FILTR_SRC, "Util.py"))
    mod = importlib.util.module_from_spec(spec)
    sys.modules["filtr_Util"] = mod
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        spec.loader.exec_module(mod)
    for k, v in saved.items():
        if v is not None:
            sys.modules[k] = v
    return mod


def test_filtr_loop_equals_mirror_on_the_same_copies_and_msa(tmp_path, monkeypatch):
    from hite_b200 import _lib, filtr
    FU = _load_filtr_util()
    flank = 100
    wl = synth.make_mixed_workload(3_000_000, 2, 24, mix=(("ltr", 0.5), ("tir", 0.5)), copies=("uniform", 5, 14), cands_per_family=3,
                                   seed=21, lengths={"ltr": (4000, 6000), "tir": (300, 900)}, sub=(0.0, 0.04), indel=(0.0, 0.004))
    ref = str(tmp_path / "genome.fa")
    wl.write_fasta(ref, str(tmp_path / "unused.fa"))
    # FiLTR's candidates are LTR terminal sequences: here the first 600 bp of LTR-like candidates, and whole TIR candidates
    cands = {}
    for i, (n, s) in enumerate(zip(wl.cand_names, wl.cand_seqs)):
        cands["cand_%d" % i] = (s[:600] if wl.truth[i]["kind"] == "ltr" else s).decode()
    cpath = str(tmp_path / "left_ltr.fa")
    with open(cpath, "w") as f:
        for n, s in cands.items():
            f.write(">%s\n%s\n" % (n, s))
    G = O.Genome(wl.contig_seqs)
    G.set_tie_rule(2)                                              # FiLTR's extract_copies: the first 100 members
    lens = [len(c) for c in wl.contig_seqs]
    all_copies, db, per_cand = {}, {}, {}
    for name, seq in cands.items():
        cps, rows = G.candidate_msa(seq.encode(), flank=flank)
        kept = [c for c in cps if c["keep"]]
        if not rows:
            continue
        all_copies[name] = [(wl.contig_names[c["contig"]], c["start1"], c["end1"], c["end1"] - c["start1"] + 1, "-" if c["rel"] else "+")
                            for c in kept]
        t = O.truncated_rows(cps, rows)
        if t:
            db[name + ".trunc.blast.bed.fa"] = t
        else:
            db[name + ".blast.bed.fa"] = rows
        per_cand[seq] = (kept, rows, t)
    G.close()
    assert len(all_copies) >= 18 and any(k.endswith(".trunc.blast.bed.fa") for k in db) and any(not k.endswith(".trunc.blast.bed.fa") for k in db)
    bindir = tmp_path / "bin"
    bindir.mkdir()
    mafft = bindir / "mafft"
    mafft.write_text(FAKE_MAFFT % sys.executable)
    mafft.chmod(mafft.stat().st_mode | stat.S_IEXEC)
    dbpath = str(tmp_path / "msa_db.json")
    json.dump(db, open(dbpath, "w"))
    monkeypatch.setenv("PATH", str(bindir) + os.pathsep + os.environ["PATH"])
    monkeypatch.setenv("HB_FAKE_MAFFT_DB", dbpath)
    monkeypatch.setattr(FU, "get_full_length_copies_minimap2", lambda q, r, t, m, th: dict(all_copies))
    ref_out, ref_full = str(tmp_path / "ref_frames"), str(tmp_path / "ref_full")
    FU.generate_both_ends_frame_from_seq_minimap2(cpath, ref, flank, 3, str(tmp_path / "ref_tmp"), ref_out, ref_full, 100)

    class OracleBackedCtx:
        device = 0
        contig_names = wl.contig_names

        def adjust_raw(self, seqs, te, plant, flank=50, max_loci=300, max_rows=100, rows_first=False, no_judge=False):
            assert flank == 100 and rows_first and no_judge and max_loci == 300
            self.batch = list(seqs)
            return None, None

        def last_copies(self):
            recs = []
            for ci, s in enumerate(self.batch):
                if s not in per_cand:
                    continue
                kept, rows, t = per_cand[s]
                trow = 0
                for r, c in enumerate(kept):
                    long_ = c["end1"] - c["start1"] + 1 + 2 * flank > 1000
                    recs.append((ci, c["rank"], c["rel"], 0, c["score"], c["i_s"], c["i_e"], 1, r, c["contig"], 0, c["j_s"], c["j_e"],
                                 c["start1"], c["end1"], c["end1"] - c["start1"] + 1 + 2 * flank, trow if long_ else -1))
                    trow += long_
            return np.array(recs, dtype=_lib.COPY_DTYPE)

        def last_msa(self, ci, trunc=False):
            kept, rows, t = per_cand[self.batch[ci]]
            return t if trunc else rows
    my_out, my_full = str(tmp_path / "my_frames"), str(tmp_path / "my_full")
    filtr.generate_both_ends_frame_from_seq_minimap2(cpath, ref, flank, 3, str(tmp_path / "my_tmp"), my_out, my_full, 100,
                                                     ctx=OracleBackedCtx(), util=FU)
    for a, b in ((ref_out, my_out), (ref_full, my_full)):
        fa, fb = sorted(os.listdir(a)), sorted(os.listdir(b))
        assert fa == fb and len(fa) >= 15, (len(fa), len(fb))
        for n in fa:
            assert open(os.path.join(a, n)).read() == open(os.path.join(b, n)).read(), n
    assert not os.path.exists(str(tmp_path / "my_tmp"))                   # :1033 temp_dir removed
    print("FiLTR loop: %d candidates, %d with copies -> %d frame files (reference == mirror)" % (len(cands), len(all_copies), len(os.listdir(ref_out))))


def test_import_hook_rebinds_only_filtrs_util(tmp_path):
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import types, unittest.mock as m\n"
            "for n in ('pysam','intervaltree','fuzzysearch','Levenshtein'): sys.modules[n] = m.MagicMock()\n"
            "import Util\n"
            "print(Util.generate_both_ends_frame_from_seq_minimap2.__module__)\n" % FILTR_SRC)
    hook = os.path.join(ROOT, "module", "filtr_hook")
    env = dict(os.environ, PYTHONPATH=os.pathsep.join([hook, ROOT]), HITE_B200_FILTR="1")
    out = subprocess.run([sys.executable, "-W", "ignore", "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0, out.stderr[-2000:]
    assert out.stdout.strip().splitlines()[-1] == "hite_b200.filtr"
    env.pop("HITE_B200_FILTR")
    out = subprocess.run([sys.executable, "-W", "ignore", "-c", code], capture_output=True, text=True, env=env, timeout=300)
    assert out.returncode == 0 and out.stdout.strip().splitlines()[-1] == "Util"          # the hook is inert without the switch
