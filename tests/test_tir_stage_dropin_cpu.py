"""The TIR module's TSD / TIR pre-filter stage end to end on the CPU: hite_b200.tir_prefilter.multi_process_tsd_v1 with an
oracle-backed stand-in for the GPU context against the REFERENCE's own multi_process_tsd_v1 (module/Util.py:6630, imported
unmodified, process pool running), both driving the same stand-in for the external `itrsearch` binary.  The reference's choice
among variants of EQUAL distance follows Python's set order (it changes with the hash seed of the run), so a candidate's record
must be the reference's whenever its closest variant is unique, and one of the tied closest variants otherwise.  Needs the
reference checkout (skipped on the GPU box)."""
import os
import stat
import sys

import numpy as np
import pytest

import tsd_synth
from oracle import oracle as O
from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference checkout not mounted")

FAKE_ITRSEARCH = r'''#!%s
# stand-in for tools/itrsearch -i 0.7 -l 7: reports the longest exact terminal inverted repeat (>= 7 bp) of each record
import sys
path = sys.argv[-1]
comp = {"A": "T", "C": "G", "G": "C", "T": "A"}
out = open(path + ".itr", "w")
name = None
for line in open(path):
    line = line.strip()
    if line.startswith(">"):
        name = line[1:]
    elif line and name:
        rc = "".join(comp.get(b, "N") for b in reversed(line))
        k = 0
        while k < len(line) // 2 and line[k] == rc[k]:
            k += 1
        print("load sequence " + name)
        if k >= 7:
            out.write(">%%s ITR(1,%%d)..(%%d,%%d) Length itr=%%d\n%%s\n" %% (name, k, len(line), len(line) - k + 1, k, line))
        name = None
out.close()
'''


def test_tir_stage_prefilter_equals_the_reference_up_to_its_own_tie_order(tmp_path, monkeypatch):
    from hite_b200 import _lib, tir_prefilter as TP
    monkeypatch.setenv("HITE_B200_ITRSEARCH", "binary")         # both sides run the same (stand-in) itrsearch binary here
    U = ref_import.load_util()
    rng = np.random.default_rng(3)
    comp = str.maketrans("ACGT", "TGCA")
    recs = {}
    for i in range(260):                                        # flanked candidates: 50 bp + TSD-bounded element with a TIR + 50 bp
        c = tsd_synth.synth_tsd_case(40_000 + i)
        if c["d"] != 50 or len(c["seq"]) < 220:
            continue
        recs["chr1:%d-%d" % (i * 5000, i * 5000 + len(c["seq"]))] = c["seq"]
    flanked = str(tmp_path / "longest_repeats_0.flanked.fa")
    with open(flanked, "w") as f:
        for n, s in recs.items():
            f.write(">%s\n%s\n" % (n, s))
    tools = tmp_path / "tools"
    tools.mkdir()
    itr = tools / "itrsearch"
    itr.write_text(FAKE_ITRSEARCH % sys.executable)
    itr.chmod(itr.stat().st_mode | stat.S_IEXEC)

    class OracleBackedCtx:
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

    for plant in (1, 0):
        ref_out = str(tmp_path / ("ref_tsd_%d.fa" % plant))
        U.multi_process_tsd_v1(flanked, ref_out, str(tmp_path / ("ref_dir_%d" % plant)), 50, 4, str(tools), plant)
        mine_out = str(tmp_path / ("mine_tsd_%d.fa" % plant))
        ctx = OracleBackedCtx()
        TP.multi_process_tsd_v1(flanked, mine_out, str(tmp_path / ("mine_dir_%d" % plant)), 50, 4, str(tools), plant, ctx=ctx)
        names, groups, tir_len = TP.tsd_variant_groups(flanked, str(tmp_path / ("g_dir_%d" % plant)), 50, 4, str(tools), plant, ctx=ctx)

        def records(p):
            lines = open(p).read().splitlines()
            return dict(zip((l[1:] for l in lines[0::2]), lines[1::2]))
        ref, mine = records(ref_out), records(mine_out)
        cand_of = lambda rec_name: rec_name.split("-tir_")[0]
        assert {cand_of(n) for n in ref} == {cand_of(n) for n in mine} == set(groups)
        assert len(ref) > 40
        n_unique = n_tied = 0
        ref_by_cand = {cand_of(n): (n, s) for n, s in ref.items()}
        for n, s in mine.items():
            q = cand_of(n)
            dists = sorted(int(v.split("-distance_")[1]) for v in groups[q])
            tied = {"%s-tir_%d-tsd_%s" % (q, tir_len.get(v, 0), v.split("-tsd_")[1].split("-")[0]): groups[q][v]
                    for v in groups[q] if int(v.split("-distance_")[1]) == dists[0]}
            if len(tied) == 1:
                assert ref_by_cand[q] == (n, s), q
                n_unique += 1
            else:
                assert ref_by_cand[q][0] in tied and tied[ref_by_cand[q][0]] == ref_by_cand[q][1], q
                n_tied += 1
        assert n_unique >= 30, (n_unique, n_tied)
        print("plant %d: %d candidates kept, %d with a unique closest variant (identical records), %d tied" % (plant, len(ref), n_unique, n_tied))
