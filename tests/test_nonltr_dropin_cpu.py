"""Non-LTR candidate generator end to end on the CPU: hite_b200.nonltr_prefilter.get_candidate_non_ltr_parallel with an
oracle-backed stand-in for the GPU context against the REFERENCE's own get_candidate_non_ltr_parallel (module/Util.py:11047,
imported unmodified; its process pool runs for real): candidate_SINE.fa / candidate_LINE.fa must be byte-identical.  Needs the
reference checkout (skipped on the GPU box)."""
import os

import numpy as np
import pytest

from oracle import oracle as O
from oracle import ref_import
from polya_synth import polya_case

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference checkout not mounted")


def test_candidate_files_equal_the_reference(tmp_path):
    from hite_b200 import _lib, nonltr_prefilter as NP
    U = ref_import.load_util()
    cases = [c for c in (polya_case(s) for s in range(60_000, 60_900)) if c["flank"] == 50]
    path = str(tmp_path / "longest_repeats_0.flanked.fa")
    with open(path, "w") as f:
        for c in cases:
            f.write(">chr1:%d-%d\n%s\n" % (c["seed"] * 10_000, c["seed"] * 10_000 + len(c["seq"]), c["seq"]))

    class OracleBackedCtx:
        def nonltr_candidates(self, seqs, flanking_len=50):
            out = np.zeros(len(seqs), dtype=_lib.POLYA_DTYPE)
            for i, s in enumerate(seqs):
                rec = O.search_polyA_TSD(s, flanking_len)[3]
                for k, v in rec.items():
                    out[i][k] = v
            return out

        def timing(self):
            return dict(kernel_ms=0.0)

    mine_dir, ref_dir = str(tmp_path / "mine"), str(tmp_path / "ref")
    os.makedirs(mine_dir); os.makedirs(ref_dir)
    sp, lp = NP.get_candidate_non_ltr_parallel(path, mine_dir, 4, ctx=OracleBackedCtx())
    rsp, rlp = U.get_candidate_non_ltr_parallel(path, ref_dir, 4)
    # the reference merges its partitions in completion order (as_completed): compare as record sets, names included
    def records(p):
        lines = open(p).read().splitlines()
        return sorted(zip(lines[0::2], lines[1::2]))
    assert records(sp) == records(rsp) and records(lp) == records(rlp)
    assert len(records(sp)) > 20 and len(records(lp)) > 60
