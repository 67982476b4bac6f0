"""GPU, BASELINE.json configs[4] scale: a maize-sized 2.3 Gbp synthetic genome (positions beyond 2^31 in the
padded coordinate space, a 430 M-entry minimizer index) -- index build + boundary adjustment of candidates whose
copies sit in the last contigs, checked against planted truth and the reference's copy rules.  The oracle is not
run at this size (its index build alone takes minutes); bit-exactness is covered at 1 / 100 Mbp."""
import numpy as np
import pytest

from workloads import synth

pytestmark = pytest.mark.gpu


def test_maize_scale_index_and_adjust():
    import hite_b200
    wl = synth.make_workload(genome_bp=2_300_000_000, n_contigs=10, n_families=60, cands_per_family=5, copies=(10, 40),
                             length=(200, 2000), seed=4)
    ctx = hite_b200.Context(0)
    try:
        ctx.load_genome(wl.contig_seqs)
        ctx.build_index(15, 10)
        info = ctx.genome_info()
        assert info["total_padded"] > (1 << 31)
        # (k=15, w=10) minimizers of i.i.d. sequence: density 2 / (w + 1)
        assert 0.95 < info["n_minimizers"] / (2 * wl.genome_bp / 11.0) < 1.05
        calls, cons = ctx.adjust_raw(wl.cand_seqs, "tir", 1)
        copies = ctx.last_copies()
    finally:
        ctx.close()
    text = cons.tobytes().decode("latin-1")
    te = calls["is_te"] != 0
    n_exact = 0
    for i in np.nonzero(te)[0]:
        got = text[int(calls["cons_off"][i]):int(calls["cons_off"][i]) + int(calls["cons_len"][i])].encode()
        c = wl.truth[i]["consensus"]
        n_exact += got in (c, synth.revcomp_arr(np.frombuffer(c, dtype=np.uint8)).tobytes())
    assert te.sum() >= 0.9 * len(te) and n_exact >= 0.8 * te.sum(), (int(te.sum()), n_exact)
    kept = copies[copies["keep"] != 0]
    clen = np.array([len(c) for c in wl.contig_seqs])
    assert np.all(kept["start1"] - 1 - 50 >= 0) and np.all(kept["end1"] + 50 <= clen[kept["contig"]])
    # copies are found on every contig, also those stored beyond position 2^31, and on both strands
    assert set(np.unique(kept["contig"]).tolist()) == set(range(10))
    assert kept["j_e"].max() > (1 << 31) and set(np.unique(kept["rel"]).tolist()) == {0, 1}
    # the kept copies of a candidate are the planted copies of its family (random families do not cross-hit)
    per = np.bincount(kept["cand"], minlength=len(wl.cand_seqs))
    planted = np.array([t["n_copies"] for t in wl.truth])
    assert 0.8 * planted.sum() <= per.sum() <= 1.05 * planted.sum(), (int(per.sum()), int(planted.sum()))
