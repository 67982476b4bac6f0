"""GPU parity of kernels (i) and (ii) + star MSA through the C-ABI against the plain-C restatement
(oracle/hbo_align.c) on the same seeded synthetic genome: index keys, copy records, raw MSA bytes
and final boundary calls must be bit-identical; planted elements must be recovered."""
import numpy as np
import pytest

from workloads import synth
from oracle import oracle

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def small(ctx):
    wl = synth.make_workload(genome_bp=1_000_000, n_contigs=2, n_families=12, cands_per_family=3, seed=3)
    G = oracle.Genome(wl.contig_seqs)
    ctx.load_genome(wl.contig_seqs)
    ctx.build_index(15, 10)
    return wl, G


def test_index_bit_exact(ctx, small):
    wl, G = small
    info = ctx.genome_info()
    assert info["total_padded"] == G.total
    keys = ctx.debug_index()
    assert len(keys) == len(G.index)
    assert np.array_equal(keys, G.index)


def test_copies_msa_and_calls_bit_exact(ctx, small):
    wl, G = small
    res = ctx.adjust(wl.cand_seqs, "tir", 1)
    copies = ctx.last_copies()
    n_match = 0
    for ci, (seq, tr) in enumerate(zip(wl.cand_seqs, wl.truth)):
        cps, rows = G.candidate_msa(seq)
        mine = copies[copies["cand"] == ci]
        assert len(mine) == len(cps), ci
        for a, b in zip(mine, cps):
            for k in ("rank", "rel", "score", "i_s", "i_e", "j_s", "j_e", "keep"):
                assert int(a[k]) == int(b[k]), (ci, k, int(a[k]), int(b[k]))
            if b["keep"]:
                assert (int(a["contig"]), int(a["start1"]), int(a["end1"])) == (b["contig"], b["start1"], b["end1"])
        got_rows = ctx.last_msa(ci)
        assert got_rows == rows, ci
        if rows:
            exp = oracle.judge_with_trunc_gate(seq.decode(), cps, rows)
            r = res[ci]
            for k in ("is_te", "info", "cons", "copy_num", "homo_start", "homo_end"):
                assert r[k] == exp[k], (ci, k)
        cons = tr["consensus"].decode()
        rc = synth.revcomp_arr(np.frombuffer(tr["consensus"], dtype=np.uint8)).tobytes().decode()
        n_match += res[ci]["cons"] in (cons, rc)
    assert n_match >= 0.75 * len(wl.cand_seqs)      # planted truth recovered


def test_pinned_input_and_result_view_return_the_same_bytes(ctx, small):
    """hb_host_register + hb_last_cons_view (zero-copy ends of the public call) change nothing in the result, also for an
    overlapped run whose consensus comes back in several pieces."""
    wl, G = small
    flat = np.frombuffer(b"".join(wl.cand_seqs) + b"\0", dtype=np.uint8).copy()
    offs = np.zeros(len(wl.cand_seqs) + 1, dtype=np.uint64)
    offs[1:] = np.cumsum([len(c) for c in wl.cand_seqs])
    calls0, cons0 = ctx.adjust_raw((flat, offs), "tir", 1)
    assert ctx.pin_host(flat)
    try:
        calls1, view = ctx.adjust_raw((flat, offs), "tir", 1, cons_view=True)
        assert not view.flags.writeable
        assert calls1.tobytes() == calls0.tobytes() and view.tobytes() == cons0.tobytes() and len(cons0) > 0
    finally:
        assert ctx.unpin_host(flat)
    many = wl.cand_seqs * 8                                    # >= 64 candidates per part: the overlapped path really splits
    c0, b0 = ctx.adjust_raw(many, "tir", 1, overlap=2)
    c1, b1 = ctx.adjust_raw(many, "tir", 1, overlap=2, cons_view=True)
    assert c1.tobytes() == c0.tobytes() and b1.tobytes() == b0.tobytes() and len(b0) > len(cons0)


def test_edge_cases(ctx, small):
    wl, G = small
    rng = np.random.default_rng(5)
    junk = bytes(np.frombuffer(b"ACGT", dtype=np.uint8)[rng.integers(0, 4, 300)])
    withn = bytearray(wl.cand_seqs[0]); withn[40:45] = b"NNNNN"
    cands = [junk, bytes(withn), wl.cand_seqs[1][:90], wl.cand_seqs[2].lower()]
    res = ctx.adjust(cands, "tir", 1)
    assert res[0]["no_copies"] and not res[0]["is_te"]
    for ci, seq in enumerate(cands):
        cps, rows = G.candidate_msa(seq.upper())
        assert ctx.last_msa(ci) == rows, ci
        if rows:
            exp = oracle.judge_with_trunc_gate(seq.decode().upper(), cps, rows)
            for k in ("is_te", "info", "cons", "copy_num"):
                assert res[ci][k] == exp[k], (ci, k)


def test_traceback_budget_chunking_is_invariant(ctx, small):
    wl, G = small
    a = ctx.adjust(wl.cand_seqs[:12], "tir", 1)
    b = ctx.adjust(wl.cand_seqs[:12], "tir", 1, tb_budget_bytes=4 << 20)
    assert a == b


def test_stress_shapes_bit_exact(ctx):
    """High-copy families (> max_loci clusters, > 100 surviving copies), long elements (wide band,
    multi-super-step traceback), N runs and lower-case input: GPU == C port on every record."""
    wl = synth.make_workload(genome_bp=3_000_000, n_contigs=3, n_families=3, cands_per_family=2, copies=(150, 220),
                             length=(300, 700), seed=21, indel=0.002)
    wl2 = synth.make_workload(genome_bp=1_500_000, n_contigs=1, n_families=4, cands_per_family=2, copies=(6, 12),
                              length=(4000, 9000), seed=22, indel=0.004, sub=(0.0, 0.05))
    # > 768 valid seed clusters for one candidate: the cluster ranking leaves its shared-memory fast path
    wl3 = synth.make_workload(genome_bp=1_500_000, n_contigs=1, n_families=1, cands_per_family=2, copies=(900, 950),
                              length=(250, 300), seed=23, indel=0.002, sub=(0.0, 0.03))
    contigs = [c.copy() for c in wl.contig_seqs] + [c.copy() for c in wl2.contig_seqs] + [c.copy() for c in wl3.contig_seqs]
    contigs[0][1000:1400] = ord("N")
    contigs[1][5000:5003] = ord("n")
    contigs[2] = np.frombuffer(contigs[2].tobytes().lower(), dtype=np.uint8).copy()
    cands = list(wl.cand_seqs) + list(wl2.cand_seqs) + list(wl3.cand_seqs)
    G = oracle.Genome(contigs)
    ctx.load_genome(contigs)
    ctx.build_index(15, 10)
    assert np.array_equal(ctx.debug_index(), G.index)
    res = ctx.adjust(cands, "tir", 1)
    copies = ctx.last_copies()
    n_big = 0
    for ci, seq in enumerate(cands):
        cps, rows = G.candidate_msa(seq)
        mine = copies[copies["cand"] == ci]
        assert len(mine) == len(cps), ci
        for a, b in zip(mine, cps):
            for k in ("rank", "rel", "score", "i_s", "i_e", "j_s", "j_e", "keep"):
                assert int(a[k]) == int(b[k]), (ci, k, int(a[k]), int(b[k]))
        assert ctx.last_msa(ci) == rows, ci
        n_big += len(rows) == 100
        if rows:
            exp = oracle.judge_with_trunc_gate(seq.decode(), cps, rows)
            for k in ("is_te", "info", "cons", "copy_num"):
                assert res[ci][k] == exp[k], (ci, k)
    assert n_big >= 1                       # the 100-longest rule was exercised
    assert sum(r["is_te"] for r in res[len(wl.cand_seqs):len(wl.cand_seqs) + len(wl2.cand_seqs)]) >= 1      # long elements come through


def test_100_longest_tie_rule_follows_member_names(ctx):
    """Families with 130-180 kept copies of EQUAL length (no indels): every member ties at the 100-record cut, so the subset
    is decided by tools/ready_for_MSA.sh's last-resort order -- member names `contig:start-end(strand)` bytewise descending
    (reference module/Util.py:10409-10412; pinned on the real script in tests/test_flank_loop_dropin_cpu.py).  Twelve contigs
    named chr1..chr12 make the name order differ from the contig index ('chr10:' < 'chr1:').  GPU rows == oracle rows under
    that rule, and != the rows the earlier-ranked rule would have kept."""
    wl = synth.make_workload(genome_bp=6_000_000, n_contigs=12, n_families=4, cands_per_family=2, copies=(130, 180),
                             length=(150, 600), seed=78, indel=0.0)
    G = oracle.Genome(wl.contig_seqs)
    G.set_tie_rule(1, wl.contig_names)
    ctx.load_genome(wl.contig_seqs, wl.contig_names)
    ctx.build_index(15, 10)
    res = ctx.adjust(wl.cand_seqs, "tir", 1)
    copies = ctx.last_copies()
    n_cut = n_differs = 0
    for ci, seq in enumerate(wl.cand_seqs):
        cps, rows = G.candidate_msa(seq)
        mine = copies[copies["cand"] == ci]
        assert [int(a["keep"]) for a in mine] == [int(b["keep"]) for b in cps], ci
        assert ctx.last_msa(ci) == rows, ci
        if len(rows) == 100:
            n_cut += 1
            G.set_tie_rule(0)
            n_differs += G.candidate_msa(seq)[1] != rows
            G.set_tie_rule(1, wl.contig_names)
        exp = oracle.judge_with_trunc_gate(seq.decode(), cps, rows)
        for k in ("is_te", "info", "cons", "copy_num"):
            assert res[ci][k] == exp[k], (ci, k)
    assert n_cut >= 4 and n_differs >= 2
    # without names the library falls back to the contig index: still the oracle's rule under the same convention
    G.set_tie_rule(1)
    ctx.load_genome(wl.contig_seqs)
    ctx.build_index(15, 10)
    ctx.adjust(wl.cand_seqs[:2], "tir", 1)
    for ci, seq in enumerate(wl.cand_seqs[:2]):
        assert ctx.last_msa(ci) == G.candidate_msa(seq)[1]


def test_scan_total_is_exact_beyond_32_bits(ctx):
    """The size guards of hb_adjust (seed hits < 2^31) and hb_coarse_boundary (seed pairs < 2^32) read the grand total of a
    device scan whose offsets are 32-bit: the total must stay exact where the offsets wrap."""
    n, v = 1 << 20, 9000                      # 9.4e9 > 2^33
    total, last = ctx.debug_scan_total(n, v)
    assert total == n * v
    assert last == ((n - 1) * v) % (1 << 32)
    total, last = ctx.debug_scan_total(5000, 7)
    assert (total, last) == (35000, 4999 * 7)


def test_chained_pieces_bit_exact(ctx):
    """HB-align S5b / S8b: copies that differ from the candidate by a block indel longer than the band (20 - 300 bp) or whose
    diagonal drifts (LTR-sized, 3 % indels) come back from the banded extension as several colinear pieces; they are chained
    into one copy (what minimap2 reports as one alignment with a long I / D).  GPU == C port on every copy record, MSA row
    and call; chains really occur and recover copies the single pieces would not."""
    import os
    wl = synth.make_mixed_workload(3_000_000, 2, 60, copies=("uniform", 8, 20), cands_per_family=6, seed=11,
                                   lengths={"tir": (400, 2500)}, sub=(0.03, 0.03), indel=(0.002, 0.002), block=(1, 20, 120))
    wl2 = synth.make_mixed_workload(3_000_000, 1, 12, mix=(("ltr", 1.0),), copies=("uniform", 6, 10), cands_per_family=4, seed=12,
                                    sub=(0.02, 0.02), indel=(0.03, 0.03))
    contigs = list(wl.contig_seqs) + list(wl2.contig_seqs)
    names = list(wl.contig_names) + ["ltr1"]
    cands = list(wl.cand_seqs) + list(wl2.cand_seqs)
    G = oracle.Genome(contigs)
    G.set_tie_rule(1, names)
    ctx.load_genome(contigs, names)
    ctx.build_index(15, 10)
    res = ctx.adjust(cands, "tir", 1)
    copies = ctx.last_copies()
    n_kept = 0
    for ci, seq in enumerate(cands):
        cps, rows = G.candidate_msa(seq)
        mine = copies[copies["cand"] == ci]
        assert len(mine) == len(cps), ci
        for a, b in zip(mine, cps):
            for k in ("rank", "rel", "score", "i_s", "i_e", "j_s", "j_e", "keep"):
                assert int(a[k]) == int(b[k]), (ci, k, int(a[k]), int(b[k]))
        assert ctx.last_msa(ci) == rows, ci
        n_kept += len(rows)
        if rows:
            exp = oracle.judge_with_trunc_gate(seq.decode(), cps, rows)
            for k in ("is_te", "info", "cons", "copy_num"):
                assert res[ci][k] == exp[k], (ci, k)
    os.environ["HBO_NO_CHAIN"] = "1"                       # the port without S8b: fewer copies survive the coverage rules
    try:
        n_unchained = sum(len(G.candidate_msa(seq)[1]) for seq in cands)
    finally:
        del os.environ["HBO_NO_CHAIN"]
    assert n_kept >= n_unchained + 30, (n_kept, n_unchained)


def test_filtr_mode_first_100_rows_no_judge_and_truncated_matrix(ctx):
    """HB_PRM_ROWS_FIRST | HB_PRM_NO_JUDGE (the FiLTR fork of the loop, bin/FiLTR-main/src/Util.py:956-1043, :1306, :1315-1339):
    flank 100, the FIRST 100 kept copies, no judge; full and truncated-member matrices == the C port's."""
    wl = synth.make_workload(genome_bp=3_000_000, n_contigs=2, n_families=3, cands_per_family=2, copies=(120, 150),
                             length=(300, 700), seed=41, indel=0.002)
    wl2 = synth.make_workload(genome_bp=1_500_000, n_contigs=1, n_families=3, cands_per_family=2, copies=(6, 12),
                              length=(1500, 4000), seed=42, indel=0.003, sub=(0.0, 0.04))
    contigs = list(wl.contig_seqs) + list(wl2.contig_seqs)
    cands = list(wl.cand_seqs) + list(wl2.cand_seqs)
    G = oracle.Genome(contigs)
    G.set_tie_rule(2)
    ctx.load_genome(contigs)
    ctx.build_index(15, 10)
    calls, cons = ctx.adjust_raw(cands, "tir", 1, flank=100, rows_first=True, no_judge=True)
    assert len(cons) == 0
    n_cut = n_trunc = 0
    for ci, seq in enumerate(cands):
        cps, rows = G.candidate_msa(seq, flank=100)
        assert ctx.last_msa(ci) == rows, ci
        assert int(calls["copy_num"][ci]) == len(rows)
        t = oracle.truncated_rows(cps, rows)
        assert ctx.last_msa(ci, trunc=True) == t, ci
        n_cut += len(rows) == 100
        n_trunc += len(t) > 0
    assert n_cut >= 3 and n_trunc >= 3
