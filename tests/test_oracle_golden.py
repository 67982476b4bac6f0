"""CPU: the plain-C oracle (oracle/hbo_judge.c) against golden vectors produced by the imported,
unmodified reference functions (tests/golden/make_golden.py)."""
import pytest

from oracle import fnm, oracle


def test_oracle_judge_v5_matches_reference_goldens(golden_v5):
    assert len(golden_v5) >= 100
    for rec in golden_v5:
        exp = rec["expect"]
        clean = oracle.remove_sparse_cols(rec["rows"])
        assert clean == exp["clean_rows"], rec["seed"]
        r = oracle.judge_v5(rec["cand"], clean, rec["plant"])
        for k in ("is_te", "info", "cons", "copy_num"):
            assert r[k] == exp[k], (rec["seed"], rec["scenario"], k)
        if exp["homo_start"] != -1:
            assert r["homo_start"] == exp["homo_start"] and r["homo_end"] == exp["homo_end"], rec["seed"]


def test_golden_scenarios_cover_edge_cases(golden_v5):
    kinds = {(r["scenario"], r["expect"]["is_te"], r["expect"]["info"]) for r in golden_v5}
    assert any(k[2] == "nb" for k in kinds) and any(k[2] == "fl1" for k in kinds)
    assert any(k[0] == "many_rows" for k in kinds) and any(k[0] == "fp_homolog_flank" and not k[1] for k in kinds)
    # planted truth: clean scenarios recover the planted element exactly most of the time
    clean = [r for r in golden_v5 if r["scenario"] == "clean"]
    assert sum(r["expect"]["cons"] == r["truth"] for r in clean) >= 0.6 * len(clean)


@pytest.mark.parametrize("seed", range(40))
def test_fnm_c_equals_python_shim(seed):
    import random
    rng = random.Random(seed)
    p = "".join(rng.choice("ACGT") for _ in range(20))
    s = "".join(rng.choice("ACGT") for _ in range(rng.randint(0, 300)))
    # plant 0..3 noisy occurrences
    for _ in range(rng.randint(0, 3)):
        q = list(p)
        for _ in range(rng.randint(0, 3)):
            op = rng.choice("sid")
            i = rng.randrange(len(q))
            if op == "s":
                q[i] = rng.choice("ACGT")
            elif op == "i":
                q.insert(i, rng.choice("ACGT"))
            elif len(q) > 1:
                del q[i]
        at = rng.randint(0, len(s))
        s = s[:at] + "".join(q) + s[at:]
    if seed % 7 == 0:
        p = "A" * 20
        s = s[:10] + "A" * 45 + s[10:]
    ms = fnm.find_near_matches(p, s, max_l_dist=2)
    n, first, last = oracle.find_near_matches_ends(p, s, 2)
    assert n == len(ms)
    if ms:
        assert first == (ms[0].start, ms[0].end)
        assert last == (ms[-1].start, ms[-1].end)


def test_levenshtein():
    assert oracle.levenshtein("ACGTA", "ACGTA") == 0
    assert oracle.levenshtein("ACGTA", "TCGTT") == 2
    assert oracle.levenshtein("", "ACG") == 3
    assert fnm.levenshtein("kitten", "sitting") == 3 == oracle.levenshtein("kitten", "sitting")


def test_oracle_judge_v6_matches_reference_goldens(golden_v6):
    assert len(golden_v6) >= 100
    for rec in golden_v6:
        exp = rec["expect"]
        clean = oracle.remove_sparse_cols(rec["rows"])
        assert clean == exp["clean_rows"], rec["seed"]
        r = oracle.judge_v6(rec["cand"], clean)
        for k in ("is_te", "info", "cons", "copy_num"):
            assert r[k] == exp[k], (rec["seed"], rec["scenario"], k)


def test_oracle_judge_v9_matches_reference_goldens(golden_v9):
    assert len(golden_v9) >= 100
    for rec in golden_v9:
        exp = rec["expect"]
        clean = oracle.remove_sparse_cols(rec["rows"])
        assert clean == exp["clean_rows"], rec["seed"]
        r = oracle.judge_v9(rec["cand"], clean)
        for k in ("is_te", "info", "cons", "copy_num"):
            assert r[k] == exp[k], (rec["seed"], rec["scenario"], k)


def test_oracle_fmea_matches_reference_goldens():
    import gzip, json, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    recs = json.loads(gzip.open(os.path.join(root, "tests", "golden", "fmea.json.gz")).read())
    assert len(recs) >= 150
    for rec in recs:
        r = oracle.fmea(tuple(rec["query"]), [tuple(s) for s in rec["subjects"]], rec["self_subj"],
                        [tuple(h) for h in rec["hsps"]], rec["skip_gap"], rec["max_len"])
        assert r == rec["expect"], rec["seed"]


def test_oracle_query_copies_matches_reference_goldens():
    import gzip, json, os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    recs = json.loads(gzip.open(os.path.join(root, "tests", "golden", "query_copies.json.gz")).read())
    for rec in recs:
        r = oracle.query_copies([tuple(h) for h in rec["hsps"]], rec["idents"], rec["qlen"], rec["subj_lens"],
                                rec["query_coverage"], rec["subject_coverage"], max_copy_num=rec["max_copy_num"])
        assert r == [tuple(x) for x in rec["expect"]], rec["seed"]
