"""CPU: host logic, C-ABI surface, FASTA contract, no-CPU-fallback behaviour, gloo sharding."""
import ctypes
import os
import re
import subprocess
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    txt = open(os.path.join(ROOT, "include", "hite_b200.h")).read()
    txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
    return sorted(set(re.findall(r"\b(hb_[a-z_0-9]+)\s*\(", txt)))


def test_cabi_library_exports_every_declared_symbol():
    import __graft_entry__ as g
    g.build()
    from hite_b200 import _lib
    lib = ctypes.CDLL(_lib.lib_path())
    syms = _declared_symbols()
    assert len(syms) >= 14
    for s in syms:
        assert hasattr(lib, s), s
    assert lib.hb_version() >= 100


def test_no_cpu_fallback_without_gpu():
    import torch
    import hite_b200
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(hite_b200.HiteB200Error):
        hite_b200.Context(0)
    from hite_b200 import boundary
    with pytest.raises(hite_b200.HiteB200Error):
        boundary.get_context(0)


def test_product_never_imports_oracle():
    for dp, _, fs in os.walk(os.path.join(ROOT, "hite_b200")):
        for f in fs:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dp, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f
                assert not re.search(r'#include\s+"[^"]*oracle', src) and "hbo.h" not in src, f


def test_fasta_contract(tmp_path):
    from hite_b200 import fasta
    p = tmp_path / "a.fa"
    p.write_text(">chr1 some desc\nacgt\nNNac\n>chr2\tx\nGG\n>empty\n>chr3\nTT\n")
    names, contigs = fasta.read_fasta(str(p))
    assert names == ["chr1", "chr2", "chr3"]
    assert contigs == {"chr1": "ACGTNNAC", "chr2": "GG", "chr3": "TT"}
    q = tmp_path / "b.fa"
    fasta.store_fasta(contigs, str(q))
    assert q.read_text() == ">chr1\nACGTNNAC\n>chr2\nGG\n>chr3\nTT\n"
    assert fasta.file_exist(str(q)) and not fasta.file_exist(str(tmp_path / "nope.fa"))
    ref = "/root/reference/module/Util.py"
    if os.path.exists(ref):
        from oracle import ref_import
        U = ref_import.load_util()
        assert U.read_fasta(str(p)) == (names, contigs)
        r = tmp_path / "c.fa"
        U.store_fasta(contigs, str(r))
        assert r.read_bytes() == q.read_bytes()


def test_post_filter_rules():
    from hite_b200 import boundary
    names = ["a", "b", "c", "d"]
    res = [dict(is_te=True, cons="TGAAAACA", copy_num=30, info=""), dict(is_te=True, cons="ACGT", copy_num=5, info=""),
           dict(is_te=True, cons="ACGTT", copy_num=6, info=""), dict(is_te=False, cons="", copy_num=0, info="nb")]
    true_tes, low = boundary._post_filter(names, None, res, "tir")
    assert true_tes == {"c": "ACGTT"} and low == {"b": "ACGT"}          # Util.py:8175, 8179-8188
    true_tes, low = boundary._post_filter(names, None, res, "helitron")
    assert set(true_tes) == {"b", "c"} and low == {}


def test_shard_indices_deterministic_partition():
    from hite_b200 import dist as hd
    lens = [5, 9, 1, 7, 7, 3, 8]
    parts = [hd.shard_indices(len(lens), r, 3, lens) for r in range(3)]
    assert sorted(sum(parts, [])) == list(range(len(lens)))
    assert parts == [hd.shard_indices(len(lens), r, 3, lens) for r in range(3)]
    assert hd.shard_indices(5, 1, 2) == [1, 3]
    # cost deal (heaviest first, to the least loaded rank): a partition, deterministic, and balanced within one item
    cost = [900, 20, 30, 880, 10, 15, 870, 5, 300, 310, 290, 1]
    parts = [hd.shard_by_cost(cost, r, 3) for r in range(3)]
    assert sorted(sum(parts, [])) == list(range(len(cost))) and parts == [hd.shard_by_cost(cost, r, 3) for r in range(3)]
    loads = [sum(cost[i] for i in p_) for p_ in parts]
    assert max(loads) - min(loads) <= 30, loads
    assert hd.shard_by_cost([], 0, 2) == [] and hd.shard_by_cost([7], 1, 2) == []


GLOO_WORKER = r"""
import os, sys
sys.path.insert(0, %r)
import torch.distributed as dist
from hite_b200 import dist as hd
dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
names = ["te%%d" %% i for i in range(11)]
seqs = ["ACGT" * (i + 1) for i in range(11)]
mine = hd.shard_indices(len(seqs), rank, world, [len(s) for s in seqs])
recs = [(names[i], seqs[i]) for i in mine if i %% 3 != 0]          # "accepted" subset
merged = hd.allgather_records(recs)
exp = []
for r in range(world):
    exp += [(names[i], seqs[i]) for i in hd.shard_indices(len(seqs), r, world, [len(s) for s in seqs]) if i %% 3 != 0]
assert merged == exp, (merged, exp)
empty = hd.allgather_records([] if rank == 0 else [("x", "A")])
assert empty == [("x", "A")] * (world - 1)
class FakeCtx:                                                        # sharded_adjust: shard -> adjust -> all-gather of the accepted
    def adjust(self, seqs, **kw):
        return [dict(is_te=len(s) %% 8 != 0, cons=s.lower()) for s in seqs]
got = hd.sharded_adjust(FakeCtx(), names, seqs)
exp = []
for r in range(world):
    exp += [(names[i], seqs[i].lower()) for i in hd.shard_indices(len(seqs), r, world, [len(s) for s in seqs]) if len(seqs[i]) %% 8 != 0]
assert got == exp, (got, exp)
# the fused exchange bench.py and the stages use: one fixed-capacity all_gather_into_tensor, sizes in the header
import numpy as np, torch
shards = [hd.shard_indices(len(seqs), r, world, [len(s) for s in seqs]) for r in range(world)]
g = hd.RecordGather([len(s) for s in shards], [sum(len(seqs[i]) for i in s) for s in shards], "cpu")
for step in range(3):
    keep = [i for i in shards[rank] if (i + step) %% 3 != 0]
    cons = "".join(seqs[i].lower() for i in keep).encode()
    g.start(np.array(keep, dtype=np.int32), np.array([len(seqs[i]) for i in keep], dtype=np.int32),
            torch.frombuffer(bytearray(cons), dtype=torch.uint8) if cons else None)
    got = g.finish()
    if rank == 0:
        idx, ln, body = got
        exp_keep = [i for r in range(world) for i in shards[r] if (i + step) %% 3 != 0]
        assert idx.tolist() == exp_keep and ln.tolist() == [len(seqs[i]) for i in exp_keep]
        assert body.tobytes().decode() == "".join(seqs[i].lower() for i in exp_keep)
    else:
        assert got is None
    if step == 2:                                                     # zero-copy read-back: per-rank views, same content
        g.start(np.array(keep, dtype=np.int32), np.array([len(seqs[i]) for i in keep], dtype=np.int32),
                torch.frombuffer(bytearray(cons), dtype=torch.uint8) if cons else None)
        g.finish_async()                                              # same, finished on the helper thread
        got = g.result()
        if rank == 0:
            assert len(got[0]) == world and np.concatenate(got[0]).tolist() == exp_keep
            assert b"".join(c.tobytes() for c in got[2]).decode() == "".join(seqs[i].lower() for i in exp_keep)
try:
    g.start(np.arange(99, dtype=np.int32), np.ones(99, dtype=np.int32), None)
    raise SystemExit("capacity overflow not detected")
except RuntimeError:
    pass
dist.destroy_process_group()
print("rank", rank, "ok")
"""


def test_allgather_records_gloo_world2(tmp_path):
    script = tmp_path / "w.py"
    script.write_text(GLOO_WORKER % ROOT)
    env = dict(os.environ, MASTER_ADDR="127.0.0.1", MASTER_PORT="29531")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                          "--master-addr", "127.0.0.1", "--master-port", "29531", str(script)],
                         capture_output=True, text=True, env=env, timeout=240)
    assert out.returncode == 0, out.stdout + out.stderr
    assert out.stdout.count("ok") == 2


def test_product_fmea_matches_reference_goldens():
    """hb_fmea_chain (host C++ inside libhite_b200.so) vs goldens from the imported reference
    get_longest_repeats_v4 (tests/golden/make_golden.py): bit-identical interval lists, same order."""
    import gzip
    import json
    import numpy as np
    from hite_b200 import _lib
    recs = json.loads(gzip.open(os.path.join(ROOT, "tests", "golden", "fmea.json.gz")).read())
    for rec in recs:
        chr_ids = {}
        for name in [rec["query"][0]] + [s[0] for s in rec["subjects"]]:
            chr_ids.setdefault(name, len(chr_ids))
        subjects = [tuple(s) for s in rec["subjects"]]
        q = tuple(rec["query"])
        # the query must be one of the segments: index self_subj if present, else appended
        segs = list(subjects)
        qi = rec["self_subj"]
        if qi < 0:
            segs.append(q)
            qi = len(segs) - 1
        seg_arr = np.zeros(len(segs), dtype=_lib.SEG_DTYPE)
        for i, (c, o) in enumerate(segs):
            seg_arr[i] = (chr_ids[c], 0, o)
        hs = np.zeros(len(rec["hsps"]), dtype=_lib.HSP_DTYPE)
        for i, h in enumerate(rec["hsps"]):
            hs[i] = tuple(h)
        q_off = np.zeros(len(segs) + 1, dtype=np.uint64)
        q_off[qi + 1:] = len(hs)
        out = _lib.fmea_chain(hs, q_off, seg_arr, rec["skip_gap"], rec["max_len"], threads=2)
        names = ["%s:%d-%d" % (q[0], int(o["start"]), int(o["end"])) for o in out]
        assert names == rec["expect"], rec["seed"]
        assert all(int(o["qseg"]) == qi for o in out)


def test_flanking_seq_matches_reference(tmp_path):
    ref_util = "/root/reference/module/Util.py"
    from hite_b200 import coarse
    ref = tmp_path / "g.fa"
    ref.write_text(">c1\n" + "ACGT" * 100 + "\n>c2\n" + "TTGCA" * 50 + "\n")
    lr = tmp_path / "lr.fa"
    lr.write_text(">c1:10-120\nA\n>c1:0-90\nA\n>c1:300-399\nA\n>c2:100-250\nA\n")
    out = tmp_path / "b.fa"
    coarse.flanking_seq(str(lr), str(out), str(ref), 50)
    text = out.read_text()
    assert text.startswith(">c1:")            # names: chr:(s+1-50)-(e+50), clamped (Util.py:4620-4633)
    assert ">c1:1-170\n" in text and ">c1:251-400\n" in text
    if os.path.exists(ref_util):
        from oracle import ref_import
        U = ref_import.load_util()
        exp = tmp_path / "a.fa"
        U.flanking_seq(str(lr), str(exp), str(ref), 50)
        assert exp.read_text() == text


def test_product_query_copies_matches_reference_goldens(tmp_path):
    """hb_query_copies / boundary.get_query_copies vs goldens from the imported reference get_query_copies."""
    import gzip
    import json
    from hite_b200 import boundary
    recs = json.loads(gzip.open(os.path.join(ROOT, "tests", "golden", "query_copies.json.gz")).read())
    assert len(recs) >= 150
    for rec in recs:
        subject_dict = {}
        for (si, qs, qe, ss, se), idn in zip(rec["hsps"], rec["idents"]):
            subject_dict.setdefault("s%d" % si, []).append((qs, qe, ss, se, idn))
        subj = tmp_path / "s.fa"
        subj.write_text("".join(">s%d\n%s\n" % (i, "A" * ln) for i, ln in enumerate(rec["subj_lens"])))
        res = boundary.get_query_copies([("q", subject_dict)], {"q": "A" * rec["qlen"]}, str(subj), rec["query_coverage"],
                                        rec["subject_coverage"], max_copy_num=rec["max_copy_num"])
        got = [[int(c[0][1:]), c[1], c[2], c[3], c[4]] for c in res["q"]]
        assert got == rec["expect"], rec["seed"]


def test_drop_in_stage_scripts_patch_the_reference_namespace(monkeypatch, tmp_path):
    """module/_hite_b200_patch.run_stage rebinds exactly the hot-path names in the reference's Util and then
    runs the reference's own stage script (runpy), so main.py's call is unchanged."""
    ref = "/root/reference"
    if not os.path.exists(os.path.join(ref, "module", "Util.py")):
        pytest.skip("reference checkout not mounted")
    from oracle import ref_import
    U = ref_import.load_util()                         # Util with the missing third-party imports stubbed
    sys.path.insert(0, os.path.join(ROOT, "module"))
    import _hite_b200_patch as patch
    import runpy
    ran = {}
    monkeypatch.setenv("HITE_HOME", ref)
    monkeypatch.setattr(runpy, "run_path", lambda path, run_name=None: ran.setdefault("path", path))
    saved = {n: getattr(U, n) for n in ("flank_region_align_v5", "run_find_members_v8", "is_TE_from_align_file",
                                        "determine_repeat_boundary_v5")}
    try:
        patch.run_stage("judge_TIR_transposons.py")
        import hite_b200.boundary as hb
        assert U.flank_region_align_v5 is hb.flank_region_align_v5
        assert U.run_find_members_v8 is hb.run_find_members_v8
        assert ran["path"].endswith("module/judge_TIR_transposons.py") and ran["path"].startswith(ref)
        patch.run_stage("coarse_boundary.py", patch_names=("determine_repeat_boundary_v5",), source="hite_b200.coarse")
        import hite_b200.coarse as hc
        assert U.determine_repeat_boundary_v5 is hc.determine_repeat_boundary_v5
        saved["multi_process_align_and_get_copies"] = U.multi_process_align_and_get_copies
        patch.run_stage("judge_Other_transposons.py", patch_names=("multi_process_align_and_get_copies",), source="hite_b200.other")
        import hite_b200.other as ho
        assert U.multi_process_align_and_get_copies is ho.multi_process_align_and_get_copies
        assert ran["path"].endswith("module/judge_TIR_transposons.py")      # setdefault keeps the first; the call itself ran
        saved["get_candidate_non_ltr_parallel"] = U.get_candidate_non_ltr_parallel
        saved["mask_genome_with_TE"] = U.mask_genome_with_TE
        patch.run_stage("judge_Non_LTR_transposons.py", extra=(("get_candidate_non_ltr_parallel", "hite_b200.nonltr_prefilter"),))
        import hite_b200.nonltr_prefilter as hn
        assert U.get_candidate_non_ltr_parallel is hn.get_candidate_non_ltr_parallel
        saved["multi_process_tsd_v1"] = U.multi_process_tsd_v1
        patch.run_stage("judge_TIR_transposons.py", extra=(("multi_process_tsd_v1", "hite_b200.tir_prefilter"),))
        import hite_b200.tir_prefilter as ht
        assert U.multi_process_tsd_v1 is ht.multi_process_tsd_v1
        patch.run_stage("judge_LTR_transposons.py", patch_names=("mask_genome_with_TE",), source="hite_b200.masking")
        import hite_b200.masking as hm
        assert U.mask_genome_with_TE is hm.mask_genome_with_TE
    finally:
        for n, f in saved.items():
            setattr(U, n, f)
    for s in ("coarse_boundary.py", "judge_TIR_transposons.py", "judge_Helitron_transposons.py", "judge_Non_LTR_transposons.py",
              "judge_Other_transposons.py", "judge_LTR_transposons.py"):
        assert os.access(os.path.join(ROOT, "module", s), os.X_OK)


def test_stage_process_shards_candidates_over_gpus(monkeypatch, tmp_path):
    """hite_b200.boundary.adjust_on_devices / flank_region_align_v5: one context and host thread per GPU of HITE_GPU_DEVICES,
    longest-first deal, results back in input order and independent of the GPU count (fake contexts: no GPU here)."""
    import threading
    from hite_b200 import boundary, fasta
    made, calls = [], []
    sync = {"barrier": None}

    class FakeCtx:
        def __init__(self, device):
            self.device = device
            self.genome = None
            made.append(device)

        def load_genome(self, contigs, names=None):
            self.genome = [bytes(c) for c in contigs]

        def build_index(self, k, w):
            assert self.genome is not None

        def adjust(self, seqs, te_type, plant, result_type, flank=50):
            if sync["barrier"] is not None:
                sync["barrier"].wait(timeout=20)          # every GPU's call is in flight at the same time: one thread each
            calls.append((self.device, threading.get_ident(), list(seqs)))
            return [dict(is_te=len(s) % 3 != 0, info="", cons=s[2:-2], copy_num=7 + len(s) % 5) for s in seqs]

    monkeypatch.setattr(boundary._lib, "Context", FakeCtx)
    monkeypatch.setattr(boundary, "_CTX", {})
    monkeypatch.setattr(boundary, "_GENOME", {})
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    ref = tmp_path / "genome.fa"
    ref.write_text(">chr1\nACGTACGTAC\n>chr2\nGGGGCCCC\n")
    cand = tmp_path / "cand.fa"
    seqs = {"c%d" % i: "ACGT" * (3 + (i * 7) % 11) + "A" * (i % 4) for i in range(23)}
    fasta.store_fasta(seqs, str(cand))
    outs = {}
    for devs in ("0", "0,1,2", "3,1"):
        monkeypatch.setenv("HITE_GPU_DEVICES", devs)
        made.clear(); calls.clear()
        sync["barrier"] = threading.Barrier(len(devs.split(",")))
        monkeypatch.setattr(boundary, "_CTX", {})
        monkeypatch.setattr(boundary, "_GENOME", {})
        assert boundary.stage_devices() == [int(x) for x in devs.split(",")]
        real, low = str(tmp_path / ("real_%s.fa" % devs)), str(tmp_path / ("low_%s.fa" % devs))
        res = boundary.flank_region_align_v5(str(cand), real, 50, str(ref), "", "tir", str(tmp_path), 8, 0, None, "", 1, 0, 0, low)
        outs[devs] = (res, open(real).read(), open(low).read())
        assert sorted(made) == sorted(int(x) for x in devs.split(","))
        assert sorted(d for d, _, _ in calls) == sorted(made)                    # one adjust call per GPU
        if len(made) > 1:
            assert len({t for _, t, _ in calls}) == len(made)                    # each from its own host thread
            sizes = sorted(sum(len(s) for s in sq) for _, _, sq in calls)
            assert sizes[-1] - sizes[0] <= max(len(s) for s in seqs.values())    # longest-first deal balances the bases
        assert sorted(s for _, _, sq in calls for s in sq) == sorted(seqs.values())
    assert outs["0"] == outs["0,1,2"] == outs["3,1"]
    monkeypatch.setenv("WORLD_SIZE", "4"); monkeypatch.setenv("LOCAL_RANK", "1"); monkeypatch.setenv("HITE_GPU_DEVICES", "4,5,6,7")
    assert boundary.stage_devices() == [5]                                       # under torchrun a rank owns one GPU

    sync["barrier"] = None

    class Boom(FakeCtx):
        def adjust(self, *a, **k):
            if self.device == 2:
                raise RuntimeError("device 2 failed")
            return FakeCtx.adjust(self, *a, **k)
    monkeypatch.delenv("WORLD_SIZE"); monkeypatch.setenv("HITE_GPU_DEVICES", "0,1,2")
    monkeypatch.setattr(boundary._lib, "Context", Boom)
    monkeypatch.setattr(boundary, "_CTX", {}); monkeypatch.setattr(boundary, "_GENOME", {})
    with pytest.raises(RuntimeError):                                            # a failing GPU fails the stage loudly
        boundary.adjust_on_devices(list(seqs.values()), str(ref), "tir", 1, "cons", 50)

    # HB_ENOMEM (too many seed hits / MSA columns for one call): the batch is halved and retried, same results in the same order
    sizes = []

    class Tight(FakeCtx):
        def adjust(self, sq, *a, **k):
            sizes.append(len(sq))
            if len(sq) > 4:
                raise boundary._lib.HiteB200Error("hb_adjust failed (-4): hb_adjust: too many seed hits in one call; pass fewer candidates")
            return FakeCtx.adjust(self, sq, *a, **k)
    monkeypatch.setenv("HITE_GPU_DEVICES", "0")
    monkeypatch.setattr(boundary._lib, "Context", Tight)
    monkeypatch.setattr(boundary, "_CTX", {}); monkeypatch.setattr(boundary, "_GENOME", {})
    got = boundary.adjust_on_devices(list(seqs.values()), str(ref), "tir", 1, "cons", 50)
    assert got == outs["0"][0] and max(sizes) == 23 and sizes.count(23) == 1 and all(n <= 4 for n in sizes if n not in (23, 11, 12, 5, 6))


def test_other_and_masking_host_logic_with_fake_context():
    """hite_b200.other / hite_b200.masking above the C-ABI: record arrays -> reference dict shapes, self hits dropped
    (module/Util.py:7045), '#' cut and strand dropped before the full-length chaining (:5913-5924, :6351-6354), empty inputs."""
    import numpy as np
    from hite_b200 import _lib, masking, other

    class Fake:
        contig_names = ["chr1", "chr2"]
        recs = np.zeros(0, dtype=_lib.QHSP_DTYPE)

        def query_hsps(self, seqs, occ_cap=512, max_loci=128):
            self.n = len(seqs)
            return self.recs

        def timing(self):
            return dict(kernel_ms=0.0, k_ms=[0] * 8)

    f = Fake()
    rec = other.query_hsp_records(f, [], {})
    assert len(rec) == 0 and other.copies_from_records(rec, [], {}, f.contig_names) == {}
    assert masking.intervals_from_hsp_records(rec, [], {}) == []
    contigs = {"a": "AC", "b": "ACGT" * 10}
    rec = other.query_hsp_records(f, ["a", "b"], contigs)          # "a" is shorter than a seed: never sent
    assert f.n == 1 and len(rec) == 0 and other.query_hsps(f, ["b"], contigs) == {}
    r = np.zeros(3, dtype=_lib.QHSP_DTYPE)
    r["query"] = [0, 0, 1]; r["contig"] = [0, 1, 0]; r["qs"] = 1; r["qe"] = 40
    r["ss"] = [101, 240, 501]; r["se"] = [140, 201, 540]; r["n_match"] = [40, 39, 40]; r["aln_len"] = 40
    f.recs = r
    names = ["x#DNA/TIR", "chr1"]
    contigs = {"x#DNA/TIR": "ACGT" * 10, "chr1": "ACGT" * 10}
    rec = other.query_hsp_records(f, names, contigs)
    assert [tuple(x) for x in rec.tolist()] == [(0, 0, 1, 40, 101, 140, 100.0), (0, 1, 1, 40, 240, 201, 97.5)]   # self hit gone
    assert other.query_hsps(f, names, contigs) == {"x#DNA/TIR": {"chr1": [(1, 40, 101, 140, 100.0)], "chr2": [(1, 40, 240, 201, 97.5)]}}
    cps = other.copies_from_records(rec, names, contigs, f.contig_names)
    assert [c[:3] + c[4:] for c in cps["x#DNA/TIR"]] == [("chr1", 101, 140, "+"), ("chr2", 201, 240, "-")]
    assert masking.intervals_from_hsp_records(rec, names, contigs) == [("x", 0, 100, 140), ("x", 1, 200, 240)]
    # a library entry beyond the kernel's query limit is tiled and its coordinates shifted back
    big = {"big": "A" * 70_000}
    r = np.zeros(3, dtype=_lib.QHSP_DTYPE)
    r["query"] = [0, 1, 2]; r["qs"] = 1; r["qe"] = [32000, 32000, 6000]; r["ss"] = [1, 32001, 64001]; r["se"] = [32000, 64000, 70000]
    r["n_match"] = r["aln_len"] = [32000, 32000, 6000]
    f.recs = r
    rec = other.query_hsp_records(f, ["big"], big)
    assert f.n == 3 and [(int(x["qs"]), int(x["qe"])) for x in rec] == [(1, 32000), (32001, 64000), (64001, 70000)]
    assert [c[:3] for c in other.copies_from_records(rec, ["big"], big, f.contig_names)["big"]] == [("chr1", 1, 70000)]


def test_genomes_past_the_32_bit_position_space_are_refused_early():
    """2^32 bp and more cannot be indexed (32-bit positions in the keys); the wrapper says so before it concatenates and
    ships the bases -- no device, no library call involved."""
    import numpy as np
    import pytest as _pytest
    import hite_b200
    from hite_b200 import _lib

    class Big:                                   # stands for a 1.5 Gbp contig without holding it
        def __init__(self, n): self.n = n
        def __len__(self): return self.n

    ctx = object.__new__(_lib.Context)
    real_asarray = np.asarray
    try:
        np.asarray = lambda x, dtype=None: x if isinstance(x, Big) else real_asarray(x, dtype=dtype)
        with _pytest.raises(hite_b200.HiteB200Error, match="32-bit positions"):
            _lib.Context.load_genome(ctx, [Big(1_500_000_000)] * 3)
    finally:
        np.asarray = real_asarray


def test_product_fmea_shares_state_inside_a_blast_file():
    """A contig's short tail segment shares its BLAST result file -- and FMEA's de-duplication dicts -- with the records that
    follow it (reference split_and_store_sequences module/Util.py:4987, chr_pos_candidates :4155).  Product host code
    (hb_fmea_chain_files) == oracle on a genome with 600 kbp tails; the per-record form differs there, as it must."""
    import numpy as np
    from hite_b200 import _lib
    from workloads import synth
    from oracle import oracle as O
    wl = synth.make_workload(genome_bp=3_200_000, n_contigs=2, n_families=25, cands_per_family=1, copies=(5, 14),
                             length=(200, 1500), seed=31)
    lens = [len(c) for c in wl.contig_seqs]
    G = O.Genome(wl.contig_seqs)
    hsps, segs, iv = G.coarse_boundary(lens, seg_len=1_000_000, skip_gap=2000, max_len=30000, occ_cap=64)
    G.close()
    files = O.split_file_groups([min(1_000_000, lens[c] - o) for c, o in segs])
    assert files == [0, 1, 1, 2]                                   # chr1 tail + chr2 head share a file
    h = np.asarray(hsps)
    h = h[np.argsort(h[:, 0], kind="stable")]
    q_off = np.searchsorted(h[:, 0], np.arange(len(segs) + 1)).astype(np.uint64)
    hs = np.zeros(len(h), dtype=_lib.HSP_DTYPE)
    for k, f in enumerate(("subj", "qs", "qe", "ss", "se")):
        hs[f] = h[:, k + 1]
    sg = np.zeros(len(segs), dtype=_lib.SEG_DTYPE)
    sg["chr"] = [s[0] for s in segs]; sg["start"] = [s[1] for s in segs]
    for threads in (1, 4):
        out = _lib.fmea_chain_files(hs, q_off, sg, files, 2000, 30000, threads=threads)
        assert [(int(o["chr"]), int(o["start"]), int(o["end"])) for o in out] == iv
    solo = _lib.fmea_chain(hs, q_off, sg, 2000, 30000, threads=2)
    assert [(int(o["chr"]), int(o["start"]), int(o["end"])) for o in solo] != iv
    same = _lib.fmea_chain_files(hs, q_off, sg, list(range(len(segs))), 2000, 30000, threads=2)
    assert [tuple(x) for x in same.tolist()] == [tuple(x) for x in solo.tolist()]


def test_product_fmea_task_graph_is_independent_of_blocks_and_threads():
    """The chaining runs as a task graph (group query -> chain (query, subject) pairs -> de-duplicate file) on one pool of host
    threads and is fed block by block by the coarse stage's GPU pipeline.  Whatever the number of submissions and of threads --
    including more threads than files and more blocks than files -- the intervals are the oracle's (hbo_fmea.c driven per file),
    on a genome with shared files (contig tails) and several files per block."""
    import numpy as np
    from hite_b200 import _lib
    from workloads import synth
    from oracle import oracle as O
    wl = synth.make_workload(genome_bp=7_500_000, n_contigs=3, n_families=30, cands_per_family=1, copies=(5, 16),
                             length=(200, 1500), seed=41)
    lens = [len(c) for c in wl.contig_seqs]
    G = O.Genome(wl.contig_seqs)
    hsps, segs, iv = G.coarse_boundary(lens, seg_len=1_000_000, skip_gap=2000, max_len=30000, occ_cap=64)
    G.close()
    files = O.split_file_groups([min(1_000_000, lens[c] - o) for c, o in segs])
    assert len(set(files)) < len(segs) and len(iv) > 100          # tails share files; there is something to chain
    h = np.asarray(hsps)
    h = h[np.argsort(h[:, 0], kind="stable")]
    q_off = np.searchsorted(h[:, 0], np.arange(len(segs) + 1)).astype(np.uint64)
    hs = np.zeros(len(h), dtype=_lib.HSP_DTYPE)
    for k, f in enumerate(("subj", "qs", "qe", "ss", "se")):
        hs[f] = h[:, k + 1]
    sg = np.zeros(len(segs), dtype=_lib.SEG_DTYPE)
    sg["chr"] = [s[0] for s in segs]; sg["start"] = [s[1] for s in segs]
    for n_blocks, threads in ((1, 1), (1, 7), (2, 3), (3, 16), (len(set(files)), 5), (64, 2)):
        for _ in range(3):                                         # scheduling differs from run to run; the result must not
            out = _lib.debug_fmea_blocks(hs, q_off, sg, files, 2000, 30000, threads, n_blocks)
            assert [(int(o["chr"]), int(o["start"]), int(o["end"])) for o in out] == iv, (n_blocks, threads)
    # no HSPs at all: nothing to do, nothing hangs
    none = _lib.debug_fmea_blocks(hs[:0], np.zeros(len(segs) + 1, dtype=np.uint64), sg, files, 2000, 30000, 4, 3)
    assert len(none) == 0


def test_product_fmea_long_segments_take_the_comparison_sorts():
    """The chaining orders fragments through packed 64-bit keys while segment coordinates fit 21 bits (the default 1 Mbp
    segments do) and through comparison sorts beyond that.  `chrom_seg_length` is a user option (main.py:41): with 3 Mbp
    segments most coordinates are past 2^21 and the result must still be the oracle's."""
    import numpy as np
    from hite_b200 import _lib
    from workloads import synth
    from oracle import oracle as O
    wl = synth.make_workload(genome_bp=7_500_000, n_contigs=2, n_families=30, cands_per_family=1, copies=(5, 16),
                             length=(200, 1500), seed=43)
    lens = [len(c) for c in wl.contig_seqs]
    G = O.Genome(wl.contig_seqs)
    hsps, segs, iv = G.coarse_boundary(lens, seg_len=3_000_000, skip_gap=2000, max_len=30000, occ_cap=64)
    G.close()
    h = np.asarray(hsps)
    assert (h[:, 2:] >= (1 << 21)).any() and len(iv) > 100          # coordinates past the packed-key range, work to do
    files = O.split_file_groups([min(3_000_000, lens[c] - o) for c, o in segs])
    h = h[np.argsort(h[:, 0], kind="stable")]
    q_off = np.searchsorted(h[:, 0], np.arange(len(segs) + 1)).astype(np.uint64)
    hs = np.zeros(len(h), dtype=_lib.HSP_DTYPE)
    for k, f in enumerate(("subj", "qs", "qe", "ss", "se")):
        hs[f] = h[:, k + 1]
    sg = np.zeros(len(segs), dtype=_lib.SEG_DTYPE)
    sg["chr"] = [s[0] for s in segs]; sg["start"] = [s[1] for s in segs]
    for threads in (1, 5):
        out = _lib.fmea_chain_files(hs, q_off, sg, files, 2000, 30000, threads=threads)
        assert [(int(o["chr"]), int(o["start"]), int(o["end"])) for o in out] == iv


def test_coarse_masking_leaves_border_crossing_copies(monkeypatch):
    """masking.mask_context(record_len=...): a full-length copy lying across a 1 Mbp record border is not masked, as in the
    reference, where the two halves are partial hits on two subject records (module/Util.py:4654 on chr$start records)."""
    import numpy as np
    from hite_b200 import _lib, masking

    class Fake:
        contig_names = ["chr1"]
        masked = None

        def query_hsps(self, seqs, occ_cap=512, max_loci=128):
            r = np.zeros(3, dtype=_lib.QHSP_DTYPE)
            r["query"] = 0; r["contig"] = 0; r["qs"] = 1; r["qe"] = 1000; r["n_match"] = r["aln_len"] = 1000
            r["ss"] = [5001, 999_501, 1_999_001]; r["se"] = [6000, 1_000_500, 2_000_000]
            return r

        def timing(self):
            return dict(kernel_ms=0.0, k_ms=[0] * 8)

        def mask_genome(self, contig, start, end):
            self.masked = sorted(zip(contig, start, end))
            return sum(b - a for _, a, b in self.masked)

        def build_index(self, k, w):
            return dict(kernel_ms=0.0)

    f = Fake()
    iv, fresh = masking.mask_context(f, ["te1"], {"te1": "A" * 1000}, record_len=1_000_000)
    assert f.masked == [(0, 5000, 6000), (0, 1_999_000, 2_000_000)] and fresh == 2000     # the copy across 1 000 000 stays
    iv, fresh = masking.mask_context(f, ["te1"], {"te1": "A" * 1000})
    assert len(f.masked) == 3


def test_smoke_next_rows_block_dry_run():
    """__graft_entry__._smoke_next_rows (the part of smoke() that exercises the section-8f kernels on the GPU box) run on the CPU
    against a stand-in context backed by the oracle: catches a typo in the checker code before the box does."""
    import numpy as np
    import __graft_entry__ as entry
    from hite_b200 import _lib
    from workloads import synth
    from oracle import oracle as O
    wl = synth.make_workload(genome_bp=400_000, n_contigs=2, n_families=4, cands_per_family=2, copies=(6, 10),
                             length=(200, 1200), seed=7)
    G = O.Genome(wl.contig_seqs)

    class OracleBackedCtx:
        def query_hsps(self, seqs, occ_cap=512, max_loci=128):
            rows = [(qi, h["contig"], h["qs"], h["qe"], h["ss"], h["se"], h["n_match"], h["aln_len"], h["score"], h["rank"])
                    for qi, q in enumerate(seqs) for h in G.query_hsps(q)]
            return np.array(rows, dtype=_lib.QHSP_DTYPE)

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

        def nonltr_candidates(self, seqs, flanking_len=50):
            out = np.zeros(len(seqs), dtype=_lib.POLYA_DTYPE)
            for i, s in enumerate(seqs):
                for k, v in O.search_polyA_TSD(s, flanking_len)[3].items():
                    out[i][k] = v
            return out
    entry._smoke_next_rows(OracleBackedCtx(), G, wl)
    G.close()
