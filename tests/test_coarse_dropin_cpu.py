"""The coarse-boundary stage end to end on the CPU, reference against mirror, on the same HSPs.

The REFERENCE's own determine_repeat_boundary_v5 (module/Util.py:4637-4669, imported unmodified) runs its segment files,
process pools, get_longest_repeats_v4 (FMEA) and generate_final_result.  Only its externals are replaced: the trf tandem filter
passes the chunk through, and the per-pair blastn (sequence2sequenceBlastn, :4068) writes the oracle's HSPs of that query file
in outfmt 6.  The MIRROR hite_b200.coarse.determine_repeat_boundary_v5 runs with a stand-in context whose intervals are
the oracle's HSPs chained by the product's host FMEA.  Both must write the same longest_repeats records ('chr:start-end' names and sequences): this pins the segment
naming (chr$start), the coordinate conventions between HSPs and interval names and the final slicing on the reference's own
code, and runs the reference's FMEA itself on a realistic HSP set.  Needs the reference checkout (skipped on the GPU box)."""
import os

import numpy as np
import pytest

from workloads import synth
from oracle import oracle as O
from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference checkout not mounted")


def _fake_blastn(query_path, target_path):
    """Stand-in for sequence2sequenceBlastn (module/Util.py:4068): module level, because the reference submits it to a process
    pool.  Every HSP of the query file is written once (for the first target file), in the oracle's order."""
    import json
    out = "%s_vs_%s.out" % (query_path, os.path.basename(target_path))
    with open(out, "w") as f:
        if os.path.basename(target_path) == "0_target.fa":
            by_q = json.load(open(os.environ["HB_FAKE_BLAST_DB"]))
            for line in open(query_path):
                if line.startswith(">"):
                    f.writelines(by_q.get(line[1:].strip(), []))
    return out


@pytest.mark.parametrize("genome_bp,n_contigs,seed,skip_gap", [(3_200_000, 2, 31, 2000), (3_900_000, 3, 57, 4000)])
def test_reference_coarse_stage_equals_mirror_on_the_same_hsps(tmp_path, monkeypatch, genome_bp, n_contigs, seed, skip_gap):
    from hite_b200 import _lib, boundary, coarse, fasta
    U = ref_import.load_util()
    wl = synth.make_workload(genome_bp=genome_bp, n_contigs=n_contigs, n_families=25, cands_per_family=1, copies=(5, 14),
                             length=(200, 1500), seed=seed)
    ref = str(tmp_path / "genome.fa")
    wl.write_fasta(ref, str(tmp_path / "cand.fa"))
    lens = [len(c) for c in wl.contig_seqs]
    G = O.Genome(wl.contig_seqs)
    hsps, segs, iv = G.coarse_boundary(lens, seg_len=1_000_000, skip_gap=skip_gap, max_len=30000, occ_cap=64)
    G.close()
    assert len(hsps) > 2000 and len(iv) > 100
    seg_name = ["%s$%d" % (wl.contig_names[c], o) for c, o in segs]
    cut = str(tmp_path / "genome.cut0.fa")                        # split_genome_chunks.py:44-52 layout
    with open(cut, "w") as f:
        for (c, o), n in zip(segs, seg_name):
            f.write(">%s\n%s\n" % (n, wl.contig_seqs[c][o:o + 1_000_000].tobytes().decode()))

    # ---- the reference, externals replaced --------------------------------------------------------------------
    def pass_through(repeat_names, repeat_contigs, tmp_output_dir, ref_index, threads):
        fasta.store_fasta({n: repeat_contigs[n] for n in repeat_names}, os.path.join(tmp_output_dir, "filter_tandem_%s.fa" % ref_index))

    by_q = {}
    for q, s, qs, qe, ss, se in hsps.tolist():
        by_q.setdefault(seg_name[q], []).append("%s\t%s\t100.0\t%d\t0\t0\t%d\t%d\t%d\t%d\t0.0\t100\n"
                                                % (seg_name[q], seg_name[s], qe - qs + 1, qs, qe, ss, se))

    import json
    dbpath = str(tmp_path / "hsp_db.json")
    json.dump(by_q, open(dbpath, "w"))
    monkeypatch.setenv("HB_FAKE_BLAST_DB", dbpath)
    monkeypatch.setattr(U, "filter_tandem_repeats", pass_through)
    monkeypatch.setattr(U, "sequence2sequenceBlastn", _fake_blastn)
    out_ref = str(tmp_path / "ref_longest_repeats_0.fa")
    U.determine_repeat_boundary_v5(cut, out_ref, None, skip_gap, 30000, str(tmp_path / "refwork"), 4, 0, ref, 1)

    # ---- the mirror, intervals from the oracle ----------------------------------------------------------------
    class OracleBackedCtx:
        device = 0

        def load_genome(self, chroms):
            assert [len(c) for c in chroms] == lens

        def build_index(self, k, w):
            return dict(kernel_ms=0.0)

        def coarse_boundary(self, seg_len, skip_gap, max_len, occ_cap, threads):
            # the HSPs are the oracle's; the chaining is the PRODUCT's host code (hb_fmea_chain_files), files as
            # hb_coarse_boundary derives them from the segment lengths
            assert (seg_len, max_len) == (1_000_000, 30000)
            h = np.asarray(hsps)
            h = h[np.argsort(h[:, 0], kind="stable")]
            q_off = np.searchsorted(h[:, 0], np.arange(len(segs) + 1)).astype(np.uint64)
            hs = np.zeros(len(h), dtype=_lib.HSP_DTYPE)
            for k, f in enumerate(("subj", "qs", "qe", "ss", "se")):
                hs[f] = h[:, k + 1]
            sg = np.zeros(len(segs), dtype=_lib.SEG_DTYPE)
            sg["chr"] = [x[0] for x in segs]; sg["start"] = [x[1] for x in segs]
            files = O.split_file_groups([min(seg_len, lens[c] - o) for c, o in segs])
            out = _lib.fmea_chain_files(hs, q_off, sg, files, skip_gap, max_len, threads=threads)
            assert [(int(o["chr"]), int(o["start"]), int(o["end"])) for o in out] == iv
            return out

    monkeypatch.setattr(coarse, "get_context", lambda device=None, role="reference": OracleBackedCtx())
    out_mine = str(tmp_path / "mine_longest_repeats_0.fa")
    coarse.determine_repeat_boundary_v5(cut, out_mine, None, skip_gap, 30000, str(tmp_path / "minework"), 4, 0, ref, 0)

    def records(p):
        lines = open(p).read().splitlines()
        return sorted(zip(lines[0::2], lines[1::2]))
    assert records(out_mine) == records(out_ref)
    assert len(records(out_ref)) > 100
    # flanking_seq on top (module/Util.py:4614): byte-identical files for the same input
    U.flanking_seq(out_ref, out_ref + ".flanked", ref, 50)
    coarse.flanking_seq(out_ref, out_mine + ".flanked", ref, 50)
    assert open(out_ref + ".flanked").read() == open(out_mine + ".flanked").read()


def test_mirror_reads_the_reference_chunkers_output(tmp_path, monkeypatch):
    """split_genome_chunks.py (reference, run unmodified in-process) -> genome.cut{i}.fa -> hite_b200.coarse._chunk_to_chromosomes:
    every chunk's records come back at their real coordinates (what is not in the chunk stays N), upper-cased, in file order --
    the input contract of determine_repeat_boundary_v5 (row a1)."""
    import runpy
    import sys
    from hite_b200 import coarse
    ref_import.load_util()                                       # puts the reference's Util (third-party imports stubbed) in sys.modules
    rng = np.random.default_rng(12)
    B = np.frombuffer(b"ACGTacgtN", dtype=np.uint8)
    genome = {"chrA": B[rng.integers(0, 9, 730_123)].tobytes().decode(), "chrB": B[rng.integers(0, 9, 250_000)].tobytes().decode(),
              "scaffold_7 some description": B[rng.integers(0, 9, 99_999)].tobytes().decode()}
    gpath = str(tmp_path / "genome.fa")
    with open(gpath, "w") as f:
        for n, s in genome.items():
            f.write(">%s\n" % n)
            for o in range(0, len(s), 80):
                f.write(s[o:o + 80] + "\n")
    out = tmp_path / "chunks"
    out.mkdir()
    script = os.path.join(ref_import.REF_ROOT, "module", "split_genome_chunks.py")
    monkeypatch.setattr(sys, "argv", [script, "-g", gpath, "--tmp_output_dir", str(out), "--chrom_seg_length", "100000", "--chunk_size", "0.3"])
    runpy.run_path(script, run_name="__main__")
    cuts = sorted(p for p in os.listdir(out) if p.startswith("genome.cut"))
    assert len(cuts) >= 3                                        # 0.3 MiB chunks: the genome spans several
    seen = {}
    for c in cuts:
        names, chroms = coarse._chunk_to_chromosomes(str(out / c))
        for n, arr in zip(names, chroms):
            full = genome[[k for k in genome if k.split(" ")[0] == n][0]].upper()
            s = arr.tobytes().decode()
            assert len(s) <= len(full)
            for i in range(0, len(s), 100_000):                  # a segment is either the real bases or not part of this chunk
                piece = s[i:i + 100_000]
                assert piece == full[i:i + 100_000] or set(piece) == {"N"}, (c, n, i)
                if piece == full[i:i + 100_000] and set(piece) != {"N"}:
                    assert (n, i) not in seen, "a segment appears in two chunks"
                    seen[(n, i)] = c
    total = sum((len(s) + 99_999) // 100_000 for s in genome.values())
    assert len(seen) == total                                    # every segment is in exactly one chunk


def test_drop_in_script_runs_the_reference_stage_with_main_py_arguments(tmp_path, monkeypatch):
    """module/coarse_boundary.py of this repo exactly as main.py invokes it (main.py:520-530): the reference's own stage script
    runs -- its argparse, temp directory, logging, copy-back -- with determine_repeat_boundary_v5 rebound to the mirror (context
    stood in, HSPs and chaining as above).  longest_repeats_{i}.fa and .flanked.fa appear in --tmp_output_dir."""
    import sys
    from hite_b200 import _lib, coarse, fasta
    ref_import.load_util()
    wl = synth.make_workload(genome_bp=2_400_000, n_contigs=2, n_families=12, cands_per_family=1, copies=(5, 10),
                             length=(200, 1200), seed=63)
    ref = str(tmp_path / "genome.fa")
    wl.write_fasta(ref, str(tmp_path / "cand.fa"))
    lens = [len(c) for c in wl.contig_seqs]
    G = O.Genome(wl.contig_seqs)
    hsps, segs, iv = G.coarse_boundary(lens, seg_len=1_000_000, skip_gap=2000, max_len=30000, occ_cap=64)
    G.close()
    cut = str(tmp_path / "genome.cut0.fa")
    with open(cut, "w") as f:
        for c, o in segs:
            f.write(">%s$%d\n%s\n" % (wl.contig_names[c], o, wl.contig_seqs[c][o:o + 1_000_000].tobytes().decode()))

    class OracleBackedCtx:
        device = 0

        def load_genome(self, chroms):
            pass

        def build_index(self, k, w):
            return dict(kernel_ms=0.0)

        def coarse_boundary(self, seg_len, skip_gap, max_len, occ_cap, threads):
            out = np.zeros(len(iv), dtype=_lib.INTERVAL_DTYPE)
            out["chr"] = [c for c, _, _ in iv]; out["start"] = [a for _, a, _ in iv]; out["end"] = [b for _, _, b in iv]
            return out
    monkeypatch.setattr(coarse, "get_context", lambda device=None, role="reference": OracleBackedCtx())
    monkeypatch.setenv("HITE_HOME", ref_import.REF_ROOT)
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    monkeypatch.syspath_prepend(os.path.join(root, "module"))
    import _hite_b200_patch as patch
    outdir, work = tmp_path / "out", tmp_path / "work"
    outdir.mkdir(); work.mkdir()
    prev_te = str(tmp_path / "prev_TE.fa")
    open(prev_te, "w").close()
    monkeypatch.setattr(sys, "argv", ["coarse_boundary.py", "-g", cut, "--tmp_output_dir", str(outdir), "--prev_TE", prev_te,
                                      "--fixed_extend_base_threshold", "2000", "--max_repeat_len", "30000", "--thread", "4",
                                      "--flanking_len", "50", "--tandem_region_cutoff", "0.5", "--ref_index", "0", "-r", ref,
                                      "--recover", "0", "--debug", "0", "-w", str(work)])
    import Util
    saved = Util.determine_repeat_boundary_v5
    try:
        patch.run_stage("coarse_boundary.py", patch_names=("determine_repeat_boundary_v5",), source="hite_b200.coarse")
    finally:
        Util.determine_repeat_boundary_v5 = saved
    names, contigs = fasta.read_fasta(str(outdir / "longest_repeats_0.fa"))
    assert sorted(names) == sorted({"%s:%d-%d" % (wl.contig_names[c], a, b) for c, a, b in iv}) and len(names) > 50
    fnames, fcontigs = fasta.read_fasta(str(outdir / "longest_repeats_0.flanked.fa"))
    assert len(fnames) == len(names) and all(len(s) >= 180 for s in fcontigs.values())
    assert os.path.exists(str(outdir / "HiTE_coarse.log")) and not os.listdir(str(work))


FAKE_TRF_MASK = r'''#!%s
# stand-in for `trf <file> 2 7 7 80 10 50 500 -f -d -m -h`: N-masks every (AC)n run of >= 60 bp, writes <file>...mask in the cwd
import os, re, sys
src = sys.argv[1]
with open(os.path.basename(src) + ".2.7.7.80.10.50.500.mask", "w") as out:
    for line in open(src):
        out.write(line if line.startswith(">") else re.sub(r"(?:AC){30,}", lambda m: "N" * len(m.group(0)), line))
'''


def test_tandem_repeat_filter_is_the_references_own_code(tmp_path, monkeypatch, capfd):
    """determine_repeat_boundary_v5 (mirror) hands the chunk to the reference's own filter_tandem_repeats (module/Util.py:4650-4651,
    :4672-4697: 100 kbp files, `trf` under a process pool, merged in completion order) when trf is on $PATH, and loads the
    N-masked result at the records' real coordinates; without trf it says so and loads the chunk as given."""
    import stat
    import sys
    from hite_b200 import coarse, fasta
    U = ref_import.load_util()
    rng = np.random.default_rng(4)
    B = np.frombuffer(b"ACGT", dtype=np.uint8)
    chroms = {"chrA": B[rng.integers(0, 4, 450_000)].copy(), "chrB": B[rng.integers(0, 4, 230_000)].copy()}
    planted = [("chrA", 120_000, 400), ("chrA", 330_500, 90), ("chrB", 10_000, 1000)]
    for c, a, n in planted:
        chroms[c][a:a + n] = np.frombuffer(b"AC" * (n // 2), dtype=np.uint8)
    ref = str(tmp_path / "genome.fa")
    fasta.store_fasta({c: s.tobytes().decode() for c, s in chroms.items()}, ref)
    cut = str(tmp_path / "genome.cut0.fa")
    with open(cut, "w") as f:
        for c, s in chroms.items():
            for o in range(0, len(s), 100_000):
                f.write(">%s$%d\n%s\n" % (c, o, s[o:o + 100_000].tobytes().decode()))
    loaded = []

    class Ctx:
        device = 0

        def load_genome(self, ch):
            loaded.append([bytes(x) for x in ch])

        def build_index(self, k, w):
            return dict(kernel_ms=0.0)

        def coarse_boundary(self, **kw):
            return np.zeros(0, dtype=coarse_lib.INTERVAL_DTYPE)
    from hite_b200 import _lib as coarse_lib
    monkeypatch.setattr(coarse, "get_context", lambda device=None, role="reference": Ctx())
    bindir = tmp_path / "bin"
    bindir.mkdir()
    p = bindir / "trf"
    p.write_text(FAKE_TRF_MASK % sys.executable)
    p.chmod(p.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(bindir) + os.pathsep + os.environ["PATH"])
    coarse.determine_repeat_boundary_v5(cut, str(tmp_path / "lr.fa"), None, 2000, 30000, str(tmp_path / "w1"), 3, 0, ref, 0,
                                        chrom_seg_length=100_000)
    got = dict(zip(["chrA", "chrB"], loaded[-1]))
    for c, s in chroms.items():
        exp = bytearray(s.tobytes())
        for cc, a, n in planted:
            if cc == c:
                exp[a:a + n] = b"N" * n
        assert got[c] == bytes(exp), c
    assert os.path.exists(str(tmp_path / "w1" / "filter_tandem_0.fa"))          # the reference's own intermediate file
    # no trf: loud message, unfiltered chunk
    monkeypatch.setenv("PATH", os.pathsep.join(x for x in os.environ["PATH"].split(os.pathsep) if x != str(bindir)))
    capfd.readouterr()
    coarse.determine_repeat_boundary_v5(cut, str(tmp_path / "lr2.fa"), None, 2000, 30000, str(tmp_path / "w2"), 3, 0, ref, 0,
                                        chrom_seg_length=100_000)
    assert "SKIPPED" in capfd.readouterr().err
    assert dict(zip(["chrA", "chrB"], loaded[-1])) == {c: s.tobytes() for c, s in chroms.items()}
