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

from workloads import synth
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
    monkeypatch.setattr(boundary, "get_context", lambda device=None, role="reference": ctx)
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


FAKE_SAMTOOLS = r'''#!%s
import sys
assert sys.argv[1] == "faidx"
path = sys.argv[2]
out, off = [], 0
data = open(path, "rb").read().split(b"\n")
i = 0
while i + 1 < len(data):
    if data[i].startswith(b">"):
        name = data[i][1:].split()[0].decode()
        off += len(data[i]) + 1
        out.append("%%s\t%%d\t%%d\t%%d\t%%d" %% (name, len(data[i + 1]), off, len(data[i + 1]), len(data[i + 1]) + 1))
        off += len(data[i + 1]) + 1
        i += 2
    else:
        i += 1
open(path + ".fai", "w").write("\n".join(out) + "\n")
'''


def test_ready_for_msa_subset_rule(tmp_path, monkeypatch):
    """What the reference's tools/ready_for_MSA.sh (run for real; only `samtools faidx` is a stand-in) keeps of a member file
    with more than 100 records (module/Util.py:10409-10412): the 100 longest, output in file order, and among records of EQUAL
    length at the cut the ones whose NAME is larger bytewise (`sort -nk 2 -r` falls back to a reversed whole-line comparison;
    C / POSIX locale).  HB-align's S9 keeps the earlier-ranked records on such a tie instead -- with real high-copy families
    (many copies of identical length) the two subsets differ, which is the first item of DESIGN.md's "Next"."""
    import functools
    import random
    import subprocess
    rng = random.Random(5)
    recs = []
    for _ in range(130):
        chrom = rng.choice(["chr1", "chr10", "chr2", "scaffold_3"])
        s0, L = rng.randint(1, 2_000_000), rng.choice([300, 300, 300, 300, 301, 299, 305, 310, 290])
        recs.append(("%s:%d-%d(%s)" % (chrom, s0, s0 + L - 101, rng.choice("+-")), "A" * L))
    assert len({n for n, _ in recs}) == len(recs)
    member = tmp_path / "q.blast.bed.fa"
    member.write_text("".join(">%s\n%s\n" % r for r in recs))
    bindir = tmp_path / "bin"
    bindir.mkdir()
    tool = bindir / "samtools"
    tool.write_text(FAKE_SAMTOOLS % sys.executable)
    tool.chmod(tool.stat().st_mode | stat.S_IEXEC)
    env = dict(os.environ, PATH=str(bindir) + os.pathsep + os.environ["PATH"], LC_ALL="C")
    subprocess.run(["sh", os.path.join(ref_import.REF_ROOT, "tools", "ready_for_MSA.sh"), str(member), "100", "100"], env=env,
                   stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL, check=True)
    kept = [l[1:].strip() for l in open(str(member) + ".rdmSubset.fa") if l.startswith(">")]

    def cmp(i, j):
        (ni, si), (nj, sj) = recs[i], recs[j]
        if len(si) != len(sj):
            return -1 if len(si) > len(sj) else 1
        return -1 if ni.encode() > nj.encode() else 1
    order = sorted(range(len(recs)), key=functools.cmp_to_key(cmp))[:100]
    assert kept == [recs[i][0] for i in sorted(order)]
    earlier = sorted(sorted(range(len(recs)), key=lambda i: (-len(recs[i][1]), i))[:100])
    assert kept != [recs[i][0] for i in earlier]                  # the current S9 tie rule picks other members
    # the host mirror's subset for member FILES (run_find_members_v8 / is_TE_from_align_file on unaligned members) follows the script
    from hite_b200 import boundary
    assert boundary.subset_100_longest([n for n, _ in recs], dict(recs)) == kept


def test_reference_loop_with_more_than_100_copies_under_the_scripts_tie_rule(tmp_path, monkeypatch):
    """Families with 110-140 kept copies through the reference's loop: it subsets the member file with its own
    tools/ready_for_MSA.sh (run for real, `samtools faidx` stood in) before mafft.  With the oracle switched to the script's tie
    rule (Genome.set_tie_rule(1): DESIGN.md section 9 item 1, what the kernels are to follow next) the stand-in mafft finds
    exactly the oracle's 100 members, in file order, and both sides write the same records."""
    from hite_b200 import boundary, fasta
    U = ref_import.load_util()
    wl = synth.make_workload(genome_bp=2_000_000, n_contigs=2, n_families=4, cands_per_family=1, copies=(110, 140),
                             length=(150, 700), seed=77)
    ref, cand = str(tmp_path / "genome.fa"), str(tmp_path / "cand.fa")
    wl.write_fasta(ref, cand)
    names, contigs = fasta.read_fasta(cand)
    lens = [len(c) for c in wl.contig_seqs]
    G = O.Genome(wl.contig_seqs)
    G.set_tie_rule(1, wl.contig_names)
    all_copies, db, verdicts = {}, {}, {}
    n_cut = n_tie = 0
    for name in names:
        seq = contigs[name]
        cps, rows = G.candidate_msa(seq.encode())
        pre, seen = [], set()                                        # every copy that reaches the member file (:8095-8115)
        for c in cps:
            if c["contig"] < 0 or c["start1"] - 1 - 50 < 0 or c["end1"] + 50 > lens[c["contig"]] or c["end1"] - c["start1"] + 101 < 100:
                continue
            if (c["rel"], c["j_s"], c["j_e"]) in seen:
                continue
            seen.add((c["rel"], c["j_s"], c["j_e"]))
            pre.append(c)
        all_copies[name] = [(wl.contig_names[c["contig"]], c["start1"], c["end1"], c["end1"] - c["start1"] + 1, "-" if c["rel"] else "+")
                            for c in pre]
        ext = sorted((c["end1"] - c["start1"] + 101 for c in pre), reverse=True)
        fname = re.sub(r'[<>:"/\\|?*]', "-", name)
        if len(pre) > 100:
            n_cut += 1
            n_tie += ext[99] == ext[100]
            assert len(rows) == 100
            db[fname + ".blast.bed.fa.rdmSubset.fa"] = rows
        else:
            db[fname + ".blast.bed.fa"] = rows
        verdicts[seq] = O.judge_with_trunc_gate(seq, cps, rows, judge=lambda c, r: O.judge_v5(c, r, 1), plant=1)
    G.close()
    assert n_cut >= 3 and n_tie >= 2                                 # the tie at the cut is the normal case
    bindir = tmp_path / "bin"
    bindir.mkdir()
    for tool, body in (("mafft", FAKE_MAFFT), ("samtools", FAKE_SAMTOOLS)):
        p = bindir / tool
        p.write_text(body % sys.executable)
        p.chmod(p.stat().st_mode | stat.S_IEXEC)
    dbpath = str(tmp_path / "msa_db.json")
    json.dump(db, open(dbpath, "w"))
    monkeypatch.setenv("PATH", str(bindir) + os.pathsep + os.environ["PATH"])
    monkeypatch.setenv("HB_FAKE_MAFFT_DB", dbpath)
    monkeypatch.setenv("LC_ALL", "C")
    monkeypatch.setattr(U, "get_full_length_copies_minimap2", lambda q, r, t, m, th: dict(all_copies))

    def no_tirs(confident_tir_path, plant, TRsearch_dir, temp_dir):
        a, b = confident_tir_path + ".with", confident_tir_path + ".without"
        open(a, "w").close()
        open(b, "w").close()
        return a, b
    monkeypatch.setattr(U, "remove_no_tirs", no_tirs)
    monkeypatch.setattr(U, "get_domain_info", lambda cons, lib, table, th, t: open(table, "w").write("#\n#\n"))
    subset_script = os.path.join(ref_import.REF_ROOT, "tools", "ready_for_MSA.sh")
    out_ref, low_ref = str(tmp_path / "ref_real.fa"), str(tmp_path / "ref_low.fa")
    U.flank_region_align_v5(cand, out_ref, 50, ref, None, "tir", str(tmp_path / "refwork"), 2, 0, _Log, subset_script, 1, 0, 0, low_ref, "cons")

    class OracleBackedCtx:
        device = 0

        def adjust(self, seqs, te, plant, result_type, flank=50):
            return [dict(verdicts[s]) for s in seqs]
    ctx = OracleBackedCtx()
    monkeypatch.setattr(boundary, "get_context", lambda device=None, role="reference": ctx)
    monkeypatch.setattr(boundary, "ensure_genome", lambda c, r, k=15, w=10, _parsed=None: c)
    monkeypatch.setenv("HITE_GPU_DEVICES", "0")
    out_mine, low_mine = str(tmp_path / "mine_real.fa"), str(tmp_path / "mine_low.fa")
    boundary.flank_region_align_v5(cand, out_mine, 50, ref, None, "tir", str(tmp_path / "minework"), 2, 0, _Log, "", 1, 0, 0, low_mine, "cons")

    def records(p):
        lines = open(p).read().splitlines() if os.path.exists(p) else []
        return sorted(zip(lines[0::2], lines[1::2]))
    assert records(out_mine) == records(out_ref) and records(low_mine) == records(low_ref)
    assert len(records(out_ref)) >= 3


FAKE_TRF = r'''#!%s
# stand-in for `trf <file> 2 7 7 80 10 50 500 -f -d -m -h`: no tandem repeats found, the .mask file is the input
import shutil, sys, os
src = sys.argv[1]
shutil.copy(src, os.path.basename(src) + ".2.7.7.80.10.50.500.mask")
'''

FAKE_BLASTX = r'''#!%s
# stand-in for `blastx -db <protein fasta> ... -query <fa> -outfmt 6`: one full-length protein hit for every query whose
# name is listed in $HB_FAKE_DOMAIN_NAMES, nothing for the others
import os, sys
a = sys.argv
db, query = a[a.index("-db") + 1], a[a.index("-query") + 1]
pname, plen = None, 0
for line in open(db):
    if line.startswith(">"):
        if pname is not None:
            break
        pname = line[1:].split()[0]
    else:
        plen += len(line.strip())
want = set(os.environ.get("HB_FAKE_DOMAIN_NAMES", "").split(","))
for line in open(query):
    if line.startswith(">") and line[1:].strip() in want:
        q = line[1:].strip()
        print("\t".join(str(x) for x in (q, pname, 100.0, plen, 0, 0, 1, 3 * plen, 1, plen, 1e-50, 500)))
'''


def test_low_copy_rescue_is_the_references_own_code(tmp_path, monkeypatch):
    """Low-copy elements (copy_num <= 5) through the rescue of module/Util.py:8195-8281 -- trf mask, remove_no_tirs (short-TIR
    rules + itrsearch), TIRPeps domain recall -- on both sides: the reference's unmodified flank_region_align_v5, and the
    mirror, which executes the reference's own rescue statements (hite_b200.boundary._reference_rescue) when the tools exist.
    Stand-in trf / itrsearch / blastx; both sides must promote the same elements to the confident set and leave the same
    remainder in the all_low_copy file, and elements must actually be rescued by TIR structure and by domain."""
    from hite_b200 import boundary, fasta
    U = ref_import.load_util()
    wl = synth.make_workload(genome_bp=1_500_000, n_contigs=2, n_families=30, cands_per_family=1, copies=(3, 9),
                             length=(150, 900), seed=31, sub=(0.0, 0.01), indel=0.0)
    ref, cand = str(tmp_path / "genome.fa"), str(tmp_path / "cand.fa")
    wl.write_fasta(ref, cand)
    names, contigs = fasta.read_fasta(cand)
    G = O.Genome(wl.contig_seqs)
    all_copies, db, verdicts = {}, {}, {}
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
            verdicts[seq] = O.judge_with_trunc_gate(seq, cps, rows, judge=lambda c, r: O.judge_v5(c, r, 1), plant=1)
        else:
            verdicts[seq] = dict(is_te=False, info="", cons="", copy_num=0)
    G.close()
    low = [n for n in names if verdicts[contigs[n]]["is_te"] and verdicts[contigs[n]]["copy_num"] <= 5]
    assert len(low) >= 6
    bindir = tmp_path / "bin"
    bindir.mkdir()
    from test_tir_stage_dropin_cpu import FAKE_ITRSEARCH
    for tool, body in (("mafft", FAKE_MAFFT), ("trf", FAKE_TRF), ("blastx", FAKE_BLASTX), ("makeblastdb", "#!%s\n"),
                       ("itrsearch", FAKE_ITRSEARCH)):
        p = bindir / tool
        p.write_text(body % sys.executable)
        p.chmod(p.stat().st_mode | stat.S_IEXEC)
    dbpath = str(tmp_path / "msa_db.json")
    json.dump(db, open(dbpath, "w"))
    monkeypatch.setenv("PATH", str(bindir) + os.pathsep + os.environ["PATH"])
    monkeypatch.setenv("HB_FAKE_MAFFT_DB", dbpath)
    monkeypatch.setenv("HB_FAKE_DOMAIN_NAMES", ",".join(low[-2:]))            # two low-copy elements carry an intact domain
    monkeypatch.setattr(U, "get_full_length_copies_minimap2", lambda q, r, t, m, th: dict(all_copies))
    real_itr = U.run_itrsearch
    monkeypatch.setattr(U, "run_itrsearch", lambda d, inp, inp_dir: real_itr(str(bindir), inp, inp_dir))   # $HITE_HOME/tools is read-only here
    out_ref, low_ref = str(tmp_path / "ref_real.fa"), str(tmp_path / "ref_low.fa")
    U.flank_region_align_v5(cand, out_ref, 50, ref, None, "tir", str(tmp_path / "refwork"), 4, 0, _Log, "", 1, 0, 0, low_ref, "cons")

    class OracleBackedCtx:
        device = 0

        def adjust(self, seqs, te, plant, result_type, flank=50):
            return [dict(verdicts[s]) for s in seqs]
    ctx = OracleBackedCtx()
    monkeypatch.setattr(boundary, "get_context", lambda device=None, role="reference": ctx)
    monkeypatch.setattr(boundary, "ensure_genome", lambda c, r, k=15, w=10, _parsed=None: c)
    monkeypatch.setenv("HITE_GPU_DEVICES", "0")
    assert boundary._rescue_tools_missing(U, "tir") == []
    out_mine, low_mine = str(tmp_path / "mine_real.fa"), str(tmp_path / "mine_low.fa")
    boundary.flank_region_align_v5(cand, out_mine, 50, ref, None, "tir", str(tmp_path / "minework"), 4, 0, _Log, "", 1, 0, 0, low_mine, "cons")

    def records(p):
        lines = open(p).read().splitlines() if os.path.exists(p) else []
        return sorted(zip(lines[0::2], lines[1::2]))
    assert records(out_mine) == records(out_ref) and records(low_mine) == records(low_ref)
    confident = {n[1:] for n, _ in records(out_ref)}
    rescued = [n for n in low if n in confident]
    assert len(rescued) >= 3 and set(low[-2:]) <= confident                   # by TIR structure, and the two by domain
    assert 1 <= len(records(low_ref)) <= len(low) - len(rescued)              # ('TG..CA' elements are dropped before the split)
    print("low-copy rescue: %d low-copy elements, %d rescued (reference == mirror), %d left in all_low_copy"
          % (len(low), len(rescued), len(records(low_ref))))
    # without the tools the mirror says so and rescues nothing
    monkeypatch.setenv("PATH", os.pathsep.join(p for p in os.environ["PATH"].split(os.pathsep) if p != str(bindir)))
    msgs = []

    class Log2:
        class logger:
            info = staticmethod(lambda m: None)
            warning = staticmethod(msgs.append)
    out2, low2 = str(tmp_path / "mine2_real.fa"), str(tmp_path / "mine2_low.fa")
    boundary.flank_region_align_v5(cand, out2, 50, ref, None, "tir", str(tmp_path / "minework2"), 4, 0, Log2, "", 1, 0, 0, low2, "cons")
    assert len(records(low2)) == len(records(low_ref)) + len(rescued) and any("SKIPPED" in m and "trf" in m for m in msgs)
