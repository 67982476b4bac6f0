#!/usr/bin/env python
"""Golden vectors for HB-itr from the reference's own closed binary tools/itrsearch (run here, in the build container, as a
black box: `itrsearch -i 0.7 -l 7`, exactly how module/Util.py:216-224 runs it).  For every seeded record the file keeps
what the BINARY printed (saved or not, the two arm lengths, score, identity); tests/test_itr_oracle.py measures HB-itr's
agreement with it.  The binary cannot travel to the GPU box; these vectors can.

    python tests/golden/make_golden_itr.py        # needs /root/reference/tools/itrsearch"""
import gzip
import json
import os
import random
import re
import shutil
import stat
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("HITE_REFERENCE", "/root/reference")
COMP = {"A": "T", "C": "G", "G": "C", "T": "A", "N": "N"}


def rc(s):
    return "".join(COMP.get(b, "N") for b in reversed(s))


def make_records(seed, n):
    rng = random.Random(seed)
    rnd = lambda k: "".join(rng.choice("ACGT") for _ in range(k))

    def mutate(s, sub, indel):
        o = []
        for ch in s:
            r = rng.random()
            if r < sub:
                o.append(rng.choice([c for c in "ACGT" if c != ch]))
            elif r < sub + indel / 2:
                pass
            elif r < sub + indel:
                o.append(ch); o.append(rng.choice("ACGT"))
            else:
                o.append(ch)
        return "".join(o)
    recs = []
    for i in range(n):
        k = rng.choice([5, 6, 7, 8, 9, 10, 12, 15, 20, 30, 40, 120])
        tir = rnd(k)
        body = rnd(rng.choice([0, 10, 60, 200, 1500]))
        left, right = rnd(rng.choice([0, 0, 0, 1, 2, 3])), rnd(rng.choice([0, 0, 0, 1, 2]))
        s = left + mutate(tir, rng.choice([0, 0.05, 0.15, 0.3]), rng.choice([0, 0, 0.05, 0.1])) + body + rc(tir) + right
        if rng.random() < 0.5 and len(s) > 80:
            s = s[:40] + s[-40:]                      # the 40 + 40 bp end pairs of search_confident_tir_batch_v1 (:6556-6573)
        if rng.random() < 0.05 and len(s) > 20:
            p = rng.randrange(len(s)); s = s[:p] + "N" + s[p + 1:]
        recs.append(("r%d_%d" % (seed, i), s))
    return recs


def run_binary(recs):
    work = tempfile.mkdtemp(prefix="itr_")
    exe = os.path.join(work, "itrsearch")
    shutil.copy(os.path.join(REF, "tools", "itrsearch"), exe)
    os.chmod(exe, os.stat(exe).st_mode | stat.S_IEXEC)
    with open(os.path.join(work, "f.fa"), "w") as f:
        f.write("".join(">%s\n%s\n" % r for r in recs))
    out = subprocess.run([exe, "-i", "0.7", "-l", "7", "f.fa"], cwd=work, capture_output=True, text=True).stdout
    res, cur = {}, None
    for line in out.splitlines():
        m = re.match(r"load sequence #\d+ (\S+)\s+length=(\d+)", line)
        if m:
            cur = m.group(1); res[cur] = dict(saved=False, len5=0, len3=0, score=0, ident=0.0)
            continue
        m = re.search(r"Similarity Score : (-?\d+)", line)
        if m:
            res[cur]["score"] = int(m.group(1))
        m = re.search(r"Identity percentage : (\S+)", line)
        if m:
            res[cur]["ident"] = float(m.group(1))
        m = re.search(r"Ends at \((\d+),(\d+)\)", line)
        if m:
            res[cur]["len5"], res[cur]["len3"] = int(m.group(1)), int(m.group(2))
        if "saved" in line:
            res[cur]["saved"] = True
    headers = [l[1:].strip() for l in open(os.path.join(work, "f.fa.itr")) if l.startswith(">")]
    shutil.rmtree(work, ignore_errors=True)
    return res, headers


def main():
    cases = []
    for seed in (101, 102, 103):
        recs = make_records(seed, 500)
        res, headers = run_binary(recs)
        hdr = {h.split(" ")[0]: h for h in headers}
        for name, s in recs:
            r = res[name]
            cases.append(dict(name=name, seq=s, saved=r["saved"], len5=r["len5"], len3=r["len3"], score=r["score"], ident=r["ident"],
                              header=hdr.get(name)))
    with gzip.open(os.path.join(HERE, "itrsearch.json.gz"), "wt") as f:
        json.dump(dict(command="itrsearch -i 0.7 -l 7", cases=cases), f)
    print("cases", len(cases), "saved", sum(c["saved"] for c in cases))


if __name__ == "__main__":
    sys.exit(main())
