"""Host mirror of the TSD / short-TIR pre-filter that runs right before the boundary-adjustment path in HiTE's TIR
module (reference module/judge_TIR_transposons.py:71-77 -> Util.py:6630 ``multi_process_tsd_v1`` ->
``search_confident_tir_batch`` :7358): same function names, arguments and return shapes; the k-mer matching runs in
``hb_tsd_candidates`` (hite_b200/csrc/hb_tsd.cu).  The two ``itrsearch`` calls of the reference loop (a closed-source
binary, Util.py:216-224) are NOT part of this module: it returns exactly what the reference hands to them.

Record names: the reference numbers the variants of a candidate by their position in a distance-sorted Python *set*,
so among equal distances its ``-C_<i>-`` is hash-seed dependent; here ties are ordered by (start, end, k).
"""
from . import _lib

_CTX = None


def _ctx():
    """The stage process's context (shared with hite_b200.boundary: one context per process per GPU, HITE_GPU_DEVICES honoured)."""
    from . import boundary
    return boundary.get_context()


def _records_to_dict(query_name, seq, recs, n):
    out = {}
    for r in recs[:n]:
        s, e, k = int(r["start"]), int(r["end"]), int(r["k"])
        name = "%s-C_%d-tsd_%s-distance_%d" % (query_name, int(r["idx"]), seq[e + 1:e + 1 + k], int(r["dist"]))   # :7836
        out[name] = seq[s:e + 1]
    return out


def search_confident_tir_v4(orig_seq, raw_tir_start, raw_tir_end, tsd_search_distance, query_name, plant, ctx=None):
    """reference module/Util.py:7734 -- {new_query_name: tir_seq}, at most 100, closest to the raw boundaries first."""
    recs, cnt = (ctx or _ctx()).tsd_candidates([orig_seq], [raw_tir_start], [raw_tir_end], tsd_search_distance, plant)
    return _records_to_dict(query_name, orig_seq, recs[0], int(cnt[0]))


def get_short_tir_contigs(cur_itr_contigs, plant, classes=None):
    """reference module/Util.py:7297.  ``classes`` (name -> short_class from ``search_confident_tir_batch``) avoids
    re-deriving what the kernel already computed; without it the rule is evaluated here on the strings."""
    if classes is not None:
        return {n: s for n, s in cur_itr_contigs.items() if classes.get(n, 0)}
    comp = {"A": "T", "T": "A", "C": "G", "G": "C"}
    rc = lambda x: "".join(comp.get(b, "N") for b in reversed(x))
    out = {}
    for name, tir in cur_itr_contigs.items():
        parts = name.split("-tsd_")
        tsd_len = len(parts[1].split("-")[0]) if len(parts) > 1 else 0
        if tir[:5] == rc(tir[-5:]):
            if (tsd_len == 8 and len(tir) < 4000) or 9 <= tsd_len <= 11 or \
                    (plant == 1 and tsd_len == 3 and tir[:5] in ("CACTA", "CACTG")):
                out[name] = tir
        elif tir[:3] == rc(tir[-3:]) and tir[:3] == "CCC":
            out[name] = tir
    return out


def search_confident_tir_batch(cur_contigs, flanking_len, plant, ctx=None):
    """The per-partition loop of the reference's ``search_confident_tir_batch`` (module/Util.py:7358-7378) up to the
    first ``itrsearch`` call: (all_copies_itr_contigs, short_itr_contigs), insertion-ordered like the reference."""
    names, seqs = [], []
    for query_name, seq in cur_contigs.items():
        if "NNNNNNNNNN" in seq:                                  # :7368
            continue
        names.append(query_name)
        seqs.append(seq)
    starts = [flanking_len + 1] * len(seqs)                      # :7371
    ends = [len(s) - flanking_len for s in seqs]                 # :7372
    keep = [i for i in range(len(seqs)) if 1 <= ends[i] <= len(seqs[i]) and starts[i] <= len(seqs[i])]
    recs, cnt = (ctx or _ctx()).tsd_candidates([seqs[i] for i in keep], [starts[i] for i in keep], [ends[i] for i in keep],
                                              flanking_len, plant)
    all_contigs, classes = {}, {}
    for row, i in enumerate(keep):
        d = _records_to_dict(names[i], seqs[i], recs[row], int(cnt[row]))
        for (name, _), r in zip(d.items(), recs[row][:int(cnt[row])]):
            classes[name] = int(r["short_class"])
        all_contigs.update(d)
    return all_contigs, get_short_tir_contigs(all_contigs, plant, classes)


def filter_dup_itr_v3(cur_copies_out_contigs, TIR_len_dict):
    """reference module/Util.py:2791-2812: of one candidate's TSD-bounded variants keep the one closest to the raw boundaries
    (first among equals, in the dict's order), renamed ``<query>-tir_<len>-tsd_<tsd>``; nothing if it is 30 000 bp or longer.
    Host-side selection only (one string split per variant)."""
    best_name, best_dist = "", 100000
    for name in cur_copies_out_contigs:
        dist = int(name.split("-distance_")[1])
        if dist < best_dist:
            best_name, best_dist = name, dist
    query, rest = best_name.split("-C_")[0], best_name.split("-C_")[1]
    tsd = rest.split("-tsd_")[1].split("-")[0]
    seq = cur_copies_out_contigs[best_name]
    if len(seq) >= 30_000:
        return {}
    return {"%s-tir_%d-tsd_%s" % (query, TIR_len_dict.get(best_name, 0), tsd): seq}


def run_itrsearch(TRsearch_dir, input, input_dir, ctx=None):
    """Same signature, file names and output format as the reference's run_itrsearch (module/Util.py:216-224, which runs the
    closed binary ``<TRsearch_dir>/itrsearch -i 0.7 -l 7 <input>``): writes ``<input>.itr`` -- one record per input record with
    a terminal inverted repeat, header ``><name> ITR(1,a)..(L,b) Length itr=<a-1>``, sequence in 60-column lines -- and
    ``<input>.log``; returns both paths.  The search itself is HB-itr on the GPU (``hb_itr_search``; parity with the binary is
    unpinned, see hite_b200/csrc/hb_itr.cu).  ``HITE_B200_ITRSEARCH=binary`` runs the reference's binary instead."""
    import os
    os.makedirs(input_dir, exist_ok=True)
    if os.environ.get("HITE_B200_ITRSEARCH", "").lower() == "binary":
        return _run_itrsearch_binary(TRsearch_dir, input, input_dir)
    names, seqs = [], []
    if os.path.exists(input):
        name = None
        for line in open(input):                              # read_fasta_v1 semantics: the whole header line is the name
            line = line.rstrip("\n")
            if line.startswith(">"):
                name = line[1:]
                names.append(name); seqs.append([])
            elif name is not None and line.strip():
                seqs[-1].append(line.strip())
    seqs = ["".join(x) for x in seqs]
    keep = [i for i, s in enumerate(seqs) if s]
    recs = (ctx or _ctx()).itr_search([seqs[i].upper() for i in keep]) if keep else []
    with open(input + ".itr", "w") as out, open(input + ".log", "w") as log:
        for k, i in enumerate(keep):
            r, s, nm = recs[k], seqs[i], names[i].split()[0] if names[i].split() else names[i]
            log.write("load sequence #%d %s\tlength=%d\n" % (k + 1, nm, len(s)))
            if not int(r["found"]):
                log.write("\n no alignment!\n\n")
                continue
            a, b = int(r["len5"]), len(s) - int(r["len3"]) + 1
            log.write("\t(1,%d)..(%d,%d) Length itr=%d\n\tLength TE=%d\n\t\t==> saved !\n" % (a, b, len(s), a, len(s)))
            out.write(">%s ITR(1,%d)..(%d,%d) Length itr=%d\n" % (nm, a, len(s), b, a - 1))
            for o in range(0, len(s), 60):
                out.write(s[o:o + 60] + "\n")
    return input + ".itr", input + ".log"


def _run_itrsearch(TRsearch_dir, input_path, work_dir):
    return run_itrsearch(TRsearch_dir, input_path, work_dir)


def _run_itrsearch_binary(TRsearch_dir, input_path, work_dir):
    """The reference's run_itrsearch (module/Util.py:216-224): the external ``itrsearch -i 0.7 -l 7`` of the HiTE tools
    directory (closed binary, not part of this repository) on a FASTA of 80 bp end pairs -> (<input>.itr, <input>.log)."""
    import os
    import subprocess
    os.makedirs(work_dir, exist_ok=True)
    log = input_path + ".log"
    with open(log, "w") as f:
        subprocess.run([os.path.join(TRsearch_dir, "itrsearch"), "-i", "0.7", "-l", "7", input_path], cwd=work_dir, stdout=f,
                       stderr=subprocess.DEVNULL)
    return input_path + ".itr", log


def _itr_headers(path):
    """Header lines of an itrsearch output (read_fasta_v1 semantics, module/Util.py:1694: the whole line is the name; a
    record without sequence is dropped)."""
    import os
    names = []
    if os.path.exists(path):
        name, has_seq = "", False
        for line in open(path):
            if line.startswith(">"):
                if name and has_seq:
                    names.append(name)
                name, has_seq = line.strip()[1:], False
            elif line.strip():
                has_seq = True
        if name and has_seq:
            names.append(name)
    return names


def tsd_variant_groups(longest_repeats_flanked_path, tir_tsd_dir, flanking_len, threads, TRsearch_dir, plant, ctx=None,
                       partition=20_000):
    """Everything of ``multi_process_tsd_v1`` up to the per-candidate selection -> (candidate names in file order,
    {candidate: {variant name: sequence}}, {variant name: TIR length}).  A candidate's dict lists the variants itrsearch
    reported first, then its short-TIR variants, like the reference's dict (:6600-6609)."""
    import os
    import shutil
    from concurrent.futures import ThreadPoolExecutor
    from .fasta import read_fasta, store_fasta
    if os.path.exists(tir_tsd_dir):
        shutil.rmtree(tir_tsd_dir, ignore_errors=True)
    os.makedirs(tir_tsd_dir, exist_ok=True)
    names, contigs = read_fasta(longest_repeats_flanked_path)
    all_contigs, short_contigs = search_confident_tir_batch({n: contigs[n] for n in names}, int(flanking_len), int(plant), ctx=ctx)
    rest = {n: s for n, s in all_contigs.items() if n not in short_contigs}

    def parts(d, tag):
        items = list(d.items())
        out = []
        for k in range(0, len(items), partition):
            path = os.path.join(tir_tsd_dir, "%d%s.fa" % (k // partition, tag))
            store_fasta({n: s[:40] + s[-40:] for n, s in items[k:k + partition]}, path)     # :6556-6560, :6569-6573
            out.append(path)
        return out
    jobs = parts(short_contigs, "_s") + parts(rest, "")
    with ThreadPoolExecutor(max(1, int(threads))) as ex:
        outs = list(ex.map(lambda p: _run_itrsearch(TRsearch_dir, p, tir_tsd_dir), jobs))
    tir_len, with_tir = {}, set()
    for (itr_path, _), job in zip(outs, jobs):
        for header in _itr_headers(itr_path):
            name = header.split(" ")[0]
            tir_len[name] = int(header.split("Length itr=")[1])                             # :6586-6593
            if not job.endswith("_s.fa"):
                with_tir.add(name)
    groups = {}
    for name, seq in rest.items():                                                          # :6603-6605
        if name in with_tir:
            groups.setdefault(name.split("-C_")[0], {})[name] = seq
    for name, seq in short_contigs.items():                                                 # :6606 update(short_itr_contigs)
        groups.setdefault(name.split("-C_")[0], {})[name] = seq
    shutil.rmtree(tir_tsd_dir, ignore_errors=True)
    return names, groups, tir_len


def multi_process_tsd_v1(longest_repeats_flanked_path, tir_tsd_path, tir_tsd_dir, flanking_len, threads, TRsearch_dir, plant,
                         ctx=None):
    """reference module/Util.py:6630-6655 (what judge_TIR_transposons.py:76 calls) with ``search_confident_tir_batch_v1``
    (:6533-6628) inside: per flanked candidate the TSD-bounded variants (``hb_tsd_candidates``, ONE launch for the whole file
    instead of a process pool over 10 kB partitions), the 40 + 40 bp end pairs of the variants through the external
    ``itrsearch`` (``threads`` parallel partitions; short-TIR classes and the rest separately, as the reference does), the
    TIR length of every variant from its output, and per candidate the variant closest to the raw boundaries
    (``filter_dup_itr_v3``) -> ``tir_tsd_path`` FASTA, records in candidate order.  Among variants of equal distance the
    reference's choice follows Python's set order (it changes with the hash seed of the run); here it is (start, end, k)."""
    from .fasta import store_fasta
    names, groups, tir_len = tsd_variant_groups(longest_repeats_flanked_path, tir_tsd_dir, flanking_len, threads, TRsearch_dir,
                                                plant, ctx=ctx)
    result = {}
    for q in names:
        if q in groups:
            result.update(filter_dup_itr_v3(groups[q], tir_len))
    store_fasta(result, tir_tsd_path)
