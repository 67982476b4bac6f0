"""Host-side mirror of the reference's operator interface for the dynamic boundary-adjustment path.

Same names, argument order and meaning, and the same FASTA-in / FASTA-out contract as

* ``flank_region_align_v5``   reference module/Util.py:8032-8299
* ``run_find_members_v8``     reference module/Util.py:10439-10449
* ``is_TE_from_align_file``   reference module/Util.py:10407-10437

but the copy search (minimap2), the MSA (mafft) and the pure-Python column voting are replaced by
one ``hb_adjust`` call into libhite_b200.so (sm_100a kernels).  There is no CPU fallback: without
the built extension or without a GPU these functions raise ``HiteB200Error``.

The low-copy rescue of module/Util.py:8195-8281 (``trf`` mask, then ``remove_no_tirs`` + TIRPeps domains for TIR, or the
Helitron / non-LTR protein-domain check) is NOT rewritten here: ``_reference_rescue`` runs the reference's OWN statements
for it -- cut out of the reference's ``flank_region_align_v5`` by position in its syntax tree and executed in the
reference's ``Util`` namespace -- whenever the reference is importable and the tools those statements call (``trf``,
``tools/itrsearch``, ``blastx`` / ``makeblastdb``) exist on the host.  Otherwise it says so loudly and every low-copy
element goes to ``all_low_copy``, which is what the reference does with the ones it could not rescue (:8284-8287).
"""
import ast
import os
import re
import shutil
import sys
import threading
import time

from . import _lib
from .fasta import read_fasta, store_fasta

_CTX = {}          # device -> Context
_GENOME = {}       # device -> (path, mtime, size)


def _device():
    env = os.environ.get("HITE_GPU_DEVICES")
    if env and env.strip().lower() != "all":
        devs = [int(x) for x in env.split(",") if x.strip() != ""]
        rank = int(os.environ.get("LOCAL_RANK", os.environ.get("RANK", "0")))
        return devs[rank % len(devs)]
    return int(os.environ.get("LOCAL_RANK", "0"))


def stage_devices():
    """The GPUs ONE stage process spreads its candidates over.  Under torchrun (WORLD_SIZE > 1) a rank owns one GPU (its
    shard comes from hite_b200.dist); a plain stage process -- how main.py runs the stages (main.py:532-628) -- takes every
    GPU listed in HITE_GPU_DEVICES ("0,1,2,3" or "all"), one context and one host thread per GPU, default GPU 0."""
    if int(os.environ.get("WORLD_SIZE", "1")) > 1:
        return [_device()]
    env = os.environ.get("HITE_GPU_DEVICES", "").strip()
    if env.lower() == "all":
        return list(range(max(1, _lib.load().hb_device_count())))
    if env:
        return [int(x) for x in env.split(",") if x.strip() != ""]
    return [0]


_LOCK = threading.Lock()


def get_context(device=None, role="reference"):
    """One context per process, GPU and role.  ``reference``: holds the stage's reference genome + index (loaded once, see
    ensure_genome).  ``members``: a second, small context for sequence sets that come and go -- the member file of one
    candidate (is_TE_from_align_file), the sequence set of a clustering run -- so that they never evict the reference."""
    d = _device() if device is None else device
    key = d if role == "reference" else (d, role)
    with _LOCK:
        if key not in _CTX:
            _CTX[key] = _lib.Context(d)
        return _CTX[key]


def ensure_genome(ctx, reference, k=15, w=10, _parsed=None):
    """Load + index the reference once per process and GPU (the analogue of the ``ref.mmi`` existence check,
    module/Util.py:7938-7947).  ``_parsed``: (names, contigs) already read by the caller."""
    st = os.stat(reference)
    key = (os.path.abspath(reference), st.st_mtime_ns, st.st_size, k, w)
    if _GENOME.get(ctx.device) != key:
        names, contigs = _parsed if _parsed is not None else read_fasta(reference)
        ctx.load_genome([contigs[n].encode("ascii") for n in names], names)
        ctx.build_index(k, w)
        _GENOME[ctx.device] = key
    return ctx


def adjust_on_devices(seqs, reference, TE_type, plant, result_type, flanking_len, devices=None):
    """One adjust round over ``seqs`` spread over the stage's GPUs: candidates are dealt longest-first to the GPUs
    (hite_b200.dist.shard_indices, the reference's own longest-processing-time deal, module/Util.py:1771-1798), every
    GPU holds the whole genome + index, each shard runs on its own context from its own host thread (ctypes releases the
    GIL inside hb_adjust), and the per-candidate results come back in input order.  No collective: the shards meet in host
    memory.  Results do not depend on the number of GPUs."""
    from .dist import shard_indices
    devices = stage_devices() if devices is None else list(devices)
    if len(seqs) < 2 * len(devices):
        devices = devices[:1]
    lens = [len(s) for s in seqs]
    shards = [shard_indices(len(seqs), r, len(devices), lens) for r in range(len(devices))]
    parsed = None
    st = os.stat(reference)
    if any(_GENOME.get(d) != (os.path.abspath(reference), st.st_mtime_ns, st.st_size, 15, 10) for d in devices):
        parsed = read_fasta(reference)                       # read the FASTA once, not once per GPU
    results = [None] * len(seqs)
    errors = []

    def work(dev, mine):
        try:
            ctx = ensure_genome(get_context(dev), reference, _parsed=parsed)
            todo = [mine[b:b + 65536] for b in range(0, len(mine), 65536)]
            while todo:
                chunk = todo.pop(0)
                try:
                    res = ctx.adjust([seqs[i] for i in chunk], TE_type, int(plant), result_type, flank=int(flanking_len))
                except _lib.HiteB200Error as e:
                    # HB_ENOMEM (-4: too many seed hits / MSA columns for one call, or the device is full): halve and retry
                    if "(-4)" not in str(e) or len(chunk) < 2:
                        raise
                    todo[:0] = [chunk[:len(chunk) // 2], chunk[len(chunk) // 2:]]
                    continue
                for i, r in zip(chunk, res):
                    results[i] = r
        except Exception as e:                               # surfaced in the caller's thread: no silent partial result
            errors.append(e)

    threads = [threading.Thread(target=work, args=(d, m)) for d, m in zip(devices[1:], shards[1:]) if m]
    for t in threads:
        t.start()
    if shards[0]:
        work(devices[0], shards[0])
    for t in threads:
        t.join()
    if errors:
        raise errors[0]
    return results


def _post_filter(names, seqs, results, TE_type):
    """module/Util.py:8159-8193: drop 'TG...CA', split by copy-count threshold."""
    true_tes, low_copy = {}, {}
    for name, r in zip(names, results):
        if not r["is_te"]:
            continue
        cons = r["cons"]
        if TE_type in ("tir", "helitron", "non_ltr"):
            if cons.startswith("TG") and cons.endswith("CA"):                     # :8175
                continue
            threshold = 5 if TE_type in ("tir", "non_ltr") else 2                  # :8179-8183
            if r["copy_num"] <= threshold:
                low_copy[name] = cons
            else:
                true_tes[name] = cons
        else:
            true_tes[name] = cons
    return true_tes, low_copy


def _reference_util():
    """The reference's own ``Util`` module: already imported when this runs inside a HiTE stage (module/_hite_b200_patch.py),
    else importable from ``$HITE_HOME/module``; None when there is no HiTE checkout."""
    U = sys.modules.get("Util")
    if U is not None and hasattr(U, "remove_no_tirs") and hasattr(U, "get_domain_info"):
        return U
    home = os.environ.get("HITE_HOME")
    if home and os.path.exists(os.path.join(home, "module", "Util.py")):
        for d in (home, os.path.join(home, "module")):
            if d not in sys.path:
                sys.path.insert(0, d)
        import importlib
        try:
            return importlib.import_module("Util")
        except Exception:
            return None
    return None


_RESCUE_CODE = {}


def _rescue_code(U):
    """The statements of the reference's ``flank_region_align_v5`` between ``store_fasta(low_copy_contigs, low_copy_path)`` and
    ``store_fasta(true_tes, real_TEs)`` (module/Util.py:8195-8281), compiled from the reference's source file as they are."""
    path = U.__file__
    if path in _RESCUE_CODE:
        return _RESCUE_CODE[path]
    code = None
    try:
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")                   # the reference has invalid escape sequences in string literals
            tree = ast.parse(open(path, encoding="utf-8").read(), filename=path)
        fn = [n for n in tree.body if isinstance(n, ast.FunctionDef) and n.name == "flank_region_align_v5"][0]

        def is_store(n, a, b):
            return (isinstance(n, ast.Expr) and isinstance(n.value, ast.Call) and getattr(n.value.func, "id", "") == "store_fasta"
                    and [getattr(x, "id", None) for x in n.value.args] == [a, b])
        i0 = [i for i, n in enumerate(fn.body) if is_store(n, "low_copy_contigs", "low_copy_path")][0]
        i1 = [i for i, n in enumerate(fn.body) if is_store(n, "true_tes", "real_TEs")][0]
        if i1 > i0 + 1:
            code = compile(ast.Module(body=fn.body[i0 + 1:i1], type_ignores=[]), path, "exec")
    except Exception:
        code = None
    _RESCUE_CODE[path] = code
    return code


def _rescue_tools_missing(U, TE_type):
    need = ["trf", "blastx", "makeblastdb"]
    missing = [t for t in need if shutil.which(t) is None]
    rebound = getattr(getattr(U, "run_itrsearch", None), "__module__", "").startswith("hite_b200")   # HB-itr stands in
    if TE_type == "tir" and not rebound and not os.path.exists(os.path.join(U.cur_dir, "tools", "itrsearch")):
        missing.append("tools/itrsearch")
    return missing


def _reference_rescue(low_copy, true_tes, TE_type, temp_dir, threads, plant, log):
    """Low-copy rescue by the reference's own code (module/Util.py:8195-8281): names it promotes are added to ``true_tes``.
    Returns the set of rescued names (empty, with a loud message, when the reference or its tools are not there)."""
    def say(msg):
        if log is not None:
            getattr(log.logger, "warning", log.logger.info)(msg)
        else:
            sys.stderr.write(msg + "\n")
    if not low_copy:
        return set()
    U = _reference_util()
    if U is None:
        say("hite_b200: low-copy rescue SKIPPED for %d %s element(s): the HiTE checkout is not importable (set HITE_HOME); "
            "they go to the all_low_copy file" % (len(low_copy), TE_type))
        return set()
    missing = _rescue_tools_missing(U, TE_type)
    if missing:
        say("hite_b200: low-copy rescue SKIPPED for %d %s element(s): %s not found on this host; they go to the all_low_copy file"
            % (len(low_copy), TE_type, ", ".join(missing)))
        return set()
    code = _rescue_code(U)
    if code is None:
        say("hite_b200: low-copy rescue SKIPPED: the rescue statements were not found in %s (another HiTE version?)" % U.__file__)
        return set()
    low_copy_dir = temp_dir + "/low_copy_itr"                                      # :8155-8158
    os.makedirs(low_copy_dir, exist_ok=True)
    low_copy_path = low_copy_dir + "/low_copy_elements.fa"
    U.store_fasta(low_copy, low_copy_path)                                         # :8194
    before = set(true_tes)
    env = dict(low_copy_contigs=low_copy, low_copy_path=low_copy_path, low_copy_dir=low_copy_dir, temp_dir=temp_dir,
               TE_type=TE_type, plant=plant, threads=threads, true_tes=true_tes, true_te_names=set(true_tes))
    exec(code, U.__dict__, env)
    return set(true_tes) - before


def flank_region_align_v5(candidate_sequence_path, real_TEs, flanking_len, reference, split_ref_dir, TE_type,
                          tmp_output_dir, threads, ref_index, log, subset_script_path, plant, debug, iter_num,
                          all_low_copy, result_type="cons"):
    if log is not None:
        log.logger.info("------Determination of homology in regions outside the boundaries of " + TE_type + " copies")
    starttime = time.time()
    names, contigs = read_fasta(candidate_sequence_path)
    seqs = [contigs[n] for n in names]
    # HB-align limits: a candidate must be 3..32000 bp (max_repeat_len defaults to 30000, main.py:25)
    ok = [i for i, s in enumerate(seqs) if 3 <= len(s) <= 32000]
    results = [dict(is_te=False, info="", cons="", copy_num=0)] * len(seqs)
    if ok:
        res = adjust_on_devices([seqs[i] for i in ok], reference, TE_type, plant, result_type, flanking_len)
        results = list(results)
        for i, r in zip(ok, res):
            results[i] = r
    true_tes, low_copy = _post_filter(names, seqs, results, TE_type)
    temp_dir = tmp_output_dir + "/" + TE_type + "_copies_" + str(ref_index) + "_" + str(iter_num)      # :8037
    rescued = _reference_rescue(low_copy, true_tes, TE_type, temp_dir, threads, plant, log)           # :8195-8281
    store_fasta(true_tes, real_TEs)                                                # :8282
    with open(all_low_copy, "a") as f_save:                                        # :8284-8287
        for te_name, seq in low_copy.items():
            if te_name not in rescued and te_name not in true_tes:
                f_save.write(">" + te_name + "\n" + seq + "\n")
    if debug != 1 and os.path.isdir(temp_dir):                                     # :8295-8296
        shutil.rmtree(temp_dir, ignore_errors=True)
    if debug:
        nb = sum(1 for r in results if r["info"] == "nb")
        fl1 = sum(1 for r in results if r["info"] == "fl1")
        print("not found boundary num: " + str(nb) + ", full length 1: " + str(fl1))
    if log is not None:
        log.logger.info("Running time of determination of homology in regions outside the boundaries of  " + TE_type +
                        " copies: %.8s s" % (time.time() - starttime))
    return results


def subset_100_longest(member_names, member_contigs, keep=100):
    """What ``sh ready_for_MSA.sh <file> 100 100`` keeps of a member file (reference tools/ready_for_MSA.sh, module/Util.py:
    10409-10412): the 100 longest records in file order; records of equal length at the cut are taken in DESCENDING bytewise
    order of their names (the script's ``sort -nk 2 -r`` falls back to a reversed whole-line comparison; C / POSIX locale --
    checked against the real script in tests/test_flank_loop_dropin_cpu.py)."""
    import functools

    def cmp(i, j):
        li, lj = len(member_contigs[member_names[i]]), len(member_contigs[member_names[j]])
        if li != lj:
            return -1 if li > lj else 1
        a, b = member_names[i].encode(), member_names[j].encode()
        return -1 if a > b else (1 if a < b else 0)
    chosen = sorted(range(len(member_names)), key=functools.cmp_to_key(cmp))[:keep]
    return [member_names[i] for i in sorted(chosen)]


def _is_aligned(member_contigs):
    lens = {len(s) for s in member_contigs.values()}
    return len(lens) == 1 and any("-" in s for s in member_contigs.values())


def is_TE_from_align_file(cur_align_file, query_name, cur_seq, temp_dir, subset_script_path, plant, TE_type, debug,
                          result_type):
    """Mirror of module/Util.py:10407-10437.  An aligned member file (equal-length rows with '-') is judged
    directly (kernel iii); an unaligned one is aligned by kernels (i)+(ii) with the members as the target set."""
    member_names, member_contigs = read_fasta(cur_align_file)
    if not member_names:
        return (None, None, "", 0, cur_align_file)
    ctx = get_context(role="members")
    if _is_aligned(member_contigs):
        rows = [member_contigs[n] for n in member_names]
        r = ctx.vote_matrices([cur_seq], [rows], TE_type, int(plant), result_type)[0]
    else:
        if len(member_names) > 100:                                              # :10409-10412, 100 longest
            member_names = subset_100_longest(member_names, member_contigs)
        ctx.load_genome([member_contigs[n].encode("ascii") for n in member_names], member_names)
        ctx.build_index(15, 10)
        # the truncated-member gate is run_find_members_v8's job (it owns the two files), not this call's
        r = ctx.adjust([cur_seq], TE_type, int(plant), result_type, flank=50, trunc_gate=False)[0]
    if r["is_te"]:
        return (query_name, r["cons"], r["info"], r["copy_num"], cur_align_file)
    return (None, None, r["info"], 0, cur_align_file)


def run_find_members_v8(batch_member_file, temp_dir, subset_script_path, plant, TE_type, debug, result_type):
    """Mirror of module/Util.py:10439-10449 (truncated-member gate first, then the full members)."""
    (query_name, cur_seq, trunc_member_file, extend_member_file) = batch_member_file
    if trunc_member_file is not None:
        query_name, cons_seq, info, copy_num, trunc_member_file = is_TE_from_align_file(
            trunc_member_file, query_name, cur_seq, temp_dir, subset_script_path, plant, TE_type, debug, result_type)
        if query_name is not None:
            return is_TE_from_align_file(extend_member_file, query_name, cur_seq, temp_dir, subset_script_path, plant,
                                         TE_type, debug, result_type)
        return query_name, cons_seq, info, copy_num, trunc_member_file
    return is_TE_from_align_file(extend_member_file, query_name, cur_seq, temp_dir, subset_script_path, plant, TE_type,
                                 debug, result_type)


def member_file_name(temp_dir, query_name, trunc=False):
    """module/Util.py:8127-8133 file naming of the per-candidate member files."""
    valid = re.sub(r'[<>:"/\\|?*]', "-", query_name)
    return temp_dir + "/" + valid + (".trunc.blast.bed.fa" if trunc else ".blast.bed.fa")


def get_query_copies(cur_segments, query_contigs, subject_path, query_coverage, subject_coverage,
                     query_fixed_extend_base_threshold=200, subject_fixed_extend_base_threshold=200, max_copy_num=100):
    """Mirror of module/Util.py:6828-7030 (same signature and return value): ``cur_segments`` is a list of
    ``(query_name, {subject_name: [(q_start, q_end, s_start, s_end, identity), ...]})``; returns
    ``{query_name: [(subject_name, start, end, length, '+'|'-'), ...]}``.  The chaining runs in the C-ABI
    library (``hb_query_copies``); HSPs come from BLAST as in the reference, or from ``Context.last_hsps``."""
    import numpy as np
    subject_lens = {}
    if subject_coverage > 0:
        names, contigs = read_fasta(subject_path)
        subject_lens = {n: len(contigs[n]) for n in names}
    all_copies = {}
    for query_name, subject_dict in cur_segments:
        subj_names = list(subject_dict.keys())
        rows, idents = [], []
        for si, sn in enumerate(subj_names):
            for (qs, qe, ss, se, idn) in subject_dict[sn]:
                rows.append((si, qs, qe, ss, se)); idents.append(float(idn))
        hs = np.array(rows, dtype=_lib.HSP_DTYPE) if rows else np.zeros(0, dtype=_lib.HSP_DTYPE)
        sl = [subject_lens.get(sn, 1) for sn in subj_names] or [1]
        out = _lib.query_copies(hs, idents, len(query_contigs[query_name]), sl, query_coverage, subject_coverage,
                                query_fixed_extend_base_threshold, subject_fixed_extend_base_threshold, max_copy_num)
        all_copies[query_name] = [(subj_names[int(o["subj"])], int(o["start"]), int(o["end"]), int(o["qlen"]),
                                   "-" if int(o["strand"]) else "+") for o in out]
    return all_copies
