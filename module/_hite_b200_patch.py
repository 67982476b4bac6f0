"""Shared launcher of the drop-in stage scripts in this directory.

HiTE's ``main.py`` runs its stages by script name found on ``$PATH`` (reference main.py:441-725,
configure.py:48).  Putting this directory *ahead* of HiTE's own ``module/`` on ``$PATH`` makes
``main.py`` call these scripts unchanged; each one executes the reference's own stage script (same
argparse flags, same file names, same logs) after replacing exactly the hot-path functions in the
reference's ``Util`` namespace with the B200 implementations:

    Util.flank_region_align_v5   -> hite_b200.boundary.flank_region_align_v5
    Util.run_find_members_v8     -> hite_b200.boundary.run_find_members_v8
    Util.is_TE_from_align_file   -> hite_b200.boundary.is_TE_from_align_file
    Util.determine_repeat_boundary_v5 -> hite_b200.coarse.determine_repeat_boundary_v5   (coarse_boundary.py)
    Util.multi_process_tsd_v1        -> hite_b200.tir_prefilter.multi_process_tsd_v1     (judge_TIR_transposons.py)
    Util.run_itrsearch               -> hite_b200.tir_prefilter.run_itrsearch            (judge_TIR_transposons.py; HB-itr)
    Util.get_candidate_non_ltr_parallel -> hite_b200.nonltr_prefilter.*                  (judge_Non_LTR_transposons.py)
    Util.multi_process_align_and_get_copies -> hite_b200.other.*                         (judge_Other_transposons.py)
    Util.mask_genome_with_TE         -> hite_b200.masking.mask_genome_with_TE            (judge_LTR_transposons.py)
    Util.run_FiLTR                   -> hite_b200.filtr.run_FiLTR  (judge_LTR_transposons.py; FiLTR's own Util gets
                                        generate_both_ends_frame_from_seq_minimap2 rebound through module/filtr_hook)

Everything else on the stage (TSD pre-filter, cd-hit-est, renaming, prev_TE bookkeeping) stays the
reference's code.  ``HITE_HOME`` must point at the HiTE checkout.  There is no CPU fallback: if the
extension or the GPU is missing the stage fails loudly (main.py ignores exit codes but checks that
the output file exists, reference main.py:541-562).
"""
import os
import runpy
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def hite_home():
    home = os.environ.get("HITE_HOME")
    if not home:
        for cand in (os.path.join(ROOT, "..", "HiTE"), "/root/reference"):
            if os.path.exists(os.path.join(cand, "module", "Util.py")):
                home = cand
                break
    if not home or not os.path.exists(os.path.join(home, "module", "Util.py")):
        sys.stderr.write("hite_b200: set HITE_HOME to the HiTE checkout (module/Util.py not found)\n")
        sys.exit(2)
    return os.path.abspath(home)


def run_stage(script_name, patch_names=("flank_region_align_v5", "run_find_members_v8", "is_TE_from_align_file"),
              source="hite_b200.boundary", extra=()):
    """``extra``: further (name, source module) pairs to rebind."""
    home = hite_home()
    if ROOT not in sys.path:
        sys.path.insert(0, ROOT)
    moddir = os.path.join(home, "module")
    sys.path.insert(0, moddir)
    sys.path.insert(0, home)
    import importlib
    hb = importlib.import_module(source)
    import Util  # the reference's own module, unmodified
    for name in patch_names:
        setattr(Util, name, getattr(hb, name))
    for name, src in extra:
        setattr(Util, name, getattr(importlib.import_module(src), name))
    target = os.path.join(moddir, script_name)
    sys.argv[0] = target
    runpy.run_path(target, run_name="__main__")
