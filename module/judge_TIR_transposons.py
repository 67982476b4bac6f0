#!/usr/bin/env python
"""Drop-in for HiTE's module/judge_TIR_transposons.py: same CLI (the reference's own argparse runs), hot path on B200.
See _hite_b200_patch.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _hite_b200_patch import run_stage  # noqa: E402

if __name__ == "__main__":
    # besides the boundary-adjustment loop, the stage's TSD / TIR pre-filter (multi_process_tsd_v1, judge_TIR_transposons.py:76)
    # and the terminal-inverted-repeat matcher every other caller on the stage uses (run_itrsearch, module/Util.py:216:
    # remove_no_tirs in the low-copy rescue) -- HB-itr on the GPU instead of the closed tools/itrsearch binary
    run_stage("judge_TIR_transposons.py", extra=(("multi_process_tsd_v1", "hite_b200.tir_prefilter"),
                                                 ("run_itrsearch", "hite_b200.tir_prefilter")))
