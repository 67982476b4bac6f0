#!/usr/bin/env python
"""Drop-in for HiTE's module/judge_LTR_transposons.py: same CLI (the reference's own argparse runs).  The LTR detectors
(LtrDetector, LTR_harvest / LTR_finder) stay the reference's; on B200 run the known-TE genome masking (mask_genome_with_TE,
judge_LTR_transposons.py:111 -> module/Util.py:4699) and, inside FiLTR, its fork of the flank-copy alignment loop
(run_FiLTR :146 -> module/Util.py:13847 -> bin/FiLTR-main/src/Util.py:956: hite_b200.filtr).  See _hite_b200_patch.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _hite_b200_patch import run_stage  # noqa: E402

if __name__ == "__main__":
    run_stage("judge_LTR_transposons.py", patch_names=("mask_genome_with_TE",), source="hite_b200.masking",
              extra=(("run_FiLTR", "hite_b200.filtr"),))
