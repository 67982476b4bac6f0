#!/usr/bin/env python
"""Drop-in for HiTE's module/judge_LTR_transposons.py: same CLI (the reference's own argparse runs).  The LTR module itself
(LTR_harvest / LTR_finder / FiLTR, external tools) stays the reference's; its known-TE genome masking
(mask_genome_with_TE, judge_LTR_transposons.py:111 -> module/Util.py:4699) runs on B200.  See _hite_b200_patch.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _hite_b200_patch import run_stage  # noqa: E402

if __name__ == "__main__":
    run_stage("judge_LTR_transposons.py", patch_names=("mask_genome_with_TE",), source="hite_b200.masking")
