#!/usr/bin/env python
"""Drop-in for HiTE's module/judge_Non_LTR_transposons.py: same CLI (the reference's own argparse runs), hot path on B200.
See _hite_b200_patch.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _hite_b200_patch import run_stage  # noqa: E402

if __name__ == "__main__":
    run_stage("judge_Non_LTR_transposons.py", extra=(("get_candidate_non_ltr_parallel", "hite_b200.nonltr_prefilter"),))
