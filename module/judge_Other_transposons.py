#!/usr/bin/env python
"""Drop-in for HiTE's module/judge_Other_transposons.py (homology non-LTR module): same CLI (the reference's own
argparse runs); its library-vs-genome copy search (multi_process_align_and_get_copies, module/Util.py:8382) runs on B200.
See _hite_b200_patch.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _hite_b200_patch import run_stage  # noqa: E402

if __name__ == "__main__":
    run_stage("judge_Other_transposons.py", patch_names=("multi_process_align_and_get_copies",), source="hite_b200.other")
