#!/usr/bin/env python
"""Drop-in for HiTE's module/coarse_boundary.py: same CLI (the reference's own argparse runs); the
all-vs-all + FMEA step runs on the B200 (hite_b200.coarse).  See _hite_b200_patch.py."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from _hite_b200_patch import run_stage  # noqa: E402

if __name__ == "__main__":
    run_stage("coarse_boundary.py", patch_names=("determine_repeat_boundary_v5",), source="hite_b200.coarse")
