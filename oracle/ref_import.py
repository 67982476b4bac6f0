"""TEST INFRASTRUCTURE (oracle) -- not part of the product path.

Imports the *unmodified* reference module ``/root/reference/module/Util.py`` read-only so that
golden fixtures can be generated from the reference's own functions
(``judge_boundary_v5/v6/v9``, ``remove_sparse_col_in_align_file``, ``search_boundary_homo_v3/v4``,
``get_longest_repeats_v4``, ``get_query_copies`` ...).  Only usable in the build container:
``/root/reference`` does not exist on the GPU box, so nothing under ``-m gpu`` may import this.

Third-party modules the reference imports at top level but which are not installed here are
stubbed; the two whose *behaviour* matters are replaced by ``oracle.fnm`` (see there for the
"parity unpinned" statement about fuzzysearch).
"""
import importlib
import os
import sys
import types
from unittest.mock import MagicMock

REF_ROOT = os.environ.get("HITE_REFERENCE", "/root/reference")


def available():
    return os.path.exists(os.path.join(REF_ROOT, "module", "Util.py"))


def load_util():
    if "Util" in sys.modules and getattr(sys.modules["Util"], "_hb_oracle", False):
        return sys.modules["Util"]
    if not available():
        raise RuntimeError("reference not mounted at " + REF_ROOT)
    from . import fnm
    for name in ("pysam", "seaborn", "PyPDF2", "matplotlib", "matplotlib.pyplot"):
        if name not in sys.modules:
            try:
                importlib.import_module(name)
            except Exception:
                sys.modules[name] = MagicMock()
    fz = types.ModuleType("fuzzysearch")
    fz.find_near_matches = fnm.find_near_matches
    sys.modules["fuzzysearch"] = fz
    lv = types.ModuleType("Levenshtein")
    lv.distance = fnm.levenshtein
    sys.modules["Levenshtein"] = lv
    moddir = os.path.join(REF_ROOT, "module")
    if moddir not in sys.path:
        sys.path.insert(0, moddir)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        util = importlib.import_module("Util")
    util._hb_oracle = True
    return util
