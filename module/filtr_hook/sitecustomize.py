"""Import hook for the interpreters FiLTR spawns (`python main.py` -> `python LTR_filter.py`, reference
bin/FiLTR-main/main.py:165-177).  Active only under HITE_B200_FILTR=1 (set by hite_b200.filtr.run_FiLTR): when FiLTR's own
`Util` (bin/FiLTR-main/src/Util.py) is imported, its generate_both_ends_frame_from_seq_minimap2 -- the FiLTR fork of HiTE's
flank-copy alignment loop -- is rebound to hite_b200.filtr's mirror; everything else stays FiLTR's code."""
import os
import sys

if os.environ.get("HITE_B200_FILTR") == "1":
    import importlib.abc
    import importlib.util

    class _FiltrUtilHook(importlib.abc.MetaPathFinder):
        def find_spec(self, name, path, target=None):
            if name != "Util":
                return None
            sys.meta_path.remove(self)
            try:
                spec = importlib.util.find_spec(name)
            finally:
                sys.meta_path.insert(0, self)
            if spec is None or not spec.origin or not spec.origin.replace("\\", "/").endswith("FiLTR-main/src/Util.py"):
                return spec
            loader = spec.loader
            orig_exec = loader.exec_module

            def exec_module(module):
                orig_exec(module)
                from hite_b200 import filtr
                module.generate_both_ends_frame_from_seq_minimap2 = filtr.generate_both_ends_frame_from_seq_minimap2
            loader.exec_module = exec_module
            return spec

    sys.meta_path.insert(0, _FiltrUtilHook())
