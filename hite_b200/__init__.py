"""hite_b200 -- Blackwell-native dynamic boundary adjustment for HiTE.

Python host side of ``libhite_b200.so`` (hand-written sm_100a CUDA kernels behind a C-ABI,
``include/hite_b200.h``).  Mirrors the reference's operator interface for this path:

* ``hite_b200.boundary.flank_region_align_v5``        (reference module/Util.py:8032)
* ``hite_b200.boundary.run_find_members_v8``          (reference module/Util.py:10439)
* ``hite_b200.boundary.is_TE_from_align_file``        (reference module/Util.py:10407)
* ``hite_b200.coarse.determine_repeat_boundary_v5``   (reference module/Util.py:4637)
* ``hite_b200.coarse.flanking_seq``                   (reference module/Util.py:4614)
* ``hite_b200.tir_prefilter.multi_process_tsd_v1`` / ``search_confident_tir_v4`` / ``get_short_tir_contigs`` / ``filter_dup_itr_v3``
  (reference module/Util.py:6630, :7734, :7297, :2791 -- the step right before the path in the TIR module)
* ``hite_b200.other.multi_process_align_and_get_copies``          (reference module/Util.py:8382; judge_Other_transposons.py)
* ``hite_b200.masking.mask_genome_intactTE`` / ``mask_genome_with_TE`` (reference module/Util.py:6389, :4699)
* ``hite_b200.nonltr_prefilter.search_polyA_TSD`` / ``get_candidate_non_LTR`` / ``get_candidate_non_ltr_parallel``
  (reference module/Util.py:10915, :11008, :11047)

There is no CPU fallback: without the built extension and a CUDA device every entry point raises.
"""
__version__ = "0.1.0"

from ._lib import HiteB200Error, Context, lib_path  # noqa: F401
