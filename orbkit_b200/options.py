"""Module-level option globals, mirroring the subset of orbkit/options.py:495-525 that the
grid-based hot path reads (library users set these attributes directly, e.g. options.quiet=True).

`numproc` and `slice_length` are kept for drop-in compatibility; on B200 they are advisory:
the point range is tiled on-chip and sharded over torch.distributed ranks instead of over a
process pool (orbkit/core.py:437-447,503-536).
"""
quiet = False        #: suppress terminal output (display.py:27-42)
no_log = True        #: no .oklog file (the CLI/logging layer is out of scope)
numproc = 1          #: advisory
slice_length = 1e4   #: advisory: upper bound of points per device launch when > 0
outputname = 'orbkit_b200'  #: base name of output files (extras.calc_mo / calc_ao / mo_set with otype='cb')
no_output = False           #: skip file output even if an otype is given (options.py:512)
exact_mixed_derivatives = False  #: opt-in analytically correct xy/xz/yz AO derivatives
                                 #  (the reference drops cross terms, c_support.c:121-168)
ci_merge_terms = False           #: detci.ci_core: merge duplicate orbital pairs before the launch (faster;
                                 #  summation order then differs from the reference at the 1e-16 level)
ci_fast = None                   #: detci.ci_core summed contractions (rho, jab, a_nabla_b): None = the fused *_from_qc calls
                                 #  use the re-ordered device sums (OKB_FLAG_CI_FAST: dense orbital-pair matrix on the FP64
                                 #  tensor path when there are many terms per pair, terms split over warps otherwise;
                                 #  agreement ~1e-15 of the largest term) and rho()/jab()/a_nabla_b() on given MO arrays keep
                                 #  the reference's summation order bit for bit; True / False force one for both
