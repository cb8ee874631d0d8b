"""umpa_b200 -- B200 (sm_100a) implementation of UMPA++'s per-pixel window-matching path.

Drop-in for the ``model`` part of the reference package::

    from umpa_b200 import UMPAModelDF
    res = UMPAModelDF(sam, ref, window_size=2, max_shift=5).match()
    res['dx'], res['dy'], res['T'], res['df'], res['f'], res['err']
"""
from . import model                                    # noqa: F401
from .model import (UMPAModelBase, UMPAModelDF, UMPAModelDFKernel,   # noqa: F401
                    UMPAModelNoDF)
from .speckle_matching import match, match_unbiased    # noqa: F401
from . import align                                    # noqa: F401  (correct_bad_pixels, UMPA_normal, UMPA_nobias)

__version__ = "0.1"
