"""The two convenience entry points of the reference's ``UMPA/speckle_matching.py`` (lines 12-75),
same names, arguments and result dictionaries, on top of the CUDA models."""
import numpy as np

from . import model

_MODELS = {True: model.UMPAModelDF, False: model.UMPAModelNoDF}


def _contiguous(frames, what):
    """The reference warns and copies when a frame is not C-contiguous (speckle_matching.py:33-39)."""
    if all(np.asarray(f).flags.c_contiguous for f in frames):
        return frames
    print('Warning: provided list of %s frames are not c contiguous - working with a copy.' % what)
    return [np.ascontiguousarray(f) for f in frames]


def _run(sample, reference, Nw, mask, step, dark_field):
    pm = _MODELS[bool(dark_field)](sam_list=_contiguous(sample, 'sample'), ref_list=_contiguous(reference, 'reference'),
                                   mask_list=mask, window_size=Nw)
    return pm.match(step=step)


def match(Isample, Iref, Nw, mask=None, step=1, max_shift=4, df=True):
    """speckle_matching.py:12-48.  ``max_shift`` is accepted and, as in the reference (which documents it
    as "currently ignored"), not forwarded: the model keeps its default of 4."""
    return _run(Isample, Iref, Nw, mask, step, df)


def match_unbiased(Isample, Iref, Nw, mask=None, step=1, max_shift=4, df=True, bias=True):
    """speckle_matching.py:51-75: ``bias=True`` matches the reference stack against itself and subtracts
    the displacement it finds, ``bias=False`` subtracts nothing, a ``(dx, dy)`` pair is subtracted as given."""
    if bias is True:
        self_match = _run(Iref, Iref, Nw, mask, step, df)
        offset = (self_match['dx'], self_match['dy'])
    elif bias is False:
        offset = (0., 0.)
    else:
        offset = bias
    out = _run(Isample, Iref, Nw, mask, step, df)
    out['dx'] -= offset[0]
    out['dy'] -= offset[1]
    return out
