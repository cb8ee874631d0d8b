"""Convenience wrappers with the reference's signatures (UMPA/speckle_matching.py:12-75)."""
from . import model


def match(Isample, Iref, Nw, mask=None, step=1, max_shift=4, df=True):
    """UMPA/speckle_matching.py:12-48.  Like the reference, ``max_shift`` is accepted but NOT
    forwarded to the model (it is documented there as "currently ignored")."""
    if any(not x.flags.c_contiguous for x in Isample):
        print('Warning: provided list of sample frames are not c contiguous - working with a copy.')
        Isample = [x.copy() for x in Isample]
    if any(not x.flags.c_contiguous for x in Iref):
        print('Warning: provided list of reference frames are not c contiguous - working with a copy.')
        Iref = [x.copy() for x in Iref]
    cls = model.UMPAModelDF if df else model.UMPAModelNoDF
    PM = cls(sam_list=Isample, ref_list=Iref, mask_list=mask, window_size=Nw)
    return PM.match(step=step)


def match_unbiased(Isample, Iref, Nw, mask=None, step=1, max_shift=4, df=True, bias=True):
    """UMPA/speckle_matching.py:51-75: subtract the bias found by matching Iref against itself."""
    if bias is True:
        cls = model.UMPAModelDF if df else model.UMPAModelNoDF
        PMref = cls(sam_list=Iref, ref_list=Iref, mask_list=mask, window_size=Nw)
        bias_result = PMref.match(step=step)
        dx, dy = bias_result['dx'], bias_result['dy']
    elif bias is False:
        dx, dy = 0., 0.
    else:
        dx, dy = bias
    result = match(Isample=Isample, Iref=Iref, Nw=Nw, mask=mask, step=step, max_shift=max_shift, df=df)
    result['dx'] -= dx
    result['dy'] -= dy
    return result
