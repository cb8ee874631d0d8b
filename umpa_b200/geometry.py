"""Output geometry of a match: extent, ROI / step handling, output shape.

Pure-Python restatement of the geometry helpers of the reference's Cython class
(UMPA/model.pyx:531-582 `_calculate_extent` / `_convert_ROI_slice`, 601-646 `_set_ROI`,
`set_step`, `ROI`, `sh`), kept free of any GPU dependency so it is testable on its own.
Quirks kept on purpose:
  * a match() with ROI= or step= overwrites the stored ROI ("sticky", model.pyx:406);
  * step= re-slices the STORED ROI's start/stop (model.pyx:576-580);
  * ROI and step together: step is ignored by match() (model.pyx:372-375) but
    `_convert_ROI_slice` itself raises (model.pyx:566-568).
"""
import numpy as np


class Geometry:
    def __init__(self, shape_list, pos_list, padding, ROI=None):
        self.shape_list = [np.asarray(s, dtype=np.int32) for s in shape_list]
        self.pos_list = [np.asarray(p, dtype=np.int32) for p in pos_list]
        self.padding = int(padding)
        self.ROI = None
        self.set_ROI(ROI)

    def extent(self):
        """model.pyx:531-549: rectangle circumscribing all frames minus 2*padding."""
        pmax = np.max(np.array(self.pos_list) + np.array(self.shape_list), axis=0)
        N0 = 1 + (int(pmax[0]) - 2 * self.padding - 1)
        N1 = 1 + (int(pmax[1]) - 2 * self.padding - 1)
        return N0, N1

    def convert(self, ROI=None, step=None):
        """model.pyx:551-582 -> ((start, stop, step), (start, stop, step)); does not store."""
        N0, N1 = self.extent()
        if ROI is not None:
            if step is not None:
                raise RuntimeError('Step and ROI should not be specified simultaneously.')
            s0, s1 = ROI
            if type(s0) is slice:
                s0 = s0.indices(N0)
            if type(s1) is slice:
                s1 = s1.indices(N1)
        else:
            s0, s1 = self.ROI
            if step is not None:
                s0 = slice(s0[0], s0[1], step).indices(N0)
                s1 = slice(s1[0], s1[1], step).indices(N1)
        return tuple(int(v) for v in s0), tuple(int(v) for v in s1)

    def set_ROI(self, ROI=None):
        """model.pyx:601-616"""
        N0, N1 = self.extent()
        if ROI is None:
            self.ROI = ((0, N0, 1), (0, N1, 1))
        else:
            s0, s1 = ROI
            if type(s0) is slice:
                s0 = s0.indices(N0)
            if type(s1) is slice:
                s1 = s1.indices(N1)
            self.ROI = (tuple(int(v) for v in s0), tuple(int(v) for v in s1))

    def set_step(self, step):
        """model.pyx:618-623"""
        self.set_ROI(self.convert(step=step))
        return self.ROI

    @staticmethod
    def shape_of(s0, s1):
        """Number of elements of range(start, stop, step) per axis (model.pyx:414-415, 641-646)."""
        N0 = 1 + (s0[1] - s0[0] - 1) // s0[2]
        N1 = 1 + (s1[1] - s1[0] - 1) // s1[2]
        return max(int(N0), 0), max(int(N1), 0)

    @property
    def sh(self):
        return self.shape_of(*self.ROI)

    def coords(self, ROI=None):
        """model.pyx:588-599 (the reference's ROI branch passes `self` twice and raises; fixed here)."""
        s0, s1 = self.convert(ROI=ROI) if ROI is not None else self.ROI
        return self.padding + np.arange(*s0), self.padding + np.arange(*s1)

    def match_roi(self, ROI=None, step=None):
        """What _match does first (model.pyx:372-406): resolve, store, return the ROI."""
        if ROI is not None and step is not None:
            step = None
        s0, s1 = self.convert(ROI, step)
        self.set_ROI((s0, s1))
        return s0, s1
