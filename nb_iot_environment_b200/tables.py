"""The two data tables on the hot path (converted from the reference's files by tools/make_tables.py)."""
import os

import numpy as np

_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'data')
_CACHE = None


def load_tables():
    """(lut2 uint16[7*8*501] = BLER * 2^15 of system/LUT2.csv, mcs uint8[500*30*3] of system/loss_tbs_to_mcs.p)"""
    global _CACHE
    if _CACHE is None:
        lut = np.fromfile(os.path.join(_DIR, 'lut2_u16.bin'), dtype='<u2')
        mcs = np.fromfile(os.path.join(_DIR, 'mcs_u8.bin'), dtype='u1')
        if lut.size != 7 * 8 * 501 or mcs.size != 500 * 30 * 3:
            raise RuntimeError('data tables are corrupt: rerun tools/make_tables.py')
        _CACHE = (lut, mcs)
    return _CACHE
