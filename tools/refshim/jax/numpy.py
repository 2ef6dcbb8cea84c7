from numpy import *  # noqa: F401,F403
import numpy as _np
from numpy import ndarray, float32, float64, int32, int64, bool_  # noqa: F401

bfloat16 = _np.float16    # placeholder so dtype comparisons do not fail; never used for arithmetic here


def asarray(x, dtype=None):
    return _np.asarray(x, dtype=dtype)


def finfo(dt):
    return _np.finfo(dt)
