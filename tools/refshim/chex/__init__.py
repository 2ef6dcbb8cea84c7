import dataclasses as _dc
import numpy as _np

Array = _np.ndarray
PRNGKey = int


def dataclass(cls=None, **kw):
    kw.pop("mappable_dataclass", None)
    if cls is None:
        return lambda c: _dc.dataclass(c, **kw)
    return _dc.dataclass(cls, **kw)
