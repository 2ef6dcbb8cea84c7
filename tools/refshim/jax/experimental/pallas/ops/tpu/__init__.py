"""Import-only stub of the TPU splash-attention module (the reference subclasses one of its types at
import time, gencast/sparse_transformer.py:217; the default attention type never calls into it)."""


class _Namespace:
    def __getattr__(self, name):
        if name[:1].isupper():
            return type(name, (), {})
        return _Namespace()

    def __call__(self, *a, **k):
        raise NotImplementedError("refshim: TPU splash attention is not available")


splash_attention = _Namespace()
