"""Stub of the `nvtx` module so the reference tree imports in this container
(golden-vector generation only; see tests/golden/make_golden.py)."""


def annotate(message=None, color=None, domain=None):
    def deco(fn):
        return fn
    return deco
