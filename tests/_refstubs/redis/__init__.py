"""Stub of `redis` (never instantiated by the golden generator)."""


class Redis:  # pragma: no cover
    def __init__(self, *a, **k):
        raise NotImplementedError("redis stub")


class Sentinel:  # pragma: no cover
    def __init__(self, *a, **k):
        raise NotImplementedError("redis stub")
