"""Serializer / Deserializer plugin ABCs and their timing wrappers (lmcache/storage_backend/serde/serde.py:12-72)."""
import abc
import time

import torch

from lmcache_b200.logging import init_logger

logger = init_logger(__name__)


class Serializer(metaclass=abc.ABCMeta):

    @abc.abstractmethod
    def to_bytes(self, t: torch.Tensor) -> bytes:
        """Serialize a tensor (any device / shape / dtype the plugin supports) to bytes that carry
        both data and metadata."""
        raise NotImplementedError


class Deserializer(metaclass=abc.ABCMeta):

    @abc.abstractmethod
    def from_bytes(self, bs: bytes) -> torch.Tensor:
        """Inverse of Serializer.to_bytes (accepts bytes or bytearray)."""
        raise NotImplementedError


class SerializerDebugWrapper(Serializer):

    def __init__(self, s: Serializer):
        self.s = s

    def to_bytes(self, t: torch.Tensor) -> bytes:
        start = time.perf_counter()
        bs = self.s.to_bytes(t)
        logger.debug(f"Serialization took {time.perf_counter() - start:.2f} seconds")
        return bs

    def __getattr__(self, name):   # batched / buffer fast paths of the wrapped plugin stay reachable
        return getattr(self.s, name)


class DeserializerDebugWrapper(Deserializer):

    def __init__(self, d: Deserializer):
        self.d = d

    def from_bytes(self, bs: bytes) -> torch.Tensor:
        start = time.perf_counter()
        ret = self.d.from_bytes(bs)
        logger.debug(f"Deserialization took {(time.perf_counter() - start) * 1000:.2f} ms")
        return ret

    def __getattr__(self, name):
        return getattr(self.d, name)
