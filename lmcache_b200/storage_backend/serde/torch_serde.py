"""Lossless torch.save/torch.load serde (lmcache/storage_backend/serde/torch_serde.py:16-32).

Byte format is unchanged (torch's zip+pickle container), so a reference TorchDeserializer reads
what this TorchSerializer writes and vice versa.  What changes is the device->host hop: the
reference's pageable `t.cpu().clone()` becomes one async copy into page-locked memory on the
caller's stream (the mover of include/b200kv.h)."""
import ctypes
import io

import torch

from lmcache_b200 import _native as N
from lmcache_b200.storage_backend.serde.serde import Deserializer, Serializer


def _to_host(t: torch.Tensor) -> torch.Tensor:
    if not t.is_cuda:
        return t.detach().clone()
    N.require_cuda()
    src = t.detach().contiguous()
    host = torch.empty(src.shape, dtype=src.dtype, pin_memory=True)
    with torch.cuda.device(src.device):
        sp = torch.cuda.current_stream().cuda_stream
        N.check(N.lib().b200kv_copy_async(ctypes.c_void_p(host.data_ptr()), ctypes.c_void_p(src.data_ptr()),
                                          src.numel() * src.element_size(), sp), "copy_async")
        N.check(N.lib().b200kv_stream_sync(sp), "stream_sync")
    return host


class TorchSerializer(Serializer):

    def to_bytes(self, t: torch.Tensor) -> bytes:
        with io.BytesIO() as f:
            torch.save(_to_host(t), f)
            return f.getvalue()


class TorchDeserializer(Deserializer):

    def from_bytes_normal(self, b: bytes) -> torch.Tensor:
        with io.BytesIO(b) as f:
            return torch.load(f)

    def from_bytes(self, b: bytes) -> torch.Tensor:
        return self.from_bytes_normal(b)
