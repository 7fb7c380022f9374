"""LMCRemoteBackend -- serde + connector (lmcache/storage_backend/remote_backend.py:24-180) and the pipelined
variant (:183-275).  This is the caller of the serde plugins: put = to_bytes -> connection.set,
get = connection.get -> from_bytes.  Non-blocking puts keep the reference's single worker thread, so at most
one encode runs at a time per backend (the codec owns per-direction buffers under that assumption)."""
import queue
import threading
from typing import Iterable, Iterator, List, Optional, Set, Tuple, Union

import torch

from lmcache_b200.config import LMCacheEngineConfig, LMCacheEngineMetadata
from lmcache_b200.logging import init_logger
from lmcache_b200.storage_backend.abstract_backend import LMCBackendInterface
from lmcache_b200.storage_backend.connector import CreateConnector
from lmcache_b200.storage_backend.serde import CreateSerde
from lmcache_b200.utils import CacheEngineKey, _lmcache_nvtx_annotate

logger = init_logger(__name__)


class RemoteBackendEndSignal:
    pass


class LMCRemoteBackend(LMCBackendInterface):

    def __init__(self, config: LMCacheEngineConfig, metadata: LMCacheEngineMetadata):
        super().__init__()
        self.existing_keys: Set[CacheEngineKey] = set()
        self.put_thread = None
        self.connection = None
        assert config.remote_url is not None, "Need to provide remote_url when using LMCRemoteBackend"
        assert config.remote_serde is not None, "Need to provide remote_serde when using LMCRemoteBackend"
        self.connection = CreateConnector(config.remote_url)
        self.serializer, self.deserializer = CreateSerde(config.remote_serde, config, metadata)
        self.dst_device = "cuda"
        self._device = torch.cuda.current_device() if torch.cuda.is_available() else None
        self.put_queue: "queue.Queue[Union[Tuple[CacheEngineKey, torch.Tensor], RemoteBackendEndSignal]]" = \
            queue.Queue()
        self.put_thread = threading.Thread(target=self.put_worker, args=(), daemon=True)
        self.put_thread.start()

    @_lmcache_nvtx_annotate
    def put_worker(self):
        if self._device is not None:
            torch.cuda.set_device(self._device)
        while True:
            item = self.put_queue.get()
            if isinstance(item, RemoteBackendEndSignal):
                self.put_queue.task_done()
                break
            try:
                if item[0] == "view":
                    _, keys, view, tok_begin, chunk_size = item
                    self._put_view_blocking(keys, view, tok_begin, chunk_size)
                else:
                    key, value = item
                    self.put_blocking(key, value)
            except Exception as e:   # a failed background put is a cache miss later, not a crash
                logger.error(f"background put failed: {e}")
            finally:
                self.put_queue.task_done()

    def _combine_key(self, key: CacheEngineKey) -> str:
        return key.to_string()

    def _split_key(self, key: str) -> CacheEngineKey:
        return CacheEngineKey.from_string(key)

    def list(self) -> List[CacheEngineKey]:
        keys = [self._split_key(k) for k in self.connection.list()]
        self.existing_keys.update(keys)
        return keys

    def contains(self, key: CacheEngineKey) -> bool:
        if key in self.existing_keys:
            return True
        flag = self.connection.exists(self._combine_key(key))
        if flag:
            self.existing_keys.add(key)
        return flag

    def put_blocking(self, key: CacheEngineKey, kv_chunk: torch.Tensor) -> None:
        bs = self.serializer.to_bytes(kv_chunk)
        self.connection.set(self._combine_key(key), bs)
        self.existing_keys.add(key)

    def put(self, key: CacheEngineKey, kv_chunk: torch.Tensor, blocking: bool = True) -> None:
        if blocking:
            self.put_blocking(key, kv_chunk)
        else:
            self.put_queue.put((key, kv_chunk))

    @_lmcache_nvtx_annotate
    def get(self, key: CacheEngineKey) -> Optional[torch.Tensor]:
        if not self.contains(key):
            return None
        bs = self.connection.get(self._combine_key(key))
        if bs is None or len(bs) == 0:
            return None
        return self.deserializer.from_bytes(bs).to(self.dst_device)

    # ------------------------------------------------------------------ engine fast paths (no per-chunk blobs)
    def supports_kv_view(self) -> bool:
        """True when the serde plugin can encode / decode straight from / into the engine's KV tensors."""
        return hasattr(self.serializer, "view_to_bytes_batch") and hasattr(self.deserializer, "decode_into")

    def _put_view_blocking(self, keys, view, tok_begin: int, chunk_size: int) -> None:
        n_tokens = view.ntokens - tok_begin
        if hasattr(self.serializer, "view_to_pinned_batch"):
            # containers are sent from the page-locked slab the device->host copies landed in
            with self.serializer.view_to_pinned_batch(view, chunk_size, tok_begin, n_tokens) as blobs:
                assert len(blobs) == len(keys)
                for key, mv in zip(keys, blobs):
                    self.connection.set(self._combine_key(key), mv)
                    self.existing_keys.add(key)
            return
        blobs = self.serializer.view_to_bytes_batch(view, chunk_size, tok_begin, n_tokens)
        assert len(blobs) == len(keys)
        for key, bs in zip(keys, blobs):
            self.connection.set(self._combine_key(key), bs)
            self.existing_keys.add(key)

    def put_kv_chunks(self, keys: List[CacheEngineKey], view, tok_begin: int, chunk_size: int,
                      blocking: bool = True) -> int:
        """Store tokens [tok_begin, T) of `view` as len(keys) chunks: one batched encode (all chunks in one kernel
        launch sequence), then one set() per chunk.  Replaces the engine's blob pack + per-chunk to_bytes."""
        if blocking:
            self._put_view_blocking(keys, view, tok_begin, chunk_size)
        else:
            self.put_queue.put(("view", list(keys), view, tok_begin, chunk_size))
        return len(keys)

    def get_kv_into(self, keys: List[CacheEngineKey], dst, dst_tok0: int, chunk_size: int) -> int:
        """Fetch consecutive chunks until the first miss and decode them with ONE batched launch straight into `dst`
        (chunk i lands at token dst_tok0 + i * chunk_size).  Returns (number of chunks, tokens written)."""
        if hasattr(self.connection, "get_into") and hasattr(self.deserializer, "pinned_staging"):
            # receive straight into a page-locked slab and upload from there (true async copies, no bytearrays)
            bound = (self.deserializer.container_bound(dst.L, dst.H, dst.D, chunk_size) + 63) & ~63
            with self.deserializer.pinned_staging(bound * len(keys)) as pin:
                blobs = []
                for i, key in enumerate(keys):
                    if not self.contains(key):
                        break
                    n = self.connection.get_into(self._combine_key(key), pin.host_ptr + i * bound, bound)
                    if n is None or n == 0:
                        break
                    blobs.append(pin.view(i * bound, n))
                if blobs:
                    self.deserializer.decode_into(blobs, dst, [dst_tok0 + i * chunk_size for i in range(len(blobs))])
                return len(blobs)
        blobs = []
        for key in keys:
            if not self.contains(key):
                break
            bs = self.connection.get(self._combine_key(key))
            if bs is None or len(bs) == 0:
                break
            blobs.append(bs)
        if blobs:
            self.deserializer.decode_into(blobs, dst, [dst_tok0 + i * chunk_size for i in range(len(blobs))])
        return len(blobs)

    def close(self):
        if self.put_thread is not None and self.put_thread.is_alive():
            self.put_queue.put(RemoteBackendEndSignal())
            self.put_thread.join()
        if self.connection is not None:
            self.connection.close()
            self.connection = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class LMCPipelinedRemoteBackend(LMCRemoteBackend):
    """batched_get with the network fetch of chunk i+1 overlapped with the decode of chunk i
    (remote_backend.py:183-275: network thread + deserialize thread)."""

    def __init__(self, config: LMCacheEngineConfig, metadata: LMCacheEngineMetadata):
        super().__init__(config, metadata)

    @_lmcache_nvtx_annotate
    def batched_get(self, keys: Iterator[CacheEngineKey]) -> Iterable[Optional[torch.Tensor]]:
        keys = list(keys)
        fetched: "queue.Queue" = queue.Queue()

        def network_worker():
            for key in keys:
                data = None
                if self.contains(key):
                    data = self.connection.get(self._combine_key(key))
                fetched.put(data)
                if data is None:
                    break

        th = threading.Thread(target=network_worker, daemon=True)
        th.start()
        results: List[Optional[torch.Tensor]] = []
        for _ in keys:
            data = fetched.get()
            if data is None or len(data) == 0:
                results.append(None)
                break
            results.append(self.deserializer.from_bytes(data).to(self.dst_device))
        th.join()
        return results
