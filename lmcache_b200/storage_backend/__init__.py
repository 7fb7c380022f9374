"""Backend factory (lmcache/storage_backend/__init__.py:13-44): same (local_device, remote_url) dispatch."""
from lmcache_b200.config import LMCacheEngineConfig, LMCacheEngineMetadata
from lmcache_b200.storage_backend.abstract_backend import LMCBackendInterface


def CreateStorageBackend(config: LMCacheEngineConfig, metadata: LMCacheEngineMetadata) -> LMCBackendInterface:
    local, remote = config.local_device, config.remote_url
    if local is None and isinstance(remote, str):
        from lmcache_b200.storage_backend.remote_backend import LMCPipelinedRemoteBackend, LMCRemoteBackend
        return (LMCPipelinedRemoteBackend if config.pipelined_backend else LMCRemoteBackend)(config, metadata)
    if isinstance(local, str) and remote is None:
        if local == "cpu" and config.local_serde == "cachegen":
            from lmcache_b200.storage_backend.local_backend import LMCLocalCompressedBackend
            return LMCLocalCompressedBackend(config, metadata)
        if local in ("cpu", "cuda"):
            from lmcache_b200.storage_backend.local_backend import LMCLocalBackend
            return LMCLocalBackend(config)
        # disk tier (LMCLocalDiskBackend, local_backend.py:163-310) is file I/O, not GPU path: SURVEY 8(f) "next" row 4
        raise ValueError(f"Invalid configuration: local disk backend '{local}' is not provided by lmcache_b200")
    if isinstance(local, str) and isinstance(remote, str):
        # hybrid = write-through composition of local + remote (hybrid_backend.py): outside the rebuilt hot path
        raise ValueError("Invalid configuration: hybrid (local + remote) backend is not provided by lmcache_b200")
    raise ValueError(f"Invalid configuration: {config}")


__all__ = ["CreateStorageBackend", "LMCBackendInterface"]
