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
        # a directory: the disk tier (LMCLocalDiskBackend, local_backend.py:163-310), here with CacheGen containers
        from lmcache_b200.storage_backend.local_backend import LMCLocalDiskBackend
        return LMCLocalDiskBackend(config, metadata)
    if isinstance(local, str) and isinstance(remote, str):
        from lmcache_b200.storage_backend.hybrid_backend import LMCHybridBackend
        return LMCHybridBackend(config, metadata)
    raise ValueError(f"Invalid configuration: {config}")


__all__ = ["CreateStorageBackend", "LMCBackendInterface"]
