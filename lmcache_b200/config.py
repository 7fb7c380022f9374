"""Engine configuration / metadata types -- same names, fields, constructors and validation
as lmcache/config.py:8-139 so existing callers (lmcache-vllm adapter, YAML files) keep working."""
from __future__ import annotations

import re
from dataclasses import dataclass
from typing import Optional

import yaml

_DISK_RE = re.compile(r"file://(.*)/")
_URL_RE = re.compile(r"(.*)://(.*):(\d+)")


@dataclass
class LMCacheEngineMetadata:
    model_name: str   # LLM name; selects the CacheGen bin table
    world_size: int   # tensor-parallel world size (part of the key)
    worker_id: int    # tensor-parallel rank (part of the key)
    fmt: str          # "vllm" | "huggingface"
    dtype: str        # dtype of the kv tensors


@dataclass
class LMCacheEngineConfig:
    chunk_size: int
    local_device: Optional[str]
    remote_url: Optional[str]
    remote_serde: Optional[str]   # "torch" | "cachegen" | ...
    pipelined_backend: bool
    save_decode_cache: bool
    # not in the reference: how the LOCAL host tier keeps chunks.  None = raw blobs, as the reference does
    # (local_backend.py:95-100); "cachegen" = CacheGen containers in page-locked memory (LMCLocalCompressedBackend).
    # The environment variable LMCACHE_B200_LOCAL_SERDE sets the default for configurations that do not name it.
    local_serde: Optional[str] = None

    def __post_init__(self):
        if self.local_serde is None:
            import os
            self.local_serde = os.environ.get("LMCACHE_B200_LOCAL_SERDE") or None
        if self.local_serde not in (None, "cachegen"):
            raise ValueError(f"Invalid local serde: {self.local_serde}")

    @staticmethod
    def from_defaults(chunk_size: int = 256, local_device: str = "cuda",
                      remote_url: str = "redis://localhost:6379", remote_serde: str = "torch",
                      pipelined_backend: bool = False, save_decode_cache: bool = False,
                      local_serde: Optional[str] = None) -> "LMCacheEngineConfig":
        return LMCacheEngineConfig(chunk_size, local_device, remote_url, remote_serde, pipelined_backend,
                                   save_decode_cache, local_serde)

    @staticmethod
    def from_legacy(chunk_size: int = 256, backend: str = "cuda", persist_path: Optional[str] = None,
                    remote_serde: Optional[str] = "torch", pipelined_backend: bool = False,
                    save_decode_cache: bool = False, local_serde: Optional[str] = None) -> "LMCacheEngineConfig":
        """backend: "cpu" | "cuda" | "file://<dir>/" | "<scheme>://<host>:<port>" (config.py:51-82)."""
        local_device: Optional[str] = None
        remote_url: Optional[str] = None
        if backend in ("cpu", "cuda"):
            local_device = backend
        elif _DISK_RE.match(backend):
            local_device = backend[7:]
        elif _URL_RE.match(backend):
            remote_url = backend
        return LMCacheEngineConfig(chunk_size, local_device, remote_url, remote_serde, pipelined_backend,
                                   save_decode_cache, local_serde)

    @staticmethod
    def from_file(file_path: str) -> "LMCacheEngineConfig":
        """YAML loader with the reference's validation rules (config.py:84-124)."""
        with open(file_path, "r") as fin:
            cfg = yaml.safe_load(fin)
        chunk_size = cfg.get("chunk_size", 256)
        local_device = cfg.get("local_device", None)
        remote_url = cfg.get("remote_url", None)
        remote_serde = cfg.get("remote_serde", "torch")
        pipelined_backend = cfg.get("pipelined_backend", False)
        save_decode_cache = cfg.get("save_decode_cache", False)
        local_serde = cfg.get("local_serde", None)

        if local_device in ("cpu", "cuda", None):
            pass
        elif isinstance(local_device, str) and _DISK_RE.match(local_device):
            local_device = local_device[7:]
        else:
            raise ValueError(f"Invalid local storage device: {local_device}")

        if remote_url is not None and not (isinstance(remote_url, str) and _URL_RE.match(remote_url)):
            raise ValueError(f"Invalid remote storage url: {remote_url}")

        return LMCacheEngineConfig(chunk_size, local_device, remote_url, remote_serde, pipelined_backend,
                                   save_decode_cache, local_serde)


class GlobalConfig:
    """Process-wide switches (config.py:130-139)."""
    enable_debug: bool = True

    @classmethod
    def set_debug(cls, enable: bool):
        cls.enable_debug = enable

    @classmethod
    def is_debug(cls) -> bool:
        return cls.enable_debug
