"""Remote key/value connector interface (lmcache/storage_backend/connector/base_connector.py:11-70)."""
import abc
from typing import List, Optional


class RemoteConnector(metaclass=abc.ABCMeta):

    @abc.abstractmethod
    def exists(self, key: str) -> bool:
        raise NotImplementedError

    @abc.abstractmethod
    def get(self, key: str) -> Optional[bytes]:
        """Bytes stored under `key`, or None when absent / on a broken connection."""
        raise NotImplementedError

    @abc.abstractmethod
    def set(self, key: str, obj: bytes) -> None:
        raise NotImplementedError

    @abc.abstractmethod
    def list(self) -> List[str]:
        raise NotImplementedError

    @abc.abstractmethod
    def close(self) -> None:
        raise NotImplementedError
