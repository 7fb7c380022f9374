"""lm:// client (lmcache/storage_backend/connector/lm_connector.py:15-84): blocking TCP, fixed headers
from lmcache_b200.protocol.  One lock covers a whole request/response exchange, so concurrent put / get
threads cannot interleave on the socket (the reference locks sends only, see its TODO:1)."""
import ctypes
import socket
import threading
from typing import List, Optional

from lmcache_b200.protocol import ClientMetaMessage, Constants, ServerMetaMessage
from lmcache_b200.storage_backend.connector.base_connector import RemoteConnector


class LMCServerConnector(RemoteConnector):

    def __init__(self, host: str, port: int):
        self.sock = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self.sock.connect((host, port))
        self.lock = threading.Lock()

    def _recv_exact(self, n: int) -> Optional[bytearray]:
        buf = bytearray(n)
        view, got = memoryview(buf), 0
        while got < n:
            k = self.sock.recv_into(view[got:], n - got)
            if k == 0:
                return None
            got += k
        return buf

    def _request(self, command: int, key: str, payload=None) -> None:
        length = 0 if payload is None else len(payload)
        self.sock.sendall(ClientMetaMessage(command, key, length).serialize())
        if payload is not None:
            self.sock.sendall(payload)

    def exists(self, key: str) -> bool:
        with self.lock:
            self._request(Constants.CLIENT_EXIST, key)
            hdr = self._recv_exact(ServerMetaMessage.packlength())
        return hdr is not None and ServerMetaMessage.deserialize(bytes(hdr)).code == Constants.SERVER_SUCCESS

    def set(self, key: str, obj) -> None:
        with self.lock:
            self._request(Constants.CLIENT_PUT, key, obj)   # the server sends no ack for PUT (server/__main__.py:46-48)

    def get(self, key: str) -> Optional[bytes]:
        with self.lock:
            self._request(Constants.CLIENT_GET, key)
            hdr = self._recv_exact(ServerMetaMessage.packlength())
            if hdr is None:
                return None
            meta = ServerMetaMessage.deserialize(bytes(hdr))
            if meta.code != Constants.SERVER_SUCCESS:
                return None
            return self._recv_exact(meta.length)

    def get_into(self, key: str, dst_ptr: int, cap: int) -> Optional[int]:
        """GET straight into caller memory (e.g. a page-locked slab): returns the payload length, None on a miss.
        A payload larger than `cap` is drained and reported as a miss."""
        with self.lock:
            self._request(Constants.CLIENT_GET, key)
            hdr = self._recv_exact(ServerMetaMessage.packlength())
            if hdr is None:
                return None
            meta = ServerMetaMessage.deserialize(bytes(hdr))
            if meta.code != Constants.SERVER_SUCCESS:
                return None
            n = meta.length
            if n > cap:
                self._recv_exact(n)
                return None
            if n == 0:
                return 0
            view = memoryview((ctypes.c_char * n).from_address(dst_ptr)).cast("B")
            got = 0
            while got < n:
                k = self.sock.recv_into(view[got:], n - got)
                if k == 0:
                    return None
                got += k
            return n

    def list(self) -> List[str]:
        with self.lock:
            self._request(Constants.CLIENT_LIST, "")
            hdr = self._recv_exact(ServerMetaMessage.packlength())
            if hdr is None:
                return []
            meta = ServerMetaMessage.deserialize(bytes(hdr))
            if meta.code != Constants.SERVER_SUCCESS or meta.length == 0:
                return []
            data = self._recv_exact(meta.length)
        return [] if data is None else [k for k in bytes(data).decode().split("\n") if k]

    def close(self) -> None:
        try:
            self.sock.close()
        except OSError:
            pass
