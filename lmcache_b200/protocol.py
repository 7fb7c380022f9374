"""lm:// wire headers -- byte-compatible with lmcache/protocol.py:4-70 so the reference's
`python -m lmcache.server` and this package's client/server interoperate.

client -> server: struct "ii150s" = (command, payload length, key padded to 150 bytes)  (158 bytes)
server -> client: struct "ii"     = (status code, payload length)                        (8 bytes)
"""
import struct
from dataclasses import dataclass

MAX_KEY_LENGTH = 150
_CLIENT_FMT = f"ii{MAX_KEY_LENGTH}s"
_SERVER_FMT = "ii"


class Constants:
    CLIENT_PUT = 1
    CLIENT_GET = 2
    CLIENT_EXIST = 3
    CLIENT_LIST = 4

    SERVER_SUCCESS = 200
    SERVER_FAIL = 400


@dataclass
class ClientMetaMessage:
    command: int
    key: str
    length: int

    def serialize(self) -> bytes:
        assert len(self.key) <= MAX_KEY_LENGTH, f"Key length {len(self.key)} exceeds maximum {MAX_KEY_LENGTH}"
        return struct.pack(_CLIENT_FMT, self.command, self.length, self.key.encode().ljust(MAX_KEY_LENGTH))

    @staticmethod
    def deserialize(s: bytes) -> "ClientMetaMessage":
        command, length, key = struct.unpack(_CLIENT_FMT, s)
        return ClientMetaMessage(command, key.decode().strip(), length)

    @staticmethod
    def packlength() -> int:
        return struct.calcsize(_CLIENT_FMT)


@dataclass
class ServerMetaMessage:
    code: int
    length: int

    def serialize(self) -> bytes:
        return struct.pack(_SERVER_FMT, self.code, self.length)

    @staticmethod
    def packlength() -> int:
        return struct.calcsize(_SERVER_FMT)

    @staticmethod
    def deserialize(s: bytes) -> "ServerMetaMessage":
        code, length = struct.unpack(_SERVER_FMT, s)
        return ServerMetaMessage(code, length)
