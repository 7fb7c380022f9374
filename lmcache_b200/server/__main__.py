"""Minimal lm:// cache server: `python -m lmcache_b200.server <host> <port>`.

Speaks the reference wire protocol (lmcache/protocol.py, lmcache/server/__main__.py:29-93): opaque bytes
in an in-memory dict, thread per client, no ack on PUT.  It never touches KV math; it exists so the
engine-over-socket configurations can be exercised on a box that does not have the reference tree.
EXIST is a dict lookup (the reference scans list_keys(), server/__main__.py:78-80)."""
import socket
import sys
import threading

from lmcache_b200.protocol import ClientMetaMessage, Constants, ServerMetaMessage


class LMCacheServer:

    def __init__(self, host: str, port: int):
        self.store = {}
        self.lock = threading.Lock()
        self.sock = socket.socket(socket.AF_INET, socket.SOCK_STREAM)
        self.sock.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self.sock.bind((host, port))
        self.sock.listen()

    @staticmethod
    def _recv_exact(conn, n):
        buf = bytearray(n)
        view, got = memoryview(buf), 0
        while got < n:
            k = conn.recv_into(view[got:], n - got)
            if k == 0:
                return None
            got += k
        return buf

    def handle_client(self, conn):
        try:
            while True:
                header = self._recv_exact(conn, ClientMetaMessage.packlength())
                if header is None:
                    break
                meta = ClientMetaMessage.deserialize(bytes(header))
                if meta.command == Constants.CLIENT_PUT:
                    data = self._recv_exact(conn, meta.length)
                    if data is None:
                        break
                    with self.lock:
                        self.store[meta.key] = data
                elif meta.command == Constants.CLIENT_GET:
                    with self.lock:
                        data = self.store.get(meta.key)
                    if data is None:
                        conn.sendall(ServerMetaMessage(Constants.SERVER_FAIL, 0).serialize())
                    else:
                        conn.sendall(ServerMetaMessage(Constants.SERVER_SUCCESS, len(data)).serialize())
                        conn.sendall(data)
                elif meta.command == Constants.CLIENT_EXIST:
                    with self.lock:
                        ok = meta.key in self.store
                    conn.sendall(ServerMetaMessage(Constants.SERVER_SUCCESS if ok else Constants.SERVER_FAIL,
                                                   0).serialize())
                elif meta.command == Constants.CLIENT_LIST:
                    with self.lock:
                        data = "\n".join(self.store.keys()).encode()
                    conn.sendall(ServerMetaMessage(Constants.SERVER_SUCCESS, len(data)).serialize())
                    conn.sendall(data)
                else:
                    break
        finally:
            conn.close()

    def run(self):
        try:
            while True:
                conn, _ = self.sock.accept()
                threading.Thread(target=self.handle_client, args=(conn,), daemon=True).start()
        finally:
            self.sock.close()


def main():
    if len(sys.argv) not in (3, 4):
        print(f"Usage: {sys.argv[0]} <host> <port> [cpu]")
        sys.exit(1)
    LMCacheServer(sys.argv[1], int(sys.argv[2])).run()


if __name__ == "__main__":
    main()
