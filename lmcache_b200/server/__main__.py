"""lm:// cache server: `python -m lmcache_b200.server <host> <port> [--python]`.
Default: the native server of libb200kv (csrc/lmnet.cu); `--python` runs the pure-Python one below.

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


def run_native(host: str, port: int) -> None:
    """The same server in C++ (csrc/lmnet.cu): thread per client, O(1) EXIST / GET under a reader-writer lock,
    payloads received into / sent from one buffer each.  Runs until the process is terminated."""
    import ctypes
    import signal
    import time

    from lmcache_b200 import _native as N
    lib = N.lib()
    h = ctypes.c_void_p()
    N.check(lib.b200kv_lm_server_start(host.encode(), port, ctypes.byref(h)), "lm_server_start")
    stop = []
    signal.signal(signal.SIGTERM, lambda *_: stop.append(1))
    try:
        while not stop:
            time.sleep(0.2)
    except KeyboardInterrupt:
        pass
    lib.b200kv_lm_server_stop(h)


def main():
    args = [a for a in sys.argv[1:] if a != "--python"]
    if len(args) not in (2, 3):
        print(f"Usage: {sys.argv[0]} <host> <port> [cpu] [--python]")
        sys.exit(1)
    if "--python" in sys.argv[1:]:
        LMCacheServer(args[0], int(args[1])).run()
    else:
        run_native(args[0], int(args[1]))


if __name__ == "__main__":
    main()
