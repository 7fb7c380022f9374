"""lm:// wire protocol: the native client / server of libb200kv (csrc/lmnet.cu) against the pure-Python ones (which
mirror lmcache/storage_backend/connector/lm_connector.py and lmcache/server/__main__.py byte for byte) in all four
pairings.  Host-only code: runs without a GPU."""
import ctypes
import os
import socket
import threading
import time

import numpy as np
import pytest

import __graft_entry__ as ge

ge.build_cuda()
from lmcache_b200 import _native as N  # noqa: E402
from lmcache_b200.protocol import ClientMetaMessage, Constants, ServerMetaMessage  # noqa: E402
from lmcache_b200.server.__main__ import LMCacheServer  # noqa: E402
from lmcache_b200.storage_backend.connector import CreateConnector  # noqa: E402
from lmcache_b200.storage_backend.connector.lm_connector import LMCServerConnector  # noqa: E402
from lmcache_b200.storage_backend.connector.native_connector import LMCNativeConnector  # noqa: E402


@pytest.fixture(params=["native", "python"])
def server(request):
    if request.param == "native":
        lib = N.lib()
        h = ctypes.c_void_p()
        N.check(lib.b200kv_lm_server_start(b"127.0.0.1", 0, ctypes.byref(h)))
        port = lib.b200kv_lm_server_port(h)
        assert port > 0
        yield ("native", port, h)
        N.check(lib.b200kv_lm_server_stop(h))
    else:
        srv = LMCacheServer("127.0.0.1", 0)
        port = srv.sock.getsockname()[1]
        threading.Thread(target=srv.run, daemon=True).start()
        yield ("python", port, srv)
        srv.sock.close()


@pytest.fixture(params=["native", "python"])
def client(request, server):
    c = (LMCNativeConnector if request.param == "native" else LMCServerConnector)("127.0.0.1", server[1])
    yield c
    c.close()


def _wait_exists(c, key, want=True, tries=200):
    for _ in range(tries):          # PUT has no ack: the server may still be reading the payload
        if c.exists(key) == want:
            return True
        time.sleep(0.005)
    return False


def test_put_get_exists_list_all_pairings(client):
    rng = np.random.default_rng(1)
    blobs = {f"vllm@model/x@1@0@{i:064x}": rng.integers(0, 256, n, dtype=np.uint8).tobytes()
             for i, n in enumerate([0, 1, 157, 158, 65536, 3 * 1024 * 1024 + 5])}
    assert client.get("missing") is None and not client.exists("missing")
    assert client.list() == []
    for k, v in blobs.items():
        client.set(k, v)
    for k, v in blobs.items():
        assert _wait_exists(client, k)
        got = client.get(k)
        assert got is not None and bytes(got) == v and isinstance(got, (bytes, bytearray))
    assert sorted(client.list()) == sorted(blobs)
    k0 = next(iter(blobs))
    client.set(k0, b"replaced")                                   # overwrite
    for _ in range(200):
        if bytes(client.get(k0)) == b"replaced":
            break
        time.sleep(0.005)
    assert bytes(client.get(k0)) == b"replaced"
    # bytes-like payloads: bytearray, memoryview (read-only and writable), numpy buffer
    for i, obj in enumerate([bytearray(b"abc" * 1000), memoryview(b"ro-view" * 99), memoryview(bytearray(b"rw" * 7)),
                             np.arange(1000).astype(np.uint8)]):
        client.set(f"k{i}", obj)
        assert _wait_exists(client, f"k{i}")
        assert bytes(client.get(f"k{i}")) == bytes(memoryview(obj).cast("B"))


def test_key_limits_and_padding(client):
    key = "k" * 150
    client.set(key, b"x")
    assert _wait_exists(client, key)
    with pytest.raises(AssertionError):
        client.set("k" * 151, b"x")
    client.set("with space inside", b"y")        # inner spaces survive, the padding is stripped (Python's .strip())
    assert _wait_exists(client, "with space inside") and "with space inside" in client.list()


def test_native_server_speaks_the_reference_header_bytes():
    """raw socket against the native server: the exact struct layouts of lmcache/protocol.py"""
    lib = N.lib()
    h = ctypes.c_void_p()
    N.check(lib.b200kv_lm_server_start(b"127.0.0.1", 0, ctypes.byref(h)))
    try:
        s = socket.create_connection(("127.0.0.1", lib.b200kv_lm_server_port(h)))
        assert ClientMetaMessage.packlength() == 158 and ServerMetaMessage.packlength() == 8
        s.sendall(ClientMetaMessage(Constants.CLIENT_PUT, "a@b", 5).serialize() + b"hello")
        s.sendall(ClientMetaMessage(Constants.CLIENT_EXIST, "a@b", 0).serialize())
        rep = ServerMetaMessage.deserialize(s.recv(8, socket.MSG_WAITALL))
        assert (rep.code, rep.length) == (Constants.SERVER_SUCCESS, 0)
        s.sendall(ClientMetaMessage(Constants.CLIENT_GET, "a@b", 0).serialize())
        rep = ServerMetaMessage.deserialize(s.recv(8, socket.MSG_WAITALL))
        assert (rep.code, rep.length) == (Constants.SERVER_SUCCESS, 5) and s.recv(5, socket.MSG_WAITALL) == b"hello"
        s.sendall(ClientMetaMessage(Constants.CLIENT_GET, "nope", 0).serialize())
        rep = ServerMetaMessage.deserialize(s.recv(8, socket.MSG_WAITALL))
        assert (rep.code, rep.length) == (Constants.SERVER_FAIL, 0)
        assert lib.b200kv_lm_server_num_keys(h) == 1
        s.close()
    finally:
        N.check(lib.b200kv_lm_server_stop(h))


def test_concurrent_clients_and_threads(server):
    """8 connections x interleaved put / get / exists of 256 KiB values; one shared connection used by 4 threads"""
    port = server[1]
    errs = []

    def worker(i, conn=None):
        c = conn or LMCNativeConnector("127.0.0.1", port)
        try:
            rng = np.random.default_rng(i)
            for r in range(20):
                k, v = f"w{i}_r{r}", rng.integers(0, 256, 262144, dtype=np.uint8).tobytes()
                c.set(k, v)
                assert _wait_exists(c, k)
                assert bytes(c.get(k)) == v
        except Exception as e:          # noqa: BLE001
            errs.append(repr(e))
        finally:
            if conn is None:
                c.close()
    ths = [threading.Thread(target=worker, args=(i,)) for i in range(8)]
    shared = LMCNativeConnector("127.0.0.1", port)
    ths += [threading.Thread(target=worker, args=(100 + i, shared)) for i in range(4)]
    [t.start() for t in ths]
    [t.join() for t in ths]
    shared.close()
    assert not errs, errs


def test_factory_and_errors():
    with pytest.raises(ValueError):
        CreateConnector("redis://127.0.0.1:6379")
    with pytest.raises(ValueError):
        CreateConnector("lm://nohostport")
    with pytest.raises(N.NativeError):
        LMCNativeConnector("127.0.0.1", 1)                       # nothing listens there
    lib = N.lib()
    h = ctypes.c_void_p()
    N.check(lib.b200kv_lm_server_start(b"127.0.0.1", 0, ctypes.byref(h)))
    port = lib.b200kv_lm_server_port(h)
    c = CreateConnector(f"lmn://127.0.0.1:{port}")
    assert isinstance(c, LMCNativeConnector)
    p = CreateConnector(f"lm://localhost:{port}")
    assert isinstance(p, LMCServerConnector)
    c.set("x", b"1")
    assert _wait_exists(p, "x") and bytes(p.get("x")) == b"1"
    # protocol misuse is an error, not a hang: read without begin, wrong length
    assert lib.b200kv_lm_read(c._h, None, 5) < 0
    n = lib.b200kv_lm_get_begin(c._h, b"x")
    assert n == 1 and lib.b200kv_lm_put(c._h, b"y", None, 0) < 0      # payload pending
    buf = ctypes.create_string_buffer(1)
    assert lib.b200kv_lm_read(c._h, buf, 1) == 0 and buf.raw == b"1"
    c.close(); p.close()
    N.check(lib.b200kv_lm_server_stop(h))
    assert c.get("x") is None and not c.exists("x")              # closed connector: miss, not a crash


@pytest.mark.parametrize("kind", ["native", "python"])
def test_get_into_caller_memory(server, kind):
    c = (LMCNativeConnector if kind == "native" else LMCServerConnector)("127.0.0.1", server[1])
    v = bytes(range(256)) * 64
    c.set("blob", v)
    assert _wait_exists(c, "blob")
    buf = ctypes.create_string_buffer(len(v) + 8)
    assert c.get_into("blob", ctypes.addressof(buf), len(buf)) == len(v) and buf.raw[:len(v)] == v
    assert c.get_into("nope", ctypes.addressof(buf), len(buf)) is None
    assert c.get_into("blob", ctypes.addressof(buf), 10) is None          # too small: drained, reported as a miss
    assert bytes(c.get("blob")) == v                                      # ... and the connection is still in step
    c.close()

