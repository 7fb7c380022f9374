"""Wave pipelines between the codec and a host-side tier: encode || device->host on the way out, host->device ||
decode on the way in (BASELINE configs[2]).

A store is cut into waves of a few chunks.  Each wave owns a *slot*: a device staging area for its containers and a
page-locked array for their sizes.  The caller's thread only enqueues kernels (CacheGenCodec.encode_async on the
caller's stream: the KV is consumed in stream order, nothing is synchronised) and hands the slot to a worker thread;
the worker waits for the wave's event, learns the container sizes, and moves exactly those bytes to the tier (device ->
page-locked slab on a copy stream, or a socket).  While it does, the next wave is already encoding.  Scratch is bounded
by slots x wave size instead of the whole block (round 1 staged every chunk of a store at once: ~4 GB of device scratch
for a 4 GiB block).

The reference does none of this: LMCLocalBackend.put_nonblocking hands whole chunk tensors to a queue and its worker
calls tensor.to("cpu") + torch.cuda.synchronize() per chunk (lmcache/storage_backend/local_backend.py:82-117).
"""
from __future__ import annotations

import os
import queue
import threading
from typing import Callable, List, Optional, Sequence

import torch

from lmcache_b200 import _native as N
from lmcache_b200.codec import CacheGenCodec, EncodeTicket, KvView, PinnedBuffer


def wave_chunks_default() -> int:
    return max(1, int(os.environ.get("LMCACHE_B200_WAVE_CHUNKS", "4")))


def wave_slots_default() -> int:
    return max(2, int(os.environ.get("LMCACHE_B200_WAVE_SLOTS", "3")))


class WaveSlot:
    """Device staging + sizes for one wave of at most `wave` chunks of one geometry."""

    def __init__(self, nbytes: int, wave: int, device):
        self.dev = torch.empty(nbytes + N.READ_SLACK, dtype=torch.uint8, device=device)
        self.sizes = PinnedBuffer(max(64, 8 * wave))
        self.ticket: Optional[EncodeTicket] = None

    def close(self):
        self.sizes.close()
        self.dev = None


class EncodeRing:
    """A fixed set of WaveSlots for one (L, H, D, chunk_size, device); acquire() blocks the calling host thread while
    every slot is still in flight -- the only back-pressure of the store pipeline."""

    def __init__(self, codec: CacheGenCodec, L: int, H: int, D: int, chunk_size: int, device, wave: Optional[int] = None,
                 slots: Optional[int] = None):
        self.codec = codec
        self.geom = (L, H, D, chunk_size, torch.device(device))
        self.wave = wave or wave_chunks_default()
        self.stride = codec.out_stride(L, H, D, chunk_size)
        n = slots or wave_slots_default()
        self._free: "queue.Queue[WaveSlot]" = queue.Queue()
        self._all: List[WaveSlot] = []
        for _ in range(n):
            s = WaveSlot(self.stride * self.wave, self.wave, device)
            self._all.append(s)
            self._free.put(s)

    def matches(self, L, H, D, chunk_size, device) -> bool:
        return self.geom == (L, H, D, chunk_size, torch.device(device))

    def scratch_bytes(self) -> int:
        return sum(s.dev.numel() for s in self._all)

    def acquire(self) -> WaveSlot:
        return self._free.get()

    def release(self, slot: WaveSlot) -> None:
        slot.ticket = None
        self._free.put(slot)

    def drain(self) -> None:
        """Wait until every slot is back (no wave in flight)."""
        got = [self._free.get() for _ in self._all]
        for s in got:
            self._free.put(s)

    def close(self) -> None:
        self.drain()
        for s in self._all:
            s.close()
        self._all = []


class StoreJob:
    """Completion of one put_kv_chunks call: counts its waves; `wait()` blocks until the sink took all of them."""

    def __init__(self, n_waves: int):
        self._left = n_waves
        self._cv = threading.Condition()
        self.error: Optional[BaseException] = None

    def wave_done(self, err: Optional[BaseException] = None) -> None:
        with self._cv:
            if err is not None and self.error is None:
                self.error = err
            self._left -= 1
            if self._left <= 0:
                self._cv.notify_all()

    def wait(self) -> None:
        with self._cv:
            while self._left > 0:
                self._cv.wait()
        if self.error is not None:
            raise self.error


class EncodePipeline:
    """encode waves on the caller's stream; a worker thread hands every finished wave to `sink`.

    sink(slot, batch, first_chunk, items) is called on the worker thread with the wave's EncodedBatch (sizes known,
    containers still in slot.dev) and items = the caller's per-chunk payload (keys); it must be done with slot.dev
    when it returns -- the slot is recycled right after."""

    def __init__(self, codec: CacheGenCodec, sink: Callable, name: str = "b200kv-store"):
        self.codec = codec
        self.sink = sink
        self.ring: Optional[EncodeRing] = None
        self._q: "queue.Queue" = queue.Queue()
        self._device = torch.cuda.current_device()
        self._thread = threading.Thread(target=self._worker, name=name, daemon=True)
        self._thread.start()
        self._closed = False

    # ------------------------------------------------------------------ caller side
    def _ring_for(self, view: KvView, chunk_size: int) -> EncodeRing:
        if self.ring is None or not self.ring.matches(view.L, view.H, view.D, chunk_size, view.device):
            if self.ring is not None:
                self.ring.close()
            self.ring = EncodeRing(self.codec, view.L, view.H, view.D, chunk_size, view.device)
        return self.ring

    def submit(self, view: KvView, tok_begin: int, chunk_size: int, items: Sequence,
               stream: Optional[torch.cuda.Stream] = None) -> StoreJob:
        """Enqueue the encode of tokens [tok_begin, T) of `view`, len(items) chunks, in waves.  Returns once every wave
        is enqueued on `stream` (default: the current stream); the job completes when the sink has taken them all."""
        n_tok = view.ntokens - tok_begin
        n_chunks = len(items)
        assert n_chunks == (n_tok + chunk_size - 1) // chunk_size and n_chunks > 0
        ring = self._ring_for(view, chunk_size)
        W = ring.wave
        job = StoreJob((n_chunks + W - 1) // W)
        with torch.cuda.device(view.device):
            for c0 in range(0, n_chunks, W):
                k = min(W, n_chunks - c0)
                t0 = tok_begin + c0 * chunk_size
                nt = min(k * chunk_size, view.ntokens - t0)
                slot = ring.acquire()
                try:
                    slot.ticket = self.codec.encode_async(view, t0, nt, chunk_size, stream, out=slot.dev, sizes=slot.sizes)
                except BaseException as e:       # noqa: BLE001 -- give the slot back, fail the job, re-raise
                    ring.release(slot)
                    job.wave_done(e)
                    raise
                self._q.put((ring, slot, c0, list(items[c0:c0 + k]), job))
        return job

    # ------------------------------------------------------------------ worker side
    def _worker(self) -> None:
        torch.cuda.set_device(self._device)
        while True:
            item = self._q.get()
            if item is None:
                return
            ring, slot, c0, items, job = item
            err = None
            try:
                batch = slot.ticket.wait()          # host wait on this wave's kernels -- on the worker thread only
                self.sink(slot, batch, c0, items)
            except BaseException as e:              # noqa: BLE001 -- a failed background store is a miss later
                err = e
            finally:
                ring.release(slot)
                job.wave_done(err)

    def close(self) -> None:
        if self._closed:
            return
        self._closed = True
        self._q.put(None)
        self._thread.join()
        if self.ring is not None:
            self.ring.close()
            self.ring = None


class UploadRing:
    """Device staging for containers on their way in: two slots filled by host->device copies on a copy stream while the
    decoder works on the other one.  Slot reuse is ordered by events on the streams -- the host never waits."""

    def __init__(self, device):
        self.device = torch.device(device)
        self.copy_stream = torch.cuda.Stream(device=device)
        self._bufs: List[Optional[torch.Tensor]] = [None, None]
        self._busy: List[Optional[torch.cuda.Event]] = [None, None]     # decode that last read the slot
        self._i = 0

    def next_slot(self, nbytes: int):
        """(slot index, device buffer of >= nbytes + read slack); the copy stream already waits for the slot's last reader"""
        i = self._i
        self._i ^= 1
        need = nbytes + N.READ_SLACK
        if self._busy[i] is not None:
            if self._bufs[i] is None or self._bufs[i].numel() < need:
                self._busy[i].synchronize()          # growing: the old buffer must not be freed under a kernel
            else:
                self.copy_stream.wait_event(self._busy[i])
        if self._bufs[i] is None or self._bufs[i].numel() < need:
            self._bufs[i] = torch.empty(max(need, need * 5 // 4), dtype=torch.uint8, device=self.device)
        return i, self._bufs[i]

    def mark_read(self, i: int, stream: torch.cuda.Stream) -> None:
        ev = torch.cuda.Event()
        ev.record(stream)
        self._busy[i] = ev


def fetch_decode(codec: CacheGenCodec, upload: UploadRing, futures: Sequence, dst: KvView, dst_tok0: int, chunk_size: int,
                 inflight: list, on_more: Optional[Callable[[int], None]] = None, wave: Optional[int] = None) -> int:
    """Consume container fetches IN ORDER -- futures[i].result() is (slab block, nbytes) or None for a miss; an entry may
    also be such a tuple directly -- and, wave by wave, upload the containers on the copy stream and decode them into
    `dst` on the current stream (chunk i lands at token dst_tok0 + i * chunk_size).  Fetches of later chunks keep running
    in their pool while earlier waves upload and decode: network / disk || H2D || decode.  Stops at the first miss or
    damaged / mismatching container; everything fetched past that point is released.  Blocks of uploaded waves are
    appended to `inflight` as (event, [blocks]) for the caller to free once the event has completed.
    `on_more(i)` is called before chunk i is awaited (lets the caller keep a window of fetches in flight)."""
    import ctypes

    from lmcache_b200.codec import parse_header
    W = wave or wave_chunks_default()
    lib = N.lib()
    n_done = 0
    wave_items: list = []                  # (block, nbytes, header, chunk index)
    with torch.cuda.device(dst.device):
        cur = torch.cuda.current_stream()

        def flush():
            if not wave_items:
                return
            offs, o = [], 0
            for _, n, _, _ in wave_items:
                offs.append(o)
                o += (n + 15) & ~15
            slot, buf = upload.next_slot(o)
            for (blk, n, _, _), off in zip(wave_items, offs):
                N.check(lib.b200kv_copy_async(ctypes.c_void_p(buf.data_ptr() + off), ctypes.c_void_p(blk.host_ptr), n,
                                              upload.copy_stream.cuda_stream), "copy_async")
            ev = torch.cuda.Event()
            ev.record(upload.copy_stream)
            inflight.append((ev, [w[0] for w in wave_items]))
            cur.wait_event(ev)
            h0 = wave_items[0][2]
            codec.decode_raw(buf.data_ptr(), buf.numel(), offs, [w[1] for w in wave_items],
                             [int(w[2].ntokens) for w in wave_items], dst,
                             [dst_tok0 + w[3] * chunk_size for w in wave_items], int(h0.max_dtype), int(h0.version) - 1, cur)
            upload.mark_read(slot, cur)
            wave_items.clear()

        stop_at = len(futures)
        for i in range(len(futures)):
            if on_more is not None:
                on_more(i)
            f = futures[i]
            got = f.result() if hasattr(f, "result") else f
            if got is None:
                stop_at = i
                break
            blk, n = got
            try:
                hd = parse_header(blk.view()[:n])
                ok = codec.accepts(hd) and (hd.L, hd.H, hd.D) == (dst.L, dst.H, dst.D) and \
                    dst_tok0 + i * chunk_size + hd.ntokens <= dst.ntokens and \
                    (not wave_items or (hd.max_dtype, hd.version) == (wave_items[0][2].max_dtype, wave_items[0][2].version))
            except ValueError:
                ok = False                      # damaged container: a miss, not an error
            if not ok:
                blk.free()
                stop_at = i
                break
            wave_items.append((blk, n, hd, i))
            n_done += 1
            if len(wave_items) == W:
                flush()
        flush()
    for f in futures[stop_at + 1:]:              # fetches past the first miss: let them finish, drop their blocks
        r = f.result() if hasattr(f, "result") else f
        if r is not None:
            r[0].free()
    return n_done
