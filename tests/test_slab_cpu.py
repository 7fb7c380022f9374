"""CPU: the page-locked slab's allocator (first fit, coalescing, oversized requests) on plain memory."""
import numpy as np

from lmcache_b200.slab import ALIGN, PinnedSlab


class _FakeSeg:
    def __init__(self, n):
        self.nbytes = n
        self.buf = np.zeros(n, np.uint8)
        self.host_ptr = self.buf.ctypes.data
        self.dev_ptr = self.host_ptr
        self.closed = False

    def view(self, off, n):
        return memoryview(self.buf)[off:off + n]

    def close(self):
        self.closed = True


def test_slab_alloc_free_coalesce_and_reuse():
    slab = PinnedSlab(segment_bytes=1 << 20, alloc_fn=_FakeSeg)
    a = slab.alloc(1000)
    b = slab.alloc(5000)
    c = slab.alloc(300)
    assert (a.seg, b.seg, c.seg) == (0, 0, 0)
    assert a.offset % ALIGN == 0 and b.offset % ALIGN == 0 and c.offset % ALIGN == 0
    assert b.offset >= a.offset + 1000 and c.offset >= b.offset + 5000
    a.view()[:4] = b"abcd"
    assert bytes(slab._segs[0].buf[a.offset:a.offset + 4]) == b"abcd" and a.host_ptr == slab._segs[0].host_ptr + a.offset
    used = slab.stats()[2]
    b.free()
    b.free()                                   # idempotent
    assert slab.stats()[2] == used - 5120
    d = slab.alloc(4000)                       # first fit: lands in b's hole
    assert d.offset == b.offset
    a.free(); c.free(); d.free()
    assert slab.stats()[2] == 0
    assert len(slab._free[0].offs) == 1 and slab._free[0].lens[0] == 1 << 20      # everything coalesced again
    big = slab.alloc(3 << 20)                  # larger than a segment: gets its own
    assert big.seg == 1 and slab.stats()[0] == 2
    slab.reserve(4 << 20)
    assert slab.stats()[0] >= 5
    slab.close()
    assert slab.stats() == (0, 0, 0)


def test_slab_random_stress_no_overlap():
    rng = np.random.default_rng(0)
    slab = PinnedSlab(segment_bytes=1 << 18, alloc_fn=_FakeSeg)
    live = []
    for step in range(3000):
        if live and rng.random() < 0.45:
            live.pop(int(rng.integers(len(live)))).free()
        else:
            n = int(rng.integers(1, 40000))
            blk = slab.alloc(n)
            for o in live:
                if o.seg == blk.seg:
                    assert blk.offset >= o.offset + o.cap or o.offset >= blk.offset + blk.cap
            live.append(blk)
    total = sum(b.cap for b in live)
    assert slab.stats()[2] == total
    for b in live:
        b.free()
    assert slab.stats()[2] == 0
    for fl, seg in zip(slab._free, slab._segs):
        assert fl.offs == [0] and fl.lens == [seg.nbytes]


def test_shrink_returns_the_tail():
    """a receive block is reserved for the largest possible container and trimmed to what arrived"""
    slab = PinnedSlab(segment_bytes=1 << 20, alloc_fn=_FakeSeg)
    a = slab.alloc(600_000)
    used = slab.bytes_in_use
    a.shrink(10_000)
    assert slab.bytes_in_use < used and 10_000 <= a.cap < 600_000 and a.nbytes <= a.cap
    b = slab.alloc(500_000)                    # fits next to it in the same 1 MiB segment only after the shrink
    assert b.seg == a.seg
    a.shrink(1_000_000)                        # growing is not a thing: no-op
    a.free()
    b.free()
    assert slab.bytes_in_use == 0
