"""The streaming FIFO (galileo-sdr-sim_b200/host/e1_fifo.h) against the reference's contract
(src/fifo.cpp:3-61, producer at src/galileo-sdr.cpp:581-596): a Python model of those functions with
the same head/tail arithmetic must see the same lengths and the same samples for any interleaving of
block writes and odd-sized reads, including reads across the wrap; a blocked writer resumes when the
reader frees a block; finish() wakes everybody."""
import ctypes as C
import sys
import threading
import time
from pathlib import Path

import numpy as np
import pytest

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "galileo-sdr-sim_b200"))
import build as B  # noqa: E402


@pytest.fixture(scope="module")
def lib():
    l = C.CDLL(str(B.build_host()))
    l.e1_fifo_create.argtypes = [C.c_size_t, C.c_void_p]
    l.e1_fifo_create.restype = C.c_void_p
    l.e1_fifo_destroy.argtypes = [C.c_void_p]
    for f in ("e1_fifo_sample_length",):
        getattr(l, f).argtypes = [C.c_void_p]
        getattr(l, f).restype = C.c_size_t
    for f in ("e1_fifo_read", "e1_fifo_read_wait"):
        getattr(l, f).argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
        getattr(l, f).restype = C.c_size_t
    l.e1_fifo_write_ready.argtypes = [C.c_void_p, C.c_size_t]
    l.e1_fifo_write.argtypes = [C.c_void_p, C.c_void_p, C.c_size_t]
    l.e1_fifo_finish.argtypes = [C.c_void_p]
    l.e1_fifo_finished.argtypes = [C.c_void_p]
    return l


class RefFifo:
    """src/fifo.cpp with FIFO_LENGTH = 2 * NUM_IQ_SAMPLES (include/constants.h:83), transcribed as a model."""

    def __init__(self, block):
        self.block, self.length = block, 2 * block
        self.fifo = np.zeros(self.length * 2, np.int16)
        self.head = self.tail = 0

    def sample_length(self):
        n = self.head - self.tail
        return n + self.length if n < 0 else n

    def write_ready(self):
        return self.sample_length() < self.block

    def write(self, iq):
        self.fifo[self.head * 2:(self.head + self.block) * 2] = iq
        self.head = (self.head + self.block) % self.length

    def read(self, samples):
        samples = min(samples, self.sample_length())
        out = np.zeros(samples * 2, np.int16)
        rem = self.length - self.tail
        done = 0
        if samples > rem:
            out[:rem * 2] = self.fifo[self.tail * 2:]
            self.tail, done, samples = 0, rem, samples - rem
        out[done * 2:(done + samples) * 2] = self.fifo[self.tail * 2:(self.tail + samples) * 2]
        self.tail = (self.tail + samples) % self.length
        return out


def test_same_samples_as_the_reference_ring(lib):
    block = 1000
    ref = RefFifo(block)
    f = lib.e1_fifo_create(2 * block, None)
    rng = np.random.default_rng(3)
    for step in range(400):
        # the reference's ring cannot tell full from empty (head == tail), so its producer only writes
        # while fewer than one block is queued; follow that discipline here
        if ref.write_ready() and rng.random() < 0.6:
            iq = rng.integers(-3000, 3000, block * 2).astype(np.int16)
            assert lib.e1_fifo_write_ready(f, block) == 1
            ref.write(iq)
            assert lib.e1_fifo_write(f, iq.ctypes.data, block) == 0
        else:
            want = int(rng.integers(0, 1500))
            a = ref.read(want)
            buf = np.zeros(want * 2 + 2, np.int16)
            got = lib.e1_fifo_read(f, buf.ctypes.data, want)
            assert got * 2 == len(a) and np.array_equal(buf[:got * 2], a), step
        assert lib.e1_fifo_sample_length(f) == ref.sample_length()
    lib.e1_fifo_destroy(f)


def test_writer_blocks_until_the_reader_makes_room_and_finish_wakes(lib):
    block = 5000
    f = lib.e1_fifo_create(2 * block, None)
    data = [np.full(block * 2, k + 1, np.int16) for k in range(6)]
    t_written = []

    def producer():
        for d in data:
            assert lib.e1_fifo_write(f, d.ctypes.data, block) == 0
            t_written.append(time.time())
        lib.e1_fifo_finish(f)

    th = threading.Thread(target=producer)
    th.start()
    time.sleep(0.3)
    assert len(t_written) == 2 and lib.e1_fifo_write_ready(f, block) == 0     # full: the third write waits
    out = []
    buf = np.zeros(3000 * 2, np.int16)
    while True:
        got = lib.e1_fifo_read_wait(f, buf.ctypes.data, 3000)
        if got == 0 and lib.e1_fifo_finished(f):
            break
        out.append(buf[:got * 2].copy())
    th.join()
    assert np.array_equal(np.concatenate(out), np.concatenate(data))
    lib.e1_fifo_destroy(f)


def test_pinned_storage_from_the_caller(lib):
    store = np.zeros(64 * 2, np.int16)
    f = lib.e1_fifo_create(64, store.ctypes.data)
    iq = np.arange(40 * 2, dtype=np.int16)
    assert lib.e1_fifo_write(f, iq.ctypes.data, 40) == 0
    assert np.array_equal(store[:80], iq)
    out = np.zeros(80, np.int16)
    assert lib.e1_fifo_read(f, out.ctypes.data, 40) == 40
    assert lib.e1_fifo_write(f, iq.ctypes.data, 40) == 0                      # wraps: 24 at the end, 16 at the start
    assert np.array_equal(store[80:128], iq[:48]) and np.array_equal(store[:32], iq[48:])
    assert lib.e1_fifo_write(f, iq.ctypes.data, 65) == -1                     # larger than the ring
    lib.e1_fifo_destroy(f)
