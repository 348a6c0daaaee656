"""ctypes binding of the C-ABI in include/e1b200.h (libe1b200.so: sm_100a kernels).

This is the thin Python mirror used by the tests, bench.py and __graft_entry__; the reference
is C++ and its own integration is the patch in INTEGRATION.md.  There is no CPU fallback:
constructing a Synth without the built library or without a CUDA device raises.
"""
import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = Path(os.environ.get("E1B200_LIB") or PKG / "lib" / "libe1b200.so")   # E1B200_LIB: tuning builds (tools/variants.py)

E1_REC_SET_PHASE = 1
PAGE_BYTES = 64
REF_DT = 0.10000002314200000  # src/galileo-sdr.cpp:347

REC_DTYPE = np.dtype([
    ("prn", "<i4"), ("ibit0", "<i4"), ("flags", "<u4"), ("gain_q7", "<i4"),
    ("code_phase0", "<f8"), ("f_code", "<f8"), ("f_carr", "<f8"), ("carr_phase_init", "<f8"),
    ("page_cur", "u1", PAGE_BYTES), ("page_next", "u1", PAGE_BYTES),
])
RANGE_DTYPE = np.dtype([
    ("prn", "<i4"), ("flags", "<u4"),
    ("rho_prev", "<f8"), ("rho_cur", "<f8"), ("grx_sec", "<f8"), ("carr_phase_init", "<f8"),
    ("page_cur", "u1", PAGE_BYTES), ("page_next", "u1", PAGE_BYTES),
])
assert REC_DTYPE.itemsize == 176 and RANGE_DTYPE.itemsize == 168


class Config(C.Structure):
    _fields_ = [("fs_hz", C.c_double), ("samples_per_epoch", C.c_int32), ("max_chan", C.c_int32),
                ("device", C.c_int32), ("flags", C.c_uint32), ("dt_epoch", C.c_double)]


class Timing(C.Structure):
    _fields_ = [("plan_ms", C.c_float), ("synth_ms", C.c_float), ("total_ms", C.c_float),
                ("kernel_launches", C.c_int32), ("synth_launches", C.c_int32)]


class Stats(C.Structure):
    _fields_ = [("exact_samples", C.c_uint64), ("planner_errors", C.c_uint64), ("serial_epochs", C.c_uint64),
                ("hat_epochs", C.c_uint64), ("tile", C.c_int32), ("tiles_per_epoch", C.c_int32),
                ("batch_epochs", C.c_int32), ("plan_epochs", C.c_int32), ("sm_count", C.c_int32),
                ("ctas_per_sm", C.c_int32), ("smem_bytes", C.c_int32), ("synth_kernel", C.c_int32)]

    @property
    def kernel_name(self):
        k, teams = self.synth_kernel & 0xff, self.synth_kernel >> 8
        return {0: "e1_synth_kernel<R>", 1: "e1_synth_cw_kernel<2,2>", 2: "e1_synth_cw_kernel<4,3>", 3: "e1_synth_float_kernel",
                4: f"e1_synth_ev_kernel<{teams}>"}.get(k, "?")


SYMBOLS = [
    "e1b200_create", "e1b200_destroy", "e1b200_set_channel", "e1b200_clear_channel",
    "e1b200_get_carrier_phase", "e1b200_set_carrier_phase", "e1b200_synth_epochs",
    "e1b200_synth_epochs_device", "e1b200_sync", "e1b200_synth_ranges", "e1b200_synth_ranges_device", "e1b200_plan_phases",
    "e1b200_restate", "e1b200_get_timing", "e1b200_get_stats", "e1b200_stream", "e1b200_last_error",
    "e1b200_version", "e1b200_host_alloc", "e1b200_host_free", "e1b200_selftest_any_hit",
    "e1b200_code_wraps", "e1b200_host_register", "e1b200_host_unregister",
    "e1b200_get_carrier_phases", "e1b200_set_carrier_phases", "e1b200_plan_phases_device",
    "e1b200_peer_alloc", "e1b200_peer_open", "e1b200_peer_close", "e1b200_peer_free",
]

_lib = None


class E1B200Error(RuntimeError):
    pass


def load():
    """dlopen libe1b200.so and declare the prototypes.  Raises if the library is not built."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise E1B200Error(f"{LIB_PATH} is missing: run `python galileo-sdr-sim_b200/build.py` (nvcc, sm_100a); "
                          "there is no CPU fallback")
    lib = C.CDLL(str(LIB_PATH))
    vp, dp = C.c_void_p, C.POINTER(C.c_double)
    lib.e1b200_create.argtypes = [C.POINTER(Config), C.POINTER(vp)]
    lib.e1b200_destroy.argtypes = [vp]
    lib.e1b200_set_channel.argtypes = [vp, C.c_int, C.c_int, C.c_double]
    lib.e1b200_clear_channel.argtypes = [vp, C.c_int]
    lib.e1b200_get_carrier_phase.argtypes = [vp, C.c_int, dp]
    lib.e1b200_set_carrier_phase.argtypes = [vp, C.c_int, C.c_double]
    for name in ("e1b200_synth_epochs", "e1b200_synth_epochs_device", "e1b200_synth_ranges", "e1b200_synth_ranges_device"):
        getattr(lib, name).argtypes = [vp, C.c_int, vp, vp]
    lib.e1b200_plan_phases.argtypes = [vp, C.c_int, vp]
    lib.e1b200_sync.argtypes = [vp]
    lib.e1b200_restate.argtypes = [C.c_double] * 4 + [dp, dp, dp, C.POINTER(C.c_int32), C.POINTER(C.c_int32)]
    lib.e1b200_get_timing.argtypes = [vp, C.POINTER(Timing)]
    lib.e1b200_get_stats.argtypes = [vp, C.POINTER(Stats)]
    lib.e1b200_stream.argtypes = [vp]
    lib.e1b200_stream.restype = vp
    lib.e1b200_last_error.argtypes = [vp]
    lib.e1b200_last_error.restype = C.c_char_p
    lib.e1b200_version.restype = C.c_char_p
    lib.e1b200_host_alloc.argtypes = [C.POINTER(vp), C.c_size_t]
    lib.e1b200_host_free.argtypes = [vp]
    lib.e1b200_selftest_any_hit.argtypes = [C.c_int, C.c_int, vp, vp]
    lib.e1b200_code_wraps.argtypes = [C.c_double, C.c_int32, C.c_double, C.c_double, C.POINTER(C.c_int32)]
    lib.e1b200_get_carrier_phases.argtypes = [vp, C.c_int, dp]
    lib.e1b200_set_carrier_phases.argtypes = [vp, C.c_int, dp]
    lib.e1b200_plan_phases_device.argtypes = [vp, C.c_int, vp]
    lib.e1b200_peer_alloc.argtypes = [C.c_int, C.c_size_t, C.POINTER(vp), C.c_char_p]
    lib.e1b200_peer_open.argtypes = [C.c_int, C.c_char_p, C.POINTER(vp)]
    lib.e1b200_peer_close.argtypes = [vp]
    lib.e1b200_peer_free.argtypes = [vp]
    lib.e1b200_host_register.argtypes = [vp, C.c_size_t]
    lib.e1b200_host_unregister.argtypes = [vp]
    _lib = lib
    return lib


def restate(rho_prev, rho_cur, dt, grx_sec):
    """computeCodePhase (src/gal-sig.cpp:308-347) -> (f_carr, f_code, code_phase0, ibit0, ipage)."""
    lib = load()
    f = [C.c_double() for _ in range(3)]
    ib, ip = C.c_int32(), C.c_int32()
    rc = lib.e1b200_restate(rho_prev, rho_cur, dt, grx_sec, f[0], f[1], f[2], ib, ip)
    if rc:
        raise E1B200Error(f"e1b200_restate: {rc}")
    return f[0].value, f[1].value, f[2].value, ib.value, ip.value


def code_wraps(fs_hz, n_samp, code_phase0, f_code):
    """How often the reference's loop wraps the code phase inside one block (src/galileo-sdr.cpp:491-494), exactly."""
    n = C.c_int32()
    rc = load().e1b200_code_wraps(fs_hz, n_samp, code_phase0, f_code, n)
    if rc:
        raise E1B200Error(f"e1b200_code_wraps: {rc}")
    return n.value


class PinnedBuffer:
    """Page-locked host memory from e1b200_host_alloc, viewed as a numpy array."""

    def __init__(self, nbytes):
        self._lib = load()
        p = C.c_void_p()
        if self._lib.e1b200_host_alloc(C.byref(p), nbytes):
            raise E1B200Error("e1b200_host_alloc failed")
        self.ptr, self.nbytes = p.value, nbytes
        self.u8 = np.ctypeslib.as_array(C.cast(p, C.POINTER(C.c_uint8)), shape=(nbytes,))

    def view(self, dtype, count=None):
        dtype = np.dtype(dtype)
        n = self.nbytes // dtype.itemsize if count is None else count
        return self.u8[: n * dtype.itemsize].view(dtype)

    def free(self):
        if self.ptr:
            self.u8 = None
            self._lib.e1b200_host_free(self.ptr)
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class Synth:
    """One synthesiser context = one GPU, one producer thread (like the reference's generator)."""

    def __init__(self, fs_hz, samples_per_epoch, max_chan, device=0, dt_epoch=0.0, flags=0):
        self._lib = load()
        self.cfg = Config(float(fs_hz), int(samples_per_epoch), int(max_chan), int(device), int(flags), float(dt_epoch))
        h = C.c_void_p()
        rc = self._lib.e1b200_create(C.byref(self.cfg), C.byref(h))
        self._h = h
        if rc:
            msg = self._lib.e1b200_last_error(h).decode() if h.value else ""
            if h.value:
                self._lib.e1b200_destroy(h)
            self._h = None
            raise E1B200Error(f"e1b200_create failed ({rc}) {msg}: no usable CUDA device / bad config; no CPU fallback")
        self.n_samp, self.max_chan = int(samples_per_epoch), int(max_chan)

    def _check(self, rc, what):
        if rc:
            raise E1B200Error(f"{what}: {rc} {self._lib.e1b200_last_error(self._h).decode()}")

    def close(self):
        if self._h:
            self._lib.e1b200_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_channel(self, slot, prn, carr_phase0):
        self._check(self._lib.e1b200_set_channel(self._h, slot, prn, carr_phase0), "set_channel")

    def clear_channel(self, slot):
        self._check(self._lib.e1b200_clear_channel(self._h, slot), "clear_channel")

    def set_carrier_phase(self, slot, phase):
        self._check(self._lib.e1b200_set_carrier_phase(self._h, slot, phase), "set_carrier_phase")

    def get_carrier_phase(self, slot):
        v = C.c_double()
        self._check(self._lib.e1b200_get_carrier_phase(self._h, slot, C.byref(v)), "get_carrier_phase")
        return v.value

    def carrier_phases(self):
        out = np.zeros(self.max_chan)
        self._check(self._lib.e1b200_get_carrier_phases(self._h, self.max_chan, out.ctypes.data_as(C.POINTER(C.c_double))), "get_carrier_phases")
        return out

    def set_carrier_phases(self, phases):
        ph = np.ascontiguousarray(np.asarray(phases, dtype=np.float64)[: self.max_chan])
        self._check(self._lib.e1b200_set_carrier_phases(self._h, len(ph), ph.ctypes.data_as(C.POINTER(C.c_double))), "set_carrier_phases")

    def plan_phases_device(self, n_epochs, d_recs_ptr):
        """Carrier planner only, records on the device; asynchronous (sync() waits)."""
        self._check(self._lib.e1b200_plan_phases_device(self._h, n_epochs, d_recs_ptr), "plan_phases_device")

    def plan_phases(self, recs):
        """Advance the carrier phases over recs [n_epochs, max_chan] without synthesising (shard hand-off)."""
        recs = np.ascontiguousarray(recs, dtype=REC_DTYPE)
        assert recs.ndim == 2 and recs.shape[1] == self.max_chan, recs.shape
        self._check(self._lib.e1b200_plan_phases(self._h, recs.shape[0], recs.ctypes.data), "plan_phases")
        return self.carrier_phases()

    def synth_epochs_to(self, recs, out_ptr):
        """e1b200_synth_epochs with a raw destination pointer: host memory, or device memory of this GPU or of a peer
        (PeerBuffer): the finished slices are copied there by the copy engines behind the kernels."""
        recs = np.ascontiguousarray(recs, dtype=REC_DTYPE)
        assert recs.ndim == 2 and recs.shape[1] == self.max_chan, recs.shape
        self._check(self._lib.e1b200_synth_epochs(self._h, recs.shape[0], recs.ctypes.data, out_ptr), "synth_epochs")

    def _host_call(self, fn, recs, dtype, out):
        recs = np.ascontiguousarray(recs, dtype=dtype)
        assert recs.ndim == 2 and recs.shape[1] == self.max_chan, recs.shape
        n = recs.shape[0]
        if out is None:
            out = np.empty((n * self.n_samp, 2), np.int16)
        assert out.dtype == np.int16 and out.size >= n * self.n_samp * 2 and out.flags.c_contiguous
        self._check(fn(self._h, n, recs.ctypes.data, out.ctypes.data), fn.__name__)
        return out

    def synth_epochs(self, recs, out=None):
        """recs: REC_DTYPE [n_epochs, max_chan] (host).  Returns int16 [n_epochs*N, 2] (host)."""
        return self._host_call(self._lib.e1b200_synth_epochs, recs, REC_DTYPE, out)

    def synth_ranges(self, ranges, out=None):
        return self._host_call(self._lib.e1b200_synth_ranges, ranges, RANGE_DTYPE, out)

    def synth_epochs_device(self, n_epochs, d_recs_ptr, d_out_ptr):
        self._check(self._lib.e1b200_synth_epochs_device(self._h, n_epochs, d_recs_ptr, d_out_ptr), "synth_epochs_device")

    def synth_ranges_device(self, n_epochs, d_ranges_ptr, d_out_ptr):
        self._check(self._lib.e1b200_synth_ranges_device(self._h, n_epochs, d_ranges_ptr, d_out_ptr), "synth_ranges_device")

    def sync(self):
        self._check(self._lib.e1b200_sync(self._h), "sync")

    def stream(self):
        return self._lib.e1b200_stream(self._h)

    def timing(self):
        t = Timing()
        self._check(self._lib.e1b200_get_timing(self._h, C.byref(t)), "get_timing")
        return t

    def stats(self):
        s = Stats()
        self._check(self._lib.e1b200_get_stats(self._h, C.byref(s)), "get_stats")
        return s


class PeerBuffer:
    """The writer rank's stream buffer in HBM, shared with the other ranks of the node (e1b200_peer_*): rank 0
    PeerBuffer.alloc()s and ships .handle (64 bytes); the others PeerBuffer.open() it.  .ptr is a device pointer valid
    in this process; synthesis kernels store straight into it over NVLink."""

    def __init__(self, ptr, handle, owner, nbytes):
        self.ptr, self.handle, self.owner, self.nbytes = ptr, handle, owner, nbytes

    @classmethod
    def alloc(cls, device, nbytes):
        p, h = C.c_void_p(), C.create_string_buffer(64)
        rc = load().e1b200_peer_alloc(device, nbytes, C.byref(p), h)
        if rc:
            raise E1B200Error(f"e1b200_peer_alloc: {rc}")
        return cls(p.value, h.raw, True, nbytes)

    @classmethod
    def open(cls, device, handle, nbytes):
        p = C.c_void_p()
        rc = load().e1b200_peer_open(device, bytes(handle), C.byref(p))
        if rc:
            raise E1B200Error(f"e1b200_peer_open: {rc} (are the GPUs peers?)")
        return cls(p.value, bytes(handle), False, nbytes)

    def close(self):
        if self.ptr:
            (load().e1b200_peer_free if self.owner else load().e1b200_peer_close)(self.ptr)
            self.ptr = None
