"""ctypes binding of galileo-sdr-sim_b200/host/e1_scenario.h (libe1host.so): the CPU-side half of the
drop-in -- RINEX 3 navigation file -> e1_epoch_rec[n_blocks][max_chan] -- as plain C++."""
import ctypes as C
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
LIB_PATH = PKG / "lib" / "libe1host.so"

REC_DTYPE = np.dtype([
    ("prn", "<i4"), ("ibit0", "<i4"), ("flags", "<u4"), ("gain_q7", "<i4"),
    ("code_phase0", "<f8"), ("f_code", "<f8"), ("f_carr", "<f8"), ("carr_phase_init", "<f8"),
    ("page_cur", "u1", 64), ("page_next", "u1", 64),
])


RANGE_DTYPE = np.dtype([
    ("prn", "<i4"), ("flags", "<u4"), ("rho_prev", "<f8"), ("rho_cur", "<f8"), ("grx_sec", "<f8"), ("carr_phase_init", "<f8"),
    ("page_cur", "u1", 64), ("page_next", "u1", 64),
])


class Options(C.Structure):
    _fields_ = [("navfile", C.c_char * 512), ("llh", C.c_double * 3), ("have_start", C.c_int32),
                ("y", C.c_int32), ("m", C.c_int32), ("d", C.c_int32), ("hh", C.c_int32), ("mm", C.c_int32),
                ("sec", C.c_double), ("iduration", C.c_int32), ("iono_enable", C.c_int32), ("max_chan", C.c_int32),
                ("fs_hz", C.c_double), ("samples_per_epoch", C.c_int32), ("verbose", C.c_int32), ("elev_mask_deg", C.c_double)]


_lib = None


def load():
    global _lib
    if _lib is None:
        if not LIB_PATH.exists():
            raise RuntimeError(f"{LIB_PATH} is missing: python -c 'import build; build.build_host()' in galileo-sdr-sim_b200/")
        lib = C.CDLL(str(LIB_PATH))
        lib.e1h_default_options.argtypes = [C.POINTER(Options)]
        lib.e1h_open.argtypes = [C.POINTER(Options), C.c_char_p, C.c_int]
        lib.e1h_open.restype = C.c_void_p
        lib.e1h_close.argtypes = [C.c_void_p]
        lib.e1h_total_epochs.argtypes = [C.c_void_p]
        lib.e1h_next.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        lib.e1h_next_ex.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
        lib.e1h_set_location.argtypes = [C.c_void_p, C.c_double, C.c_double, C.c_double]
        lib.e1h_set_motion.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
        lib.e1h_ecef_to_llh_deg.argtypes = [C.c_void_p, C.c_void_p]
        lib.e1h_page_symbols.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_int, C.c_void_p]
        lib.e1h_crc24q_bits.argtypes = [C.c_void_p, C.c_int]
        lib.e1h_crc24q_bits.restype = C.c_uint
        lib.e1h_encode_page.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        _lib = lib
    return _lib


class Scenario:
    """navfile + receiver position (+ start time) -> records, exactly as the reference's galileo_task()
    derives its channel state for every 0.1 s block."""

    def __init__(self, navfile, llh=None, start=None, duration_s=300.0, iono=True, max_chan=16, verbose=False, fs_hz=None,
                 elev_mask_deg=0.0):
        lib = load()
        o = Options()
        lib.e1h_default_options(C.byref(o))
        o.navfile = str(navfile).encode()
        if llh is not None:
            o.llh[:] = list(map(float, llh))
        if start is not None:                  # (y, m, d, hh, mm, sec)
            o.have_start = 1
            o.y, o.m, o.d, o.hh, o.mm = map(int, start[:5])
            o.sec = float(start[5])
        o.iduration = int(duration_s * 10.0 + 0.5)
        o.iono_enable = 1 if iono else 0
        o.max_chan = max_chan
        o.verbose = 1 if verbose else 0
        if fs_hz is not None:                  # a build of the reference with another SAMP_RATE (oracle/ref_patches/fs25.diff)
            o.fs_hz = float(np.float32(fs_hz))
            o.samples_per_epoch = int(np.float32(fs_hz) / 10)      # NUM_IQ_SAMPLES = TX_SAMPLERATE / 10
        o.elev_mask_deg = float(elev_mask_deg)  # 0 = the reference's 10 degrees
        err = C.create_string_buffer(256)
        self._h = lib.e1h_open(C.byref(o), err, 256)
        if not self._h:
            raise RuntimeError(err.value.decode() or "e1h_open failed")
        self.max_chan, self.n_epochs = max_chan, lib.e1h_total_epochs(self._h)
        self.fs_hz, self.samples_per_epoch = o.fs_hz, o.samples_per_epoch

    def next(self, n):
        recs = np.zeros((n, self.max_chan), REC_DTYPE)
        grx = np.zeros(n)
        got = load().e1h_next(self._h, n, recs.ctypes.data, grx.ctypes.data)
        return recs[:got], grx[:got]

    def next_ranges(self, n, with_recs=False):
        """The next n blocks as pseudoranges (e1_range_rec) for the device-side restate; with_recs also
        returns the host-restated e1_epoch_rec of the same blocks."""
        rng = np.zeros((n, self.max_chan), RANGE_DTYPE)
        recs = np.zeros((n, self.max_chan), REC_DTYPE) if with_recs else None
        grx = np.zeros(n)
        got = load().e1h_next_ex(self._h, n, recs.ctypes.data if with_recs else None, rng.ctypes.data, grx.ctypes.data)
        return (rng[:got], recs[:got], grx[:got]) if with_recs else (rng[:got], grx[:got])

    def set_location(self, lat_deg, lon_deg, height_m):
        """What the reference's UDP location thread writes into llhr (include/socket.h:165-178)."""
        load().e1h_set_location(self._h, float(lat_deg), float(lon_deg), float(height_m))

    def set_motion(self, llh_deg):
        """Table of (lat, lon [deg], height [m]) indexed by block number (entry 0 unused)."""
        a = np.ascontiguousarray(llh_deg, np.float64).reshape(-1, 3)
        load().e1h_set_motion(self._h, a.shape[0], a.ctypes.data)

    def all(self):
        return self.next(self.n_epochs)

    def close(self):
        if self._h:
            load().e1h_close(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
