"""Shared helpers for the tests, bench.py and __graft_entry__: ctypes views of the C-ABI
structs, a loader for the CPU oracle (test infrastructure) and synthetic-input generators
(SURVEY.md section 8d).  Nothing here is on the product path."""
import ctypes as C
import os
import subprocess
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent.parent
PKG = ROOT / "galileo-sdr-sim_b200"
ORACLE_DIR = ROOT / "oracle"

E1_REC_SET_PHASE = 1
PAGE_BYTES = 64

REC_DTYPE = np.dtype([
    ("prn", "<i4"), ("ibit0", "<i4"), ("flags", "<u4"), ("reserved", "<u4"),
    ("code_phase0", "<f8"), ("f_code", "<f8"), ("f_carr", "<f8"), ("carr_phase_init", "<f8"),
    ("page_cur", "u1", PAGE_BYTES), ("page_next", "u1", PAGE_BYTES),
])
assert REC_DTYPE.itemsize == 176

RANGE_DTYPE = np.dtype([
    ("prn", "<i4"), ("flags", "<u4"),
    ("rho_prev", "<f8"), ("rho_cur", "<f8"), ("grx_sec", "<f8"), ("carr_phase_init", "<f8"),
    ("page_cur", "u1", PAGE_BYTES), ("page_next", "u1", PAGE_BYTES),
])
assert RANGE_DTYPE.itemsize == 168

TRACE_DTYPE = np.dtype([
    ("iumd", "<i4"), ("slot", "<i4"), ("prn", "<i4"), ("ibit", "<i4"), ("ipage", "<i4"), ("pad", "<i4"),
    ("code_phase", "<f8"), ("f_code", "<f8"), ("f_carr", "<f8"), ("carr_phase", "<f8"),
    ("grx", "<f8"), ("rho", "<f8"), ("page", "u1", 504),
])

REF_DT = 0.10000002314200000  # src/galileo-sdr.cpp:347


def fs_as_reference(fs):
    """delt = 1/(double)(float)SAMP_RATE (include/constants.h:96, src/galileo-sdr.cpp:162)."""
    return float(np.float32(fs))


def pack_page(sym):
    """500 symbols (0/1) -> 64 bytes, bit k at byte k>>3, bit k&7."""
    sym = np.asarray(sym, dtype=np.uint8)[:500]
    return np.packbits(np.concatenate([sym, np.zeros(512 - 500, np.uint8)]), bitorder="little")


# ----------------------------------------------------------------------------- oracle loader
_oracle = None


def build_oracle(force=False):
    so = ORACLE_DIR / "libe1oracle.so"
    src = ORACLE_DIR / "e1_oracle.c"
    if force or not so.exists() or so.stat().st_mtime < src.stat().st_mtime:
        subprocess.check_call(["make", "-s", "-C", str(ORACLE_DIR), "libe1oracle.so"])
    return so


def oracle():
    """ctypes handle on oracle/libe1oracle.so (CPU restatement; checker only)."""
    global _oracle
    if _oracle is None:
        lib = C.CDLL(str(build_oracle()))
        dp = C.POINTER(C.c_double)
        lib.e1o_carrier_lut.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.e1o_halfchip_table.argtypes = [C.c_int, C.c_int, C.POINTER(C.c_short)]
        lib.e1o_restate.argtypes = [C.c_double] * 4 + [dp, dp, dp, C.POINTER(C.c_int), C.POINTER(C.c_int)]
        lib.e1o_synth_epochs.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p, dp, C.c_void_p]
        lib.e1o_synth_epochs_mt.argtypes = lib.e1o_synth_epochs.argtypes + [C.c_int]
        _oracle = lib
    return _oracle


def oracle_synth(fs_hz, n_samp, recs, carr_phase=None, threads=1):
    """recs: REC_DTYPE array [n_epochs, max_chan].  Returns (int16 [n_epochs*n_samp, 2], final phases)."""
    recs = np.ascontiguousarray(recs)
    n_epochs, max_chan = recs.shape
    ph = np.zeros(max_chan) if carr_phase is None else np.array(carr_phase, dtype=np.float64)
    out = np.empty((n_epochs * n_samp, 2), np.int16)
    lib = oracle()
    args = [fs_hz, n_samp, max_chan, n_epochs, recs.ctypes.data, ph.ctypes.data_as(C.POINTER(C.c_double)), out.ctypes.data]
    if threads > 1:
        lib.e1o_synth_epochs_mt(*args, threads)
    else:
        lib.e1o_synth_epochs(*args)
    return out, ph


def oracle_restate(rho_prev, rho_cur, dt, grx_sec):
    f = [C.c_double() for _ in range(3)]
    ib, ip = C.c_int(), C.c_int()
    oracle().e1o_restate(rho_prev, rho_cur, dt, grx_sec, f[0], f[1], f[2], ib, ip)
    return f[0].value, f[1].value, f[2].value, ib.value, ip.value


# ------------------------------------------------------------------- reference trace -> recs
def trace_to_recs(trace, max_chan):
    """Channel-state trace of the reference (oracle/ref_hooks) -> e1_epoch_rec[n_epochs][max_chan].

    page_next of (epoch e, slot s) is the page the reference holds at the start of epoch e+1: the
    in-loop generateINavMsg (src/galileo-sdr.cpp:497-506) is the only writer in between.  The first
    record of a slot carries E1_REC_SET_PHASE with the traced carrier phase (src/channel.cpp:98-99);
    after that the phase is integrated by the synthesiser.  Returns (recs, traced_phase[n_epochs][max_chan])."""
    e0, e1 = int(trace["iumd"].min()), int(trace["iumd"].max())
    n_epochs = e1 - e0 + 1
    recs = np.zeros((n_epochs, max_chan), REC_DTYPE)
    phase = np.full((n_epochs, max_chan), np.nan)
    for t in trace:
        e, s = int(t["iumd"]) - e0, int(t["slot"])
        r = recs[e, s]
        r["prn"], r["ibit0"] = t["prn"], t["ibit"]
        r["code_phase0"], r["f_code"], r["f_carr"] = t["code_phase"], t["f_code"], t["f_carr"]
        r["page_cur"] = pack_page(t["page"][:500])
        phase[e, s] = t["carr_phase"]
    for e in range(n_epochs):
        for s in range(max_chan):
            r = recs[e, s]
            if r["prn"] == 0:
                continue
            nxt = recs[e + 1, s] if e + 1 < n_epochs else None
            r["page_next"] = nxt["page_cur"] if (nxt is not None and nxt["prn"] == r["prn"]) else r["page_cur"]
            if e == 0 or recs[e - 1, s]["prn"] != r["prn"]:
                r["flags"] = E1_REC_SET_PHASE
                r["carr_phase_init"] = phase[e, s]
    return recs, phase


# ------------------------------------------------------------------------- synthetic inputs
def synthetic_recs(n_epochs, n_chan, fs_hz, seed=0, max_chan=None, f_max=4000.0):
    """Trace-level synthetic inputs, SURVEY.md section 8(d): PRN 1..C, f_carr ~ U(-f_max,f_max) with a
    per-epoch drift, code_phase0 ~ U(0,4092), ibit0 ~ U{0..499}, pages = sync + Bernoulli(1/2)."""
    rng = np.random.default_rng(seed)
    max_chan = max_chan or n_chan
    recs = np.zeros((n_epochs, max_chan), REC_DTYPE)
    sync = np.array([0, 1, 0, 1, 1, 0, 0, 0, 0, 0], np.uint8)
    f0 = rng.uniform(-f_max, f_max, n_chan)
    cp0 = rng.uniform(0, 4092, n_chan)
    ib0 = rng.integers(0, 500, n_chan)
    ph0 = rng.uniform(0, 1, n_chan)

    def page():
        return pack_page(np.concatenate([sync, rng.integers(0, 2, 240), sync, rng.integers(0, 2, 240)]).astype(np.uint8))

    cur = [page() for _ in range(n_chan)]
    nxt = [page() for _ in range(n_chan)]
    for e in range(n_epochs):
        for c in range(n_chan):
            r = recs[e, c]
            f = f0[c] + rng.uniform(-0.1, 0.1) * (e + 1)
            r["prn"] = (c % 50) + 1
            # 25 symbols per epoch, plus a small forward code-phase snap like the reference's dt constant
            ib = int(ib0[c] + 25 * e) % 500
            r["ibit0"] = ib
            r["code_phase0"] = (cp0[c] + 0.024 * e) % 4092
            r["f_carr"] = f
            r["f_code"] = 1.023e6 + f * 0.0006493506493506494
            r["page_cur"], r["page_next"] = cur[c], nxt[c]
            if e == 0:
                r["flags"] = E1_REC_SET_PHASE
                r["carr_phase_init"] = ph0[c]
            if ib + 25 >= 500:  # the page turns inside (or at the end of) this epoch
                cur[c], nxt[c] = nxt[c], page()
    return recs
