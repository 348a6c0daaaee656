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
    ("prn", "<i4"), ("ibit0", "<i4"), ("flags", "<u4"), ("gain_q7", "<i4"),
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
    ("iumd", "<i4"), ("slot", "<i4"), ("prn", "<i4"), ("ibit", "<i4"), ("ipage", "<i4"), ("gain", "<i4"),
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
        lib.e1o_synth_epochs_float.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p, dp, C.c_void_p, C.c_double, C.c_double, C.c_int, C.c_int]
        lib.e1o_carrier_phases.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p, dp, dp, C.c_int]
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


ALPHA_CBOC, BETA_CBOC = (10.0 / 11.0) ** 0.5, (1.0 / 11.0) ** 0.5     # Galileo OS SIS ICD: CBOC(6,1,1/11)


def oracle_synth_float(fs_hz, n_samp, recs, carr_phase=None, cboc=True, use_gain=False, threads=1):
    """SURVEY 8 f4 oracle (parity unpinned): CBOC sub-carrier and/or per-record gain, double accumulate, nearest int16."""
    recs = np.ascontiguousarray(recs)
    n_epochs, max_chan = recs.shape
    ph = np.zeros(max_chan) if carr_phase is None else np.array(carr_phase, dtype=np.float64)
    out = np.empty((n_epochs * n_samp, 2), np.int16)
    a, b = (ALPHA_CBOC, BETA_CBOC) if cboc else (1.0, 0.0)
    oracle().e1o_synth_epochs_float(fs_hz, n_samp, max_chan, n_epochs, recs.ctypes.data, ph.ctypes.data_as(C.POINTER(C.c_double)),
                                    out.ctypes.data, a, b, int(use_gain), threads)
    return out, ph


def oracle_carrier_phases(fs_hz, n_samp, recs, carr_phase=None, threads=1):
    """The literal carrier recurrence alone: (phase at the top of every block [n_epochs, max_chan], final phases)."""
    recs = np.ascontiguousarray(recs)
    n_epochs, max_chan = recs.shape
    ph = np.zeros(max_chan) if carr_phase is None else np.array(carr_phase, dtype=np.float64)
    out = np.zeros((n_epochs, max_chan))
    dp = C.POINTER(C.c_double)
    oracle().e1o_carrier_phases(fs_hz, n_samp, max_chan, n_epochs, recs.ctypes.data, ph.ctypes.data_as(dp), out.ctypes.data_as(dp), threads)
    return out, ph


def oracle_restate(rho_prev, rho_cur, dt, grx_sec):
    f = [C.c_double() for _ in range(3)]
    ib, ip = C.c_int(), C.c_int()
    oracle().e1o_restate(rho_prev, rho_cur, dt, grx_sec, f[0], f[1], f[2], ib, ip)
    return f[0].value, f[1].value, f[2].value, ib.value, ip.value


# ----------------------------------------------------------------------------- hostsim loader
_hostsim = None
HOSTSIM_DIR = Path(__file__).resolve().parent / "hostsim"


def hostsim():
    """tests/hostsim/libe1hostsim.so: the product's exact-arithmetic header (csrc/e1_core.h) compiled
    for the host and driven like the kernels drive it.  Test scaffolding for GPU-less boxes."""
    global _hostsim
    if _hostsim is None:
        so, src = HOSTSIM_DIR / "libe1hostsim.so", HOSTSIM_DIR / "hostsim.cpp"
        deps = [src, PKG / "csrc" / "e1_core.h", ROOT / "include" / "e1b200.h"]
        if not so.exists() or any(so.stat().st_mtime < d.stat().st_mtime for d in deps):
            subprocess.check_call(["g++", "-O2", "-ffp-contract=off", "-fPIC", "-shared", "-Wall", "-Wextra",
                                   "-o", str(so), str(src)])
        hs = C.CDLL(str(so))
        hs.hs_carr_advance.restype = C.c_double
        hs.hs_carr_advance.argtypes = [C.c_double, C.c_double, C.c_long, C.c_long]
        hs.hs_carr_literal.restype = C.c_double
        hs.hs_carr_literal.argtypes = [C.c_double, C.c_double, C.c_long]
        hs.hs_code_advance.restype = C.c_double
        hs.hs_code_advance.argtypes = [C.c_double, C.c_double, C.c_long, C.POINTER(C.c_long)]
        hs.hs_code_literal.restype = C.c_double
        hs.hs_code_literal.argtypes = [C.c_double, C.c_double, C.c_long, C.POINTER(C.c_long)]
        hs.hs_to_fixed.restype = C.c_ulonglong
        hs.hs_to_fixed.argtypes = [C.c_double, C.c_int]
        hs.hs_build_codes.argtypes = [C.c_void_p]
        hs.hs_build_lut.argtypes = [C.POINTER(C.c_int), C.POINTER(C.c_int), C.c_void_p]
        hs.hs_synth_epochs.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p,
                                       C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        hs.hs_synth_epochs_p.argtypes = hs.hs_synth_epochs.argtypes + [C.c_int]
        hs.hs_lut_oob.restype = C.c_ulonglong
        hs.hs_chain_scan_compare.restype = C.c_long
        hs.hs_chain_scan_compare.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        hs.hs_first_hit.restype = C.c_longlong
        hs.hs_first_hit.argtypes = [C.c_longlong] * 5
        for f in (hs.hs_clean_tiles, hs.hs_checked_tiles, hs.hs_clean_violations, hs.hs_cw_pairs, hs.hs_ev_pairs, hs.hs_ev_rest):
            f.restype = C.c_ulonglong
        hs.hs_max_code_ab_units.restype = C.c_double
        hs.hs_plan_compare.restype = C.c_long
        hs.hs_plan_compare.argtypes = [C.c_double, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        _hostsim = hs
    return _hostsim


def product_lut():
    """int32[642][32] carrier table in the product's layout (E1C_LUT_IDX x E1C_LUT_REP), built from the ORACLE's tables."""
    c, s = (C.c_int * 512)(), (C.c_int * 512)()
    oracle().e1o_carrier_lut(c, s)
    lut = np.zeros(hostsim().hs_lut_entries(), np.int32)
    hostsim().hs_build_lut(c, s, lut.ctypes.data)
    return lut


CFG_CBOC, CFG_GAIN = 2, 4     # E1B200_CFG_CBOC / E1B200_CFG_GAIN (include/e1b200.h)


def hostsim_synth(fs_hz, n_samp, recs, carr_phase=None, groups=4, amb_scale=1, planner=2, cfg_flags=0):
    """Same contract as oracle_synth, through the product's core header on the host.
    planner 1 = serial exact walk per channel, 2 = the parallel planner's passes (default, what the
    kernels run).  Returns (int16 [n_epochs*n_samp, 2], final phases, stats[5]): stats = fallback
    samples, planner errors, slow-path threads, epochs walked serially by the chain, HAT epochs."""
    recs = np.ascontiguousarray(recs)
    n_epochs, max_chan = recs.shape
    ph = np.zeros(max_chan) if carr_phase is None else np.array(carr_phase, dtype=np.float64)
    out = np.zeros((n_epochs * n_samp, 2), np.int16)
    st = np.zeros(5, np.uint64)
    lut = product_lut()
    hostsim().hs_set_cfg_flags(cfg_flags)
    try:
        rc = hostsim().hs_synth_epochs_p(fs_hz, n_samp, max_chan, n_epochs, recs.ctypes.data, ph.ctypes.data, out.ctypes.data,
                                         groups, amb_scale, lut.ctypes.data, st.ctypes.data, planner)
    finally:
        hostsim().hs_set_cfg_flags(0)
    assert rc == 0, f"hostsim planner errors: {st}"
    assert hostsim().hs_lut_oob() == 0, "a fast-form lookup left the carrier table"
    assert hostsim().hs_clean_violations() == 0, "a run of a tile marked E1_PAR_CLEAN was flagged by the tracking sample loop"
    return out, ph, st


def hostsim_plan_compare(fs_hz, n_samp, recs, carr_phase=None, groups=4):
    """Serial vs parallel carrier planner on the host: (mismatching checkpoints, [serial epochs, HAT epochs, active epochs])."""
    recs = np.ascontiguousarray(recs)
    n_epochs, max_chan = recs.shape
    ph = np.zeros(max_chan) if carr_phase is None else np.array(carr_phase, dtype=np.float64)
    st = np.zeros(3, np.uint64)
    bad = hostsim().hs_plan_compare(fs_hz, n_samp, max_chan, n_epochs, recs.ctypes.data, ph.ctypes.data, groups, st.ctypes.data)
    return int(bad), [int(x) for x in st]


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
        r["gain_q7"] = t["gain"]           # the reference's gain[i] of this block (src/galileo-sdr.cpp:477; it never uses it)
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


def synthetic_ranges(n_epochs, n_chan, seed=0, max_chan=None, dt=REF_DT, grx0=43200.0):
    """Pseudorange-level synthetic inputs for the device-side restate (BASELINE config 4): per channel a
    range of ~2.3e7 m with a range rate of +-800 m/s and a small acceleration; pages as in synthetic_recs.
    Returns (RANGE_DTYPE [n_epochs, max_chan], REC_DTYPE [n_epochs, max_chan] restated by the ORACLE)."""
    rng = np.random.default_rng(seed)
    max_chan = max_chan or n_chan
    rr = np.zeros((n_epochs, max_chan), RANGE_DTYPE)
    recs = np.zeros((n_epochs, max_chan), REC_DTYPE)
    sync = np.array([0, 1, 0, 1, 1, 0, 0, 0, 0, 0], np.uint8)
    rho0 = rng.uniform(2.2e7, 2.6e7, n_chan)
    v = rng.uniform(-800, 800, n_chan)
    a = rng.uniform(-0.5, 0.5, n_chan)
    ph0 = rng.uniform(0, 1, n_chan)

    def page():
        return pack_page(np.concatenate([sync, rng.integers(0, 2, 240), sync, rng.integers(0, 2, 240)]).astype(np.uint8))

    cur = [page() for _ in range(n_chan)]
    nxt = [page() for _ in range(n_chan)]
    for e in range(n_epochs):
        t0, t1 = e * dt, (e + 1) * dt
        for c in range(n_chan):
            r = rr[e, c]
            r["prn"] = (c % 50) + 1
            r["rho_prev"] = rho0[c] + v[c] * t0 + 0.5 * a[c] * t0 * t0
            r["rho_cur"] = rho0[c] + v[c] * t1 + 0.5 * a[c] * t1 * t1
            r["grx_sec"] = grx0 + t1
            fc, fcode, cp, ib, _ = oracle_restate(float(r["rho_prev"]), float(r["rho_cur"]), dt, float(r["grx_sec"]))
            o = recs[e, c]
            o["prn"], o["ibit0"], o["code_phase0"], o["f_code"], o["f_carr"] = r["prn"], ib, cp, fcode, fc
            if e == 0:
                r["flags"] = o["flags"] = E1_REC_SET_PHASE
                r["carr_phase_init"] = o["carr_phase_init"] = ph0[c]
            r["page_cur"], r["page_next"] = cur[c], nxt[c]
            o["page_cur"], o["page_next"] = cur[c], nxt[c]
            if ib + 26 >= 500:
                cur[c], nxt[c] = nxt[c], page()
    return rr, recs


def synthetic_recs_fast(n_epochs, n_chan, fs_hz, seed=0, max_chan=None, f_max=4000.0):
    """Vectorised generator with the SURVEY.md section 8(d) distribution, for bench-size inputs
    (thousands of epochs).  Not sample-identical to synthetic_recs(); same statistics."""
    rng = np.random.default_rng(seed)
    max_chan = max_chan or n_chan
    recs = np.zeros((n_epochs, max_chan), REC_DTYPE)
    e = np.arange(n_epochs)[:, None]
    f0 = rng.uniform(-f_max, f_max, n_chan)[None, :]
    # smooth Doppler: a constant rate of up to +-0.1 Hz per 0.1 s block (MEO passes stay below ~1 Hz/s)
    f = f0 + rng.uniform(-0.1, 0.1, n_chan)[None, :] * e
    ib0 = rng.integers(0, 500, n_chan)[None, :]
    ib = (ib0 + 25 * e) % 500
    v = recs[:, :n_chan]
    v["prn"] = (np.arange(n_chan) % 50 + 1)[None, :]
    v["ibit0"] = ib
    v["code_phase0"] = (rng.uniform(0, 4092, n_chan)[None, :] + 0.024 * e) % 4092
    v["f_carr"] = f
    v["f_code"] = 1.023e6 + f * 0.0006493506493506494
    v["flags"][0, :] = E1_REC_SET_PHASE
    v["carr_phase_init"][0, :] = rng.uniform(0, 1, n_chan)
    sync = np.array([0, 1, 0, 1, 1, 0, 0, 0, 0, 0], np.uint8)
    # a page lasts 20 epochs; page k of a channel is in force while floor((ib0 + 25 e) / 500) == k
    pidx = (ib0 + 25 * e) // 500
    n_pages = int(pidx.max()) + 2
    sym = rng.integers(0, 2, (n_chan, n_pages, 500), dtype=np.uint8)
    sym[:, :, 0:10] = sync
    sym[:, :, 250:260] = sync
    packed = np.packbits(np.concatenate([sym, np.zeros((n_chan, n_pages, 12), np.uint8)], axis=2), axis=2, bitorder="little")
    c = np.arange(n_chan)[None, :]
    v["page_cur"] = packed[c, pidx]
    v["page_next"] = packed[c, pidx + 1]
    return recs


# ----------------------------------------------------------------------------- shard-test engine
class OracleEngine:
    """The engine interface galileo-sdr-sim_b200/shard.py drives (set_carrier_phases / plan_phases /
    synth_epochs / carrier_phases / max_chan), backed by the CPU oracle.  TEST ONLY: lets the GPU-less
    suite cover the split / phase hand-off / file-offset logic of the sharded path."""

    def __init__(self, fs_hz, n_samp, max_chan, threads=1):
        self.fs, self.n_samp, self.max_chan, self.threads = fs_hz, n_samp, max_chan, threads
        self.ph = np.zeros(max_chan)

    def set_carrier_phases(self, phases):
        self.ph = np.array(phases, dtype=np.float64)[: self.max_chan].copy()

    def carrier_phases(self):
        return self.ph.copy()

    def plan_phases(self, recs):
        _, self.ph = oracle_synth(self.fs, self.n_samp, recs, self.ph, threads=self.threads)
        return self.ph.copy()

    def synth_epochs(self, recs, out=None):
        res, self.ph = oracle_synth(self.fs, self.n_samp, recs, self.ph, threads=self.threads)
        if out is not None:
            out.reshape(-1)[: res.size] = res.reshape(-1)
            return out
        return res


def hostsim_chain_scan_compare(fs_hz, n_samp, recs, carr_phase=None, round_spans=512):
    """The chain kernel's rounds (optimistic scan + serial step) transcribed for the host against the serial
    chain: (differing translations, [spans accepted by the scan, spans through the serial step])."""
    recs = np.ascontiguousarray(recs)
    n_epochs, max_chan = recs.shape
    ph = np.zeros(max_chan) if carr_phase is None else np.array(carr_phase, dtype=np.float64)
    st = np.zeros(2, np.uint64)
    bad = hostsim().hs_chain_scan_compare(fs_hz, n_samp, max_chan, n_epochs, recs.ctypes.data, ph.ctypes.data, round_spans, st.ctypes.data)
    return int(bad), [int(v) for v in st]
