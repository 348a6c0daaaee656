"""CBOC(6,1,1/11) sub-carrier and per-satellite gain -- the FLOAT path (E1B200_CFG_CBOC /
E1B200_CFG_GAIN, include/e1b200.h).  The reference transmits BOC(1,1) at unit gain, so this mode has no reference
bytes (PARITY UNPINNED, oracle/e1_oracle.c: e1o_synth_epochs_float, written from the Galileo OS SIS ICD); what can be
pinned is pinned: with the reference's weights (1, 0) and unit gains the float path -- same kernel, same sub-chip
indexing, FP32 sums, float -> int16 store -- writes the REFERENCE's bytes (md5 419622c8...).  Against the oracle in
CBOC mode the criterion is the north star's +-1 LSB (FP32 against double accumulation decides a rounding now and
then), stated below.  CPU tests drive the product's core header through tests/hostsim, -m gpu tests the CUDA path."""
import hashlib
import sys
from pathlib import Path

import numpy as np
import pytest

import e1util as U

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "galileo-sdr-sim_b200"))
GOLD = Path(__file__).parent / "golden"
FS26 = U.fs_as_reference(2.6e6)
FS25 = U.fs_as_reference(25e6)
TOL_LSB = 1                 # north star: "within +-1 LSB (int16)"; integer modes stay at 0
MAX_DIFF_FRACTION = 1e-3    # ... and only where a sum lands within FP32 rounding of a half: a few 1e-5 of the values


def with_gains(recs, seed=0):
    """gain_q7 like the reference's gain[i] (src/galileo-sdr.cpp:477): path loss x antenna pattern x 128, 3..91 on the goldens"""
    rng = np.random.default_rng(seed)
    recs = recs.copy()
    g = rng.integers(3, 260, recs.shape[1])
    recs["gain_q7"] = np.where(recs["prn"] > 0, g[None, :], 0)
    return recs


def check_tolerance(got, ref):
    d = got.astype(np.int32) - ref.astype(np.int32)
    assert np.abs(d).max() <= TOL_LSB, f"max |diff| {np.abs(d).max()} LSB"
    assert np.count_nonzero(d) <= MAX_DIFF_FRACTION * d.size, f"{np.count_nonzero(d)} of {d.size} values differ"


# ------------------------------------------------------------------------------------------ CPU (hostsim)
def test_float_path_with_reference_weights_writes_the_reference_bytes():
    """E1B200_CFG_GAIN with every gain_q7 = 0 (unit): BOC(1,1) weights through the float path, on the reference's own
    trace of BASELINE configs[0] -- the first 12 blocks' SHA-256 equal the reference's."""
    recs = np.load(GOLD / "cfg1_recs.npz")["recs"][:12].copy()
    recs["gain_q7"] = 0
    sha = (GOLD / "cfg1_sha256.txt").read_text().splitlines()[1:]
    out, _, _ = U.hostsim_synth(FS26, 260000, recs, cfg_flags=U.CFG_GAIN)
    out = out.reshape(12, 260000, 2)
    for e in range(12):
        assert hashlib.sha256(out[e].tobytes()).hexdigest() == sha[e], e


def test_float_oracle_with_reference_weights_is_the_integer_oracle():
    recs = np.load(GOLD / "cfg1_recs.npz")["recs"]
    o, ph = U.oracle_synth_float(FS26, 260000, recs, cboc=False, threads=8)
    assert hashlib.md5(o.tobytes()).hexdigest() == "419622c87f06f4048858bce54df72d29"
    _, ph2 = U.oracle_synth(FS26, 260000, recs, threads=8)
    assert np.array_equal(ph, ph2)


@pytest.mark.parametrize("fs,n,nch,n_ep,flags", [
    (FS26, 260000, 8, 2, U.CFG_CBOC),
    (FS26, 100000, 12, 2, U.CFG_CBOC | U.CFG_GAIN),
    (FS25, 300000, 6, 2, U.CFG_CBOC | U.CFG_GAIN),
    (FS26, 100000, 12, 2, U.CFG_GAIN),
])
def test_hostsim_float_path_against_the_float_oracle(fs, n, nch, n_ep, flags):
    recs = with_gains(U.synthetic_recs(n_ep, nch, fs, seed=nch, max_chan=nch + 2))
    got, ph, st = U.hostsim_synth(fs, n, recs, cfg_flags=flags)
    ref, ph_ref = U.oracle_synth_float(fs, n, recs, cboc=bool(flags & U.CFG_CBOC), use_gain=bool(flags & U.CFG_GAIN), threads=8)
    assert np.array_equal(ph, ph_ref)
    if flags == U.CFG_GAIN:          # integer gains over 128 on integer terms: sums of multiples of 1/64 -- exact in FP32, ties to even on both sides
        assert np.array_equal(got, ref)
    else:
        check_tolerance(got, ref)
    if flags & U.CFG_CBOC:           # it IS another signal: the BOC(6,1) component carries 1/11 of the power
        boc, _ = U.oracle_synth_float(fs, n, recs, cboc=False, use_gain=bool(flags & U.CFG_GAIN), threads=8)
        c = np.corrcoef(boc[:, 0].astype(np.float64), ref[:, 0].astype(np.float64))[0, 1]
        assert (0.94 < c < 0.965) if fs == FS25 else (0.90 < c < 0.99), c     # alpha = 0.9535


def test_sub_chip_division_and_sign_algebra():
    """(s * 43691) >> 19 == s // 12 over every sub-chip index, and the kernel's two-case form of
    m = eB - eC = alpha a (Bd - Cs) + beta b (Bd + Cs) against the ICD expression, all 16 sign combinations."""
    s = np.arange(0, 12 * 4092 + 24, dtype=np.uint64)
    assert np.array_equal((s * 43691) >> 19, s // 12)
    a_, b_ = U.ALPHA_CBOC, U.BETA_CBOC
    for bd in (-1, 1):
        for cs in (-1, 1):
            for a in (-1, 1):
                for b in (-1, 1):
                    icd = bd * (a_ * a + b_ * b) - cs * (a_ * a - b_ * b)
                    mine = 2 * a_ * a * bd if bd != cs else 2 * b_ * b * bd
                    assert abs(icd - mine) < 1e-15


# ------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
def test_cuda_float_path_with_reference_weights_writes_the_reference_file():
    """The float kernel with BOC(1,1) weights and unit gains on the reference's trace of configs[0]: md5 419622c8...,
    and on 24 satellites at 25 MS/s (the patched reference build): md5 542bc088..."""
    import e1b200 as E
    for name, fs, n, md5 in (("cfg1", FS26, 260000, "419622c87f06f4048858bce54df72d29"), ("fs25ch36", FS25, 2500000, "542bc0887b3f81897f2af436ac42ab89")):
        recs = np.load(GOLD / f"{name}_recs.npz")["recs"].copy()
        recs["gain_q7"] = 0
        s = E.Synth(fs, n, recs.shape[1], flags=U.CFG_GAIN)
        out = s.synth_epochs(recs)
        s.close()
        assert hashlib.md5(out.tobytes()).hexdigest() == md5, name


@pytest.mark.gpu
@pytest.mark.parametrize("fs,n,nch,n_ep,flags", [
    (FS26, 260000, 36, 3, U.CFG_CBOC),
    (FS26, 260000, 36, 3, U.CFG_CBOC | U.CFG_GAIN),
    (FS25, 2500000, 24, 1, U.CFG_CBOC | U.CFG_GAIN),
    (FS25, 250001, 5, 2, U.CFG_CBOC),                    # ragged block: scalar stores
])
def test_cuda_cboc_and_gain_within_one_lsb_of_the_oracle(fs, n, nch, n_ep, flags):
    import e1b200 as E
    recs = with_gains(U.synthetic_recs(n_ep, nch, fs, seed=nch))
    ref, ph_ref = U.oracle_synth_float(fs, n, recs, cboc=True, use_gain=bool(flags & U.CFG_GAIN), threads=16)
    s = E.Synth(fs, n, nch, flags=flags)
    got = s.synth_epochs(recs)
    assert np.array_equal(s.carrier_phases(), ph_ref)
    s.close()
    check_tolerance(got, ref)
    hs, _, _ = U.hostsim_synth(fs, n, recs[:1], cfg_flags=flags) if n <= 300000 else (None, None, None)
    if hs is not None:               # same FP32 operations in the same order on host and device: identical, not just close
        assert np.array_equal(got[:n], hs)


@pytest.mark.gpu
def test_cuda_gain_on_the_reference_trace_with_the_reference_gains():
    """E1B200_CFG_GAIN with the gain[i] the reference computed for these very blocks (traced; it never applies them):
    equal to the float oracle exactly (integer terms x gain/128), and every satellite's share of the stream scales
    with its gain (the loudest is 91/128, the weakest 46/128 on configs[0])."""
    import e1b200 as E
    recs = np.load(GOLD / "cfg1_recs.npz")["recs"][:10]
    assert recs["gain_q7"][0][:8].tolist() == [67, 67, 46, 91, 62, 61, 61, 46]
    ref, _ = U.oracle_synth_float(FS26, 260000, recs, cboc=False, use_gain=True, threads=16)
    s = E.Synth(FS26, 260000, 16, flags=U.CFG_GAIN)
    got = s.synth_epochs(recs)
    s.close()
    assert np.array_equal(got, ref)
    unit, _ = U.oracle_synth(FS26, 260000, recs, threads=16)
    assert 0.3 < np.abs(got).mean() / np.abs(unit).mean() < 0.7


@pytest.mark.gpu
def test_receiver_acquires_the_cboc_stream_with_boc_and_cboc_replicas():
    """25 MS/s (the BOC(6,1) component needs > 12 MHz), the 8 satellites of configs[0] from the patched reference's
    trace, 4 code periods: a blind FFT acquisition finds every satellite at the records' Doppler and code phase with the
    BOC(1,1) replica (what the bundled receiver conf uses, Acquisition_1B.cboc=false) and with the CBOC replica; the
    matched CBOC replica collects 11/10 of the BOC(1,1) replica's correlation power (alpha^2 = 10/11)."""
    import e1b200 as E
    import rx_e1 as RX
    recs = np.load(GOLD / "fs25_recs.npz")["recs"][:1].copy()
    s = E.Synth(FS25, 2500000, 16, flags=U.CFG_CBOC)
    iq = s.synth_epochs(recs)
    s.close()
    x = iq[:500000, 0].astype(np.float32) + 1j * iq[:500000, 1].astype(np.float32)
    present = [int(p) for p in recs[0]["prn"] if p > 0]
    ratios = []
    for slot, prn in enumerate(recs[0]["prn"]):
        if prn <= 0:
            continue
        r0 = recs[0, slot]
        res = {}
        for name, w in (("boc", None), ("cboc", (U.ALPHA_CBOC, U.BETA_CBOC))):
            m, fd, cp, peak = RX.acquire(x, int(prn), FS25, periods=4, cboc=w, raw=True, f_step=250.0)
            assert m > 20.0, (prn, name, m)
            assert abs(fd - r0["f_carr"]) <= 250.0, (prn, name, fd, r0["f_carr"])
            d = (cp - r0["code_phase0"] + 2046.0) % 4092.0 - 2046.0
            assert abs(d) < 0.1, (prn, name, cp, r0["code_phase0"])
            res[name] = peak
        ratios.append(res["cboc"] / res["boc"])
    assert 1.05 < np.mean(ratios) < 1.15, ratios
    for prn in (3, 10):                                  # absent satellites stay in the noise with either replica
        assert prn not in present
        assert RX.acquire(x, prn, FS25, periods=4, f_step=250.0)[0] < 12.0
