"""Pins the CPU oracle (oracle/e1_oracle.c) to the reference's own output.

The fixtures in tests/golden/ were produced by tools/make_golden.py from a build of the
reference's unmodified arithmetic (oracle/Makefile `ref`): per-0.1 s-block SHA-256 of the ishort
file, raw sample slices, and the channel state the reference held at the top of every block.
"""
import hashlib
from pathlib import Path

import numpy as np
import pytest

import e1util as U

GOLD = Path(__file__).parent / "golden"
N = 260000
FS = U.fs_as_reference(2.6e6)


def load(name):
    z = np.load(GOLD / f"{name}_recs.npz")
    lines = (GOLD / f"{name}_sha256.txt").read_text().splitlines()
    return z["recs"], z["phase"], lines[0], lines[1:]


@pytest.mark.parametrize("name,epochs", [("cfg1", slice(0, 99)), ("paris45", slice(0, 40))])
def test_oracle_matches_reference_blocks(name, epochs):
    recs, phase, header, sha = load(name)
    recs = recs[epochs]
    out, _ = U.oracle_synth(FS, N, recs)
    out = out.reshape(recs.shape[0], N, 2)
    for e in range(recs.shape[0]):
        assert hashlib.sha256(out[e].tobytes()).hexdigest() == sha[epochs.start + e], f"{name} block {e}"


# Patched builds of the reference (oracle/ref_patches/*.diff applied by `make -C oracle refp`; fixtures by
# tools/make_golden.py): the rates / channel counts of BASELINE configs[1]-[4].
PATCHED = {  # name: (fs, samples per block, MAX_CHAN, satellites, blocks, md5 of the reference's file)
    "fs25": (U.fs_as_reference(25e6), 2500000, 16, 8, 29, "038c859da001e2cf45c0528f3864adc1"),
    "ch36": (FS, 260000, 36, 24, 349, "09ea87819b295a90d2295b0554d7a5c2"),
    "fs25ch36": (U.fs_as_reference(25e6), 2500000, 36, 24, 29, "542bc0887b3f81897f2af436ac42ab89"),
}


@pytest.mark.parametrize("name,epochs", [("fs25", slice(0, 12)), ("ch36", slice(0, 349)), ("fs25ch36", slice(0, 10))])
def test_oracle_matches_patched_reference_blocks(name, epochs):
    """25 MS/s, 24 satellites in 36 slots with the elevation mask off (35 s: every channel turns its page,
    the 30 s re-allocation runs), and both together: the oracle's blocks equal, hash for hash, the bytes the
    reference ITSELF wrote when built with those constants.  The whole ch36 file (its md5) is checked here;
    the 25 MS/s ones in full on the GPU side (tests/test_gpu_parity.py), a slice here to keep the CPU suite short."""
    fs, n, max_chan, n_sat, n_blocks, md5 = PATCHED[name]
    recs, phase, header, sha = load(name)
    assert recs.shape == (n_blocks, max_chan) and header.split()[2] == md5
    assert len(set(int(x) for x in recs["prn"].ravel()) - {0}) == n_sat
    ph0 = None
    r = recs[epochs]
    out, _ = U.oracle_synth(fs, n, r, ph0, threads=8)
    out = out.reshape(r.shape[0], n, 2)
    for e in range(r.shape[0]):
        assert hashlib.sha256(out[e].tobytes()).hexdigest() == sha[epochs.start + e], f"{name} block {e}"
    if r.shape[0] == n_blocks:
        assert hashlib.md5(out.tobytes()).hexdigest() == md5


def test_oracle_cfg1_file_md5():
    recs, _, header, _ = load("cfg1")
    out, _ = U.oracle_synth(FS, N, recs)
    assert hashlib.md5(out.tobytes()).hexdigest() == header.split()[2] == "419622c87f06f4048858bce54df72d29"
    # facts recorded in BASELINE.md section 2
    assert out.shape[0] == 99 * N
    assert out[:4].tolist() == [[-606, 1130], [-638, 374], [16, 404], [172, -868]]
    assert int(out[:, 0].min()) == -3646 and int(out[:, 0].max()) == 3826
    assert not (out & 1).any()


def test_oracle_carries_reference_carrier_phase():
    """The phase the oracle integrates must equal, bit for bit, the phase the reference held at the
    top of every later block (chan[i].carr_phase, src/galileo-sdr.cpp:531-532)."""
    recs, phase, _, _ = load("cfg1")
    ph = np.zeros(16)
    for e in range(12):
        active = recs[e]["prn"] > 0
        if e > 0:
            assert np.array_equal(ph[active], phase[e][active]), f"epoch {e}"
        _, ph = U.oracle_synth(FS, N, recs[e:e + 1], ph)


def test_oracle_sample_slices():
    for name in ("cfg1", "paris45"):
        recs, _, _, _ = load(name)
        z = np.load(GOLD / f"{name}_samples.npz")
        last = int(z["epochs"][1]) if name == "paris45" else int(z["epochs"].max())
        out, _ = U.oracle_synth(FS, N, recs[:last + 1])
        out = out.reshape(-1, N, 2)
        for i, e in enumerate(z["epochs"]):
            if e <= last:
                assert np.array_equal(out[e, :4096], z["head"][i])
                assert np.array_equal(out[e, -4096:], z["tail"][i])


def test_mt_oracle_identical():
    recs = U.synthetic_recs(3, 12, FS, seed=5)
    a, pa = U.oracle_synth(FS, 26000, recs)
    b, pb = U.oracle_synth(FS, 26000, recs, threads=5)
    assert np.array_equal(a, b) and np.array_equal(pa, pb)


def test_carrier_lut_and_code_tables():
    import ctypes as C
    lib = U.oracle()
    c, s = (C.c_int * 512)(), (C.c_int * 512)()
    lib.e1o_carrier_lut(c, s)
    c, s = np.array(c), np.array(s)
    assert c[0] == 250 and s[0] == 2 and c[92] == 105 and s[35] == 105 and c[163] == -105 and s[476] == -105
    assert np.abs(c).max() == 250 and np.abs(c * c + s * s - 62500).max() < 400
    t = (C.c_short * 8184)()
    lib.e1o_halfchip_table(1, 0, t)
    t = np.array(t)
    # E1-B PRN 1 starts with hex F5D7... -> bits 1111 0101 -> chips -1 -1 -1 -1 +1 -1 +1 -1
    assert (t[1:16:2] == [-1, -1, -1, -1, 1, -1, 1, -1]).all() and (t[0::2] == -t[1::2]).all()


def test_restate_self_consistency():
    """computeCodePhase restatement: the documented formulas on one hand-made pseudorange pair (the comparison
    with the reference's traced f_carr / code_phase / ibit is tests/test_host_scenario.py, which derives them
    from the RINEX file through the host pipeline and holds them bit-identical to the trace)."""
    fc, fcode, cp, ib, ip = U.oracle_restate(25771939.7, 25771939.7 - 16.5, U.REF_DT, 0.2000000462)
    assert abs(fc - (16.5 / U.REF_DT / 0.1902936727983649)) < 1e-6
    assert abs(fcode - (1.023e6 + fc * 0.0006493506493506494)) < 1e-9
    assert 0 <= cp < 4092 and 0 <= ib < 500
