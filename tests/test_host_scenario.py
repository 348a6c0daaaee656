"""Host half of the drop-in (galileo-sdr-sim_b200/host/: RINEX 3 -> ephemerides -> pseudoranges ->
I/NAV pages -> e1_epoch_rec) against the reference: the records it derives must equal, bit for bit,
the channel state the reference itself held at the top of every 0.1 s block (tests/golden/*_recs.npz,
dumped from a build of the reference's own sources by oracle/ref_hooks + tools/make_golden.py), and the
command-line tool built on it must write the reference's file (-m gpu)."""
import hashlib
import subprocess
import sys
from pathlib import Path

import numpy as np
import pytest

import e1util as U

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "galileo-sdr-sim_b200"))
import build as B  # noqa: E402

GOLD = Path(__file__).parent / "golden"
NAV = GOLD / "week171_subset.rnx"
SCENARIOS = {
    # BASELINE configs[0]: -l -6,51,100 -e week171.rnx -d 10 (start = first clock epoch in the file)
    "cfg1": dict(llh=(-6, 51, 100), duration_s=10),
    # tools/make_golden.py's second scenario: crosses the 30 s re-allocation and many page turns
    "paris45": dict(llh=(48.85, 2.35, 35), start=(2021, 6, 20, 11, 59, 40), duration_s=45),
    # the patched builds of the reference (oracle/ref_patches): 25 MS/s; 36 slots with the mask off (24 satellites,
    # the 30 s re-allocation, every page turn); both
    "fs25": dict(llh=(-6, 51, 100), duration_s=3, fs_hz=25e6),
    "ch36": dict(llh=(-6, 51, 100), duration_s=35, max_chan=36, elev_mask_deg=-90.0),
    "fs25ch36": dict(llh=(-6, 51, 100), duration_s=3, fs_hz=25e6, max_chan=36, elev_mask_deg=-90.0),
}


@pytest.fixture(scope="module")
def H():
    B.build_host()
    import e1host
    return e1host


@pytest.mark.parametrize("name", sorted(SCENARIOS))
def test_records_equal_reference_trace(H, name):
    gold = np.load(GOLD / f"{name}_recs.npz")["recs"]
    s = H.Scenario(NAV, **SCENARIOS[name])
    assert s.n_epochs == gold.shape[0]
    recs, grx = s.all()
    s.close()
    n = gold.shape[0] - 1              # the trace cannot know the last block's page_next (no block after it)
    for f in recs.dtype.names:
        a, b = recs[f], gold[f]
        if f == "page_next":
            a, b = a[:n], b[:n]
        assert np.array_equal(a, b), f"{name}.{f}: {np.count_nonzero((a != b).reshape(a.shape[0], -1).any(1))} blocks differ"
    assert np.allclose(np.diff(grx), 0.100000023142)


def test_records_in_pieces_equal_one_call(H):
    a, _ = H.Scenario(NAV, **SCENARIOS["paris45"]).all()
    s = H.Scenario(NAV, **SCENARIOS["paris45"])
    parts = []
    while True:
        r, _ = s.next(37)
        if not len(r):
            break
        parts.append(r)
    assert np.array_equal(np.concatenate(parts), a)


def test_pages_have_sync_and_both_halves_change(H):
    recs, _ = H.Scenario(NAV, **SCENARIOS["paris45"]).all()
    sync = np.array([0, 1, 0, 1, 1, 0, 0, 0, 0, 0], np.uint8)
    act = recs["prn"] > 0
    sym = np.unpackbits(recs["page_cur"][act], axis=-1, bitorder="little")[:, :500]
    assert (sym[:, :10] == sync).all() and (sym[:, 250:260] == sync).all()
    assert len({hashlib.md5(p.tobytes()).hexdigest() for p in sym}) > 20      # page content follows TOW / word type


def test_crc24q_matches_bitwise_definition(H):
    """CRC-24Q (polynomial 0x1864CFB) over a bit string, as the reference's page builder computes it
    for the 196 page bits (src/inav-msg.cpp:134-162; trailing partial byte = 4 bits)."""
    lib = H.load()
    rng = np.random.default_rng(1)

    def ref(bits):
        crc = 0
        for b in bits:
            crc ^= int(b) << 23
            crc <<= 1
            if crc & 0x1000000:
                crc ^= 0x1864CFB
        return crc & 0xFFFFFF

    for n in (196, 12, 100, 53):      # whole-byte lengths take a path of the reference's routine that is never used (and undefined)
        bits = rng.integers(0, 2, n).astype(np.int32)
        assert lib.e1h_crc24q_bits(bits.ctypes.data, n) == ref(bits), n


def test_bad_inputs(H):
    with pytest.raises(RuntimeError):
        H.Scenario(GOLD / "does_not_exist.rnx")
    with pytest.raises(RuntimeError):
        H.Scenario(NAV, start=(2021, 6, 25, 0, 0, 0))         # outside the file's span


def test_range_records_restate_to_the_epoch_records(H):
    """e1h_next_ex hands the same blocks over as pseudoranges; computeCodePhase (src/gal-sig.cpp:308-347)
    applied to them -- e1b200_restate, the host twin of the device kernel -- gives the epoch records'
    f_carr, f_code, code_phase0 and ibit0 bit for bit, pages and phase initialisation are copied."""
    import ctypes as C
    lib = C.CDLL(str(B.build_lib()))
    lib.e1b200_restate.argtypes = [C.c_double] * 4 + [C.c_void_p] * 5
    rng, recs, grx = H.Scenario(NAV, **SCENARIOS["paris45"]).next_ranges(449, with_recs=True)
    gold = np.load(GOLD / "paris45_recs.npz")["recs"]
    assert np.array_equal(recs["f_carr"], gold["f_carr"]) and np.array_equal(rng["prn"], gold["prn"])
    out = [C.c_double() for _ in range(3)]
    ib, ip = C.c_int32(), C.c_int32()
    for e in list(range(0, 449, 7)) + [299, 300, 301]:
        for c in np.nonzero(rng[e]["prn"] > 0)[0]:
            r = rng[e, c]
            assert r["grx_sec"] == grx[e]
            lib.e1b200_restate(r["rho_prev"], r["rho_cur"], 0.100000023142, r["grx_sec"], *[C.byref(x) for x in out], C.byref(ib), C.byref(ip))
            g = recs[e, c]
            assert (out[0].value, out[1].value, out[2].value, ib.value) == (g["f_carr"], g["f_code"], g["code_phase0"], g["ibit0"]), (e, c)
    for f in ("carr_phase_init", "page_cur", "page_next"):
        assert np.array_equal(rng[f], recs[f]), f
    # e1_range_rec has no spare field: E1_REC_* in the low byte of flags, the block's gain[i] above it
    assert np.array_equal(rng["flags"] & 0xFF, recs["flags"]) and np.array_equal(rng["flags"] >> 8, recs["gain_q7"].astype(np.uint32))


def test_receiver_motion_table_and_location_updates(H):
    """Dynamic mode: the reference re-reads the location thread's llhr at every block
    (src/galileo-sdr.cpp:443-448).  A table that repeats the static position changes nothing; a table
    and per-block set_location calls are the same thing; a receiver moving east at 100 m/s shifts every
    channel's Doppler by -(v . line of sight)/lambda, i.e. by at most 100 / 0.1903 = 525 Hz."""
    kw = dict(SCENARIOS["cfg1"])
    static, _ = H.Scenario(NAV, **kw).all()
    n = static.shape[0] + 1
    s = H.Scenario(NAV, **kw)
    s.set_motion(np.tile(np.array(kw["llh"], float), (n, 1)))
    same, _ = s.all()
    assert np.array_equal(same, static)

    lat, lon, h = kw["llh"]
    dlon = np.degrees(100.0 * 0.1 / (6378137.0 * np.cos(np.radians(lat))))       # 10 m per block
    table = np.array([[lat, lon + k * dlon, h] for k in range(n)])
    s = H.Scenario(NAV, **kw)
    s.set_motion(table)
    moving, _ = s.all()
    s = H.Scenario(NAV, **kw)
    parts = []
    for k in range(1, n):
        s.set_location(*table[k])
        parts.append(s.next(1)[0])
    assert np.array_equal(np.concatenate(parts), moving)
    act = static["prn"][5:] > 0
    assert np.array_equal(moving["prn"], static["prn"])
    df = (moving["f_carr"] - static["f_carr"])[5:][act]                          # block 1 sees a position JUMP from xyz[0]
    assert 20.0 < np.abs(df).max() < 526.0 and np.abs(df).min() < 500.0
    assert len(np.unique(np.sign(df))) == 2                                      # satellites ahead and behind


def test_oracle_on_host_records_reproduces_reference_file(H):
    """End to end on the CPU: host records -> oracle sample loop -> the reference's own bytes."""
    recs, _ = H.Scenario(NAV, **SCENARIOS["cfg1"]).all()
    out, _ = U.oracle_synth(U.fs_as_reference(2.6e6), 260000, recs[:20], threads=8)
    sha = (GOLD / "cfg1_sha256.txt").read_text().splitlines()[1:]
    blocks = out.reshape(20, 260000, 2)
    for e in range(20):
        assert hashlib.sha256(blocks[e].tobytes()).hexdigest() == sha[e], e


@pytest.mark.gpu
def test_cli_writes_the_reference_file(tmp_path):
    """e1sim with the reference's own command line for BASELINE configs[0] -> md5 of the reference's output."""
    exe = B.build_cli()
    out = tmp_path / "cfg1.ishort"
    r = subprocess.run([str(exe), "-e", str(NAV), "-l", "-6,51,100", "-d", "10", "-U", "1", "-b", "1", "-o", str(out)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    data = out.read_bytes()
    assert len(data) == 102960000
    assert hashlib.md5(data).hexdigest() == "419622c87f06f4048858bce54df72d29"
    assert "Done!" in r.stderr


@pytest.mark.gpu
def test_cli_device_restate_and_motion_file(tmp_path):
    """-R (pseudoranges to the GPU, computeCodePhase on the device) and -u with a motion file that
    stays put both write the reference's bytes; -r paces a 1 s run to about a second."""
    import time
    exe = B.build_cli()
    out = tmp_path / "r.ishort"
    r = subprocess.run([str(exe), "-e", str(NAV), "-l", "-6,51,100", "-d", "10", "-R", "-o", str(out)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert hashlib.md5(out.read_bytes()).hexdigest() == "419622c87f06f4048858bce54df72d29"
    mot = tmp_path / "static.csv"
    mot.write_text("".join("-6,51,100\n" for _ in range(100)))
    r = subprocess.run([str(exe), "-e", str(NAV), "-u", str(mot), "-o", str(out)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    assert hashlib.md5(out.read_bytes()).hexdigest() == "419622c87f06f4048858bce54df72d29"
    t0 = time.time()
    r = subprocess.run([str(exe), "-e", str(NAV), "-l", "-6,51,100", "-d", "2", "-r", "-B", "2", "-o", str(out)], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0 and len(out.read_bytes()) == 19 * 260000 * 4
    assert time.time() - t0 > 1.5


@pytest.mark.gpu
def test_cli_second_scenario_blocks(tmp_path):
    exe = B.build_cli()
    out = tmp_path / "paris.ishort"
    r = subprocess.run([str(exe), "-e", str(NAV), "-l", "48.85,2.35,35", "-t", "2021/06/20,11:59:40", "-d", "45", "-B", "100", "-o", str(out)],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    sha = (GOLD / "paris45_sha256.txt").read_text().splitlines()[1:]
    blk = 260000 * 4
    with open(out, "rb") as f:
        for e in range(449):
            assert hashlib.sha256(f.read(blk)).hexdigest() == sha[e], e
        assert f.read(1) == b""


@pytest.mark.gpu
@pytest.mark.parametrize("name,extra,n", [
    ("ch36", ["-c", "36", "-m", "-90", "-d", "35"], 260000),
    ("fs25ch36", ["-f", "25e6", "-c", "36", "-m", "-90", "-d", "3"], 2500000),
])
def test_cli_at_the_rates_and_channel_counts_of_the_patched_reference(tmp_path, name, extra, n):
    """e1sim -f / -c / -m: the rates, slot counts and mask the reference fixes at compile time -- the files equal
    the patched reference builds' (oracle/ref_patches), block for block."""
    exe = B.build_cli()
    out = tmp_path / f"{name}.ishort"
    r = subprocess.run([str(exe), "-e", str(NAV), "-l", "-6,51,100", "-o", str(out)] + extra, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = (GOLD / f"{name}_sha256.txt").read_text().splitlines()
    assert hashlib.md5(out.read_bytes()).hexdigest() == lines[0].split()[2]


@pytest.mark.gpu
def test_cli_cboc_and_gain_options(tmp_path):
    """-C / -A select the float path: another stream than the default, the same size, louder satellites louder."""
    exe = B.build_cli()
    outs = {}
    for tag, extra in (("boc", []), ("cboc", ["-C"]), ("gain", ["-A"])):
        out = tmp_path / f"{tag}.ishort"
        r = subprocess.run([str(exe), "-e", str(NAV), "-l", "-6,51,100", "-d", "1", "-o", str(out)] + extra, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        outs[tag] = np.frombuffer(out.read_bytes(), np.int16).astype(np.float64)
    assert outs["boc"].size == outs["cboc"].size == outs["gain"].size == 9 * 260000 * 2
    c = np.corrcoef(outs["boc"], outs["cboc"])[0, 1]
    assert 0.90 < c < 0.99
    assert 0.3 < np.abs(outs["gain"]).mean() / np.abs(outs["boc"]).mean() < 0.7
