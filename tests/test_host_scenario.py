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
