"""Receiver-side check of the generated stream: tests/rx_e1.py acquires,
tracks and decodes BASELINE configs[0] blind.  What must come out: exactly the satellites the
scenario allocated, at the Doppler and code phase of the first block's records; the symbol stream of
every channel equal to the I/NAV symbols that went in; pages that pass Viterbi decoding with zero
channel errors and the CRC-24Q of the 196 page bits, with the word types of the reference's
allocation table.  The CPU variant feeds the receiver from the oracle's sample loop, the GPU variant
from the CUDA path (same records)."""
import sys
from pathlib import Path

import numpy as np
import pytest

import e1util as U
import rx_e1 as RX

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "galileo-sdr-sim_b200"))
import build as B  # noqa: E402

GOLD = Path(__file__).parent / "golden"
NAV = GOLD / "week171_subset.rnx"
FS = U.fs_as_reference(2.6e6)
N = 260000
BLOCKS = 45                      # 4.5 s: one complete even + odd page pair even if the first 0.3 s are pull-in


def scenario_records():
    B.build_host()
    import e1host
    s = e1host.Scenario(NAV, llh=(-6, 51, 100), duration_s=10)
    recs, grx = s.next(BLOCKS)
    s.close()
    return recs, grx


def check_stream(iq, recs):
    x = iq[:, 0].astype(np.float32) + 1j * iq[:, 1].astype(np.float32)
    present = sorted(int(p) for p in recs[0]["prn"] if p > 0)
    # --- acquisition: every PRN 1..36 is tried; the allocated ones and only those stand out
    found = {}
    metrics = {}
    for prn in range(1, 37):
        m, fd, cp = RX.acquire(x, prn, FS)
        metrics[prn] = m
        if m > 20.0:
            found[prn] = (fd, cp)
    assert sorted(found) == present, (sorted(found), present, {k: round(v, 1) for k, v in metrics.items()})
    n_periods = len(x) // 10400
    for slot, prn in enumerate(recs[0]["prn"]):
        if prn <= 0:
            continue
        prn = int(prn)
        fd, cp = found[prn]
        r0 = recs[0, slot]
        assert abs(fd - r0["f_carr"]) <= 125.0, (prn, fd, r0["f_carr"])
        d = (cp - r0["code_phase0"] + 2046.0) % 4092.0 - 2046.0
        assert abs(d) < 1.0, (prn, cp, r0["code_phase0"])
        # --- tracking: stays locked for the whole stream, Doppler follows the records
        prompts, f_hist, cp_hist, start = RX.track(x, prn, FS, fd, cp, n_periods)
        assert len(prompts) >= n_periods - 2
        amp = np.abs(prompts[25:])
        assert amp.min() > 0.7 * amp.mean(), (prn, amp.min(), amp.mean())           # no loss of lock
        f_true = np.repeat(recs[:, slot]["f_carr"], 25)[:len(f_hist)]
        assert np.abs(f_hist[75:] - f_true[75:]).max() < 5.0, prn
        # --- symbols: what went in comes out.  The receiver's symbol k is the code period that starts at
        # sample `start` + k * 10400: block e = sample // 260000 of that instant, symbol index = the
        # block's ibit0 plus the code periods begun inside the block up to there
        sym = RX.symbols_from_prompts(prompts)
        bits_in = np.unpackbits(recs[:, slot]["page_cur"], axis=-1, bitorder="little")[:, :500]
        nxt_in = np.unpackbits(recs[:, slot]["page_next"], axis=-1, bitorder="little")[:, :500]
        sent = []
        for k in range(len(sym)):
            mid = start + k * 10400 + 5200                  # a sample safely inside the symbol
            e = mid // N
            r = recs[e, slot]
            wraps = int(np.floor((r["code_phase0"] + (mid - e * N) * r["f_code"] / FS) / 4092.0))
            j = int(r["ibit0"]) + wraps
            sent.append(nxt_in[e, j - 500] if j >= 500 else bits_in[e, j])
        sent = np.array(sent, np.int8)
        agree = (sym[75:] == sent[75:]).mean()
        assert agree == 1.0 or agree == 0.0, (prn, agree)                          # global sign is the carrier's
        # --- pages: blind sync, de-interleave, Viterbi, CRC
        pages, info = RX.decode_pages(sym)
        assert info["sync_quality"] == 10.0, (prn, info)
        assert len(pages) >= 1, (prn, info)
        for pg in pages:
            assert pg["crc_ok"] and pg["channel_errors"] == 0 and pg["tails_zero"], (prn, pg["start"], pg["word_type"])
            assert pg["word_type"] in (0, 1, 2, 3, 4, 5, 6, 63)
            assert int("".join(map(str, pg["bits"][220:228])), 2) in (4, 43, 47)    # SSP byte, src/inav-msg.cpp:393-395
    return found


def sent_bits(recs, slot, start, n_sym):
    """Page bit (0/1) and page symbol index of the receiver's symbols k = 0 .. n_sym-1 (code periods from sample `start`)."""
    bits_in = np.unpackbits(recs[:, slot]["page_cur"], axis=-1, bitorder="little")[:, :500]
    nxt_in = np.unpackbits(recs[:, slot]["page_next"], axis=-1, bitorder="little")[:, :500]
    sent, idx = [], []
    for k in range(n_sym):
        mid = start + k * 10400 + 5200
        e = mid // N
        r = recs[e, slot]
        wraps = int(np.floor((r["code_phase0"] + (mid - e * N) * r["f_code"] / FS) / 4092.0))
        j = int(r["ibit0"]) + wraps
        sent.append(nxt_in[e, j - 500] if j >= 500 else bits_in[e, j])
        idx.append(j % 500)
    return np.array(sent, np.int8), np.array(idx)


def check_pilot(iq, recs, found, prns=None):
    """What GNSS-SDR does with the bundled configuration (track_pilot=true): the loops run on E1-C.  The pilot must carry
    the 25-chip secondary code, aligned to the page symbols (symbol index mod 25), with the ICD's sign (the composite is
    e_B d - e_C s); wiping it off resolves the carrier's half-cycle ambiguity, so the data symbols demodulated with the
    pilot's carrier must equal the page bits that went in ABSOLUTELY, not just up to a global sign -- a sign or
    secondary-code error in the E1-C half of the generator cannot hide here -- and the pages must still decode."""
    x = iq[:, 0].astype(np.float32) + 1j * iq[:, 1].astype(np.float32)
    n_periods = len(x) // 10400
    for slot, prn in enumerate(recs[0]["prn"]):
        prn = int(prn)
        if prn <= 0 or (prns is not None and prn not in prns):
            continue
        fd, cp = found[prn] if found else (float(recs[0, slot]["f_carr"]), float(recs[0, slot]["code_phase0"]))
        pp, f_hist, cp_hist, start, dp = RX.track(x, prn, FS, fd, cp, n_periods, pilot=True)
        assert len(pp) >= n_periods - 2
        amp = np.abs(pp[25:])
        assert amp.min() > 0.7 * amp.mean(), (prn, amp.min(), amp.mean())           # pilot stays in lock
        f_true = np.repeat(recs[:, slot]["f_carr"], 25)[:len(f_hist)]
        assert np.abs(f_hist[75:] - f_true[75:]).max() < 5.0, prn
        bits, shift, agree = RX.pilot_symbols(pp, dp)
        assert agree == 1.0, (prn, agree)                                            # every pilot prompt is a secondary-code chip
        sent, idx = sent_bits(recs, slot, start, len(bits))
        assert shift == idx[0] % 25, (prn, shift, idx[0])                            # ... aligned to the page symbols
        assert np.array_equal(bits[75:], sent[75:]), (prn, float((bits[75:] == sent[75:]).mean()))   # absolute polarity
        pages, info = RX.decode_pages(bits)
        assert not info["inverted"] and info["sync_quality"] == 10.0 and len(pages) >= 1, (prn, info)
        for pg in pages:
            assert pg["crc_ok"] and pg["channel_errors"] == 0 and pg["tails_zero"], (prn, pg["start"], pg["word_type"])


def test_receiver_acquires_tracks_and_decodes_oracle_stream():
    recs, _ = scenario_records()
    iq, _ = U.oracle_synth(FS, N, recs, threads=8)
    check_stream(iq, recs)


def test_pilot_tracking_secondary_code_and_absolute_symbols_oracle_stream():
    recs, _ = scenario_records()
    iq, _ = U.oracle_synth(FS, N, recs, threads=8)
    check_pilot(iq, recs, None)


@pytest.mark.gpu
def test_receiver_acquires_tracks_and_decodes_cuda_stream():
    import e1b200 as E
    recs, _ = scenario_records()
    s = E.Synth(FS, N, 16)
    iq = s.synth_epochs(recs)
    s.close()
    found = check_stream(iq, recs)
    check_pilot(iq, recs, found)
