"""The reference's live-sky I/NAV pages (tv/<date>/<svid>.csv, sampled into tests/golden/tv_pages.csv by
tools/make_tv_fixture.py) as known answers for the navigation-message side of the drop-in:
pages a real Galileo satellite sent must pass the CRC-24Q both the host page builder and the receiver stand-in compute,
must survive the host's channel coding (tail, rate-1/2 K=7 code, 30 x 8 interleaver, sync) followed by the stand-in's
blind decoder, and must carry week number, time of week, SVID and IODnav at the bit positions the host page builder
(galileo-sdr-sim_b200/host/e1_scenario.cpp, page_symbols; reference: src/inav-msg.cpp:165-409) writes them to."""
import ctypes as C
import sys
from pathlib import Path

import numpy as np

import e1util as U  # noqa: F401  (path set-up)
import rx_e1 as RX

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT / "galileo-sdr-sim_b200"))
import build as B  # noqa: E402

GOLD = Path(__file__).parent / "golden"


def val(bits):
    return int("".join(map(str, bits)), 2)


def tv_pages():
    """-> list of (day, tow, wn, svid, even[114], odd[114], tails[12])."""
    out = []
    for ln in (GOLD / "tv_pages.csv").read_text().splitlines():
        day, tow, wn, svid, hx = ln.split(",")
        b = np.array([int(c) for c in bin(int(hx, 16))[2:].zfill(240)], np.int32)
        out.append((day, int(tow), int(wn), int(svid), b[:114].copy(), b[120:234].copy(), np.concatenate([b[114:120], b[234:240]])))
    return out


def host():
    B.build_host()
    import e1host
    return e1host


def test_live_sky_pages_pass_both_crc_implementations():
    H = host().load()
    pages = tv_pages()
    assert len(pages) == 360
    for day, tow, wn, svid, even, odd, tails in pages:
        assert even[0] == 0 and odd[0] == 1 and even[1] == 0 and odd[1] == 0 and not tails.any(), (day, tow, svid)
        body = np.ascontiguousarray(np.concatenate([even, odd[:82]]), np.int32)          # 196 bits under the CRC
        crc_sent = val(odd[82:106])
        assert RX.crc24q(body) == crc_sent, (day, tow, svid)
        assert H.e1h_crc24q_bits(body.ctypes.data, 196) == crc_sent, (day, tow, svid)
        body[17] ^= 1                                                                     # and it is a check: one flipped bit fails
        assert H.e1h_crc24q_bits(body.ctypes.data, 196) != crc_sent and RX.crc24q(body) != crc_sent


def test_live_sky_pages_through_host_encoder_and_receiver_decoder():
    H = host().load()
    by_sat = {}
    for day, tow, wn, svid, even, odd, tails in tv_pages():
        sym = np.zeros(500, np.int32)
        assert H.e1h_encode_page(np.ascontiguousarray(even).ctypes.data, np.ascontiguousarray(odd).ctypes.data, sym.ctypes.data) == 0
        assert np.array_equal(sym[:10], RX.SYNC) and np.array_equal(sym[250:260], RX.SYNC)
        e, err_e, tail_e = RX.decode_half(sym[:250].astype(np.int8))
        o, err_o, tail_o = RX.decode_half(sym[250:].astype(np.int8))
        assert np.array_equal(e, even) and np.array_equal(o, odd) and err_e == 0 and err_o == 0 and not tail_e.any() and not tail_o.any()
        by_sat.setdefault((day, svid), []).append((sym, even, odd))
    # one satellite's 30 pages as a symbol stream cut at an arbitrary symbol and inverted (the carrier's half-cycle
    # ambiguity): the blind decoder finds the sync, the polarity and every page, CRC included
    (day, svid), pg = sorted(by_sat.items())[5]
    stream = 1 - np.concatenate([p[0] for p in pg]).astype(np.int8)[137:]
    pages, info = RX.decode_pages(stream)
    assert info["inverted"] and info["sync_quality"] == 10.0 and info["sync_offset"] == (500 - 137) % 250
    assert len(pages) == 29 and all(p["crc_ok"] and p["channel_errors"] == 0 and p["tails_zero"] for p in pages)
    sent = [np.concatenate([p[1], p[2]]) for p in pg[1:]]
    assert all(np.array_equal(p["bits"], s) for p, s in zip(pages, sent))


def test_time_svid_and_iod_fields_sit_where_the_host_page_builder_puts_them():
    """Word 0: time = 2, WN at word bits 96..107, TOW behind it (split over the half pages); word 5: WN / TOW behind the
    health bits; word 6: TOW behind dtLSF; word 4: SVID behind IODnav; IODnav at the same place in words 1-4.  Checked on
    the live-sky pages against the CSV's own WN / TOW / SVID columns, then on pages the host builder makes."""
    H = host().load()
    iod = {}
    seen = set()
    for day, tow, wn, svid, even, odd, tails in tv_pages():
        wt = val(even[2:8])
        if wt == 0:
            assert val(even[8:10]) == 2 and val(even[98:110]) == wn and val(np.concatenate([even[110:114], odd[2:18]])) == tow
        elif wt == 5:
            assert val(even[75:87]) == wn and val(even[87:107]) == tow
        elif wt == 6:
            assert val(np.concatenate([even[107:114], odd[2:15]])) == tow
        elif wt == 4:
            assert val(even[18:24]) == svid
        if wt in (1, 2, 3, 4):
            iod.setdefault((day, svid, tow // 30), set()).add(val(even[8:18]))
        seen.add(wt)
    assert {0, 1, 2, 3, 4, 5, 6} <= seen
    assert all(len(v) == 1 for v in iod.values())                 # one issue of data per 30 s sub-frame
    # the host builder, same positions (any scenario that holds an ephemeris of the satellite will do: the time
    # fields come from the requested page time; GST week = week - 1024, src/inav-msg.cpp:190)
    e1host = host()
    s = e1host.Scenario(GOLD / "week171_subset.rnx", llh=(-6, 51, 100), duration_s=2)
    week, prn = 2163, 1
    got = set()
    for tow in range(1, 61, 2):                                   # one 60 s cycle of word types from 2021-06-20 00:00:01
        sym = np.zeros(500, np.int32)
        assert H.e1h_page_symbols(s._h, prn, float(tow), week, sym.ctypes.data) == 0
        even, e0, _ = RX.decode_half(sym[:250].astype(np.int8))
        odd, e1, _ = RX.decode_half(sym[250:].astype(np.int8))
        assert e0 == 0 and e1 == 0 and even[0] == 0 and odd[0] == 1
        wt = val(even[2:8])
        got.add(wt)
        if wt == 0:
            assert val(even[8:10]) == 2 and val(even[98:110]) == week - 1024 and val(np.concatenate([even[110:114], odd[2:18]])) == tow
        elif wt == 5:
            assert val(even[75:87]) == week - 1024 and val(even[87:107]) == tow
        elif wt == 6:
            assert val(np.concatenate([even[107:114], odd[2:15]])) == tow
        elif wt == 4:
            assert val(even[18:24]) == prn
        body = np.ascontiguousarray(np.concatenate([even, odd[:82]]), np.int32)
        assert RX.crc24q(body) == val(odd[82:106])
    assert {0, 1, 2, 3, 4, 5, 6} <= got
    s.close()
