"""Host-side checks of the product's exact arithmetic (galileo-sdr-sim_b200/csrc/e1_core.h): the
binade-jump walkers against the literal reference recurrences, and the whole
planner -> closed form -> ambiguity fallback chain against the CPU oracle and the reference's
golden blocks.  Runs without a GPU (the header is compiled for the host by tests/hostsim)."""
import ctypes as C
import hashlib
from pathlib import Path

import numpy as np
import pytest

import e1util as U

GOLD = Path(__file__).parent / "golden"
FS26 = U.fs_as_reference(2.6e6)
FS25 = U.fs_as_reference(25e6)


def test_carrier_walker_equals_literal_loop():
    hs = U.hostsim()
    rng = np.random.default_rng(11)
    for fs in (FS26, FS25):
        delt = 1.0 / fs
        for i in range(150):
            f = rng.uniform(-5000, 5000) if i % 3 else rng.uniform(-20, 20)
            if i % 50 == 0:
                f = 0.0
            sp = f * delt
            phi = rng.uniform(-1, 1) if i % 2 else rng.uniform(0, 1) * (1 if f >= 0 else -1)
            n = int(rng.integers(1, 400000))
            assert hs.hs_carr_advance(phi, sp, 0, n) == hs.hs_carr_literal(phi, sp, n), (fs, f, phi, n)


def test_carrier_walker_ties_and_sign_changes():
    """Steps with many trailing zero bits make exact half-ulp ties in several binades; opposite signs of
    phase and Doppler walk the magnitude down through zero."""
    hs = U.hostsim()
    for sp in (2.0 ** -10, 2.0 ** -10 + 2.0 ** -53, 3 * 2.0 ** -12, 2.0 ** -9 + 2.0 ** -54, 1e-3, 0.25, 0.499, 2.0 ** -60, 1e-300):
        for phi in (0.0, 0.3, 0.999999, 2.0 ** -30, 0.5 + 2.0 ** -53, 1 - 2.0 ** -53):
            for s in (1, -1):
                for ps in (1, -1):
                    n = 50003
                    assert hs.hs_carr_advance(ps * phi, s * sp, 0, n) == hs.hs_carr_literal(ps * phi, s * sp, n), (sp, phi, s, ps)


def test_carrier_walker_adversarial_steps():
    """The jump length comes from RN_even(t / ulp) without probing: steps that are exact half-ulp ties in
    one binade (lowest set bit of t = half that binade's ulp), steps with few mantissa bits, steps so
    small the integer-division path or the stuck path is taken, and steps of nearly half a cycle."""
    hs = U.hostsim()
    rng = np.random.default_rng(99)
    steps = []
    for e in range(-30, -1):                      # one set bit, and neighbours that make ties in the binade two above
        steps += [2.0 ** e, 2.0 ** e + 2.0 ** (e - 52), 2.0 ** e - 2.0 ** (e - 53), 3 * 2.0 ** e * 0.25 + 2.0 ** -54]
    steps += [k * 2.0 ** -54 for k in (1, 3, 5, 2 ** 20 + 1, 2 ** 43 + 1, 2 ** 44 + 3)]   # ties in [1/2, 1)
    steps += [k * 2.0 ** -55 for k in (1, 3, 2 ** 40 + 1)]                                 # ties in [1/4, 1/2)
    steps += [1e-12, 7.3e-9, 2.0 ** -27 * 1.0000001, 2.0 ** -26, 1e-7, 0.4999999, 0.3333333333333333]
    steps += list(rng.uniform(0, 2e-3, 40)) + list(rng.uniform(0, 1e-6, 20))
    for sp in steps:
        for phi in (0.0, rng.uniform(0, 1), 1 - 2.0 ** -53, 2.0 ** -52, 0.5, 0.25 - 2.0 ** -55):
            n = int(rng.integers(1000, 30000)) if sp > 1e-4 else int(rng.integers(1000, 200000))
            for sg in (1, -1):
                assert hs.hs_carr_advance(sg * phi, sg * sp, 0, n) == hs.hs_carr_literal(sg * phi, sg * sp, n), (sp, phi, n, sg)


def test_code_walker_adversarial_steps():
    hs = U.hostsim()
    rng = np.random.default_rng(98)
    for sc in (0.5, 0.25, 0.39346153846153847, 2.0 ** -2 + 2.0 ** -43, 2.0 ** -2 + 2.0 ** -42, 1.0, 1.5, 0.04092, 2.0 ** -20,
               0.4092 + 2.0 ** -41, 3.0 * 2.0 ** -3, 1e-9):
        for cp in (0.0, 4091.999999, 2048.0, 2047.9999999999998, rng.uniform(0, 4092), 2.0 ** -40):
            n = 60000
            w1, w2 = C.c_long(), C.c_long()
            a = hs.hs_code_advance(cp, sc, n, w1)
            b = hs.hs_code_literal(cp, sc, n, w2)
            assert a == b and w1.value == w2.value, (sc, cp)


def test_code_walker_equals_literal_loop():
    hs = U.hostsim()
    rng = np.random.default_rng(12)
    for fs in (FS26, FS25, 4.0e6):
        delt = 1.0 / fs
        for i in range(150):
            f = rng.uniform(-5000, 5000)
            sc = (1.023e6 + f * 0.0006493506493506494) * delt
            cp = rng.uniform(0, 4092) if i % 5 else rng.uniform(0, 1e-6)
            n = int(rng.integers(1, 400000))
            w1, w2 = C.c_long(), C.c_long()
            a = hs.hs_code_advance(cp, sc, n, w1)
            b = hs.hs_code_literal(cp, sc, n, w2)
            assert a == b and w1.value == w2.value, (fs, f, cp, n)


def test_fixed_point_conversion():
    hs = U.hostsim()
    assert hs.hs_to_fixed(0.5, 64) == 1 << 63
    assert hs.hs_to_fixed(-0.25, 64) == 1 << 62
    assert hs.hs_to_fixed(4091.5, 52) == int(4091.5 * 2 ** 52)
    assert hs.hs_to_fixed(1e-30, 64) == 0
    assert hs.hs_to_fixed(2.0 ** -64, 64) == 1
    x = 0.123456789012345678
    assert hs.hs_to_fixed(x, 64) == int(x * 2 ** 64)


def test_code_table_matches_oracle_halfchips():
    """Packed half-chip table (product layout: first half-chip in the top bits; high bit = 1 where the
    E1-B half-chip is -1, low bit = high bit XOR (E1-C half-chip is -1)) -> the oracle's 8184-entry
    BOC(1,1) tables, all 50 PRNs."""
    hs = U.hostsim()
    W = hs.hs_code_words_per_prn()
    codes = np.zeros(50 * W, np.uint32)
    hs.hs_build_codes(codes.ctypes.data)
    lib = U.oracle()
    tb, tc = (C.c_short * 8184)(), (C.c_short * 8184)()
    hh = np.arange(8184)
    for prn in range(1, 51):
        f = (codes[(prn - 1) * W + (hh >> 4)] >> (30 - (hh & 15) * 2)) & 3
        lib.e1o_halfchip_table(prn, 0, tb)
        lib.e1o_halfchip_table(prn, 1, tc)
        eb, ec = np.array(tb), np.array(tc)          # +-1 including the sub-carrier sign
        bneg, cneg = f >> 1, (f >> 1) ^ (f & 1)
        assert np.array_equal(eb, 1 - 2 * bneg.astype(np.int16)), prn
        assert np.array_equal(ec, 1 - 2 * cneg.astype(np.int16)), prn
        assert not codes[(prn - 1) * W + 512:(prn - 1) * W + W].any()


def test_carrier_table_layout():
    """int32[642][32]: 2*(cos + 65536*sin) of the reference's tables, 32 copies per entry (one per lane).
    Entry E = e + 64: e in [0, 512] -> table index e & 511; below 0 -> 511 + e; above 512 -> e - 511
    (where runs that are about to wrap start; see E1C_LUT_IDX in e1_core.h).  Checked against what the
    reference reads (src/galileo-sdr.cpp:509-510) for every index of a wrapped and an unwrapped phase."""
    c, s = (C.c_int * 512)(), (C.c_int * 512)()
    U.oracle().e1o_carrier_lut(c, s)
    c, s = np.array(c), np.array(s)
    lut = U.product_lut().reshape(642, 32)
    w2 = 2 * (c + 65536 * s)
    assert (lut == lut[:, :1]).all()
    L, EXT = lut[:, 0], 64
    i = np.arange(511)                      # trunc(511 |phi|) of a wrapped phase
    assert np.array_equal(L[EXT + i], w2[i])                    # phi >= 0: index i
    assert np.array_equal(L[EXT + 512 - i], w2[(-i) & 511])     # phi < 0: index (-i) & 511
    m = np.arange(1, EXT + 1)               # runs that start m index steps before the wrap
    assert np.array_equal(L[EXT - m], w2[511 - m])              # phi >= 0, i = 511 - m, read at e = i - 511
    assert np.array_equal(L[EXT + 512 + m], w2[(-(511 - m)) & 511])   # phi < 0, read at e = 512 - i + 511


@pytest.mark.parametrize("name,epochs", [("cfg1", (0, 1, 2, 28, 29, 30, 98)), ("paris45", (0, 1, 190, 191, 300, 301, 448))])
def test_pipeline_matches_reference_blocks(name, epochs):
    """Planner + closed form + fallback reproduce the reference's own 0.1 s blocks (SHA-256 fixtures
    generated from oracle/_ref), starting each probe from the reference's traced carrier phase."""
    z = np.load(GOLD / f"{name}_recs.npz")
    recs, phase = z["recs"], z["phase"]
    sha = (GOLD / f"{name}_sha256.txt").read_text().splitlines()[1:]
    N = 260000
    for e in epochs:
        ph = np.where(np.isnan(phase[e]), 0.0, phase[e])
        out, _, st = U.hostsim_synth(FS26, N, recs[e:e + 1], ph)
        assert hashlib.sha256(out.tobytes()).hexdigest() == sha[e], (name, e, st)


def test_pipeline_carries_phase_like_the_oracle():
    z = np.load(GOLD / "cfg1_recs.npz")
    recs = z["recs"][:5]
    a, pa = U.oracle_synth(FS26, 260000, recs)
    b, pb, st = U.hostsim_synth(FS26, 260000, recs)
    assert np.array_equal(a, b) and np.array_equal(pa, pb)
    # reference's own traced phase at the top of epoch 4
    act = recs[4]["prn"] > 0
    _, p4, _ = U.hostsim_synth(FS26, 260000, recs[:4])
    assert np.array_equal(p4[act], z["phase"][4][act])


@pytest.mark.parametrize("fs,n_samp,n_chan,groups", [(FS26, 26000, 9, 4), (FS25, 250000, 12, 4), (FS25, 100001, 5, 2),
                                                       (4.0e6, 40000, 3, 1)])
def test_pipeline_matches_oracle_synthetic(fs, n_samp, n_chan, groups):
    recs = U.synthetic_recs(4, n_chan, fs, seed=n_chan, max_chan=n_chan + 2)
    a, pa = U.oracle_synth(fs, n_samp, recs)
    b, pb, st = U.hostsim_synth(fs, n_samp, recs, groups=groups)
    assert np.array_equal(a, b) and np.array_equal(pa, pb), st


def test_result_independent_of_ambiguity_threshold():
    """Inflating the ambiguity bound only sends more samples through the exact walk: the output
    must not change.  This is the check that the closed form and the walk agree where it matters."""
    recs = U.synthetic_recs(2, 10, FS26, seed=21)
    a, _ = U.oracle_synth(FS26, 52000, recs)
    counts = []
    for scale in (1, 1000, 100000, 10000000):
        b, _, st = U.hostsim_synth(FS26, 52000, recs, amb_scale=scale)
        assert np.array_equal(a, b), scale
        counts.append(int(st[0]))
    assert counts == sorted(counts) and counts[-1] > counts[0]


def test_doppler_sign_flip_and_phase_reset_mid_run():
    """f_carr changes sign across epochs (phase and Doppler of opposite sign -> magnitude walks down
    through zero), a channel is re-initialised mid-run, another goes idle and comes back."""
    fs, N = FS26, 30000
    recs = U.synthetic_recs(8, 4, fs, seed=33, max_chan=5)
    for e in range(8):
        recs[e, 0]["f_carr"] = (3.5 - e) * 700.0          # crosses zero between epochs 3 and 4
        recs[e, 0]["f_code"] = 1.023e6 + recs[e, 0]["f_carr"] * 0.0006493506493506494
        recs[e, 1]["f_carr"] = -(3.5 - e) * 0.4           # tiny Doppler, sign flip
        recs[e, 1]["f_code"] = 1.023e6 + recs[e, 1]["f_carr"] * 0.0006493506493506494
    recs[5, 2]["flags"] = U.E1_REC_SET_PHASE
    recs[5, 2]["carr_phase_init"] = 0.987654321
    recs[2:4, 3]["prn"] = 0                               # idle for two epochs
    recs[6, 3]["f_carr"] = 0.0
    a, pa = U.oracle_synth(fs, N, recs)
    b, pb, st = U.hostsim_synth(fs, N, recs)
    assert np.array_equal(a, b) and np.array_equal(pa, pb), st


def test_fast_path_edges_of_the_carrier_table():
    """Doppler just inside / outside the fast path's step bound (a run of 16 samples may move the
    table index by at most 63 entries: 20 kHz at 2.6 MS/s), both signs and both phase signs, so runs
    start anywhere in the table's extensions and wrap inside; plus steps of +-1 table index per
    sample and exactly zero."""
    fs, N = FS26, 26000
    sp_max = 63.0 / (511.0 * 16)
    f_edge = sp_max * fs
    freqs = [0.999 * f_edge, -0.999 * f_edge, 1.001 * f_edge, -1.001 * f_edge, fs / 511.0, -fs / 511.0, 0.5 * f_edge, -0.5 * f_edge, 0.0]
    recs = U.synthetic_recs(3, len(freqs), fs, seed=5)
    for c, f in enumerate(freqs):
        recs[:, c]["f_carr"] = f
        recs[:, c]["f_code"] = 1.023e6 + f * 0.0006493506493506494
        recs[0, c]["carr_phase_init"] = (0.3 + 0.07 * c) * (1 if c % 3 else -1)   # phase sign independent of Doppler sign
    a, pa = U.oracle_synth(fs, N, recs)
    b, pb, st = U.hostsim_synth(fs, N, recs)
    assert np.array_equal(a, b) and np.array_equal(pa, pb), st


def test_mirrored_run_that_starts_on_an_ambiguous_sample_stays_inside_the_table():
    """Negative phase, first sample of a run with trunc(511 |phi|) = 447 (the lowest index from which a run
    is started 511 entries up because it may wrap) and a fraction below the ambiguity limit: the bias
    that marks the sample ambiguous carries into the entry word, one past the highest entry any
    unambiguous sample reads.  The run is redone exactly, but the fast form's terms come back out by
    recomputing them, so that read must be inside the table: on the GPU the bytes behind the table are
    a parameter buffer that another team's bulk copy rewrites (one wrong sample per ~10 runs of 780 M
    samples before the table got its extra entry).  hostsim_synth asserts that no lookup left the table."""
    fs, N = FS26, 20000
    for frac in (1e-9, 5e-8, 1.5e-6):
        for i0 in (446, 447, 448, 510):
            recs = U.synthetic_recs(1, 2, fs, seed=1)
            recs[0, 0]["f_carr"] = -2500.0
            recs[0, 0]["f_code"] = 1.023e6 - 2500.0 * 0.0006493506493506494
            recs[0, 0]["carr_phase_init"] = -(i0 + frac) / 511.0
            a, pa = U.oracle_synth(fs, N, recs)
            b, pb, st = U.hostsim_synth(fs, N, recs)
            assert np.array_equal(a, b) and np.array_equal(pa, pb), (frac, i0, st)


# ------------------------------------------------------------------ parallel carrier planner
@pytest.mark.parametrize("fs,n_samp,n_chan,n_epochs,seed", [(FS26, 260000, 36, 150, 1), (FS25, 2500000, 36, 40, 2),
                                                             (4.0e6, 400000, 8, 60, 3)])
def test_parallel_planner_equals_serial_walk(fs, n_samp, n_chan, n_epochs, seed):
    """Every tile checkpoint (hat + translation) and the carried phase of the parallel planner equal
    the serial exact walk bit for bit; nearly all epochs are accepted without a serial walk."""
    recs = U.synthetic_recs_fast(n_epochs, n_chan, fs, seed=seed)
    bad, (serial, hat, active) = U.hostsim_plan_compare(fs, n_samp, recs)
    assert bad == 0
    # counts are in planner units (spans: an epoch is cut into S of them)
    assert active % (n_chan * n_epochs) == 0 and 1 <= active // (n_chan * n_epochs) <= 8
    assert hat > 0.95 * (active - n_chan) and serial < 0.05 * active


def low_doppler_recs(n_epochs, fs):
    """Dopplers that sit near zero or run through it: spans without a carrier wrap, anchors several
    spans back, sign changes, a channel below the look-back reach."""
    recs = U.synthetic_recs_fast(n_epochs, 8, fs, seed=3)
    e = np.arange(n_epochs)
    for c, (f0, rate) in enumerate([(30, -0.1), (-5, 0.02), (0.3, 0.0), (100, -0.5), (-0.7, 0.01), (12, 0.0), (-45, 0.15), (2, -0.01)]):
        f = f0 + rate * e
        recs[:, c]["f_carr"] = f
        recs[:, c]["f_code"] = 1.023e6 + f * 0.0006493506493506494
    return recs


def test_parallel_planner_low_doppler_anchors_further_back():
    recs = low_doppler_recs(600, FS26)
    bad, (serial, hat, active) = U.hostsim_plan_compare(FS26, 260000, recs)
    assert bad == 0 and hat > 0.8 * active
    a, pa = U.oracle_synth(FS26, 26000, recs[:40], threads=8)
    b, pb, st = U.hostsim_synth(FS26, 26000, recs[:40])
    assert np.array_equal(a, b) and np.array_equal(pa, pb), st


def test_parallel_planner_on_reference_trace():
    z = np.load(GOLD / "paris45_recs.npz")
    bad, (serial, hat, active) = U.hostsim_plan_compare(FS26, 260000, z["recs"])
    assert bad == 0 and hat > 0.98 * active - 20


def _step_multiple_of_2m53(fs, f_target):
    """A Doppler whose per-sample phase step fl(f*delt) is an exact multiple of 2^-53 (round-half-even
    ties at the wrap step: the planner must walk such epochs serially)."""
    delt = 1.0 / fs
    k0 = round(f_target * delt * 2.0 ** 53)
    for k in range(k0, k0 + 4000):        # not every product is reachable: try neighbouring multiples
        sp = k * 2.0 ** -53
        f = sp / delt
        for _ in range(8):
            got = f * delt
            if got == sp:
                return f
            f = np.nextafter(f, np.inf if got < sp else -np.inf)
    raise AssertionError("no Doppler found")


def test_parallel_planner_ties_idle_gaps_resets_and_low_doppler():
    fs, n_samp, n_ep = FS26, 260000, 24
    recs = U.synthetic_recs_fast(n_ep, 8, fs, seed=9, max_chan=9)
    code = lambda f: 1.023e6 + f * 0.0006493506493506494
    for e in range(n_ep):
        f = _step_multiple_of_2m53(fs, 2500.0 + e)                 # ch 0: every epoch is a tie epoch
        recs[e, 0]["f_carr"], recs[e, 0]["f_code"] = f, code(f)
        f = 3.0 - 0.4 * e                                          # ch 1: |f| < 10 Hz (no wrap), sign flip
        recs[e, 1]["f_carr"], recs[e, 1]["f_code"] = f, code(f)
        f = (e - 11.5) * 300.0                                     # ch 2: fast sweep through zero
        recs[e, 2]["f_carr"], recs[e, 2]["f_code"] = f, code(f)
    recs[6:9, 3]["prn"] = 0                                        # ch 3: idle gap, comes back without a reset
    recs[12, 4]["flags"] = U.E1_REC_SET_PHASE                      # ch 4: re-allocated mid-run
    recs[12, 4]["carr_phase_init"] = 0.4321
    recs[5, 5]["f_carr"] = 0.0                                     # ch 5: one epoch with zero Doppler
    recs[:, 6]["f_carr"] = -np.abs(recs[:, 6]["f_carr"]) - 50.0    # ch 6: negative Doppler throughout
    recs[:, 6]["f_code"] = code(recs[:, 6]["f_carr"])
    bad, (serial, hat, active) = U.hostsim_plan_compare(fs, n_samp, recs, carr_phase=np.linspace(-0.9, 0.9, 9))
    assert bad == 0
    assert serial >= n_ep - 1          # at least the tie channel went through the serial path
    # and the synthesised bytes agree with the oracle on a short prefix
    a, pa = U.oracle_synth(fs, 26000, recs[:14], np.linspace(-0.9, 0.9, 9))
    b, pb, st = U.hostsim_synth(fs, 26000, recs[:14], np.linspace(-0.9, 0.9, 9))
    assert np.array_equal(a, b) and np.array_equal(pa, pb)


@pytest.mark.parametrize("round_spans", [512, 64, 7])
def test_chain_kernel_rounds_equal_the_serial_chain(round_spans):
    """e1_v2_chain_kernel's algorithm (rounds of spans: optimistic scan of T = the true value after the most
    recent wrap, D = T - guess for anchors any number of spans back, prefix commit, one serial step at the
    first span that fails) transcribed for the host: the translations and the carried phase equal the
    serial chain's on ordinary, low-Doppler (anchors up to 64 spans back, spans without a wrap in reach),
    sign-changing and tie inputs, whatever the round length."""
    fs = FS26
    cases = [(low_doppler_recs(300, fs), None),
             (U.synthetic_recs_fast(200, 12, fs, seed=5), None),
             (U.synthetic_recs_fast(300, 12, fs, seed=6, f_max=80.0), np.linspace(-0.9, 0.9, 12))]
    recs = U.synthetic_recs_fast(24, 8, fs, seed=9, max_chan=9)
    for e in range(24):
        f = _step_multiple_of_2m53(fs, 2500.0 + e)
        recs[e, 0]["f_carr"], recs[e, 0]["f_code"] = f, 1.023e6 + f * 0.0006493506493506494
        f = (e - 11.5) * 300.0
        recs[e, 2]["f_carr"], recs[e, 2]["f_code"] = f, 1.023e6 + f * 0.0006493506493506494
    recs[6:9, 3]["prn"] = 0
    recs[12, 4]["flags"] = U.E1_REC_SET_PHASE
    recs[12, 4]["carr_phase_init"] = 0.4321
    cases.append((recs, np.linspace(-0.9, 0.9, 9)))
    scan = serial = 0
    for r, ph in cases:
        bad, (a, b) = U.hostsim_chain_scan_compare(fs, 260000, r, ph, round_spans)
        assert bad == 0
        scan, serial = scan + a, serial + b
    assert scan > 4 * serial
    # 25 MS/s: 306 tiles per block, spans of 39 tiles, slow channels beside ordinary ones
    r25 = U.synthetic_recs_fast(40, 10, FS25, seed=8)
    e = np.arange(40)
    for c, (f0, rate) in enumerate([(30, -0.01), (-5, 0.002), (2.5, 0.0), (100, -0.05)]):
        f = f0 + rate * e
        r25[:, c]["f_carr"] = f
        r25[:, c]["f_code"] = 1.023e6 + f * 0.0006493506493506494
    bad, _ = U.hostsim_chain_scan_compare(FS25, 2500000, r25, None, round_spans)
    assert bad == 0


def test_planner_is_exact_whatever_the_drift_estimates(monkeypatch):
    """The drift pass (e1_v2_drift_unit: the first 1/16 of each span walked exactly, its rounding drift
    scaled to the span) only feeds the guesses; the chain validates every span exactly.  With a constant
    Doppler the drift is systematic and grows linearly: with the estimates the chain accepts nearly
    every span, without them (HS_NO_DRIFT: estimates = the ideal line) it walks more of them serially --
    and the checkpoints equal the serial walk bit for bit either way."""
    fs, n_ep = FS26, 1500
    recs = U.synthetic_recs_fast(n_ep, 6, fs, seed=3)
    for c in range(6):
        recs[:, c]["f_carr"] = recs[0, c]["f_carr"]
        recs[:, c]["f_code"] = recs[0, c]["f_code"]
    bad, (serial, hat, active) = U.hostsim_plan_compare(fs, 260000, recs)
    assert bad == 0 and serial <= 6
    monkeypatch.setenv("HS_NO_DRIFT", "1")
    bad2, (serial2, hat2, active2) = U.hostsim_plan_compare(fs, 260000, recs)
    assert bad2 == 0 and active2 == active
    assert serial2 > serial + 20 and serial2 < 0.02 * active


def _first_hit_brute(a, d, M, L, n):
    x = a
    for j in range(n):
        if x < L:
            return j
        x = (x + d) % M
    return -1


def test_first_hit_equals_brute_force():
    """e1_first_hit (the Euclid-like descent behind the tile-level ambiguity test) against the literal
    loop: small moduli exhaustively-ish, the product's 2^40 modulus with steps near 0, near M, near
    M/k and random, zones from one unit to M/16."""
    import random
    hs = U.hostsim()
    rnd = random.Random(5)
    for trial in range(6000):
        kind = trial % 4
        if kind == 0:
            M = rnd.randrange(2, 120); L = rnd.randrange(1, M + 1); n = rnd.randrange(0, 50)
        elif kind == 1:
            M = 1 << 40; L = rnd.randrange(1, 1 << rnd.randrange(1, 37)); n = rnd.randrange(1, 8193)
        elif kind == 2:
            M = rnd.randrange(2, 1 << 40); L = rnd.randrange(1, min(M, 1 << 30) + 1); n = rnd.randrange(1, 8193)
        else:
            M = 1 << 40; L = rnd.randrange(1 << 20, 1 << 28); n = 8192
        a = rnd.randrange(M)
        d = [rnd.randrange(M), rnd.randrange(min(M, 1000)), M - 1 - rnd.randrange(min(M, 1000)),
             (M // rnd.randrange(1, 50) + rnd.randrange(-3, 4)) % M,
             (M * rnd.randrange(1, 30) // rnd.randrange(30, 60) + rnd.randrange(-2, 3)) % M][rnd.randrange(5)]
        assert hs.hs_first_hit(a, d, M, L, n) == _first_hit_brute(a, d, M, L, n), (a, d, M, L, n)


def test_tiles_marked_clean_never_hold_a_flagged_run():
    """The carry-walked kernel drops the per-sample ambiguity tracking for the channels of a tile that
    e1_par_clean marks.  The host build runs the tracking loop beside the plain one on every run of
    such a tile: it must flag nothing and add the same terms (hostsim_synth asserts the counter);
    here, that the marking is neither vacuous nor universal, and that flagged runs do occur (in the
    unmarked tiles) on the same input."""
    hs = U.hostsim()
    c0, k0 = hs.hs_clean_tiles(), hs.hs_checked_tiles()
    slow = 0
    for fs, n_samp, seed, kw in ((FS26, 260000, 41, {}), (FS26, 260000, 42, {"f_max": 60.0}), (FS25, 250000, 43, {})):
        recs = U.synthetic_recs(3, 36, fs, seed=seed, max_chan=36, **kw)
        a, pa = U.oracle_synth(fs, n_samp, recs, threads=8)
        b, pb, st = U.hostsim_synth(fs, n_samp, recs)
        assert np.array_equal(a, b) and np.array_equal(pa, pb)
        slow += int(st[2])
    clean, checked = hs.hs_clean_tiles() - c0, hs.hs_checked_tiles() - k0
    assert hs.hs_clean_violations() == 0
    # one code sequence per tile: the fraction stepped after the tile's code wrap stays within thr_code of the
    # one stepped before it (what e1_par_clean widens its zone by)
    assert 0.0 < hs.hs_max_code_ab_units() <= 8192 / 512.0 * 1.001 + 3.0
    assert clean > 0 and checked > 0 and slow > 0
    assert 0.9 < clean / (clean + checked) < 0.995, (clean, checked)


def test_records_with_carrier_fields_outside_the_contract_are_rejected():
    """A carr_phase_init outside (-1, 1) or a carrier step of a cycle or more per sample (or NaN) would be
    walked with ONE wrap per step where the reference does `phi -= (long)phi`: the planner rejects the record
    (E1_CK_ERROR -> E1B200_EINVAL) instead of producing other samples than the reference."""
    fs, n = U.fs_as_reference(2.6e6), 26000
    good = U.synthetic_recs(2, 3, fs, seed=4)
    U.hostsim_synth(fs, n, good)                                   # in contract: fine
    for field, val, e in (("carr_phase_init", 1.5, 0), ("carr_phase_init", -1.0, 0), ("carr_phase_init", float("nan"), 0),
                          ("f_carr", 1.5 * fs, 1), ("f_carr", -fs, 1), ("f_carr", float("nan"), 0)):
        bad = good.copy()
        bad[e, 1][field] = val
        if field == "carr_phase_init":
            bad[e, 1]["flags"] = U.E1_REC_SET_PHASE
        with pytest.raises(AssertionError, match="planner errors"):
            U.hostsim_synth(fs, n, bad)
    ok = good.copy()                                               # an out-of-range init without the flag is never read
    ok[1, 1]["carr_phase_init"] = 7.0
    U.hostsim_synth(fs, n, ok)


# ------------------------------------------------------------------ event-driven runs (fs >= 10 MS/s)
def ev_edge_recs(fs, n_epochs=4):
    """Everything the event-driven form has to get right in one input: Dopplers of both signs against phases of both
    signs (all four table walks), zero and near-zero Doppler (no carrier event in a thread's samples), a Doppler that
    changes sign (the phase runs through zero: generic form), a step of one table entry per sample and beyond (not
    E1_PAR_SLOW: generic form), a phase reset, an idle slot, the code wrap inside a thread's samples (every block)."""
    freqs = [3900.0, -3900.0, 2500.0, -2500.0, 0.0, 0.37, -12.0, 900.0, -900.0, fs / 511.0 * 0.98, -fs / 511.0 * 1.02, 11000.0]
    recs = U.synthetic_recs(n_epochs, len(freqs), fs, seed=77, max_chan=len(freqs) + 1)
    for c, f in enumerate(freqs):
        recs[:, c]["f_carr"] = f
        recs[:, c]["f_code"] = 1.023e6 + f * 0.0006493506493506494
        recs[0, c]["carr_phase_init"] = (0.11 + 0.07 * c) * (1 if c % 3 else -1)
    for e in range(n_epochs):
        recs[e, 7]["f_carr"] = (1.5 - e) * 600.0
        recs[e, 7]["f_code"] = 1.023e6 + recs[e, 7]["f_carr"] * 0.0006493506493506494
    recs[2, 2]["flags"] = U.E1_REC_SET_PHASE
    recs[2, 2]["carr_phase_init"] = -0.9999999
    recs[1:3, 3]["prn"] = 0
    return recs


@pytest.mark.parametrize("fs0,n_samp", [(25e6, 250000), (10.2e6, 102000), (16e6, 70001), (50e6, 200000)])
def test_event_driven_form_matches_oracle(fs0, n_samp):
    """e1_ev_run64 / e1_ev_sub / e1_ev_rest_impl through the host build, as e1_synth_ev_kernel drives them: bit-identical to
    the oracle; the lean form really is what most (thread, channel) pairs take and the out-of-line forms are exercised too."""
    hs = U.hostsim()
    fs = U.fs_as_reference(fs0)
    p0, r0 = hs.hs_ev_pairs(), hs.hs_ev_rest()
    recs = ev_edge_recs(fs)
    a, pa = U.oracle_synth(fs, n_samp, recs, threads=8)
    b, pb, st = U.hostsim_synth(fs, n_samp, recs)          # asserts: no lookup outside the tables, no clean-tile violation
    assert np.array_equal(a, b) and np.array_equal(pa, pb), (st, int((a != b).any(1).sum()))
    ev, rest = hs.hs_ev_pairs() - p0, hs.hs_ev_rest() - r0
    assert ev > rest > 0, (ev, rest)


def test_event_driven_and_per_sample_forms_write_the_same_stream():
    """The same 25 MS/s input through the event-driven kernel's functions and through the carry-walked ones (hs_set_ev_mode)."""
    hs = U.hostsim()
    fs, n_samp = FS25, 180000
    recs = U.synthetic_recs(3, 9, fs, seed=5, max_chan=10)
    try:
        hs.hs_set_ev_mode(0)
        a, pa, _ = U.hostsim_synth(fs, n_samp, recs)
        hs.hs_set_ev_mode(1)
        p0 = hs.hs_ev_pairs()
        b, pb, _ = U.hostsim_synth(fs, n_samp, recs)
        assert hs.hs_ev_pairs() > p0
    finally:
        hs.hs_set_ev_mode(-1)
    assert np.array_equal(a, b) and np.array_equal(pa, pb)
    ref, _ = U.oracle_synth(fs, n_samp, recs, threads=8)
    assert np.array_equal(a, ref)


def test_event_driven_form_degenerate_steps_fall_back():
    """Steps the event form does not take: a code step that divides 2^32 (fs = 32 f_chip at zero Doppler: dF = 2^28 exactly,
    n0 steps land ON 2^32 and the carry test would miss it), a carrier step of a quarter entry per sample and more
    (dG >= 2^30: the first-carry remainder test needs 3 dG < 2^32), both next to channels just inside the limits.  Such
    (tile, channel) sets are not marked E1_PAR_EV and go through the generic form; the stream is the oracle's either way."""
    hs = U.hostsim()
    fs = 32.0 * 1.023e6                      # exactly representable in float: fs_as_reference leaves it alone
    assert U.fs_as_reference(fs) == fs
    f_edge = 0.25 * fs / 511.0               # dG = 2^30
    freqs = [0.0, 1e-3, f_edge * 0.999, f_edge * 1.001, -f_edge * 0.999, -f_edge * 1.3, 2000.0]
    recs = U.synthetic_recs(2, len(freqs), fs, seed=91)
    for c, f in enumerate(freqs):
        recs[:, c]["f_carr"] = f
        recs[:, c]["f_code"] = 1.023e6 + f * 0.0006493506493506494
    p0, r0 = hs.hs_ev_pairs(), hs.hs_ev_rest()
    a, pa = U.oracle_synth(fs, 150000, recs, threads=8)
    b, pb, st = U.hostsim_synth(fs, 150000, recs)
    assert np.array_equal(a, b) and np.array_equal(pa, pb), st
    ev, rest = hs.hs_ev_pairs() - p0, hs.hs_ev_rest() - r0
    # channels 0 (power-of-two code step), 3 and 5 (fast carrier) take the generic form, the other four the event form
    assert rest >= 3 * (ev // 4) * 0.9 and ev > 0, (ev, rest)


def test_event_driven_form_independent_of_ambiguity_threshold():
    """The same at 25 MS/s: with inflated bounds fewer tiles are clean and more threads are flagged, so the tracking form of
    the events, its take-back and the generic form behind it all run -- and write the bytes of the lean form."""
    hs = U.hostsim()
    recs = U.synthetic_recs(2, 8, FS25, seed=21)
    a, _ = U.oracle_synth(FS25, 120000, recs, threads=8)
    exact, rest = [], []
    for scale in (1, 300, 30000):
        r0 = hs.hs_ev_rest()
        b, _, st = U.hostsim_synth(FS25, 120000, recs, amb_scale=scale)
        assert np.array_equal(a, b), scale
        exact.append(int(st[0]))
        rest.append(hs.hs_ev_rest() - r0)
    assert exact == sorted(exact) and exact[-1] > exact[0] and rest == sorted(rest) and rest[-1] > rest[0], (exact, rest)


def test_event_driven_form_random_rates_and_dopplers():
    """Seeded sweep: 24 sample rates between 10 and 60 MS/s (as the reference would round them: (double)(float)fs), block
    lengths that are no multiple of anything, Dopplers up to +-12 kHz, a few blocks each: the oracle's bytes every time."""
    hs = U.hostsim()
    rng = np.random.default_rng(2024)
    p0 = hs.hs_ev_pairs()
    for i in range(24):
        fs = U.fs_as_reference(float(rng.uniform(10.0e6, 60.0e6)))
        n_samp = int(rng.integers(30000, 90000))
        recs = U.synthetic_recs(2, 6, fs, seed=300 + i, max_chan=7, f_max=float(rng.choice([500.0, 4000.0, 12000.0])))
        a, pa = U.oracle_synth(fs, n_samp, recs, threads=8)
        b, pb, st = U.hostsim_synth(fs, n_samp, recs)
        assert np.array_equal(a, b) and np.array_equal(pa, pb), (i, fs, n_samp, st, int((a != b).any(1).sum()))
    assert hs.hs_ev_pairs() - p0 > 20000
