"""A small self-contained Galileo E1-B receiver (numpy), test infrastructure only.

SURVEY.md section 8(f)-2: the north star asks that the generated stream be "acquirable/trackable by
GNSS-SDR with the bundled conf"; GNSS-SDR is not in this image, so this module does what its E1 chain
does, blind (it is told nothing but the sample rate and a PRN to try):

  acquire()   FFT search over code delay and Doppler against a BOC(1,1) E1-B replica
  track()     per 4 ms code period: early / prompt / late correlators, DLL on the code phase,
              Costas PLL on the carrier -> one prompt value per I/NAV symbol
  decode()    page synchronisation on the 10-symbol pattern, 30x8 de-interleaver, Viterbi decoder of
              the rate-1/2 K=7 code (G1 = 171o, G2 = 133o inverted), CRC-24Q over the 196 page bits

Nothing here shares code with the generator: replicas come from the oracle's chip tables, the decoder
and the CRC are written from the Galileo OS SIS ICD definitions the reference follows
(src/inav-msg.cpp:4-26 interleaver + sync, :53-125 encoder, :134-162 CRC, :383-399 page layout)."""
import ctypes as C

import numpy as np

import e1util as U

CODE_LEN = 4092
F_CODE = 1.023e6
F_L1 = 1575.42e6
SYNC = np.array([0, 1, 0, 1, 1, 0, 0, 0, 0, 0], np.int8)


def halfchips(prn, component=0):
    """+-1 per BOC(1,1) half-chip (8184), E1-B (component 0) or E1-C (1), from the oracle's tables."""
    buf = (C.c_short * (2 * CODE_LEN))()
    U.oracle().e1o_halfchip_table(prn, component, buf)
    return np.array(buf, np.float32)


def replica(hc, n, fs, code_phase, f_code):
    """hc sampled at fs: sample k carries half-chip floor(2 (code_phase + k f_code / fs)) mod 8184."""
    k = np.arange(n, dtype=np.float64)
    return hc[np.floor(2.0 * (code_phase + k * (f_code / fs))).astype(np.int64) % (2 * CODE_LEN)]


def cboc_replica(prn, n, fs, code_phase, f_code, alpha, beta, component=0):
    """E1-B (component 0: alpha sc_a + beta sc_b) or E1-C (1: alpha sc_a - beta sc_b) CBOC(6,1,1/11) replica sampled
    at fs (Galileo OS SIS ICD): sc_a / sc_b the BOC(1,1) / BOC(6,1) sub-carriers, negative on their first half period."""
    hc = halfchips(prn, component)
    chips = hc[1::2]                                            # +chip sits on the odd half-chip
    k = np.arange(n, dtype=np.float64)
    sub = np.floor(12.0 * (code_phase + k * (f_code / fs))).astype(np.int64) % (12 * CODE_LEN)
    tw = sub % 12
    sc_a = np.where(tw < 6, -1.0, 1.0)
    sc_b = np.where(tw & 1, 1.0, -1.0)
    return (chips[sub // 12] * (alpha * sc_a + (beta if component == 0 else -beta) * sc_b)).astype(np.float32)


def acquire(x, prn, fs, periods=4, f_max=5000.0, f_step=125.0, cboc=None, raw=False):
    """-> (metric = peak / mean of the search grid, doppler [Hz], code phase of sample 0 [chips]).
    cboc = (alpha, beta): correlate against the E1-B CBOC replica instead of BOC(1,1); raw: also return the peak power."""
    n = int(round(fs * CODE_LEN / F_CODE))                      # samples per 4 ms code period
    hc = halfchips(prn)
    rep = replica(hc, n, fs, 0.0, F_CODE) if cboc is None else cboc_replica(prn, n, fs, 0.0, F_CODE, cboc[0], cboc[1])
    L = np.conj(np.fft.fft(rep))
    t = np.arange(n * periods) / fs
    best = (0.0, 0.0, 0)
    total, cells = 0.0, 0
    for fd in np.arange(-f_max, f_max + 1.0, f_step):
        y = (x[:n * periods] * np.exp(-2j * np.pi * fd * t)).reshape(periods, n)
        p = (np.abs(np.fft.ifft(np.fft.fft(y, axis=1) * L, axis=1)) ** 2).sum(0)
        total += p.sum()
        cells += n
        i = int(p.argmax())
        if p[i] > best[0]:
            best = (float(p[i]), float(fd), i)
    peak, fd, delay = best
    # the replica's chip 0 sits at sample `delay`: sample 0 is (n - delay) samples into a code period
    code_phase = ((n - delay) % n) * F_CODE / fs
    if raw:
        return peak / (total / cells), fd, code_phase % CODE_LEN, peak
    return peak / (total / cells), fd, code_phase % CODE_LEN


def track(x, prn, fs, doppler, code_phase, n_periods, spacing=0.25, bn_pll=25.0, bn_dll=2.0, pilot=False):
    """Code/carrier tracking from the acquisition estimate.  Returns the complex prompt of every whole
    code period (= I/NAV symbol) after sample `start`, the carrier frequency and code phase histories,
    and `start` (the sample at which the first whole code period begins).
    pilot=True: the loops run on the E1-C (pilot) replica, as GNSS-SDR does with the bundled configuration
    (gnss-sdr_Galileo_E1_ishort.conf:60, track_pilot=true); a fifth return value holds the E1-B (data) prompts
    taken with the same code and carrier NCOs (see pilot_symbols)."""
    hc = halfchips(prn, 1 if pilot else 0)
    hc_data = halfchips(prn, 0) if pilot else None
    n = int(round(fs * CODE_LEN / F_CODE))
    T = n / fs
    # second-order loop coefficients (natural frequency from the noise bandwidth, damping 0.707)
    wn_p, wn_d = bn_pll / 0.53, bn_dll / 0.53
    k1_p, k2_p = 1.414 * wn_p, wn_p * wn_p * T
    fd, theta, cp = float(doppler), 0.0, float(code_phase)
    # pull-in: refine the Doppler of the 125 Hz acquisition bin from the phase slope of squared 1 ms
    # partial correlations (squaring removes the symbol sign; 1 ms spacing is unambiguous to +-250 Hz)
    k = np.arange(n)
    q = n // 4
    pr = []
    for p in range(16):
        seg = x[p * n:(p + 1) * n] * np.exp(-2j * np.pi * fd * (k + p * n) / fs)
        rep = replica(hc, n, fs, cp + p * n * F_CODE / fs, F_CODE)
        pr.extend(np.vdot(rep[i * q:(i + 1) * q], seg[i * q:(i + 1) * q]) for i in range(4))
    sq = np.array(pr) ** 2
    fd += float(np.angle(np.sum(sq[1:] * np.conj(sq[:-1])))) / (2.0 * 2.0 * np.pi * (q / fs))
    prompts = np.zeros(n_periods, np.complex128)
    data_prompts = np.zeros(n_periods, np.complex128)
    f_hist, cp_hist = np.zeros(n_periods), np.zeros(n_periods)
    f_int = fd
    # integrate over whole code periods (= symbols): start at the first sample after the next code
    # boundary and size every window so that it ends at the following one (10400 +- 1 samples)
    pos = int(np.ceil((CODE_LEN - cp) * fs / F_CODE)) % n
    cp = (cp + pos * F_CODE / fs) % CODE_LEN
    if cp > CODE_LEN / 2:
        cp -= CODE_LEN
    start = pos
    done = 0
    for p in range(n_periods):
        f_code = F_CODE * (1.0 + fd / F_L1)
        m = int(round((CODE_LEN - cp) * fs / f_code))
        seg = x[pos:pos + m]
        if len(seg) < m:
            break
        km = np.arange(m)
        base = seg * np.exp(-1j * (theta + 2.0 * np.pi * fd * km / fs))
        e = np.vdot(replica(hc, m, fs, cp + spacing, f_code), base)
        pm = np.vdot(replica(hc, m, fs, cp, f_code), base)
        l = np.vdot(replica(hc, m, fs, cp - spacing, f_code), base)
        prompts[p], f_hist[p], cp_hist[p] = pm, fd, cp
        if pilot:
            data_prompts[p] = np.vdot(replica(hc_data, m, fs, cp, f_code), base)
        done = p + 1
        # Costas discriminator (insensitive to the symbol sign), loop filter
        err = np.arctan(pm.imag / pm.real) / (2.0 * np.pi) if pm.real != 0.0 else 0.0
        theta = (theta + 2.0 * np.pi * fd * m / fs) % (2.0 * np.pi)
        f_int += k2_p * err
        fd = f_int + k1_p * err
        # DLL: normalised early-minus-late envelope; an early replica (ahead in phase) that correlates
        # better means the local code lags the signal
        ae, al = abs(e), abs(l)
        d = 0.5 * (ae - al) / (ae + al) if ae + al > 0 else 0.0
        cp = cp + m * f_code / fs - CODE_LEN + 1.414 * wn_d * T * d
        pos += m
    if pilot:
        return prompts[:done], f_hist[:done], cp_hist[:done], start, data_prompts[:done]
    return prompts[:done], f_hist[:done], cp_hist[:done], start


SEC25 = np.array([int(c) for c in "0011100000001010110110010"], np.int8)   # Galileo OS SIS ICD, E1-C secondary code CS25_1


def pilot_symbols(pilot_prompts, data_prompts, skip=75):
    """Secondary-code synchronisation on the pilot and pilot-aided data demodulation.
    The E1 composite is e_B(t) d - e_C(t) s (ICD: the pilot enters with a minus sign; src/galileo-sdr.cpp:520), so after
    correlation the pilot prompt is -s e^(j theta) and the data prompt d e^(j theta), with s = +-1 the secondary-code chip
    (bit 1 -> -1) and d the I/NAV symbol (bit 1 -> -1): the Costas loop leaves theta ambiguous by pi, the secondary code
    -- a KNOWN sequence -- removes that ambiguity, and the data symbols come out with their absolute sign.
    -> (page bits (0/1, absolute), secondary-code phase of symbol 0 (0..24), fraction of the pilot prompts after `skip`
        that agree with the aligned secondary code)."""
    ps = (pilot_prompts.real < 0).astype(np.int8)               # sign bit of -s e^(j theta), theta in {0, pi}
    n = len(ps)
    best = (-1.0, 0, 0)
    for shift in range(25):
        exp = SEC25[(np.arange(n) + shift) % 25]                # bit 1 -> s = -1 -> pilot prompt -s = +1 -> sign bit 0 (theta = 0)
        agree = float(((1 - exp)[skip:] == ps[skip:]).mean())
        for inv, a in ((0, agree), (1, 1.0 - agree)):
            if a > best[0]:
                best = (a, shift, inv)
    agree, shift, inv = best
    # wipe the secondary code off the pilot: what is left is the carrier phase, e^(j theta) without ambiguity
    s = 1.0 - 2.0 * SEC25[(np.arange(n) + shift) % 25]
    carrier = -pilot_prompts * s
    d = (data_prompts * np.conj(carrier)).real
    return (d < 0).astype(np.int8), shift, agree


def symbols_from_prompts(prompts):
    """Hard symbols with the reference's polarity convention up to a global sign: page bit 1 -> -1."""
    return (prompts.real < 0).astype(np.int8)


def find_sync(sym):
    """Offsets (mod 250) where the 10-symbol pattern (or its inverse) repeats every 250 symbols.
    -> (offset, inverted)"""
    best = (-1, 0, False)
    n_half = (len(sym) - 10) // 250
    for off in range(250):
        hits = 0
        inv = 0
        cnt = 0
        for h in range(n_half + 1):
            a = off + 250 * h
            if a + 10 > len(sym):
                break
            m = int((sym[a:a + 10] == SYNC).sum())
            hits += max(m, 10 - m)
            inv += m < 5
            cnt += 1
        if cnt and hits / cnt > best[1]:
            best = (off, hits / cnt, inv * 2 > cnt)
    return best[0], best[2], best[1]


_G1 = np.array([1, 1, 1, 1, 0, 0, 1], np.int8)      # taps on the newest bit first (171 octal)
_G2 = np.array([1, 0, 1, 1, 0, 1, 1], np.int8)      # 133 octal, output inverted


def viterbi_k7(coded):
    """Hard-decision Viterbi decoder of the I/NAV code: 240 symbols -> 120 bits (114 + 6 tail zeros),
    encoder starts and ends in state 0."""
    n = len(coded) // 2
    n_states = 64
    # state = the 6 previous bits, newest in the MSB (bit 5)
    out = np.zeros((n_states, 2, 2), np.int8)
    nxt = np.zeros((n_states, 2), np.int64)
    for s in range(n_states):
        prev = [(s >> (5 - j)) & 1 for j in range(6)]           # newest first
        for b in (0, 1):
            reg = [b] + prev
            g1 = sum(r & int(t) for r, t in zip(reg, _G1)) & 1
            g2 = 1 - (sum(r & int(t) for r, t in zip(reg, _G2)) & 1)
            out[s, b] = (g1, g2)
            nxt[s, b] = (b << 5) | (s >> 1)
    INF = 10 ** 9
    metric = np.full(n_states, INF, np.int64)
    metric[0] = 0
    back = np.zeros((n, n_states), np.int8)
    prev_state = np.zeros((n, n_states), np.int64)
    for t in range(n):
        r = coded[2 * t:2 * t + 2]
        new = np.full(n_states, INF, np.int64)
        for s in range(n_states):
            if metric[s] >= INF:
                continue
            for b in (0, 1):
                m = metric[s] + int(out[s, b, 0] != r[0]) + int(out[s, b, 1] != r[1])
                ns = nxt[s, b]
                if m < new[ns]:
                    new[ns] = m
                    back[t, ns] = b
                    prev_state[t, ns] = s
        metric = new
    s = 0                                                       # tail bits drive the encoder back to zero
    bits = np.zeros(n, np.int8)
    for t in range(n - 1, -1, -1):
        bits[t] = back[t, s]
        s = prev_state[t, s]
    return bits, int(metric[0])


def crc24q(bits):
    crc = 0
    for b in bits:
        crc ^= int(b) << 23
        crc <<= 1
        if crc & 0x1000000:
            crc ^= 0x1864CFB
    return crc & 0xFFFFFF


def decode_half(sym250):
    """250 symbols (sync + 240) -> 114 page bits and the number of channel errors Viterbi corrected."""
    inter = sym250[10:250]
    coded = np.zeros(240, np.int8)
    for r in range(8):
        for c in range(30):
            coded[c * 8 + r] = inter[r * 30 + c]
    bits, errs = viterbi_k7(coded)
    return bits[:114], errs, bits[114:]


def decode_pages(sym):
    """-> list of dicts {start (symbol index of the even half), word_type, crc_ok, bits (228)}."""
    off, inverted, quality = find_sync(sym)
    if inverted:
        sym = 1 - sym
    halves = []
    a = off
    while a + 250 <= len(sym):
        bits, errs, tail = decode_half(sym[a:a + 250])
        halves.append((a, bits, errs, tail))
        a += 250
    pages = []
    for (a0, even, e0, t0), (a1, odd, e1, t1) in zip(halves[:-1], halves[1:]):
        if even[0] != 0 or odd[0] != 1:                         # even/odd flag is the first bit of a half page
            continue
        page = np.concatenate([even, odd])
        crc_rx = int("".join(map(str, page[196:220])), 2)
        pages.append(dict(start=a0, word_type=int("".join(map(str, even[2:8])), 2), crc_ok=crc24q(page[:196]) == crc_rx,
                          channel_errors=e0 + e1, tails_zero=not (t0.any() or t1.any()), bits=page))
    return pages, dict(sync_offset=off, inverted=inverted, sync_quality=quality)
