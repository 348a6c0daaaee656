/* e1_core.h -- exact arithmetic shared by the planner and the synthesis kernels.
 *
 * Everything here is `host device`: the CUDA kernels (e1_kernels.cu) are the product; the
 * same functions are also compiled for the host by tests/hostsim (g++ -ffp-contract=off) so
 * the exactness logic can be checked against the oracle in the GPU-less build container.
 * That host build is test scaffolding: the C-ABI library has no CPU path.
 *
 * The reference integrates two serial double-precision recurrences per channel
 * (src/galileo-sdr.cpp:491-494, 528-532):
 *      code_phase = fl(code_phase + fl(f_code*delt));   wrap by -4092 before the next use
 *      carr_phase = fl(carr_phase + fl(f_carr*delt));   carr_phase -= (long)carr_phase
 * and turns them into table indices by truncation (:509-515).  A 1-ulp difference flips an
 * index, so a parallel version has to reproduce the serial roundings, not approximate them.
 * Two building blocks do that:
 *
 *  1. e1_walk_*: "binade jumps".  While x stays inside one binade [2^e, 2^(e+1)) it is a
 *     multiple of that binade's ulp, so fl(x+s) = x + D with a constant integer D (in ulps),
 *     i.e. the raw IEEE bit pattern advances linearly.  D is measured with two real additions
 *     (which also settles round-half-even parity), the rest of the binade is one integer
 *     multiply, and the step that crosses a binade edge / wrap threshold is a real addition
 *     again.  Cost: O(binades crossed), result: bit-identical to the serial loop.
 *
 *  2. e1_fast_* / e1_exact_*: inside a tile of T samples that starts from an exact
 *     checkpoint, sample j is evaluated in closed form with 64-bit fixed point.  The closed
 *     form differs from the serial value by at most j half-ulps; whenever the scaled phase is
 *     within that bound of an integer (where truncation could disagree) the sample is flagged
 *     and re-evaluated by literally stepping the reference recurrence from the checkpoint.
 */
#ifndef E1_CORE_H
#define E1_CORE_H

#include <stddef.h>
#include <stdint.h>
#include <string.h>

#include "../../include/e1b200.h"

#if defined(__CUDACC__)
#define E1_HD __host__ __device__ __forceinline__
#else
#define E1_HD static inline
#endif

#define E1C_CODE_LEN 4092
#define E1C_SYM_PER_PAGE 500

/* ------------------------------------------------------------------ IEEE helpers */
E1_HD double e1_add(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dadd_rn(a, b); /* never contracted into an FMA */
#else
    return a + b; /* host builds use -ffp-contract=off */
#endif
}
E1_HD double e1_mul(double a, double b)
{
#if defined(__CUDA_ARCH__)
    return __dmul_rn(a, b);
#else
    return a * b;
#endif
}
E1_HD int64_t e1_bits(double x)
{
#if defined(__CUDA_ARCH__)
    return __double_as_longlong(x);
#else
    int64_t b;
    memcpy(&b, &x, 8);
    return b;
#endif
}
E1_HD double e1_from_bits(int64_t b)
{
#if defined(__CUDA_ARCH__)
    return __longlong_as_double(b);
#else
    double x;
    memcpy(&x, &b, 8);
    return x;
#endif
}
E1_HD double e1_fabs(double x) { return e1_from_bits(e1_bits(x) & 0x7fffffffffffffffLL); }

/* floor(num / d) for 0 <= num < 2^53, 1 <= d < 2^53 (mantissa distances inside one binade).  On the
 * device a 64-bit integer division is a ~100-instruction subroutine and was half of the planner's
 * instructions; both operands are exact as doubles, and a quotient rounded DOWN has the same integer
 * part as the true one (every integer below 2^53 is representable, rounding down is monotone). */
E1_HD int64_t e1_div_binade(int64_t num, int64_t d)
{
#if defined(__CUDA_ARCH__)
    return __double2ll_rz(__ddiv_rd(__ll2double_rn(num), __ll2double_rn(d)));
#else
    return num / d;
#endif
}

/* One literal reference step of the carrier recurrence (src/galileo-sdr.cpp:531-532). */
E1_HD double e1_carr_step(double phi, double sp)
{
    phi = e1_add(phi, sp);
    return e1_add(phi, -(double)(long long)phi); /* exact: |phi| < 2 */
}
/* One literal reference step of the code recurrence: add (:528), then the wrap test the
 * reference performs at the top of the next sample (:491-494).  *wrapped is set on a wrap. */
E1_HD double e1_code_step(double cp, double sc, int *wrapped)
{
    cp = e1_add(cp, sc);
    if (cp >= (double)E1C_CODE_LEN) {
        cp = e1_add(cp, -(double)E1C_CODE_LEN);
        *wrapped = 1;
    }
    return cp;
}

/* ------------------------------------------------------------------ binade-jump walkers
 *
 * e1_walk_up: x >= 0 grows by s >= 0 per sample; when the sum reaches `limit` the caller's
 * wrap rule applies (limit = 1.0 for |carrier phase|, 4092.0 for the code phase).  Advances
 * from sample *k towards k_target and returns early (with *wrapped = 1) right after a wrap,
 * so the caller can record where it happened.  On return *k is the sample whose value is the
 * returned x.  `limit` must be a finite positive double. */
E1_HD double e1_walk_up(double x, double s, double limit, int64_t *k, int64_t k_target, int *wrapped)
{
    int64_t kk = *k;
    const int64_t lim_bits = e1_bits(limit);
    *wrapped = 0;
    while (kk < k_target) {
        /* real step */
        double x1 = e1_add(x, s);
        kk++;
        if (x1 >= limit) {
            x = e1_add(x1, -limit); /* exact (Sterbenz) for both users of this function */
            *wrapped = 1;
            break;
        }
        x = x1;
        if (kk + 1 >= k_target)
            continue;
        /* try to jump through the rest of x1's binade.  x1 may have entered the binade with an
           odd mantissa; if s is an exact half-ulp tie here, round-half-even makes x1->x2 differ
           from all later steps, so the constant step is measured from x2->x3. */
        double x2 = e1_add(x1, s), x3 = e1_add(x2, s);
        int64_t b1 = e1_bits(x1), b2 = e1_bits(x2), b3 = e1_bits(x3);
        if ((((b1 ^ b2) | (b1 ^ b3)) >> 52) != 0 || x3 >= limit)
            continue; /* a binade edge or the wrap is at most two real steps away */
        int64_t d = b3 - b2;
        if (d == 0) { /* s below half an ulp: the value is stuck for good */
            x = x2;
            kk = k_target;
            break;
        }
        /* first bit pattern that is NOT reachable by plain integer stepping */
        int64_t end = ((b1 >> 52) + 1) << 52;
        if (lim_bits < end)
            end = lim_bits;
        kk++;                           /* now at x2 */
        int64_t n = e1_div_binade(end - 1 - b2, d); /* >= 1 because x3 qualified */
        int64_t room = k_target - kk;
        if (n > room)
            n = room;
        x = e1_from_bits(b2 + n * d);
        kk += n;
    }
    *k = kk;
    return x;
}

/* e1_walk_down: x > 0 shrinks by s > 0 per sample (carrier phase and Doppler of opposite
 * sign).  Stops early with *crossed = 1 right after the step whose result is <= 0; the
 * returned value is then that (non-positive) result.  */
E1_HD double e1_walk_down(double x, double s, int64_t *k, int64_t k_target, int *crossed)
{
    int64_t kk = *k;
    *crossed = 0;
    while (kk < k_target) {
        double x1 = e1_add(x, -s);
        kk++;
        x = x1;
        if (!(x1 > 0.0)) {
            *crossed = 1;
            break;
        }
        if (kk >= k_target)
            break;
        if (kk + 1 >= k_target)
            continue;
        double x2 = e1_add(x1, -s), x3 = e1_add(x2, -s);
        int64_t b1 = e1_bits(x1), b2 = e1_bits(x2), b3 = e1_bits(x3);
        if (!(x3 > 0.0) || (((b1 ^ b2) | (b1 ^ b3)) >> 52) != 0)
            continue;
        int64_t d = b2 - b3;
        if (d == 0) {
            x = x2;
            kk = k_target;
            break;
        }
        int64_t floor_bits = (b1 >> 52) << 52; /* 2^e: still inside the binade */
        kk++;                                  /* now at x2 */
        int64_t n = e1_div_binade(b2 - floor_bits, d); /* >= 1 */
        int64_t room = k_target - kk;
        if (n > room)
            n = room;
        x = e1_from_bits(b2 - n * d);
        kk += n;
    }
    *k = kk;
    return x;
}

/* Carrier phase after advancing from sample k0 to k1 with step sp (any signs), bit-identical
 * to applying e1_carr_step (k1-k0) times.  FP addition is sign-symmetric, so the walk is done
 * on the magnitude. */
E1_HD double e1_carr_advance(double phi, double sp, int64_t k0, int64_t k1)
{
    if (sp == 0.0 || k1 <= k0)
        return phi;
    int neg = (phi < 0.0) || (phi == 0.0 && sp < 0.0);
    double a = e1_fabs(phi);
    const double t = e1_fabs(sp);
    int64_t k = k0;
    while (k < k1) {
        int aligned = (a == 0.0) || (neg == (sp < 0.0));
        int ev;
        if (aligned) {
            neg = (sp < 0.0);
            a = e1_walk_up(a, t, 1.0, &k, k1, &ev);
        } else {
            a = e1_walk_down(a, t, &k, k1, &ev);
            if (ev) { /* went through zero: magnitude continues on the other side */
                a = -a;
                neg = !neg;
            }
        }
    }
    return neg ? -a : a;
}

/* ------------------------------------------------------------------ fixed-point closed form */

/* |x| * 2^sh truncated to 64 bits, for 0 <= |x| * 2^sh < 2^64. */
E1_HD uint64_t e1_to_fixed(double x, int sh)
{
    int64_t b = e1_bits(x) & 0x7fffffffffffffffLL;
    int e = (int)(b >> 52);
    uint64_t m = (uint64_t)(b & 0xfffffffffffffLL);
    if (e == 0)
        return 0; /* denormals are far below one unit */
    m |= 1ULL << 52;
    int shift = e - 1075 + sh; /* value = m * 2^(e-1075) */
    if (shift >= 0)
        return shift > 11 ? ~0ULL : (m << shift);
    shift = -shift;
    return shift > 63 ? 0 : (m >> shift);
}

#define E1_CARR_FIX 64 /* carrier phase magnitude in units of 2^-64 cycle           */
#define E1_CODE_FIX 52 /* code phase in units of 2^-52 chip = 2^-51 half-chip, < 2^64 */

/* Ambiguity thresholds, in units of 2^-32 of the truncated quantity, for a tile of T samples.
 *  carrier: |phi_serial - phi_closed| <= T*(2^-53 + 2^-64) + 2^-64 cycles; y = 511*phi, and the
 *           reference truncates fl(511*phi), which can round up across an integer from at most
 *           2^-45 below -> thr = 511*T*2^-21 (1 + 2^-10) + 2.
 *  code   : per step error <= 2^-42 chip (values < 4096) plus 2^-53 from the quantised step;
 *           h = 2*cp -> thr = T*2^-9 (1 + 2^-10) + 2.                                        */
E1_HD uint32_t e1_thr_carr(int T, int scale)
{
    double v = 511.0 * (double)T / 2097152.0 * 1.001 + 2.0;
    v *= (double)scale;
    return v > 1.0e9 ? 1000000000u : (uint32_t)v + 1u;
}
E1_HD uint32_t e1_thr_code(int T, int scale)
{
    double v = (double)T / 512.0 * 1.001 + 2.0;
    v *= (double)scale;
    return v > 1.0e9 ? 1000000000u : (uint32_t)v + 1u;
}

/* ------------------------------------------------------------------ records shared by planner and synthesis */
#define E1C_N_PRN 50
/* Code table: per PRN, one 2-bit field per BOC(1,1) half-chip hh = 0..8183 (chip c = hh/2), sixteen
 * half-chips per 32-bit word, FIRST half-chip in the TOP bits: field of hh at bits 31-2k, 30-2k of
 * word hh/16, k = hh & 15.  With bneg / cneg = 1 where the E1-B / E1-C half-chip value is -1
 * (src/gal-sig.cpp:9-233: chip = 1 - 2*bit, sboc negates every even half-chip):
 *   high bit = bneg          low bit = bneg XOR cneg
 * XORing a word with the symbol pattern (D on the high bits, D^S on the low bits; D = nav symbol,
 * S = secondary-code bit, src/galileo-sdr.cpp:517-518) turns the fields into (x, x^y) with
 * x / y = 1 where B*d / C*s is -1, and `w & ((w << 1) | 0x55555555)` turns that into the 2-bit
 * two's-complement value of  y - x = (B*d - C*s)/2  in {-1, 0, +1}  (:520-521).
 * 512 data words plus zero padding (a window read touches word i and i+1).                       */
#define E1C_CODE_WORDS_PER_PRN 516
/* Carrier table: int32 [E1C_LUT_IDX][E1C_LUT_REP]; entry E = e + E1C_LUT_EXT holds, in every copy l,
 * 2*(cos + 65536*sin) of the reference's table index t(e) (include/constants.h:216-284):
 *     e in [0, 512]          t = e & 511
 *     e in [-EXT, -1]        t = 511 + e        (below: where a run that will wrap starts, phase >= 0)
 *     e in [513, 513 + EXT]  t = e - 511        (above: the same for phase < 0)
 * With i = trunc(511 |phi|) in [0, 510] the reference reads index i for phi >= 0 and (-i) & 511 for
 * phi < 0 (src/galileo-sdr.cpp:509-510).  The sample loop multiplies an UNWRAPPED magnitude across a
 * carrier wrap (phi -= (long)phi, :532): 511 (|phi| + 1) has the same fraction and an index exactly
 * 511 higher.  So phase >= 0 reads e = i, or e = i - 511 from the start when the run is about to
 * wrap (i - 511 < 0 before the wrap, the true index after it); phase < 0 reads e = 512 - i, or
 * e = 512 - i + 511 when about to wrap (1 + m, m = 511 - i, before the wrap -- which is (-i) & 511 --
 * and 512 - i' after it).  One table serves both signs, which leaves room for THIRTY-TWO copies of
 * every entry, one per lane: a warp's 32 lookups fall into 32 distinct banks whatever the indices
 * are (with 16 copies lanes l and l+16 collided on almost every load).                           */
#define E1C_LUT_REP 32
#define E1C_LUT_EXT 64
/* 513 + 2 EXT positions are reachable by unambiguous samples; one more at the top because an AMBIGUOUS
 * sample of a mirrored run reads one entry too high (see e1_carrier_start): the run is redone, but the
 * fast form's terms are taken back out by recomputing them, so that read has to be inside the table
 * (what lies behind it in shared memory -- a parameter buffer -- can change between the two reads). */
#define E1C_LUT_IDX (514 + 2 * E1C_LUT_EXT)
#define E1C_LUT_ENTRIES (E1C_LUT_IDX * E1C_LUT_REP)
#define E1C_THREADS 512 /* synthesis CTA: thread t owns samples [t*R, (t+1)*R) of the tile */
#define E1C_MAX_RUN 16
/* fast path: |carrier step| such that E1C_MAX_RUN samples move the table index by < E1C_LUT_EXT */
#define E1C_FAST_SP_MAX ((E1C_LUT_EXT - 1.0) / (511.0 * E1C_MAX_RUN))
/* GALILEO_E1_SECONDARY_CODE (include/constants.h:213), bit i = symbol i */
#define E1C_SEC25_MASK 0x009B501Cu
#define E1C_NO_WRAP 0x7fffffff

typedef struct e1_tile_ck { /* 32 bytes, one per (epoch, tile, channel) */
    double phi;   /* carrier phase at the tile's first sample (exact)                     */
    double cp;    /* code phase at the tile's first sample, after the wrap test (exact)  */
    double cp_w;  /* code phase at sample j_w (exact)                                     */
    int32_t j_w;  /* tile-relative first sample after the code wrap, E1C_NO_WRAP if none */
    uint32_t sym; /* bit0 nav symbol, bit1 secondary-code bit before the wrap; bits 2,3
                     after it; bit 4 = channel active; bit 5 = planner error             */
} e1_tile_ck;
#define E1_CK_ACTIVE 16u
#define E1_CK_ERROR 32u

typedef struct e1_chan_par { /* 112 bytes, one per active channel of a tile (HBM -> shared memory by bulk copy) */
    uint64_t U0, dU;        /* |carrier phase| and its per-sample step, 2^-64 cycle                    */
    uint64_t HA, HB;        /* code phase at sample 0 / extrapolated back from j_w to sample 0,
                               2^-51 half-chip, plus the fast path's bias (e1_bias_h)                 */
    uint64_t dH;            /* code phase step per sample, 2^-51 half-chip                            */
    uint32_t dF;            /* dH >> 19: step of the fast path's 32-bit half-chip fraction            */
    int32_t j_w;
    uint32_t pat_a, pat_b;  /* symbol XOR pattern for the code words before / after the code wrap;
                               FLOAT path (E1B200_CFG_CBOC / _GAIN): pat_a = the channel's linear gain as
                               IEEE float bits, pat_b unused (the symbols are in misc)                */
    uint32_t code_off;      /* word offset of this PRN's code words                                   */
    uint32_t misc;          /* bits 0-1 symbol field (D<<1 | D^S) before the wrap, bits 2-3 after it,
                               bit 4 negative-phase regime, bit 5 force the generic path, bit 6: the
                               phase runs through zero at tile sample j_z = bits 16-30 (from there on
                               the magnitude is the two's complement of U and the regime flips), bit 7:
                               no sample of the tile is near an index boundary (e1_par_clean)          */
    double phi, sp, cp, sc; /* exact checkpoint for the exact fallback                                */
    uint32_t Dlo;           /* E1_PAR_SLOW: low word of the table position's step per sample, 2^-32 entry:
                               floor(511 dU / 2^32) of the SIGNED step, negated for phase < 0 (the mirrored walk);
                               its sign is E1_PAR_DOWN                                                 */
    uint32_t ev_n0;         /* E1_PAR_EV: samples between two carries of the 32-bit fraction accumulators, rounded down:
                               floor(2^32 / dF) in bits 0-15 (code), floor(2^32 / dG) capped at 0xffff in bits 16-31
                               (carrier; dG = Dlo walking up, 2^32 - Dlo walking down)                                 */
    uint32_t ev_rcp_f;      /* 1.0f / (float)dF as IEEE bits (first carry of a run: e1_ev_first)                       */
    uint32_t ev_rcp_g;      /* 1.0f / (float)dG, +inf for dG = 0                                                        */
} e1_chan_par;
#define E1_PAR_NEG 16u
#define E1_PAR_FORCE 32u
#define E1_PAR_HASZ 64u
#define E1_PAR_CLEAN 128u /* no run of this tile can be ambiguous (e1_par_clean): the sample loop skips its tracking */
#define E1_PAR_SLOW 256u  /* the carrier moves by less than one table entry per sample (|step| < 1/512 cycle: 5 kHz at
                             2.6 MS/s) and the tile is regular (no FORCE / HASZ): the carry-walked kernel walks the table
                             by CARRIES of the index fraction, one carrier start per 32 or 64 samples (e1_run_cw)      */
#define E1_PAR_DOWN 512u  /* with E1_PAR_SLOW: the table position decreases from sample to sample                      */
#define E1_PAR_EV 1024u   /* with E1_PAR_SLOW, set when the context runs the event-driven kernel (E1_INT_EV): a half-chip
                             lasts several samples, a thread's E1C_EV_RUN samples fit one 16-field code window, and the
                             ev_* fields are filled in (e1_ev_add)                                                      */
/* internal bit of the cfg_flags argument of e1_make_par (never in e1b200_config.flags): fill in the event fields */
#define E1_INT_EV 0x40000000u
#if !defined(E1C_EV_RUN)
#define E1C_EV_RUN 64 /* consecutive samples per thread of the event-driven kernel, walked from ONE carrier start */
#endif
#define E1C_EV_EXT E1C_EV_RUN              /* ... so its (single-copy) carrier tables run out that many entries on either side */
#define E1C_LUT1_IDX (514 + 2 * E1C_EV_EXT) /* entries of one of them (cf. E1C_LUT_IDX) */

/* One tile's parameter block in HBM: header (16 bytes: n_active, 3 x pad) + max_chan e1_chan_par,
 * active channels first. */
#define E1C_BLK_HEADER 16
E1_HD size_t e1_blk_bytes(int max_chan) { return E1C_BLK_HEADER + (size_t)max_chan * sizeof(e1_chan_par); }

E1_HD uint32_t e1_umulhi(uint32_t a, uint32_t b)
{
#if defined(__CUDA_ARCH__)
    return __umulhi(a, b);
#else
    return (uint32_t)(((uint64_t)a * b) >> 32);
#endif
}
E1_HD int e1_d2i_rz(double x)
{
#if defined(__CUDA_ARCH__)
    return __double2int_rz(x);
#else
    return (int)x;
#endif
}

E1_HD uint32_t e1_sym_bits(const e1_epoch_rec *r, int ibit, int page_sel)
{
    const uint8_t *pg = page_sel ? r->page_next : r->page_cur;
    uint32_t d = (uint32_t)(pg[ibit >> 3] >> (ibit & 7)) & 1u;      /* src/galileo-sdr.cpp:517 */
    uint32_t s = (E1C_SEC25_MASK >> (ibit % E1_SEC_CODE_LEN)) & 1u; /* :518 */
    return d | (s << 1);
}

/* Carrier fields of a record inside the contract (include/e1b200.h): the walkers and the closed form take
 * ONE wrap per step (x1 - 1.0), which equals the reference's `phi -= (long)phi` (:532) only while the
 * phase stays inside (-1, 1) and the step is below one cycle per sample; anything else (NaN included) is
 * rejected like a bad code phase, instead of silently giving other samples than the reference would. */
E1_HD int e1_rec_carrier_ok(const e1_epoch_rec *r, double delt)
{
    if ((r->flags & E1_REC_SET_PHASE) && !(e1_fabs(r->carr_phase_init) < 1.0))
        return 0;
    return e1_fabs(e1_mul(r->f_carr, delt)) < 1.0;
}

/* Code-phase plan of one (epoch, channel): the code phase restarts every epoch
 * (computeCodePhase, src/gal-sig.cpp:308-347), so epochs are independent.  Walks the
 * recurrence (:491-507, :528) exactly and leaves one checkpoint per tile in o[t*stride]. */
E1_HD void e1_plan_code_epoch(const e1_epoch_rec *r, e1_tile_ck *o, int stride, int n_samp, int tile,
                              int tiles_per_epoch, double delt)
{
    const int prn = r->prn;
    if (prn < 1 || prn > E1C_N_PRN) {
        for (int t = 0; t < tiles_per_epoch; t++)
            o[(size_t)t * stride].sym = 0;
        return;
    }
    const double sc = e1_mul(r->f_code, delt);
    double cp = r->code_phase0;
    int ibit = r->ibit0, page_sel = 0;
    uint32_t err = 0;
    if (!(cp >= 0.0) || !(sc > 0.0) || !(sc < 2048.0) || ibit < 0 || ibit >= E1C_SYM_PER_PAGE || !e1_rec_carrier_ok(r, delt)) {
        err = E1_CK_ERROR;
        ibit = 0;
    }
    if (cp >= (double)E1C_CODE_LEN) { /* the reference's wrap test at sample 0 (:491) */
        cp = e1_add(cp, -(double)E1C_CODE_LEN);
        if (++ibit >= E1C_SYM_PER_PAGE) {
            ibit = 0;
            page_sel = 1;
        }
        if (cp >= (double)E1C_CODE_LEN)
            err = E1_CK_ERROR;
    }
    for (int t = 0; t < tiles_per_epoch; t++) {
        int64_t k0 = (int64_t)t * tile, k1 = k0 + tile;
        if (k1 > n_samp)
            k1 = n_samp;
        const double cp_t = cp;
        double cp_w = 0.0;
        int32_t j_w = E1C_NO_WRAP;
        uint32_t sym_a = e1_sym_bits(r, ibit, page_sel), sym_b = sym_a;
        int64_t k = k0;
        while (k < k1 && !err) {
            int wrapped;
            cp = e1_walk_up(cp, sc, (double)E1C_CODE_LEN, &k, k1, &wrapped);
            if (wrapped) {
                if (++ibit >= E1C_SYM_PER_PAGE) { /* :495-506: generateINavMsg -> page_next */
                    ibit = 0;
                    page_sel = 1;
                }
                if (k < k1) {
                    if (j_w != E1C_NO_WRAP)
                        err = E1_CK_ERROR; /* two code wraps in one tile: tile too long for fs */
                    j_w = (int32_t)(k - k0);
                    cp_w = cp;
                    sym_b = e1_sym_bits(r, ibit, page_sel);
                }
            }
        }
        e1_tile_ck *dst = &o[(size_t)t * stride];
        dst->cp = cp_t;
        dst->cp_w = cp_w;
        dst->j_w = j_w;
        dst->sym = sym_a | (sym_b << 2) | E1_CK_ACTIVE | err;
    }
}

/* Carrier plan of one (epoch, channel): checkpoints phi at every tile start and returns the
 * phase after the epoch's last sample (the only state that crosses epochs, :531-532). */
E1_HD double e1_plan_carr_epoch(const e1_epoch_rec *r, e1_tile_ck *o, int stride, double phi, int n_samp, int tile,
                                int tiles_per_epoch, double delt)
{
    const int prn = r->prn;
    if (prn < 1 || prn > E1C_N_PRN)
        return phi;
    if (r->flags & E1_REC_SET_PHASE)
        phi = r->carr_phase_init; /* src/channel.cpp:98-99 */
    const double sp = e1_mul(r->f_carr, delt);
    for (int t = 0; t < tiles_per_epoch; t++) {
        int64_t k0 = (int64_t)t * tile, k1 = k0 + tile;
        if (k1 > n_samp)
            k1 = n_samp;
        o[(size_t)t * stride].phi = phi;
        phi = e1_carr_advance(phi, sp, k0, k1);
    }
    return phi;
}

/* ------------------------------------------------------------------ parallel carrier planner
 *
 * The carrier recurrence is serial over the whole run, but it is *translation invariant*: two
 * trajectories whose starts differ by D, a multiple of 2^-52 cycle, stay exactly D apart for as
 * long as every pair of corresponding values lies in the same binade -- every rounding grid
 * below 1.0 divides 2^-52, so fl(x + D + s) = fl(x + s) + D -- with one exception: a step s that
 * is a multiple of 2^-53 can make exact ties at the wrap step (x + s an odd multiple of 2^-53 in
 * [1,2), where the grid is 2^-52), whose round-half-even direction depends on the parity of
 * d = D / 2^-52.  With d odd the translated walk lands one grid step to the other side, i.e. its
 * translation becomes d + 1 (guess rounded down) or d - 1 (guess rounded up) -- even either way, so
 * only the FIRST tie wrap of an epoch matters: the span pass records where it is and which way the
 * guess went (e1_unit.tie_k / tie_dir) and the chain splits the epoch's translation there.
 * Values right after a wrap are multiples of 2^-52.  Hence:
 *
 *   drift pass   (parallel, per epoch)   walk each epoch from an *ideal* start phase; the
 *                                        measured end-start gives that epoch's rounding drift
 *   estimate     (serial, O(1)/epoch)    prefix of the drifts -> start phase of every epoch to
 *                                        ~1e-15 cycle
 *   span pass    (parallel, per epoch)   from a *guessed* post-wrap value at the last wrap of
 *                                        the previous epoch ("anchor") walk exactly to the end
 *                                        of the epoch, writing hat checkpoints and the interval
 *                                        [lo,hi) of translations D for which the walk is valid
 *   chain        (serial, O(1)/epoch)    D = true anchor value - guess; if lo <= D < hi the
 *                                        epoch's checkpoints are hat + D, else (rare) the epoch
 *                                        is walked serially from its true start
 * Correctness never depends on the guesses, only speed does.                                   */
/* The unit of parallel work is a SPAN: a run of span_tiles consecutive tiles of one epoch of one
 * channel (an epoch is cut into S = ceil(tiles_per_epoch / span_tiles) spans, e1_span_geometry), so
 * even a single epoch gives every channel several walks to run side by side and the serial
 * dependency per thread is a span, not an epoch.  Everything above holds with "epoch" read as "span":
 * a span's records are its epoch's, E1_REC_SET_PHASE applies to the epoch's first span only.      */
typedef struct e1_span_geo {
    int span_tiles;      /* tiles per span (the epoch's last span may be shorter) */
    int spans_per_epoch; /* S */
} e1_span_geo;
E1_HD e1_span_geo e1_span_geometry(int tiles_per_epoch)
{
    e1_span_geo g;
    int st = (tiles_per_epoch + 7) / 8; /* at most 8 spans per epoch ... */
    if (st < 8)
        st = 8; /* ... of at least 8 tiles: the guesses need a wrap inside the span before */
    if (st > tiles_per_epoch)
        st = tiles_per_epoch;
    g.span_tiles = st;
    g.spans_per_epoch = (tiles_per_epoch + st - 1) / st;
    return g;
}
/* samples in span s of an epoch of n_samp samples */
E1_HD int e1_span_samples(const e1_span_geo *g, int s, int n_samp, int tile)
{
    const int64_t k0 = (int64_t)s * g->span_tiles * tile, k1 = k0 + (int64_t)g->span_tiles * tile;
    return (int)((k1 < n_samp ? k1 : n_samp) - k0);
}

#define E1_UNIT_NONE 0   /* slot idle this epoch                                                   */
#define E1_UNIT_EXACT 1  /* walked from an exactly known start (batch start / E1_REC_SET_PHASE)    */
#define E1_UNIT_HAT 2    /* walked from a guessed anchor: needs the chain's translation            */
#define E1_UNIT_SERIAL 3 /* not eligible for a guess: the chain walks it                           */

typedef struct e1_unit { /* one per (span, channel), 64 bytes */
    double anchor_p;  /* HAT: guessed |phase| right after the anchor wrap (multiple of 2^-52)   */
    double end_phi;   /* signed phase after the epoch's last sample (hat for HAT units)         */
    double last_p;    /* |phase| right after the last wrap inside this epoch (hat for HAT)      */
    double lo, hi;    /* HAT: valid translations, lo <= D < hi                                  */
    int32_t anchor_k; /* HAT: sample index (1..N) of the anchor wrap inside its span            */
    int32_t last_k;   /* sample index (1..N) of the last wrap inside this span; -1: none, the walk stayed
                         in the aligned regime (an earlier wrap can still serve as anchor); -2: none and
                         the walk was not aligned throughout                                       */
    int32_t type;
    int32_t neg;      /* sign of the walk (1: phase <= 0)                                       */
    int32_t tie;         /* HAT: +-k, k = sample index (1..N) of the first tie wrap inside this span; + the
                            guess rounded down there, - it rounded up; 0: no tie wrap               */
    int32_t anchor_back; /* HAT: the anchor wrap lies in the span this many units before this one (>= 1) */
} e1_unit;
#define E1_MAX_BACK 64 /* how far back the span pass looks for a wrap to anchor on */

/* Translation of one (channel, epoch): its carrier checkpoints are ck.phi + a for tile starts before
 * sample k_split and ck.phi + b from there on (k_split = 0 and a = b unless a tie wrap split it). */
typedef struct e1_trans {
    double a, b;
    int32_t k_split, pad;
} e1_trans;

typedef struct e1_span_track {
    double lo, hi;
    int64_t last_k;
    double last_p;
    int64_t tie_k; /* first wrap whose sum was an exact tie (set by the caller to -1) */
    int tie_dir;
} e1_span_track;

E1_HD double e1_binade_floor(double x) { return e1_from_bits((e1_bits(x) >> 52) << 52); }
E1_HD double e1_binade_top(double x) { return e1_from_bits(((e1_bits(x) >> 52) + 1) << 52); }

/* Aligned-regime walk of the phase magnitude: a in [0,1) grows by t in (0,0.5) per sample and
 * wraps at 1.0, from sample k to k_end.  Bit-identical to the literal loop (same binade jumps
 * as e1_walk_up) and in addition
 *   - writes sign*a to o[(kt/tile)*stride].phi for every tile start kt with k < kt < n_emit
 *     (o == NULL: no checkpoints),
 *   - narrows tr->[lo,hi) to the translations that keep every visited value in its binade,
 *   - records the last wrap in tr->last_k / last_p and the first tie wrap in tr->tie_k / tie_dir. */
E1_HD double e1_span_walk(double a, double t, int32_t k, int32_t k_end, int tile, int32_t n_emit, e1_tile_ck *o,
                          int stride, int neg, e1_span_track *tr)
{
    /* sample counters are 32-bit: a span is a few hundred thousand samples at most (the 64-bit versions cost two
       instructions for every compare and add of the walk) */
    int32_t ti = k / tile + 1; /* next tile index to checkpoint */
    int32_t kt = o ? ti * tile : 0x7fffffff;
    double lo = tr->lo, hi = tr->hi;
#define E1_EMIT(val)                                                                                                   \
    do {                                                                                                               \
        if (kt < n_emit)                                                                                               \
            o[(size_t)ti * stride].phi = neg ? -(val) : (val);                                                         \
        kt += tile;                                                                                                    \
        ti++;                                                                                                          \
    } while (0)
    while (k < k_end) {
        double x1 = e1_add(a, t);
        k++;
        if (x1 >= 1.0) {
            double m = e1_add(1.0, -x1);
            if (m > lo)
                lo = m;
            if (tr->tie_k < 0) {
                /* Fast2Sum (a >= 1/2 > t): err = (a + t) - x1 exactly.  x1 is on the 2^-52 grid, so the
                   rounding was an exact tie iff |err| = 2^-53 */
                const double err = e1_add(t, -e1_add(x1, -a));
                if (err == 1.1102230246251565e-16 || err == -1.1102230246251565e-16) {
                    tr->tie_k = k;
                    tr->tie_dir = err > 0.0 ? 1 : -1;
                }
            }
            a = e1_add(x1, -1.0);
            tr->last_k = k;
            tr->last_p = a;
            if (k == kt)
                E1_EMIT(a);
            continue;
        }
        {
            double up = e1_add(e1_binade_top(x1), -x1), dn = e1_add(e1_binade_floor(x1), -x1);
            if (up < hi)
                hi = up;
            if (dn > lo)
                lo = dn;
        }
        a = x1;
        if (k == kt)
            E1_EMIT(a);
        if (k + 1 >= k_end)
            continue;
        double x2 = e1_add(x1, t), x3 = e1_add(x2, t);
        int64_t b1 = e1_bits(x1), b2 = e1_bits(x2), b3 = e1_bits(x3);
        if ((((b1 ^ b2) | (b1 ^ b3)) >> 52) != 0 || x3 >= 1.0)
            continue;
        int64_t d = b3 - b2;
        int64_t end = ((b1 >> 52) + 1) << 52;
        int32_t n;
        k++; /* now at x2 */
        a = x2;
        if (k == kt)
            E1_EMIT(a);
        if (d == 0)
            n = k_end - k; /* stuck for good */
        else {
            const int64_t nn = e1_div_binade(end - 1 - b2, d);
            n = nn > (int64_t)(k_end - k) ? k_end - k : (int32_t)nn;
        }
        while (kt <= k + n && kt < n_emit) {
            double v = e1_from_bits(b2 + (int64_t)(kt - k) * d);
            E1_EMIT(v);
        }
        a = e1_from_bits(b2 + (int64_t)n * d);
        k += n;
        {
            double up = e1_add(e1_from_bits(end), -a);
            if (up < hi)
                hi = up;
        }
    }
#undef E1_EMIT
    tr->lo = lo;
    tr->hi = hi;
    return a;
}

/* One whole epoch from an exactly known start phase (any sign combination): checkpoints are
 * final.  Returns the phase after the last sample; last_k and last_p describe the last wrap when
 * the whole epoch ran in the aligned regime (else *last_k = -1).                              */
E1_HD double e1_carr_epoch_exact(double phi, double sp, int n_samp, int tile, int tiles_per_epoch, e1_tile_ck *o,
                                 int stride, int32_t *last_k, double *last_p, int32_t *neg_out)
{
    const int neg = (phi < 0.0) || (phi == 0.0 && sp < 0.0);
    const int aligned = (phi == 0.0) || (neg == (sp < 0.0));
    *last_k = -2;
    *last_p = 0.0;
    *neg_out = neg;
    if (sp == 0.0 || !aligned) {
        for (int t = 0; t < tiles_per_epoch; t++) {
            int64_t k0 = (int64_t)t * tile, k1 = k0 + tile;
            if (k1 > n_samp)
                k1 = n_samp;
            o[(size_t)t * stride].phi = phi;
            phi = e1_carr_advance(phi, sp, k0, k1);
        }
        return phi;
    }
    e1_span_track tr;
    tr.lo = -2.0;
    tr.hi = 2.0;
    tr.last_k = -1;
    tr.last_p = 0.0;
    tr.tie_k = 0; /* not tracked */
    tr.tie_dir = 0;
    o[0].phi = phi;
    double a = e1_span_walk(e1_fabs(phi), e1_fabs(sp), 0, n_samp, tile, n_samp, o, stride, neg, &tr);
    *last_k = (int32_t)tr.last_k;
    *last_p = tr.last_p;
    return neg ? -a : a;
}

E1_HD int e1_rec_active(const e1_epoch_rec *r) { return r->prn >= 1 && r->prn <= E1C_N_PRN; }

/* What the carrier planner needs from one epoch record, stored channel-major ([channel][epoch])
 * so the serial per-channel passes read contiguous memory. */
typedef struct e1_prep {
    double sp;      /* fl(f_carr * delt): carrier phase step per sample (:531)                  */
    double init;    /* carr_phase_init (valid with E1_PREP_SET_PHASE)                            */
    uint32_t flags; /* E1_PREP_*                                                                 */
    int32_t n;      /* samples in this span                                                      */
} e1_prep;
#define E1_PREP_ACTIVE 1u
#define E1_PREP_SET_PHASE 2u

/* span s (n samples) of the epoch described by r */
E1_HD void e1_v2_prep(const e1_epoch_rec *r, double delt, int s, int n, e1_prep *p)
{
    const int act = e1_rec_active(r);
    p->sp = act ? e1_mul(r->f_carr, delt) : 0.0;
    p->init = r->carr_phase_init;
    p->flags = (act ? E1_PREP_ACTIVE : 0u) | ((act && s == 0 && (r->flags & E1_REC_SET_PHASE)) ? E1_PREP_SET_PHASE : 0u);
    p->n = n;
}

/* phase + whole-epoch advance, folded back into (-1,1) the way :532 does (sign kept) */
E1_HD double e1_ideal_next(double g, double sp, int n_samp)
{
    double v = g + (double)n_samp * sp;
    if (v >= 1.0 || v <= -1.0)
        v -= (double)(long long)v;
    return v;
}

/* K0: ideal (rounding-free, double precision) start phase of every epoch of one channel. */
E1_HD void e1_v2_ideal_prefix(const e1_prep *pp, int n_epochs, double phi0, double *g_out)
{
    double g = phi0;
    for (int e = 0; e < n_epochs; e++) {
        const uint32_t f = pp[e].flags;
        if (f & E1_PREP_SET_PHASE)
            g = pp[e].init;
        g_out[e] = g;
        if (f & E1_PREP_ACTIVE)
            g = e1_ideal_next(g, pp[e].sp, pp[e].n);
    }
}

/* drift pass: where the span ends when it starts at its ideal phase g, as an ESTIMATE (the chain
 * validates every guess exactly; this only decides how many it accepts).  The rounding drift of the
 * serial recurrence against the ideal line is systematic -- inside a binade every addition rounds the
 * same way by the same amount, the step's bits below the binade's ulp -- so the first
 * n / E1_DRIFT_DECIM samples are walked exactly and their drift is scaled to the span.  (With a Doppler
 * that changes from epoch to epoch the accumulated drift is a random walk of a few 1e-10 cycles over
 * 300 s and the pass hardly matters; with a constant Doppler it grows linearly to ~1e-8 and, left out,
 * costs about one span in a hundred a serial walk.) */
#define E1_DRIFT_DECIM 16
E1_HD double e1_v2_drift_unit(const e1_prep *p, double g)
{
    if (!(p->flags & E1_PREP_ACTIVE))
        return g;
    const int q = p->n / E1_DRIFT_DECIM;
    if (q < 1024)
        return e1_carr_advance(g, p->sp, 0, p->n);
    double d = e1_carr_advance(g, p->sp, 0, q) - e1_ideal_next(g, p->sp, q);
    if (d > 0.5)
        d -= 1.0;
    else if (d < -0.5)
        d += 1.0;
    if (!(d < 1e-6 && d > -1e-6)) /* not a drift: the two folded differently (sign change near zero) */
        d = 0.0;
    double v = e1_ideal_next(g, p->sp, p->n) + d * ((double)p->n / (double)q);
    if (v >= 1.0)
        v -= 1.0;
    else if (v <= -1.0)
        v += 1.0;
    return v;
}

/* K1: refined start-phase estimates of one channel.  The drift pass walked epoch e exactly from
 * the ideal start g[e] to end[e]; the true start est[e] differs from g[e] by a tiny eta, and by
 * translation the true end is end[e] + eta (to within an ulp or two).                          */
E1_HD double e1_v2_estimate_prefix(const e1_prep *pp, int n_epochs, double phi0, const double *g, const double *end,
                                   double *est)
{
    double cur = phi0;
    for (int e = 0; e < n_epochs; e++) {
        const uint32_t f = pp[e].flags;
        if (f & E1_PREP_SET_PHASE)
            cur = pp[e].init;
        est[e] = cur;
        if (!(f & E1_PREP_ACTIVE))
            continue;
        double eta = cur - g[e];
        if (eta > 0.5)
            eta -= 1.0;
        else if (eta < -0.5)
            eta += 1.0;
        cur = end[e] + eta;
        if (cur >= 1.0)
            cur -= 1.0;
        else if (cur <= -1.0)
            cur += 1.0;
    }
    return cur; /* estimate for the epoch after the last one: lets a caller continue chunk by chunk */
}

/* span pass for one (epoch, channel): p / p_prev are this channel's prep records of epoch e and
 * e-1, o its first tile checkpoint of epoch e (tile stride `stride`). */
/* span pass for one (span, channel): pr = &prep[u] and est = &est[u] of this channel (earlier spans of
 * the batch at negative offsets), o its first tile checkpoint (tile stride `stride`).  The anchor is the
 * last wrap before this span according to the estimates -- normally inside the span just before, at
 * low Doppler up to E1_MAX_BACK spans back. */
E1_HD void e1_v2_span_unit(const e1_prep *pr, int e, double phi_batch_start, const double *est, int tile, e1_tile_ck *o,
                           int stride, e1_unit *u)
{
    const int n_samp = pr->n, tiles_per_epoch = (pr->n + tile - 1) / tile; /* of this span */
    u->type = E1_UNIT_NONE;
    u->last_k = -1;
    u->anchor_k = -1;
    u->anchor_p = 0.0;
    u->end_phi = 0.0;
    u->last_p = 0.0;
    u->lo = 0.0;
    u->hi = 0.0;
    u->neg = 0;
    u->tie = 0;
    u->anchor_back = 0;
    if (!(pr->flags & E1_PREP_ACTIVE))
        return;
    const double sp = pr->sp;
    if (e == 0 || (pr->flags & E1_PREP_SET_PHASE)) {
        const double phi = (pr->flags & E1_PREP_SET_PHASE) ? pr->init : phi_batch_start;
        u->type = E1_UNIT_EXACT;
        u->end_phi = e1_carr_epoch_exact(phi, sp, n_samp, tile, tiles_per_epoch, o, stride, &u->last_k, &u->last_p, &u->neg);
        return;
    }
    u->type = E1_UNIT_SERIAL;
    if (sp == 0.0 || !(e1_fabs(sp) < 0.5))
        return;
    const int neg = sp < 0.0;
    const double t1 = e1_fabs(sp);
    /* find the span with the last wrap: unwrapped phase a0 + k*t0 along the direction of motion, a0 < 0
       when that span started in the mixed regime (phase and Doppler of opposite sign) */
    int back = 0;
    double p = 0.0, t0 = 0.0;
    int64_t kL = 0;
    for (int b = 1; b <= E1_MAX_BACK && b <= e; b++) {
        const e1_prep *q = pr - b;
        if (!(q->flags & E1_PREP_ACTIVE) || q->sp == 0.0 || (q->sp < 0.0) != neg || !(e1_fabs(q->sp) < 0.5))
            return;
        const double eq = est[-b], tq = e1_fabs(q->sp);
        const double a0 = (eq == 0.0 || (eq < 0.0) == neg) ? e1_fabs(eq) : -e1_fabs(eq);
        if (!(a0 < 1.0) || !(a0 > -1.0))
            return;
        const double total = a0 + (double)q->n * tq;
        const double W = (double)(long long)total;
        if (W >= 1.0) {
            double kd = (W - a0) / tq;
            kL = (int64_t)kd;
            if ((double)kL < kd)
                kL++;
            for (int tries = 0; tries < 3; tries++) { /* p = a0 + kL*tq - W in double-double, want 0 <= p < tq */
#if defined(__CUDA_ARCH__)
                const double hi_ = __dmul_rn((double)kL, tq), lo_ = __fma_rn((double)kL, tq, -hi_);
#else
                const double hi_ = (double)kL * tq, lo_ = __builtin_fma((double)kL, tq, -hi_);
#endif
                p = ((hi_ - W) + a0) + lo_;
                if (p < 0.0)
                    kL++;
                else if (p >= tq)
                    kL--;
                else
                    break;
            }
            if (!(p >= 0.0) || !(p < tq) || kL < 1 || kL > q->n)
                return;
            back = b;
            t0 = tq;
            break;
        }
        if (q->flags & E1_PREP_SET_PHASE)
            return; /* the trajectory starts at q: nothing earlier to anchor on */
    }
    if (!back)
        return; /* no wrap within reach */
    /* round the guess to the post-wrap grid */
    const double two52 = 4503599627370496.0;
    p = (double)(long long)(p * two52 + 0.5) / two52;
    e1_span_track tr;
    tr.lo = -p;
    tr.hi = 2.0;
    tr.last_k = -1;
    tr.last_p = 0.0;
    tr.tie_k = 0;
    tr.tie_dir = 0;
    double a = e1_span_walk(p, t0, kL, (pr - back)->n, tile, 0, (e1_tile_ck *)0, 0, neg, &tr);
    for (int b = back - 1; b >= 1 && tr.last_k == -1; b--) /* the spans in between: no wrap expected */
        a = e1_span_walk(a, e1_fabs((pr - b)->sp), 0, (pr - b)->n, tile, 0, (e1_tile_ck *)0, 0, neg, &tr);
    if (tr.last_k != -1)
        return; /* the guessed anchor was not the last wrap after all */
    tr.tie_k = -1; /* ties only matter at this span's own wraps */
    o[0].phi = neg ? -a : a;
    a = e1_span_walk(a, t1, 0, n_samp, tile, n_samp, o, stride, neg, &tr);
    u->type = E1_UNIT_HAT;
    u->neg = neg;
    u->anchor_k = (int32_t)kL;
    u->anchor_back = back;
    u->anchor_p = p;
    u->end_phi = neg ? -a : a;
    u->last_k = (int32_t)tr.last_k;
    u->last_p = tr.last_p;
    u->lo = tr.lo;
    u->hi = tr.hi;
    u->tie = tr.tie_k >= 0 ? (int32_t)tr.tie_k * tr.tie_dir : 0;
}

/* chain of one channel: validates HAT units, walks the others, writes the per-epoch translation
 * delta[e] (signed; the epoch's checkpoints are ck.phi + delta) and returns the final phase.
 * (e1_trans).  The state that runs along the chain is e1_chain_state; e1_v2_chain_step consumes one epoch, so
 * the kernel can stage units / deltas through shared memory chunk by chunk.
 * stats[0] += epochs walked serially, stats[1] += HAT units accepted. */
typedef struct e1_chain_state {
    double phi;    /* signed phase at the first sample of the next epoch                */
    double prev_p; /* |phase| right after the most recent wrap                           */
    int32_t prev_k; /* ... its sample index inside span prev_u                            */
    int32_t prev_u;
    int prev_ok, prev_neg; /* prev_ok: there is such a wrap and the walk has been aligned since */
} e1_chain_state;

E1_HD void e1_chain_init(e1_chain_state *s, double phi0)
{
    s->phi = phi0;
    s->prev_p = 0.0;
    s->prev_k = -1;
    s->prev_u = -1;
    s->prev_ok = 0;
    s->prev_neg = 0;
}

/* u: this epoch's unit; sp: its carrier step; ck_e: this channel's first checkpoint of this epoch
 * (tiles `stride` apart), only touched when the epoch has to be walked serially. */
/* ui: index of this span in the batch. */
E1_HD e1_trans e1_v2_chain_step(e1_chain_state *s, const e1_unit *u, int ui, double sp, int n_samp, int tile, int tiles_per_epoch,
                                e1_tile_ck *ck_e, int stride, unsigned long long *stats)
{
    const int type = u->type;
    e1_trans tr;
    tr.a = tr.b = 0.0;
    tr.k_split = 0;
    tr.pad = 0;
    if (type == E1_UNIT_NONE) {
        s->prev_ok = 0;
        return tr;
    }
    int32_t last_k = u->last_k, neg = u->neg;
    double last_p = u->last_p;
    if (type == E1_UNIT_EXACT) {
        s->phi = u->end_phi;
        s->prev_ok = 0; /* whatever wrap came before belongs to another trajectory */
    } else {
        int ok = 0;
        if (type == E1_UNIT_HAT && s->prev_ok && s->prev_neg == neg && s->prev_u == ui - u->anchor_back && s->prev_k == u->anchor_k) {
            const double D = e1_add(s->prev_p, -u->anchor_p); /* a multiple of 2^-52, |D| < 1 */
            double D2 = D;
            if (u->tie != 0 && ((long long)e1_mul(D, 4503599627370496.0) & 1LL))
                D2 = e1_add(D, u->tie > 0 ? 2.220446049250313e-16 : -2.220446049250313e-16);
            if (D >= u->lo && D < u->hi && D2 >= u->lo && D2 < u->hi) {
                ok = 1;
                tr.a = neg ? -D : D;
                tr.b = neg ? -D2 : D2;
                tr.k_split = D2 != D ? (u->tie > 0 ? u->tie : -u->tie) : 0;
                if (tr.k_split == 0)
                    tr.a = tr.b;
                s->phi = e1_add(u->end_phi, tr.b);
                last_p = e1_add(last_p, D2); /* the last wrap is at or after the first tie wrap */
                stats[1]++;
            }
        }
        if (!ok) {
            s->phi = e1_carr_epoch_exact(s->phi, sp, n_samp, tile, tiles_per_epoch, ck_e, stride, &last_k, &last_p, &neg);
            stats[0]++;
        }
    }
    if (last_k >= 1) {
        s->prev_ok = 1;
        s->prev_neg = neg;
        s->prev_k = last_k;
        s->prev_u = ui;
        s->prev_p = last_p;
    } else if (last_k != -1 || s->prev_neg != neg) {
        s->prev_ok = 0; /* no wrap and not (or not in the same direction) aligned: the older wrap is no anchor any more */
    }
    return tr;
}

/* ck points at this channel's first checkpoint of the batch; the tiles of consecutive spans follow
 * each other `stride` apart (tile-major layout [epoch][tile][channel], stride = max_chan). */
E1_HD double e1_v2_chain(const e1_prep *pp, int n_units, double phi0, int tile, e1_unit *units, e1_tile_ck *ck, int stride,
                         e1_trans *delta, unsigned long long *stats)
{
    e1_chain_state s;
    e1_chain_init(&s, phi0);
    size_t tile0 = 0;
    for (int u = 0; u < n_units; u++) {
        const int tiles = (pp[u].n + tile - 1) / tile;
        delta[u] = e1_v2_chain_step(&s, &units[u], u, pp[u].sp, pp[u].n, tile, tiles, ck + tile0 * (size_t)stride, stride, stats);
        tile0 += (size_t)tiles;
    }
    return s.phi;
}

/* Ambiguity half-widths of the fast path for a run of R consecutive samples of a tile of T samples,
 * in units of 2^-32 of the truncated quantity.  The fast path steps 32-bit truncations of the
 * fixed-point phases, so sample i of the run is below the exact closed form by less than i+1 <= R
 * units of 2^-32 cycle (carrier; times 511 in index units) / 2^-32 half-chip (code); the closed form
 * itself is within e1_thr_carr / e1_thr_code of the serial value.  Adding the half-width as a bias
 * makes "possibly on the other side of an integer" read as "fraction < 2*half-width + 1".        */
E1_HD uint32_t e1_tc_carr(uint32_t thr_carr, int R) { return thr_carr + 511u * (uint32_t)R; }
/* carry-walked runs (e1_run_cw) step floor(511 U / 2^32) by floor(511 dU / 2^32): sample i of the run is below the exact
   closed form by less than i + 1 <= n_samples units (one floor at the start, one per step) */
E1_HD uint32_t e1_tc_carr_cw(uint32_t thr_carr, int n_samples) { return thr_carr + (uint32_t)n_samples + 1u; }
E1_HD uint32_t e1_lim_carr_cw(uint32_t thr_carr, int n_samples) { return e1_tc_carr_cw(thr_carr, n_samples) + thr_carr + 1u; }
/* code: a run that contains the code wrap keeps stepping the pre-wrap closed form, which is within
 * thr_code of the serial value at the wrap; the serial recurrence restarts there and drifts by at
 * most thr_code more -> 2*thr_code on either side */
E1_HD uint32_t e1_tc_code(uint32_t thr_code, int R) { return 2u * thr_code + (uint32_t)R + 1u; }
E1_HD uint64_t e1_bias_h(uint32_t tc_code) { return (uint64_t)tc_code << 19; }
E1_HD uint32_t e1_lim_carr(uint32_t tc_carr, uint32_t thr_carr) { return tc_carr + thr_carr + 1u; }
E1_HD uint32_t e1_lim_code(uint32_t tc_code, uint32_t thr_code) { return tc_code + 2u * thr_code + 1u; }

/* translation of the checkpoint of the tile that starts at sample k0 of its epoch */
E1_HD double e1_trans_at(const e1_trans *t, int k0) { return k0 >= t->k_split ? t->b : t->a; }

/* Tile checkpoint + epoch record -> the per-channel parameters the sample loop reads. */
E1_HD uint32_t e1_float_bits(float f)
{
#if defined(__CUDA_ARCH__)
    return __float_as_uint(f);
#else
    uint32_t u;
    memcpy(&u, &f, 4);
    return u;
#endif
}
E1_HD float e1_bits_float(uint32_t u)
{
#if defined(__CUDA_ARCH__)
    return __uint_as_float(u);
#else
    float f;
    memcpy(&f, &u, 4);
    return f;
#endif
}

E1_HD void e1_make_par(const e1_tile_ck *c, const e1_epoch_rec *r, double delt, int tile, double delta, uint32_t tc_code,
                       e1_chan_par *p, uint32_t cfg_flags = 0u)
{
    p->phi = e1_add(c->phi, delta); /* planner translation of this epoch (exact, see e1_v2_chain) */
    p->cp = c->cp;
    p->sp = e1_mul(r->f_carr, delt);
    p->sc = e1_mul(r->f_code, delt);
    const int neg = (p->phi < 0.0) || (p->phi == 0.0 && p->sp < 0.0);
    const int aligned = (p->phi == 0.0) || (p->sp == 0.0) || (neg == (p->sp < 0.0));
    p->U0 = e1_to_fixed(p->phi, E1_CARR_FIX);
    uint64_t s = e1_to_fixed(p->sp, E1_CARR_FIX);
    /* planner (D, S) pairs -> field order of the code words: D on the high bit, D^S on the low bit */
    const uint32_t sa = c->sym & 3u, sb = (c->sym >> 2) & 3u;
    const uint32_t da = ((sa & 1u) << 1) | ((sa ^ (sa >> 1)) & 1u), db = ((sb & 1u) << 1) | ((sb ^ (sb >> 1)) & 1u);
    uint32_t misc = da | (db << 2) | (neg ? E1_PAR_NEG : 0u);
    /* outside the fast path's domain (|phase| < 1, a run's index excursion inside the table's
       extension -- 20 kHz of Doppler at 2.6 MS/s --, at most one half-chip per sample so the code
       window advances by carries): generic path */
    if (!(e1_fabs(p->phi) < 1.0) || !(e1_fabs(p->sp) < E1C_FAST_SP_MAX) || !(p->sc < 0.4999))
        misc |= E1_PAR_FORCE;
    if (!aligned) {
        /* |phi| shrinks by s per sample and, where it would go below zero, the phase changes sign and
           the magnitude grows again (:531-532 keeps the sign of the sum).  In fixed point that is the
           two's complement of U from the first sample j_z with U0 - j_z*s < 0 on. */
        if (!(misc & E1_PAR_FORCE) && s != 0ull) {
            const uint64_t jz = p->U0 / s + 1ull;
            if (jz < (uint64_t)tile + 2ull && jz < 0x7fffull)
                misc |= E1_PAR_HASZ | ((uint32_t)jz << 16);
        }
        s = 0ull - s;
    }
    p->dU = s;
    p->dH = e1_to_fixed(p->sc, E1_CODE_FIX);
    p->dF = (uint32_t)(p->dH >> 19);
    const uint64_t bias = e1_bias_h(tc_code);
    p->HA = e1_to_fixed(p->cp, E1_CODE_FIX) + bias;
    p->j_w = c->j_w;
    p->HB = (c->j_w == E1C_NO_WRAP) ? p->HA : e1_to_fixed(c->cp_w, E1_CODE_FIX) + bias - (uint64_t)(uint32_t)c->j_w * p->dH;
    p->pat_a = ((da & 2u) ? 0xAAAAAAAAu : 0u) | ((da & 1u) ? 0x55555555u : 0u);
    p->pat_b = ((db & 2u) ? 0xAAAAAAAAu : 0u) | ((db & 1u) ? 0x55555555u : 0u);
    p->code_off = (uint32_t)(r->prn - 1) * E1C_CODE_WORDS_PER_PRN;
    p->Dlo = 0u;
    p->ev_n0 = p->ev_rcp_f = p->ev_rcp_g = 0u;
    if (!(misc & (E1_PAR_FORCE | E1_PAR_HASZ)) && e1_fabs(p->sp) < 1.0 / 512.0) {
        /* step of the table position 511 |phase| / 2^32 (entry in the high word): floor(511 dU / 2^32) with dU read as
           a signed 64-bit number -- |D| < 2^32 because |sp| < 1/512 -- and negated when the table is walked mirrored */
        const uint64_t lo = (p->dU & 0xffffffffull) * 511ull;                          /* < 2^41 */
        int64_t D = (int64_t)(int32_t)(uint32_t)(p->dU >> 32) * 511 + (int64_t)(lo >> 32); /* floor: the dropped bits are non-negative */
        if (misc & E1_PAR_NEG)
            D = -D;
        p->Dlo = (uint32_t)(uint64_t)D;
        misc |= E1_PAR_SLOW | (D < 0 ? E1_PAR_DOWN : 0u);
        if (cfg_flags & E1_INT_EV) {
            /* event-driven runs: both index fractions are 32-bit accumulators whose CARRIES are the events.  Not for a
               step that divides 2^32 (n0 steps land exactly on 2^32: the carry test of e1_ev_sub would miss it) and
               only while E1C_EV_RUN samples cross fewer than 15 half-chips (one code window) */
            const uint32_t dF = p->dF, dG = D < 0 ? (uint32_t)(uint64_t)(-D) : (uint32_t)(uint64_t)D;
            if (dF != 0u && (uint64_t)dF * E1C_EV_RUN < (14ull << 32) && dG < 0x40000000u) {
                const uint64_t n0c = 0x100000000ull / dF, n0r = dG ? 0x100000000ull / dG : 0xffffull;
                if (n0c <= 0xffffull && n0c * dF != 0x100000000ull && (dG == 0u || n0r * dG != 0x100000000ull)) {
                    p->ev_n0 = (uint32_t)n0c | ((n0r > 0xffffull ? 0xffffu : (uint32_t)n0r) << 16);
                    p->ev_rcp_f = e1_float_bits(1.0f / (float)dF);
                    p->ev_rcp_g = dG ? e1_float_bits(1.0f / (float)dG) : 0x7f800000u;
                    misc |= E1_PAR_EV;
                }
            }
        }
    }
    p->misc = misc;
    if (cfg_flags & (E1B200_CFG_CBOC | E1B200_CFG_GAIN)) { /* float path: gain[i] / 2^7 (src/galileo-sdr.cpp:477), exact in float below 2^24 */
        const int32_t g = ((cfg_flags & E1B200_CFG_GAIN) && r->gain_q7 != 0) ? r->gain_q7 : 128;
        p->pat_a = e1_float_bits((float)g * 0.0078125f);
        p->pat_b = 0u;
    }
}

/* ------------------------------------------------------------------ tile-level ambiguity test
 * The sample loop's ambiguity tracking (two running minima, one instruction per channel-sample) asks,
 * run by run, whether a biased index fraction comes closer to an integer than lim_*.  Both fractions
 * are, up to the fast form's truncation slack, ARITHMETIC SEQUENCES modulo one over the whole tile:
 *   carrier  frac(511 (U0 + j dU) / 2^64)      code  frac((H + j dH) / 2^51)
 * so "does any of the tile's samples come that close" is the classic question about the first term of
 * (a + j d) mod M that falls below L, answered in O(log) steps by a Euclid-like descent.  A tile whose
 * widened test finds no such sample is marked E1_PAR_CLEAN and its runs go through the sample loop
 * without the tracking; every other tile (about 3 % of the (tile, channel) pairs: lim_carr + slack is
 * 4e-6 of an index and a tile has 8192 samples) keeps it.  The implication "clean => no run of the
 * tile would have been flagged" is what has to hold; the host test build asserts it run by run
 * (hs_clean_violations in tests/hostsim).
 *
 * e1_first_hit: smallest j in [0, n) with (a + j d) mod M < L, or -1.  0 <= a, d < M <= 2^40,
 * 0 < L <= M, n <= 8192 keeps every intermediate below 2^53 (e1_div_binade).  One level: reflect the
 * circle (x -> L-1-x keeps the zone, reverses the direction) so that the step goes up by d <= M/2;
 * walking up from a >= L nothing can hit until the sequence wraps, and the landing points after the
 * wraps are again an arithmetic sequence, modulo d, with step -(M mod d): recurse on those (at most
 * n d / M of them), then turn the landing number back into a sample number. */
#define E1_AMB_BITS 40
/* floor(num / d) from a reciprocal of d computed once per level (a level divides twice by its step and
 * once by its modulus, which is the previous level's step): 0 <= num < 2^53 and a quotient below 2^41
 * put the product within 2^-11 of the true quotient, so its integer part is off by at most one. */
E1_HD double e1_rcp(int64_t d)
{
#if defined(__CUDA_ARCH__)
    return __drcp_rn(__ll2double_rn(d));
#else
    (void)d;
    return 0.0;
#endif
}
E1_HD int64_t e1_div_rcp(int64_t num, int64_t d, double rcp)
{
#if defined(__CUDA_ARCH__)
    int64_t q = __double2ll_rz(__dmul_rn(__ll2double_rn(num), rcp));
    const int64_t r = num - q * d;
    if (r < 0)
        q--;
    else if (r >= d)
        q++;
    return q;
#else
    (void)rcp;
    return num / d;
#endif
}
template <bool INDEX>
E1_HD int64_t e1_hit_descent(int64_t a, int64_t d, int64_t M, int64_t L, int64_t n)
{
    double rM = e1_rcp(M);
    int64_t sa[INDEX ? 20 : 1], sd[INDEX ? 20 : 1], sM[INDEX ? 20 : 1]; /* INDEX = false: "is there one" only, no unwinding */
    int depth = 0;
    int64_t res = -1;
    for (;;) {
        if (n <= 0)
            break;
        if (a < L) {
            res = 0;
            break;
        }
        if (d == 0)
            break;
        if (2 * d > M) {
            a = M - (a - L + 1);
            d = M - d;
        }
        const double rd = e1_rcp(d);
        const int64_t k0 = e1_div_rcp(M - a + d - 1, d, rd); /* steps to the first wrap */
        if (k0 >= n)
            break;
        const int64_t a0 = a + k0 * d - M; /* first landing, in [0, d) */
        if (a0 < L) {
            res = k0;
            break;
        }
        if (INDEX) {
            if (depth == 20) /* cannot happen (n halves per level); "hit" is the safe answer */
                return 0;
            sa[depth] = a, sd[depth] = d, sM[depth] = M;
            depth++;
        }
        const int64_t n1 = e1_div_rcp(a + (n - 1) * d, M, rM); /* landings within n samples */
        const int64_t r = M - e1_div_rcp(M, d, rd) * d;
        a = a0;
        M = d;
        rM = rd;
        d = r ? M - r : 0;
        n = n1;
    }
    if (res < 0)
        return -1;
    /* res is a sample number of the level it was found on: a landing number of the level above */
    while (INDEX && depth > 0) { /* landing number res of a pushed level -> its sample number */
        depth--;
        res = e1_div_binade((res + 1) * sM[depth] - sa[depth] + sd[depth] - 1, sd[depth]);
    }
    return res;
}
E1_HD int64_t e1_first_hit(int64_t a, int64_t d, int64_t M, int64_t L, int64_t n) { return e1_hit_descent<true>(a, d, M, L, n); }
E1_HD int e1_any_hit(int64_t a, int64_t d, int64_t M, int64_t L, int64_t n) { return e1_hit_descent<false>(a, d, M, L, n) >= 0; }

/* 1 when no sample of the tile (T samples from the checkpoint in p) can make a run of R = E1C_MAX_RUN
 * samples ambiguous in e1_sample_loop.  Slack, in units of 2^-32 of an index:
 *   carrier  the loop's biased fraction f = low32(511 hi32(U_run) + tc_carr + i D) lies in
 *            (r + tc_carr - 511 R, r + tc_carr] with r the exact 511 (U0 + j dU) / 2^32 (mod 2^32);
 *   code     F = low32(H_run >> 19) + i dF lies in (r - R, r] with r = (H + j dH) / 2^19, H = HA for
 *            runs that start before the code wrap, HB (= HA to within thr_code) after;
 *   both     the 40-bit sequence used here is below the exact one by less than (T + 1) / 256 <= 33.
 * FORCE / HASZ tiles are not examined (the kernel takes them through the single-run path anyway). */
E1_HD int e1_par_clean(const e1_chan_par *p, int T, uint32_t tc_carr, uint32_t lim_carr, uint32_t lim_code, uint32_t thr_code,
                        int cw_samples = 2 * E1C_MAX_RUN, int code_run = E1C_MAX_RUN)
{
    if ((p->misc & (E1_PAR_FORCE | E1_PAR_HASZ)) || T > 8192)
        return 0;
    const int64_t M = (int64_t)1 << E1_AMB_BITS;
    const int R = E1C_MAX_RUN, sh = 64 - E1_AMB_BITS, slack = 33;
    {
        /* E1_PAR_SLOW tiles step the carrier position through BOTH halves of a pair from one start: the same
           statement with a run of 2 R samples (bias and limit 511 R larger, e1_tc_carr_pair / e1_lim_carr_pair) */
        /* E1_PAR_SLOW tiles step the full-precision position 511 U / 2^32 through both halves of a pair from one start:
           the loop's biased fraction lies in (r + tc - (cw_samples + 1), r + tc] (e1_carrier_start_p), tc and lim are
           those of a run of cw_samples (e1_tc_carr_cw / e1_lim_carr_cw; 32 or 64 samples per thread) */
        int64_t step_slack = 511 * R;
        if (p->misc & E1_PAR_SLOW) {
            const uint32_t thr_carr = tc_carr - 511u * (uint32_t)R;
            tc_carr = e1_tc_carr_cw(thr_carr, cw_samples);
            lim_carr = e1_lim_carr_cw(thr_carr, cw_samples);
            step_slack = cw_samples + 1;
        }
        const uint64_t G0 = p->U0 * 511ull, GB = p->dU * 511ull; /* mod 2^64: the index fraction, 2^-64 */
        const int64_t a = (int64_t)(((G0 >> sh) + (((uint64_t)tc_carr + slack) << (E1_AMB_BITS - 32))) & (uint64_t)(M - 1));
        const int64_t L = ((int64_t)lim_carr + step_slack + slack + 1) << (E1_AMB_BITS - 32);
        if (L >= M / 2)
            return 0;
        if (e1_any_hit(a, (int64_t)(GB >> sh), M, L, T))
            return 0;
    }
    {
        /* one sequence for the whole tile, from HA: the runs after the code wrap step HB + j dH, which is
           the pre-wrap closed form at the wrap to within thr_code (the serial recurrence restarts there,
           e1_tc_code) and has the same step -- the zone grows by thr_code on either side */
        const int hs = 51 - E1_AMB_BITS;
        const uint64_t m51 = ((uint64_t)1 << 51) - 1ull;
        /* code_run: samples the code fraction is stepped from one start (R in the run kernels, E1C_EV_RUN in the
           event-driven one, whose bias and limit are those of runs of that length as well) */
        const int64_t L = ((int64_t)lim_code + code_run + slack + 1 + 2 * (int64_t)thr_code) << (E1_AMB_BITS - 32);
        const uint64_t bias = ((uint64_t)slack + thr_code) << (E1_AMB_BITS - 32);
        if (L >= M / 2)
            return 0;
        if (e1_any_hit((int64_t)((((p->HA & m51) >> hs) + bias) & (uint64_t)(M - 1)), (int64_t)((p->dH & m51) >> hs), M, L, T))
            return 0;
    }
    return 1;
}

/* Exact table indices of sample j of the tile by walking the reference recurrences exactly
 * from the tile checkpoint (only for samples the closed form flags as ambiguous). */
E1_HD void e1_exact_indices_impl(const e1_chan_par *p, int j, uint32_t *h_out, uint32_t *it_out, double per_chip = 2.0)
{
    /* same values as j literal e1_carr_step / e1_code_step calls, in O(binades) steps */
    const double phi = e1_carr_advance(p->phi, p->sp, 0, j);
    double cp = p->cp;
    int64_t k = 0;
    while (k < j) {
        int w;
        cp = e1_walk_up(cp, p->sc, (double)E1C_CODE_LEN, &k, j, &w);
    }
    *it_out = (uint32_t)(e1_d2i_rz(e1_mul(511.0, phi)) & 511); /* src/galileo-sdr.cpp:509-510 */
    *h_out = (uint32_t)e1_d2i_rz(e1_mul(cp, per_chip));        /* :512 (per_chip = 2; 12 = CBOC sub-chips) */
}
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ void e1_exact_indices(const e1_chan_par *p, int j, uint32_t *h_out, uint32_t *it_out)
{
    e1_exact_indices_impl(p, j, h_out, it_out);
}
static __device__ __noinline__ void e1_exact_indices12(const e1_chan_par *p, int j, uint32_t *s_out, uint32_t *it_out)
{
    e1_exact_indices_impl(p, j, s_out, it_out, 12.0);
}
#else
#define e1_exact_indices e1_exact_indices_impl
E1_HD void e1_exact_indices12(const e1_chan_par *p, int j, uint32_t *s_out, uint32_t *it_out) { e1_exact_indices_impl(p, j, s_out, it_out, 12.0); }
#endif

E1_HD uint32_t e1_funnel_l(uint32_t lo, uint32_t hi, uint32_t n) /* high word of (hi:lo) << (n & 31) */
{
#if defined(__CUDA_ARCH__)
    return __funnelshift_l(lo, hi, n);
#else
    n &= 31u;
    return n ? (hi << n) | (lo >> (32u - n)) : hi;
#endif
}

/* a*b + c on 32 bits.  On the device this is spelled in PTX so the 64 * index + base address of the
 * table lookup stays ONE multiply-add on the FMA pipe instead of being strength-reduced into a
 * 64-bit shift, a mask and an add on the (busier) ALU pipe. */
E1_HD uint32_t e1_mad_u32(uint32_t a, uint32_t b, uint32_t c)
{
#if defined(__CUDA_ARCH__)
    uint32_t d;
    asm("mad.lo.u32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(c));
    return d;
#else
    return a * b + c;
#endif
}

/* Generic form of one channel's contribution to the R consecutive samples starting at tile-relative
 * j0 (src/galileo-sdr.cpp:509-525): full-precision closed form, full-precision ambiguity test, any
 * position of the code wrap; ambiguous samples are resolved by e1_exact_indices and counted in
 * *n_exact.  add[i] receives the term I + 65536*Q of sample j0+i.  lut_lane = carrier table + the
 * caller's copy offset (4 * lane).  Used for the rare runs the fast form hands back. */
E1_HD void e1_channel_run(const e1_chan_par *p, const uint32_t *codes, const unsigned char *lut_lane, int j0, int R, int *add,
                          uint32_t thr_carr, uint32_t thr_code, uint64_t bias_h, unsigned long long *n_exact)
{
    const uint64_t U0 = p->U0, dU = p->dU, HA = p->HA, dH = p->dH, HB = p->HB;
    const int jw = p->j_w;
    const uint32_t misc = p->misc;
    const int jz = (misc & E1_PAR_HASZ) ? (int)(misc >> 16) : E1C_NO_WRAP;
    const uint32_t *code = codes + p->code_off;
    const uint32_t force = (misc >> 5) & 1u;
    uint64_t Ua = U0 + (uint64_t)(uint32_t)j0 * dU;
    for (int i = 0; i < R; i++) {
        const int j = j0 + i;
        const int after = j >= jw;
        const uint64_t H = (after ? HB : HA) + (uint64_t)(uint32_t)j * dH - bias_h;
        const uint32_t ds = after ? (misc >> 2) & 3u : misc & 3u;
        const uint64_t U = j >= jz ? 0ull - Ua : Ua;
        const uint32_t neg = ((misc & E1_PAR_NEG) ? 1u : 0u) ^ (j >= jz ? 1u : 0u);
        /* carrier: y = 511*|phi| in 9.32 fixed point */
        const uint32_t lo511 = e1_umulhi((uint32_t)U, 511u);
        const uint64_t y = (uint64_t)(uint32_t)(U >> 32) * 511u + lo511;
        uint32_t it = (uint32_t)(y >> 32);
        const uint32_t yf = (uint32_t)y;
        /* code: h = trunc(2*code_phase) */
        uint32_t h = (uint32_t)(H >> 51);
        const uint32_t hf = (uint32_t)(H >> 19);
        const uint32_t a = (uint32_t)((yf + thr_carr) < 2u * thr_carr + 1u) | (uint32_t)((hf + thr_code) < 2u * thr_code + 1u) | force;
        if (neg)
            it = (0u - it) & 511u; /* src/galileo-sdr.cpp:509-510 with phi < 0 */
        if (a) {
            e1_exact_indices(p, j, &h, &it);
            (*n_exact)++;
        }
        const uint32_t f = ((code[h >> 4] >> (30u - 2u * (h & 15u))) ^ ds) & 3u; /* (x, x^y) */
        const int sgn = (f & 1u) ? ((f & 2u) ? -1 : 1) : 0;                      /* y - x */
        add[i] = sgn * *(const int32_t *)(lut_lane + (it + E1C_LUT_EXT) * (4u * E1C_LUT_REP));
        Ua += dU;
    }
}

/* FLOAT path (E1B200_CFG_CBOC / E1B200_CFG_GAIN, see include/e1b200.h): one channel's contribution to the R
 * consecutive samples from tile-relative j0, added to the FP32 sums fi / fq.  Same exact phases as e1_channel_run
 * (full-precision closed form, ambiguity test, exact fallback), with the code phase resolved to the SUB-CHIP
 * s = trunc(12 code_phase) (the reference's `(int)(code_phase * 2)`, :512, at CBOC's resolution):
 *     chip = s / 12, k = s % 12, a = (k < 6 ? -1 : +1), b = (k odd ? +1 : -1)          (sboc's sign convention)
 *     Bd = B d, Cs = C s in {-1,+1};  m = eB - eC = alpha a (Bd - Cs) + beta b (Bd + Cs)
 *       = 2 alpha a Bd  when Bd != Cs,   2 beta b Bd  when Bd == Cs                      (one of the two vanishes)
 * so a sample adds  +-(g alpha) or +-(g beta)  times the table's 2 cos / 2 sin.  wa = g alpha, wb = g beta.
 * Closed form vs serial recurrence in sub-chip units: six times the half-chip bound, plus the 2^-44 truncation
 * below and the rounding of fl(12 code_phase) (< 1 unit of 2^-32 together). */
E1_HD void e1_channel_run_float(const e1_chan_par *p, const uint32_t *codes, const unsigned char *lut_lane, int j0, int R, float *fi,
                                float *fq, float wa, float wb, uint32_t thr_carr, uint32_t thr_code, uint64_t bias_h,
                                unsigned long long *n_exact)
{
    const uint64_t U0 = p->U0, dU = p->dU, HA = p->HA, dH = p->dH, HB = p->HB;
    const int jw = p->j_w;
    const uint32_t misc = p->misc;
    const int jz = (misc & E1_PAR_HASZ) ? (int)(misc >> 16) : E1C_NO_WRAP;
    const uint32_t *code = codes + p->code_off;
    const uint32_t force = (misc >> 5) & 1u;
    const uint32_t thr12 = 6u * thr_code + 2u;
    uint64_t Ua = U0 + (uint64_t)(uint32_t)j0 * dU;
    for (int i = 0; i < R; i++) {
        const int j = j0 + i;
        const int after = j >= jw;
        const uint64_t H = (after ? HB : HA) + (uint64_t)(uint32_t)j * dH - bias_h; /* code phase, 2^-52 chip */
        const uint32_t ds = after ? (misc >> 2) & 3u : misc & 3u;                     /* (D << 1) | (D ^ S) */
        const uint64_t U = j >= jz ? 0ull - Ua : Ua;
        const uint32_t neg = ((misc & E1_PAR_NEG) ? 1u : 0u) ^ (j >= jz ? 1u : 0u);
        const uint32_t lo511 = e1_umulhi((uint32_t)U, 511u);
        const uint64_t y = (uint64_t)(uint32_t)(U >> 32) * 511u + lo511;
        uint32_t it = (uint32_t)(y >> 32);
        const uint32_t yf = (uint32_t)y;
        const uint64_t P = (H >> 8) * 12ull; /* 12 code_phase, 2^-44 sub-chip; < 49104 * 2^44 */
        uint32_t s = (uint32_t)(P >> 44);
        const uint32_t sf = (uint32_t)(P >> 12);
        const uint32_t amb = (uint32_t)((yf + thr_carr) < 2u * thr_carr + 1u) | (uint32_t)((sf + thr12) < 2u * thr12 + 1u) | force;
        if (neg)
            it = (0u - it) & 511u;
        if (amb) {
            e1_exact_indices12(p, j, &s, &it);
            (*n_exact)++;
        }
        const uint32_t chip = (s * 43691u) >> 19; /* s / 12 for s < 98304 */
        const uint32_t k = s - 12u * chip;
        const uint32_t hh = 2u * chip + 1u;                                        /* the odd half-chip carries +chip */
        const uint32_t f = (code[hh >> 4] >> (30u - 2u * (hh & 15u))) & 3u;        /* (bneg, bneg ^ cneg) */
        const uint32_t bd = (f >> 1) ^ (ds >> 1);                                  /* 1: B d = -1 */
        const uint32_t differ = (f ^ ds) & 1u;                                     /* (bneg ^ cneg) ^ (D ^ S): B d != C s */
        const uint32_t aneg = k < 6u ? 1u : 0u, bneg = (k & 1u) ^ 1u;
        const uint32_t sneg = bd ^ (differ ? aneg : bneg);
        const float w = differ ? wa : wb;
        const int32_t t = *(const int32_t *)(lut_lane + (it + E1C_LUT_EXT) * (4u * E1C_LUT_REP)); /* 2 (cos + 65536 sin) */
        const int32_t c2 = (int32_t)(int16_t)(uint16_t)t, s2 = (t - c2) >> 16;
        const float ws = sneg ? -w : w;
        fi[i] += ws * (float)c2;
        fq[i] += ws * (float)s2;
        Ua += dU;
    }
}
/* FP32 sum -> the sink's int16: round to nearest even, saturate (the float->int16 store of the north star) */
E1_HD int32_t e1_f2i16(float v)
{
#if defined(__CUDA_ARCH__)
    int32_t r = __float2int_rn(v);
#else
    int32_t r = (int32_t)__builtin_rintf(v);
#endif
    return r > 32767 ? 32767 : (r < -32768 ? -32768 : r);
}
#define E1C_ALPHA_CBOC 0.95346258924559231545 /* sqrt(10/11) */
#define E1C_BETA_CBOC 0.30151134457776362265  /* sqrt(1/11)  */

/* Fast form for the common run: the channel is inside the closed form's domain.  Per sample it costs
 * one 32x32+64 multiply-add (carrier index and its fraction, straight from the run's start value),
 * one address multiply-add, one shared-memory load, one 32-bit add whose carry-out says "next
 * half-chip" and predicates the shift that moves the code window up by one 2-bit field, one
 * arithmetic shift (the signed chip value), one multiply-add into the accumulator, and half a
 * 3-input minimum for each of the two ambiguity fractions: 8 instructions, 4 on the multiplier
 * pipe and 3-4 on the ALU pipe (tools/sass_mix.py).  Everything is conservative with respect to
 * the generic form:
 *   - carrier: only the high words of U and dU are used; sample i is low by less than i+1 units,
 *     the bias tc_carr = thr_carr + 511 R covers it;
 *   - code: the fraction F (2^-32 half-chip) is stepped by dF = dH >> 19, same argument, bias in HA/HB;
 *   - a run is ambiguous when the smallest biased fraction it saw is below lim = tc + thr + 1
 *     (the serial value lies in [biased - tc - thr, biased)).
 * A code wrap inside the run only changes how the 16-field window is assembled.
 * Returns 0: terms added, final.  1: terms added but some sample is ambiguous (the caller takes them
 * back out and uses the generic form).  2: nothing added, generic form needed.                      */
/* Shared-memory operands of the sample loop.  On the device the tables are addressed as 32-bit
 * shared-window addresses and read with ld.shared spelled out: through generic pointers the compiler
 * may carry the window base in a register and add it to every lookup address (it did, in the
 * two-team kernel: one extra add per channel-sample).  On the host they are plain pointers. */
#if defined(__CUDA_ARCH__)
typedef uint32_t e1_sptr;
static __device__ __forceinline__ e1_sptr e1_sp(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
static __device__ __forceinline__ uint32_t e1_ld32(e1_sptr a)
{
    uint32_t v;
    asm("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(a));
    return v;
}
#else
#if defined(E1_CHECK_LUT_BOUNDS)
static unsigned long long e1_lut_oob = 0;
static const unsigned char *e1_lut_base = 0; /* carrier table of the running test (lane copy 0): address-walking loops check against it */
#endif
typedef const unsigned char *e1_sptr;
static inline e1_sptr e1_sp(const void *p) { return (const unsigned char *)p; }
static inline uint32_t e1_ld32(e1_sptr a) { return *(const uint32_t *)a; }
#endif

/* Table position of the first sample of a run and its per-sample step, from the carrier magnitude U
 * (2^-64 cycle) and its step dU.  y_i = 511 * (uh + i*duh) + tc with the 32-bit truncations uh, duh:
 * index i in the high word, its fraction in the low word; stepped from the run's start value without
 * reducing uh + i*duh mod 2^32 when the phase wraps inside the run -- the table is laid out for that
 * (E1C_LUT_IDX; e1_make_par sends steps too large for its extension to the generic form).  A magnitude
 * that shrinks towards zero has dU = -step (it cannot reach zero inside a run that gets here) and
 * never wraps.  The loop steps the table POSITION, entry in the high word:
 *   phase >= 0:  EXT + i (- 511 when the run may wrap)                  = y + EXT 2^32 - wrap
 *   phase <  0:  EXT + 512 - i (+ 511 ...): (513 + EXT) 2^32 - 1 - y + wrap has exactly that high word
 *                and the COMPLEMENT of the fraction below it; adding lim_carr carries into the high
 *                word iff fraction < lim_carr, so "ambiguous" reads "low word < lim_carr" here too
 *                (the entry is then one too high, in a run that is redone anyway).               */
E1_HD uint64_t e1_carrier_start(uint64_t U, uint64_t dU, uint32_t neg, uint32_t tc_carr, uint32_t lim_carr, int64_t *D_out)
{
    const uint64_t y0 = (uint64_t)(uint32_t)(U >> 32) * 511u + tc_carr;
    const int64_t D = (int64_t)(int32_t)(uint32_t)(dU >> 32) * 511; /* floor of the signed step, times 511 */
    const uint64_t wrap = (D >= 0 && (uint32_t)(y0 >> 32) >= 511u - E1C_LUT_EXT) ? (511ull << 32) : 0ull;
    if (!neg) {
        *D_out = D;
        return y0 + ((uint64_t)E1C_LUT_EXT << 32) - wrap;
    }
    *D_out = -D;
    return (((uint64_t)(513 + E1C_LUT_EXT) << 32) - 1ull) - y0 + wrap + lim_carr;
}

/* The sample loop proper: R samples from table position y (step D), code fraction F (step dF) and the
 * window of signed chip fields win.  Returns the ambiguity flag of the run. */
template <int R, bool CHECK = true>
E1_HD uint32_t e1_sample_loop(uint64_t y, int64_t D, e1_sptr lut_lane, uint32_t F, uint32_t dF, uint32_t win, int *acc,
                              uint32_t lim_carr, uint32_t lim_code)
{
    uint32_t mY = 0xffffffffu, mF = 0xffffffffu;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < R; i++) {
        if (CHECK) { /* off for the runs of an E1_PAR_CLEAN tile: none of them can be ambiguous */
            mY = (uint32_t)y < mY ? (uint32_t)y : mY;
            mF = F < mF ? F : mF;
        }
#if defined(__CUDA_ARCH__)
        const int w = (int)e1_ld32(e1_mad_u32((uint32_t)(y >> 32), 4u * E1C_LUT_REP, lut_lane));
#else
#if defined(E1_CHECK_LUT_BOUNDS) /* host test build: every lookup of the fast form must stay inside the table */
        if ((uint32_t)(y >> 32) >= (uint32_t)E1C_LUT_IDX)
            e1_lut_oob++;
#endif
        const int w = (int)e1_ld32(lut_lane + (uint32_t)(y >> 32) * (4u * E1C_LUT_REP));
#endif
        acc[i] += w * ((int)win >> 30);
        y += (uint64_t)D;
        /* F += dF; a carry = the next half-chip: the window moves up by one field.  On the device the
           add is spelled with its carry flag so it is ONE add on the ALU pipe whose carry-out
           predicates the shift (instead of an add on the multiplier pipe plus a compare). */
#if defined(__CUDA_ARCH__)
        uint32_t cy;
        asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, 0, 0;" : "+r"(F), "=r"(cy) : "r"(dF));
        if (cy)
            win <<= 2;
#else
        const uint32_t F2 = F + dF;
        if (F2 < F)
            win <<= 2;
        F = F2;
#endif
    }
    return CHECK ? (uint32_t)(mY < lim_carr) | (uint32_t)(mF < lim_code) : 0u;
}

/* The 16-field code window of the run that starts at tile sample j (code phase H, 2^-51 half-chip, bias
 * included): half-chips h0..h0+15 with h0 on top, symbol pattern applied, fields turned into the signed
 * chip value.  A code wrap inside the run (:491-494) only changes how the window is assembled:
 * half-chip 8184 is half-chip 0 of the next code period, under the next symbol.  The fraction keeps
 * running from the pre-wrap checkpoint; what that costs in accuracy is inside tc_code (e1_tc_code). */
E1_HD uint32_t e1_code_window(const e1_chan_par *p, e1_sptr code, uint64_t H, int j, int R, int jw)
{
    const uint32_t h0 = (uint32_t)(H >> 51);
    const e1_sptr cw = code + 4u * (h0 >> 4);
    uint32_t win = e1_funnel_l(e1_ld32(cw + 4u), e1_ld32(cw), 2u * (h0 & 15u));
    win ^= j >= jw ? p->pat_b : p->pat_a;
    if (j < jw && jw < j + R) {
        int k0 = 2 * E1C_CODE_LEN - (int)h0; /* window position of half-chip 0 */
        k0 = k0 < 0 ? 0 : (k0 > 16 ? 16 : k0);
        const uint32_t keep = k0 >= 16 ? 0xffffffffu : ~(0xffffffffu >> (2 * k0));
        const uint32_t wb = (k0 >= 16 ? 0u : (e1_ld32(code) >> (2 * k0))) ^ p->pat_b;
        win = (win & keep) | (wb & ~keep);
    }
    return win & ((win << 1) | 0x55555555u); /* fields are now y - x in two's complement */
}

template <int R>
E1_HD uint32_t e1_run_fast(const e1_chan_par *p, const uint32_t *codes, const unsigned char *lut_lane, int j0, int *acc,
                           uint32_t tc_carr, uint32_t lim_carr, uint32_t lim_code)
{
    const int jw = p->j_w;
    const uint32_t misc = p->misc;
    if (misc & E1_PAR_FORCE)
        return 2u;
    const int after = j0 >= jw;
    const uint64_t H = (after ? p->HB : p->HA) + (uint64_t)(uint32_t)j0 * p->dH;
    const uint32_t F = (uint32_t)(H >> 19);
    const uint32_t win = e1_code_window(p, e1_sp(codes) + 4u * p->code_off, H, j0, R, jw);
    uint64_t U = p->U0 + (uint64_t)(uint32_t)j0 * p->dU, dU = p->dU;
    uint32_t neg = misc & E1_PAR_NEG;
    if (misc & E1_PAR_HASZ) { /* rare: the phase changes sign inside this tile */
        const int jz = (int)(misc >> 16);
        /* the run that contains the crossing, and the one that ends on the last sample before it
           (the magnitude there can be smaller than the stepping error below): generic form */
        if (j0 < jz && jz <= j0 + R)
            return 2u;
        if (j0 >= jz) {
            U = 0ull - U;
            dU = 0ull - dU;
            neg ^= E1_PAR_NEG;
        }
    }
    int64_t D;
    const uint64_t y = e1_carrier_start(U, dU, neg, tc_carr, lim_carr, &D);
    return e1_sample_loop<R>(y, D, e1_sp(lut_lane), F, p->dF, win, acc, lim_carr, lim_code);
}

/* ---- several runs from one carrier start, the table walked by carries (E1_PAR_SLOW tiles) -------------------------------------
 * The table position of e1_sample_loop is a 64-bit number, entry in the high word, index fraction in the low
 * word, stepped by D per sample.  When |D| < 2^32 (less than one entry per sample) the high word changes by at
 * most one: the loop keeps the fraction and the table ADDRESS instead, adds the low word of D to the fraction
 * with its carry flag, and moves the address by one entry (128 bytes) on a carry (D >= 0) or on a missing carry
 * (D < 0, i.e. a borrow) -- the same integers as the 64-bit add, one instruction per sample less (no
 * multiply-add from entry number to address).  At most NH R <= E1C_LUT_EXT entries are walked from one start, so
 * ONE carrier start serves all NH runs of a thread (bias and limit for NH R samples).  The code side is what
 * e1_run_fast does, run by run.  Returns 1 when some sample is ambiguous (CHECK): the caller drops the terms and
 * redoes the samples with the generic form (e1_cw_rest_impl). */

/* Sample loop with the table walked by carries: see e1_cw_step below.  `one` is the value 1 in a register ptxas
 * cannot see through (the kernel derives it from a launch argument): the address step is spelled
 * mad.lo(one, +-128, addr) so that it is issued to the multiplier pipe.  The loop is bound by its two integer pipes
 * (each takes an instruction every other clock): with the step as a plain add the ALU pipe carries four of the six
 * arithmetic instructions per sample, this way three and three. */
/* e1_carrier_start at full precision, for the carry-walked pairs: table position of the run's first sample from
 * floor(511 U / 2^32) (U: |phase|, 2^-64 cycle), entry in the high word, fraction in the low word.  up = the
 * unmirrored position increases (only then can a carrier wrap fall into the run).  Same layout rules as
 * e1_carrier_start. */
E1_HD uint64_t e1_carrier_start_p(uint64_t U, uint32_t neg, uint32_t up, uint32_t tc_carr, uint32_t lim_carr, uint32_t ext = E1C_LUT_EXT)
{
    const uint64_t lo = (U & 0xffffffffull) * 511ull;
    const uint64_t y0 = (U >> 32) * 511ull + (lo >> 32) + tc_carr;
    const uint64_t wrap = (up && (uint32_t)(y0 >> 32) >= 511u - ext) ? (511ull << 32) : 0ull;
    if (!neg)
        return y0 + ((uint64_t)ext << 32) - wrap;
    return (((uint64_t)(513u + ext) << 32) - 1ull) - y0 + wrap + lim_carr;
}

/* One step of the carry-walked sample loop (see e1_sample_loop_cw), on one run's state. */
template <bool CHECK, bool DOWN>
E1_HD void e1_cw_step(uint32_t &ylo, e1_sptr &addr, uint32_t Dlo, uint32_t &F, uint32_t dF, uint32_t &win, int &acc, uint32_t &mY, uint32_t &mF,
                      uint32_t one)
{
    if (CHECK) {
        mY = ylo < mY ? ylo : mY;
        mF = F < mF ? F : mF;
    }
#if !defined(__CUDA_ARCH__) && defined(E1_CHECK_LUT_BOUNDS)
    if ((size_t)(addr - e1_lut_base) >= (size_t)E1C_LUT_IDX * 4u * E1C_LUT_REP)
        e1_lut_oob++;
#endif
    const int w = (int)e1_ld32(addr);
    acc += w * ((int)win >> 30);
#if defined(__CUDA_ARCH__)
    uint32_t cy, cz;
    asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, 0, 0;" : "+r"(ylo), "=r"(cy) : "r"(Dlo));
    if (DOWN) {
        if (!cy)
            asm("mad.lo.u32 %0, %1, -128, %0;" : "+r"(addr) : "r"(one));
    } else {
        if (cy)
            asm("mad.lo.u32 %0, %1, 128, %0;" : "+r"(addr) : "r"(one));
    }
    asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, 0, 0;" : "+r"(F), "=r"(cz) : "r"(dF));
    if (cz)
        win <<= 2;
#else
    (void)one;
    const uint32_t y2 = ylo + Dlo;
    if (DOWN) {
        if (!(y2 < ylo))
            addr -= 4u * E1C_LUT_REP;
    } else {
        if (y2 < ylo)
            addr += 4u * E1C_LUT_REP;
    }
    ylo = y2;
    const uint32_t F2 = F + dF;
    if (F2 < F)
        win <<= 2;
    F = F2;
#endif
}

/* NH consecutive runs of R = 16 samples (NH = 2: 32 samples per thread, NH = 4: 64) from ONE carrier start: every
 * run's table position is the start's plus a multiple of 16 D (|D| < 1 entry per sample and NH R <= E1C_LUT_EXT keep all of
 * them inside the table's extension); the code side is set up run by run as in e1_run_fast.  All set-up comes
 * first, then ONE loop of R iterations steps the NH runs side by side: NH independent dependency chains for the
 * scheduler instead of one of NH R steps.  codes_s / lut_s: the code words and this lane's copy of the carrier table
 * as shared-memory operands (e1_sptr).  Returns 1 when some sample is ambiguous. */
template <int NH, bool CHECK, bool DOWN>
E1_HD uint32_t e1_run_cw(const e1_chan_par *p, e1_sptr codes_s, e1_sptr lut_s, int j0, int *acc, uint32_t tc_carr, uint32_t lim_carr,
                         uint32_t lim_code, uint32_t one)
{
    const int R = E1C_MAX_RUN;
    const int jw = p->j_w;
    const uint64_t dH = p->dH;
    const uint32_t dF = p->dF;
    const e1_sptr code = codes_s + 4u * p->code_off;
    const uint32_t neg = p->misc & E1_PAR_NEG;
    const uint64_t U = p->U0 + (uint64_t)(uint32_t)j0 * p->dU;
    /* the mirrored walk goes down when the unmirrored one goes up */
    const uint64_t y = e1_carrier_start_p(U, neg, neg ? (DOWN || p->Dlo == 0u) : !DOWN, tc_carr, lim_carr);
    const uint32_t Dlo = p->Dlo; /* |D| < 2^32 and sign(D) = DOWN (e1_make_par): the low word says it all */
    const uint64_t D16 = ((uint64_t)Dlo << 4) - (DOWN ? (1ull << 36) : 0ull); /* 16 D as a 64-bit two's complement number */
    uint32_t ylo[NH], F[NH], win[NH];
    e1_sptr addr[NH];
    uint32_t mY = 0xffffffffu, mF = 0xffffffffu;
    int after_prev = j0 >= jw;
    uint64_t H = (after_prev ? p->HB : p->HA) + (uint64_t)(uint32_t)j0 * dH;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int h = 0; h < NH; h++) {
        const int jh = j0 + h * R;
        if (h > 0) { /* a run continues the previous one's code phase unless the code wrapped in between */
            const int after = jh >= jw;
            H = after != after_prev ? p->HB + (uint64_t)(uint32_t)jh * dH : H + (uint64_t)R * dH;
            after_prev = after;
        }
        win[h] = e1_code_window(p, code, H, jh, R, jw);
        F[h] = (uint32_t)(H >> 19);
        const uint64_t yh = y + (uint64_t)h * D16;
        ylo[h] = (uint32_t)yh;
        addr[h] = lut_s + (uint32_t)(yh >> 32) * (4u * E1C_LUT_REP);
    }
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < R; i++) {
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
        for (int h = 0; h < NH; h++)
            e1_cw_step<CHECK, DOWN>(ylo[h], addr[h], Dlo, F[h], dF, win[h], acc[h * R + i], mY, mF, one);
    }
    return CHECK ? (uint32_t)(mY < lim_carr) | (uint32_t)(mF < lim_code) : 0u;
}

/* One channel's contribution to the NH R samples of a thread: the INLINE part.  Tiles that are E1_PAR_SLOW and
 * E1_PAR_CLEAN -- 99 % of them at the Dopplers of a terrestrial receiver -- are walked by carries without tracking
 * (two instances: position up / down), their terms go to acc and are final: returns 0.  Everything else returns a
 * code for the out-of-line part (e1_cw_rest_impl) and adds nothing, which keeps the hot loop's code small:
 *   E1_RC_CHECK  E1_PAR_SLOW tile with a sample near an index boundary: the same walk with the tracking on
 *   E1_RC_OLD    not E1_PAR_SLOW (fast carrier, forced generic form, zero crossing): runs of R in the 64-bit form */
#define E1_RC_CHECK 16u
#define E1_RC_OLD 32u
template <int NH>
E1_HD uint32_t e1_cw_add(const e1_chan_par *p, e1_sptr codes_s, e1_sptr lut_s, int j0, int *acc, uint32_t thr_carr, uint32_t lim_code,
                         uint32_t one)
{
    const uint32_t misc = p->misc;
    if ((misc & (E1_PAR_SLOW | E1_PAR_CLEAN)) != (E1_PAR_SLOW | E1_PAR_CLEAN))
        return (misc & E1_PAR_SLOW) ? E1_RC_CHECK : E1_RC_OLD;
    const uint32_t tc = e1_tc_carr_cw(thr_carr, NH * E1C_MAX_RUN), lim = e1_lim_carr_cw(thr_carr, NH * E1C_MAX_RUN);
    if (misc & E1_PAR_DOWN)
        e1_run_cw<NH, false, true>(p, codes_s, lut_s, j0, acc, tc, lim, lim_code, one);
    else
        e1_run_cw<NH, false, false>(p, codes_s, lut_s, j0, acc, tc, lim, lim_code, one);
    return 0u;
}

/* The out-of-line part: d[0 .. NH R) receives what has to be ADDED to the accumulators.
 *   E1_RC_CHECK  the carry walk with the tracking on; if it flags a sample, the generic form of all NH R samples
 *   E1_RC_OLD    the samples as runs of R in the 64-bit-position form (e1_run_fast, tracking on), each flagged or
 *                unhandled run replaced by the generic form
 * check_clean (host test build): a tile marked clean is walked with the tracking on as well and must not be flagged. */
template <int NH>
E1_HD void e1_cw_rest_impl(const e1_chan_par *p, const uint32_t *codes, const unsigned char *lut_lane, int j0, int *d, uint32_t rc,
                           uint32_t thr_carr, uint32_t thr_code, uint32_t tc_carr, uint32_t tc_code, unsigned long long *n_exact,
                           unsigned long long *n_slow)
{
    const int R = E1C_MAX_RUN;
    const uint32_t lim_carr = e1_lim_carr(tc_carr, thr_carr), lim_code = e1_lim_code(tc_code, thr_code);
    for (int i = 0; i < NH * R; i++)
        d[i] = 0;
    if (rc == E1_RC_CHECK) {
        const uint32_t tc = e1_tc_carr_cw(thr_carr, NH * R), lim = e1_lim_carr_cw(thr_carr, NH * R);
        const uint32_t flagged = (p->misc & E1_PAR_DOWN) ? e1_run_cw<NH, true, true>(p, e1_sp(codes), e1_sp(lut_lane), j0, d, tc, lim, lim_code, 1u)
                                                         : e1_run_cw<NH, true, false>(p, e1_sp(codes), e1_sp(lut_lane), j0, d, tc, lim, lim_code, 1u);
        if (flagged) { /* the walk's terms are dropped: the generic form decides every sample */
            e1_channel_run(p, codes, lut_lane, j0, NH * R, d, thr_carr, thr_code, e1_bias_h(tc_code), n_exact);
            if (n_slow)
                (*n_slow)++;
        }
        return;
    }
    for (int h = 0; h < NH; h++) {
        const uint32_t rc2 = e1_run_fast<E1C_MAX_RUN>(p, codes, lut_lane, j0 + h * R, d + h * R, tc_carr, lim_carr, lim_code);
        if (rc2) {
            int t[E1C_MAX_RUN], g[E1C_MAX_RUN];
            for (int i = 0; i < R; i++)
                t[i] = 0;
            e1_run_fast<E1C_MAX_RUN>(p, codes, lut_lane, j0 + h * R, t, tc_carr, lim_carr, lim_code);
            e1_channel_run(p, codes, lut_lane, j0 + h * R, R, g, thr_carr, thr_code, e1_bias_h(tc_code), n_exact);
            for (int i = 0; i < R; i++)
                d[h * R + i] += g[i] - t[i];
            if (n_slow)
                (*n_slow)++;
        }
    }
}

/* ---- event-driven runs (E1_PAR_EV; sample rates where a half-chip lasts several samples) ---------------------
 * At 25 MS/s a BOC(1,1) half-chip lasts 12.2 samples and, at 4 kHz of Doppler, so does an entry of the carrier
 * table: the term  m x (cos, sin)  a channel adds to the stream (src/galileo-sdr.cpp:509-525) is piecewise constant and
 * the per-sample loop recomputes the same product five times out of six.  Here a thread owns E1C_EV_RUN consecutive
 * samples of a tile as a column of DIFFERENCES in shared memory; per channel it adds the term of its first sample
 * and, at every sample where the term changes, the change; after the last channel ONE running sum over the column
 * gives the samples (integer additions are associative: the same sums as the per-sample loop, bit for bit).
 *
 * The change points are the CARRIES of the two 32-bit index fractions the carry-walked loop (e1_run_cw) steps:
 *   code     F += dF, a carry = the next half-chip (the window moves up by one field)
 *   carrier  G += dG, a carry = the next table entry (G = the position's low word walking up, its complement walking
 *            down, dG = |D|: the borrow of the low word is the carry of its complement)
 * -- the same integers as that loop, so everything proved for it (bias, limits, E1_PAR_CLEAN) carries over; the code
 * fraction is stepped from ONE start per E1C_EV_RUN samples instead of one per 16, which the bias and the limits of an
 * event-driven context are sized for (tc_code of a run of E1C_EV_RUN, e1_par_clean's code_run).  Between carries
 * nothing is computed: the first carry of a run comes from one reciprocal multiply (e1_ev_first), the following ones
 * are n0 or n0 + 1 steps apart (n0 = floor(2^32 / step): one add of n0 x step and its carry flag say which).  The
 * two event streams run as two independent loops: at a code event the table entry is evaluated directly (number of
 * carrier carries so far = high word of G_a + n dG), at a carrier event the chip field likewise (high word of
 * F_a + (n - 1) dF: the chip of the sample BEFORE, so that coinciding events add up to new - old exactly).
 * A code wrap inside the thread's samples splits them into two sub-ranges (HA / pat_a before, HB / pat_b after).   */
#if defined(__CUDA_ARCH__)
typedef uint32_t e1_dptr; /* this thread's column: shared-memory address of entry 0, entries 128 bytes apart */
static __device__ __forceinline__ void e1_diff_add(e1_dptr d, int k, int v)
{
    asm volatile("red.shared.add.u32 [%0], %1;" : : "r"(d + 128u * (uint32_t)k), "r"(v) : "memory");
}
#else
typedef int *e1_dptr;
static inline void e1_diff_add(e1_dptr d, int k, int v) { d[k] += v; }
#endif

/* a code word of the event-driven kernel: it reads them from GLOBAL memory (read-only path, L1 / L2: two words per
   thread and channel), which leaves the shared memory to the columns -- five tiles in flight per SM instead of three */
E1_HD uint32_t e1_ldg32(const uint32_t *p)
{
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}

E1_HD uint32_t e1_hi_mad(uint32_t a, uint32_t b, uint32_t c) /* high word of a * b + c */
{
    return (uint32_t)(((uint64_t)a * b + c) >> 32);
}

/* Smallest n >= 1 with g + n d >= 2^32 (the first carry of the accumulator g stepped by d < 2^30), or a number above
 * E1C_EV_RUN when there is none within reach.  rcp = 1.0f / (float)d (+inf for d = 0): the float quotient is within one
 * of floor((2^32 - 1 - g) / d) as long as that is small, and one remainder settles it (mod 2^32: the true remainder of
 * the estimate lies in [-d, 2 d), and 3 d < 2^32 keeps "negative" and "d or more" apart). */
E1_HD uint32_t e1_ev_first(uint32_t g, uint32_t d, float rcp)
{
    const uint32_t x = ~g;
    const float qf = (float)x * rcp;
    if (!(qf < (float)(E1C_EV_RUN + 2)))
        return 0xffffu;
    uint32_t q = (uint32_t)qf;
    const uint32_t r = x - q * d;
    if (r >= d)
        q += r >= 0u - d ? 0xffffffffu : 1u;
    return q + 1u;
}

/* F += E0 (n0 steps at once); a carry: the next event is n0 samples on; none: one more step, n0 + 1.  n01 = n0 + 1. */
E1_HD void e1_ev_next(uint32_t &F, uint32_t &krel, uint32_t E0, uint32_t dF, uint32_t n01)
{
#if defined(__CUDA_ARCH__)
    uint32_t cy;
    asm("add.cc.u32 %0, %0, %2;\n\taddc.u32 %1, 0, 0;" : "+r"(F), "=r"(cy) : "r"(E0));
    krel += n01 - cy;
    if (!cy)
        F += dF;
#else
    const uint32_t F2 = F + E0;
    if (F2 < F) {
        F = F2;
        krel += n01 - 1u;
    } else {
        F = F2 + dF;
        krel += n01;
    }
#endif
}

/* One sub-range [ka, kb) of a thread's samples (tile-relative; the thread's first sample is j0): adds the channel's
 * differences to the column (CHECK: scale x them).  after: the sub-range lies behind the tile's code wrap.  close: the term
 * of sample kb - 1 is taken back out at kb (the next sub-range starts from nothing).  Returns 1 when CHECK and some
 * sample's biased fraction is below its limit (the caller takes the differences back out: scale = -1).
 * Both event loops run a trip count that is the same for every thread working on the channel (the most carries
 * E1C_EV_RUN - 1 steps can produce) and are straight-line inside: an iteration past the thread's last event adds zero to
 * the column's last entry.  lut1_s: the single-copy carrier table and its two difference tables (e1_build_lut1). */
#define E1C_LUT1_WORDS ((E1C_LUT1_IDX + 3) / 4 * 4)
template <bool CHECK>
E1_HD uint32_t e1_ev_sub(const e1_chan_par *p, const uint32_t *codes, e1_sptr lut1_s, int ka, int kb, int j0, int after, int close, e1_dptr diff,
                         int scale, uint32_t tc_carr, uint32_t lim_carr, uint32_t lim_code)
{
    const uint32_t misc = p->misc, dF = p->dF, Dlo = p->Dlo;
    const uint32_t neg = misc & E1_PAR_NEG, down = misc & E1_PAR_DOWN;
    const uint32_t n0c = p->ev_n0 & 0xffffu, n0r = p->ev_n0 >> 16;
    const uint32_t kn = (uint32_t)(kb - ka); /* samples of the sub-range */
    /* code side at ka */
    const uint64_t H = (after ? p->HB : p->HA) + (uint64_t)(uint32_t)ka * p->dH;
    const uint32_t Fa = (uint32_t)(H >> 19), h0 = (uint32_t)(H >> 51);
    const uint32_t *cw = codes + p->code_off + (h0 >> 4);
    uint32_t win = e1_funnel_l(e1_ldg32(cw + 1), e1_ldg32(cw), 2u * (h0 & 15u)) ^ (after ? p->pat_b : p->pat_a);
    win &= (win << 1) | 0x55555555u; /* fields are now y - x in two's complement (e1_code_window) */
    /* carrier side at ka: the table position of e1_run_cw, one start for the whole sub-range */
    const uint64_t U = p->U0 + (uint64_t)(uint32_t)ka * p->dU;
    const uint64_t y = e1_carrier_start_p(U, neg, neg ? (down || Dlo == 0u) : !down, tc_carr, lim_carr, E1C_EV_EXT);
    const uint32_t Ga = down ? ~(uint32_t)y : (uint32_t)y, dG = down ? 0u - Dlo : Dlo;
    const int sstep = down ? -4 : 4; /* bytes per entry of the single-copy table */
    const e1_sptr addr_a = lut1_s + 4u * (uint32_t)(y >> 32);
    /* walking up entry E takes over from E - 1, walking down from E + 1: one difference table each */
    const e1_sptr daddr_a = addr_a + (down ? 8u : 4u) * E1C_LUT1_WORDS;
#if !defined(__CUDA_ARCH__) && defined(E1_CHECK_LUT_BOUNDS)
    {
        const int64_t last = (int64_t)(uint32_t)(y >> 32) + (down ? -1 : 1) * (int64_t)e1_hi_mad(kn - 1u, dG, Ga);
        if ((uint32_t)(y >> 32) >= (uint32_t)E1C_LUT1_IDX || last < 0 || last >= E1C_LUT1_IDX)
            e1_lut_oob++;
    }
#endif
    uint32_t mF = Fa, mY = (uint32_t)y;
    if (CHECK && down) /* walking down the low word shrinks until it borrows: the last sample's is a minimum too */
        mY = ~(Ga + (kn - 1u) * dG);
    int sgn = (int)win >> 30;
    {
        const int w0 = (int)e1_ld32(addr_a);
        e1_diff_add(diff, ka - j0, CHECK ? scale * sgn * w0 : sgn * w0);
    }
    const int kbase = ka - j0;
    { /* code events: the chip changes, the table entry is whatever it is at that sample */
        const uint32_t nmax = e1_hi_mad((uint32_t)(E1C_EV_RUN - 1), dF, 0xffffffffu);
        const uint32_t E0 = n0c * dF, n01 = n0c + 1u;
        uint32_t krel = e1_ev_first(Fa, dF, e1_bits_float(p->ev_rcp_f));
        uint32_t F = Fa + krel * dF, wn = win;
        for (uint32_t i = 0; i < nmax; i++) {
            const uint32_t valid = krel < kn, kc = valid ? krel : kn - 1u;
            if (CHECK && valid)
                mF = F < mF ? F : mF;
            wn <<= 2;
            const int s2 = (int)wn >> 30;
            const int w = (int)e1_ld32(addr_a + sstep * (int)e1_hi_mad(kc, dG, Ga));
            int d = valid ? (s2 - sgn) * w : 0;
            if (CHECK)
                d *= scale;
            e1_diff_add(diff, kbase + (int)kc, d);
            sgn = valid ? s2 : sgn;
            e1_ev_next(F, krel, E0, dF, n01);
        }
    }
    { /* carrier events: the table entry changes under the chip of the sample before */
        const uint32_t nmax = e1_hi_mad((uint32_t)(E1C_EV_RUN - 1), dG, 0xffffffffu);
        const uint32_t E0 = n0r * dG, n01 = n0r + 1u;
        uint32_t krel = e1_ev_first(Ga, dG, e1_bits_float(p->ev_rcp_g));
        uint32_t G = Ga + krel * dG;
        for (uint32_t i = 0; i < nmax; i++) {
            const uint32_t valid = krel < kn, kc = valid ? krel : kn - 1u;
            if (CHECK && valid) { /* low word at this sample (up: G) / at the sample before (down: the complement of G - dG) */
                const uint32_t lw = down ? dG - 1u - G : G;
                mY = lw < mY ? lw : mY;
            }
            const int dw = (int)e1_ld32(daddr_a + sstep * (int)e1_hi_mad(kc, dG, Ga));
            const uint32_t idx = e1_hi_mad(kc - 1u, dF, Fa) & 15u; /* (the mask only matters in an iteration that adds nothing) */
            const int s = (int)(win << (2u * idx)) >> 30;
            int d = valid ? s * dw : 0;
            if (CHECK)
                d *= scale;
            e1_diff_add(diff, kbase + (int)kc, d);
            e1_ev_next(G, krel, E0, dG, n01);
        }
    }
    if (close) { /* the sub-range's last term, evaluated directly */
        const int w = (int)e1_ld32(addr_a + sstep * (int)e1_hi_mad(kn - 1u, dG, Ga));
        const int d = sgn * w;
        e1_diff_add(diff, kb - j0, CHECK ? -(scale * d) : -d);
    }
    return CHECK ? (uint32_t)(mY < lim_carr) | (uint32_t)(mF < lim_code) : 0u;
}

/* krel += the distance to the next carry of an accumulator stepped by d, n0 or n0 + 1 steps.  The accumulator's value
 * right after a carry, F in [0, d), is kept as its complement Fc = d - 1 - F.  E1 = (n0 + 1) d - 2^32 in (0, d): n0 + 1
 * steps on the accumulator reads F + E1 < 2 d; below d the carry came with that last step, otherwise one step earlier
 * and the value there was d less.  On the complement: Fc - E1 does not borrow <=> n0 + 1 steps.  (sub.cc leaves the
 * hardware carry, 1 = no borrow, in CC.CF.) */
E1_HD void e1_ev_step(uint32_t &Fc, uint32_t &krel, uint32_t E1, uint32_t d, uint32_t n0)
{
#if defined(__CUDA_ARCH__)
    asm("{\n\t.reg .u32 t;\n\t"
        "sub.cc.u32 %0, %0, %2;\n\t"
        "addc.u32 %1, %1, %4;\n\t"
        "add.u32 t, %0, %3;\n\t"
        "min.u32 %0, %0, t;\n\t}"
        : "+r"(Fc), "+r"(krel)
        : "r"(E1), "r"(d), "r"(n0));
#else
    if (Fc >= E1) {
        Fc -= E1;
        krel += n0 + 1u;
    } else {
        Fc = Fc - E1 + d;
        krel += n0;
    }
#endif
}

/* The common case of e1_ev_sub, lean: all E1C_EV_RUN samples of the thread (first one: tile sample j0), no code wrap among
 * them, clean tile (no tracking).  E1C_EV_RUN - 1 steps from any start produce nmax or nmax - 1 carries (nmax = the most
 * they can): all iterations but the last are events of every thread and run unguarded, only the last one is clamped
 * and zeroed where the thread has no event left.
 * lut1_s: the carrier table, E1C_LUT1_WORDS words, then the difference table walking up (entry E = table[E] -
 * table[E - 1]) and the one walking down (table[E] - table[E + 1]). */
/* The two code words a thread's window is cut from (e1_ev_run64) and H, the thread's code phase at its first sample.
 * (Fetching them one channel ahead was tried: the extra live registers cost more than the load latency, 6.26 -> 6.48 ms.) */
E1_HD void e1_ev_fetch(const e1_chan_par *p, const uint32_t *codes, int j0, uint64_t *H, uint32_t *w0, uint32_t *w1)
{
    *H = (j0 >= p->j_w ? p->HB : p->HA) + (uint64_t)(uint32_t)j0 * p->dH;
    const uint32_t *cw = codes + p->code_off + (((uint32_t)(*H >> 51) >> 4) & 511u);
    *w0 = e1_ldg32(cw);
    *w1 = e1_ldg32(cw + 1);
}

#if !defined(E1_EV_UNROLL)
#define E1_EV_UNROLL 4
#endif
#define E1_PRAGMA_(x) _Pragma(#x)
#if defined(__CUDA_ARCH__)
#define E1_UNROLL(n) E1_PRAGMA_(unroll n)
#else
#define E1_UNROLL(n)
#endif
E1_HD void e1_ev_run64(const e1_chan_par *p, uint64_t H, uint32_t cw0, uint32_t cw1, e1_sptr lut1_s, int j0, e1_dptr diff, uint32_t tc_carr,
                       uint32_t lim_carr)
{
    const uint32_t misc = p->misc, dF = p->dF, Dlo = p->Dlo;
    const uint32_t neg = misc & E1_PAR_NEG, down = misc & E1_PAR_DOWN;
    const uint32_t n0c = p->ev_n0 & 0xffffu, n0r = p->ev_n0 >> 16;
    const int after = j0 >= p->j_w;
    const uint32_t Fa = (uint32_t)(H >> 19), h0 = (uint32_t)(H >> 51);
    uint32_t win = e1_funnel_l(cw1, cw0, 2u * (h0 & 15u)) ^ (after ? p->pat_b : p->pat_a);
    win &= (win << 1) | 0x55555555u;
    const uint64_t U = p->U0 + (uint64_t)(uint32_t)j0 * p->dU;
    const uint64_t y = e1_carrier_start_p(U, neg, neg ? (down || Dlo == 0u) : !down, tc_carr, lim_carr, E1C_EV_EXT);
    const uint32_t Ga = down ? ~(uint32_t)y : (uint32_t)y, dG = down ? 0u - Dlo : Dlo;
    const int sstep = down ? -4 : 4;
    const e1_sptr addr_a = lut1_s + 4u * (uint32_t)(y >> 32);
    const e1_sptr daddr_a = addr_a + (down ? 8u : 4u) * E1C_LUT1_WORDS;
    const uint32_t last = (uint32_t)(E1C_EV_RUN - 1);
#if !defined(__CUDA_ARCH__) && defined(E1_CHECK_LUT_BOUNDS)
    {
        const int64_t le = (int64_t)(uint32_t)(y >> 32) + (down ? -1 : 1) * (int64_t)e1_hi_mad(last, dG, Ga);
        if ((uint32_t)(y >> 32) >= (uint32_t)E1C_LUT1_IDX || le < 0 || le >= E1C_LUT1_IDX)
            e1_lut_oob++;
    }
#endif
    int sgn = (int)win >> 30;
    e1_diff_add(diff, 0, sgn * (int)e1_ld32(addr_a));
    { /* code events */
        const uint32_t nmax = e1_hi_mad(last, dF, 0xffffffffu);
        const uint32_t E1 = (n0c + 1u) * dF;
        uint32_t krel = e1_ev_first(Fa, dF, e1_bits_float(p->ev_rcp_f));
        uint32_t F = dF - 1u - (Fa + krel * dF), wn = win; /* complement of the post-carry value (e1_ev_step) */
        E1_UNROLL(E1_EV_UNROLL)
        for (uint32_t i = 1; i < nmax; i++) {
#if !defined(__CUDA_ARCH__) && defined(E1_CHECK_LUT_BOUNDS)
            if (krel > last) /* an unguarded iteration without an event: must not happen */
                e1_lut_oob++;
#endif
            wn <<= 2;
            const int s2 = (int)wn >> 30;
            const int w = (int)e1_ld32(addr_a + sstep * (int)e1_hi_mad(krel, dG, Ga));
            e1_diff_add(diff, (int)krel, (s2 - sgn) * w);
            sgn = s2;
            e1_ev_step(F, krel, E1, dF, n0c);
        }
        if (nmax) {
            const uint32_t kc = krel < last ? krel : last;
            const int s2 = (int)(wn << 2) >> 30;
            const int w = (int)e1_ld32(addr_a + sstep * (int)e1_hi_mad(kc, dG, Ga));
            e1_diff_add(diff, (int)kc, krel <= last ? (s2 - sgn) * w : 0);
        }
    }
    { /* carrier events */
        const uint32_t nmax = e1_hi_mad(last, dG, 0xffffffffu);
        const uint32_t E1 = (n0r + 1u) * dG;
        const uint64_t C = (uint64_t)Fa - (uint64_t)dF; /* chip field of the sample before: high word of (k - 1) dF + Fa */
        uint32_t krel = e1_ev_first(Ga, dG, e1_bits_float(p->ev_rcp_g));
        uint32_t G = dG - 1u - (Ga + krel * dG);
        E1_UNROLL(E1_EV_UNROLL)
        for (uint32_t i = 1; i < nmax; i++) {
#if !defined(__CUDA_ARCH__) && defined(E1_CHECK_LUT_BOUNDS)
            if (krel > last)
                e1_lut_oob++;
#endif
            const int dw = (int)e1_ld32(daddr_a + sstep * (int)e1_hi_mad(krel, dG, Ga));
            const uint32_t idx = (uint32_t)(((uint64_t)krel * dF + C) >> 32);
            const int s = (int)(win << (2u * idx)) >> 30;
            e1_diff_add(diff, (int)krel, s * dw);
            e1_ev_step(G, krel, E1, dG, n0r);
        }
        if (nmax) {
            const uint32_t kc = krel < last ? krel : last;
            const int dw = (int)e1_ld32(daddr_a + sstep * (int)e1_hi_mad(kc, dG, Ga));
            const uint32_t idx = (uint32_t)(((uint64_t)kc * dF + C) >> 32) & 15u;
            const int s = (int)(win << (2u * idx)) >> 30;
            e1_diff_add(diff, (int)kc, krel <= last ? s * dw : 0);
        }
    }
}

/* One E1_PAR_EV channel's differences for the n (<= E1C_EV_RUN) samples of a thread that start at tile sample j0. */
template <bool CHECK>
E1_HD uint32_t e1_ev_add(const e1_chan_par *p, const uint32_t *codes, e1_sptr lut1_s, int j0, int n, e1_dptr diff, int scale, uint32_t thr_carr,
                         uint32_t lim_code)
{
    const uint32_t tc = e1_tc_carr_cw(thr_carr, E1C_EV_RUN), lim = e1_lim_carr_cw(thr_carr, E1C_EV_RUN);
    const int jw = p->j_w, kend = j0 + n;
    uint32_t rc = 0;
    int ka = j0;
    do {
        const int kb = (jw > ka && jw < kend) ? jw : kend;
        rc |= e1_ev_sub<CHECK>(p, codes, lut1_s, ka, kb, j0, ka >= jw, kb < kend, diff, scale, tc, lim, lim_code);
        ka = kb;
    } while (ka < kend);
    return rc;
}

/* Everything that does not go through e1_ev_run64, out of line:
 *   E1_PAR_EV, clean       the thread with the tile's code wrap among its samples, the threads of an epoch's last,
 *                          shorter tile: the general form of the events (e1_ev_sub)
 *   E1_PAR_EV, not clean   (about 1 % of the (tile, channel) sets) the same events with the tracking on; a flagged thread takes its differences back out and
 *   anything else          goes through the generic form (e1_channel_run), 16 samples at a time, term by term
 * codes / lut_lane: the code words and this lane's copy of the replicated carrier table as plain pointers (the kernel
 * passes the arrays in global memory: it keeps neither in shared memory). */
E1_HD void e1_ev_rest_impl(const e1_chan_par *p, e1_sptr lut1_s, const uint32_t *codes, const unsigned char *lut_lane,
                           int j0, int n, e1_dptr diff, uint32_t thr_carr, uint32_t thr_code, uint32_t tc_code, unsigned long long *n_exact,
                           unsigned long long *n_slow)
{
    const uint32_t lim_code = e1_lim_code(tc_code, thr_code);
    if ((p->misc & (E1_PAR_EV | E1_PAR_CLEAN)) == (E1_PAR_EV | E1_PAR_CLEAN)) { /* clean, but a code wrap among the thread's samples or fewer than E1C_EV_RUN of them */
        e1_ev_add<false>(p, codes, lut1_s, j0, n, diff, 1, thr_carr, lim_code);
        return;
    }
    if (p->misc & E1_PAR_EV) {
        if (!e1_ev_add<true>(p, codes, lut1_s, j0, n, diff, 1, thr_carr, lim_code))
            return;
        e1_ev_add<true>(p, codes, lut1_s, j0, n, diff, -1, thr_carr, lim_code);
    }
    if (n_slow)
        (*n_slow)++;
    int prev = 0;
    for (int c = 0; c < n; c += E1C_MAX_RUN) {
        int g[E1C_MAX_RUN];
        const int m = n - c < E1C_MAX_RUN ? n - c : E1C_MAX_RUN;
        e1_channel_run(p, codes, lut_lane, j0 + c, m, g, thr_carr, thr_code, e1_bias_h(tc_code), n_exact);
        for (int i = 0; i < m; i++) {
            e1_diff_add(diff, c + i, g[i] - prev);
            prev = g[i];
        }
    }
}

/* The event-driven kernel serves contexts with 8192-sample tiles whose sample rate puts at least ~5 samples into a
 * half-chip (a thread's E1C_EV_RUN samples must cross fewer than 15 of them, e1_make_par). */
E1_HD int e1_ev_context(double fs_hz, int run) { return run == E1C_MAX_RUN && fs_hz >= 10.0e6; }

/* carrier tables of the event-driven kernel, 3 E1C_LUT1_WORDS words: the table of e1_build_lut, one copy, with a run-out of
   E1C_EV_EXT entries on either side (entry E = e + E1C_EV_EXT, same index map t(e)); behind it the differences walking up,
   table[E] - table[E - 1]; behind those the differences walking down, table[E] - table[E + 1] (a walk only arrives at an
   entry from a neighbour: the entries without one are 0 and never read).  lut: the replicated table (its entries
   E1C_LUT_EXT .. E1C_LUT_EXT + 511 are the reference's 512 values). */
E1_HD void e1_build_lut1(const int32_t *lut, int32_t *lut1)
{
    for (int E = 0; E < 3 * E1C_LUT1_WORDS; E++)
        lut1[E] = 0;
    for (int E = 0; E < E1C_LUT1_IDX; E++) {
        const int e = E - E1C_EV_EXT;
        const int t = e < 0 ? 511 + e : (e <= 512 ? (e & 511) : e - 511);
        lut1[E] = lut[(t + E1C_LUT_EXT) * E1C_LUT_REP];
    }
    for (int E = 0; E < E1C_LUT1_IDX; E++) {
        if (E)
            lut1[E1C_LUT1_WORDS + E] = lut1[E] - lut1[E - 1];
        if (E + 1 < E1C_LUT1_IDX)
            lut1[2 * E1C_LUT1_WORDS + E] = lut1[E] - lut1[E + 1];
    }
}

/* acc = I + 65536*Q  ->  the sink's little-endian (int16 I, int16 Q) pair (:536-537) */
E1_HD uint32_t e1_pack_iq(int acc)
{
    const uint32_t x = (uint32_t)acc;
    return x + ((x & 0x8000u) << 1);
}

/* Host-side table builders (the product's C-ABI calls them at create(); tests/hostsim too).
 * cos512/sin512: the reference's carrier tables (include/constants.h:216-284); b_words/c_words:
 * the E1-B / E1-C primary codes of one PRN, chip j = bit 31-(j%32) of word j/32.              */
E1_HD void e1_build_lut(const int *cos512, const int *sin512, int32_t *lut)
{
    for (int E = 0; E < E1C_LUT_IDX; E++) {
        const int e = E - E1C_LUT_EXT;
        const int t = e < 0 ? 511 + e : (e <= 512 ? (e & 511) : e - 511);
        const int32_t w2 = 2 * (cos512[t] + 65536 * sin512[t]);
        for (int l = 0; l < E1C_LUT_REP; l++)
            lut[E * E1C_LUT_REP + l] = w2;
    }
}
E1_HD void e1_build_code_words(const uint32_t *b_words, const uint32_t *c_words, uint32_t *out)
{
    for (int i = 0; i < E1C_CODE_WORDS_PER_PRN; i++)
        out[i] = 0;
    for (int hh = 0; hh < 2 * E1C_CODE_LEN; hh++) {
        const int c = hh >> 1;
        const uint32_t b = (b_words[c >> 5] >> (31 - (c & 31))) & 1u, q = (c_words[c >> 5] >> (31 - (c & 31))) & 1u;
        /* value = -chip on even half-chips, +chip on odd ones; chip = -1 where the bit is 1 */
        const uint32_t flip = (uint32_t)(hh & 1) ^ 1u, bneg = b ^ flip, cneg = q ^ flip;
        const uint32_t f = (bneg << 1) | (bneg ^ cneg);
        out[hh >> 4] |= f << (30 - (hh & 15) * 2);
    }
}

#endif /* E1_CORE_H */
