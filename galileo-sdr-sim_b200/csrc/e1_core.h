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
        int64_t n = (end - 1 - b2) / d; /* >= 1 because x3 qualified */
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
        int64_t n = (b2 - floor_bits) / d;     /* >= 1 */
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
#define E1C_CODE_WORDS_PER_PRN 256
#define E1C_RUN 4 /* consecutive samples per thread per group -> one 128-bit store */
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

typedef struct e1_chan_par { /* per active channel of the current tile (shared memory) */
    uint64_t U0, dU;        /* |carrier phase| and its per-sample step, 2^-64 cycle   */
    uint64_t H0, dH;        /* code phase and step, 2^-51 half-chip                   */
    uint64_t Hw;            /* code phase at j_w                                      */
    int32_t j_w;
    uint32_t misc;          /* bits 0-3 = ck.sym symbols, bit 4 reflect LUT, bit 5 force
                               exact, bits 8-15 prn-1                                 */
    double phi, sp, cp, sc; /* exact checkpoint for the literal fallback              */
} e1_chan_par;

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
    if (!(cp >= 0.0) || !(sc > 0.0) || !(sc < 2048.0) || ibit < 0 || ibit >= E1C_SYM_PER_PAGE) {
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

/* Tile checkpoint + epoch record -> the per-channel parameters the sample loop reads. */
E1_HD void e1_make_par(const e1_tile_ck *c, const e1_epoch_rec *r, double delt, int tile, e1_chan_par *p)
{
    p->phi = c->phi;
    p->cp = c->cp;
    p->sp = e1_mul(r->f_carr, delt);
    p->sc = e1_mul(r->f_code, delt);
    const int neg = (p->phi < 0.0) || (p->phi == 0.0 && p->sp < 0.0);
    const int aligned = (p->phi == 0.0) || (p->sp == 0.0) || (neg == (p->sp < 0.0));
    p->U0 = e1_to_fixed(p->phi, E1_CARR_FIX);
    uint64_t s = e1_to_fixed(p->sp, E1_CARR_FIX);
    uint32_t misc = (c->sym & 15u) | (neg ? 16u : 0u) | ((uint32_t)(r->prn - 1) << 8);
    if (!(e1_fabs(p->phi) < 1.0) || !(e1_fabs(p->sp) < 0.5))
        misc |= 32u; /* outside the closed form's domain: literal path */
    if (!aligned) {
        /* |phi| shrinks; if it can reach zero inside the tile the sign regime changes
           mid-tile: leave that (rare) tile to the literal path */
        if (e1_fabs(p->phi) <= e1_mul(e1_fabs(p->sp), (double)(tile + 2)))
            misc |= 32u;
        s = 0ull - s;
    }
    p->dU = s;
    p->H0 = e1_to_fixed(p->cp, E1_CODE_FIX);
    p->dH = e1_to_fixed(p->sc, E1_CODE_FIX);
    p->Hw = e1_to_fixed(c->cp_w, E1_CODE_FIX);
    p->j_w = c->j_w;
    p->misc = misc;
}

/* Exact table indices of sample j of the tile by walking the reference recurrences exactly
 * from the tile checkpoint (only for samples the closed form flags as ambiguous). */
E1_HD void e1_exact_indices_impl(const e1_chan_par *p, int j, uint32_t *h_out, uint32_t *it_out)
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
    *h_out = (uint32_t)e1_d2i_rz(e1_mul(cp, 2.0));             /* :512 */
}
#if defined(__CUDA_ARCH__)
static __device__ __noinline__ void e1_exact_indices(const e1_chan_par *p, int j, uint32_t *h_out, uint32_t *it_out)
{
    e1_exact_indices_impl(p, j, h_out, it_out);
}
#else
#define e1_exact_indices e1_exact_indices_impl
#endif

/* One channel's contribution to the E1C_RUN consecutive samples starting at tile-relative j0
 * (src/galileo-sdr.cpp:509-525).  acc[i] accumulates I + 65536*Q.  `codes` is the 2-bit chip
 * table (all PRNs), `lut` the 1024-entry carrier table.  With exact == 0 the return value is
 * nonzero if any sample was ambiguous; with exact != 0 ambiguous samples are resolved by
 * e1_exact_indices and counted in *n_exact. */
E1_HD uint32_t e1_channel_run(const e1_chan_par *p, const uint32_t *codes, const int32_t *lut_base, int j0, int *acc,
                              uint32_t thr_carr, uint32_t thr_code, const int exact, unsigned long long *n_exact)
{
    const uint64_t U0 = p->U0, dU = p->dU, H0 = p->H0, dH = p->dH, Hw = p->Hw;
    const int jw = p->j_w;
    const uint32_t misc = p->misc;
    const uint32_t *code = codes + ((misc >> 8) & 0xffu) * E1C_CODE_WORDS_PER_PRN;
    const int32_t *lut = lut_base + ((misc & 16u) ? 512 : 0);
    const uint32_t force = (misc >> 5) & 1u;
    uint32_t amb = 0;
    uint64_t U = U0 + (uint64_t)(uint32_t)j0 * dU;
#if defined(__CUDA_ARCH__)
#pragma unroll
#endif
    for (int i = 0; i < E1C_RUN; i++) {
        const int j = j0 + i;
        const int after = j >= jw;
        const uint64_t H = (after ? Hw : H0) + (uint64_t)(uint32_t)(after ? j - jw : j) * dH;
        const uint32_t ds = after ? (misc >> 2) & 3u : misc & 3u;
        /* carrier: y = 511*|phi| in 9.32 fixed point */
        const uint32_t lo511 = e1_umulhi((uint32_t)U, 511u);
        const uint64_t y = (uint64_t)(uint32_t)(U >> 32) * 511u + lo511;
        uint32_t it = (uint32_t)(y >> 32);
        const uint32_t yf = (uint32_t)y;
        /* code: h = trunc(2*code_phase) */
        uint32_t h = (uint32_t)(H >> 51);
        const uint32_t hf = (uint32_t)(H >> 19);
        const uint32_t a = (uint32_t)((yf + thr_carr) < 2u * thr_carr + 1u) | (uint32_t)((hf + thr_code) < 2u * thr_code + 1u) | force;
        if (exact) {
            if (a) {
                uint32_t itx;
                e1_exact_indices(p, j, &h, &itx);
                it = (misc & 16u) ? ((0u - itx) & 511u) : itx; /* lut is already reflected */
                (*n_exact)++;
            }
        } else {
            amb |= a;
        }
        const uint32_t c = h >> 1;
        const uint32_t v = code[c >> 4] >> ((c & 15u) * 2u);
        /* sign bits of E1B*data and E1C*secondary; the BOC(1,1) sub-carrier negates even half-chips */
        const uint32_t t = v ^ ds ^ ((h & 1u) ? 0u : 3u);
        const int m = (int)((t >> 1) & 1u) - (int)(t & 1u);
        acc[i] += m * lut[it];
        U += dU;
    }
    return amb;
}

/* acc = I + 65536*Q  ->  the sink's little-endian (int16 I, int16 Q) pair (:536-537) */
E1_HD uint32_t e1_pack_iq(int acc)
{
    const uint32_t x = (uint32_t)acc;
    return x + ((x & 0x8000u) << 1);
}

#endif /* E1_CORE_H */
