/* e1_oracle.c -- see e1_oracle.h.  TEST INFRASTRUCTURE, not product code.
 *
 * A from-scratch restatement of what the reference computes, written from the behavioural
 * spec in SURVEY.md Appendix A.  Every function names the reference lines it follows.
 * Build with -O2 -ffp-contract=off: the reference's phases are plain (unfused) IEEE double
 * operations and a fused multiply-add changes the output.
 */
#include "e1_oracle.h"
#include "../galileo-sdr-sim_b200/data/e1_prn_codes.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

/* ---- a3: carrier tables (include/constants.h:216-284) ------------------------------------
 * The shipped tables are round(250*cos(2*pi*(i+1/2)/512)) (same for sin) except at the four
 * places per table where the exact value is +-105.5 and the shipped entry is +-105.       */
void e1o_carrier_lut(int cos_out[512], int sin_out[512])
{
    static const int cos_fix[4] = {92, 163, 348, 419};
    static const int sin_fix[4] = {35, 220, 291, 476};
    const double two_pi = 6.283185307179586476925286766559;
    for (int i = 0; i < 512; i++) {
        double a = two_pi * ((double)i + 0.5) / 512.0;
        cos_out[i] = (int)lround(250.0 * cos(a));
        sin_out[i] = (int)lround(250.0 * sin(a));
    }
    for (int k = 0; k < 4; k++) {
        cos_out[cos_fix[k]] = cos_out[cos_fix[k]] > 0 ? 105 : -105;
        sin_out[sin_fix[k]] = sin_out[sin_fix[k]] > 0 ? 105 : -105;
    }
}

/* ---- a11: hex -> chips -> BOC(1,1) half-chips (src/gal-sig.cpp:9-233) ----------------------
 * chip j = +1 for logic 0, -1 for logic 1, MSB of each hex digit first; each chip becomes two
 * half-chips and sboc() negates every even-indexed entry: E[2j] = -chip, E[2j+1] = +chip.  */
void e1o_halfchip_table(int prn, int is_e1c, short *out)
{
    const uint32_t *w = is_e1c ? E1C_PRN_WORDS[prn - 1] : E1B_PRN_WORDS[prn - 1];
    for (int j = 0; j < E1_CODE_LEN; j++) {
        int bit = (w[j >> 5] >> (31 - (j & 31))) & 1u;
        short chip = bit ? -1 : 1;
        out[2 * j] = (short)-chip;
        out[2 * j + 1] = chip;
    }
}

/* ---- a8: computeCodePhase (src/gal-sig.cpp:308-347) ------------------------------------- */
void e1o_restate(double rho_prev, double rho_cur, double dt, double grx_sec,
                 double *f_carr, double *f_code, double *code_phase0, int *ibit0, int *ipage_out)
{
    const double lambda_e1 = 0.1902936727983649;       /* LAMBDA_E1, constants.h:119       */
    const double carr_to_code = 0.0006493506493506494; /* CARR_TO_CODE_E1, constants.h:125 */
    const double c_light = 2.99792458e8;               /* SPEED_OF_LIGHT, constants.h:60   */
    double rhorate = (rho_cur - rho_prev) / dt;                    /* :315 */
    double fc = -rhorate / lambda_e1;                              /* :318 */
    *f_carr = fc;
    *f_code = 1.023e6 + fc * carr_to_code;                         /* :320 */
    double ms = (grx_sec - rho_cur / c_light) * 1000.0;            /* :322 */
    int ipage = (int)(ms / 2000.0);                                /* :324 */
    ms -= ipage * 2000;                                            /* :326 */
    int ibit = (int)((unsigned int)ms / 4);                        /* :328 */
    ms -= ibit * 4;                                                /* :329 */
    *code_phase0 = ms / 4 * E1_CODE_LEN;                           /* :330 */
    *ibit0 = (ibit + E1_SYM_PER_PAGE / 2) % E1_SYM_PER_PAGE;       /* :334 */
    *ipage_out = ipage % 360;                                      /* :339 */
}

/* ---- a1..a7: the sample loop (src/galileo-sdr.cpp:481-539) --------------------------------- */
static const unsigned char SEC25[E1_SEC_CODE_LEN] = /* GALILEO_E1_SECONDARY_CODE, constants.h:213 */
    {0, 0, 1, 1, 1, 0, 0, 0, 0, 0, 0, 0, 1, 0, 1, 0, 1, 1, 0, 1, 1, 0, 0, 1, 0};

typedef struct {
    short b[2 * E1_CODE_LEN];
    short c[2 * E1_CODE_LEN];
} code_pair;

static int g_lut_ready = 0;
static int g_cos[512], g_sin[512];
static code_pair *g_codes[E1_N_PRN_CODES + 1];
static pthread_mutex_t g_lock = PTHREAD_MUTEX_INITIALIZER;

static const code_pair *codes_for(int prn)
{
    pthread_mutex_lock(&g_lock);
    if (!g_lut_ready) {
        e1o_carrier_lut(g_cos, g_sin);
        g_lut_ready = 1;
    }
    if (!g_codes[prn]) {
        code_pair *p = (code_pair *)malloc(sizeof *p);
        e1o_halfchip_table(prn, 0, p->b);
        e1o_halfchip_table(prn, 1, p->c);
        g_codes[prn] = p;
    }
    pthread_mutex_unlock(&g_lock);
    return g_codes[prn];
}

static inline int page_bit(const uint8_t *pg, int k) { return (pg[k >> 3] >> (k & 7)) & 1; }

/* one epoch, channels [c0,c1): adds into acc[2*n_samp] (int32 I,Q) */
static void synth_epoch_channels(double delt, int n_samp, const e1_epoch_rec *rec, int c0, int c1,
                                 double *carr_phase, int *acc)
{
    for (int ch = c0; ch < c1; ch++) {
        const e1_epoch_rec *r = &rec[ch];
        if (r->prn <= 0)
            continue;
        const code_pair *cp = codes_for(r->prn);
        if (r->flags & E1_REC_SET_PHASE)
            carr_phase[ch] = r->carr_phase_init;
        double code_phase = r->code_phase0;
        double phi = carr_phase[ch];
        int ibit = r->ibit0;
        const uint8_t *page = r->page_cur;
        const double f_code = r->f_code, f_carr = r->f_carr;
        for (int k = 0; k < n_samp; k++) {
            if (code_phase >= E1_CODE_LEN) {               /* :491-507 */
                code_phase -= E1_CODE_LEN;
                ibit++;
                if (ibit >= E1_SYM_PER_PAGE) {
                    ibit = 0;
                    page = r->page_next;                   /* generateINavMsg() result */
                }
            }
            int it = ((int)(511 * phi)) & 511;             /* :509-510 */
            int cosv = g_cos[it], sinv = g_sin[it];
            int icode = (int)(code_phase * 2);             /* :512 */
            int eb = cp->b[icode], ec = cp->c[icode];      /* :514-515 */
            int databit = page_bit(page, ibit) ? -1 : 1;   /* :517 */
            int sec = SEC25[ibit % E1_SEC_CODE_LEN] ? -1 : 1; /* :518 */
            int m = eb * databit - ec * sec;               /* :520-521 */
            acc[2 * k] += m * cosv;                        /* :524-525 */
            acc[2 * k + 1] += m * sinv;
            code_phase += f_code * delt;                   /* :528 */
            phi += f_carr * delt;                          /* :531 */
            phi -= (long)phi;                              /* :532 */
        }
        carr_phase[ch] = phi;
    }
}

void e1o_synth_epochs(double fs_hz, int n_samp, int max_chan, int n_epochs,
                      const e1_epoch_rec *recs, double *carr_phase, int16_t *out)
{
    const double delt = 1.0 / fs_hz;                       /* :162 */
    int *acc = (int *)malloc(sizeof(int) * 2 * (size_t)n_samp);
    for (int e = 0; e < n_epochs; e++) {
        memset(acc, 0, sizeof(int) * 2 * (size_t)n_samp);
        synth_epoch_channels(delt, n_samp, recs + (size_t)e * max_chan, 0, max_chan, carr_phase, acc);
        int16_t *o = out + (size_t)e * n_samp * 2;
        for (int k = 0; k < 2 * n_samp; k++)
            o[k] = (short)acc[k];                          /* :536-537 */
    }
    free(acc);
}

/* ---- multi-threaded variant: the reference is single-threaded; this spreads channels over
 * host threads so the CPU baseline can use every core the box has.  Per epoch: every thread
 * synthesises its channels into a private int32 block, barrier, every thread sums a slice of
 * samples over all private blocks and casts to int16 (integer addition is associative, so the
 * bytes equal the serial version's), barrier.                                                */
typedef struct {
    double delt;
    int n_samp, max_chan, n_epochs, n_threads, tid, c0, c1;
    const e1_epoch_rec *recs;
    double *carr_phase;
    int **acc; /* [n_threads] -> private [2*n_samp] */
    int16_t *out;
    pthread_barrier_t *bar;
} mt_job;

static void *mt_worker(void *arg)
{
    mt_job *j = (mt_job *)arg;
    const size_t len = 2 * (size_t)j->n_samp;
    const size_t k0 = len * (size_t)j->tid / (size_t)j->n_threads, k1 = len * (size_t)(j->tid + 1) / (size_t)j->n_threads;
    int *mine = j->acc[j->tid];
    for (int e = 0; e < j->n_epochs; e++) {
        memset(mine, 0, sizeof(int) * len);
        synth_epoch_channels(j->delt, j->n_samp, j->recs + (size_t)e * j->max_chan, j->c0, j->c1, j->carr_phase, mine);
        pthread_barrier_wait(j->bar);
        int16_t *o = j->out + (size_t)e * len;
        for (size_t k = k0; k < k1; k++) {
            int s = 0;
            for (int t = 0; t < j->n_threads; t++)
                s += j->acc[t][k];
            o[k] = (short)s;
        }
        pthread_barrier_wait(j->bar);
    }
    return NULL;
}

void e1o_synth_epochs_mt(double fs_hz, int n_samp, int max_chan, int n_epochs,
                         const e1_epoch_rec *recs, double *carr_phase, int16_t *out, int n_threads)
{
    if (max_chan < 1 || n_samp < 1)
        return;
    if (n_threads > max_chan) n_threads = max_chan;
    if (n_threads < 1) n_threads = 1;
    for (int c = 0; c < max_chan; c++) /* build the shared tables before the threads start */
        for (int e = 0; e < n_epochs; e++)
            if (recs[(size_t)e * max_chan + c].prn > 0)
                codes_for(recs[(size_t)e * max_chan + c].prn);
    mt_job *jobs = (mt_job *)calloc((size_t)n_threads, sizeof *jobs);
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof *th);
    int **acc = (int **)calloc((size_t)n_threads, sizeof *acc);
    pthread_barrier_t bar;
    pthread_barrier_init(&bar, NULL, (unsigned)n_threads);
    for (int t = 0; t < n_threads; t++)
        acc[t] = (int *)malloc(sizeof(int) * 2 * (size_t)n_samp);
    for (int t = 0; t < n_threads; t++) {
        mt_job *j = &jobs[t];
        j->delt = 1.0 / fs_hz; j->n_samp = n_samp; j->max_chan = max_chan; j->n_epochs = n_epochs;
        j->n_threads = n_threads; j->tid = t;
        j->c0 = (int)((long)max_chan * t / n_threads);
        j->c1 = (int)((long)max_chan * (t + 1) / n_threads);
        j->recs = recs; j->carr_phase = carr_phase; j->acc = acc; j->out = out; j->bar = &bar;
        pthread_create(&th[t], NULL, mt_worker, j);
    }
    for (int t = 0; t < n_threads; t++)
        pthread_join(th[t], NULL);
    pthread_barrier_destroy(&bar);
    for (int t = 0; t < n_threads; t++)
        free(acc[t]);
    free(acc); free(jobs); free(th);
}

/* ---- carrier recurrence alone (src/galileo-sdr.cpp:531-532, src/channel.cpp:98-99) -----------------
 * The literal `phi += f_carr * delt; phi -= (long)phi` of every sample of every block, nothing else:
 * phases[e][ch] receives the phase each channel holds at the top of block e.  Lets a checker start the
 * full loop at any block of a long run from a phase that no planner had a hand in.  Channels split
 * over threads (each channel's walk is independent). */
typedef struct {
    double delt;
    int n_samp, max_chan, n_epochs, c0, c1;
    const e1_epoch_rec *recs;
    double *carr_phase, *phases;
} ph_job;

static void *ph_worker(void *arg)
{
    ph_job *j = (ph_job *)arg;
    for (int ch = j->c0; ch < j->c1; ch++) {
        double phi = j->carr_phase[ch];
        for (int e = 0; e < j->n_epochs; e++) {
            const e1_epoch_rec *r = &j->recs[(size_t)e * j->max_chan + ch];
            if (r->prn > 0) {
                if (r->flags & E1_REC_SET_PHASE)
                    phi = r->carr_phase_init;
                j->phases[(size_t)e * j->max_chan + ch] = phi;
                const double f_carr = r->f_carr, delt = j->delt;
                for (int k = 0; k < j->n_samp; k++) {
                    phi += f_carr * delt; /* :531 */
                    phi -= (long)phi;     /* :532 */
                }
            } else {
                j->phases[(size_t)e * j->max_chan + ch] = phi;
            }
        }
        j->carr_phase[ch] = phi;
    }
    return NULL;
}

void e1o_carrier_phases(double fs_hz, int n_samp, int max_chan, int n_epochs, const e1_epoch_rec *recs,
                        double *carr_phase, double *phases, int n_threads)
{
    if (max_chan < 1)
        return;
    if (n_threads > max_chan) n_threads = max_chan;
    if (n_threads < 1) n_threads = 1;
    ph_job *jobs = (ph_job *)calloc((size_t)n_threads, sizeof *jobs);
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof *th);
    for (int t = 0; t < n_threads; t++) {
        ph_job *j = &jobs[t];
        j->delt = 1.0 / fs_hz; j->n_samp = n_samp; j->max_chan = max_chan; j->n_epochs = n_epochs;
        j->c0 = (int)((long)max_chan * t / n_threads);
        j->c1 = (int)((long)max_chan * (t + 1) / n_threads);
        j->recs = recs; j->carr_phase = carr_phase; j->phases = phases;
        pthread_create(&th[t], NULL, ph_worker, j);
    }
    for (int t = 0; t < n_threads; t++)
        pthread_join(th[t], NULL);
    free(jobs); free(th);
}

/* ---- CBOC(6,1,1/11) sub-carrier and per-satellite gain --------------------------------
 * PARITY UNPINNED for alpha/beta != (1, 0) and for gains: the reference transmits BOC(1,1) at unit gain
 * (sboc(.., 1, 1), src/gal-sig.cpp:224,232; `* gain[i]` commented out, src/galileo-sdr.cpp:520-521), so there
 * are no reference bytes to hold this mode to.  Written from the Galileo OS SIS ICD (2.1.2, E1 CBOC):
 *     e_B(t) = c_B d (alpha sc_a + beta sc_b),   e_C(t) = c_C s (alpha sc_a - beta sc_b),   s = e_B - e_C
 * with the reference's own loop around it (code wrap / symbol advance :491-507, table index :509-510, phase
 * advance :528-532) and its sign convention for a sub-carrier (sboc negates the FIRST half period): sc_a < 0 on
 * the first half of a chip, sc_b < 0 on every even twelfth.  The sub-chip is (int)(code_phase * 12), the
 * resolution-12 analogue of :512.  Accumulated in double, stored as the nearest int16 (ties to even,
 * saturating).  With (alpha, beta) = (1, 0) and unit gains it writes the reference's bytes (tested).
 * Deterministic whatever the thread count: threads take whole blocks, each block sums its channels in slot
 * order from the phase the literal carrier recurrence (e1o_carrier_phases) reaches at its top. */
typedef struct {
    double delt, alpha, beta;
    int n_samp, max_chan, n_epochs, use_gain, tid, n_threads;
    const e1_epoch_rec *recs;
    const double *phases;
    int16_t *out;
} fl_job;

static void *fl_worker(void *arg)
{
    fl_job *j = (fl_job *)arg;
    double *acc = (double *)malloc(sizeof(double) * 2 * (size_t)j->n_samp);
    for (int e = j->tid; e < j->n_epochs; e += j->n_threads) {
        memset(acc, 0, sizeof(double) * 2 * (size_t)j->n_samp);
        for (int ch = 0; ch < j->max_chan; ch++) {
            const e1_epoch_rec *r = &j->recs[(size_t)e * j->max_chan + ch];
            if (r->prn <= 0)
                continue;
            const code_pair *cp = codes_for(r->prn);
            const double g = (j->use_gain && r->gain_q7 != 0) ? (double)r->gain_q7 / 128.0 : 1.0; /* gain[i] is scaled by 2^7 (:477) */
            double code_phase = r->code_phase0, phi = j->phases[(size_t)e * j->max_chan + ch];
            int ibit = r->ibit0;
            const uint8_t *page = r->page_cur;
            for (int k = 0; k < j->n_samp; k++) {
                if (code_phase >= E1_CODE_LEN) {
                    code_phase -= E1_CODE_LEN;
                    if (++ibit >= E1_SYM_PER_PAGE) {
                        ibit = 0;
                        page = r->page_next;
                    }
                }
                int it = ((int)(511 * phi)) & 511;
                int sub = (int)(code_phase * 12);
                int chip = sub / 12, tw = sub % 12;
                double sc_a = tw < 6 ? -1.0 : 1.0, sc_b = (tw & 1) ? 1.0 : -1.0;
                double c_b = cp->b[2 * chip + 1], c_c = cp->c[2 * chip + 1]; /* +chip sits on the odd half-chip */
                double d = page_bit(page, ibit) ? -1.0 : 1.0, sec = SEC25[ibit % E1_SEC_CODE_LEN] ? -1.0 : 1.0;
                double e_b = c_b * d * (j->alpha * sc_a + j->beta * sc_b);
                double e_c = c_c * sec * (j->alpha * sc_a - j->beta * sc_b);
                double m = g * (e_b - e_c);
                acc[2 * k] += m * g_cos[it];
                acc[2 * k + 1] += m * g_sin[it];
                code_phase += r->f_code * j->delt;
                phi += r->f_carr * j->delt;
                phi -= (long)phi;
            }
        }
        int16_t *o = j->out + (size_t)e * j->n_samp * 2;
        for (int k = 0; k < 2 * j->n_samp; k++) {
            double v = nearbyint(acc[k]); /* default rounding mode: to nearest, ties to even */
            o[k] = (int16_t)(v > 32767.0 ? 32767.0 : (v < -32768.0 ? -32768.0 : v));
        }
    }
    free(acc);
    return NULL;
}

void e1o_synth_epochs_float(double fs_hz, int n_samp, int max_chan, int n_epochs, const e1_epoch_rec *recs,
                            double *carr_phase, int16_t *out, double alpha, double beta, int use_gain, int n_threads)
{
    if (max_chan < 1 || n_samp < 1 || n_epochs < 1)
        return;
    if (n_threads < 1) n_threads = 1;
    for (int c = 0; c < max_chan; c++)
        for (int e = 0; e < n_epochs; e++)
            if (recs[(size_t)e * max_chan + c].prn > 0)
                codes_for(recs[(size_t)e * max_chan + c].prn);
    double *phases = (double *)malloc(sizeof(double) * (size_t)n_epochs * max_chan);
    e1o_carrier_phases(fs_hz, n_samp, max_chan, n_epochs, recs, carr_phase, phases, n_threads);
    if (n_threads > n_epochs) n_threads = n_epochs;
    fl_job *jobs = (fl_job *)calloc((size_t)n_threads, sizeof *jobs);
    pthread_t *th = (pthread_t *)calloc((size_t)n_threads, sizeof *th);
    for (int t = 0; t < n_threads; t++) {
        fl_job *j = &jobs[t];
        j->delt = 1.0 / fs_hz; j->alpha = alpha; j->beta = beta; j->n_samp = n_samp; j->max_chan = max_chan;
        j->n_epochs = n_epochs; j->use_gain = use_gain; j->tid = t; j->n_threads = n_threads;
        j->recs = recs; j->phases = phases; j->out = out;
        pthread_create(&th[t], NULL, fl_worker, j);
    }
    for (int t = 0; t < n_threads; t++)
        pthread_join(th[t], NULL);
    free(phases); free(jobs); free(th);
}
