// hostsim.cpp -- TEST SCAFFOLDING.  Compiles galileo-sdr-sim_b200/csrc/e1_core.h (the exact
// arithmetic the CUDA kernels are made of) for the host, and drives it in the same order the
// kernels do (plan code, plan carrier, per-tile parameters, per-thread sample runs), so the
// planner / closed-form / fallback logic can be checked against the oracle on a box without a
// GPU.  It is built by tests/e1util.py into tests/hostsim/libe1hostsim.so and loaded only by
// the tests; the product library (libe1b200.so) does not contain or call any of this.
#include <algorithm>
#include <cstdlib>
#include <cstring>
#include <vector>

#define E1_CHECK_LUT_BOUNDS 1
#include "../../galileo-sdr-sim_b200/csrc/e1_core.h"
#include "../../galileo-sdr-sim_b200/data/e1_prn_codes.h"

extern "C" {

double hs_carr_advance(double phi, double sp, long k0, long k1) { return e1_carr_advance(phi, sp, k0, k1); }

double hs_carr_literal(double phi, double sp, long n)
{
    for (long i = 0; i < n; i++)
        phi = e1_carr_step(phi, sp);
    return phi;
}

// code walk: returns cp after n samples, *wraps = number of wraps
double hs_code_advance(double cp, double sc, long n, long *wraps)
{
    int64_t k = 0;
    long w = 0;
    while (k < n) {
        int wr;
        cp = e1_walk_up(cp, sc, 4092.0, &k, n, &wr);
        w += wr;
    }
    *wraps = w;
    return cp;
}
double hs_code_literal(double cp, double sc, long n, long *wraps)
{
    long w = 0;
    for (long i = 0; i < n; i++) {
        int wr = 0;
        cp = e1_code_step(cp, sc, &wr);
        w += wr;
    }
    *wraps = w;
    return cp;
}

unsigned long long hs_to_fixed(double x, int sh) { return e1_to_fixed(x, sh); }

void hs_build_codes(uint32_t *codes)
{
    for (int p = 0; p < E1C_N_PRN; p++)
        e1_build_code_words(E1B_PRN_WORDS[p], E1C_PRN_WORDS[p], codes + (size_t)p * E1C_CODE_WORDS_PER_PRN);
}

void hs_build_lut(const int *cos512, const int *sin512, int32_t *lut) { e1_build_lut(cos512, sin512, lut); }
int hs_lut_entries(void) { return E1C_LUT_ENTRIES; }
unsigned long long hs_lut_oob(void) { return e1_lut_oob; } // fast-form lookups outside the carrier table since load (must stay 0)

int hs_code_words_per_prn(void) { return E1C_CODE_WORDS_PER_PRN; }

// tile-level ambiguity test (e1_first_hit / e1_par_clean)
long long hs_first_hit(long long a, long long d, long long M, long long L, long long n) { return e1_first_hit(a, d, M, L, n); }
static unsigned long long g_clean_tiles = 0, g_checked_tiles = 0, g_clean_violations = 0;
// (tile, channel) parameter sets marked clean / not marked since load, and runs of a clean tile that the
// tracking sample loop flagged or summed differently (must stay 0)
unsigned long long hs_clean_tiles(void) { return g_clean_tiles; }
unsigned long long hs_checked_tiles(void) { return g_checked_tiles; }
unsigned long long hs_clean_violations(void) { return g_clean_violations; }
// largest distance seen, in units of 2^-32 half-chip, between the code fraction the runs before a tile's code
// wrap step (HA + j dH) and the one the runs after it step (HB + j dH): e1_par_clean searches ONE sequence
// for the whole tile and widens its zone by thr_code for this
static double g_max_ab = 0.0;
double hs_max_code_ab_units(void) { return g_max_ab; }
// E1B200_CFG_CBOC / E1B200_CFG_GAIN for the following hs_synth_epochs* calls (0 = the integer path): the float
// path of e1_synth_float_kernel -- e1_channel_run_float per (thread, channel), FP32 sums, e1_f2i16 store.
static unsigned long long g_cw_pairs = 0; // (thread, channel) pairs that went through the carry-walk form since load
unsigned long long hs_cw_pairs(void) { return g_cw_pairs; }
static int g_cw_nh = 4; // runs of 16 samples per thread of the team kernel: 4 (64 samples, 3 x 128 threads) or 2
void hs_set_cw_runs(int nh) { g_cw_nh = nh == 2 ? 2 : 4; }
static uint32_t g_cfg_flags = 0;
void hs_set_cfg_flags(unsigned int flags) { g_cfg_flags = flags; }
// event-driven kernel (e1_synth_ev_kernel): -1 = as the library decides (e1_ev_context), 0 = never, 1 = whenever the tile is 8192 samples
static int g_ev_mode = -1;
void hs_set_ev_mode(int m) { g_ev_mode = m; }
static unsigned long long g_ev_pairs = 0, g_ev_rest = 0; // (thread, channel) pairs through e1_ev_add<false> / through e1_ev_rest_impl since load
unsigned long long hs_ev_pairs(void) { return g_ev_pairs; }
unsigned long long hs_ev_rest(void) { return g_ev_rest; }

} // extern "C"
// One tile of e1_synth_cw_kernel<NH, .>: every thread's NH x 16 samples, channel by channel.
template <int NH>
static void cw_tile(const e1_chan_par *par, int nact, const uint32_t *codes, const unsigned char *lut, int n_valid, int16_t *o, uint32_t thr_carr,
                    uint32_t thr_code, uint32_t tc_carr, uint32_t tc_code, uint32_t lim_code, unsigned long long *stats)
{
    const int RUN = NH * E1C_MAX_RUN, threads = E1C_THREADS * E1C_MAX_RUN / RUN;
    for (int tid = 0; tid < threads; tid++) {
        const int j0 = tid * RUN;
        if (j0 >= n_valid)
            continue;
        const unsigned char *lut_lane = lut + 4 * (tid & (E1C_LUT_REP - 1));
        int acc[NH * E1C_MAX_RUN] = {0};
        for (int a = 0; a < nact; a++) {
            const uint32_t rc = e1_cw_add<NH>(&par[a], e1_sp(codes), e1_sp(lut_lane), j0, acc, thr_carr, lim_code, 1u);
            if (par[a].misc & E1_PAR_SLOW)
                g_cw_pairs++;
            if (!rc) { // a tile marked clean went through the walk without tracking: beside it, the tracking walk --
                       // it must not flag anything and must add the same terms
                int a1[NH * E1C_MAX_RUN], a0[NH * E1C_MAX_RUN] = {0};
                unsigned long long dummy[2] = {0, 0};
                e1_cw_rest_impl<NH>(&par[a], codes, lut_lane, j0, a1, E1_RC_CHECK, thr_carr, thr_code, tc_carr, tc_code, &dummy[0], &dummy[1]);
                e1_cw_add<NH>(&par[a], e1_sp(codes), e1_sp(lut_lane), j0, a0, thr_carr, lim_code, 1u);
                if (dummy[1] || memcmp(a0, a1, sizeof a0))
                    g_clean_violations++;
            } else { // e1_cw_rest
                int d[NH * E1C_MAX_RUN];
                e1_cw_rest_impl<NH>(&par[a], codes, lut_lane, j0, d, rc, thr_carr, thr_code, tc_carr, tc_code, &stats[0], &stats[2]);
                for (int i = 0; i < RUN; i++)
                    acc[i] += d[i];
            }
        }
        for (int i = 0; i < RUN; i++)
            if (j0 + i < n_valid) {
                uint32_t w = e1_pack_iq(acc[i]);
                o[(size_t)(j0 + i) * 2] = (int16_t)(w & 0xffffu);
                o[(size_t)(j0 + i) * 2 + 1] = (int16_t)(w >> 16);
            }
    }
}
// One tile of e1_synth_ev_kernel: every thread's E1C_EV_RUN samples as a column of differences, channel by channel, then the running sum.
static void ev_tile(const e1_chan_par *par, int nact, const uint32_t *codes, const unsigned char *lut, const int32_t *lut1, int n_valid, int16_t *o,
                    uint32_t thr_carr, uint32_t thr_code, uint32_t tc_code, uint32_t lim_code, unsigned long long *stats)
{
    const int RUN = E1C_EV_RUN, threads = E1C_THREADS * E1C_MAX_RUN / RUN;
    for (int tid = 0; tid < threads; tid++) {
        const int j0 = tid * RUN;
        if (j0 >= n_valid)
            continue;
        const int n = n_valid - j0 < RUN ? n_valid - j0 : RUN;
        const unsigned char *lut_lane = lut + 4 * (tid & (E1C_LUT_REP - 1));
        int diff[E1C_EV_RUN + 1] = {0};
        for (int a = 0; a < nact; a++) {
            const uint32_t want = E1_PAR_EV | E1_PAR_CLEAN;
            if ((par[a].misc & want) == want) {
                if (n == RUN && !(par[a].j_w > j0 && par[a].j_w < j0 + RUN)) { // the kernel's inline form
                    uint64_t H;
                    uint32_t w0, w1;
                    e1_ev_fetch(&par[a], codes, j0, &H, &w0, &w1);
                    e1_ev_run64(&par[a], H, w0, w1, e1_sp(lut1), j0, diff, e1_tc_carr_cw(thr_carr, RUN), e1_lim_carr_cw(thr_carr, RUN));
                } else
                    e1_ev_add<false>(&par[a], codes, e1_sp(lut1), j0, n, diff, 1, thr_carr, lim_code);
                g_ev_pairs++;
                // beside it: the same events with the tracking on must not flag anything, and the generic form must give the same terms
                int d1[E1C_EV_RUN + 1] = {0}, d0[E1C_EV_RUN + 1] = {0}, d2[E1C_EV_RUN + 1] = {0};
                unsigned long long dummy[2] = {0, 0};
                e1_ev_add<false>(&par[a], codes, e1_sp(lut1), j0, n, d0, 1, thr_carr, lim_code);
                const uint32_t flagged = e1_ev_add<true>(&par[a], codes, e1_sp(lut1), j0, n, d1, 1, thr_carr, lim_code);
                const bool generic_too = (tid + a) % 4 == 0; // a quarter of the pairs (the generic form is 20 times slower)
                if (generic_too) {
                    e1_chan_par q = par[a];
                    q.misc &= ~(uint32_t)E1_PAR_EV; // straight to the generic form
                    e1_ev_rest_impl(&q, e1_sp(lut1), codes, lut_lane, j0, n, d2, thr_carr, thr_code, tc_code, &dummy[0], &dummy[1]);
                }
                if (flagged || dummy[0] || memcmp(d0, d1, sizeof(int) * n) || (generic_too && memcmp(d0, d2, sizeof(int) * n)))
                    g_clean_violations++;
            } else {
                e1_ev_rest_impl(&par[a], e1_sp(lut1), codes, lut_lane, j0, n, diff, thr_carr, thr_code, tc_code, &stats[0], &stats[2]);
                g_ev_rest++;
            }
        }
        int s = 0;
        for (int i = 0; i < n; i++) {
            s += diff[i];
            uint32_t w = e1_pack_iq(s);
            o[(size_t)(j0 + i) * 2] = (int16_t)(w & 0xffffu);
            o[(size_t)(j0 + i) * 2 + 1] = (int16_t)(w >> 16);
        }
    }
}
extern "C" {

// Whole pipeline on the host.  lut: int32[642][32] in the product layout (passed in by the test
// from the oracle's tables so this file holds no second copy of them).
// stats[0] = samples resolved by the literal fallback, stats[1] = planner errors,
// stats[2] = threads that took the slow path.
// planner: 1 = serial reference planner (one exact walk per channel), 2 = parallel planner (drift
// pass, estimate prefix, span pass, chain) in the order the kernels run it.
// stats[3] = epochs the chain walked serially, stats[4] = HAT units accepted (planner 2 only).
int hs_synth_epochs_p(double fs_hz, int n_samp, int max_chan, int n_epochs, const e1_epoch_rec *recs, double *phase,
                      int16_t *out, int groups, int amb_scale, const int32_t *lut, unsigned long long *stats, int planner);

int hs_synth_epochs(double fs_hz, int n_samp, int max_chan, int n_epochs, const e1_epoch_rec *recs, double *phase,
                    int16_t *out, int groups, int amb_scale, const int32_t *lut, unsigned long long *stats)
{
    unsigned long long st[5];
    int rc = hs_synth_epochs_p(fs_hz, n_samp, max_chan, n_epochs, recs, phase, out, groups, amb_scale, lut, st, 1);
    stats[0] = st[0], stats[1] = st[1], stats[2] = st[2];
    return rc;
}

int hs_synth_epochs_p(double fs_hz, int n_samp, int max_chan, int n_epochs, const e1_epoch_rec *recs, double *phase,
                      int16_t *out, int groups, int amb_scale, const int32_t *lut, unsigned long long *stats, int planner)
{
    const double delt = 1.0 / fs_hz;
    e1_lut_base = (const unsigned char *)lut;
    const int threads = E1C_THREADS, run = 4 * groups, tile = threads * run;
    const int tpe = (n_samp + tile - 1) / tile;
    std::vector<uint32_t> codes(E1C_N_PRN * E1C_CODE_WORDS_PER_PRN);
    hs_build_codes(codes.data());
    std::vector<e1_tile_ck> ck((size_t)n_epochs * tpe * max_chan);
    memset(ck.data(), 0, ck.size() * sizeof(e1_tile_ck));
    for (int e = 0; e < n_epochs; e++)
        for (int ch = 0; ch < max_chan; ch++)
            e1_plan_code_epoch(&recs[(size_t)e * max_chan + ch], &ck[(size_t)e * tpe * max_chan + ch], max_chan, n_samp, tile,
                               tpe, delt);
    // planner units: spans of span_tiles tiles (e1_span_geometry), channel-major [max_chan][n_units]
    const e1_span_geo geo = e1_span_geometry(tpe);
    const int S = geo.spans_per_epoch, n_units = n_epochs * S;
    std::vector<e1_trans> delta((size_t)n_units * max_chan, e1_trans{0.0, 0.0, 0, 0}); // [ch][unit]
    stats[3] = stats[4] = 0;
    if (planner == 1) {
        for (int ch = 0; ch < max_chan; ch++) {
            double phi = phase[ch];
            for (int e = 0; e < n_epochs; e++)
                phi = e1_plan_carr_epoch(&recs[(size_t)e * max_chan + ch], &ck[(size_t)e * tpe * max_chan + ch], max_chan, phi,
                                         n_samp, tile, tpe, delt);
            phase[ch] = phi;
        }
    } else {
        const size_t ne = (size_t)n_units * max_chan;
        std::vector<e1_prep> prep(ne);
        std::vector<double> g(ne), dend(ne), est(ne);
        std::vector<e1_unit> units(ne);
        for (int e = 0; e < n_epochs; e++)
            for (int ch = 0; ch < max_chan; ch++)
                for (int sp = 0; sp < S; sp++)
                    e1_v2_prep(&recs[(size_t)e * max_chan + ch], delt, sp, e1_span_samples(&geo, sp, n_samp, tile),
                               &prep[(size_t)ch * n_units + (size_t)e * S + sp]);
        for (int ch = 0; ch < max_chan; ch++)
            e1_v2_ideal_prefix(&prep[(size_t)ch * n_units], n_units, phase[ch], &g[(size_t)ch * n_units]);
        for (size_t i = 0; i < ne; i++)
            dend[i] = e1_v2_drift_unit(&prep[i], g[i]);
        for (int ch = 0; ch < max_chan; ch++) {
            const size_t o = (size_t)ch * n_units;
            e1_v2_estimate_prefix(&prep[o], n_units, phase[ch], &g[o], &dend[o], &est[o]);
        }
        for (int ch = 0; ch < max_chan; ch++)
            for (int u = 0; u < n_units; u++) {
                const size_t i = (size_t)ch * n_units + u;
                const int e = u / S, sp = u - e * S;
                e1_v2_span_unit(&prep[i], u, phase[ch], &est[i], tile,
                                &ck[((size_t)e * tpe + (size_t)sp * geo.span_tiles) * max_chan + ch], max_chan, &units[i]);
            }
        for (int ch = 0; ch < max_chan; ch++) {
            unsigned long long st2[2] = {0, 0};
            const size_t o = (size_t)ch * n_units;
            phase[ch] = e1_v2_chain(&prep[o], n_units, phase[ch], tile, &units[o], &ck[ch], max_chan, &delta[o], st2);
            stats[3] += st2[0];
            stats[4] += st2[1];
        }
    }
    const uint32_t thr_carr = e1_thr_carr(tile, amb_scale), thr_code = e1_thr_code(tile, amb_scale);
    // event-driven context (e1_synth_ev_kernel): the code fraction is stepped from one start per E1C_EV_RUN samples
    const int ev = !(g_cfg_flags & (E1B200_CFG_CBOC | E1B200_CFG_GAIN)) && (g_ev_mode < 0 ? e1_ev_context(fs_hz, run) : (g_ev_mode && run == E1C_MAX_RUN));
    const int code_run = ev ? E1C_EV_RUN : run;
    const uint32_t tc_carr = e1_tc_carr(thr_carr, run), tc_code = e1_tc_code(thr_code, code_run);
    const uint32_t lim_carr = e1_lim_carr(tc_carr, thr_carr), lim_code = e1_lim_code(tc_code, thr_code);
    std::vector<e1_chan_par> par(max_chan);
    std::vector<int32_t> lut1(3 * E1C_LUT1_WORDS);
    e1_build_lut1(lut, lut1.data());
    stats[0] = stats[1] = stats[2] = 0;
    for (int e = 0; e < n_epochs; e++)
        for (int t = 0; t < tpe; t++) {
            int nact = 0; // e1_finalize_kernel
            for (int ch = 0; ch < max_chan; ch++) {
                const e1_tile_ck *c = &ck[((size_t)e * tpe + t) * max_chan + ch];
                if (!(c->sym & E1_CK_ACTIVE))
                    continue;
                if (c->sym & E1_CK_ERROR)
                    stats[1]++;
                const int sp = t / geo.span_tiles;
                e1_make_par(c, &recs[(size_t)e * max_chan + ch], delt, tile,
                            e1_trans_at(&delta[(size_t)ch * n_units + (size_t)e * S + sp], (t - sp * geo.span_tiles) * tile), tc_code,
                            &par[nact], g_cfg_flags | (ev ? E1_INT_EV : 0u));
                if (run == E1C_MAX_RUN && par[nact].j_w != E1C_NO_WRAP && !(par[nact].misc & E1_PAR_FORCE)) {
                    const int64_t d51 = (int64_t)((par[nact].HA - par[nact].HB) << 13) >> 13; // mod 2^51, signed
                    const double units = (double)(d51 < 0 ? -d51 : d51) / 524288.0;
                    if (units > g_max_ab)
                        g_max_ab = units;
                }
                if (run == E1C_MAX_RUN) { // e1_clean_kernel
                    if (e1_par_clean(&par[nact], tile, tc_carr, lim_carr, lim_code, thr_code, ev ? E1C_EV_RUN : g_cw_nh * E1C_MAX_RUN, code_run)) {
                        par[nact].misc |= E1_PAR_CLEAN;
                        g_clean_tiles++;
                    } else
                        g_checked_tiles++;
                }
                nact++;
            }
            const int n_valid = (n_samp - t * tile) < tile ? (n_samp - t * tile) : tile;
            int16_t *o = out + ((size_t)e * n_samp + (size_t)t * tile) * 2;
            if (g_cfg_flags & (E1B200_CFG_CBOC | E1B200_CFG_GAIN)) { // e1_synth_float_kernel: 512 threads x 16 samples
                const float alpha = (g_cfg_flags & E1B200_CFG_CBOC) ? (float)E1C_ALPHA_CBOC : 1.0f;
                const float beta = (g_cfg_flags & E1B200_CFG_CBOC) ? (float)E1C_BETA_CBOC : 0.0f;
                for (int tid = 0; tid < threads; tid++) {
                    const int j0 = tid * run;
                    if (j0 >= n_valid)
                        continue;
                    const unsigned char *lut_lane = (const unsigned char *)lut + 4 * (tid & (E1C_LUT_REP - 1));
                    float fi[E1C_MAX_RUN] = {0}, fq[E1C_MAX_RUN] = {0};
                    for (int a = 0; a < nact; a++) {
                        const float g = e1_bits_float(par[a].pat_a);
                        e1_channel_run_float(&par[a], codes.data(), lut_lane, j0, run, fi, fq, g * alpha, g * beta, thr_carr, thr_code,
                                             e1_bias_h(tc_code), &stats[0]);
                    }
                    for (int i = 0; i < run; i++)
                        if (j0 + i < n_valid) {
                            o[(size_t)(j0 + i) * 2] = (int16_t)e1_f2i16(fi[i]);
                            o[(size_t)(j0 + i) * 2 + 1] = (int16_t)e1_f2i16(fq[i]);
                        }
                }
                continue;
            }
            if (ev) { // e1_synth_ev_kernel
                ev_tile(par.data(), nact, codes.data(), (const unsigned char *)lut, lut1.data(), n_valid, o, thr_carr, thr_code, tc_code, lim_code, stats);
                continue;
            }
            if (run == E1C_MAX_RUN) { // e1_synth_cw_kernel<NH, TEAMS>: 8192 / (16 NH) threads per tile, NH runs of 16 samples each
                if (g_cw_nh == 4)
                    cw_tile<4>(par.data(), nact, codes.data(), (const unsigned char *)lut, n_valid, o, thr_carr, thr_code, tc_carr, tc_code, lim_code, stats);
                else
                    cw_tile<2>(par.data(), nact, codes.data(), (const unsigned char *)lut, n_valid, o, thr_carr, thr_code, tc_carr, tc_code, lim_code, stats);
                continue;
            }
            for (int tid = 0; tid < threads; tid++) { // e1_synth_kernel, one thread
                const int j0 = tid * run;
                if (j0 >= n_valid)
                    continue;
                const unsigned char *lut_lane = (const unsigned char *)lut + 4 * (tid & (E1C_LUT_REP - 1));
                int acc[E1C_MAX_RUN] = {0};
                for (int a = 0; a < nact; a++) {
                    uint32_t rc;
                    switch (run) {
                    case 4: rc = e1_run_fast<4>(&par[a], codes.data(), lut_lane, j0, acc, tc_carr, lim_carr, lim_code); break;
                    case 8: rc = e1_run_fast<8>(&par[a], codes.data(), lut_lane, j0, acc, tc_carr, lim_carr, lim_code); break;
                    default: rc = e1_run_fast<16>(&par[a], codes.data(), lut_lane, j0, acc, tc_carr, lim_carr, lim_code); break;
                    }
                    if (!rc)
                        continue;
                    stats[2]++; // e1_fix_run
                    int tt[E1C_MAX_RUN] = {0}, g[E1C_MAX_RUN];
                    switch (run) {
                    case 4: e1_run_fast<4>(&par[a], codes.data(), lut_lane, j0, tt, tc_carr, lim_carr, lim_code); break;
                    case 8: e1_run_fast<8>(&par[a], codes.data(), lut_lane, j0, tt, tc_carr, lim_carr, lim_code); break;
                    default: e1_run_fast<16>(&par[a], codes.data(), lut_lane, j0, tt, tc_carr, lim_carr, lim_code); break;
                    }
                    e1_channel_run(&par[a], codes.data(), lut_lane, j0, run, g, thr_carr, thr_code, e1_bias_h(tc_code), &stats[0]);
                    for (int i = 0; i < run; i++)
                        acc[i] += g[i] - tt[i];
                }
                for (int i = 0; i < run; i++)
                    if (j0 + i < n_valid) {
                        uint32_t w = e1_pack_iq(acc[i]);
                        o[(size_t)(j0 + i) * 2] = (int16_t)(w & 0xffffu);
                        o[(size_t)(j0 + i) * 2 + 1] = (int16_t)(w >> 16);
                    }
            }
        }
    return stats[1] ? -1 : 0;
}

// The chain kernel's round structure on the host (e1_v2_chain_kernel in e1_kernels.cuh: rounds of
// `round` spans, optimistic scan of T = true |phase| after the most recent wrap, prefix commit, one
// serial step at the first span that fails, resume), statement for statement, against the serial chain
// e1_v2_chain on the same span-pass output.  Returns the number of differing translations (+1 if the
// final phase differs); stats[0] = spans accepted by the scan, stats[1] = spans through the serial step.
long hs_chain_scan_compare(double fs_hz, int n_samp, int max_chan, int n_epochs, const e1_epoch_rec *recs, const double *phase0,
                           int round, unsigned long long *stats)
{
    const double delt = 1.0 / fs_hz;
    const int tile = 4 * 4 * E1C_THREADS;
    const int tpe = (n_samp + tile - 1) / tile;
    const e1_span_geo geo = e1_span_geometry(tpe);
    const int S_ = geo.spans_per_epoch, n_units = n_epochs * S_;
    const size_t ne = (size_t)n_units * max_chan;
    std::vector<e1_tile_ck> ck1((size_t)n_epochs * tpe * max_chan), ck2(ck1.size());
    memset(ck1.data(), 0, ck1.size() * sizeof(e1_tile_ck));
    memset(ck2.data(), 0, ck2.size() * sizeof(e1_tile_ck));
    std::vector<double> g(ne), dend(ne), est(ne);
    std::vector<e1_trans> d1(ne, e1_trans{0.0, 0.0, 0, 0}), d2(ne, e1_trans{0.0, 0.0, 0, 0});
    std::vector<e1_unit> units(ne);
    std::vector<e1_prep> prep(ne);
    for (int e = 0; e < n_epochs; e++)
        for (int ch = 0; ch < max_chan; ch++)
            for (int sp = 0; sp < S_; sp++)
                e1_v2_prep(&recs[(size_t)e * max_chan + ch], delt, sp, e1_span_samples(&geo, sp, n_samp, tile),
                           &prep[(size_t)ch * n_units + (size_t)e * S_ + sp]);
    for (int ch = 0; ch < max_chan; ch++)
        e1_v2_ideal_prefix(&prep[(size_t)ch * n_units], n_units, phase0[ch], &g[(size_t)ch * n_units]);
    for (size_t i = 0; i < ne; i++)
        dend[i] = e1_v2_drift_unit(&prep[i], g[i]);
    for (int ch = 0; ch < max_chan; ch++) {
        const size_t o = (size_t)ch * n_units;
        e1_v2_estimate_prefix(&prep[o], n_units, phase0[ch], &g[o], &dend[o], &est[o]);
    }
    for (int ch = 0; ch < max_chan; ch++)
        for (int u = 0; u < n_units; u++) {
            const size_t i = (size_t)ch * n_units + u;
            const int e = u / S_, sp = u - e * S_;
            e1_v2_span_unit(&prep[i], u, phase0[ch], &est[i], tile,
                            &ck1[((size_t)e * tpe + (size_t)sp * geo.span_tiles) * max_chan + ch], max_chan, &units[i]);
        }
    ck2 = ck1;
    long bad_total = 0;
    stats[0] = stats[1] = 0;
    for (int ch = 0; ch < max_chan; ch++) {
        const size_t o = (size_t)ch * n_units;
        unsigned long long st1[2] = {0, 0}, st2[2] = {0, 0};
        std::vector<e1_unit> u1(units.begin() + o, units.begin() + o + n_units);
        const double p1 = e1_v2_chain(&prep[o], n_units, phase0[ch], tile, u1.data(), &ck1[ch], max_chan, &d1[o], st1);
        // ---- the kernel's rounds
        e1_chain_state cs;
        e1_chain_init(&cs, phase0[ch]);
        std::vector<double> x(round), Sv(round), D(round);
        std::vector<int> Wv(round);
        for (int e0 = 0; e0 < n_units; e0 += round) {
            const int n = std::min(round, n_units - e0);
            const e1_unit *su = &units[o + e0];
            e1_trans *sd = &d2[o + e0];
            int start = 0, restarts = 0;
            while (start < n) {
                double run = 0.0;
                int wmax = -1;
                for (int tid = 0; tid < n; tid++) { // the two inclusive scans
                    x[tid] = 0.0;
                    int wi = -1;
                    if (tid >= start && su[tid].type == E1_UNIT_HAT && su[tid].last_k >= 1) {
                        x[tid] = e1_add(su[tid].last_p, -su[tid].anchor_p);
                        wi = tid;
                    }
                    run = e1_add(run, x[tid]);
                    wmax = std::max(wmax, wi);
                    Sv[tid] = run;
                    Wv[tid] = wmax;
                }
                int bad = n;
                for (int tid = start; tid < n; tid++) {
                    const e1_unit *u = &su[tid];
                    const double Tb = e1_add(cs.prev_p, e1_add(Sv[tid], -x[tid]));
                    const int Wb = tid > 0 ? Wv[tid - 1] : -1;
                    int p_ok, p_neg, p_k, p_u;
                    if (Wb >= start) {
                        const e1_unit *q = &su[Wb];
                        p_ok = 1, p_neg = q->neg, p_k = q->last_k, p_u = e0 + Wb;
                    } else {
                        p_ok = cs.prev_ok, p_neg = cs.prev_neg, p_k = cs.prev_k, p_u = cs.prev_u;
                    }
                    D[tid] = e1_add(Tb, -u->anchor_p);
                    const int ok = u->type == E1_UNIT_HAT && u->tie == 0 && p_ok && p_neg == u->neg && p_u == e0 + tid - u->anchor_back &&
                                   p_k == u->anchor_k;
                    if (!(ok && D[tid] >= u->lo && D[tid] < u->hi) && tid < bad)
                        bad = tid;
                }
                for (int tid = start; tid < bad; tid++) {
                    e1_trans tr;
                    tr.a = tr.b = su[tid].neg ? -D[tid] : D[tid];
                    tr.k_split = 0;
                    tr.pad = 0;
                    sd[tid] = tr;
                }
                if (bad > start) {
                    const int tid = bad - 1;
                    const e1_unit *u = &su[tid];
                    cs.phi = e1_add(u->end_phi, u->neg ? -D[tid] : D[tid]);
                    if (Wv[tid] >= start) {
                        const e1_unit *q = &su[Wv[tid]];
                        cs.prev_p = e1_add(cs.prev_p, Sv[tid]);
                        cs.prev_k = q->last_k;
                        cs.prev_u = e0 + Wv[tid];
                        cs.prev_ok = 1;
                        cs.prev_neg = q->neg;
                    }
                }
                st2[1] += (unsigned long long)(bad - start);
                if (bad >= n)
                    break;
                const int upto = ++restarts > 16 ? n : bad + 1;
                for (int i = bad; i < upto; i++) {
                    const int u_abs = e0 + i, ep = u_abs / S_, sp = u_abs - ep * S_;
                    sd[i] = e1_v2_chain_step(&cs, &su[i], u_abs, prep[o + u_abs].sp, prep[o + u_abs].n, tile,
                                             (prep[o + u_abs].n + tile - 1) / tile,
                                             &ck2[((size_t)ep * tpe + (size_t)sp * geo.span_tiles) * max_chan + ch], max_chan, st2);
                    stats[1]++;
                }
                start = upto;
            }
        }
        stats[0] += st2[1];
        for (int u = 0; u < n_units; u++) {
            const e1_trans &a = d1[o + u], &b = d2[o + u];
            if (e1_bits(a.a) != e1_bits(b.a) || e1_bits(a.b) != e1_bits(b.b) || a.k_split != b.k_split)
                if (!(a.a == b.a && a.b == b.b && a.k_split == b.k_split))
                    bad_total++;
        }
        if (p1 != cs.phi)
            bad_total++;
    }
    return bad_total;
}

// Carrier planners only: serial walk (planner 1) against the parallel passes (planner 2), every
// tile checkpoint and the final phase compared bit for bit.  Returns the number of mismatches;
// stats[0] = epochs the chain walked serially, stats[1] = HAT epochs accepted, stats[2] = active epochs.
long hs_plan_compare(double fs_hz, int n_samp, int max_chan, int n_epochs, const e1_epoch_rec *recs, const double *phase0,
                     int groups, unsigned long long *stats)
{
    const double delt = 1.0 / fs_hz;
    const int tile = groups * 4 * E1C_THREADS;
    const int tpe = (n_samp + tile - 1) / tile;
    const size_t nec = (size_t)n_epochs * max_chan;
    const e1_span_geo geo = e1_span_geometry(tpe);
    const int S = geo.spans_per_epoch, n_units = n_epochs * S;
    const size_t ne = (size_t)n_units * max_chan;
    std::vector<e1_tile_ck> ck1(nec * tpe), ck2(nec * tpe);
    memset(ck1.data(), 0, ck1.size() * sizeof(e1_tile_ck));
    memset(ck2.data(), 0, ck2.size() * sizeof(e1_tile_ck));
    std::vector<double> g(ne), dend(ne), est(ne), p1(max_chan), p2(max_chan);
    std::vector<e1_trans> delta(ne, e1_trans{0.0, 0.0, 0, 0});
    std::vector<e1_unit> units(ne);
    stats[0] = stats[1] = stats[2] = 0;
    for (int ch = 0; ch < max_chan; ch++) {
        double phi = phase0[ch];
        for (int e = 0; e < n_epochs; e++)
            phi = e1_plan_carr_epoch(&recs[(size_t)e * max_chan + ch], &ck1[(size_t)e * tpe * max_chan + ch], max_chan, phi, n_samp,
                                     tile, tpe, delt);
        p1[ch] = phi;
    }
    std::vector<e1_prep> prep(ne);
    for (int e = 0; e < n_epochs; e++)
        for (int ch = 0; ch < max_chan; ch++)
            for (int sp = 0; sp < S; sp++)
                e1_v2_prep(&recs[(size_t)e * max_chan + ch], delt, sp, e1_span_samples(&geo, sp, n_samp, tile),
                           &prep[(size_t)ch * n_units + (size_t)e * S + sp]);
    for (int ch = 0; ch < max_chan; ch++)
        e1_v2_ideal_prefix(&prep[(size_t)ch * n_units], n_units, phase0[ch], &g[(size_t)ch * n_units]);
    for (size_t i = 0; i < ne; i++)
        dend[i] = e1_v2_drift_unit(&prep[i], g[i]);
    for (int ch = 0; ch < max_chan; ch++) {
        const size_t o = (size_t)ch * n_units;
        e1_v2_estimate_prefix(&prep[o], n_units, phase0[ch], &g[o], &dend[o], &est[o]);
    }
    if (getenv("HS_NO_DRIFT")) // test hook: guesses from the ideal line only (no drift estimates)
        est = g;
    for (int ch = 0; ch < max_chan; ch++)
        for (int u = 0; u < n_units; u++) {
            const size_t i = (size_t)ch * n_units + u;
            const int e = u / S, sp = u - e * S;
            e1_v2_span_unit(&prep[i], u, phase0[ch], &est[i], tile,
                            &ck2[((size_t)e * tpe + (size_t)sp * geo.span_tiles) * max_chan + ch], max_chan, &units[i]);
        }
    for (int ch = 0; ch < max_chan; ch++) {
        const size_t o = (size_t)ch * n_units;
        p2[ch] = e1_v2_chain(&prep[o], n_units, phase0[ch], tile, &units[o], &ck2[ch], max_chan, &delta[o], stats);
    }
    long bad = 0;
    for (int e = 0; e < n_epochs; e++)
        for (int ch = 0; ch < max_chan; ch++) {
            const size_t i = (size_t)e * max_chan + ch;
            if (!e1_rec_active(&recs[i]))
                continue;
            stats[2] += S;
            for (int t = 0; t < tpe; t++) {
                const size_t j = ((size_t)e * tpe + t) * max_chan + ch;
                const int sp = t / geo.span_tiles;
                const double v = e1_add(ck2[j].phi, e1_trans_at(&delta[(size_t)ch * n_units + (size_t)e * S + sp], (t - sp * geo.span_tiles) * tile));
                if (e1_bits(ck1[j].phi) != e1_bits(v) && !(ck1[j].phi == 0.0 && v == 0.0))
                    bad++;
            }
        }
    for (int ch = 0; ch < max_chan; ch++)
        if (p1[ch] != p2[ch])
            bad++;
    return bad;
}
}
