// hostsim.cpp -- TEST SCAFFOLDING.  Compiles galileo-sdr-sim_b200/csrc/e1_core.h (the exact
// arithmetic the CUDA kernels are made of) for the host, and drives it in the same order the
// kernels do (plan code, plan carrier, per-tile parameters, per-thread sample runs), so the
// planner / closed-form / fallback logic can be checked against the oracle on a box without a
// GPU.  It is built by tests/e1util.py into tests/hostsim/libe1hostsim.so and loaded only by
// the tests; the product library (libe1b200.so) does not contain or call any of this.
#include <cstdlib>
#include <cstring>
#include <vector>

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
    memset(codes, 0, E1C_N_PRN * E1C_CODE_WORDS_PER_PRN * 4);
    for (int p = 0; p < E1C_N_PRN; p++)
        for (int c = 0; c < 4092; c++) {
            uint32_t b = (E1B_PRN_WORDS[p][c >> 5] >> (31 - (c & 31))) & 1u;
            uint32_t q = (E1C_PRN_WORDS[p][c >> 5] >> (31 - (c & 31))) & 1u;
            codes[p * E1C_CODE_WORDS_PER_PRN + (c >> 4)] |= (b | (q << 1)) << ((c & 15) * 2);
        }
}

// Whole pipeline on the host.  lut: int32[1024] as the product builds it (passed in by the test
// from the oracle's tables so this file holds no second copy of them).
// stats[0] = samples resolved by the literal fallback, stats[1] = planner errors,
// stats[2] = threads that took the slow path.
int hs_synth_epochs(double fs_hz, int n_samp, int max_chan, int n_epochs, const e1_epoch_rec *recs, double *phase,
                    int16_t *out, int groups, int amb_scale, const int32_t *lut, unsigned long long *stats)
{
    const double delt = 1.0 / fs_hz;
    const int threads = 256, tile = groups * threads * E1C_RUN;
    const int tpe = (n_samp + tile - 1) / tile;
    std::vector<uint32_t> codes(E1C_N_PRN * E1C_CODE_WORDS_PER_PRN);
    hs_build_codes(codes.data());
    std::vector<e1_tile_ck> ck((size_t)n_epochs * tpe * max_chan);
    memset(ck.data(), 0, ck.size() * sizeof(e1_tile_ck));
    for (int e = 0; e < n_epochs; e++)
        for (int ch = 0; ch < max_chan; ch++)
            e1_plan_code_epoch(&recs[(size_t)e * max_chan + ch], &ck[(size_t)e * tpe * max_chan + ch], max_chan, n_samp, tile,
                               tpe, delt);
    for (int ch = 0; ch < max_chan; ch++) {
        double phi = phase[ch];
        for (int e = 0; e < n_epochs; e++)
            phi = e1_plan_carr_epoch(&recs[(size_t)e * max_chan + ch], &ck[(size_t)e * tpe * max_chan + ch], max_chan, phi,
                                     n_samp, tile, tpe, delt);
        phase[ch] = phi;
    }
    const uint32_t thr_carr = e1_thr_carr(tile, amb_scale), thr_code = e1_thr_code(tile, amb_scale);
    std::vector<e1_chan_par> par(max_chan);
    stats[0] = stats[1] = stats[2] = 0;
    for (int e = 0; e < n_epochs; e++)
        for (int t = 0; t < tpe; t++) {
            int nact = 0;
            for (int ch = 0; ch < max_chan; ch++) {
                const e1_tile_ck *c = &ck[((size_t)e * tpe + t) * max_chan + ch];
                if (!(c->sym & E1_CK_ACTIVE))
                    continue;
                if (c->sym & E1_CK_ERROR)
                    stats[1]++;
                e1_make_par(c, &recs[(size_t)e * max_chan + ch], delt, tile, &par[nact++]);
            }
            const int n_valid = (n_samp - t * tile) < tile ? (n_samp - t * tile) : tile;
            int16_t *o = out + ((size_t)e * n_samp + (size_t)t * tile) * 2;
            for (int g = 0; g < groups; g++)
                for (int tid = 0; tid < threads; tid++) {
                    const int j0 = (g * threads + tid) * E1C_RUN;
                    if (j0 >= n_valid)
                        continue;
                    int acc[E1C_RUN] = {0, 0, 0, 0};
                    uint32_t amb = 0;
                    for (int a = 0; a < nact; a++)
                        amb |= e1_channel_run(&par[a], codes.data(), lut, j0, acc, thr_carr, thr_code, 0, nullptr);
                    if (amb) {
                        stats[2]++;
                        memset(acc, 0, sizeof acc);
                        for (int a = 0; a < nact; a++)
                            e1_channel_run(&par[a], codes.data(), lut, j0, acc, thr_carr, thr_code, 1, &stats[0]);
                    }
                    for (int i = 0; i < E1C_RUN; i++)
                        if (j0 + i < n_valid) {
                            uint32_t w = e1_pack_iq(acc[i]);
                            o[(size_t)(j0 + i) * 2] = (int16_t)(w & 0xffffu);
                            o[(size_t)(j0 + i) * 2 + 1] = (int16_t)(w >> 16);
                        }
                }
        }
    return stats[1] ? -1 : 0;
}
}
