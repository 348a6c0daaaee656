// Trace dump for the reference oracle build (see trace_pre.h).  Test infrastructure.
// One record per (epoch, active slot), written at the first sample of each 0.1 s block:
// exactly the state the loop at src/galileo-sdr.cpp:481-539 consumes.
#ifndef E1_REF_HEADER
#define E1_REF_HEADER "/root/reference/include/galileo-sdr.h" /* patched builds (make ref25 / ref36) point this at their scratch copy */
#endif
#include E1_REF_HEADER
#include <cstdio>
#include <cstdlib>
#include <cstring>

#pragma pack(push, 1)
struct e1_trace_rec {
    int32_t iumd, slot, prn, ibit, ipage, gain;
    double code_phase, f_code, f_carr, carr_phase, grx_sec, rho_range;
    uint8_t page[504];
};
#pragma pack(pop)

static FILE *g_fp = NULL;

void e1_oracle_trace_hook(int line, int isamp, int iumd, const channel_t *chan, const galtime_t *grx, const int *gain)
{
    if (line != 485 || isamp != 0)
        return;
    if (!g_fp) {
        const char *p = getenv("E1_TRACE_OUT");
        g_fp = fopen(p ? p : "e1_trace.bin", "wb");
        if (!g_fp) { perror("E1_TRACE_OUT"); exit(2); }
    }
    for (int i = 0; i < MAX_CHAN; i++) {
        if (chan[i].prn <= 0)
            continue;
        e1_trace_rec r;
        memset(&r, 0, sizeof r);
        r.iumd = iumd; r.slot = i; r.prn = chan[i].prn;
        r.ibit = chan[i].ibit; r.ipage = chan[i].ipage; r.gain = gain[i];
        r.code_phase = chan[i].code_phase; r.f_code = chan[i].f_code; r.f_carr = chan[i].f_carr;
        r.carr_phase = chan[i].carr_phase; r.grx_sec = grx->sec; r.rho_range = chan[i].rho0.range;
        for (int k = 0; k < PAGE_SIZE; k++)
            r.page[k] = (uint8_t)(chan[i].page[k] > 0);
        fwrite(&r, sizeof r, 1, g_fp);
    }
    fflush(g_fp);
}

int e1_oracle_pinned_rinex(std::vector<ephem_t> eph_vector[MAX_SAT], ionoutc_t *ionoutc, char *fname)
{
    int enable = ionoutc->enable;
    memset(ionoutc, 0, sizeof *ionoutc);
    ionoutc->enable = enable;
    return readRinexV3(eph_vector, ionoutc, fname);
}
