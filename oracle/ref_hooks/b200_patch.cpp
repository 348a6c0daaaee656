// b200_patch.cpp -- see b200_patch.h.  One e1b200_synth_epochs call per 0.1 s block (n_epochs = 1: the
// reference's control flow -- location thread, 30 s re-allocation, FIFO hand-off -- is left as it is).
#include "b200_patch.h"

#include <cstdio>
#include <cstdlib>
#include <cstring>

#include "../../include/e1b200.h"

struct e1b200_ref_binding {
    e1b200_ctx *gpu;
    double fs_hz;
    int n_samp;
    int last_prn[MAX_CHAN];
    std::vector<e1_epoch_rec> rec;
};

static void fatal(e1b200_ref_binding *b, const char *what, int rc)
{
    // like the reference's other fatal errors (src/galileo-sdr.cpp:335): message on stderr, exit(1)
    fprintf(stderr, "ERROR: %s failed (%d): %s\n", what, rc, b && b->gpu ? e1b200_last_error(b->gpu) : "no context");
    exit(1);
}

// channel_t::page (500 ints, > 0 means symbol 1; src/galileo-sdr.cpp:517) -> e1_epoch_rec's 64 packed bytes
static void pack_page(const int *page, uint8_t out[E1_PAGE_BYTES])
{
    memset(out, 0, E1_PAGE_BYTES);
    for (int k = 0; k < N_SYM_PAGE; k++)
        if (page[k] > 0)
            out[k >> 3] |= (uint8_t)(1u << (k & 7));
}

e1b200_ref_binding *e1b200_ref_open(short *iq_buff, int iq_buff_size, double samp_rate)
{
    e1b200_ref_binding *b = new e1b200_ref_binding();
    b->fs_hz = samp_rate; // (double)TX_SAMPLERATE: the float-rounded rate delt is computed from (:162)
    b->n_samp = iq_buff_size;
    memset(b->last_prn, 0, sizeof b->last_prn);
    b->rec.resize(MAX_CHAN);
    e1b200_config cfg;
    memset(&cfg, 0, sizeof cfg);
    cfg.fs_hz = samp_rate;
    cfg.samples_per_epoch = iq_buff_size; // NUM_IQ_SAMPLES
    cfg.max_chan = MAX_CHAN;
    cfg.device = 0;
    int rc = e1b200_create(&cfg, &b->gpu);
    if (rc != E1B200_OK)
        fatal(b, "e1b200_create", rc);
    // the reference's calloc'd buffer stays the reference's (fwrite :542, FIFO memcpy :588, free :655); pinned in place
    if (e1b200_host_register(iq_buff, (size_t)2 * iq_buff_size * sizeof(short)) != E1B200_OK)
        fprintf(stderr, "e1b200: iq_buff left pageable (cudaHostRegister failed)\n");
    return b;
}

void e1b200_ref_block(e1b200_ref_binding *b, channel_t *chan, galtime_t grx, std::vector<ephem_t> *eph_vector,
                      const std::vector<int> &current_eph, ionoutc_t *iono, short *iq_buff)
{
    int *next_page[MAX_CHAN];
    for (int i = 0; i < MAX_CHAN; i++) {
        e1_epoch_rec &r = b->rec[i];
        memset(&r, 0, sizeof r);
        next_page[i] = NULL;
        if (chan[i].prn <= 0) {
            b->last_prn[i] = 0;
            continue;
        }
        r.prn = chan[i].prn;
        r.ibit0 = chan[i].ibit; // as computeCodePhase just left them (:464, src/gal-sig.cpp:334-338)
        r.code_phase0 = chan[i].code_phase;
        r.f_code = chan[i].f_code;
        r.f_carr = chan[i].f_carr;
        if (b->last_prn[i] != chan[i].prn) { // slot (re)allocated: allocateChannel set carr_phase (src/channel.cpp:98-99)
            r.flags = E1_REC_SET_PHASE;
            r.carr_phase_init = chan[i].carr_phase;
            b->last_prn[i] = chan[i].prn;
        }
        pack_page(chan[i].page, r.page_cur);
        memcpy(r.page_next, r.page_cur, E1_PAGE_BYTES);
        // Does the loop's symbol counter pass 499 inside this block (:491-506)?  The exact count of its
        // `code_phase -= 4092; ibit++` branch decides (a wrap that lands on the block's end belongs to the next
        // block's restate and does NOT regenerate the page); only asked when the block is near the page's end.
        if (chan[i].ibit + 30 >= N_SYM_PAGE) {
            int32_t wraps = 0;
            int rc = e1b200_code_wraps(b->fs_hz, b->n_samp, chan[i].code_phase, chan[i].f_code, &wraps);
            if (rc != E1B200_OK)
                fatal(b, "e1b200_code_wraps", rc);
            if (chan[i].ibit + wraps >= N_SYM_PAGE) {
                // the call the loop makes at :503-506, with the block's grx, on a copy; its page is adopted below
                channel_t next = chan[i];
                int sv = chan[i].prn - 1;
                ephem_t eph = eph_vector[sv][current_eph[sv]];
                generateINavMsg(grx, &next, &eph, iono);
                next_page[i] = next.page;
                pack_page(next.page, r.page_next);
            }
        }
    }
    int rc = e1b200_synth_epochs(b->gpu, 1, b->rec.data(), iq_buff);
    if (rc != E1B200_OK)
        fatal(b, "e1b200_synth_epochs", rc);
    // What the loop leaves behind in chan[]: ibit / code_phase / ipage are overwritten by the next
    // computeCodePhase before anything reads them; carr_phase lives in the library from here on (it is only
    // ever read by the loop itself); the page is the new one where the loop would have generated it (the old
    // one leaks, as in the reference).
    for (int i = 0; i < MAX_CHAN; i++)
        if (next_page[i])
            chan[i].page = next_page[i];
}

void e1b200_ref_close(e1b200_ref_binding *b, short *iq_buff)
{
    if (!b)
        return;
    e1b200_host_unregister(iq_buff);
    e1b200_destroy(b->gpu);
    delete b;
}
