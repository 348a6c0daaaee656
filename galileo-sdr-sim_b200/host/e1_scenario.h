/* e1_scenario.h -- host side of the drop-in: RINEX 3 navigation file -> per-block channel records.
 *
 * This is the part of the reference that stays on the CPU (north star: "host code stays C++:
 * RINEX parse, ephemeris/geodesy, I/NAV page assembly"): everything galileo_task() does around its
 * sample loop (src/galileo-sdr.cpp:185-479, 545-564), restated from the behavioural description in
 * SURVEY.md as one scenario object that emits e1_epoch_rec[max_chan] per 0.1 s block -- the input
 * of the C-ABI in include/e1b200.h.  Floating-point expressions are evaluated in the reference's
 * order (build with -ffp-contract=off) because the records must be bit-identical: a last-bit
 * difference in f_carr or code_phase0 flips table indices in the sample loop.
 *
 * Plain C linkage so the tests can drive it through ctypes; no CUDA in this translation unit.
 */
#ifndef E1_SCENARIO_H
#define E1_SCENARIO_H

#include <stdint.h>

#include "../../include/e1b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct e1h_options {
    char navfile[512];    /* -e  RINEX 3 navigation file                                          */
    double llh[3];        /* -l  latitude, longitude [deg], height [m]   (src/main.cpp:189-191)   */
    int32_t have_start;   /* -t  given                                                             */
    int32_t y, m, d, hh, mm; /* -t  yyyy/mm/dd,hh:mm:ss (seconds floored, src/main.cpp:270)        */
    double sec;
    int32_t iduration;    /* -d  in units of 0.1 s: (int)(seconds*10 + 0.5)  (src/main.cpp:276)   */
    int32_t iono_enable;  /* cleared by -I                                                         */
    int32_t max_chan;     /* MAX_CHAN of the build being mirrored (reference: 16)                 */
    double fs_hz;         /* sample rate the records are for (reference: (float)2.6e6)            */
    int32_t samples_per_epoch;
    int32_t verbose;      /* print the reference's allocation lines to stderr                      */
    double elev_mask_deg; /* 0 = the reference's hard-coded 10 degrees (src/channel.cpp:60); e.g. -90 =
                             every satellite of the file (the patched `ch36` reference build)          */
} e1h_options;

typedef struct e1h_scenario e1h_scenario;

void e1h_default_options(e1h_options *o);

/* Reads the navigation file, fixes the start time, allocates the first channels.
 * Returns NULL (message in err, if given) on failure. */
e1h_scenario *e1h_open(const e1h_options *o, char *err, int err_len);
void e1h_close(e1h_scenario *s);

/* Blocks the scenario will produce: iduration - 1 (src/galileo-sdr.cpp:438). */
int e1h_total_epochs(const e1h_scenario *s);

/* Produces the records of the next n blocks (fewer at the end): recs[n][max_chan].  Returns the
 * number of blocks written.  grx_sec (optional, [n]) receives the receiver time of each block. */
int e1h_next(e1h_scenario *s, int n, e1_epoch_rec *recs, double *grx_sec);

/* Same blocks as pseudoranges (e1_range_rec: rho_prev, rho_cur, receiver time, pages) for the
 * device-side restate, e1b200_synth_ranges (BASELINE configs[3]).  recs and/or ranges may be NULL. */
int e1h_next_ex(e1h_scenario *s, int n, e1_epoch_rec *recs, e1_range_rec *ranges, double *grx_sec);

/* Receiver position.  The reference re-reads `llhr` -- latitude, longitude [deg], height [m], written
 * by its UDP location thread (include/socket.h:69,165-178) -- at the top of every block
 * (src/galileo-sdr.cpp:443-448).  e1h_set_location is that write; e1h_set_motion installs a table
 * indexed by block number iumd = 1 .. iduration-1 (entry 0 is unused, like xyz[0]) that is applied
 * the same way, for reproducible dynamic scenarios (e1sim -u). */
int e1h_set_location(e1h_scenario *s, double lat_deg, double lon_deg, double height_m);
int e1h_set_motion(e1h_scenario *s, int n_blocks, const double *llh_deg);
int e1h_ecef_to_llh_deg(const double *xyz, double *llh_deg);

/* Pieces exposed for the tests. */
int e1h_page_symbols(const e1h_scenario *s, int prn, double grx_sec, int week, int *symbols500);
unsigned int e1h_crc24q_bits(const int *bits, int length);
int e1h_encode_page(const int *even114, const int *odd114, int *symbols500); /* tails + FEC + interleaver + sync */

#ifdef __cplusplus
}
#endif
#endif
