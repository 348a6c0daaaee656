/* e1_oracle.h -- CPU restatement of the reference's E1B/C synthesis path.
 *
 * TEST INFRASTRUCTURE.  Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this.  The product (galileo-sdr-sim_b200/) never links it.
 *
 * Parity status: PINNED.  tests/test_oracle_golden.py feeds it the channel-state trace dumped
 * by an unmodified-arithmetic build of the reference (oracle/Makefile -> oracle/_ref/) for
 * BASELINE config 1 and checks all 99 epochs x 260000 samples against the reference's own
 * output file (md5 419622c87f06f4048858bce54df72d29) through per-epoch SHA-256 fixtures in
 * tests/golden/.
 */
#ifndef E1_ORACLE_H
#define E1_ORACLE_H
#include "../include/e1b200.h"
#ifdef __cplusplus
extern "C" {
#endif

/* include/constants.h:216-284 -- 512-entry carrier tables, amplitude 250 */
void e1o_carrier_lut(int cos_out[512], int sin_out[512]);

/* hex_to_binary_converter + sboc(...,1,1): src/gal-sig.cpp:9-233.  out has 8184 entries of +-1. */
void e1o_halfchip_table(int prn, int is_e1c, short *out);

/* computeCodePhase: src/gal-sig.cpp:308-347 */
void e1o_restate(double rho_prev, double rho_cur, double dt, double grx_sec,
                 double *f_carr, double *f_code, double *code_phase0, int *ibit0, int *ipage);

/* Sample loop, src/galileo-sdr.cpp:481-539, over n_epochs blocks.
 * recs[n_epochs][max_chan]; carr_phase[max_chan] is read and updated (chan[i].carr_phase);
 * out[n_epochs*n_samp*2] receives interleaved int16 I,Q.  Serial, one thread.            */
void e1o_synth_epochs(double fs_hz, int n_samp, int max_chan, int n_epochs,
                      const e1_epoch_rec *recs, double *carr_phase, int16_t *out);

/* Same arithmetic, channels split over n_threads host threads (int32 partial sums added at the
 * end of each epoch; integer addition is associative so the result is identical).         */
void e1o_synth_epochs_mt(double fs_hz, int n_samp, int max_chan, int n_epochs,
                         const e1_epoch_rec *recs, double *carr_phase, int16_t *out, int n_threads);

/* The carrier recurrence alone (:531-532), literally, over n_epochs blocks: phases[n_epochs][max_chan] =
 * chan[i].carr_phase at the top of every block; carr_phase[max_chan] is read and updated.  For checkers
 * that want to start the full loop in the middle of a long run.                              */
void e1o_carrier_phases(double fs_hz, int n_samp, int max_chan, int n_epochs, const e1_epoch_rec *recs,
                        double *carr_phase, double *phases, int n_threads);

/* SURVEY 8 f4, PARITY UNPINNED (the reference has no such mode): the same loop with the E1 CBOC(6,1,1/11)
 * sub-carrier of the Galileo OS SIS ICD (alpha = sqrt(10/11), beta = sqrt(1/11); (1, 0) is the reference's
 * BOC(1,1)) at sub-chip (int)(code_phase * 12) and, with use_gain, the record's gain_q7 / 128; double
 * accumulate, nearest-int16 store.  carr_phase is read and updated.                              */
void e1o_synth_epochs_float(double fs_hz, int n_samp, int max_chan, int n_epochs, const e1_epoch_rec *recs,
                            double *carr_phase, int16_t *out, double alpha, double beta, int use_gain, int n_threads);

#ifdef __cplusplus
}
#endif
#endif
