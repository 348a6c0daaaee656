// Forced-include (-include) for the reference's src/galileo-sdr.cpp ONLY, used by
// oracle/Makefile to build oracle/_ref/usrp_galileo_trace.  Test infrastructure.
//
// It changes no arithmetic.  It (1) pulls in the reference's own header first, so the
// header guard (include/galileo-sdr.h:1) turns the source's later #include into a no-op
// and the macros below touch only the body of galileo_task(); (2) re-routes the
// per-sample `get_nanos()` call at src/galileo-sdr.cpp:485 -- which sits exactly at the
// top of the sample loop, after the per-epoch restate at :450-479 -- to a hook that dumps
// the channel state the sample loop is about to consume; (3) zero-fills the `iono`
// struct the reference declares uninitialised on the stack (src/galileo-sdr.cpp:139,
// `vflg` is read at src/iono.cpp:37 but never written) before readRinexV3 fills it.
#pragma once
#ifndef E1_REF_HEADER
#define E1_REF_HEADER "/root/reference/include/galileo-sdr.h" /* patched builds (make ref25 / ref36) point this at their scratch copy */
#endif
#include E1_REF_HEADER

void e1_oracle_trace_hook(int line, int isamp, int iumd, const channel_t *chan, const galtime_t *grx, const int *gain);
int e1_oracle_pinned_rinex(std::vector<ephem_t> eph_vector[MAX_SAT], ionoutc_t *ionoutc, char *fname);

// `gain` is galileo_task()'s local gain[MAX_CHAN] (src/galileo-sdr.cpp:119,477): computed per block, never used (:520-521)
#define get_nanos() (e1_oracle_trace_hook(__LINE__, isamp, iumd, chan, &grx, gain), 0L)
#define readRinexV3(a, b, c) e1_oracle_pinned_rinex(a, b, c)
