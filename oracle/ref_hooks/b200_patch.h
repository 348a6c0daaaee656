// b200_patch.h -- the reference-side binding of libe1b200.so: what a maintainer of galileo-sdr-sim adds to
// galileo_task() to hand its sample loop (src/galileo-sdr.cpp:481-539) to the B200 library.  It is compiled
// into oracle/_ref/usrp_galileo*_b200 by `make -C oracle refb200` from the reference's own sources plus
// ref_patches/b200_dropin.diff (four changed places in src/galileo-sdr.cpp); INTEGRATION.md quotes both.
// Reference-side code: it includes the reference's header and uses its types.  Not part of the product
// library, not linked into it.
#pragma once
#ifndef E1_REF_HEADER
#define E1_REF_HEADER "/root/reference/include/galileo-sdr.h"
#endif
#include E1_REF_HEADER
#include <vector>

struct e1b200_ref_binding;

// after iq_buff is allocated (src/galileo-sdr.cpp:326): creates the context, pins iq_buff in place
e1b200_ref_binding *e1b200_ref_open(short *iq_buff, int iq_buff_size, double samp_rate);
// replaces the sample loop of one 0.1 s block (:481-539): fills iq_buff, leaves chan[i].page as the loop would
void e1b200_ref_block(e1b200_ref_binding *b, channel_t *chan, galtime_t grx, std::vector<ephem_t> *eph_vector,
                      const std::vector<int> &current_eph, ionoutc_t *iono, short *iq_buff);
// before free(iq_buff) (:655)
void e1b200_ref_close(e1b200_ref_binding *b, short *iq_buff);
