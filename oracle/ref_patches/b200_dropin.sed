# b200_dropin.sed -- the four edits that hand the reference's sample loop to libe1b200.so, as line-addressed
# sed commands on /root/reference/src/galileo-sdr.cpp @ 036185e (a unified diff would carry the 59 deleted
# lines of the reference's loop into this repository; this names them by number instead).  The Makefile
# checks the anchor lines first (b200_dropin.check) so a different revision of the file fails loudly.
#   :28   #include "../include/socket.h"            -> + the binding's header
#   :326  iq_buff = (short *)calloc(...)             -> + create the context, pin iq_buff
#   :481-539  for (isamp ...) { ... }  (sample loop) -> one call per 0.1 s block
#   :655  free(iq_buff);                              -> close the context first
28a\
#include "b200_patch.h" // B200 drop-in: reference-side binding of libe1b200.so (oracle/ref_hooks)
326a\
    e1b200_ref_binding *b200 = e1b200_ref_open(iq_buff, iq_buff_size, (double)TX_SAMPLERATE); // B200 drop-in
481,539c\
        // B200 drop-in: the sample loop (src/galileo-sdr.cpp:481-539 of the original) is one call into libe1b200.so\
        e1b200_ref_block(b200, chan, grx, eph_vector, current_eph, &iono, iq_buff);
655i\
    e1b200_ref_close(b200, iq_buff); // B200 drop-in
