/* e1_fifo.h -- the sample FIFO between the generator and a streaming consumer (radio TX thread).
 *
 * Same contract as the reference's ring (src/fifo.cpp:3-61; producer side src/galileo-sdr.cpp:581-596):
 * complex int16 samples, head/tail in samples, a reader that takes what is there (never blocks, returns
 * the count, handles the wrap with two copies), a writer that waits for room for one whole block.
 * Differences that the B200 path needs: the capacity is a parameter (the reference's is two blocks;
 * the GPU delivers hundreds of blocks per call, so the writer pushes many blocks and the ring is sized
 * for a few batches), the ring can live in pinned memory handed in by the caller so D2H copies land
 * in it directly, and waiting uses one mutex + two condition variables inside the object instead of
 * the reference's globals.  Plain C linkage (tests drive it through ctypes); no CUDA here.          */
#ifndef E1_FIFO_H
#define E1_FIFO_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct e1_fifo e1_fifo;

/* capacity_samples complex samples; storage = NULL lets the FIFO allocate, otherwise the caller's
 * buffer of capacity_samples * 2 int16 (e.g. from e1b200_host_alloc) is used and not freed. */
e1_fifo *e1_fifo_create(size_t capacity_samples, int16_t *storage);
void e1_fifo_destroy(e1_fifo *f);

size_t e1_fifo_sample_length(e1_fifo *f);                       /* get_sample_length (src/fifo.cpp:3)       */
size_t e1_fifo_read(e1_fifo *f, int16_t *buffer, size_t samples); /* fifo_read (:14): min(available, samples)   */
int e1_fifo_write_ready(e1_fifo *f, size_t block_samples);      /* is_fifo_write_ready (:50), room for a block */
/* Blocks until there is room for `samples` (<= capacity), copies them in, wakes the reader.  Returns 0,
 * or -1 if the FIFO was closed while waiting (src/galileo-sdr.cpp:581-596). */
int e1_fifo_write(e1_fifo *f, const int16_t *iq, size_t samples);
/* Blocks until at least `samples` are available or the producer has finished; then reads like
 * e1_fifo_read (what the reference's TX thread gets from its fifo_read_ready wait + fifo_read). */
size_t e1_fifo_read_wait(e1_fifo *f, int16_t *buffer, size_t samples);
void e1_fifo_finish(e1_fifo *f);                                /* producer: generation finished (s->finished) */
int e1_fifo_finished(e1_fifo *f);                               /* is_finished_generation (:48)             */

#ifdef __cplusplus
}
#endif
#endif
