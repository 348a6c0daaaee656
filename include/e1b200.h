/* e1b200.h -- C-ABI of the B200-native Galileo E1B/C baseband synthesiser.
 *
 * The reference (harshadms/galileo-sdr-sim) has no plugin/FFI seam: its IQ synthesis is the
 * body of one loop nest inside galileo_task() (src/galileo-sdr.cpp:438-564).  This header is
 * the seam a maintainer would cut there (INTEGRATION.md shows the patch):
 *
 *   reference                                             this ABI
 *   ---------------------------------------------------   -----------------------------------
 *   sim constants SAMP_RATE / NUM_IQ_SAMPLES / MAX_CHAN    e1b200_config      (create)
 *     (include/constants.h:10,75,96)
 *   allocateChannel(): chan[i].prn, .carr_phase            e1b200_set_channel / clear_channel
 *     (src/channel.cpp:69-99, 112-119)
 *   computeCodePhase(): f_carr, f_code, code_phase, ibit   e1b200_restate  -> e1_epoch_rec
 *     (src/gal-sig.cpp:308-347)
 *   chan[i].page + generateINavMsg() called in-loop        e1_epoch_rec.page_cur / page_next
 *     (src/galileo-sdr.cpp:497-506, src/inav-msg.cpp:28)
 *   sample loop + (short) store into iq_buff               e1b200_synth_epochs*
 *     (src/galileo-sdr.cpp:481-539)
 *   chan[i].carr_phase carried across blocks               e1b200_get/set_carrier_phase
 *     (src/galileo-sdr.cpp:531-532)
 *   `code_phase -= 4092; ibit++` count of a block, i.e.     e1b200_code_wraps
 *     whether the in-loop generateINavMsg ran (:491-506)
 *   calloc'd iq_buff (:326)                                 e1b200_host_register / host_alloc
 *   (nothing: one process, one thread)                      e1b200_plan_phases*, e1b200_peer_*: time-axis shards
 *                                                          over the GPUs of a node
 *   (nothing: diagnostics of this library)                 e1b200_get_timing / get_stats / last_error,
 *                                                          e1b200_selftest_any_hit
 *
 * Conventions: plain C linkage, no exceptions cross the boundary, every call returns 0 or a
 * negative E1B200_E* code, the opaque context owns all device memory, the caller owns every
 * host buffer it passes in.  One producer thread per context (the reference's generator is
 * single-threaded too).  There is NO CPU fallback: if no CUDA device is usable, create fails.
 */
#ifndef E1B200_H
#define E1B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define E1_CODE_LEN        4092  /* chips per primary code      (CA_SEQ_LEN_E1,  constants.h:121) */
#define E1_SYM_PER_PAGE    500   /* nav symbols per 2 s page    (N_SYM_PAGE,     constants.h:32)  */
#define E1_SEC_CODE_LEN    25    /* E1-C secondary code length  (constants.h:173,213)             */
#define E1_N_PRN_CODES     50
#define E1_PAGE_BYTES      64    /* 500 symbol bits, bit k -> byte k>>3, bit k&7; rest zero       */
#define E1B200_MAX_CHAN    64    /* upper bound on config.max_chan                                */

/* error codes */
#define E1B200_OK           0
#define E1B200_EINVAL     (-1)   /* bad argument / record                                          */
#define E1B200_ENODEV     (-2)   /* no usable CUDA device (there is no CPU fallback)               */
#define E1B200_ECUDA      (-3)   /* CUDA runtime error, see e1b200_last_error()                    */
#define E1B200_ENOMEM     (-4)
#define E1B200_ESTATE     (-5)   /* call sequence error                                            */

/* e1_epoch_rec.flags */
#define E1_REC_SET_PHASE   1u    /* load carr_phase_init into the slot before sample 0 of this
                                    epoch (what allocateChannel does, src/channel.cpp:98-99)       */

/* One record per (epoch, channel slot): everything the reference's sample loop reads from
 * channel_t (include/structures.h:140-162) at the top of a 0.1 s block, after
 * computeCodePhase().  The carrier phase is NOT in the record: it is integrated across
 * epochs by the synthesiser, exactly like chan[i].carr_phase.                              */
typedef struct e1_epoch_rec {
    int32_t  prn;              /* 1..50, 0 = slot idle this epoch            (channel_t.prn)      */
    int32_t  ibit0;            /* nav symbol index at sample 0, 0..499       (channel_t.ibit)     */
    uint32_t flags;            /* E1_REC_*                                                        */
    int32_t  gain_q7;          /* the reference's gain[i] (path loss x antenna pattern, scaled by 2^7,
                                  src/galileo-sdr.cpp:469-477): applied only with E1B200_CFG_GAIN
                                  (the reference computes it and leaves it unused, :520-521);
                                  0 = unit gain (128)                                             */
    double   code_phase0;      /* chips, [0,4092)                            (channel_t.code_phase)*/
    double   f_code;           /* chips/s                                    (channel_t.f_code)   */
    double   f_carr;           /* Hz, |f_carr| < fs_hz                       (channel_t.f_carr)   */
    double   carr_phase_init;  /* cycles in (-1, 1), used iff E1_REC_SET_PHASE (allocateChannel
                                  stores a fraction in [0,1), src/channel.cpp:98-99)              */
    uint8_t  page_cur[E1_PAGE_BYTES];   /* symbols in force at sample 0      (channel_t.page)     */
    uint8_t  page_next[E1_PAGE_BYTES];  /* symbols after ibit passes 499 inside this epoch        */
} e1_epoch_rec;                /* 176 bytes */

/* Device-side restate input (BASELINE config 4): pseudoranges instead of phases; the kernel
 * evaluates computeCodePhase (src/gal-sig.cpp:308-347) itself.                              */
typedef struct e1_range_rec {
    int32_t  prn;
    uint32_t flags;            /* bits 0-7 E1_REC_*, bits 8-31 gain_q7 (see e1_epoch_rec)          */
    double   rho_prev;         /* chan->rho0.range  [m]                                           */
    double   rho_cur;          /* rho1.range        [m]                                           */
    double   grx_sec;          /* receiver time of this epoch [s of week]                         */
    double   carr_phase_init;
    uint8_t  page_cur[E1_PAGE_BYTES];
    uint8_t  page_next[E1_PAGE_BYTES];
} e1_range_rec;                /* 168 bytes */

typedef struct e1b200_config {
    double   fs_hz;            /* sample rate; the reference uses (double)(float)SAMP_RATE        */
    int32_t  samples_per_epoch;/* NUM_IQ_SAMPLES = fs/10                                          */
    int32_t  max_chan;         /* MAX_CHAN, 1..E1B200_MAX_CHAN                                    */
    int32_t  device;           /* CUDA device ordinal                                             */
    uint32_t flags;            /* E1B200_CFG_*                                                    */
    double   dt_epoch;         /* receiver-time step per epoch; 0 -> reference's 0.100000023142   */
} e1b200_config;

#define E1B200_CFG_SERIAL_PLANNER 1u  /* carrier planner: single chain per channel (debug/compare) */
/* Signal options beyond what the reference transmits.  The reference modulates BOC(1,1) at unit
 * gain into exact integers (sboc(.., 1, 1), src/gal-sig.cpp:224,232; gain[i] unused, :520-521) -- that stays
 * the default and stays bit-exact.  Either flag below selects the FLOAT path: per sample
 *     I = sum_c g_c * m_c * cos_c,   m_c = eB_c - eC_c,
 *     eB = B d (alpha a + beta b),   eC = C s (alpha a - beta b)            (Galileo OS SIS ICD 2.1.2, eq. 7-9)
 * with a / b the BOC(1,1) / BOC(6,1) sub-carrier signs at sub-chip trunc(12 code_phase) (first half of the
 * chip / every even sub-chip negative, the convention of the reference's sboc), accumulated in FP32 and
 * stored as int16 by round-to-nearest-even with saturation.  Phases and table indices are the same exact
 * ones as in the integer path; the accumulate is float, so the criterion against the oracle is +-1 LSB.   */
#define E1B200_CFG_CBOC 2u  /* alpha = sqrt(10/11), beta = sqrt(1/11) instead of BOC(1,1)'s (1, 0)          */
#define E1B200_CFG_GAIN 4u  /* g_c = gain_q7 / 128 of the record instead of 1                               */

typedef struct e1b200_ctx e1b200_ctx;

typedef struct e1b200_timing {
    float plan_ms;             /* planner kernels of the last synth call (CUDA events)            */
    float synth_ms;            /* sample-synthesis kernel(s) of the last synth call               */
    float total_ms;            /* first launch -> last launch complete, incl. copies if any       */
    int32_t kernel_launches;   /* kernels launched by the last synth call                         */
    int32_t synth_launches;
} e1b200_timing;

int  e1b200_create(const e1b200_config *cfg, e1b200_ctx **out);
int  e1b200_destroy(e1b200_ctx *ctx);

/* What allocateChannel() leaves in a slot that the sample loop reads (src/channel.cpp:69-99): the carrier phase.
 * The PRN -- and with it the code tables, which the reference builds per allocation and this library holds for all
 * 50 PRNs -- travels in every record (e1_epoch_rec.prn), as chan[i].prn does in the reference; `prn` here is only
 * range-checked.  Equivalent: E1_REC_SET_PHASE + carr_phase_init in the slot's next record.  clear_channel: the slot's
 * phase is dead state (src/channel.cpp:112-119); records with prn = 0 keep it silent. */
int  e1b200_set_channel(e1b200_ctx *ctx, int slot, int prn, double carr_phase0);
int  e1b200_clear_channel(e1b200_ctx *ctx, int slot);
int  e1b200_get_carrier_phase(e1b200_ctx *ctx, int slot, double *out);
int  e1b200_set_carrier_phase(e1b200_ctx *ctx, int slot, double phase);
/* slots 0 .. n-1 in one copy (the state a time-axis shard hands to the next one)                          */
int  e1b200_get_carrier_phases(e1b200_ctx *ctx, int n, double *out);
int  e1b200_set_carrier_phases(e1b200_ctx *ctx, int n, const double *phases);

/* Host-buffer entry point (what the patched galileo_task() calls): copies recs H2D,
 * synthesises n_epochs * samples_per_epoch samples, copies int16 I/Q D2H into out.
 * recs is [n_epochs][max_chan]; out is [n_epochs * samples_per_epoch * 2] int16.
 * out may also be DEVICE memory -- of this GPU or of a peer (e1b200_peer_open): the copy engines then move
 * every finished slice there (over NVLink for a peer) behind the kernels that synthesise the next one.   */
int  e1b200_synth_epochs(e1b200_ctx *ctx, int n_epochs, const e1_epoch_rec *recs, int16_t *out);

/* Device-resident variant: d_recs and d_out are device pointers on cfg.device.  Runs on the
 * context's stream; returns after the work is enqueued and e1b200_sync() waits for it.
 * A record outside its documented range (code phase, f_code, ibit0, carrier phase / step) makes
 * e1b200_sync() -- for the host entry points the call itself -- return E1B200_EINVAL; with the device
 * entry points the samples of that call have been written by then and must be discarded.     */
int  e1b200_synth_epochs_device(e1b200_ctx *ctx, int n_epochs, const e1_epoch_rec *d_recs, int16_t *d_out);
int  e1b200_sync(e1b200_ctx *ctx);

/* Carrier planner only: advances the slots' carrier phases over n_epochs blocks (host records)
 * without synthesising them -- the value chan[i].carr_phase has after the reference's loop ran
 * over those blocks (src/galileo-sdr.cpp:531-532).  Used for the phase hand-off between
 * time-axis shards and for checkpoint/resume.                                               */
int  e1b200_plan_phases(e1b200_ctx *ctx, int n_epochs, const e1_epoch_rec *recs);
int  e1b200_plan_phases_device(e1b200_ctx *ctx, int n_epochs, const e1_epoch_rec *d_recs); /* device records, asynchronous */

/* Same two, from pseudoranges (restate evaluated on the device).                            */
int  e1b200_synth_ranges(e1b200_ctx *ctx, int n_epochs, const e1_range_rec *recs, int16_t *out);
int  e1b200_synth_ranges_device(e1b200_ctx *ctx, int n_epochs, const e1_range_rec *d_recs, int16_t *d_out);

/* Host restatement of computeCodePhase for callers that keep the restate on the CPU.       */
int  e1b200_restate(double rho_prev, double rho_cur, double dt, double grx_sec,
                    double *f_carr, double *f_code, double *code_phase0, int32_t *ibit0, int32_t *ipage);

/* Number of times the reference's loop takes its code-wrap branch (`code_phase -= 4092; ibit++`,
 * src/galileo-sdr.cpp:491-494) inside one block of n_samp samples that starts from code_phase0 -- exact, same
 * roundings as the loop.  ibit0 + *n_wraps >= 500 is the loop's condition for calling generateINavMsg
 * in-loop (:497-506): the caller uses it to decide whether chan[i].page is replaced by page_next after
 * the block (a wrap that falls exactly on the block's end is NOT counted: the next computeCodePhase absorbs
 * it and the reference keeps the stale page).  Host arithmetic, needs no device or context.              */
int  e1b200_code_wraps(double fs_hz, int32_t n_samp, double code_phase0, double f_code, int32_t *n_wraps);

int  e1b200_get_timing(e1b200_ctx *ctx, e1b200_timing *out);

/* Exactness bookkeeping and launch geometry (for tests, bench and profiles).                */
typedef struct e1b200_stats {
    uint64_t exact_samples;    /* channel-samples the closed form flagged as ambiguous and the
                                  exact walk resolved, cumulative since create                    */
    uint64_t planner_errors;   /* records the planner rejected, cumulative                        */
    uint64_t serial_epochs;    /* channel-epochs the carrier chain had to walk serially           */
    uint64_t hat_epochs;       /* channel-epochs planned in parallel and accepted by the chain    */
    int32_t tile;              /* samples per planner checkpoint / synthesis tile                 */
    int32_t tiles_per_epoch;
    int32_t batch_epochs;      /* epochs per D2H staging slice (host entry points)                */
    int32_t plan_epochs;       /* epochs per planner pass                                         */
    int32_t sm_count, ctas_per_sm, smem_bytes;
    int32_t synth_kernel;      /* which sample-loop kernel this context launches: E1B200_KERNEL_*  */
} e1b200_stats;
#define E1B200_KERNEL_RUN 0    /* e1_synth_kernel<R>: tiles shorter than 8192 samples (fs < ~2.2 MS/s)           */
#define E1B200_KERNEL_CW2 1    /* e1_synth_cw_kernel<2,2>: carry-walked runs, 32 samples per thread              */
#define E1B200_KERNEL_CW4 2    /* e1_synth_cw_kernel<4,3>: carry-walked runs, 64 samples per thread (default)    */
#define E1B200_KERNEL_FLOAT 3  /* e1_synth_float_kernel: E1B200_CFG_CBOC / _GAIN                                 */
#define E1B200_KERNEL_EV 4     /* e1_synth_ev_kernel<teams>: event-driven, fs >= 10 MS/s; teams = synth_kernel >> 8 */
int  e1b200_get_stats(e1b200_ctx *ctx, e1b200_stats *out);
void *e1b200_stream(e1b200_ctx *ctx);          /* cudaStream_t the context launches on           */
const char *e1b200_last_error(e1b200_ctx *ctx);
const char *e1b200_version(void);

/* Diagnostics: the device build of the tile-level ambiguity search (e1_any_hit in csrc/e1_core.h, the
 * arithmetic behind e1_clean_kernel) on caller-supplied cases, so a test can hold it against a literal
 * loop: out[i] = 1 when some j in [0, n[i]) has (a[i] + j d[i]) mod M[i] < L[i].  Host pointers,
 * 5 int64 per case in `cases` (a, d, M, L, n).  No counterpart in the reference. */
int  e1b200_selftest_any_hit(int device, int n_cases, const int64_t *cases, int32_t *out);

/* Time-axis shards on several GPUs of one node (one process per GPU), gather without a collective: the writer
 * rank allocates the whole stream in its HBM (e1b200_peer_alloc) and ships the 64-byte handle to the others (any
 * channel: torch.distributed, a pipe); they e1b200_peer_open it and pass `pointer + 4 * first_sample_of_my_segment`
 * as `out` to e1b200_synth_epochs / _ranges (slices travel by the copy engines over NVLink behind the synthesis
 * kernels: the fast way) or as d_out to e1b200_synth_*_device (the kernel's own stores cross NVLink: works, but 16-byte
 * remote stores run at a fifth of the link).  The reference has no counterpart (single process, single thread). */
#define E1B200_IPC_HANDLE_BYTES 64
int  e1b200_peer_alloc(int device, size_t bytes, void **d_ptr, unsigned char handle[E1B200_IPC_HANDLE_BYTES]);
int  e1b200_peer_open(int device, const unsigned char handle[E1B200_IPC_HANDLE_BYTES], void **d_ptr);
int  e1b200_peer_close(void *d_ptr);   /* a pointer from e1b200_peer_open  */
int  e1b200_peer_free(void *d_ptr);    /* a pointer from e1b200_peer_alloc */

/* pinned host allocation helpers so the reference's fwrite/FIFO memcpy consumers
 * (src/galileo-sdr.cpp:542,588) can stay unchanged while D2H runs at full PCIe rate        */
int  e1b200_host_alloc(void **p, size_t bytes);
int  e1b200_host_free(void *p);
/* ... or pin the buffer the reference calloc()s and free()s itself (:326, :655) in place               */
int  e1b200_host_register(void *p, size_t bytes);
int  e1b200_host_unregister(void *p);

#ifdef __cplusplus
}
#endif
#endif /* E1B200_H */
