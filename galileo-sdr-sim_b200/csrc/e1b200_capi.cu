/* e1b200_capi.cu -- the C-ABI declared in include/e1b200.h over the sm_100a kernels.
 *
 * Replaces the body of the reference's 0.1 s block loop (src/galileo-sdr.cpp:481-539): the
 * caller (a patched galileo_task(), or this repo's own host driver) hands over the channel
 * state it computed for each block and gets the int16 I/Q samples back.  No CPU fallback:
 * without a usable CUDA device e1b200_create() fails with E1B200_ENODEV.
 */
#include <cuda_runtime.h>
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "../../include/e1b200.h"
#include "../data/e1_prn_codes.h"
#include "e1_kernels.cuh"

#define E1B200_VERSION "e1b200 0.1 (sm_100a)"

struct e1b200_ctx {
    e1b200_config cfg;
    double delt;        /* 1/fs, as the reference computes it (src/galileo-sdr.cpp:162) */
    int tile;           /* samples per planner checkpoint / synthesis tile              */
    int run;            /* consecutive samples per thread: tile = 512 * run             */
    int elide;          /* mark tiles that cannot hold an ambiguous sample (e1_clean_kernel) so the sample loop skips its tracking */
    int pair;           /* run == 16: 8192-sample tiles through e1_synth_cw_kernel (teams, carry-walked runs) */
    int quad;           /* ... with 64 samples per thread in 3 teams of 128 threads instead of 32 in 2 teams of 256 */
    int evk;            /* event-driven synthesis (e1_synth_ev_kernel): e1_ev_context(fs, run); evk = teams per CTA (2 .. 5), 0 = off */
    int tc_run;         /* samples the code fraction is stepped from one start: sizes the code bias and limits (run, or E1C_EV_RUN) */
    int synth_threads;  /* threads per synthesis CTA */
    int tiles_per_epoch;
    int batch_epochs;   /* epochs per D2H staging buffer (host entry points)            */
    int plan_epochs;    /* epochs per planner pass, at most (bounds scratch)            */
    int scratch_epochs; /* epochs the scratch arrays hold right now (ensure_plan_scratch) */
    int sm_count, ctas_per_sm, smem_bytes;
    int use_bulk, amb_scale, serial_planner;
    int float_path;     /* E1B200_CFG_CBOC / _GAIN: e1_synth_float_kernel (FP32 accumulate, float -> int16 store) */
    float alpha, beta;  /* sub-carrier weights of the float path: (1, 0) BOC(1,1), (sqrt(10/11), sqrt(1/11)) CBOC */
    cudaStream_t stream, copy_stream;
    cudaStream_t side_stream;            /* planner: the code-phase pass runs here, beside the carrier chain */
    cudaEvent_t ev_fork, ev_join;
    std::vector<cudaEvent_t> ev_buf, ev_copy; /* per staging slot: synthesis done / D2H done */
    std::vector<cudaEvent_t> ev;   /* pairs (start, end) of the current call (borrowed from ev_pool) */
    std::vector<cudaEvent_t> ev_pool; /* every timing event this context ever created */
    std::vector<int> ev_kind;      /* 0 = planner pass, 1 = synthesis launch */
    uint32_t *d_codes;
    int32_t *d_lut;
    int32_t *d_lut1;    /* single-copy carrier table of the event-driven kernel */
    double *d_phase;
    unsigned long long *d_counters; /* [0] exact-fallback samples [1] planner errors [2] serial epochs [3] HAT epochs */
    unsigned int *d_next_tile;      /* dynamic tile counter of the synthesis kernel */
    e1_tile_ck *d_ck;
    unsigned char *d_blk;     /* per-tile parameter blocks of the current plan */
    double *d_g, *d_dend, *d_est;
    e1_trans *d_delta;
    e1_prep *d_prep;
    int plan_n;               /* spans in the current plan (stride of the channel-major arrays) */
    e1_span_geo geo;          /* how an epoch is cut into planner spans */
    e1_unit *d_units;
    e1_epoch_rec *d_recs;     /* staging for the host entry points / restate output */
    e1_range_rec *d_ranges;
    int16_t *d_stage;         /* staging ring for the host entry points: n_stage slots of batch_epochs blocks */
    int n_stage;
    e1b200_timing timing;
    unsigned long long counters[4];
    char err[256];
};

static int fail(e1b200_ctx *c, int code, const char *what, cudaError_t ce)
{
    if (c) {
        if (ce != cudaSuccess)
            snprintf(c->err, sizeof c->err, "%s: %s", what, cudaGetErrorString(ce));
        else
            snprintf(c->err, sizeof c->err, "%s", what);
    }
    return code;
}
#define CK(call)                                                                                                       \
    do {                                                                                                               \
        cudaError_t ce_ = (call);                                                                                      \
        if (ce_ != cudaSuccess)                                                                                        \
            return fail(ctx, E1B200_ECUDA, #call, ce_);                                                                \
    } while (0)

/* include/constants.h:216-284: round(250*cos(2*pi*(i+1/2)/512)), except the four entries per
 * table whose exact value is +-105.5, which the reference stores as +-105 */
static void build_lut(int32_t *lut)
{
    static const int cos_fix[4] = {92, 163, 348, 419}, sin_fix[4] = {35, 220, 291, 476};
    int c[512], s[512];
    for (int i = 0; i < 512; i++) {
        double a = 6.283185307179586476925286766559 * ((double)i + 0.5) / 512.0;
        c[i] = (int)lround(250.0 * cos(a));
        s[i] = (int)lround(250.0 * sin(a));
    }
    for (int k = 0; k < 4; k++) {
        c[cos_fix[k]] = c[cos_fix[k]] > 0 ? 105 : -105;
        s[sin_fix[k]] = s[sin_fix[k]] > 0 ? 105 : -105;
    }
    e1_build_lut(c, s, lut);
}

/* src/gal-sig.cpp:9-233 (hex -> chips -> BOC(1,1) half-chips) in the packed layout of e1_core.h */
static void build_codes(uint32_t *codes)
{
    for (int p = 0; p < E1C_N_PRN; p++)
        e1_build_code_words(E1B_PRN_WORDS[p], E1C_PRN_WORDS[p], codes + (size_t)p * E1_CODE_WORDS_PER_PRN);
}

static int env_int(const char *name, int dflt)
{
    const char *v = getenv(name);
    return v && *v ? atoi(v) : dflt;
}

typedef void (*synth_fn)(const e1_synth_args);
typedef void (*ev_fn)(const e1_synth_args, const int32_t *);
static ev_fn ev_for(int teams)
{
    switch (teams) {
    case 5: return e1_synth_ev_kernel<5>;
    case 4: return e1_synth_ev_kernel<4>;
    case 3: return e1_synth_ev_kernel<3>;
    default: return e1_synth_ev_kernel<2>;
    }
}
static synth_fn synth_for(int run, int pair = 0, int quad = 0)
{
    if (pair)
        return quad ? e1_synth_cw_kernel<4, 3> : e1_synth_cw_kernel<2, 2>;
    switch (run) {
    case 4: return e1_synth_kernel<4>;
    case 8: return e1_synth_kernel<8>;
    case 16: return e1_synth_kernel<16>;
    default: return nullptr;
    }
}

extern "C" {

const char *e1b200_version(void) { return E1B200_VERSION; }

const char *e1b200_last_error(e1b200_ctx *ctx) { return ctx ? ctx->err : "null context"; }

int e1b200_create(const e1b200_config *cfg, e1b200_ctx **out)
{
    if (!cfg || !out)
        return E1B200_EINVAL;
    *out = nullptr;
    if (!(cfg->fs_hz > 0.0) || cfg->samples_per_epoch < 1 || cfg->max_chan < 1 || cfg->max_chan > E1B200_MAX_CHAN)
        return E1B200_EINVAL;
    /* one code period must not fit twice in a tile: tile * (1.023e6+margin)/fs < 4092.  Below
       ~2.1 MS/s a sample can span more than one half-chip and every channel takes the generic
       (slow, still exact) form; the reference's only rate is 2.6 MS/s. */
    int run = E1C_MAX_RUN; /* tile = 512 * run samples */
    while (run > 4 && (double)(run * E1_SYNTH_THREADS) * 1.03e6 / cfg->fs_hz >= 4000.0)
        run >>= 1;
    if ((double)(run * E1_SYNTH_THREADS) * 1.03e6 / cfg->fs_hz >= 4000.0)
        return E1B200_EINVAL; /* fs below ~0.53 MS/s */
    run = env_int("E1B200_RUN", run);
    if (!synth_for(run))
        return E1B200_EINVAL;
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev <= cfg->device || cfg->device < 0)
        return E1B200_ENODEV;
    cudaDeviceProp prop;
    if (cudaSetDevice(cfg->device) != cudaSuccess || cudaGetDeviceProperties(&prop, cfg->device) != cudaSuccess)
        return E1B200_ENODEV;
    e1b200_ctx *ctx = new (std::nothrow) e1b200_ctx();
    if (!ctx)
        return E1B200_ENOMEM;
    ctx->cfg = *cfg;
    if (ctx->cfg.dt_epoch == 0.0)
        ctx->cfg.dt_epoch = 0.10000002314200000; /* src/galileo-sdr.cpp:347 */
    ctx->delt = 1.0 / cfg->fs_hz;
    ctx->run = run;
    ctx->float_path = (cfg->flags & (E1B200_CFG_CBOC | E1B200_CFG_GAIN)) != 0;
    ctx->alpha = (cfg->flags & E1B200_CFG_CBOC) ? (float)E1C_ALPHA_CBOC : 1.0f;
    ctx->beta = (cfg->flags & E1B200_CFG_CBOC) ? (float)E1C_BETA_CBOC : 0.0f;
    if (ctx->float_path && run != E1C_MAX_RUN) { /* the float kernel owns 16 samples per thread: 8192-sample tiles */
        delete ctx;
        return E1B200_EINVAL;
    }
    ctx->pair = (run == E1C_MAX_RUN) && !env_int("E1B200_NO_PAIR", 0) && !ctx->float_path;
    ctx->quad = ctx->pair && env_int("E1B200_QUAD", 1);
    ctx->synth_threads = ctx->pair ? (ctx->quad ? 3 * E1_CW_TEAM_THREADS(4) : 2 * E1_CW_TEAM_THREADS(2)) : E1_SYNTH_THREADS;
    ctx->evk = 0;
    if (ctx->pair && e1_ev_context(cfg->fs_hz, run) && !env_int("E1B200_NO_EV", 0)) {
        int want = env_int("E1B200_EV_TEAMS", E1_EV_MAX_TEAMS);
        want = want < 2 ? 2 : (want > E1_EV_MAX_TEAMS ? E1_EV_MAX_TEAMS : want);
        for (int teams = want; teams >= 2 && !ctx->evk; teams--)
            if ((size_t)E1_EV_SMEM(teams, cfg->max_chan) + 256 <= prop.sharedMemPerBlockOptin)
                ctx->evk = teams;
    }
    ctx->tc_run = ctx->evk ? E1C_EV_RUN : run;
    if (ctx->evk)
        ctx->synth_threads = ctx->evk * E1_EV_TEAM_THREADS;
    ctx->tile = run * E1_SYNTH_THREADS;
    ctx->tiles_per_epoch = (cfg->samples_per_epoch + ctx->tile - 1) / ctx->tile;
    ctx->geo = e1_span_geometry(ctx->tiles_per_epoch);
    ctx->use_bulk = env_int("E1B200_NO_TMA", 0) ? 0 : 1;
    ctx->amb_scale = env_int("E1B200_AMB_SCALE", 1);
    if (ctx->amb_scale < 1)
        ctx->amb_scale = 1;
    ctx->elide = !env_int("E1B200_NO_ELIDE", 0);
    ctx->serial_planner = (cfg->flags & E1B200_CFG_SERIAL_PLANNER) || env_int("E1B200_SERIAL_PLANNER", 0);
    const size_t epoch_bytes = (size_t)cfg->samples_per_epoch * 4;
    long be = (long)(((size_t)env_int("E1B200_BATCH_MB", 96) << 20) / epoch_bytes);
    ctx->batch_epochs = be < 1 ? 1 : (be > 512 ? 512 : (int)be);
    const size_t ck_epoch_bytes = (sizeof(e1_tile_ck) * (size_t)cfg->max_chan + e1_blk_bytes(cfg->max_chan)) * (size_t)ctx->tiles_per_epoch;
    long pe = (long)(((size_t)env_int("E1B200_PLAN_MB", 2048) << 20) / ck_epoch_bytes);
    ctx->plan_epochs = pe < 1 ? 1 : (pe > 4096 ? 4096 : (int)pe);
    ctx->sm_count = prop.multiProcessorCount;
    ctx->smem_bytes = E1_CODES_BYTES + E1_LUT_BYTES + (ctx->pair ? (ctx->quad ? 6 : 4) : 2) * (int)e1_blk_bytes(cfg->max_chan);
    if (ctx->evk)
        ctx->smem_bytes = E1_EV_SMEM(ctx->evk, cfg->max_chan);
    *out = ctx; /* from here on errors leave a context the caller can query and destroy */
    CK(cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->copy_stream, cudaStreamNonBlocking));
    CK(cudaStreamCreateWithFlags(&ctx->side_stream, cudaStreamNonBlocking));
    CK(cudaEventCreateWithFlags(&ctx->ev_fork, cudaEventDisableTiming));
    CK(cudaEventCreateWithFlags(&ctx->ev_join, cudaEventDisableTiming));
    synth_fn fn = synth_for(run, ctx->pair, ctx->quad);
    int occ = 0;
    if (ctx->float_path) {
        CK(cudaFuncSetAttribute(e1_synth_float_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_bytes));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, e1_synth_float_kernel, E1_SYNTH_THREADS, ctx->smem_bytes));
    } else if (ctx->evk) {
        ev_fn efn = ev_for(ctx->evk);
        CK(cudaFuncSetAttribute(efn, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_bytes));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, efn, ctx->synth_threads, ctx->smem_bytes));
    } else {
        CK(cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, ctx->smem_bytes));
        CK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, fn, ctx->synth_threads, ctx->smem_bytes));
    }
    if (occ < 1)
        return fail(ctx, E1B200_ECUDA, "synthesis kernel does not fit on this device", cudaSuccess);
    ctx->ctas_per_sm = env_int("E1B200_CTAS_PER_SM", occ);

    std::vector<uint32_t> codes_v(E1_CODES_BYTES / 4);
    std::vector<int32_t> lut_v(E1_LUT_ENTRIES);
    uint32_t *h_codes = codes_v.data();
    int32_t *h_lut = lut_v.data();
    build_codes(h_codes);
    build_lut(h_lut);
    CK(cudaMalloc(&ctx->d_codes, E1_CODES_BYTES));
    CK(cudaMalloc(&ctx->d_lut, E1_LUT_BYTES));
    CK(cudaMalloc(&ctx->d_phase, sizeof(double) * E1B200_MAX_CHAN));
    CK(cudaMalloc(&ctx->d_counters, sizeof ctx->counters));
    CK(cudaMalloc(&ctx->d_next_tile, sizeof(unsigned int)));
    CK(cudaMemcpy(ctx->d_codes, h_codes, E1_CODES_BYTES, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(ctx->d_lut, h_lut, E1_LUT_BYTES, cudaMemcpyHostToDevice));
    if (ctx->evk) {
        std::vector<int32_t> lut1_v(E1_LUT1_BYTES / 4, 0);
        e1_build_lut1(h_lut, lut1_v.data());
        CK(cudaMalloc(&ctx->d_lut1, E1_LUT1_BYTES));
        CK(cudaMemcpy(ctx->d_lut1, lut1_v.data(), E1_LUT1_BYTES, cudaMemcpyHostToDevice));
    }
    CK(cudaMemset(ctx->d_phase, 0, sizeof(double) * E1B200_MAX_CHAN));
    CK(cudaMemset(ctx->d_counters, 0, sizeof ctx->counters));
    return E1B200_OK;
}

int e1b200_destroy(e1b200_ctx *ctx)
{
    if (!ctx)
        return E1B200_EINVAL;
    cudaSetDevice(ctx->cfg.device);
    if (ctx->stream)
        cudaStreamSynchronize(ctx->stream);
    if (ctx->copy_stream)
        cudaStreamSynchronize(ctx->copy_stream);
    cudaFree(ctx->d_codes);
    cudaFree(ctx->d_lut);
    cudaFree(ctx->d_lut1);
    cudaFree(ctx->d_phase);
    cudaFree(ctx->d_counters);
    cudaFree(ctx->d_next_tile);
    cudaFree(ctx->d_ck);
    cudaFree(ctx->d_blk);
    cudaFree(ctx->d_g);
    cudaFree(ctx->d_dend);
    cudaFree(ctx->d_est);
    cudaFree(ctx->d_delta);
    cudaFree(ctx->d_units);
    cudaFree(ctx->d_prep);
    cudaFree(ctx->d_recs);
    cudaFree(ctx->d_ranges);
    cudaFree(ctx->d_stage);
    for (cudaEvent_t ev : ctx->ev_pool)
        cudaEventDestroy(ev);
    for (cudaEvent_t ev : ctx->ev_buf)
        cudaEventDestroy(ev);
    for (cudaEvent_t ev : ctx->ev_copy)
        cudaEventDestroy(ev);
    if (ctx->stream) cudaStreamDestroy(ctx->stream);
    if (ctx->copy_stream) cudaStreamDestroy(ctx->copy_stream);
    if (ctx->side_stream) cudaStreamDestroy(ctx->side_stream);
    if (ctx->ev_fork) cudaEventDestroy(ctx->ev_fork);
    if (ctx->ev_join) cudaEventDestroy(ctx->ev_join);
    delete ctx;
    return E1B200_OK;
}

/* allocateChannel() effects (src/channel.cpp:69-99): the slot's carrier phase is (re)initialised.
 * The PRN itself travels in every epoch record, as chan[i].prn does in the reference. */
int e1b200_set_channel(e1b200_ctx *ctx, int slot, int prn, double carr_phase0)
{
    if (!ctx || slot < 0 || slot >= ctx->cfg.max_chan || prn < 1 || prn > E1_N_PRN_CODES)
        return E1B200_EINVAL;
    return e1b200_set_carrier_phase(ctx, slot, carr_phase0);
}

/* src/channel.cpp:112-119: the slot goes idle; its phase is dead state. */
int e1b200_clear_channel(e1b200_ctx *ctx, int slot)
{
    if (!ctx || slot < 0 || slot >= ctx->cfg.max_chan)
        return E1B200_EINVAL;
    return e1b200_set_carrier_phase(ctx, slot, 0.0);
}

int e1b200_set_carrier_phase(e1b200_ctx *ctx, int slot, double phase)
{
    if (!ctx || slot < 0 || slot >= ctx->cfg.max_chan || !(fabs(phase) < 1.0))
        return E1B200_EINVAL;
    CK(cudaSetDevice(ctx->cfg.device));
    CK(cudaMemcpyAsync(ctx->d_phase + slot, &phase, sizeof phase, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream)); /* &phase is a stack address */
    return E1B200_OK;
}

int e1b200_get_carrier_phase(e1b200_ctx *ctx, int slot, double *out)
{
    if (!ctx || !out || slot < 0 || slot >= ctx->cfg.max_chan)
        return E1B200_EINVAL;
    CK(cudaSetDevice(ctx->cfg.device));
    CK(cudaMemcpyAsync(out, ctx->d_phase + slot, sizeof *out, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return E1B200_OK;
}

/* all slots at once (a time-axis shard's hand-off: one copy instead of max_chan round trips) */
int e1b200_get_carrier_phases(e1b200_ctx *ctx, int n, double *out)
{
    if (!ctx || !out || n < 1 || n > ctx->cfg.max_chan)
        return E1B200_EINVAL;
    CK(cudaSetDevice(ctx->cfg.device));
    CK(cudaMemcpyAsync(out, ctx->d_phase, sizeof(double) * (size_t)n, cudaMemcpyDeviceToHost, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return E1B200_OK;
}

int e1b200_set_carrier_phases(e1b200_ctx *ctx, int n, const double *phases)
{
    if (!ctx || !phases || n < 1 || n > ctx->cfg.max_chan)
        return E1B200_EINVAL;
    for (int i = 0; i < n; i++)
        if (!(fabs(phases[i]) < 1.0))
            return E1B200_EINVAL;
    CK(cudaSetDevice(ctx->cfg.device));
    CK(cudaMemcpyAsync(ctx->d_phase, phases, sizeof(double) * (size_t)n, cudaMemcpyHostToDevice, ctx->stream));
    CK(cudaStreamSynchronize(ctx->stream));
    return E1B200_OK;
}

/* Planner scratch (checkpoints, parameter blocks, per-span arrays) and the record staging of the host entry points,
 * for passes of up to scratch_epochs blocks.  Sized by the calls that arrive, not by the largest pass the context may
 * ever run: the real-time call shape (one block per call) gets along with 64 blocks' worth (a few MB) instead of
 * plan_epochs' (2 GiB); a larger call grows it, at least doubling, up to plan_epochs. */
static int ensure_plan_scratch(e1b200_ctx *ctx, int n_epochs)
{
    int want = n_epochs < 64 ? 64 : n_epochs;
    if (want > ctx->plan_epochs)
        want = ctx->plan_epochs;
    if (want <= ctx->scratch_epochs)
        return E1B200_OK;
    if (want < 2 * ctx->scratch_epochs)
        want = 2 * ctx->scratch_epochs < ctx->plan_epochs ? 2 * ctx->scratch_epochs : ctx->plan_epochs;
    /* work of an earlier asynchronous call may still be using the old arrays */
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(ctx->side_stream));
    CK(cudaStreamSynchronize(ctx->copy_stream));
    void **arrays[] = {(void **)&ctx->d_ck, (void **)&ctx->d_blk, (void **)&ctx->d_delta, (void **)&ctx->d_g, (void **)&ctx->d_dend,
                       (void **)&ctx->d_est, (void **)&ctx->d_units, (void **)&ctx->d_prep, (void **)&ctx->d_recs, (void **)&ctx->d_ranges};
    for (void **a : arrays) {
        cudaFree(*a);
        *a = nullptr;
    }
    ctx->scratch_epochs = 0;
    const size_t nec = (size_t)want * ctx->cfg.max_chan;
    const size_t ne = nec * ctx->geo.spans_per_epoch; /* planner units */
    CK(cudaMalloc(&ctx->d_ck, sizeof(e1_tile_ck) * nec * ctx->tiles_per_epoch));
    CK(cudaMalloc(&ctx->d_blk, e1_blk_bytes(ctx->cfg.max_chan) * (size_t)want * ctx->tiles_per_epoch));
    CK(cudaMalloc(&ctx->d_delta, sizeof(e1_trans) * ne));
    CK(cudaMemset(ctx->d_delta, 0, sizeof(e1_trans) * ne));
    if (!ctx->serial_planner) {
        CK(cudaMalloc(&ctx->d_g, sizeof(double) * ne));
        CK(cudaMalloc(&ctx->d_dend, sizeof(double) * ne));
        CK(cudaMalloc(&ctx->d_est, sizeof(double) * ne));
        CK(cudaMalloc(&ctx->d_units, sizeof(e1_unit) * ne));
        CK(cudaMalloc(&ctx->d_prep, sizeof(e1_prep) * ne));
    }
    ctx->scratch_epochs = want;
    return E1B200_OK;
}

/* timing marks of a call: the events are created once and reused call after call (ev_pool), so the real-time call
   shape pays no cudaEventCreate / cudaEventDestroy per block */
static int mark(e1b200_ctx *ctx, int kind, int end)
{
    cudaEvent_t ev;
    if (ctx->ev.size() < ctx->ev_pool.size()) {
        ev = ctx->ev_pool[ctx->ev.size()];
    } else {
        CK(cudaEventCreate(&ev));
        ctx->ev_pool.push_back(ev);
    }
    ctx->ev.push_back(ev);
    if (!end)
        ctx->ev_kind.push_back(kind);
    CK(cudaEventRecord(ev, ctx->stream));
    return E1B200_OK;
}

static void reset_call(e1b200_ctx *ctx)
{
    ctx->ev.clear();
    ctx->ev_kind.clear();
    memset(&ctx->timing, 0, sizeof ctx->timing);
}

/* planner kernels for n (<= plan_epochs) epochs whose records are on the device */
static int enqueue_plan(e1b200_ctx *ctx, int n, const e1_epoch_rec *d_recs, int carrier_only = 0)
{
    const e1b200_config *cfg = &ctx->cfg;
    int rc = mark(ctx, 0, 0);
    if (rc)
        return rc;
    const int nthr = n * cfg->max_chan;
    /* Span geometry of this pass.  A span is walked by ONE thread: the create-time geometry (8 tiles or more per span)
       is what a large pass wants, but a pass of a few blocks -- the real-time call shape, one 0.1 s block per call --
       has far fewer spans than the GPU has lanes and then the walk of one 65 536-sample span is its latency (0.5 ms).
       Such a pass (up to E1B200_SMALL_PASS = 8 blocks) gets shorter spans, down to one tile, as long as they fit the scratch. */
    e1_span_geo geo = ctx->geo;
    {
        const long cap = (long)ctx->scratch_epochs * ctx->geo.spans_per_epoch; /* units per channel the scratch holds */
        long units = (long)n * geo.spans_per_epoch * cfg->max_chan;
        const long fill = (long)ctx->sm_count * 128;
        const int small_max = env_int("E1B200_SMALL_PASS", 8); /* blocks; larger passes keep the long spans (their reach back to a
                                                                 wrap to anchor on -- 64 spans -- matters at low Doppler) */
        while (geo.span_tiles > 1 && n <= small_max && units < fill && !env_int("E1B200_COARSE_SPANS", 0)) {
            const int st = (geo.span_tiles + 1) / 2, sp = (ctx->tiles_per_epoch + st - 1) / st;
            if ((long)n * sp > cap)
                break;
            geo.span_tiles = st;
            geo.spans_per_epoch = sp;
            units = (long)n * sp * cfg->max_chan;
        }
    }
    const int n_units = n * geo.spans_per_epoch, nuthr = n_units * cfg->max_chan;
    ctx->plan_n = n_units;
    const bool small_pass = geo.span_tiles != ctx->geo.span_tiles;
    /* The code-phase pass writes the code fields of the tile checkpoints, the carrier passes only .phi:
       the two are independent until finalize.  With the parallel planner it runs on a side stream,
       released when the span pass is done, so that it fills the SMs the carrier chain (one block per
       channel) leaves idle. */
    const bool code_beside_chain = !carrier_only && !ctx->serial_planner;
    if (code_beside_chain && small_pass) {
        /* a small pass is a chain of short dependent kernels: the code-phase pass (one thread per block and channel, the
           longest of them) runs beside ALL of the carrier passes, not just beside the chain */
        CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
        CK(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
        e1_plan_code_kernel<<<(nthr + 31) / 32, 32, 0, ctx->side_stream>>>(d_recs, ctx->d_ck, n, cfg->max_chan, cfg->samples_per_epoch,
                                                                         ctx->tile, ctx->tiles_per_epoch, ctx->delt);
        CK(cudaEventRecord(ctx->ev_join, ctx->side_stream));
    }
    if (carrier_only) {
        e1_validate_carrier_kernel<<<(nthr + 127) / 128, 128, 0, ctx->stream>>>(d_recs, nthr, ctx->delt, ctx->d_counters);
        ctx->timing.kernel_launches += 1;
    }
    if (!carrier_only && !code_beside_chain)
        e1_plan_code_kernel<<<(nthr + 127) / 128, 128, 0, ctx->stream>>>(d_recs, ctx->d_ck, n, cfg->max_chan, cfg->samples_per_epoch,
                                                                         ctx->tile, ctx->tiles_per_epoch, ctx->delt);
    if (ctx->serial_planner) {
        e1_plan_carr_kernel<<<(cfg->max_chan + 31) / 32, 32, 0, ctx->stream>>>(d_recs, ctx->d_ck, ctx->d_phase, n, cfg->max_chan,
                                                                             cfg->samples_per_epoch, ctx->tile,
                                                                             ctx->tiles_per_epoch, ctx->delt);
        ctx->timing.kernel_launches += carrier_only ? 1 : 2;
        /* the serial planner writes final checkpoints: translations are zero */
        CK(cudaMemsetAsync(ctx->d_delta, 0, sizeof(e1_trans) * (size_t)nuthr, ctx->stream));
    } else {
        e1_plan_args P;
        P.recs = d_recs;
        P.ck = ctx->d_ck;
        P.phase = ctx->d_phase;
        P.prep = ctx->d_prep;
        P.g = ctx->d_g;
        P.dend = ctx->d_dend;
        P.est = ctx->d_est;
        P.delta = ctx->d_delta;
        P.units = ctx->d_units;
        P.counters = ctx->d_counters;
        P.delt = ctx->delt;
        P.n_epochs = n;
        P.n_samp = cfg->samples_per_epoch;
        P.max_chan = cfg->max_chan;
        P.tile = ctx->tile;
        P.tiles_per_epoch = ctx->tiles_per_epoch;
        P.geo = geo;
        P.n_units = n_units;
        const int cb = cfg->max_chan; /* one channel per block */
        e1_v2_prep_kernel<<<(nthr + 127) / 128, 128, 0, ctx->stream>>>(P);
        e1_v2_ideal_kernel<<<cb, E1_SERIAL_THREADS, 0, ctx->stream>>>(P);
        e1_v2_drift_kernel<<<(nuthr + 63) / 64, 64, 0, ctx->stream>>>(P);
        e1_v2_estimate_kernel<<<cb, E1_SERIAL_THREADS, 0, ctx->stream>>>(P);
        e1_v2_span_kernel<<<(nuthr + 63) / 64, 64, 0, ctx->stream>>>(P);
        if (code_beside_chain && !small_pass) {
            CK(cudaEventRecord(ctx->ev_fork, ctx->stream));
            CK(cudaStreamWaitEvent(ctx->side_stream, ctx->ev_fork, 0));
            e1_plan_code_kernel<<<(nthr + 127) / 128, 128, 0, ctx->side_stream>>>(d_recs, ctx->d_ck, n, cfg->max_chan, cfg->samples_per_epoch,
                                                                                  ctx->tile, ctx->tiles_per_epoch, ctx->delt);
            CK(cudaEventRecord(ctx->ev_join, ctx->side_stream));
        }
        e1_v2_chain_kernel<<<cb, E1_CHAIN_THREADS, 0, ctx->stream>>>(P);
        if (code_beside_chain)
            CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_join, 0));
        ctx->timing.kernel_launches += 7;
    }
    if (!carrier_only) {
        e1_finalize_args F;
        F.recs = d_recs;
        F.ck = ctx->d_ck;
        F.delta = ctx->d_delta;
        F.delta_stride = n_units;
        F.geo = geo;
        F.blk = ctx->d_blk;
        F.counters = ctx->d_counters;
        F.delt = ctx->delt;
        F.n_epochs = n;
        F.max_chan = cfg->max_chan;
        F.tile = ctx->tile;
        F.tiles_per_epoch = ctx->tiles_per_epoch;
        F.tc_code = e1_tc_code(e1_thr_code(ctx->tile, ctx->amb_scale), ctx->tc_run);
        F.cfg_flags = ctx->cfg.flags | (ctx->evk ? E1_INT_EV : 0u);
        const long tiles = (long)n * ctx->tiles_per_epoch;
        e1_finalize_kernel<<<(unsigned)((tiles + 3) / 4), 128, 0, ctx->stream>>>(F);
        ctx->timing.kernel_launches += 1;
        if (ctx->pair && ctx->elide) { /* tiles without a near-boundary sample: the sample loop drops its tracking */
            e1_clean_args C;
            C.blk = ctx->d_blk;
            C.n_tiles = tiles;
            C.max_chan = cfg->max_chan;
            C.tile = ctx->tile;
            const uint32_t thr_carr = e1_thr_carr(ctx->tile, ctx->amb_scale), thr_code = e1_thr_code(ctx->tile, ctx->amb_scale);
            C.tc_carr = e1_tc_carr(thr_carr, ctx->run);
            C.lim_carr = e1_lim_carr(C.tc_carr, thr_carr);
            C.cw_samples = ctx->evk ? E1C_EV_RUN : (ctx->quad ? E1_CW_RUN(4) : E1_CW_RUN(2));
            C.code_run = ctx->evk ? E1C_EV_RUN : E1C_MAX_RUN;
            C.lim_code = e1_lim_code(F.tc_code, thr_code);
            C.thr_code = thr_code;
            e1_clean_kernel<<<(unsigned)((tiles * cfg->max_chan + 127) / 128), 128, 0, ctx->stream>>>(C);
            ctx->timing.kernel_launches += 1;
        }
    }
    CK(cudaGetLastError());
    return mark(ctx, 0, 1);
}

/* synthesis of epochs [e_off, e_off+n) of the current plan into d_out */
static int enqueue_synth(e1b200_ctx *ctx, int e_off, int n, const e1_epoch_rec *d_recs_plan, int16_t *d_out)
{
    const e1b200_config *cfg = &ctx->cfg;
    e1_synth_args A;
    (void)d_recs_plan;
    A.blk = ctx->d_blk + e1_blk_bytes(cfg->max_chan) * (size_t)e_off * ctx->tiles_per_epoch;
    A.codes = ctx->d_codes;
    A.lut = ctx->d_lut;
    A.out = d_out;
    A.counters = ctx->d_counters;
    A.next_tile = ctx->d_next_tile;
    A.n_epochs = n;
    A.n_samp = cfg->samples_per_epoch;
    A.max_chan = cfg->max_chan;
    A.tile = ctx->tile;
    A.tiles_per_epoch = ctx->tiles_per_epoch;
    A.thr_carr = e1_thr_carr(ctx->tile, ctx->amb_scale);
    A.thr_code = e1_thr_code(ctx->tile, ctx->amb_scale);
    A.tc_carr = e1_tc_carr(A.thr_carr, ctx->run);
    A.tc_code = e1_tc_code(A.thr_code, ctx->tc_run);
    A.vec_ok = (cfg->samples_per_epoch % 4 == 0) && (((uintptr_t)d_out & 15u) == 0);
    A.use_bulk = ctx->use_bulk;
    long total_tiles = (long)n * ctx->tiles_per_epoch;
    long grid = (long)ctx->sm_count * ctx->ctas_per_sm;
    const int teams = ctx->evk ? ctx->evk : (ctx->pair ? (ctx->quad ? 3 : 2) : 1);
    const long cta_tiles = (total_tiles + teams - 1) / teams; /* a CTA of the team kernel starts on one tile per team */
    if (grid > cta_tiles)
        grid = cta_tiles;
    CK(cudaMemsetAsync(ctx->d_next_tile, 0, sizeof(unsigned int), ctx->stream));
    int rc = mark(ctx, 1, 0);
    if (rc)
        return rc;
    if (ctx->float_path)
        e1_synth_float_kernel<<<(unsigned)grid, E1_SYNTH_THREADS, ctx->smem_bytes, ctx->stream>>>(A, ctx->alpha, ctx->beta);
    else if (ctx->evk)
        ev_for(ctx->evk)<<<(unsigned)grid, ctx->synth_threads, ctx->smem_bytes, ctx->stream>>>(A, ctx->d_lut1);
    else
        synth_for(ctx->run, ctx->pair, ctx->quad)<<<(unsigned)grid, ctx->synth_threads, ctx->smem_bytes, ctx->stream>>>(A);
    CK(cudaGetLastError());
    ctx->timing.kernel_launches += 1;
    ctx->timing.synth_launches += 1;
    return mark(ctx, 1, 1);
}

static int finish_timing(e1b200_ctx *ctx)
{
    float plan = 0, synth = 0, total = 0;
    for (size_t i = 0; i < ctx->ev_kind.size(); i++) {
        float t = 0;
        CK(cudaEventElapsedTime(&t, ctx->ev[2 * i], ctx->ev[2 * i + 1]));
        (ctx->ev_kind[i] ? synth : plan) += t;
    }
    if (!ctx->ev.empty())
        CK(cudaEventElapsedTime(&total, ctx->ev.front(), ctx->ev.back()));
    ctx->timing.plan_ms = plan;
    ctx->timing.synth_ms = synth;
    ctx->timing.total_ms = total;
    unsigned long long before = ctx->counters[1];
    CK(cudaMemcpy(ctx->counters, ctx->d_counters, sizeof ctx->counters, cudaMemcpyDeviceToHost));
    if (ctx->counters[1] != before)
        return fail(ctx, E1B200_EINVAL, "planner rejected a record (code phase / f_code / ibit out of range, |carr_phase_init| >= 1, |f_carr / fs| >= 1, or fs too low for the tile)",
                    cudaSuccess);
    return E1B200_OK;
}

int e1b200_synth_epochs_device(e1b200_ctx *ctx, int n_epochs, const e1_epoch_rec *d_recs, int16_t *d_out)
{
    if (!ctx || n_epochs < 0 || (n_epochs && (!d_recs || !d_out)))
        return E1B200_EINVAL;
    CK(cudaSetDevice(ctx->cfg.device));
    int rc = ensure_plan_scratch(ctx, n_epochs);
    if (rc)
        return rc;
    reset_call(ctx);
    const size_t epoch_i16 = (size_t)ctx->cfg.samples_per_epoch * 2;
    for (int e0 = 0; e0 < n_epochs; e0 += ctx->plan_epochs) {
        int n = n_epochs - e0 < ctx->plan_epochs ? n_epochs - e0 : ctx->plan_epochs;
        const e1_epoch_rec *r = d_recs + (size_t)e0 * ctx->cfg.max_chan;
        if ((rc = enqueue_plan(ctx, n, r)) || (rc = enqueue_synth(ctx, 0, n, r, d_out + (size_t)e0 * epoch_i16)))
            return rc;
    }
    return E1B200_OK;
}

int e1b200_sync(e1b200_ctx *ctx)
{
    if (!ctx)
        return E1B200_EINVAL;
    CK(cudaSetDevice(ctx->cfg.device));
    CK(cudaStreamSynchronize(ctx->stream));
    CK(cudaStreamSynchronize(ctx->copy_stream));
    if (!ctx->ev.empty())
        return finish_timing(ctx);
    return E1B200_OK;
}

static int ensure_staging(e1b200_ctx *ctx, int want_ranges, int want_out)
{
    const e1b200_config *cfg = &ctx->cfg;
    size_t nrec = (size_t)ctx->scratch_epochs * cfg->max_chan; /* after ensure_plan_scratch: it frees these with the rest when it grows */
    if (!ctx->d_recs)
        CK(cudaMalloc(&ctx->d_recs, nrec * sizeof(e1_epoch_rec)));
    if (want_ranges && !ctx->d_ranges)
        CK(cudaMalloc(&ctx->d_ranges, nrec * sizeof(e1_range_rec)));
    if (want_out > ctx->n_stage) {
        /* a deep ring (up to 4 GiB of the 180 GB by default) lets the kernels run far ahead of the PCIe
           copies, so a planner pass between two synthesis slices never starves the copy engine */
        const size_t slot = (size_t)ctx->batch_epochs * cfg->samples_per_epoch * 4;
        long cap = (long)(((size_t)env_int("E1B200_STAGE_MB", 4096) << 20) / slot);
        cap = cap < 2 ? 2 : (cap > 64 ? 64 : cap);
        int ns = want_out < 2 ? 2 : (want_out > cap ? (int)cap : want_out);
        if (ns > ctx->n_stage) {
            CK(cudaStreamSynchronize(ctx->copy_stream));
            CK(cudaFree(ctx->d_stage));
            ctx->d_stage = nullptr;
            ctx->n_stage = 0;
            CK(cudaMalloc(&ctx->d_stage, slot * ns));
            while ((int)ctx->ev_buf.size() < ns) {
                cudaEvent_t a, b;
                CK(cudaEventCreateWithFlags(&a, cudaEventDisableTiming));
                ctx->ev_buf.push_back(a);
                CK(cudaEventCreateWithFlags(&b, cudaEventDisableTiming));
                ctx->ev_copy.push_back(b);
            }
            ctx->n_stage = ns;
        }
    }
    return E1B200_OK;
}

/* Host-buffer pipeline.  Per planner pass: records H2D, (restate,) planner kernels; then the
 * synthesis runs in slices of batch_epochs into a ring of staging slots, each slice's D2H on
 * `copy_stream` overlapping the following kernels. */
static int synth_host(e1b200_ctx *ctx, int n_epochs, const e1_epoch_rec *recs, const e1_range_rec *ranges, int16_t *out)
{
    if (!ctx || n_epochs < 0 || (n_epochs && ((!recs && !ranges) || !out)))
        return E1B200_EINVAL;
    CK(cudaSetDevice(ctx->cfg.device));
    int rc = ensure_plan_scratch(ctx, n_epochs);
    if (!rc)
    {
        /* ring slots: a job of many slices wants the kernels far ahead of the copies; a call of one or two slices (the
           real-time shape) gets one spare slot, not eight (a slot is E1B200_BATCH_MB of device memory) */
        const int slices = n_epochs ? (n_epochs + ctx->batch_epochs - 1) / ctx->batch_epochs : 0;
        rc = ensure_staging(ctx, ranges != nullptr, slices ? (slices <= 2 ? slices + 1 : slices + 8) : 0);
    }
    if (rc)
        return rc;
    reset_call(ctx);
    const e1b200_config *cfg = &ctx->cfg;
    const size_t epoch_i16 = (size_t)cfg->samples_per_epoch * 2;
    int slice = 0;
    const int trace = env_int("E1B200_TRACE", 0);
    std::vector<cudaEvent_t> tr_ev;
    /* planner passes grow geometrically (256, 640, 1600, ... plan_epochs blocks): the first D2H starts
       after a short pass instead of after the whole job's planning, and each pass's copies outlast the
       next pass's planning (a pass costs ~5 ms of dependent walks however small it is, a block's D2H
       ~20 us at 2.6 MS/s); the copies, not the kernels, bound this entry point */
    /* ... the first pass holds about 256 MB of output (256 blocks at 2.6 MS/s, 26 at 25 MS/s), unless E1B200_FIRST_PASS says otherwise */
    int chunk = env_int("E1B200_FIRST_PASS", 0);
    if (chunk < 1) {
        chunk = (int)(((size_t)256 * 1040000) / ((size_t)cfg->samples_per_epoch * 4));
        chunk = chunk < 9 ? 9 : (chunk > 256 ? 256 : chunk);
    }
    {
        /* a destination in DEVICE memory (this GPU's or a peer's, e1b200_peer_open) is fed over NVLink or HBM, an order
           of magnitude faster than PCIe: there the kernels bound the call, not the copies, and every extra planner pass
           (about 1 ms of dependent launches whatever its size) shows -- plan as much as the scratch holds at once */
        cudaPointerAttributes pa;
        if (out && cudaPointerGetAttributes(&pa, out) == cudaSuccess && pa.type == cudaMemoryTypeDevice)
            chunk = ctx->plan_epochs;
        else
            (void)cudaGetLastError();
    }
    for (int p0 = 0, np = 0; p0 < n_epochs; p0 += np) {
        np = chunk < ctx->plan_epochs ? chunk : ctx->plan_epochs;
        if (np > n_epochs - p0)
            np = n_epochs - p0;
        chunk = chunk > (1 << 20) ? chunk : chunk * 5 / 2;
        size_t nrec = (size_t)np * cfg->max_chan;
        if (ranges) {
            CK(cudaMemcpyAsync(ctx->d_ranges, ranges + (size_t)p0 * cfg->max_chan, nrec * sizeof(e1_range_rec),
                               cudaMemcpyHostToDevice, ctx->stream));
            e1_restate_kernel<<<(unsigned)((nrec + 127) / 128), 128, 0, ctx->stream>>>(ctx->d_ranges, ctx->d_recs, (int)nrec,
                                                                                        cfg->dt_epoch);
            ctx->timing.kernel_launches += 1;
        } else {
            CK(cudaMemcpyAsync(ctx->d_recs, recs + (size_t)p0 * cfg->max_chan, nrec * sizeof(e1_epoch_rec),
                               cudaMemcpyHostToDevice, ctx->stream));
        }
        if ((rc = enqueue_plan(ctx, np, ctx->d_recs)))
            return rc;
        for (int e0 = 0; e0 < np; e0 += ctx->batch_epochs, slice++) {
            int n = np - e0 < ctx->batch_epochs ? np - e0 : ctx->batch_epochs;
            int b = slice % ctx->n_stage;
            int16_t *d_slot = ctx->d_stage + (size_t)b * ctx->batch_epochs * epoch_i16;
            if (slice >= ctx->n_stage) /* slot b is free once its previous D2H finished */
                CK(cudaStreamWaitEvent(ctx->stream, ctx->ev_copy[b], 0));
            if ((rc = enqueue_synth(ctx, e0, n, ctx->d_recs, d_slot)))
                return rc;
            CK(cudaEventRecord(ctx->ev_buf[b], ctx->stream));
            CK(cudaStreamWaitEvent(ctx->copy_stream, ctx->ev_buf[b], 0));
            cudaEvent_t t0 = nullptr, t1 = nullptr;
            if (trace) {
                CK(cudaEventCreate(&t0));
                CK(cudaEventCreate(&t1));
                CK(cudaEventRecord(t0, ctx->copy_stream));
            }
            /* cudaMemcpyDefault: `out` is normally the caller's (pinned) host buffer -- a D2H over PCIe -- but may as well be
               device memory of this or of a PEER GPU (e1b200_peer_open): then the copy engines carry the slice over NVLink
               while the SMs synthesise the next one (time-axis shards: the gather without a collective) */
            CK(cudaMemcpyAsync(out + (size_t)(p0 + e0) * epoch_i16, d_slot, (size_t)n * epoch_i16 * 2,
                               cudaMemcpyDefault, ctx->copy_stream));
            CK(cudaEventRecord(ctx->ev_copy[b], ctx->copy_stream));
            if (trace) {
                CK(cudaEventRecord(t1, ctx->copy_stream));
                tr_ev.push_back(t0);
                tr_ev.push_back(t1);
            }
        }
    }
    rc = e1b200_sync(ctx);
    if (trace && !ctx->ev.empty()) { /* E1B200_TRACE=1: timeline of the call on stderr, ms since the first kernel */
        for (size_t i = 0; i < ctx->ev_kind.size(); i++) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, ctx->ev.front(), ctx->ev[2 * i]);
            cudaEventElapsedTime(&b, ctx->ev.front(), ctx->ev[2 * i + 1]);
            fprintf(stderr, "e1b200 trace: %-5s %8.3f .. %8.3f ms\n", ctx->ev_kind[i] ? "synth" : "plan", a, b);
        }
        for (size_t i = 0; i + 1 < tr_ev.size(); i += 2) {
            float a = 0, b = 0;
            cudaEventElapsedTime(&a, ctx->ev.front(), tr_ev[i]);
            cudaEventElapsedTime(&b, ctx->ev.front(), tr_ev[i + 1]);
            fprintf(stderr, "e1b200 trace: d2h   %8.3f .. %8.3f ms\n", a, b);
        }
    }
    for (cudaEvent_t ev : tr_ev)
        cudaEventDestroy(ev);
    return rc;
}

int e1b200_synth_epochs(e1b200_ctx *ctx, int n_epochs, const e1_epoch_rec *recs, int16_t *out)
{
    return synth_host(ctx, n_epochs, recs, nullptr, out);
}

int e1b200_synth_ranges(e1b200_ctx *ctx, int n_epochs, const e1_range_rec *recs, int16_t *out)
{
    return synth_host(ctx, n_epochs, nullptr, recs, out);
}

/* Carrier planner only: what chan[i].carr_phase would be after these blocks (src/galileo-sdr.cpp:531-532
 * integrated over n_epochs * samples_per_epoch samples), without synthesising them.  The hand-off of a
 * time-axis shard: rank r plans its segment, passes the phases on, then synthesises. */
int e1b200_plan_phases(e1b200_ctx *ctx, int n_epochs, const e1_epoch_rec *recs)
{
    if (!ctx || n_epochs < 0 || (n_epochs && !recs))
        return E1B200_EINVAL;
    CK(cudaSetDevice(ctx->cfg.device));
    int rc = ensure_plan_scratch(ctx, n_epochs);
    if (!rc)
        rc = ensure_staging(ctx, 0, 0);
    if (rc)
        return rc;
    reset_call(ctx);
    for (int p0 = 0; p0 < n_epochs; p0 += ctx->plan_epochs) {
        const int np = n_epochs - p0 < ctx->plan_epochs ? n_epochs - p0 : ctx->plan_epochs;
        CK(cudaMemcpyAsync(ctx->d_recs, recs + (size_t)p0 * ctx->cfg.max_chan, (size_t)np * ctx->cfg.max_chan * sizeof(e1_epoch_rec),
                           cudaMemcpyHostToDevice, ctx->stream));
        if ((rc = enqueue_plan(ctx, np, ctx->d_recs, 1)))
            return rc;
        CK(cudaStreamSynchronize(ctx->stream)); /* recs may be pageable: the copy must be done before the caller reuses it */
    }
    return e1b200_sync(ctx);
}

/* the same from records that are already on the device; asynchronous like the other *_device entry points */
int e1b200_plan_phases_device(e1b200_ctx *ctx, int n_epochs, const e1_epoch_rec *d_recs)
{
    if (!ctx || n_epochs < 0 || (n_epochs && !d_recs))
        return E1B200_EINVAL;
    CK(cudaSetDevice(ctx->cfg.device));
    int rc = ensure_plan_scratch(ctx, n_epochs);
    if (rc)
        return rc;
    reset_call(ctx);
    for (int p0 = 0; p0 < n_epochs; p0 += ctx->plan_epochs) {
        const int np = n_epochs - p0 < ctx->plan_epochs ? n_epochs - p0 : ctx->plan_epochs;
        if ((rc = enqueue_plan(ctx, np, d_recs + (size_t)p0 * ctx->cfg.max_chan, 1)))
            return rc;
    }
    return E1B200_OK;
}

int e1b200_synth_ranges_device(e1b200_ctx *ctx, int n_epochs, const e1_range_rec *d_rr, int16_t *d_out)
{
    if (!ctx || n_epochs < 0 || (n_epochs && (!d_rr || !d_out)))
        return E1B200_EINVAL;
    CK(cudaSetDevice(ctx->cfg.device));
    int rc = ensure_plan_scratch(ctx, n_epochs);
    if (!rc)
        rc = ensure_staging(ctx, 0, 0);
    if (rc)
        return rc;
    reset_call(ctx);
    const e1b200_config *cfg = &ctx->cfg;
    const size_t epoch_i16 = (size_t)cfg->samples_per_epoch * 2;
    for (int e0 = 0; e0 < n_epochs; e0 += ctx->plan_epochs) {
        int n = n_epochs - e0 < ctx->plan_epochs ? n_epochs - e0 : ctx->plan_epochs;
        size_t nrec = (size_t)n * cfg->max_chan;
        /* single staging slot: stream order keeps pass i+1's restate behind pass i's synthesis */
        e1_restate_kernel<<<(unsigned)((nrec + 127) / 128), 128, 0, ctx->stream>>>(d_rr + (size_t)e0 * cfg->max_chan, ctx->d_recs,
                                                                                    (int)nrec, cfg->dt_epoch);
        ctx->timing.kernel_launches += 1;
        if ((rc = enqueue_plan(ctx, n, ctx->d_recs)) || (rc = enqueue_synth(ctx, 0, n, ctx->d_recs, d_out + (size_t)e0 * epoch_i16)))
            return rc;
    }
    return E1B200_OK;
}

/* computeCodePhase (src/gal-sig.cpp:308-347) for callers that keep the restate on the CPU.
 * Plain IEEE double operations in the reference's order; this file is compiled with
 * -fmad=false / -ffp-contract=off so nothing is fused. */
int e1b200_restate(double rho_prev, double rho_cur, double dt, double grx_sec, double *f_carr, double *f_code,
                   double *code_phase0, int32_t *ibit0, int32_t *ipage_out)
{
    if (!f_carr || !f_code || !code_phase0 || !ibit0 || !(dt > 0.0))
        return E1B200_EINVAL;
    const double lambda_e1 = 0.1902936727983649, carr_to_code = 0.0006493506493506494, c_light = 2.99792458e8;
    volatile double rhorate = (rho_cur - rho_prev) / dt;
    volatile double fc = -rhorate / lambda_e1;
    volatile double prod = fc * carr_to_code;
    *f_carr = fc;
    *f_code = 1.023e6 + prod;
    volatile double tof = rho_cur / c_light;
    volatile double diff = grx_sec - tof;
    volatile double ms = diff * 1000.0;
    int ipage = (int)(ms / 2000.0);
    ms = ms - (double)(ipage * 2000);
    int ibit = (int)((unsigned int)ms / 4u);
    ms = ms - (double)(ibit * 4);
    volatile double q = ms / 4.0;
    *code_phase0 = q * (double)E1_CODE_LEN;
    *ibit0 = (ibit + E1_SYM_PER_PAGE / 2) % E1_SYM_PER_PAGE;
    if (ipage_out)
        *ipage_out = ipage % 360;
    return E1B200_OK;
}

/* How many times the reference's loop takes its code-wrap branch (src/galileo-sdr.cpp:491-494) inside one
 * block that starts from code_phase0: the wrap test runs at the TOP of samples 0 .. n_samp-1, so a sum that
 * reaches 4092 on the block's last addition is not counted (the next computeCodePhase absorbs it).  Exact:
 * the same roundings as the loop's `code_phase += f_code * delt`, walked binade by binade (e1_walk_up).
 * The caller adds it to ibit0 to know whether the loop would have called generateINavMsg (:497-506). */
int e1b200_code_wraps(double fs_hz, int32_t n_samp, double code_phase0, double f_code, int32_t *n_wraps)
{
    if (!n_wraps || !(fs_hz > 0.0) || n_samp < 0 || !(code_phase0 >= 0.0) || !(f_code > 0.0))
        return E1B200_EINVAL;
    const double sc = (1.0 / fs_hz) * f_code; /* f_code * delt, delt = 1/fs (:162, :528); one rounding each, commutative */
    if (!(sc < (double)E1C_CODE_LEN))
        return E1B200_EINVAL;
    double cp = code_phase0;
    int wraps = 0;
    if (n_samp > 0 && cp >= (double)E1C_CODE_LEN) { /* the test at sample 0 */
        cp -= (double)E1C_CODE_LEN;
        wraps++;
        if (cp >= (double)E1C_CODE_LEN)
            return E1B200_EINVAL;
    }
    int64_t k = 0;
    while (k < n_samp) {
        int w = 0;
        cp = e1_walk_up(cp, sc, (double)E1C_CODE_LEN, &k, n_samp, &w);
        if (w && k < n_samp)
            wraps++;
    }
    *n_wraps = wraps;
    return E1B200_OK;
}

int e1b200_get_timing(e1b200_ctx *ctx, e1b200_timing *out)
{
    if (!ctx || !out)
        return E1B200_EINVAL;
    *out = ctx->timing;
    return E1B200_OK;
}

int e1b200_get_stats(e1b200_ctx *ctx, e1b200_stats *out)
{
    if (!ctx || !out)
        return E1B200_EINVAL;
    out->exact_samples = ctx->counters[0];
    out->planner_errors = ctx->counters[1];
    out->serial_epochs = ctx->counters[2];
    out->hat_epochs = ctx->counters[3];
    out->plan_epochs = ctx->plan_epochs;
    out->tile = ctx->tile;
    out->tiles_per_epoch = ctx->tiles_per_epoch;
    out->batch_epochs = ctx->batch_epochs;
    out->sm_count = ctx->sm_count;
    out->ctas_per_sm = ctx->ctas_per_sm;
    out->smem_bytes = ctx->smem_bytes;
    out->synth_kernel = ctx->float_path ? E1B200_KERNEL_FLOAT
                        : ctx->evk      ? (E1B200_KERNEL_EV | (ctx->evk << 8))
                        : ctx->pair     ? (ctx->quad ? E1B200_KERNEL_CW4 : E1B200_KERNEL_CW2)
                                        : E1B200_KERNEL_RUN;
    return E1B200_OK;
}

__global__ void e1_selftest_any_hit_kernel(int n, const int64_t *cs, int32_t *out)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n)
        out[i] = e1_any_hit(cs[5 * i], cs[5 * i + 1], cs[5 * i + 2], cs[5 * i + 3], cs[5 * i + 4]);
}

int e1b200_selftest_any_hit(int device, int n_cases, const int64_t *cases, int32_t *out)
{
    if (n_cases <= 0 || !cases || !out)
        return E1B200_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess)
        return E1B200_ENODEV;
    int64_t *d_c = nullptr;
    int32_t *d_o = nullptr;
    int rc = E1B200_ECUDA;
    if (cudaMalloc(&d_c, sizeof(int64_t) * 5 * (size_t)n_cases) == cudaSuccess && cudaMalloc(&d_o, sizeof(int32_t) * (size_t)n_cases) == cudaSuccess &&
        cudaMemcpy(d_c, cases, sizeof(int64_t) * 5 * (size_t)n_cases, cudaMemcpyHostToDevice) == cudaSuccess) {
        e1_selftest_any_hit_kernel<<<(n_cases + 127) / 128, 128>>>(n_cases, d_c, d_o);
        if (cudaMemcpy(out, d_o, sizeof(int32_t) * (size_t)n_cases, cudaMemcpyDeviceToHost) == cudaSuccess)
            rc = E1B200_OK;
    }
    cudaFree(d_c);
    cudaFree(d_o);
    return rc;
}

void *e1b200_stream(e1b200_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }

/* ---- time-axis shards on several GPUs of a node: the gather without a collective ---------------------------------
 * The writer rank allocates the whole stream's buffer in its HBM and exports it; every other rank (one process per
 * GPU) opens the handle and passes `its pointer + its segment's offset` as d_out to e1b200_synth_*_device: the
 * synthesis kernel's own 128-bit stores then travel over NVLink straight into the writer's memory, tile by tile,
 * while the kernel computes -- no staging copy, no separate gather step.  (CUDA IPC; the devices must be peers.) */
int e1b200_peer_alloc(int device, size_t bytes, void **d_ptr, unsigned char handle[E1B200_IPC_HANDLE_BYTES])
{
    static_assert(sizeof(cudaIpcMemHandle_t) == E1B200_IPC_HANDLE_BYTES, "handle size");
    if (!d_ptr || !handle || !bytes)
        return E1B200_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess)
        return E1B200_ENODEV;
    if (cudaMalloc(d_ptr, bytes) != cudaSuccess)
        return E1B200_ENOMEM;
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, *d_ptr) != cudaSuccess) {
        cudaFree(*d_ptr);
        *d_ptr = nullptr;
        return E1B200_ECUDA;
    }
    memcpy(handle, &h, sizeof h);
    return E1B200_OK;
}

int e1b200_peer_open(int device, const unsigned char handle[E1B200_IPC_HANDLE_BYTES], void **d_ptr)
{
    if (!d_ptr || !handle)
        return E1B200_EINVAL;
    if (cudaSetDevice(device) != cudaSuccess)
        return E1B200_ENODEV;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, sizeof h);
    return cudaIpcOpenMemHandle(d_ptr, h, cudaIpcMemLazyEnablePeerAccess) == cudaSuccess ? E1B200_OK : E1B200_ECUDA;
}

int e1b200_peer_close(void *d_ptr) { return cudaIpcCloseMemHandle(d_ptr) == cudaSuccess ? E1B200_OK : E1B200_ECUDA; }
int e1b200_peer_free(void *d_ptr) { return cudaFree(d_ptr) == cudaSuccess ? E1B200_OK : E1B200_ECUDA; }

int e1b200_host_alloc(void **p, size_t bytes)
{
    if (!p)
        return E1B200_EINVAL;
    return cudaHostAlloc(p, bytes, cudaHostAllocDefault) == cudaSuccess ? E1B200_OK : E1B200_ENOMEM;
}

int e1b200_host_free(void *p) { return cudaFreeHost(p) == cudaSuccess ? E1B200_OK : E1B200_ECUDA; }

/* Pin a buffer the caller allocated itself (the reference's calloc'd iq_buff, src/galileo-sdr.cpp:326, which
 * it also frees itself at :655) so the D2H into it is a DMA at PCIe rate. */
int e1b200_host_register(void *p, size_t bytes)
{
    if (!p || !bytes)
        return E1B200_EINVAL;
    return cudaHostRegister(p, bytes, cudaHostRegisterDefault) == cudaSuccess ? E1B200_OK : E1B200_ECUDA;
}

int e1b200_host_unregister(void *p) { return cudaHostUnregister(p) == cudaSuccess ? E1B200_OK : E1B200_ECUDA; }

} /* extern "C" */
