/* e1_kernels.cuh -- sm_100a kernels of the Galileo E1B/C synthesiser.
 *
 *   e1_restate_kernel        computeCodePhase (src/gal-sig.cpp:308-347) per (epoch, channel)
 *   e1_plan_code_kernel      exact code-phase / symbol checkpoints per (epoch, tile, channel)
 *   e1_v2_prep/ideal/drift/estimate/span/chain_kernel
 *                            the parallel exact carrier planner (see "parallel carrier planner" in e1_core.h)
 *   e1_plan_carr_kernel      the same checkpoints by one serial walk per channel (debug / comparison)
 *   e1_finalize_kernel       checkpoints + translations -> one parameter block per tile
 *   e1_clean_kernel          tile-level ambiguity test (e1_par_clean): marks the (tile, channel) sets whose runs
 *                            cannot be ambiguous, so the sample loop skips its per-sample tracking for them
 *   e1_synth_cw_kernel<NH,T> the sample loop (src/galileo-sdr.cpp:481-539): per sample, all channels,
 *                            int32 accumulate, packed int16 I/Q, 128-bit stores; 32 or 64 samples per thread
 *   e1_synth_kernel<R>       the same with R = 4, 8 or 16 samples per thread (tiles shorter than 8192 samples)
 *   e1_synth_ev_kernel<T>    the same, event-driven, for sample rates of 10 MS/s and more: a channel's term is piecewise constant
 *                            there, only its change points are computed (e1_ev_add), one running sum per tile gives the samples
 *
 * HBM layout
 *   recs   e1_epoch_rec[n_epochs][max_chan]                         176 B each (caller / H2D)
 *   ck     e1_tile_ck[n_epochs][tiles_per_epoch][max_chan]           32 B each (scratch)
 *   blk    per tile: 16 B header + max_chan e1_chan_par (112 B), active channels first (scratch)
 *   out    int16 I,Q interleaved, sample (epoch*N + k) at byte 4*(epoch*N + k)
 *   codes  uint32[50][516]: one 2-bit field per BOC(1,1) half-chip (see e1_core.h)   103 200 B
 *   lut    int32[642][32]: carrier term by table position, one copy per lane         82 176 B
 *          both smem-resident, loaded once per persistent CTA with cp.async.bulk
 */
#ifndef E1_KERNELS_CUH
#define E1_KERNELS_CUH

#include <cuda_runtime.h>
#include <limits.h>
#include <stdint.h>

#include "../../include/e1b200.h"
#include "e1_core.h"

#define E1_CODE_WORDS_PER_PRN E1C_CODE_WORDS_PER_PRN
#define E1_CODES_BYTES (E1C_N_PRN * E1_CODE_WORDS_PER_PRN * 4)
#define E1_LUT_ENTRIES E1C_LUT_ENTRIES
#define E1_LUT_BYTES (E1_LUT_ENTRIES * 4)
#define E1_SYNTH_THREADS E1C_THREADS

struct e1_synth_args {
    const unsigned char *blk; /* parameter blocks of this launch's tiles, e1_blk_bytes(max_chan) apart */
    const uint32_t *codes;
    const int32_t *lut;
    int16_t *out;
    unsigned long long *counters; /* [0] ambiguous samples resolved exactly */
    unsigned int *next_tile;      /* zeroed before the launch: tiles beyond the first wave are handed out dynamically */
    int n_epochs, n_samp, max_chan, tile, tiles_per_epoch;
    uint32_t thr_carr, thr_code; /* closed form vs serial recurrence (e1_thr_carr / e1_thr_code)  */
    uint32_t tc_carr, tc_code;   /* + the fast path's truncation slack (e1_tc_carr / e1_tc_code) */
    int vec_ok, use_bulk;
};

struct e1_finalize_args {
    const e1_epoch_rec *recs;
    const e1_tile_ck *ck;
    const e1_trans *delta; /* planner translation of each span's carrier checkpoints, channel-major:
                              delta[ch * delta_stride + e * geo.spans_per_epoch + span]           */
    int delta_stride;
    e1_span_geo geo;
    unsigned char *blk;
    unsigned long long *counters; /* [1] planner errors */
    double delt;
    int n_epochs, max_chan, tile, tiles_per_epoch;
    uint32_t tc_code;
    uint32_t cfg_flags; /* E1B200_CFG_CBOC / _GAIN: the parameter blocks carry the channel gain (float path) */
};

/* ------------------------------------------------------------------ restate (a8) */
__global__ void e1_restate_kernel(const e1_range_rec *rr, e1_epoch_rec *recs, int n, double dt)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n)
        return;
    const e1_range_rec r = rr[i];
    e1_epoch_rec o;
    o.prn = r.prn;
    o.flags = r.flags & 0xffu;
    o.gain_q7 = (int32_t)(r.flags >> 8); /* e1_range_rec carries gain[i] in the upper flag bits */
    o.carr_phase_init = r.carr_phase_init;
    const double lambda_e1 = 0.1902936727983649;       /* constants.h:119 */
    const double carr_to_code = 0.0006493506493506494; /* constants.h:125 */
    const double c_light = 2.99792458e8;               /* constants.h:60  */
    double rhorate = __ddiv_rn(__dadd_rn(r.rho_cur, -r.rho_prev), dt);       /* :315 */
    double fc = __ddiv_rn(-rhorate, lambda_e1);                               /* :318 */
    o.f_carr = fc;
    o.f_code = __dadd_rn(1.023e6, __dmul_rn(fc, carr_to_code));               /* :320 */
    double ms = __dmul_rn(__dadd_rn(r.grx_sec, -__ddiv_rn(r.rho_cur, c_light)), 1000.0); /* :322 */
    int ipage = __double2int_rz(__ddiv_rn(ms, 2000.0));                       /* :324 */
    ms = __dadd_rn(ms, -(double)(ipage * 2000));                              /* :326 */
    int ibit = (int)(__double2uint_rz(ms) / 4u);                              /* :328 */
    ms = __dadd_rn(ms, -(double)(ibit * 4));                                  /* :329 */
    o.code_phase0 = __dmul_rn(__ddiv_rn(ms, 4.0), (double)E1C_CODE_LEN);      /* :330 */
    o.ibit0 = (ibit + E1C_SYM_PER_PAGE / 2) % E1C_SYM_PER_PAGE;               /* :334 */
#pragma unroll
    for (int k = 0; k < E1_PAGE_BYTES; k++) {
        o.page_cur[k] = r.page_cur[k];
        o.page_next[k] = r.page_next[k];
    }
    recs[i] = o;
}

/* ------------------------------------------------------------------ planners */
/* one thread per (epoch, channel) */
__global__ void e1_plan_code_kernel(const e1_epoch_rec *recs, e1_tile_ck *ck, int n_epochs, int max_chan,
                                    int n_samp, int tile, int tiles_per_epoch, double delt)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_epochs * max_chan)
        return;
    int e = i / max_chan, ch = i - e * max_chan;
    e1_plan_code_epoch(&recs[i], ck + (size_t)e * tiles_per_epoch * max_chan + ch, max_chan, n_samp, tile,
                       tiles_per_epoch, delt);
}

/* carrier-only passes (e1b200_plan_phases) do not run the code planner, which is where a record is
   validated: the carrier fields alone, one thread per (epoch, channel) */
__global__ void e1_validate_carrier_kernel(const e1_epoch_rec *recs, int n, double delt, unsigned long long *counters)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n && e1_rec_active(&recs[i]) && !e1_rec_carrier_ok(&recs[i], delt))
        atomicAdd(&counters[1], 1ull);
}

/* serial reference planner (E1B200_CFG_SERIAL_PLANNER): one thread per channel, one exact walk */
__global__ void e1_plan_carr_kernel(const e1_epoch_rec *recs, e1_tile_ck *ck, double *phase, int n_epochs,
                                    int max_chan, int n_samp, int tile, int tiles_per_epoch, double delt)
{
    int ch = blockIdx.x * blockDim.x + threadIdx.x;
    if (ch >= max_chan)
        return;
    double phi = phase[ch];
    for (int e = 0; e < n_epochs; e++)
        phi = e1_plan_carr_epoch(&recs[(size_t)e * max_chan + ch], ck + (size_t)e * tiles_per_epoch * max_chan + ch,
                                 max_chan, phi, n_samp, tile, tiles_per_epoch, delt);
    phase[ch] = phi;
}

/* parallel carrier planner (see e1_core.h): K0 ideal prefix, drift pass, K1 estimate prefix, span
 * pass, chain.  Per-(epoch, channel) kernels map consecutive threads to consecutive epochs of one
 * channel, so a warp walks similar Dopplers and stays converged.                                  */
struct e1_plan_args {
    const e1_epoch_rec *recs;
    e1_tile_ck *ck;
    double *phase;   /* [max_chan] carried carrier phase (in: batch start, out: batch end) */
    e1_prep *prep;   /* channel-major [max_chan][n_units], like g, dend, est, delta and units;
                        unit u = epoch * geo.spans_per_epoch + span                              */
    double *g, *dend, *est;
    e1_trans *delta;
    e1_unit *units;
    unsigned long long *counters; /* [2] serial epochs, [3] HAT epochs */
    double delt;
    int n_epochs, n_samp, max_chan, tile, tiles_per_epoch;
    e1_span_geo geo;
    int n_units; /* n_epochs * geo.spans_per_epoch */
};

__global__ void e1_v2_prep_kernel(const e1_plan_args P)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x; /* record order: coalesced reads */
    if (i >= P.n_epochs * P.max_chan)
        return;
    int e = i / P.max_chan, ch = i - e * P.max_chan;
    for (int sp = 0; sp < P.geo.spans_per_epoch; sp++)
        e1_v2_prep(&P.recs[i], P.delt, sp, e1_span_samples(&P.geo, sp, P.n_samp, P.tile),
                   &P.prep[(size_t)ch * P.n_units + (size_t)e * P.geo.spans_per_epoch + sp]);
}

/* first checkpoint of (unit u, channel ch): [epoch][tile][channel] */
__device__ __forceinline__ e1_tile_ck *e1_unit_ck(const e1_plan_args &P, int u, int ch)
{
    const int e = u / P.geo.spans_per_epoch, sp = u - e * P.geo.spans_per_epoch;
    return P.ck + ((size_t)e * P.tiles_per_epoch + (size_t)sp * P.geo.span_tiles) * P.max_chan + ch;
}

/* The three per-channel passes (ideal prefix, estimate prefix, chain) run one channel per block of
 * E1_SERIAL_THREADS threads over that channel's n_units spans (channel-major, contiguous).
 *
 * The two prefix passes only produce ESTIMATES (nothing exact depends on their last bit: the chain
 * validates every guess), so they are evaluated as chunked scans: every thread folds its contiguous
 * run of units from a zero start, thread 0 combines the 128 run summaries, every thread replays its
 * run from its true start. */
#define E1_SERIAL_THREADS 128
#define E1_SERIAL_CHUNK 128
#define E1_CHAIN_THREADS 512 /* chain kernel: one span per thread per round, 512 spans per round */

__device__ __forceinline__ double e1_fold1(double v) /* into (-1,1), sign kept (like :532) */
{
    if (v >= 1.0 || v <= -1.0)
        v -= (double)(long long)v;
    return v;
}
__device__ __forceinline__ double e1_fold_half(double v) /* into (-1/2, 1/2] */
{
    v -= (double)(long long)v;
    if (v > 0.5)
        v -= 1.0;
    else if (v <= -0.5)
        v += 1.0;
    return v;
}

__global__ void __launch_bounds__(E1_SERIAL_THREADS) e1_v2_ideal_kernel(const e1_plan_args P)
{
    __shared__ double s_val[E1_SERIAL_THREADS], s_start[E1_SERIAL_THREADS];
    __shared__ int s_abs[E1_SERIAL_THREADS];
    const int ch = blockIdx.x, tid = threadIdx.x;
    if (ch >= P.max_chan)
        return;
    const size_t o = (size_t)ch * P.n_units;
    const int L = (P.n_units + E1_SERIAL_THREADS - 1) / E1_SERIAL_THREADS;
    const int u0 = min(tid * L, P.n_units), u1 = min(u0 + L, P.n_units);
    double g = 0.0;
    int abs = 0;
    for (int u = u0; u < u1; u++) { /* this run from a zero start: displacement, or absolute value after a reset */
        const e1_prep p = P.prep[o + u];
        if (p.flags & E1_PREP_SET_PHASE) {
            g = p.init;
            abs = 1;
        }
        if (p.flags & E1_PREP_ACTIVE)
            g = e1_ideal_next(g, p.sp, p.n);
    }
    s_val[tid] = g;
    s_abs[tid] = abs;
    __syncthreads();
    if (tid == 0) {
        double cur = P.phase[ch];
        for (int t = 0; t < E1_SERIAL_THREADS; t++) {
            s_start[t] = cur;
            cur = s_abs[t] ? s_val[t] : e1_fold1(cur + s_val[t]);
        }
    }
    __syncthreads();
    g = s_start[tid];
    for (int u = u0; u < u1; u++) { /* e1_v2_ideal_prefix from the run's true start */
        const e1_prep p = P.prep[o + u];
        if (p.flags & E1_PREP_SET_PHASE)
            g = p.init;
        P.g[o + u] = g;
        if (p.flags & E1_PREP_ACTIVE)
            g = e1_ideal_next(g, p.sp, p.n);
    }
}

__global__ void e1_v2_drift_kernel(const e1_plan_args P)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x; /* channel-major: a warp walks one channel's spans */
    if (i >= P.n_units * P.max_chan)
        return;
    P.dend[i] = e1_v2_drift_unit(&P.prep[i], P.g[i]);
}

/* e1_v2_estimate_prefix as a scan: est[u] = g[u] + eta[u] with eta[u+1] = eta[u] + (dend[u] - g[u+1])
 * (the span's measured rounding drift), eta = 0 at the batch start and at every phase reset. */
__global__ void __launch_bounds__(E1_SERIAL_THREADS) e1_v2_estimate_kernel(const e1_plan_args P)
{
    __shared__ double s_val[E1_SERIAL_THREADS], s_start[E1_SERIAL_THREADS];
    __shared__ int s_abs[E1_SERIAL_THREADS];
    const int ch = blockIdx.x, tid = threadIdx.x;
    if (ch >= P.max_chan)
        return;
    const size_t o = (size_t)ch * P.n_units;
    const int L = (P.n_units + E1_SERIAL_THREADS - 1) / E1_SERIAL_THREADS;
    const int u0 = min(tid * L, P.n_units), u1 = min(u0 + L, P.n_units);
    /* eta at the START of unit u1 given eta = 0 at the start of unit u0 */
    double eta = 0.0;
    int abs = 0;
    for (int u = u0; u < u1; u++) {
        if (P.prep[o + u].flags & E1_PREP_SET_PHASE) {
            eta = 0.0;
            abs = 1;
        }
        if (u + 1 < P.n_units)
            eta += e1_fold_half(P.dend[o + u] - P.g[o + u + 1]);
    }
    s_val[tid] = eta;
    s_abs[tid] = abs;
    __syncthreads();
    if (tid == 0) {
        double cur = 0.0;
        for (int t = 0; t < E1_SERIAL_THREADS; t++) {
            s_start[t] = cur;
            cur = s_abs[t] ? s_val[t] : cur + s_val[t];
        }
    }
    __syncthreads();
    eta = s_start[tid];
    for (int u = u0; u < u1; u++) {
        if (P.prep[o + u].flags & E1_PREP_SET_PHASE)
            eta = 0.0;
        P.est[o + u] = e1_fold1(P.g[o + u] + eta);
        if (u + 1 < P.n_units)
            eta += e1_fold_half(P.dend[o + u] - P.g[o + u + 1]);
    }
}

__global__ void e1_v2_span_kernel(const e1_plan_args P)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_units * P.max_chan)
        return;
    int ch = i / P.n_units, u = i - ch * P.n_units;
    e1_v2_span_unit(&P.prep[i], u, P.phase[ch], &P.est[i], P.tile, e1_unit_ck(P, u, ch), P.max_chan, &P.units[i]);
}

/* Chain.  Exact, and serial in principle: the translation of span u is
 *     D[u] = (true post-wrap value at its anchor) - (guessed one)
 *          = (last_p[u-1] + D[u-1]) - anchor_p[u]
 * where last_p / anchor_p are the span pass's hat values -- all multiples of 2^-52 below 1, so the sums
 * are exact in any order.  A round of E1_CHAIN_THREADS consecutive spans is therefore first tried as a parallel
 * inclusive scan of x[u] = last_p[u-1] - anchor_p[u] (x[first] uses the carried state); if every span
 * of the chunk is an ordinary accepted guess (HAT unit, anchored on its predecessor's last wrap, no
 * tie wrap, lo <= D < hi) the chunk is done; otherwise thread 0 walks that chunk with
 * e1_v2_chain_step exactly as the serial chain would. */
__global__ void __launch_bounds__(E1_CHAIN_THREADS) e1_v2_chain_kernel(const e1_plan_args P)
{
    __shared__ __align__(16) e1_unit s_units[E1_CHAIN_THREADS];
    __shared__ e1_trans s_delta[E1_CHAIN_THREADS];
    __shared__ double s_warp[E1_CHAIN_THREADS / 32];
    __shared__ int s_wmax[E1_CHAIN_THREADS / 32];
    __shared__ e1_chain_state s_cs;
    __shared__ int s_bad;
    const int ch = blockIdx.x, tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (ch >= P.max_chan)
        return;
    const size_t o = (size_t)ch * P.n_units;
    unsigned long long st[2] = {0, 0};
    if (tid == 0)
        e1_chain_init(&s_cs, P.phase[ch]);
    /* The rounds are serial (each starts from the state the previous one leaves) and short, so the
       global-memory latency of a round's inputs would be most of its time: every thread fetches ITS span
       of the next round into registers before the current round's scan. */
    uint4 pf[sizeof(e1_unit) / 16];
    auto fetch = [&](int e0) {
        if (e0 + tid < P.n_units) {
            const uint4 *src = reinterpret_cast<const uint4 *>(P.units + o + e0 + tid);
#pragma unroll
            for (int i = 0; i < (int)(sizeof(e1_unit) / 16); i++)
                pf[i] = src[i];
        }
    };
    fetch(0);
    for (int e0 = 0; e0 < P.n_units; e0 += E1_CHAIN_THREADS) {
        const int n = min(E1_CHAIN_THREADS, P.n_units - e0);
        if (tid < n) {
            uint4 *dst = reinterpret_cast<uint4 *>(&s_units[tid]);
#pragma unroll
            for (int i = 0; i < (int)(sizeof(e1_unit) / 16); i++)
                dst[i] = pf[i];
        }
        fetch(e0 + E1_CHAIN_THREADS);
        if (tid == 0)
            s_bad = 0;
        __syncthreads();
        /* Optimistic pass over spans [start, n) of the round, thread t owns span e0 + t: assume every span
           is an accepted HAT span, scan the translations, validate.  The prefix up to the first span that
           is not (guess rejected, tie wrap, phase reset, idle slot, sign change, no wrap within reach to
           anchor on) is committed, thread 0 takes that ONE span through the serial chain step, and the
           pass resumes behind it.  (A round that keeps failing goes to the serial chain for its remainder.) */
        int start = 0, restarts = 0;
        while (start < n) {
            /* x = what span tid adds to T, the true |phase| right after the most recent wrap: a span with a
               wrap moves it to last_p + D = last_p + (T_before - anchor_p), a span without one leaves it.
               So T before span tid = T at `start` + the exclusive scan of x, and D = that - anchor_p -- whether
               the anchor wrap lies in the span just before or (low Doppler) several spans back. */
            double x = 0.0;
            int wrapidx = -1; /* this span if it holds a wrap: max-scanned into "the span with the most recent wrap" */
            if (tid >= start && tid < n) {
                const e1_unit *u = &s_units[tid];
                if (u->type == E1_UNIT_HAT && u->last_k >= 1) {
                    x = __dadd_rn(u->last_p, -u->anchor_p);
                    wrapidx = tid;
                }
            }
            double S = x; /* inclusive scans over the block (exact additions: every term is a multiple of 2^-52) */
            int W = wrapidx;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                const double y = __shfl_up_sync(0xffffffffu, S, d);
                const int w = __shfl_up_sync(0xffffffffu, W, d);
                if (lane >= d) {
                    S = __dadd_rn(S, y);
                    W = max(W, w);
                }
            }
            if (lane == 31) {
                s_warp[wid] = S;
                s_wmax[wid] = W;
            }
            if (tid == 0)
                s_bad = n; /* first span of [start, n) that fails */
            __syncthreads();
            for (int w = 0; w < wid; w++) {
                S = __dadd_rn(S, s_warp[w]);
                W = max(W, s_wmax[w]);
            }
            /* exclusive values: the state BEFORE span tid */
            const double Tb = __dadd_rn(s_cs.prev_p, __dadd_rn(S, -x));
            int Wb = __shfl_up_sync(0xffffffffu, W, 1);
            if (lane == 0) {
                Wb = -1;
                for (int w = 0; w < wid; w++)
                    Wb = max(Wb, s_wmax[w]);
            }
            double D = 0.0;
            if (tid >= start && tid < n) {
                const e1_unit *u = &s_units[tid];
                int p_ok, p_neg, p_k, p_u;
                if (Wb >= start) { /* the most recent wrap is inside this pass: every span of the pass is taken to be an accepted guess */
                    const e1_unit *q = &s_units[Wb];
                    p_ok = 1, p_neg = q->neg, p_k = q->last_k, p_u = e0 + Wb;
                } else {
                    p_ok = s_cs.prev_ok, p_neg = s_cs.prev_neg, p_k = s_cs.prev_k, p_u = s_cs.prev_u;
                }
                D = __dadd_rn(Tb, -u->anchor_p);
                const int ok = u->type == E1_UNIT_HAT && u->tie == 0 && p_ok && p_neg == u->neg && p_u == e0 + tid - u->anchor_back &&
                               p_k == u->anchor_k;
                if (!(ok && D >= u->lo && D < u->hi))
                    atomicMin(&s_bad, tid);
            }
            __syncthreads();
            const int bad = s_bad;
            if (tid >= start && tid < bad) {
                const e1_unit *u = &s_units[tid];
                e1_trans tr;
                tr.a = tr.b = u->neg ? -D : D;
                tr.k_split = 0;
                tr.pad = 0;
                s_delta[tid] = tr;
            }
            __syncthreads(); /* every thread has read s_cs */
            /* the state the serial chain (e1_v2_chain_step) would carry out of the accepted prefix: the phase
               after its last span, and the most recent wrap -- the last one inside the prefix, or, when the
               prefix has none (all its spans aligned and wrap-free), still the one carried in */
            if (tid == bad - 1 && bad > start) {
                const e1_unit *u = &s_units[tid];
                s_cs.phi = __dadd_rn(u->end_phi, u->neg ? -D : D);
                if (W >= start) {
                    const e1_unit *q = &s_units[W];
                    /* D of span W = T before W - anchor_p[W]; T after it = last_p[W] + that = T before tid + x terms:
                       the inclusive sum S holds it */
                    s_cs.prev_p = __dadd_rn(s_cs.prev_p, S);
                    s_cs.prev_k = q->last_k;
                    s_cs.prev_u = e0 + W;
                    s_cs.prev_ok = 1;
                    s_cs.prev_neg = q->neg;
                }
            }
            if (tid == 0)
                st[1] += (unsigned long long)(bad - start);
            __syncthreads();
            if (bad >= n)
                break;
            const int upto = ++restarts > 16 ? n : bad + 1; /* the same in every thread */
            if (tid == 0) {
                e1_chain_state cs = s_cs;
                for (int i = bad; i < upto; i++) { /* rare: the span's step and length come straight from global memory */
                    const double sp_i = P.prep[o + e0 + i].sp;
                    const int n_i = P.prep[o + e0 + i].n;
                    s_delta[i] = e1_v2_chain_step(&cs, &s_units[i], e0 + i, sp_i, n_i, P.tile, (n_i + P.tile - 1) / P.tile,
                                                  e1_unit_ck(P, e0 + i, ch), P.max_chan, st);
                }
                s_cs = cs;
            }
            start = upto;
            __syncthreads();
        }
        __syncthreads();
        if (tid < n)
            P.delta[o + e0 + tid] = s_delta[tid];
    }
    if (tid == 0) {
        P.phase[ch] = s_cs.phi;
        if (st[0])
            atomicAdd(&P.counters[2], st[0]);
        if (st[1])
            atomicAdd(&P.counters[3], st[1]);
    }
}

/* ------------------------------------------------------------------ finalize
 * One warp per tile: tile checkpoints (+ the chain's translation) -> the tile's parameter block,
 * active channels compacted to the front, so the synthesis CTAs only have to bulk-copy it. */
__global__ void __launch_bounds__(128) e1_finalize_kernel(const e1_finalize_args A)
{
    const int lane = threadIdx.x & 31;
    const long tile_id = (long)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    if (tile_id >= (long)A.n_epochs * A.tiles_per_epoch)
        return;
    const int e = (int)(tile_id / A.tiles_per_epoch);
    unsigned char *blk = A.blk + (size_t)tile_id * e1_blk_bytes(A.max_chan);
    e1_chan_par *par = reinterpret_cast<e1_chan_par *>(blk + E1C_BLK_HEADER);
    int base = 0;
    unsigned long long errs = 0;
    for (int c0 = 0; c0 < A.max_chan; c0 += 32) {
        const int ch = c0 + lane;
        e1_tile_ck c;
        c.sym = 0;
        if (ch < A.max_chan)
            c = A.ck[(size_t)tile_id * A.max_chan + ch];
        const bool active = (c.sym & E1_CK_ACTIVE) != 0;
        const unsigned m = __ballot_sync(0xffffffffu, active);
        if (active) {
            e1_chan_par p;
            const int t = (int)(tile_id - (long)e * A.tiles_per_epoch), sp = t / A.geo.span_tiles;
            e1_make_par(&c, &A.recs[(size_t)e * A.max_chan + ch], A.delt, A.tile,
                        e1_trans_at(&A.delta[(size_t)ch * A.delta_stride + (size_t)e * A.geo.spans_per_epoch + sp],
                                    (t - sp * A.geo.span_tiles) * A.tile),
                        A.tc_code, &p, A.cfg_flags);
            if (c.sym & E1_CK_ERROR)
                errs++;
            par[base + __popc(m & ((1u << lane) - 1u))] = p;
        }
        base += __popc(m);
    }
    if (lane == 0)
        *reinterpret_cast<uint4 *>(blk) = make_uint4((unsigned)base, 0u, 0u, 0u);
    if (errs)
        atomicAdd(&A.counters[1], errs);
}

/* ------------------------------------------------------------------ synthesis */
__device__ __forceinline__ uint32_t e1_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

/* 1-D bulk copy global -> shared through the TMA unit, completion on an mbarrier. */
__device__ __forceinline__ void e1_bulk_g2s(void *dst, const void *src, uint32_t bytes, uint64_t *bar)
{
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 :
                 : "r"(e1_smem_u32(dst)), "l"(src), "r"(bytes), "r"(e1_smem_u32(bar))
                 : "memory");
}
__device__ __forceinline__ void e1_mbar_init(uint64_t *bar, uint32_t count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" : : "r"(e1_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void e1_mbar_init_fence()
{
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void e1_mbar_expect(uint64_t *bar, uint32_t bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" : : "r"(e1_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void e1_mbar_wait(uint64_t *bar, uint32_t parity)
{
    asm volatile("{\n\t.reg .pred p;\n\tE1_WAIT_%=:\n\t"
                 "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
                 "@!p bra E1_WAIT_%=;\n\t}"
                 :
                 : "r"(e1_smem_u32(bar)), "r"(parity)
                 : "memory");
}

/* Marks the (tile, channel) parameter sets whose runs cannot be ambiguous (e1_par_clean): one thread per
 * (tile, compacted slot), after e1_finalize_kernel, for the carry-walked and the event-driven kernel. */
struct e1_clean_args {
    unsigned char *blk;
    long n_tiles;
    int max_chan, tile;
    uint32_t tc_carr, lim_carr, lim_code, thr_code;
    int cw_samples; /* samples a thread of the synthesis kernel walks from one carrier start (32 or 64) */
    int code_run;   /* ... and steps the code fraction from one start (16; E1C_EV_RUN in an event-driven context) */
};
__global__ void __launch_bounds__(128) e1_clean_kernel(const e1_clean_args A)
{
    const long i = (long)blockIdx.x * blockDim.x + threadIdx.x;
    const long tile_id = i / A.max_chan;
    const int slot = (int)(i - tile_id * A.max_chan);
    if (tile_id >= A.n_tiles)
        return;
    unsigned char *blk = A.blk + (size_t)tile_id * e1_blk_bytes(A.max_chan);
    if (slot >= *reinterpret_cast<const int *>(blk))
        return;
    e1_chan_par *p = reinterpret_cast<e1_chan_par *>(blk + E1C_BLK_HEADER) + slot;
    e1_chan_par q;
    q.U0 = p->U0, q.dU = p->dU, q.HA = p->HA, q.HB = p->HB, q.dH = p->dH, q.j_w = p->j_w, q.misc = p->misc, q.Dlo = p->Dlo;
    if (e1_par_clean(&q, A.tile, A.tc_carr, A.lim_carr, A.lim_code, A.thr_code, A.cw_samples, A.code_run))
        p->misc = q.misc | E1_PAR_CLEAN;
}

/* The rare path of one (thread, channel): the fast form flagged the run (rc 1: its terms are in the
 * accumulators and come back out) or did not handle it (rc 2).  d[i] receives the correction. */
template <int R>
__device__ __noinline__ void e1_fix_run(const e1_chan_par *p, const uint32_t *codes, const unsigned char *lut_lane, int j0,
                                        int *d, uint32_t thr_carr, uint32_t thr_code, uint32_t tc_carr, uint32_t tc_code,
                                        unsigned long long *n_exact)
{
    int t[R], g[R];
#pragma unroll
    for (int i = 0; i < R; i++)
        t[i] = 0;
    e1_run_fast<R>(p, codes, lut_lane, j0, t, tc_carr, e1_lim_carr(tc_carr, thr_carr), e1_lim_code(tc_code, thr_code));
    e1_channel_run(p, codes, lut_lane, j0, R, g, thr_carr, thr_code, e1_bias_h(tc_code), n_exact);
#pragma unroll
    for (int i = 0; i < R; i++)
        d[i] = g[i] - t[i];
}

/* The out-of-line rest of a carry-walked run (rc from e1_cw_add; see e1_cw_rest_impl).  Takes the shared-memory
 * operands the kernel holds and turns them back into generic pointers for the generic-form functions. */
template <int NH>
__device__ __noinline__ void e1_cw_rest(const e1_chan_par *p, uint32_t codes_s, uint32_t lut_s, int j0, int *d, uint32_t rc, uint32_t thr_carr,
                                        uint32_t thr_code, uint32_t tc_carr, uint32_t tc_code, unsigned long long *n_exact)
{
    const uint32_t *codes = (const uint32_t *)__cvta_shared_to_generic((size_t)codes_s);
    const unsigned char *lut_lane = (const unsigned char *)__cvta_shared_to_generic((size_t)lut_s);
    e1_cw_rest_impl<NH>(p, codes, lut_lane, j0, d, rc, thr_carr, thr_code, tc_carr, tc_code, n_exact, (unsigned long long *)0);
}

/* Persistent CTA, one per SM.  Shared memory: code words of all PRNs (103 200 B) and the replicated
 * carrier table (82 176 B), both loaded once by bulk copy; two parameter-block buffers, the next
 * tile's block in flight (bulk copy + mbarrier) while the current tile is computed.  The first wave of
 * tiles is blockIdx.x; after that CTAs draw tile numbers from an atomic counter (one draw ahead, so
 * the atomic's latency hides behind a tile), which keeps the SMs level when some tiles are slow.
 * Thread t owns the R consecutive samples [t*R, (t+1)*R) of the tile and walks the active channels
 * with the I/Q sums in registers. */
template <int R>
__global__ void __launch_bounds__(E1_SYNTH_THREADS, 1) e1_synth_kernel(const e1_synth_args A)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t *s_codes = reinterpret_cast<uint32_t *>(smem_raw);
    unsigned char *s_lut = smem_raw + E1_CODES_BYTES;
    const uint32_t blk_bytes = (uint32_t)e1_blk_bytes(A.max_chan);
    unsigned char *s_blk0 = smem_raw + E1_CODES_BYTES + E1_LUT_BYTES;
    __shared__ __align__(8) uint64_t s_bar[3]; /* [0] tables, [1],[2] parameter buffers */
    __shared__ unsigned long long s_cnt;
    __shared__ long s_tile[2]; /* tile number whose block is (being) loaded into buffer 0 / 1 */

    const int tid = threadIdx.x;
    const long total_tiles = (long)A.n_epochs * A.tiles_per_epoch;
    long next_id = 0; /* thread 0: the tile after the one in s_tile[...] */
    if (tid == 0) {
        s_cnt = 0;
        e1_mbar_init(&s_bar[0], 1);
        e1_mbar_init(&s_bar[1], 1);
        e1_mbar_init(&s_bar[2], 1);
        e1_mbar_init_fence();
        if (A.use_bulk) {
            e1_mbar_expect(&s_bar[0], E1_CODES_BYTES + E1_LUT_BYTES);
            e1_bulk_g2s(s_codes, A.codes, E1_CODES_BYTES, &s_bar[0]);
            e1_bulk_g2s(s_lut, A.lut, E1_LUT_BYTES, &s_bar[0]);
        }
        s_tile[0] = blockIdx.x;
        if ((long)blockIdx.x < total_tiles) {
            e1_mbar_expect(&s_bar[1], blk_bytes);
            e1_bulk_g2s(s_blk0, A.blk + (size_t)blockIdx.x * blk_bytes, blk_bytes, &s_bar[1]);
        }
        next_id = (long)gridDim.x + atomicAdd(A.next_tile, 1u);
    }
    __syncthreads();
    if (A.use_bulk) {
        e1_mbar_wait(&s_bar[0], 0);
    } else {
        for (int i = tid; i < E1_CODES_BYTES / 4; i += E1_SYNTH_THREADS)
            s_codes[i] = A.codes[i];
        for (int i = tid; i < E1_LUT_ENTRIES; i += E1_SYNTH_THREADS)
            reinterpret_cast<int32_t *>(s_lut)[i] = A.lut[i];
        __syncthreads();
    }

    const unsigned char *lut_lane = s_lut + 4 * (tid & (E1C_LUT_REP - 1)); /* this lane's copy of every entry */
    const uint32_t lim_carr = e1_lim_carr(A.tc_carr, A.thr_carr), lim_code = e1_lim_code(A.tc_code, A.thr_code);
    const int j0 = tid * R;
    unsigned long long n_exact = 0;
    for (int it = 0;; it++) {
        const int b = it & 1;
        const long tile_id = s_tile[b];
        if (tile_id >= total_tiles)
            break;
        /* every thread is past the previous tile (barrier at the end of the loop body), so the other
           buffer is free: start the next tile's block now, then draw the tile after it */
        if (tid == 0) {
            s_tile[1 - b] = next_id;
            if (next_id < total_tiles) {
                e1_mbar_expect(&s_bar[2 - b], blk_bytes);
                e1_bulk_g2s(s_blk0 + (size_t)(1 - b) * blk_bytes, A.blk + (size_t)next_id * blk_bytes, blk_bytes, &s_bar[2 - b]);
                next_id = (long)gridDim.x + atomicAdd(A.next_tile, 1u);
            }
        }
        e1_mbar_wait(&s_bar[1 + b], (uint32_t)(it >> 1) & 1u);
        const unsigned char *blk = s_blk0 + (size_t)b * blk_bytes;
        const int nact = *reinterpret_cast<const int *>(blk);
        const e1_chan_par *par = reinterpret_cast<const e1_chan_par *>(blk + E1C_BLK_HEADER);
        const int e = (int)(tile_id / A.tiles_per_epoch), t = (int)(tile_id - (long)e * A.tiles_per_epoch);
        const int n_valid = min(A.tile, A.n_samp - t * A.tile);

        if (j0 < n_valid) {
            int acc[R];
#pragma unroll
            for (int i = 0; i < R; i++)
                acc[i] = 0;
            for (int a = 0; a < nact; a++)
                if (e1_run_fast<R>(&par[a], s_codes, lut_lane, j0, acc, A.tc_carr, lim_carr, lim_code)) {
                    /* rare: this (thread, channel) goes through the generic form */
                    int d[R];
                    e1_fix_run<R>(&par[a], s_codes, lut_lane, j0, d, A.thr_carr, A.thr_code, A.tc_carr, A.tc_code, &n_exact);
#pragma unroll
                    for (int i = 0; i < R; i++)
                        acc[i] += d[i];
                }
            /* a6 + sink format (:536-537): (short)I, (short)Q interleaved; acc = I + 65536*Q */
            int16_t *dst = A.out + ((size_t)e * A.n_samp + (size_t)t * A.tile + j0) * 2;
            if (A.vec_ok && j0 + R <= n_valid) {
#pragma unroll
                for (int i = 0; i < R; i += 4)
                    *reinterpret_cast<uint4 *>(dst + 2 * i) =
                        make_uint4(e1_pack_iq(acc[i]), e1_pack_iq(acc[i + 1]), e1_pack_iq(acc[i + 2]), e1_pack_iq(acc[i + 3]));
            } else {
#pragma unroll
                for (int i = 0; i < R; i++)
                    if (j0 + i < n_valid)
                        *reinterpret_cast<uint32_t *>(dst + 2 * i) = e1_pack_iq(acc[i]);
            }
        }
        __syncthreads(); /* all reads of this tile's block (and of s_tile[b]) are done */
    }
    if (n_exact)
        atomicAdd(&s_cnt, n_exact);
    __syncthreads();
    if (tid == 0 && A.counters && s_cnt)
        atomicAdd(&A.counters[0], s_cnt);
}

/* ------------------------------------------------------------------ synthesis, FLOAT path
 * E1B200_CFG_CBOC / E1B200_CFG_GAIN (include/e1b200.h): CBOC(6,1,1/11) sub-carrier and the
 * reference's computed-but-unused per-satellite gain.  Same persistent-CTA frame, same parameter blocks and
 * tables as e1_synth_kernel<16>; per (thread, channel) e1_channel_run_float evaluates the exact closed form per
 * sample at sub-chip resolution and adds +-(g alpha | g beta) x (2 cos, 2 sin) into 2 x 16 FP32 sums; the store is
 * the fused float -> int16 (round to nearest even, saturating) interleaved I/Q, 128 bits at a time. */
#define E1_FLOAT_RUN 16
__global__ void __launch_bounds__(E1_SYNTH_THREADS, 1) e1_synth_float_kernel(const e1_synth_args A, const float alpha, const float beta)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t *s_codes = reinterpret_cast<uint32_t *>(smem_raw);
    unsigned char *s_lut = smem_raw + E1_CODES_BYTES;
    const uint32_t blk_bytes = (uint32_t)e1_blk_bytes(A.max_chan);
    unsigned char *s_blk0 = smem_raw + E1_CODES_BYTES + E1_LUT_BYTES;
    __shared__ __align__(8) uint64_t s_bar[3];
    __shared__ unsigned long long s_cnt;
    __shared__ long s_tile[2];
    const int R = E1_FLOAT_RUN;
    const int tid = threadIdx.x;
    const long total_tiles = (long)A.n_epochs * A.tiles_per_epoch;
    long next_id = 0;
    if (tid == 0) {
        s_cnt = 0;
        e1_mbar_init(&s_bar[0], 1);
        e1_mbar_init(&s_bar[1], 1);
        e1_mbar_init(&s_bar[2], 1);
        e1_mbar_init_fence();
        e1_mbar_expect(&s_bar[0], E1_CODES_BYTES + E1_LUT_BYTES);
        e1_bulk_g2s(s_codes, A.codes, E1_CODES_BYTES, &s_bar[0]);
        e1_bulk_g2s(s_lut, A.lut, E1_LUT_BYTES, &s_bar[0]);
        s_tile[0] = blockIdx.x;
        if ((long)blockIdx.x < total_tiles) {
            e1_mbar_expect(&s_bar[1], blk_bytes);
            e1_bulk_g2s(s_blk0, A.blk + (size_t)blockIdx.x * blk_bytes, blk_bytes, &s_bar[1]);
        }
        next_id = (long)gridDim.x + atomicAdd(A.next_tile, 1u);
    }
    __syncthreads();
    e1_mbar_wait(&s_bar[0], 0);
    const unsigned char *lut_lane = s_lut + 4 * (tid & (E1C_LUT_REP - 1));
    const uint64_t bias_h = e1_bias_h(A.tc_code);
    const int j0 = tid * R;
    unsigned long long n_exact = 0;
    for (int it = 0;; it++) {
        const int b = it & 1;
        const long tile_id = s_tile[b];
        if (tile_id >= total_tiles)
            break;
        if (tid == 0) {
            s_tile[1 - b] = next_id;
            if (next_id < total_tiles) {
                e1_mbar_expect(&s_bar[2 - b], blk_bytes);
                e1_bulk_g2s(s_blk0 + (size_t)(1 - b) * blk_bytes, A.blk + (size_t)next_id * blk_bytes, blk_bytes, &s_bar[2 - b]);
                next_id = (long)gridDim.x + atomicAdd(A.next_tile, 1u);
            }
        }
        e1_mbar_wait(&s_bar[1 + b], (uint32_t)(it >> 1) & 1u);
        const unsigned char *blk = s_blk0 + (size_t)b * blk_bytes;
        const int nact = *reinterpret_cast<const int *>(blk);
        const e1_chan_par *par = reinterpret_cast<const e1_chan_par *>(blk + E1C_BLK_HEADER);
        const int e = (int)(tile_id / A.tiles_per_epoch), t = (int)(tile_id - (long)e * A.tiles_per_epoch);
        const int n_valid = min(A.tile, A.n_samp - t * A.tile);
        if (j0 < n_valid) {
            float fi[R], fq[R];
#pragma unroll
            for (int i = 0; i < R; i++)
                fi[i] = fq[i] = 0.0f;
            for (int a = 0; a < nact; a++) {
                const float g = __uint_as_float(par[a].pat_a);
                e1_channel_run_float(&par[a], s_codes, lut_lane, j0, R, fi, fq, g * alpha, g * beta, A.thr_carr, A.thr_code, bias_h, &n_exact);
            }
            int16_t *dst = A.out + ((size_t)e * A.n_samp + (size_t)t * A.tile + j0) * 2;
            uint32_t w[R];
#pragma unroll
            for (int i = 0; i < R; i++) /* the sink's little-endian (int16 I, int16 Q) pair (:536-537) */
                w[i] = ((uint32_t)e1_f2i16(fi[i]) & 0xffffu) | ((uint32_t)e1_f2i16(fq[i]) << 16);
            if (A.vec_ok && j0 + R <= n_valid) {
#pragma unroll
                for (int i = 0; i < R; i += 4)
                    *reinterpret_cast<uint4 *>(dst + 2 * i) = make_uint4(w[i], w[i + 1], w[i + 2], w[i + 3]);
            } else {
#pragma unroll
                for (int i = 0; i < R; i++)
                    if (j0 + i < n_valid)
                        *reinterpret_cast<uint32_t *>(dst + 2 * i) = w[i];
            }
        }
        __syncthreads();
    }
    if (n_exact)
        atomicAdd(&s_cnt, n_exact);
    __syncthreads();
    if (tid == 0 && A.counters && s_cnt)
        atomicAdd(&A.counters[0], s_cnt);
}

/* ------------------------------------------------------------------ synthesis, carry-walked runs, teams
 * The sample loop for 8192-sample tiles.  A thread owns NH x 16 consecutive samples of a tile (NH = 2: 32 samples,
 * NH = 4: 64) and walks the active channels with the sums in registers (e1_cw_add: one carrier start per thread and
 * channel, the table walked by carries, the per-(thread, channel) set-up paid once per 32 / 64 samples).  A tile
 * needs 8192 / (16 NH) threads: the CTA is TEAMS teams of that many threads (2 x 256 or 3 x 128) that walk their own
 * tiles independently -- own parameter buffers, own mbarriers, own named barrier -- and share the code and carrier
 * tables.  3 x 128 threads leave 168 registers per thread for the 64 sums of NH = 4. */
#define E1_CW_RUN(NH) ((NH) * E1C_MAX_RUN)
#define E1_CW_TEAM_THREADS(NH) (E1C_THREADS * E1C_MAX_RUN / E1_CW_RUN(NH))

template <int TEAM_THREADS>
__device__ __forceinline__ void e1_team_sync(int team)
{
    asm volatile("bar.sync %0, %1;" : : "r"(1 + team), "n"(TEAM_THREADS) : "memory");
}

template <int NH, int TEAMS>
__global__ void __launch_bounds__(TEAMS *E1_CW_TEAM_THREADS(NH), 1) e1_synth_cw_kernel(const e1_synth_args A)
{
    constexpr int TT = E1_CW_TEAM_THREADS(NH), RUN = E1_CW_RUN(NH), NT = TEAMS * TT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t *s_codes = reinterpret_cast<uint32_t *>(smem_raw);
    unsigned char *s_lut = smem_raw + E1_CODES_BYTES;
    const uint32_t blk_bytes = (uint32_t)e1_blk_bytes(A.max_chan);
    __shared__ __align__(8) uint64_t s_bar[1 + 2 * TEAMS]; /* [0] tables, [1 + 2 team + b] parameter buffer b of a team */
    __shared__ unsigned long long s_cnt;
    __shared__ long s_tile[TEAMS][2]; /* [team][b]: tile number whose block is (being) loaded into that buffer */

    const int tid = threadIdx.x, team = tid / TT, t = tid - team * TT;
    unsigned char *s_blk0 = smem_raw + E1_CODES_BYTES + E1_LUT_BYTES + (size_t)team * 2 * blk_bytes;
    uint64_t *bar = &s_bar[1 + 2 * team];
    const long total_tiles = (long)A.n_epochs * A.tiles_per_epoch;
    const long first_wave = (long)TEAMS * gridDim.x;
    long next_id = 0; /* team leader: the tile after the one in s_tile[team][...] */
    if (tid == 0) {
        s_cnt = 0;
        for (int i = 0; i < 1 + 2 * TEAMS; i++)
            e1_mbar_init(&s_bar[i], 1);
        e1_mbar_init_fence();
        if (A.use_bulk) {
            e1_mbar_expect(&s_bar[0], E1_CODES_BYTES + E1_LUT_BYTES);
            e1_bulk_g2s(s_codes, A.codes, E1_CODES_BYTES, &s_bar[0]);
            e1_bulk_g2s(s_lut, A.lut, E1_LUT_BYTES, &s_bar[0]);
        }
    }
    __syncthreads();
    if (t == 0) {
        const long first = (long)TEAMS * blockIdx.x + team;
        s_tile[team][0] = first;
        if (first < total_tiles) {
            e1_mbar_expect(&bar[0], blk_bytes);
            e1_bulk_g2s(s_blk0, A.blk + (size_t)first * blk_bytes, blk_bytes, &bar[0]);
        }
        next_id = first_wave + atomicAdd(A.next_tile, 1u);
    }
    __syncthreads();
    if (A.use_bulk) {
        e1_mbar_wait(&s_bar[0], 0);
    } else {
        for (int i = tid; i < E1_CODES_BYTES / 4; i += NT)
            s_codes[i] = A.codes[i];
        for (int i = tid; i < E1_LUT_ENTRIES; i += NT)
            reinterpret_cast<int32_t *>(s_lut)[i] = A.lut[i];
        __syncthreads();
    }

    const uint32_t codes_s = e1_smem_u32(s_codes);
    const uint32_t lut_s = e1_smem_u32(s_lut) + 4u * (uint32_t)(tid & (E1C_LUT_REP - 1)); /* this lane's copy of every entry */
    const uint32_t lim_code = e1_lim_code(A.tc_code, A.thr_code);
    const uint32_t one = (uint32_t)A.use_bulk | 1u; /* the value 1, opaque to ptxas (e1_sample_loop_cw) */
    const int j0 = t * RUN;
    unsigned long long n_exact = 0;
    for (int it = 0;; it++) {
        const int b = it & 1;
        const long tile_id = s_tile[team][b];
        if (tile_id >= total_tiles)
            break;
        /* the whole team is past the previous tile (barrier at the end of the loop body), so the other
           buffer is free: start the next tile's block now, then draw the tile after it */
        if (t == 0) {
            s_tile[team][1 - b] = next_id;
            if (next_id < total_tiles) {
                e1_mbar_expect(&bar[1 - b], blk_bytes);
                e1_bulk_g2s(s_blk0 + (size_t)(1 - b) * blk_bytes, A.blk + (size_t)next_id * blk_bytes, blk_bytes, &bar[1 - b]);
                next_id = first_wave + atomicAdd(A.next_tile, 1u);
            }
        }
        e1_mbar_wait(&bar[b], (uint32_t)(it >> 1) & 1u);
        const unsigned char *blk = s_blk0 + (size_t)b * blk_bytes;
        const int nact = *reinterpret_cast<const int *>(blk);
        const e1_chan_par *par = reinterpret_cast<const e1_chan_par *>(blk + E1C_BLK_HEADER);
        const int e = (int)(tile_id / A.tiles_per_epoch), tt = (int)(tile_id - (long)e * A.tiles_per_epoch);
        const int n_valid = min(A.tile, A.n_samp - tt * A.tile);

        if (j0 < n_valid) {
            int acc[RUN];
#pragma unroll
            for (int i = 0; i < RUN; i++)
                acc[i] = 0;
            for (int a = 0; a < nact; a++) {
                const uint32_t rc = e1_cw_add<NH>(&par[a], codes_s, lut_s, j0, acc, A.thr_carr, lim_code, one);
                if (rc) { /* 1 % of the tiles: the out-of-line forms (tracking on / 64-bit positions / generic); their terms come back
                             in an array of their own so that the accumulators never have to live in local memory */
                    int d[RUN];
                    e1_cw_rest<NH>(&par[a], codes_s, lut_s, j0, d, rc, A.thr_carr, A.thr_code, A.tc_carr, A.tc_code, &n_exact);
#pragma unroll
                    for (int i = 0; i < RUN; i++)
                        acc[i] += d[i];
                }
            }
            /* a6 + sink format (:536-537): (short)I, (short)Q interleaved; acc = I + 65536*Q */
            int16_t *dst = A.out + ((size_t)e * A.n_samp + (size_t)tt * A.tile + j0) * 2;
            if (A.vec_ok && j0 + RUN <= n_valid) {
#pragma unroll
                for (int i = 0; i < RUN; i += 4)
                    *reinterpret_cast<uint4 *>(dst + 2 * i) =
                        make_uint4(e1_pack_iq(acc[i]), e1_pack_iq(acc[i + 1]), e1_pack_iq(acc[i + 2]), e1_pack_iq(acc[i + 3]));
            } else {
#pragma unroll
                for (int i = 0; i < RUN; i++)
                    if (j0 + i < n_valid)
                        *reinterpret_cast<uint32_t *>(dst + 2 * i) = e1_pack_iq(acc[i]);
            }
        }
        e1_team_sync<TT>(team); /* all reads of this tile's block (and of s_tile[team][b]) are done */
    }
    if (n_exact)
        atomicAdd(&s_cnt, n_exact);
    __syncthreads();
    if (tid == 0 && A.counters && s_cnt)
        atomicAdd(&A.counters[0], s_cnt);
}

/* ------------------------------------------------------------------ synthesis, event-driven (high sample rates)
 * e1_ev_context(fs, run): a half-chip and a carrier-table entry last several samples, so a channel's term is piecewise
 * constant; see "event-driven runs" in e1_core.h.  Same frame as e1_synth_cw_kernel<4, .> -- persistent CTA, TEAMS teams
 * of 128 threads, a team owns a tile at a time, parameter blocks double-buffered by bulk copy, thread t owns the 64
 * samples [64 t, 64 t + 64) of the tile -- but the sums live in shared memory as a column of DIFFERENCES per thread
 * (entry k of lane l of a warp at 128 k + 4 l: a warp's 32 accesses fall into 32 banks whatever the k are), the
 * channels add their change points to it (e1_ev_add: red.shared) and one running sum per tile turns the column into the
 * samples (a6 + the sink format).  Shared memory: ONE copy of the carrier table and its two difference tables (7 728 B: a
 * table read per event, not per sample), the parameter buffers, 32 KB of columns per team; the code words (two per thread
 * and channel) come from global memory through the read-only path, which is what lets five teams share an SM. */
#define E1_EV_TEAM_THREADS (E1C_THREADS * E1C_MAX_RUN / E1C_EV_RUN)
#define E1_LUT1_BYTES (3 * E1C_LUT1_WORDS * 4) /* the table and its two difference tables (e1_build_lut1) */
#define E1_EV_SMEM(teams, max_chan)                                                                                    \
    (E1_LUT1_BYTES + (teams) * 2 * (int)e1_blk_bytes(max_chan) + (teams) * E1_EV_TEAM_THREADS * E1C_EV_RUN * 4)
#define E1_EV_MAX_TEAMS 5

__device__ __noinline__ void e1_ev_rest(const e1_chan_par *p, uint32_t lut1_s, const uint32_t *codes_g, const unsigned char *lut_lane_g,
                                        int j0, int n, uint32_t col, uint32_t thr_carr, uint32_t thr_code, uint32_t tc_code,
                                        unsigned long long *n_exact)
{
    /* the casts are no-ops in the device pass (e1_sptr / e1_dptr are 32-bit shared-memory addresses there); they let the host
       pass, which sees the pointer flavour of the two types, parse this */
    e1_ev_rest_impl(p, (e1_sptr)(size_t)lut1_s, codes_g, lut_lane_g, j0, n, (e1_dptr)(size_t)col, thr_carr, thr_code, tc_code, n_exact,
                    (unsigned long long *)0);
}

template <int TEAMS>
__global__ void __launch_bounds__(TEAMS *E1_EV_TEAM_THREADS, 1) e1_synth_ev_kernel(const e1_synth_args A, const int32_t *lut1)
{
    constexpr int TT = E1_EV_TEAM_THREADS, RUN = E1C_EV_RUN;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    unsigned char *s_lut1 = smem_raw;
    const uint32_t blk_bytes = (uint32_t)e1_blk_bytes(A.max_chan);
    __shared__ __align__(8) uint64_t s_bar[1 + 2 * TEAMS]; /* [0] tables, [1 + 2 team + b] parameter buffer b of a team */
    __shared__ unsigned long long s_cnt;
    __shared__ long s_tile[TEAMS][2];

    const int tid = threadIdx.x, team = tid / TT, t = tid - team * TT;
    unsigned char *s_blk0 = smem_raw + E1_LUT1_BYTES + (size_t)team * 2 * blk_bytes;
    unsigned char *s_diff = smem_raw + E1_LUT1_BYTES + (size_t)TEAMS * 2 * blk_bytes;
    uint64_t *bar = &s_bar[1 + 2 * team];
    const long total_tiles = (long)A.n_epochs * A.tiles_per_epoch;
    const long first_wave = (long)TEAMS * gridDim.x;
    long next_id = 0;
    if (tid == 0) {
        s_cnt = 0;
        for (int i = 0; i < 1 + 2 * TEAMS; i++)
            e1_mbar_init(&s_bar[i], 1);
        e1_mbar_init_fence();
        e1_mbar_expect(&s_bar[0], E1_LUT1_BYTES);
        e1_bulk_g2s(s_lut1, lut1, E1_LUT1_BYTES, &s_bar[0]);
    }
    __syncthreads();
    if (t == 0) {
        const long first = (long)TEAMS * blockIdx.x + team;
        s_tile[team][0] = first;
        if (first < total_tiles) {
            e1_mbar_expect(&bar[0], blk_bytes);
            e1_bulk_g2s(s_blk0, A.blk + (size_t)first * blk_bytes, blk_bytes, &bar[0]);
        }
        next_id = first_wave + atomicAdd(A.next_tile, 1u);
    }
    /* this thread's column of differences: warp w of the CTA owns 8 KB, entry k of lane l at 128 k + 4 l */
    const uint32_t col = e1_smem_u32(s_diff) + (uint32_t)(tid >> 5) * (32u * RUN * 4u) + 4u * (uint32_t)(tid & 31);
#pragma unroll 8
    for (int k = 0; k < RUN; k++)
        asm volatile("st.shared.u32 [%0], %1;" : : "r"(col + 128u * (uint32_t)k), "r"(0u) : "memory");
    __syncthreads();
    e1_mbar_wait(&s_bar[0], 0);

    const uint32_t lut1_s = e1_smem_u32(s_lut1);
    const unsigned char *lut_lane_g = reinterpret_cast<const unsigned char *>(A.lut) + 4 * (tid & (E1C_LUT_REP - 1));
    const uint32_t tc_cw = e1_tc_carr_cw(A.thr_carr, RUN), lim_cw = e1_lim_carr_cw(A.thr_carr, RUN);
    const int j0 = t * RUN;
    unsigned long long n_exact = 0;
    for (int it = 0;; it++) {
        const int b = it & 1;
        const long tile_id = s_tile[team][b];
        if (tile_id >= total_tiles)
            break;
        if (t == 0) {
            s_tile[team][1 - b] = next_id;
            if (next_id < total_tiles) {
                e1_mbar_expect(&bar[1 - b], blk_bytes);
                e1_bulk_g2s(s_blk0 + (size_t)(1 - b) * blk_bytes, A.blk + (size_t)next_id * blk_bytes, blk_bytes, &bar[1 - b]);
                next_id = first_wave + atomicAdd(A.next_tile, 1u);
            }
        }
        e1_mbar_wait(&bar[b], (uint32_t)(it >> 1) & 1u);
        const unsigned char *blk = s_blk0 + (size_t)b * blk_bytes;
        const int nact = *reinterpret_cast<const int *>(blk);
        const e1_chan_par *par = reinterpret_cast<const e1_chan_par *>(blk + E1C_BLK_HEADER);
        const int e = (int)(tile_id / A.tiles_per_epoch), tt = (int)(tile_id - (long)e * A.tiles_per_epoch);
        const int n_valid = min(A.tile, A.n_samp - tt * A.tile);

        if (j0 < n_valid) {
            const int n = min(RUN, n_valid - j0);
            for (int a = 0; a < nact; a++) {
                const uint32_t want = E1_PAR_EV | E1_PAR_CLEAN;
                const int jw = par[a].j_w;
                if ((par[a].misc & want) == want && n == RUN && !(jw > j0 && jw < j0 + RUN)) {
                    uint64_t H;
                    uint32_t w0, w1;
                    e1_ev_fetch(&par[a], A.codes, j0, &H, &w0, &w1);
                    e1_ev_run64(&par[a], H, w0, w1, (e1_sptr)(size_t)lut1_s, j0, (e1_dptr)(size_t)col, tc_cw, lim_cw);
                } else
                    e1_ev_rest(&par[a], lut1_s, A.codes, lut_lane_g, j0, n, col, A.thr_carr, A.thr_code, A.tc_code, &n_exact);
            }
            /* running sum of the column = the samples; a6 + sink format (:536-537); the column is left zeroed */
            int16_t *dst = A.out + ((size_t)e * A.n_samp + (size_t)tt * A.tile + j0) * 2;
            int sum = 0;
            if (A.vec_ok && n == RUN) {
#pragma unroll
                for (int i = 0; i < RUN; i += 4) {
                    uint32_t w[4];
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        uint32_t v;
                        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(col + 128u * (uint32_t)(i + q)) : "memory");
                        asm volatile("st.shared.u32 [%0], %1;" : : "r"(col + 128u * (uint32_t)(i + q)), "r"(0u) : "memory");
                        sum += (int)v;
                        w[q] = e1_pack_iq(sum);
                    }
                    *reinterpret_cast<uint4 *>(dst + 2 * i) = make_uint4(w[0], w[1], w[2], w[3]);
                }
            } else {
                for (int i = 0; i < RUN; i++) {
                    uint32_t v;
                    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(col + 128u * (uint32_t)i) : "memory");
                    asm volatile("st.shared.u32 [%0], %1;" : : "r"(col + 128u * (uint32_t)i), "r"(0u) : "memory");
                    sum += (int)v;
                    if (i < n)
                        *reinterpret_cast<uint32_t *>(dst + 2 * i) = e1_pack_iq(sum);
                }
            }
        }
        e1_team_sync<TT>(team); /* all reads of this tile's block (and of s_tile[team][b]) are done */
    }
    if (n_exact)
        atomicAdd(&s_cnt, n_exact);
    __syncthreads();
    if (tid == 0 && A.counters && s_cnt)
        atomicAdd(&A.counters[0], s_cnt);
}

#endif /* E1_KERNELS_CUH */
