/* e1_kernels.cuh -- sm_100a kernels of the Galileo E1B/C synthesiser.
 *
 *   e1_restate_kernel     computeCodePhase (src/gal-sig.cpp:308-347) per (epoch, channel)
 *   e1_plan_code_kernel   exact code-phase / symbol checkpoints per (epoch, tile, channel)
 *   e1_plan_carr_kernel   exact carrier-phase checkpoints per (epoch, tile, channel)
 *   e1_synth_kernel       the sample loop (src/galileo-sdr.cpp:481-539): per sample, all
 *                         channels, int32 accumulate, packed int16 I/Q, 128-bit stores
 *
 * HBM layout
 *   recs   e1_epoch_rec[n_epochs][max_chan]                         176 B each (caller / H2D)
 *   ck     e1_tile_ck[n_epochs][tiles_per_epoch][max_chan]           32 B each (scratch)
 *   out    int16 I,Q interleaved, sample (epoch*N + k) at byte 4*(epoch*N + k)
 *   codes  uint32[50][516]: one 2-bit field per BOC(1,1) half-chip (see e1_core.h)   103 200 B
 *   lut    int32[2][4][512]: carrier term by (phase regime, code/symbol field, index) 16 384 B
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
#define E1_SYNTH_THREADS 512
#define E1_RUN E1C_RUN
#define E1_GROUP (E1_SYNTH_THREADS * E1_RUN)

struct e1_synth_args {
    const e1_epoch_rec *recs;
    const e1_tile_ck *ck;
    const double *delta; /* planner translation of each epoch's carrier checkpoints, channel-major:
                            delta[ch * delta_stride + delta_off + e] for epoch e of this launch          */
    int delta_stride, delta_off;
    const uint32_t *codes;
    const int32_t *lut;
    int16_t *out;
    unsigned long long *counters; /* [0] ambiguous samples resolved exactly, [1] planner errors */
    double delt;
    int n_epochs, n_samp, max_chan, tile, tiles_per_epoch;
    uint32_t thr_carr, thr_code;
    int vec_ok, use_bulk;
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
    o.flags = r.flags;
    o.reserved = 0;
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
    e1_prep *prep;   /* channel-major [max_chan][n_epochs], like g, dend, est, delta and units */
    double *g, *dend, *est, *delta;
    e1_unit *units;
    unsigned long long *counters; /* [2] serial epochs, [3] HAT epochs */
    double delt;
    int n_epochs, n_samp, max_chan, tile, tiles_per_epoch;
};

__global__ void e1_v2_prep_kernel(const e1_plan_args P)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x; /* record order: coalesced reads */
    if (i >= P.n_epochs * P.max_chan)
        return;
    int e = i / P.max_chan, ch = i - e * P.max_chan;
    e1_v2_prep(&P.recs[i], P.delt, &P.prep[(size_t)ch * P.n_epochs + e]);
}

__global__ void e1_v2_ideal_kernel(const e1_plan_args P)
{
    int ch = blockIdx.x; /* one channel per block: serial walks of different channels must not share a warp */
    if (threadIdx.x != 0)
        return;
    if (ch < P.max_chan)
        e1_v2_ideal_prefix(P.prep + (size_t)ch * P.n_epochs, P.n_epochs, P.phase[ch], P.n_samp, P.g + (size_t)ch * P.n_epochs);
}

__global__ void e1_v2_drift_kernel(const e1_plan_args P)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x; /* channel-major: a warp walks one channel's epochs */
    if (i >= P.n_epochs * P.max_chan)
        return;
    P.dend[i] = e1_v2_drift_unit(&P.prep[i], P.g[i], P.n_samp);
}

__global__ void e1_v2_estimate_kernel(const e1_plan_args P)
{
    int ch = blockIdx.x; /* one channel per block: serial walks of different channels must not share a warp */
    if (threadIdx.x != 0)
        return;
    if (ch >= P.max_chan)
        return;
    size_t o = (size_t)ch * P.n_epochs;
    e1_v2_estimate_prefix(P.prep + o, P.n_epochs, P.phase[ch], P.g + o, P.dend + o, P.est + o);
}

__global__ void e1_v2_span_kernel(const e1_plan_args P)
{
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P.n_epochs * P.max_chan)
        return;
    int ch = i / P.n_epochs, e = i - ch * P.n_epochs;
    e1_v2_span_unit(&P.prep[i], e ? &P.prep[i - 1] : nullptr, e, P.phase[ch], e ? P.est[i - 1] : 0.0, P.n_samp, P.tile,
                    P.tiles_per_epoch, P.ck + (size_t)e * P.tiles_per_epoch * P.max_chan + ch, P.max_chan, &P.units[i]);
}

__global__ void e1_v2_chain_kernel(const e1_plan_args P)
{
    int ch = blockIdx.x; /* one channel per block: serial walks of different channels must not share a warp */
    if (threadIdx.x != 0)
        return;
    if (ch >= P.max_chan)
        return;
    unsigned long long st[2] = {0, 0};
    size_t o = (size_t)ch * P.n_epochs;
    P.phase[ch] = e1_v2_chain(P.prep + o, P.n_epochs, P.phase[ch], P.n_samp, P.tile, P.tiles_per_epoch, P.units + o, P.ck + ch,
                              P.max_chan, (size_t)P.tiles_per_epoch * P.max_chan, P.delta + o, st);
    if (st[0])
        atomicAdd(&P.counters[2], st[0]);
    if (st[1])
        atomicAdd(&P.counters[3], st[1]);
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

template <int G>
__global__ void __launch_bounds__(E1_SYNTH_THREADS, 1) e1_synth_kernel(const e1_synth_args A)
{
    extern __shared__ __align__(128) unsigned char smem_raw[];
    uint32_t *s_codes = reinterpret_cast<uint32_t *>(smem_raw);
    int32_t *s_lut = reinterpret_cast<int32_t *>(smem_raw + E1_CODES_BYTES);
    e1_chan_par *s_par = reinterpret_cast<e1_chan_par *>(smem_raw + E1_CODES_BYTES + E1_LUT_BYTES);
    __shared__ __align__(8) uint64_t s_bar;
    __shared__ int s_nact;
    __shared__ unsigned long long s_cnt[2];

    const int tid = threadIdx.x;
    if (tid == 0) {
        s_cnt[0] = 0;
        s_cnt[1] = 0;
    }
    /* tables: one bulk copy each through the TMA unit, once per (persistent) CTA */
    if (A.use_bulk) {
        if (tid == 0) {
            e1_mbar_init(&s_bar, 1);
            e1_mbar_expect(&s_bar, E1_CODES_BYTES + E1_LUT_BYTES);
            e1_bulk_g2s(s_codes, A.codes, E1_CODES_BYTES, &s_bar);
            e1_bulk_g2s(s_lut, A.lut, E1_LUT_BYTES, &s_bar);
        }
        __syncthreads();
        e1_mbar_wait(&s_bar, 0);
    } else {
        for (int i = tid; i < E1_CODES_BYTES / 4; i += E1_SYNTH_THREADS)
            s_codes[i] = A.codes[i];
        for (int i = tid; i < E1_LUT_ENTRIES; i += E1_SYNTH_THREADS)
            s_lut[i] = A.lut[i];
    }

    const int total_tiles = A.n_epochs * A.tiles_per_epoch;
    for (int tile_id = blockIdx.x; tile_id < total_tiles; tile_id += gridDim.x) {
        const int e = tile_id / A.tiles_per_epoch, t = tile_id - e * A.tiles_per_epoch;
        __syncthreads(); /* previous tile done with s_par */
        if (tid == 0)
            s_nact = 0;
        __syncthreads();
        if (tid < A.max_chan) {
            const e1_tile_ck c = A.ck[(size_t)tile_id * A.max_chan + tid];
            if (c.sym & E1_CK_ACTIVE) {
                e1_chan_par p;
                e1_make_par(&c, &A.recs[(size_t)e * A.max_chan + tid], A.delt, A.tile, A.delta[(size_t)tid * A.delta_stride + A.delta_off + e], &p);
                if (c.sym & E1_CK_ERROR)
                    atomicAdd(&s_cnt[1], 1ull);
                s_par[atomicAdd(&s_nact, 1)] = p;
            }
        }
        __syncthreads();
        const int nact = s_nact;
        const int n_valid = min(A.tile, A.n_samp - t * A.tile);

        int acc[G][E1_RUN];
#pragma unroll
        for (int g = 0; g < G; g++)
#pragma unroll
            for (int i = 0; i < E1_RUN; i++)
                acc[g][i] = 0;
        uint32_t amb[G];
#pragma unroll
        for (int g = 0; g < G; g++)
            amb[g] = 0;
        const unsigned char *lutb = reinterpret_cast<const unsigned char *>(s_lut);
        for (int a = 0; a < nact; a++) {
#pragma unroll
            for (int g = 0; g < G; g++)
                amb[g] |= e1_run_fast(&s_par[a], s_codes, lutb, g * E1_GROUP + tid * E1_RUN, acc[g], A.thr_carr, A.thr_code);
        }
        /* rare: some channel flagged a sample of run g.  Find the channel(s) by re-running the fast
           form into a scratch accumulator, take their fast terms back out and add the exact ones. */
#pragma unroll
        for (int g = 0; g < G; g++) {
            if (amb[g]) {
                unsigned long long n_exact = 0;
                const int j0 = g * E1_GROUP + tid * E1_RUN;
                for (int a = 0; a < nact; a++) {
                    int t4[E1_RUN] = {0, 0, 0, 0};
                    if (e1_run_fast(&s_par[a], s_codes, lutb, j0, t4, A.thr_carr, A.thr_code)) {
#pragma unroll
                        for (int i = 0; i < E1_RUN; i++)
                            acc[g][i] -= t4[i];
                        e1_channel_run(&s_par[a], s_codes, s_lut, j0, acc[g], A.thr_carr, A.thr_code, 1, &n_exact);
                    }
                }
                atomicAdd(&s_cnt[0], n_exact);
            }
        }
        /* a6 + sink format (:536-537): (short)I, (short)Q interleaved; acc = I + 65536*Q */
        int16_t *out_tile = A.out + ((size_t)e * A.n_samp + (size_t)t * A.tile) * 2;
#pragma unroll
        for (int g = 0; g < G; g++) {
            const int j0 = g * E1_GROUP + tid * E1_RUN;
            uint32_t w[E1_RUN];
#pragma unroll
            for (int i = 0; i < E1_RUN; i++)
                w[i] = e1_pack_iq(acc[g][i]);
            if (A.vec_ok && j0 + E1_RUN <= n_valid) {
                *reinterpret_cast<uint4 *>(out_tile + (size_t)j0 * 2) = make_uint4(w[0], w[1], w[2], w[3]);
            } else {
#pragma unroll
                for (int i = 0; i < E1_RUN; i++)
                    if (j0 + i < n_valid)
                        *reinterpret_cast<uint32_t *>(out_tile + (size_t)(j0 + i) * 2) = w[i];
            }
        }
    }
    __syncthreads();
    if (tid == 0 && A.counters) {
        if (s_cnt[0])
            atomicAdd(&A.counters[0], s_cnt[0]);
        if (s_cnt[1])
            atomicAdd(&A.counters[1], s_cnt[1]);
    }
}

#endif /* E1_KERNELS_CUH */
