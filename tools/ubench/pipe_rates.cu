/* Issue-rate probe for the integer instructions the synthesis loop is made of (B200, sm_100a).
 *   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o pipe_rates pipe_rates.cu && ./pipe_rates
 * Each kernel runs N_IT iterations of 32 independent chains of ONE instruction kind per thread
 * (inline PTX; check the SASS with cuobjdump before trusting a line), with 1 CTA of 512 threads
 * per SM like the synthesis kernel (4 warps per scheduler).  Reported: warp-instructions per clock
 * per SM sub-partition. */
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define N_IT 2048
#define CH 16

#define KERNEL(NAME, DECL, INIT, STEP, FOLD)                                                            \
    __global__ void __launch_bounds__(512, 1) NAME(uint32_t *out, uint32_t a, uint32_t b, long long *clk) \
    {                                                                                                   \
        DECL;                                                                                           \
        _Pragma("unroll") for (int c = 0; c < CH; c++) { INIT; }                                        \
        __syncthreads();                                                                                \
        const long long t0 = clock64();                                                                 \
        for (int it = 0; it < N_IT; it++) {                                                             \
            _Pragma("unroll") for (int c = 0; c < CH; c++) { STEP; }                                    \
        }                                                                                               \
        const long long t1 = clock64();                                                                 \
        uint32_t r = 0;                                                                                 \
        _Pragma("unroll") for (int c = 0; c < CH; c++) { FOLD; }                                        \
        out[blockIdx.x * blockDim.x + threadIdx.x] = r;                                                 \
        if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;                                                \
    }

KERNEL(k_imad, uint32_t x[CH], x[c] = threadIdx.x + c, asm volatile("mad.lo.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(a), "r"(b)), r += x[c])
KERNEL(k_imad_hi, uint32_t x[CH], x[c] = threadIdx.x + c, asm volatile("mad.hi.u32 %0, %0, %1, %2;" : "+r"(x[c]) : "r"(a), "r"(b)), r += x[c])
KERNEL(k_imad_wide_acc, uint64_t x[CH], x[c] = threadIdx.x + c, asm volatile("mad.wide.u32 %0, %1, %2, %0;" : "+l"(x[c]) : "r"((uint32_t)x[c]), "r"(b)),
       r += (uint32_t)x[c] + (uint32_t)(x[c] >> 32))
KERNEL(k_imad_wide_imm, uint64_t x[CH], x[c] = threadIdx.x + c, asm volatile("mad.wide.u32 %0, %1, 511, %0;" : "+l"(x[c]) : "r"((uint32_t)x[c])),
       r += (uint32_t)x[c] + (uint32_t)(x[c] >> 32))
KERNEL(k_iadd3, uint32_t x[CH], x[c] = threadIdx.x + c, asm volatile("{.reg .u32 t; add.cc.u32 %0, %0, %1; addc.u32 t, 0, 0; xor.b32 %0, %0, t;}" : "+r"(x[c]) : "r"(a)), r += x[c])
KERNEL(k_add64, uint64_t x[CH], x[c] = threadIdx.x + c, asm volatile("add.u64 %0, %0, %1;" : "+l"(x[c]) : "l"(((uint64_t)b << 32) | a)),
       r += (uint32_t)x[c] + (uint32_t)(x[c] >> 32))
KERNEL(k_shf, uint32_t x[CH], x[c] = threadIdx.x + c, asm volatile("shf.l.wrap.b32 %0, %1, %0, 3;" : "+r"(x[c]) : "r"(a)), r += x[c])
KERNEL(k_lop3, uint32_t x[CH], x[c] = threadIdx.x + c, asm volatile("lop3.b32 %0, %0, %1, %2, 0x96;" : "+r"(x[c]) : "r"(a), "r"(b)), r += x[c])
KERNEL(k_min3, uint32_t x[CH], x[c] = threadIdx.x * 2654435761u + c * 40503u,
       asm volatile("{.reg .u32 t; min.u32 t, %0, %1; min.u32 %0, t, %2;}" : "+r"(x[c]) : "r"(x[(c + 1) % CH]), "r"(x[(c + 5) % CH])), r += x[c])
/* pairs on independent registers: do two kinds issue in the same clocks (separate pipes)? */
#define PAIR(NAME, OPA, OPB) \
    KERNEL(NAME, uint32_t x[CH]; uint32_t y[CH], x[c] = threadIdx.x + c; y[c] = threadIdx.x * 7 + c, asm volatile(OPA "\n\t" OPB : "+r"(x[c]), "+r"(y[c]) : "r"(a), "r"(b)), r += x[c] + y[c])
#define OP_IMAD "mad.lo.u32 %0, %0, %2, %3;"
#define OP_IMAD_Y "mad.lo.u32 %1, %1, %2, %3;"
#define OP_SHF_Y "shf.l.wrap.b32 %1, %2, %1, 3;"
#define OP_LOP_Y "lop3.b32 %1, %1, %2, %3, 0x96;"
#define OP_LOP_X "lop3.b32 %0, %0, %2, %3, 0x96;"
#define OP_ADD3_Y "{.reg .u32 t; add.u32 t, %1, %2; add.u32 %1, t, %3;}"
#define OP_ADD3_X "{.reg .u32 t; add.u32 t, %0, %2; add.u32 %0, t, %3;}"
#define OP_ADD2_Y "add.u32 %1, %1, %2;"
#define OP_MIN_Y "{.reg .u32 t; min.u32 t, %1, %2; max.u32 %1, t, %3;}"
#define OP_HI_Y "mad.hi.u32 %1, %1, %2, %3;"
PAIR(k_p_imad_imad, OP_IMAD, OP_IMAD_Y)
PAIR(k_p_imad_shf, OP_IMAD, OP_SHF_Y)
PAIR(k_p_imad_lop, OP_IMAD, OP_LOP_Y)
PAIR(k_p_imad_add3, OP_IMAD, OP_ADD3_Y)
PAIR(k_p_imad_add2, OP_IMAD, OP_ADD2_Y)
PAIR(k_p_imad_hi, OP_IMAD, OP_HI_Y)
PAIR(k_p_lop_shf, OP_LOP_X, OP_SHF_Y)
PAIR(k_p_lop_add3, OP_LOP_X, OP_ADD3_Y)
PAIR(k_p_add3_add3, OP_ADD3_X, OP_ADD3_Y)
PAIR(k_p_lop_add2, OP_LOP_X, OP_ADD2_Y)
/* shared-memory lookups like the carrier table's FIRST layout (v4): 64-byte entries, one 4-byte copy per (lane & 15),
 * random entries per lane -> lanes l and l+16 collide.  (The table now has one copy per lane.) */
__global__ void __launch_bounds__(512, 1) k_lds(uint32_t *out, uint32_t a, uint32_t b, long long *clk)
{
    extern __shared__ uint32_t sm[];
    for (int i = threadIdx.x; i < 576 * 16; i += 512) sm[i] = (i * 2654435761u) >> 7;
    uint32_t x[CH];
    const uint32_t lane_off = (threadIdx.x & 15) * 4;
    _Pragma("unroll") for (int c = 0; c < CH; c++) x[c] = threadIdx.x * 40503u + c * 977u;
    __syncthreads();
    const uint32_t base = (uint32_t)__cvta_generic_to_shared(sm);
    const long long t0 = clock64();
    for (int it = 0; it < N_IT; it++) {
        _Pragma("unroll") for (int c = 0; c < CH; c++)
            asm volatile("{.reg .u32 t; lop3.b32 t, %0, 0x7fc0, %1, 0xea; add.u32 t, t, %2; ld.shared.u32 %0, [t];}" : "+r"(x[c]) : "r"(lane_off), "r"(base));
    }
    const long long t1 = clock64();
    uint32_t r = 0;
    _Pragma("unroll") for (int c = 0; c < CH; c++) r += x[c];
    out[blockIdx.x * blockDim.x + threadIdx.x] = r + a + b;
    if (threadIdx.x == 0) clk[blockIdx.x] = t1 - t0;
}

template <typename K>
static void run(const char *name, K kern, int per_step, uint32_t *d_out, long long *d_clk, int sms)
{
    kern<<<sms, 512>>>(d_out, 3u, 5u, d_clk);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    cudaEventRecord(e0);
    kern<<<sms, 512>>>(d_out, 3u, 5u, d_clk);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    long long clk[256];
    cudaMemcpy(clk, d_clk, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
    double avg = 0;
    for (int i = 0; i < sms; i++) avg += (double)clk[i];
    avg /= sms;
    /* 16 warps per SM = 4 per sub-partition; CH * N_IT steps of per_step instructions each */
    const double inst_per_smsp = 4.0 * CH * N_IT * per_step;
    printf("%-22s %8.3f ms  %12.0f clk  %.3f warp-inst/clk/SMSP (counting %d inst per step)\n", name, ms, avg, inst_per_smsp / avg, per_step);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    const int sms = p.multiProcessorCount;
    printf("%s, %d SMs\n", p.name, sms);
    uint32_t *d_out;
    long long *d_clk;
    cudaMalloc(&d_out, sizeof(uint32_t) * 512 * sms);
    cudaMalloc(&d_clk, sizeof(long long) * sms);
    run("IMAD", k_imad, 1, d_out, d_clk, sms);
    run("IMAD.HI", k_imad_hi, 1, d_out, d_clk, sms);
    run("IMAD.WIDE reg +64", k_imad_wide_acc, 1, d_out, d_clk, sms);
    run("IMAD.WIDE imm +64", k_imad_wide_imm, 1, d_out, d_clk, sms);
    run("IADD3.cc + SEL + LOP3", k_iadd3, 3, d_out, d_clk, sms);
    run("add.u64 (IADD3+IADD3.X)", k_add64, 2, d_out, d_clk, sms);
    run("SHF", k_shf, 1, d_out, d_clk, sms);
    run("LOP3", k_lop3, 1, d_out, d_clk, sms);
    run("VIMNMX3", k_min3, 1, d_out, d_clk, sms);
    run("pair IMAD | IMAD", k_p_imad_imad, 2, d_out, d_clk, sms);
    run("pair IMAD | SHF", k_p_imad_shf, 2, d_out, d_clk, sms);
    run("pair IMAD | LOP3", k_p_imad_lop, 2, d_out, d_clk, sms);
    run("pair IMAD | add3", k_p_imad_add3, 2, d_out, d_clk, sms);
    run("pair IMAD | add2", k_p_imad_add2, 2, d_out, d_clk, sms);
    run("pair IMAD | IMAD.HI", k_p_imad_hi, 2, d_out, d_clk, sms);
    run("pair LOP3 | SHF", k_p_lop_shf, 2, d_out, d_clk, sms);
    run("pair LOP3 | add3", k_p_lop_add3, 2, d_out, d_clk, sms);
    run("pair add3 | add3", k_p_add3_add3, 2, d_out, d_clk, sms);
    run("pair LOP3 | add2", k_p_lop_add2, 2, d_out, d_clk, sms);
    {
        cudaFuncSetAttribute(k_lds, cudaFuncAttributeMaxDynamicSharedMemorySize, 576 * 64);
        for (int rep = 0; rep < 2; rep++) k_lds<<<sms, 512, 576 * 64>>>(d_out, 3u, 5u, d_clk);
        cudaDeviceSynchronize();
        long long clk[256];
        cudaMemcpy(clk, d_clk, sizeof(long long) * sms, cudaMemcpyDeviceToHost);
        double avg = 0;
        for (int i = 0; i < sms; i++) avg += (double)clk[i];
        avg /= sms;
        printf("%-22s %12.0f clk  %.3f LDS/clk/SMSP (random 64-byte entries, 16 copies: <= 2-way conflicts; + LOP3 + IADD per load)\n", "LDS table lookup", avg,
               4.0 * CH * N_IT / avg);
    }
    return 0;
}
