// Micro-benchmark of the integer issue rates the K1 scan kernel lives on (sm_100a): warp
// instructions per clock per SM sub-partition for single opcodes and for ALU + FMA mixes.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o tools/bin/pipe_bench tools/pipe_bench.cu
// Every thread runs 8 independent dependency chains of the opcode under test; 32 warps per SM.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CHAINS 8
#define ITER 2048

#define OP_LOP3(x, y) asm volatile("lop3.b32 %0, %0, %1, 0x5a5a5a5a, 0x96;" : "+r"(x) : "r"(y))
#define OP_SHF(x, y) asm volatile("shf.r.wrap.b32 %0, %0, %1, 8;" : "+r"(x) : "r"(y))
#define OP_PRMT(x, y) asm volatile("prmt.b32 %0, %0, %1, 0x4321;" : "+r"(x) : "r"(y))
#define OP_ADD(x, y) asm volatile("add.u32 %0, %0, %1;" : "+r"(x) : "r"(y))
#define OP_SEL(x, y) asm volatile("{.reg .pred p; setp.ne.u32 p, %1, 0; selp.u32 %0, %0, %1, p;}" : "+r"(x) : "r"(y))
#define OP_IMAD(x, y) asm volatile("mad.lo.u32 %0, %0, %1, %1;" : "+r"(x) : "r"(y))
#define OP_IMADC(x, y) asm volatile("mad.lo.u32 %0, %0, 0x114253d5, %1;" : "+r"(x) : "r"(y))
#define OP_MULHI(x, y) asm volatile("mul.hi.u32 %0, %0, %1;" : "+r"(x) : "r"(y))
#define OP_MULHIC(x, y) asm volatile("mul.hi.u32 %0, %0, 16777216;" : "+r"(x))
#define OP_WIDE(x, y) asm volatile("{.reg .u64 w; .reg .u32 lo, hi; mul.wide.u32 w, %0, %1; mov.b64 {lo, hi}, w; xor.b32 %0, lo, hi;}" : "+r"(x) : "r"(y))
#define OP_WIDEONLY(x, y) asm volatile("{.reg .u64 w; .reg .u32 lo; mul.wide.u32 w, %0, %1; mov.b64 {lo, %0}, w;}" : "+r"(x) : "r"(y))
#define OP_SHL(x, y) asm volatile("shl.b32 %0, %0, 3;" : "+r"(x))
#define OP_SHR(x, y) asm volatile("shr.u32 %0, %0, 3;" : "+r"(x))
#define OP_MIX_AF(x, y) do { OP_LOP3(x, y); OP_IMAD(x, y); } while (0)
#define OP_MIX_AAF(x, y) do { OP_LOP3(x, y); OP_SHF(x, y); OP_IMAD(x, y); } while (0)
#define OP_MIX_AFF(x, y) do { OP_LOP3(x, y); OP_IMAD(x, y); OP_IMADC(x, y); } while (0)
#define OP_MIX_AH(x, y) do { OP_LOP3(x, y); OP_MULHI(x, y); } while (0)
#define OP_MIX_AW(x, y) do { OP_LOP3(x, y); OP_WIDEONLY(x, y); } while (0)
#define OP_MIX_AI(x, y) do { OP_LOP3(x, y); OP_ADD(x, y); } while (0)
#define OP_MIX_FI(x, y) do { OP_IMAD(x, y); OP_ADD(x, y); } while (0)
#define OP_MIX_AFI(x, y) do { OP_LOP3(x, y); OP_IMAD(x, y); OP_ADD(x, y); } while (0)
#define OP_MIX_AAFI(x, y) do { OP_LOP3(x, y); OP_SHF(x, y); OP_IMAD(x, y); OP_ADD(x, y); } while (0)
#define OP_VIADD(x, y) asm volatile("add.u32 %0, %0, 12345;" : "+r"(x))
#define OP_SETP(x, y) asm volatile("{.reg .pred p; setp.lt.u32 p, %0, %1; @p add.u32 %0, %0, 7;}" : "+r"(x) : "r"(y))
#define OP_MIN(x, y) asm volatile("min.u32 %0, %0, %1; add.u32 %0, %0, 3;" : "+r"(x) : "r"(y))

#define KERNEL(NAME, OP, NINST)                                                                   \
    __global__ void __launch_bounds__(1024) k_##NAME(uint32_t *out, uint32_t seed, unsigned long long *cyc) { \
        uint32_t v[CHAINS];                                                                       \
        for (int c = 0; c < CHAINS; c++) v[c] = seed + threadIdx.x * 17 + c;                      \
        const uint32_t y = seed | 1;                                                              \
        __syncthreads();                                                                          \
        const long long t0 = clock64();                                                           \
        for (int i = 0; i < ITER; i++) {                                                          \
            _Pragma("unroll") for (int c = 0; c < CHAINS; c++) { OP(v[c], y); }                   \
        }                                                                                         \
        __syncthreads();                                                                          \
        const long long t1 = clock64();                                                           \
        uint32_t s = 0;                                                                           \
        for (int c = 0; c < CHAINS; c++) s ^= v[c];                                               \
        out[blockIdx.x * blockDim.x + threadIdx.x] = s;                                           \
        if (threadIdx.x == 0) cyc[blockIdx.x] = (unsigned long long)(t1 - t0);                    \
    }                                                                                             \
    static void run_##NAME(uint32_t *out, unsigned long long *cyc, int sms) {                     \
        k_##NAME<<<sms, 1024>>>(out, 12345u, cyc);                                                \
        cudaDeviceSynchronize();                                                                  \
        k_##NAME<<<sms, 1024>>>(out, 12345u, cyc);                                                \
        cudaDeviceSynchronize();                                                                  \
        unsigned long long h[256];                                                                \
        cudaMemcpy(h, cyc, sizeof(unsigned long long) * sms, cudaMemcpyDeviceToHost);             \
        double avg = 0;                                                                           \
        for (int i = 0; i < sms; i++) avg += (double)h[i];                                        \
        avg /= sms;                                                                               \
        const double winst = 32.0 * ITER * CHAINS * NINST; /* warp instructions per SM */         \
        printf("%-10s %d inst/iter  %8.0f cycles  %.3f warp-inst/clk/SM  %.3f per sub-partition\n", #NAME, NINST, avg, \
               winst / avg, winst / avg / 4.0);                                                   \
    }

KERNEL(lop3, OP_LOP3, 1)
KERNEL(shf, OP_SHF, 1)
KERNEL(prmt, OP_PRMT, 1)
KERNEL(add, OP_ADD, 1)
KERNEL(sel, OP_SEL, 2)
KERNEL(imad, OP_IMAD, 1)
KERNEL(imadc, OP_IMADC, 1)
KERNEL(mulhi, OP_MULHI, 1)
KERNEL(mulhic, OP_MULHIC, 1)
KERNEL(wide, OP_WIDE, 2)
KERNEL(wideonly, OP_WIDEONLY, 1)
KERNEL(shl, OP_SHL, 1)
KERNEL(shr, OP_SHR, 1)
KERNEL(mix_af, OP_MIX_AF, 2)
KERNEL(mix_aaf, OP_MIX_AAF, 3)
KERNEL(mix_aff, OP_MIX_AFF, 3)
KERNEL(mix_ah, OP_MIX_AH, 2)
KERNEL(mix_aw, OP_MIX_AW, 2)
KERNEL(mix_ai, OP_MIX_AI, 2)
KERNEL(mix_fi, OP_MIX_FI, 2)
KERNEL(mix_afi, OP_MIX_AFI, 3)
KERNEL(mix_aafi, OP_MIX_AAFI, 4)
KERNEL(viadd, OP_VIADD, 1)
KERNEL(setp_padd, OP_SETP, 2)
KERNEL(min_add, OP_MIN, 2)

int main() {
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    uint32_t *out;
    unsigned long long *cyc;
    cudaMalloc(&out, sizeof(uint32_t) * 1024 * sms);
    cudaMalloc(&cyc, sizeof(unsigned long long) * sms);
    printf("SMs %d, 32 warps per SM, %d chains per thread, %d iterations\n", sms, CHAINS, ITER);
    run_lop3(out, cyc, sms); run_shf(out, cyc, sms); run_prmt(out, cyc, sms); run_add(out, cyc, sms);
    run_sel(out, cyc, sms); run_imad(out, cyc, sms); run_imadc(out, cyc, sms); run_mulhi(out, cyc, sms);
    run_mulhic(out, cyc, sms); run_wide(out, cyc, sms); run_wideonly(out, cyc, sms); run_shl(out, cyc, sms);
    run_shr(out, cyc, sms); run_mix_af(out, cyc, sms); run_mix_aaf(out, cyc, sms); run_mix_aff(out, cyc, sms);
    run_mix_ah(out, cyc, sms); run_mix_aw(out, cyc, sms);
    run_mix_ai(out, cyc, sms); run_mix_fi(out, cyc, sms); run_mix_afi(out, cyc, sms); run_mix_aafi(out, cyc, sms);
    run_viadd(out, cyc, sms); run_setp_padd(out, cyc, sms); run_min_add(out, cyc, sms);
    return 0;
}
