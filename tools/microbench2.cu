// Second integer-pipe microbenchmark (B200, sm_100a): is a plain IMAD.WIDE.U32 (64-bit accumulate, no carry flag) cheaper
// than the IMAD.WIDE.U32.X carry-chain form the saturated 8x32 field multiply uses?  Decides between saturated limbs with
// carry chains and unsaturated limbs with carry-free 64-bit accumulators.  Prints one JSON object.
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned int u32; typedef unsigned long long u64;

template <int OP>
__global__ void __launch_bounds__(256) k(u32* out, u32 seed, int iters) {
    u32 a0 = seed + threadIdx.x, a1 = a0 * 3 + 1, a2 = a0 * 5 + 2, a3 = a0 * 7 + 3, a4 = a0 * 11 + 4, a5 = a0 * 13 + 5, a6 = a0 * 17 + 6, a7 = a0 * 19 + 7;
    u64 w0 = a0, w1 = a1, w2 = a2, w3 = a3, w4 = a4, w5 = a5, w6 = a6, w7 = a7;
    u32 m = seed | 1u;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (OP == 0) {        // plain mad.wide.u32, 8 independent 64-bit accumulators; multiplicands vary so nothing folds
                asm volatile("mad.wide.u32 %0, %8, %9, %0; mad.wide.u32 %1, %10, %9, %1; mad.wide.u32 %2, %11, %9, %2; mad.wide.u32 %3, %12, %9, %3;"
                             "mad.wide.u32 %4, %13, %9, %4; mad.wide.u32 %5, %14, %9, %5; mad.wide.u32 %6, %15, %9, %6; mad.wide.u32 %7, %16, %9, %7;"
                             : "+l"(w0), "+l"(w1), "+l"(w2), "+l"(w3), "+l"(w4), "+l"(w5), "+l"(w6), "+l"(w7)
                             : "r"(a0), "r"(m), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5), "r"(a6), "r"(a7));
                m += (u32)w0;     // data dependence on the full-width result (one IADD per 8 mads)
            } else if (OP == 1) { // carry chain (IMAD.WIDE.U32.X pairs)
                asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1; madc.lo.cc.u32 %2, %10, %9, %2; madc.hi.cc.u32 %3, %10, %9, %3;"
                             "madc.lo.cc.u32 %4, %11, %9, %4; madc.hi.cc.u32 %5, %11, %9, %5; madc.lo.cc.u32 %6, %12, %9, %6; madc.hi.u32 %7, %12, %9, %7;"
                             : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(m), "r"(seed), "r"(seed + 1), "r"(seed + 2), "r"(seed + 3));
            } else if (OP == 2) { // mad.hi.u32 alone
                asm volatile("mad.hi.u32 %0, %0, %8, %0; mad.hi.u32 %1, %1, %8, %1; mad.hi.u32 %2, %2, %8, %2; mad.hi.u32 %3, %3, %8, %3;"
                             "mad.hi.u32 %4, %4, %8, %4; mad.hi.u32 %5, %5, %8, %5; mad.hi.u32 %6, %6, %8, %6; mad.hi.u32 %7, %7, %8, %7;"
                             : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(m));
            } else if (OP == 3) { // plain wide mads interleaved 1:1 with 3-input adds (co-issue with the alu pipe)
                asm volatile("mad.wide.u32 %0, %8, %9, %0; mad.wide.u32 %1, %10, %9, %1; mad.wide.u32 %2, %11, %9, %2; mad.wide.u32 %3, %12, %9, %3;"
                             "add.u32 %10, %10, %9; add.u32 %11, %11, %9; add.u32 %12, %12, %9; add.u32 %13, %13, %9;"
                             "mad.wide.u32 %4, %13, %9, %4; mad.wide.u32 %5, %14, %9, %5; mad.wide.u32 %6, %15, %9, %6; mad.wide.u32 %7, %16, %9, %7;"
                             "add.u32 %14, %14, %9; add.u32 %15, %15, %9; add.u32 %16, %16, %9; add.u32 %8, %8, %9;"
                             : "+l"(w0), "+l"(w1), "+l"(w2), "+l"(w3), "+l"(w4), "+l"(w5), "+l"(w6), "+l"(w7),
                               "+r"(a0), "+r"(m), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7));
            } else if (OP == 4) { // two independent carry chains interleaved (as the E/O rows of fe_mul)
                asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1; madc.lo.cc.u32 %2, %10, %9, %2; madc.hi.u32 %3, %10, %9, %3;"
                             "mad.lo.cc.u32 %4, %11, %9, %4; madc.hi.cc.u32 %5, %11, %9, %5; madc.lo.cc.u32 %6, %12, %9, %6; madc.hi.u32 %7, %12, %9, %7;"
                             : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(m), "r"(seed), "r"(seed + 1), "r"(seed + 2), "r"(seed + 3));
            } else if (OP == 5) { // wide mads without carry-in but WITH carry-out is not expressible; mul.wide (no accumulate)
                asm volatile("mul.wide.u32 %0, %8, %9; mul.wide.u32 %1, %10, %9; mul.wide.u32 %2, %11, %9; mul.wide.u32 %3, %12, %9;"
                             "mul.wide.u32 %4, %13, %9; mul.wide.u32 %5, %14, %9; mul.wide.u32 %6, %15, %9; mul.wide.u32 %7, %16, %9;"
                             : "=l"(w0), "=l"(w1), "=l"(w2), "=l"(w3), "=l"(w4), "=l"(w5), "=l"(w6), "=l"(w7)
                             : "r"(a0), "r"(m), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5), "r"(a6), "r"(a7));
                a0 += (u32)(w0 >> 32); a1 += (u32)w1; a2 += (u32)(w2 >> 32); a3 += (u32)w3; a4 += (u32)(w4 >> 32); a5 += (u32)w5; a6 += (u32)(w6 >> 32); a7 += (u32)w7;
            }
        }
    }
    u64 x = w0 ^ w1 ^ w2 ^ w3 ^ w4 ^ w5 ^ w6 ^ w7;
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7 ^ m ^ (u32)x ^ (u32)(x >> 32);
}

template <int OP>
static double rate(u32* out, int sms, int ctas_per_sm, int clk_khz, double ops_per_iter) {
    int iters = 2048, blocks = sms * ctas_per_sm;
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, 256>>>(out, 12345u, iters); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) { cudaEventRecord(e0); k<OP><<<blocks, 256>>>(out, 12345u, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    double ops = (double)blocks * 256 * iters * ops_per_iter;
    return ops / (best * 1e-3) / sms / (clk_khz * 1e3);
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount; int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    u32* out; cudaMalloc(&out, 64 << 20);
    printf("{\"gpu\": \"%s\", \"unit\": \"warp-lane ops per clk per SM at max clock\"", prop.name);
    for (int occ : {8, 2, 1}) {
        printf(", \"mad_wide_plain_%dcta\": %.2f", occ, rate<0>(out, sms, occ, clk_khz, 64));
        printf(", \"mad_wide_carry_chain8_%dcta\": %.2f", occ, rate<1>(out, sms, occ, clk_khz, 32));
        printf(", \"mad_hi_%dcta\": %.2f", occ, rate<2>(out, sms, occ, clk_khz, 64));
        printf(", \"mad_wide_plain_with_equal_adds_%dcta\": %.2f", occ, rate<3>(out, sms, occ, clk_khz, 64));
        printf(", \"mad_wide_two_chains4_%dcta\": %.2f", occ, rate<4>(out, sms, occ, clk_khz, 32));
        printf(", \"mul_wide_%dcta\": %.2f", occ, rate<5>(out, sms, occ, clk_khz, 64));
    }
    printf("}\n");
    return 0;
}
