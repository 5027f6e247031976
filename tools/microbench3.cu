// Third pipe microbenchmark (B200, sm_100a): FP64 DFMA issue rate, and whether DFMA overlaps with the IMAD.WIDE carry chains the
// field arithmetic lives on (separate pipes?).  Input for a possible hybrid integer / double-precision limb product next round.
#include <cuda_runtime.h>
#include <cstdio>
typedef unsigned int u32;

template <int OP>
__global__ void __launch_bounds__(256) k(double* out, u32* iout, u32 seed, int iters) {
    double d0 = seed + threadIdx.x, d1 = d0 * 1.5, d2 = d0 * 2.5, d3 = d0 * 3.5, d4 = d0 * 4.5, d5 = d0 * 5.5, d6 = d0 * 6.5, d7 = d0 * 7.5;
    double m = 1.0000001 + 1e-9 * seed, c = 0.5;
    u32 a0 = seed + threadIdx.x, a1 = a0 * 3 + 1, a2 = a0 * 5 + 2, a3 = a0 * 7 + 3, a4 = a0 * 11 + 4, a5 = a0 * 13 + 5, a6 = a0 * 17 + 6, a7 = a0 * 19 + 7;
    u32 mm = seed | 1u;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (OP == 0 || OP == 2) {   // 8 independent DFMA chains
                asm volatile("fma.rn.f64 %0, %0, %8, %9; fma.rn.f64 %1, %1, %8, %9; fma.rn.f64 %2, %2, %8, %9; fma.rn.f64 %3, %3, %8, %9;"
                             "fma.rn.f64 %4, %4, %8, %9; fma.rn.f64 %5, %5, %8, %9; fma.rn.f64 %6, %6, %8, %9; fma.rn.f64 %7, %7, %8, %9;"
                             : "+d"(d0), "+d"(d1), "+d"(d2), "+d"(d3), "+d"(d4), "+d"(d5), "+d"(d6), "+d"(d7) : "d"(m), "d"(c));
            }
            if (OP == 1 || OP == 2) {   // the carry-chain wide products of fe_mul (4 fused IMAD.WIDE.U32.X per block)
                asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1; madc.lo.cc.u32 %2, %10, %9, %2; madc.hi.cc.u32 %3, %10, %9, %3;"
                             "madc.lo.cc.u32 %4, %11, %9, %4; madc.hi.cc.u32 %5, %11, %9, %5; madc.lo.cc.u32 %6, %12, %9, %6; madc.hi.u32 %7, %12, %9, %7;"
                             : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(mm), "r"(seed), "r"(seed + 1), "r"(seed + 2), "r"(seed + 3));
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = d0 + d1 + d2 + d3 + d4 + d5 + d6 + d7;
    iout[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7;
}
template <int OP>
static float run(double* out, u32* iout, int blocks, int iters) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<OP><<<blocks, 256>>>(out, iout, 12345u, iters); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) { cudaEventRecord(e0); k<OP><<<blocks, 256>>>(out, iout, 12345u, iters); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best;
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount, clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    double* out; u32* iout; cudaMalloc(&out, 64 << 20); cudaMalloc(&iout, 32 << 20);
    int iters = 2048, blocks = sms * 8;
    double lanes = (double)blocks * 256 * iters;
    float t_d = run<0>(out, iout, blocks, iters), t_i = run<1>(out, iout, blocks, iters), t_b = run<2>(out, iout, blocks, iters);
    double per = sms * (clk_khz * 1e3);
    printf("{\"gpu\": \"%s\", \"dfma_per_clk_per_sm\": %.2f, \"imad_wide_chain_per_clk_per_sm\": %.2f, "
           "\"both_interleaved_ms\": %.3f, \"dfma_alone_ms\": %.3f, \"imad_wide_alone_ms\": %.3f, \"overlap\": \"%s\"}\n",
           prop.name, lanes * 64 / (t_d * 1e-3) / per, lanes * 32 / (t_i * 1e-3) / per, t_b, t_d, t_i,
           t_b < 0.8f * (t_d + t_i) ? "the two pipes overlap" : "no useful overlap");
    return 0;
}
