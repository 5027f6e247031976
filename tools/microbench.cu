// Integer-pipe microbenchmark for B200 (sm_100a): issue rates of IMAD, IMAD.WIDE.U32, IADD3 and the throughput of the
// engine's field multiply / square.  Fixes the roofline denominator ("IMAD peak").  Prints one JSON object.
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include "../aeonflux_b200/csrc/fe.cuh"
using namespace afx;

template <int OP>
__global__ void __launch_bounds__(256) k_rate(u32* out, u32 seed, int iters) {
    u32 a0 = seed + threadIdx.x, a1 = a0 * 3 + 1, a2 = a0 * 5 + 2, a3 = a0 * 7 + 3, a4 = a0 * 11 + 4, a5 = a0 * 13 + 5, a6 = a0 * 17 + 6, a7 = a0 * 19 + 7;
    u64 w0 = a0, w1 = a1, w2 = a2, w3 = a3, w4 = a4, w5 = a5, w6 = a6, w7 = a7;
    u32 m = seed | 1u;
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++) {
            if (OP == 0) {  // IMAD (32-bit low multiply-add)
                asm volatile("mad.lo.u32 %0, %0, %8, %0; mad.lo.u32 %1, %1, %8, %1; mad.lo.u32 %2, %2, %8, %2; mad.lo.u32 %3, %3, %8, %3;"
                             "mad.lo.u32 %4, %4, %8, %4; mad.lo.u32 %5, %5, %8, %5; mad.lo.u32 %6, %6, %8, %6; mad.lo.u32 %7, %7, %8, %7;"
                             : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(m));
            } else if (OP == 1) {  // IMAD.WIDE.U32 (32x32+64)
                asm volatile("mad.wide.u32 %0, %8, %9, %0; mad.wide.u32 %1, %10, %9, %1; mad.wide.u32 %2, %11, %9, %2; mad.wide.u32 %3, %12, %9, %3;"
                             "mad.wide.u32 %4, %13, %9, %4; mad.wide.u32 %5, %14, %9, %5; mad.wide.u32 %6, %15, %9, %6; mad.wide.u32 %7, %16, %9, %7;"
                             : "+l"(w0), "+l"(w1), "+l"(w2), "+l"(w3), "+l"(w4), "+l"(w5), "+l"(w6), "+l"(w7)
                             : "r"(a0), "r"(m), "r"(a1), "r"(a2), "r"(a3), "r"(a4), "r"(a5), "r"(a6), "r"(a7));
            } else if (OP == 2) {  // IADD3
                asm volatile("add.u32 %0, %0, %8; add.u32 %1, %1, %8; add.u32 %2, %2, %8; add.u32 %3, %3, %8;"
                             "add.u32 %4, %4, %8; add.u32 %5, %5, %8; add.u32 %6, %6, %8; add.u32 %7, %7, %8;"
                             : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(m));
            } else {  // carry chain: IMAD.WIDE.U32.X pairs as the field multiply uses them
                asm volatile("mad.lo.cc.u32 %0, %8, %9, %0; madc.hi.cc.u32 %1, %8, %9, %1; madc.lo.cc.u32 %2, %8, %9, %2; madc.hi.cc.u32 %3, %8, %9, %3;"
                             "madc.lo.cc.u32 %4, %8, %9, %4; madc.hi.cc.u32 %5, %8, %9, %5; madc.lo.cc.u32 %6, %8, %9, %6; madc.hi.u32 %7, %8, %9, %7;"
                             : "+r"(a0), "+r"(a1), "+r"(a2), "+r"(a3), "+r"(a4), "+r"(a5), "+r"(a6), "+r"(a7) : "r"(m), "r"(seed));
            }
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 ^ a1 ^ a2 ^ a3 ^ a4 ^ a5 ^ a6 ^ a7 ^ (u32)(w0 ^ w1 ^ w2 ^ w3 ^ w4 ^ w5 ^ w6 ^ w7);
}

template <int OP>
__global__ void __launch_bounds__(128) k_fe(u32* out, const u32* in, int iters) {
    u32 t = blockIdx.x * blockDim.x + threadIdx.x;
    fe a, b;
    for (int i = 0; i < 8; i++) { a.v[i] = in[(16 * t + i) & 1023]; b.v[i] = in[(16 * t + 8 + i) & 1023]; }
    for (int i = 0; i < iters; i++) {
        if (OP == 0) { fe c = fe_mul(a, b); a = b; b = c; }
        else { a = fe_sq(a); }
    }
    for (int i = 0; i < 8; i++) out[8 * t + i] = a.v[i] ^ b.v[i];
}

static float time_ms(void (*launch)(void*), void* arg) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    launch(arg); cudaDeviceSynchronize();
    float best = 1e30f;
    for (int r = 0; r < 5; r++) { cudaEventRecord(e0); launch(arg); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best;
}
struct Args { u32* out; u32* in; int iters; int blocks; int which; };
static void launch_rate(void* p) {
    Args* a = (Args*)p;
    switch (a->which) {
        case 0: k_rate<0><<<a->blocks, 256>>>(a->out, 12345u, a->iters); break;
        case 1: k_rate<1><<<a->blocks, 256>>>(a->out, 12345u, a->iters); break;
        case 2: k_rate<2><<<a->blocks, 256>>>(a->out, 12345u, a->iters); break;
        case 3: k_rate<3><<<a->blocks, 256>>>(a->out, 12345u, a->iters); break;
        case 10: k_fe<0><<<a->blocks, 128>>>(a->out, a->in, a->iters); break;
        case 11: k_fe<1><<<a->blocks, 128>>>(a->out, a->in, a->iters); break;
    }
}
int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount; int clk_khz = 0; cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0);
    u32 *out, *in; cudaMalloc(&out, 64 << 20); cudaMalloc(&in, 4096);
    std::vector<u32> h(1024); for (int i = 0; i < 1024; i++) h[i] = 0x9e3779b9u * (i + 1);
    cudaMemcpy(in, h.data(), 4096, cudaMemcpyHostToDevice);
    printf("{\"gpu\": \"%s\", \"sms\": %d, \"max_clock_mhz\": %.0f", prop.name, sms, clk_khz / 1000.0);
    const char* names[4] = {"imad", "imad_wide_u32", "iadd3", "imad_wide_carry_chain"};
    for (int w = 0; w < 4; w++) {
        Args a{out, in, 2048, sms * 8, w};   // 8 CTAs x 256 threads per SM = full occupancy
        float ms = time_ms(launch_rate, &a);
        double ops = (double)a.blocks * 256 * a.iters * 64;   // 8 unrolled x 8 ops
        if (w == 3) ops /= 2;                                  // count fused lo/hi pairs as one wide op
        printf(", \"%s_Tops\": %.3f, \"%s_per_clk_per_sm_at_max_clock\": %.2f", names[w], ops / ms / 1e9, names[w], ops / (ms * 1e-3) / sms / (clk_khz * 1e3));
    }
    for (int occ = 1; occ <= 4; occ++) {
        for (int w = 10; w <= 11; w++) {
            Args a{out, in, 4096, sms * occ, w};
            float ms = time_ms(launch_rate, &a);
            double muls = (double)a.blocks * 128 * a.iters;
            printf(", \"%s_G_per_s_%dcta\": %.2f", w == 10 ? "fe_mul" : "fe_sq", occ, muls / ms / 1e6);
        }
    }
    for (int w = 10; w <= 11; w++) {
        Args a{out, in, 4096, sms * 16, w};
        float ms = time_ms(launch_rate, &a);
        printf(", \"%s_G_per_s_16cta\": %.2f", w == 10 ? "fe_mul" : "fe_sq", (double)a.blocks * 128 * a.iters / ms / 1e6);
    }
    printf("}\n");
    return 0;
}
