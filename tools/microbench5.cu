// How fast can the ladder's window body run with nothing around it?  One "window" = 4 doublings + 2 additions of a
// projective-Niels entry + 1 mixed addition, all in registers (no table loads, no digit extraction, no barrier, no aMAC
// CTAs), in the completed-coordinates forms k_ladders uses; 256-thread CTAs, 2 per SM, like k_ladders.  Reports windows/s and
// the share of the wide-product rate (4 x 392 + 2 x 576 + 504 = 3,224 products per window; 9.12 T products/s measured peak).
#include <cuda_runtime.h>
#include <cstdio>
#include <vector>
#include "../aeonflux_b200/csrc/ge.cuh"
using namespace afx;

template <int SYNC>
__global__ void __launch_bounds__(256, 2) k_window(u32* out, const u32* in, int iters) {
    u32 tid = blockIdx.x * blockDim.x + threadIdx.x;
    pniels n; aniels m; gc acc = gc_identity();
    for (int i = 0; i < 8; i++) {
        n.YpX.v[i] = in[(tid * 8 + i) & 1023]; n.YmX.v[i] = in[(tid * 8 + i + 8) & 1023]; n.Z.v[i] = in[(tid * 8 + i + 16) & 1023]; n.T2d.v[i] = in[(tid * 8 + i + 24) & 1023];
        m.ypx.v[i] = in[(tid * 8 + i + 32) & 1023]; m.ymx.v[i] = in[(tid * 8 + i + 40) & 1023]; m.xy2d.v[i] = in[(tid * 8 + i + 48) & 1023];
        acc.E.v[i] = in[(tid * 8 + i + 56) & 1023]; acc.F.v[i] = in[(tid * 8 + i + 64) & 1023];
    }
#pragma unroll 1
    for (int w = 0; w < iters; w++) {
        if (SYNC) __syncthreads();
        gc_dbl4(acc);
#pragma unroll 1
        for (int k = 0; k < 2; k++) { pniels e = pniels_cneg(n, (u32)(w + k) & 1u); GE_LADDER_ADD(acc, e); }
        aniels f = aniels_cneg(m, (u32)w & 1u);
        GE_LADDER_MADD(acc, f);
    }
    ge r = gc_to_ge(acc);
    for (int i = 0; i < 8; i++) out[8 * tid + i] = r.X.v[i] ^ r.Y.v[i] ^ r.Z.v[i] ^ r.T.v[i];
}

int main() {
    cudaDeviceProp prop; cudaGetDeviceProperties(&prop, 0);
    int sms = prop.multiProcessorCount;
    u32 *out, *in; cudaMalloc(&out, 64 << 20); cudaMalloc(&in, 4096);
    std::vector<u32> h(1024); for (int i = 0; i < 1024; i++) h[i] = 0x9e3779b9u * (i + 1) ^ (i * 0x85ebca6bu);
    cudaMemcpy(in, h.data(), 4096, cudaMemcpyHostToDevice);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const int iters = 512;
    const double products = 4 * 392 + 2 * 576 + 504;
    printf("{\"gpu\": \"%s\", \"products_per_window\": %.0f", prop.name, products);
    for (int sync = 0; sync < 2; sync++) {
        for (int waves : {1, 4}) {
            float best = 1e30f;
            for (int rep = 0; rep < 4; rep++) {
                cudaEventRecord(e0);
                if (sync) k_window<1><<<sms * 2 * waves, 256>>>(out, in, iters); else k_window<0><<<sms * 2 * waves, 256>>>(out, in, iters);
                cudaEventRecord(e1); cudaEventSynchronize(e1);
                float ms; cudaEventElapsedTime(&ms, e0, e1); if (rep && ms < best) best = ms;
            }
            double wps = (double)sms * 2 * waves * 256 * iters / (best * 1e-3);
            printf(", \"%s_%dwave_Gwindows_per_s\": %.3f, \"%s_%dwave_frac_of_wide_product_peak\": %.3f", sync ? "barrier" : "free", waves, wps / 1e9,
                   sync ? "barrier" : "free", waves, wps * products / 9.123e12);
        }
    }
    printf("}\n");
    return 0;
}
