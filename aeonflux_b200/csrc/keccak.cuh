// Keccak-f[1600] and the block-template form of a Merlin/STROBE-128 transcript.
//
// Replaces merlin's Transcript as driven by the zkp toolbox (call sites
// /root/reference/src/nizk/presentation.rs:355-435, encryption.rs:160-209, issuance.rs:142-217).  All STROBE
// framing (labels, lengths, begin_op headers, the run_f padding bytes) is static per proof shape, so the host compiles
// a transcript into a list of 168-byte XOR masks ("blocks", one Keccak-f each) with 32-byte holes for the per-item
// point encodings; the per-issuer constant prefix is folded into a precomputed midstate.  The device just XORs,
// fills holes and permutes, then reads the 64 challenge bytes from the head of the state.
#pragma once
#include "fe.cuh"

namespace afx {

AFX_HD u64 rotl64(u64 x, int n) { return (x << n) | (x >> (64 - n)); }

AFX_HD void keccak_f1600(u64* A) {
    const u64 RC[24] = {0x0000000000000001ull, 0x0000000000008082ull, 0x800000000000808aull, 0x8000000080008000ull, 0x000000000000808bull,
                        0x0000000080000001ull, 0x8000000080008081ull, 0x8000000000008009ull, 0x000000000000008aull, 0x0000000000000088ull,
                        0x0000000080008009ull, 0x000000008000000aull, 0x000000008000808bull, 0x800000000000008bull, 0x8000000000008089ull,
                        0x8000000000008003ull, 0x8000000000008002ull, 0x8000000000000080ull, 0x000000000000800aull, 0x800000008000000aull,
                        0x8000000080008081ull, 0x8000000000008080ull, 0x0000000080000001ull, 0x8000000080008008ull};
#if defined(__CUDA_ARCH__)
#pragma unroll 1
#endif
    for (int r = 0; r < 24; r++) {
        u64 C0 = A[0] ^ A[5] ^ A[10] ^ A[15] ^ A[20], C1 = A[1] ^ A[6] ^ A[11] ^ A[16] ^ A[21], C2 = A[2] ^ A[7] ^ A[12] ^ A[17] ^ A[22];
        u64 C3 = A[3] ^ A[8] ^ A[13] ^ A[18] ^ A[23], C4 = A[4] ^ A[9] ^ A[14] ^ A[19] ^ A[24];
        u64 D0 = C4 ^ rotl64(C1, 1), D1 = C0 ^ rotl64(C2, 1), D2 = C1 ^ rotl64(C3, 1), D3 = C2 ^ rotl64(C4, 1), D4 = C3 ^ rotl64(C0, 1);
        // theta + rho + pi: B[y + 5*((2x+3y)%5)] = rotl(A[x+5y] ^ D[x], rho[x+5y])
        u64 B0 = A[0] ^ D0;
        u64 B10 = rotl64(A[1] ^ D1, 1), B20 = rotl64(A[2] ^ D2, 62), B5 = rotl64(A[3] ^ D3, 28), B15 = rotl64(A[4] ^ D4, 27);
        u64 B16 = rotl64(A[5] ^ D0, 36), B1 = rotl64(A[6] ^ D1, 44), B11 = rotl64(A[7] ^ D2, 6), B21 = rotl64(A[8] ^ D3, 55), B6 = rotl64(A[9] ^ D4, 20);
        u64 B7 = rotl64(A[10] ^ D0, 3), B17 = rotl64(A[11] ^ D1, 10), B2 = rotl64(A[12] ^ D2, 43), B12 = rotl64(A[13] ^ D3, 25), B22 = rotl64(A[14] ^ D4, 39);
        u64 B23 = rotl64(A[15] ^ D0, 41), B8 = rotl64(A[16] ^ D1, 45), B18 = rotl64(A[17] ^ D2, 15), B3 = rotl64(A[18] ^ D3, 21), B13 = rotl64(A[19] ^ D4, 8);
        u64 B14 = rotl64(A[20] ^ D0, 18), B24 = rotl64(A[21] ^ D1, 2), B9 = rotl64(A[22] ^ D2, 61), B19 = rotl64(A[23] ^ D3, 56), B4 = rotl64(A[24] ^ D4, 14);
        // chi + iota
        A[0] = B0 ^ (~B1 & B2) ^ RC[r]; A[1] = B1 ^ (~B2 & B3); A[2] = B2 ^ (~B3 & B4); A[3] = B3 ^ (~B4 & B0); A[4] = B4 ^ (~B0 & B1);
        A[5] = B5 ^ (~B6 & B7); A[6] = B6 ^ (~B7 & B8); A[7] = B7 ^ (~B8 & B9); A[8] = B8 ^ (~B9 & B5); A[9] = B9 ^ (~B5 & B6);
        A[10] = B10 ^ (~B11 & B12); A[11] = B11 ^ (~B12 & B13); A[12] = B12 ^ (~B13 & B14); A[13] = B13 ^ (~B14 & B10); A[14] = B14 ^ (~B10 & B11);
        A[15] = B15 ^ (~B16 & B17); A[16] = B16 ^ (~B17 & B18); A[17] = B17 ^ (~B18 & B19); A[18] = B18 ^ (~B19 & B15); A[19] = B19 ^ (~B15 & B16);
        A[20] = B20 ^ (~B21 & B22); A[21] = B21 ^ (~B22 & B23); A[22] = B22 ^ (~B23 & B24); A[23] = B23 ^ (~B24 & B20); A[24] = B24 ^ (~B20 & B21);
    }
}

}  // namespace afx
