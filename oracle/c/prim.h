/* ORACLE (test infrastructure, NOT the product).
 *
 * CPU restatement of the third-party layers under aeonflux's hot path, following the
 * *reference's own CPU schedule* so it can double as the reported CPU baseline:
 *   - GF(2^255-19) in 5x51-bit limbs with unsigned __int128 products, like
 *     curve25519-dalek's default u64_backend (/root/reference/Cargo.toml:34,48);
 *   - ristretto255 compress / decompress / Elligator, extended Edwards
 *     add/double through projective-Niels and completed points;
 *   - constant-time radix-16 variable-base scalar mult and multiscalar mult, and the
 *     vartime Straus NAF-5 multiscalar mult (dalek's choices below 190 points);
 *   - scalars mod l (Barrett on 64-bit limbs);
 *   - Keccak-f[1600], STROBE-128, Merlin transcripts; SHA-512; SHAKE-256.
 * The algorithms are the published ones; nothing here is copied from dalek/merlin/zkp
 * (they are not even present in /root/reference).  Validated against the big-int Python
 * oracle and libsodium in tests/test_oracle_c.py.
 */
#pragma once
#include <stdint.h>
#include <string.h>
#include "consts.h"

typedef unsigned __int128 u128;
typedef uint64_t fe[5];
#define M51 ((1ULL << 51) - 1)

/* ------------------------------------------------------------------ field */
static inline void fe_copy(fe h, const fe f) { memcpy(h, f, sizeof(fe)); }
static inline void fe_0(fe h) { memset(h, 0, sizeof(fe)); }
static inline void fe_1(fe h) { fe_0(h); h[0] = 1; }

static inline void fe_weak_reduce(fe h) {
    uint64_t c;
    c = h[0] >> 51; h[0] &= M51; h[1] += c;
    c = h[1] >> 51; h[1] &= M51; h[2] += c;
    c = h[2] >> 51; h[2] &= M51; h[3] += c;
    c = h[3] >> 51; h[3] &= M51; h[4] += c;
    c = h[4] >> 51; h[4] &= M51; h[0] += 19 * c;
}

static inline uint64_t load64(const uint8_t* p) { uint64_t v; memcpy(&v, p, 8); return v; }
static inline void store64(uint8_t* p, uint64_t v) { memcpy(p, &v, 8); }

static inline void fe_frombytes(fe h, const uint8_t s[32]) { /* bit 255 ignored */
    uint64_t w0 = load64(s), w1 = load64(s + 8), w2 = load64(s + 16), w3 = load64(s + 24);
    h[0] = w0 & M51;
    h[1] = ((w0 >> 51) | (w1 << 13)) & M51;
    h[2] = ((w1 >> 38) | (w2 << 26)) & M51;
    h[3] = ((w2 >> 25) | (w3 << 39)) & M51;
    h[4] = (w3 >> 12) & M51;
}

static inline void fe_tobytes(uint8_t s[32], const fe f) {
    fe h; fe_copy(h, f);
    fe_weak_reduce(h); fe_weak_reduce(h);
    uint64_t q = (h[0] + 19) >> 51;
    q = (h[1] + q) >> 51; q = (h[2] + q) >> 51; q = (h[3] + q) >> 51; q = (h[4] + q) >> 51;
    h[0] += 19 * q;
    uint64_t c;
    c = h[0] >> 51; h[0] &= M51; h[1] += c;
    c = h[1] >> 51; h[1] &= M51; h[2] += c;
    c = h[2] >> 51; h[2] &= M51; h[3] += c;
    c = h[3] >> 51; h[3] &= M51; h[4] += c;
    h[4] &= M51;
    store64(s, h[0] | (h[1] << 51));
    store64(s + 8, (h[1] >> 13) | (h[2] << 38));
    store64(s + 16, (h[2] >> 26) | (h[3] << 25));
    store64(s + 24, (h[3] >> 39) | (h[4] << 12));
}

static inline void fe_add(fe h, const fe f, const fe g) { for (int i = 0; i < 5; i++) h[i] = f[i] + g[i]; }
static inline void fe_sub(fe h, const fe f, const fe g) { /* f + 16p - g, then carry */
    h[0] = f[0] + 36028797018963664ULL - g[0];
    for (int i = 1; i < 5; i++) h[i] = f[i] + 36028797018963952ULL - g[i];
    fe_weak_reduce(h);
}
static inline void fe_neg(fe h, const fe f) { fe z; fe_0(z); fe_sub(h, z, f); }

static inline void fe_mul(fe h, const fe f, const fe g) {
    uint64_t f0 = f[0], f1 = f[1], f2 = f[2], f3 = f[3], f4 = f[4];
    uint64_t g0 = g[0], g1 = g[1], g2 = g[2], g3 = g[3], g4 = g[4];
    uint64_t g1_19 = 19 * g1, g2_19 = 19 * g2, g3_19 = 19 * g3, g4_19 = 19 * g4;
    u128 c0 = (u128)f0 * g0 + (u128)f4 * g1_19 + (u128)f3 * g2_19 + (u128)f2 * g3_19 + (u128)f1 * g4_19;
    u128 c1 = (u128)f1 * g0 + (u128)f0 * g1 + (u128)f4 * g2_19 + (u128)f3 * g3_19 + (u128)f2 * g4_19;
    u128 c2 = (u128)f2 * g0 + (u128)f1 * g1 + (u128)f0 * g2 + (u128)f4 * g3_19 + (u128)f3 * g4_19;
    u128 c3 = (u128)f3 * g0 + (u128)f2 * g1 + (u128)f1 * g2 + (u128)f0 * g3 + (u128)f4 * g4_19;
    u128 c4 = (u128)f4 * g0 + (u128)f3 * g1 + (u128)f2 * g2 + (u128)f1 * g3 + (u128)f0 * g4;
    c1 += (uint64_t)(c0 >> 51); h[0] = (uint64_t)c0 & M51;
    c2 += (uint64_t)(c1 >> 51); h[1] = (uint64_t)c1 & M51;
    c3 += (uint64_t)(c2 >> 51); h[2] = (uint64_t)c2 & M51;
    c4 += (uint64_t)(c3 >> 51); h[3] = (uint64_t)c3 & M51;
    uint64_t carry = (uint64_t)(c4 >> 51); h[4] = (uint64_t)c4 & M51;
    h[0] += carry * 19;
    h[1] += h[0] >> 51; h[0] &= M51;
}

static inline void fe_sq(fe h, const fe f) {
    uint64_t f0 = f[0], f1 = f[1], f2 = f[2], f3 = f[3], f4 = f[4];
    uint64_t f0_2 = 2 * f0, f1_2 = 2 * f1, f3_19 = 19 * f3, f4_19 = 19 * f4;
    u128 c0 = (u128)f0 * f0 + (u128)f1_2 * f4_19 + (u128)(2 * f2) * f3_19;
    u128 c1 = (u128)f0_2 * f1 + (u128)(2 * f2) * f4_19 + (u128)f3 * f3_19;
    u128 c2 = (u128)f0_2 * f2 + (u128)f1 * f1 + (u128)(2 * f3) * f4_19;
    u128 c3 = (u128)f0_2 * f3 + (u128)f1_2 * f2 + (u128)f4 * f4_19;
    u128 c4 = (u128)f0_2 * f4 + (u128)f1_2 * f3 + (u128)f2 * f2;
    c1 += (uint64_t)(c0 >> 51); h[0] = (uint64_t)c0 & M51;
    c2 += (uint64_t)(c1 >> 51); h[1] = (uint64_t)c1 & M51;
    c3 += (uint64_t)(c2 >> 51); h[2] = (uint64_t)c2 & M51;
    c4 += (uint64_t)(c3 >> 51); h[3] = (uint64_t)c3 & M51;
    uint64_t carry = (uint64_t)(c4 >> 51); h[4] = (uint64_t)c4 & M51;
    h[0] += carry * 19;
    h[1] += h[0] >> 51; h[0] &= M51;
}

static inline void fe_sqn(fe h, const fe f, int n) { fe_sq(h, f); for (int i = 1; i < n; i++) fe_sq(h, h); }

/* x^(2^250-1) and x^11 -- the shared prefix of inversion and pow_p58 */
static inline void fe_pow22501(fe t19, fe t3, const fe x) {
    fe t0, t1, t2, t4, t5, t6, t7, t8, t9, t10, t11, t12, t13, t14, t15, t16, t17, t18;
    fe_sq(t0, x); fe_sqn(t1, t0, 2); fe_mul(t2, x, t1); fe_mul(t3, t0, t2);
    fe_sq(t4, t3); fe_mul(t5, t2, t4);
    fe_sqn(t6, t5, 5); fe_mul(t7, t6, t5);
    fe_sqn(t8, t7, 10); fe_mul(t9, t8, t7);
    fe_sqn(t10, t9, 20); fe_mul(t11, t10, t9);
    fe_sqn(t12, t11, 10); fe_mul(t13, t12, t7);
    fe_sqn(t14, t13, 50); fe_mul(t15, t14, t13);
    fe_sqn(t16, t15, 100); fe_mul(t17, t16, t15);
    fe_sqn(t18, t17, 50); fe_mul(t19, t18, t13);
}
static inline void fe_pow_p58(fe h, const fe x) { /* x^((p-5)/8) = x^(2^252-3) */
    fe t19, t3, t20; fe_pow22501(t19, t3, x); fe_sqn(t20, t19, 2); fe_mul(h, x, t20);
}

static inline int fe_is_negative(const fe f) { uint8_t s[32]; fe_tobytes(s, f); return s[0] & 1; }
static inline int fe_equal(const fe f, const fe g) { uint8_t a[32], b[32]; fe_tobytes(a, f); fe_tobytes(b, g); return memcmp(a, b, 32) == 0; }
static inline int fe_is_zero(const fe f) { uint8_t a[32]; static const uint8_t z[32] = {0}; fe_tobytes(a, f); return memcmp(a, z, 32) == 0; }
static inline void fe_cneg(fe h, int b) { if (b) { fe t; fe_neg(t, h); fe_copy(h, t); } }
static inline void fe_abs(fe h) { fe_cneg(h, fe_is_negative(h)); }

/* sqrt_ratio_i(u, v) -- SURVEY A.1.  returns was_square; r = non-negative root */
static inline int fe_sqrt_ratio_i(fe r, const fe u, const fe v) {
    fe v3, v7, t, check, mu, mui;
    fe_sq(t, v); fe_mul(v3, t, v);
    fe_sq(t, v3); fe_mul(v7, t, v);
    fe_mul(t, u, v7); fe_pow_p58(t, t);
    fe_mul(r, u, v3); fe_mul(r, r, t);
    fe_sq(t, r); fe_mul(check, v, t);
    fe_neg(mu, u); fe_mul(mui, mu, FE_SQRT_M1);
    int correct = fe_equal(check, u), flipped = fe_equal(check, mu), flipped_i = fe_equal(check, mui);
    if (flipped | flipped_i) { fe_mul(t, r, FE_SQRT_M1); fe_copy(r, t); }
    fe_abs(r);
    return correct | flipped;
}
static inline int fe_invsqrt(fe r, const fe v) { fe one; fe_1(one); return fe_sqrt_ratio_i(r, one, v); }

/* ------------------------------------------------------------------ group */
typedef struct { fe X, Y, Z, T; } ge;        /* extended */
typedef struct { fe YpX, YmX, Z, T2d; } pniels; /* projective Niels */
typedef struct { fe X, Y, Z, T; } completed;

static inline void ge_identity(ge* p) { fe_0(p->X); fe_1(p->Y); fe_1(p->Z); fe_0(p->T); }
static inline void ge_neg(ge* r, const ge* p) { fe_neg(r->X, p->X); fe_copy(r->Y, p->Y); fe_copy(r->Z, p->Z); fe_neg(r->T, p->T); }
static inline void ge_to_pniels(pniels* n, const ge* p) {
    fe_add(n->YpX, p->Y, p->X); fe_sub(n->YmX, p->Y, p->X); fe_copy(n->Z, p->Z); fe_mul(n->T2d, p->T, FE_D2);
}
static inline void pniels_identity(pniels* n) { fe_1(n->YpX); fe_1(n->YmX); fe_1(n->Z); fe_0(n->T2d); }
static inline void pniels_neg(pniels* r, const pniels* n) { fe_copy(r->YpX, n->YmX); fe_copy(r->YmX, n->YpX); fe_copy(r->Z, n->Z); fe_neg(r->T2d, n->T2d); }
static inline void completed_to_ge(ge* r, const completed* c) {
    fe_mul(r->X, c->X, c->T); fe_mul(r->Y, c->Y, c->Z); fe_mul(r->Z, c->Z, c->T); fe_mul(r->T, c->X, c->Y);
}
static inline void completed_to_proj(ge* r, const completed* c) { /* T left stale */
    fe_mul(r->X, c->X, c->T); fe_mul(r->Y, c->Y, c->Z); fe_mul(r->Z, c->Z, c->T);
}
static inline void ge_add_pn(completed* r, const ge* p, const pniels* n) {
    fe a, b, PP, MM, TT, ZZ, ZZ2;
    fe_add(a, p->Y, p->X); fe_sub(b, p->Y, p->X);
    fe_mul(PP, a, n->YpX); fe_mul(MM, b, n->YmX); fe_mul(TT, p->T, n->T2d); fe_mul(ZZ, p->Z, n->Z);
    fe_add(ZZ2, ZZ, ZZ);
    fe_sub(r->X, PP, MM); fe_add(r->Y, PP, MM); fe_add(r->Z, ZZ2, TT); fe_sub(r->T, ZZ2, TT);
}
static inline void ge_sub_pn(completed* r, const ge* p, const pniels* n) {
    fe a, b, PM, MP, TT, ZZ, ZZ2;
    fe_add(a, p->Y, p->X); fe_sub(b, p->Y, p->X);
    fe_mul(PM, a, n->YmX); fe_mul(MP, b, n->YpX); fe_mul(TT, p->T, n->T2d); fe_mul(ZZ, p->Z, n->Z);
    fe_add(ZZ2, ZZ, ZZ);
    fe_sub(r->X, PM, MP); fe_add(r->Y, PM, MP); fe_sub(r->Z, ZZ2, TT); fe_add(r->T, ZZ2, TT);
}
static inline void ge_dbl(completed* r, const ge* p) { /* uses X,Y,Z only */
    fe XX, YY, ZZ2, XpY, XpY2, YYpXX, YYmXX;
    fe_sq(XX, p->X); fe_sq(YY, p->Y); fe_sq(ZZ2, p->Z); fe_add(ZZ2, ZZ2, ZZ2);
    fe_add(XpY, p->X, p->Y); fe_sq(XpY2, XpY);
    fe_add(YYpXX, YY, XX); fe_sub(YYmXX, YY, XX);
    fe_sub(r->X, XpY2, YYpXX); fe_copy(r->Y, YYpXX); fe_weak_reduce(r->Y); fe_copy(r->Z, YYmXX); fe_sub(r->T, ZZ2, YYmXX);
}
static inline void ge_add(ge* r, const ge* p, const ge* q) { pniels n; completed c; ge_to_pniels(&n, q); ge_add_pn(&c, p, &n); completed_to_ge(r, &c); }
static inline void ge_sub(ge* r, const ge* p, const ge* q) { pniels n; completed c; ge_to_pniels(&n, q); ge_sub_pn(&c, p, &n); completed_to_ge(r, &c); }
static inline void ge_double(ge* r, const ge* p) { completed c; ge_dbl(&c, p); completed_to_ge(r, &c); }
static inline void ge_mul_pow2(ge* r, const ge* p, int k) { /* dalek mul_by_pow_2 */
    completed c; ge s = *p;
    for (int i = 0; i < k - 1; i++) { ge_dbl(&c, &s); completed_to_proj(&s, &c); }
    ge_dbl(&c, &s); completed_to_ge(r, &c);
}

static inline int ge_decompress(ge* p, const uint8_t b[32]) { /* SURVEY A.2; 1 = ok */
    fe s, ss, u1, u2, u2s, v, t, I, Dx, Dy, one; uint8_t chk[32];
    fe_frombytes(s, b); fe_tobytes(chk, s);
    if (memcmp(chk, b, 32) != 0 || (b[0] & 1)) return 0;
    fe_1(one); fe_sq(ss, s); fe_sub(u1, one, ss); fe_add(u2, one, ss); fe_sq(u2s, u2);
    fe_sq(t, u1); fe_mul(t, t, FE_D); fe_neg(t, t); fe_sub(v, t, u2s);
    fe_mul(t, v, u2s);
    int ok = fe_invsqrt(I, t);
    fe_mul(Dx, I, u2); fe_mul(t, I, Dx); fe_mul(Dy, t, v);
    fe_add(t, s, s); fe_mul(p->X, t, Dx); fe_abs(p->X);
    fe_mul(p->Y, u1, Dy); fe_1(p->Z); fe_mul(p->T, p->X, p->Y);
    if (!ok || fe_is_negative(p->T) || fe_is_zero(p->Y)) return 0;
    return 1;
}
static inline void ge_compress(uint8_t out[32], const ge* p) { /* SURVEY A.2 */
    fe u1, u2, t, t2, inv, i1, i2, zinv, X, Y, den, s;
    fe_add(t, p->Z, p->Y); fe_sub(t2, p->Z, p->Y); fe_mul(u1, t, t2);
    fe_mul(u2, p->X, p->Y);
    fe_sq(t, u2); fe_mul(t, t, u1); fe_invsqrt(inv, t);
    fe_mul(i1, inv, u1); fe_mul(i2, inv, u2);
    fe_mul(t, i2, p->T); fe_mul(zinv, i1, t);
    fe_copy(X, p->X); fe_copy(Y, p->Y); fe_copy(den, i2);
    fe_mul(t, p->T, zinv);
    if (fe_is_negative(t)) { fe_mul(X, p->Y, FE_SQRT_M1); fe_mul(Y, p->X, FE_SQRT_M1); fe_mul(den, i1, FE_INVSQRT_A_MINUS_D); }
    fe_mul(t, X, zinv);
    if (fe_is_negative(t)) fe_neg(Y, Y);
    fe_sub(t, p->Z, Y); fe_mul(s, den, t); fe_abs(s);
    fe_tobytes(out, s);
}
static inline void ge_elligator(ge* p, const fe r0) { /* SURVEY A.2 */
    fe r, Ns, c, Dn, s, sp, Nt, t, t2, one, W0, W1, W2, W3;
    fe_1(one);
    fe_sq(t, r0); fe_mul(r, t, FE_SQRT_M1);
    fe_add(t, r, one); fe_mul(Ns, t, FE_ONE_MINUS_D_SQ);
    fe_neg(c, one);
    fe_mul(t, FE_D, r); fe_sub(t, c, t); fe_add(t2, r, FE_D); fe_mul(Dn, t, t2);
    int sq = fe_sqrt_ratio_i(s, Ns, Dn);
    fe_mul(sp, s, r0); fe_abs(sp); fe_neg(sp, sp);
    if (!sq) { fe_copy(s, sp); fe_copy(c, r); }
    fe_sub(t, r, one); fe_mul(t, c, t); fe_mul(t, t, FE_D_MINUS_ONE_SQ); fe_sub(Nt, t, Dn);
    fe_add(t, s, s); fe_mul(W0, t, Dn);
    fe_mul(W1, Nt, FE_SQRT_AD_MINUS_ONE);
    fe_sq(t, s); fe_sub(W2, one, t); fe_add(W3, one, t);
    fe_mul(p->X, W0, W3); fe_mul(p->Y, W2, W1); fe_mul(p->Z, W1, W3); fe_mul(p->T, W0, W2);
}
static inline void ge_from_uniform(ge* p, const uint8_t b[64]) {
    fe r; ge a, c; fe_frombytes(r, b); ge_elligator(&a, r); fe_frombytes(r, b + 32); ge_elligator(&c, r); ge_add(p, &a, &c);
}

/* ---- scalar recodings ---- */
static inline void sc_radix16(int8_t d[64], const uint8_t s[32]) { /* digits in [-8,8), s < 2^255 */
    for (int i = 0; i < 32; i++) { d[2 * i] = s[i] & 15; d[2 * i + 1] = (s[i] >> 4) & 15; }
    for (int i = 0; i < 63; i++) { int8_t carry = (d[i] + 8) >> 4; d[i] -= carry << 4; d[i + 1] += carry; }
}
static inline void sc_naf5(int8_t naf[257], const uint8_t s[32]) {
    uint64_t x[5] = {load64(s), load64(s + 8), load64(s + 16), load64(s + 24), 0};
    memset(naf, 0, 257);
    int pos = 0; uint64_t carry = 0;
    while (pos < 256) {
        int wi = pos / 64, bi = pos % 64;
        uint64_t buf = (bi < 59) ? (x[wi] >> bi) : ((x[wi] >> bi) | (x[wi + 1] << (64 - bi)));
        uint64_t window = carry + (buf & 31);
        if ((window & 1) == 0) { pos++; continue; }
        if (window < 16) { carry = 0; naf[pos] = (int8_t)window; }
        else { carry = 1; naf[pos] = (int8_t)((int)window - 32); }
        pos += 5;
    }
    naf[256] = (int8_t)carry; /* cannot be set for s < 2^255 */
}

/* constant-time lookup table of [1P..8P] (dalek LookupTable<ProjectiveNielsPoint>) */
typedef struct { pniels e[8]; } lut8;
static inline void lut8_build(lut8* t, const ge* p) {
    ge q = *p; ge_to_pniels(&t->e[0], p);
    for (int j = 0; j < 7; j++) { completed c; ge_add_pn(&c, p, &t->e[j]); completed_to_ge(&q, &c); ge_to_pniels(&t->e[j + 1], &q); }
}
static inline void lut8_select(pniels* out, const lut8* t, int8_t x) { /* scan + masked select */
    int xm = x >> 7; unsigned xabs = (unsigned)((x + xm) ^ xm);
    pniels_identity(out);
    for (unsigned j = 1; j <= 8; j++) {
        uint64_t m = (uint64_t)0 - (uint64_t)(j == xabs);
        const uint64_t* src = (const uint64_t*)&t->e[j - 1]; uint64_t* dst = (uint64_t*)out;
        for (int k = 0; k < 20; k++) dst[k] ^= m & (dst[k] ^ src[k]);
    }
    if (xm) { pniels n; pniels_neg(&n, out); *out = n; }
}
/* constant-time variable-base scalar mult, radix 16 (dalek `P * s`) */
static inline void ge_scalarmult_ct(ge* r, const ge* p, const uint8_t s[32]) {
    lut8 t; int8_t d[64]; pniels n; completed c; ge q;
    lut8_build(&t, p); sc_radix16(d, s);
    ge_identity(&q); lut8_select(&n, &t, d[63]); ge_add_pn(&c, &q, &n); completed_to_ge(&q, &c);
    for (int i = 62; i >= 0; i--) {
        ge_mul_pow2(&q, &q, 4);
        lut8_select(&n, &t, d[i]); ge_add_pn(&c, &q, &n); completed_to_ge(&q, &c);
    }
    *r = q;
}
#define MSM_MAX 40
/* constant-time Straus multiscalar mult (dalek RistrettoPoint::multiscalar_mul) */
static inline void ge_msm_ct(ge* r, int n, const uint8_t (*s)[32], const ge* p) {
    static __thread lut8 t[MSM_MAX]; int8_t d[MSM_MAX][64]; pniels nn; completed c; ge q;
    for (int k = 0; k < n; k++) { lut8_build(&t[k], &p[k]); sc_radix16(d[k], s[k]); }
    ge_identity(&q);
    for (int i = 63; i >= 0; i--) {
        if (i != 63) ge_mul_pow2(&q, &q, 4);
        for (int k = 0; k < n; k++) { lut8_select(&nn, &t[k], d[k][i]); ge_add_pn(&c, &q, &nn); completed_to_ge(&q, &c); }
    }
    *r = q;
}
/* vartime Straus, width-5 NAF (dalek vartime_multiscalar_mul below 190 points) */
static inline void ge_msm_vartime(ge* r, int n, const uint8_t (*s)[32], const ge* p) {
    static __thread pniels t[MSM_MAX][8]; static __thread int8_t naf[MSM_MAX][257];
    completed c; ge q, p2;
    for (int k = 0; k < n; k++) {
        sc_naf5(naf[k], s[k]);
        ge_to_pniels(&t[k][0], &p[k]); ge_double(&p2, &p[k]);
        for (int j = 0; j < 7; j++) { ge_add_pn(&c, &p2, &t[k][j]); completed_to_ge(&q, &c); ge_to_pniels(&t[k][j + 1], &q); }
    }
    ge_identity(&q);
    int top = 256;
    for (; top >= 0; top--) { int any = 0; for (int k = 0; k < n; k++) any |= naf[k][top]; if (any) break; }
    for (int i = top; i >= 0; i--) {
        ge_dbl(&c, &q);
        for (int k = 0; k < n; k++) {
            int8_t x = naf[k][i];
            if (x > 0) { completed_to_ge(&q, &c); ge_add_pn(&c, &q, &t[k][x / 2]); }
            else if (x < 0) { completed_to_ge(&q, &c); ge_sub_pn(&c, &q, &t[k][(-x) / 2]); }
        }
        if (i == 0) completed_to_ge(&q, &c); else completed_to_proj(&q, &c);
    }
    *r = q;
}

/* ------------------------------------------------------------------ scalars mod l (little-endian u64[4]) */
static inline int sc_geq_l(const uint64_t a[4]) {
    for (int i = 3; i >= 0; i--) { if (a[i] > SC_L[i]) return 1; if (a[i] < SC_L[i]) return 0; }
    return 1;
}
static inline int sc_is_canonical(const uint8_t s[32]) { uint64_t a[4] = {load64(s), load64(s + 8), load64(s + 16), load64(s + 24)}; return !sc_geq_l(a); }
/* Barrett (HAC 14.42), b = 2^64, k = 4, x < 2^512 */
static inline void sc_reduce512(uint64_t r[4], const uint64_t x[8]) {
    uint64_t q2[10] = {0};
    for (int i = 0; i < 5; i++) { u128 c = 0; for (int j = 0; j < 5; j++) { c += (u128)x[3 + i] * SC_MU[j] + q2[i + j]; q2[i + j] = (uint64_t)c; c >>= 64; } q2[i + 5] = (uint64_t)c; }
    const uint64_t* q3 = q2 + 5; /* 5 limbs */
    uint64_t r2[5] = {0};
    static const uint64_t L5[5] = {0x5812631a5cf5d3edULL, 0x14def9dea2f79cd6ULL, 0, 0x1000000000000000ULL, 0};
    for (int i = 0; i < 5; i++) { u128 c = 0; for (int j = 0; i + j < 5; j++) { c += (u128)q3[i] * L5[j] + r2[i + j]; r2[i + j] = (uint64_t)c; c >>= 64; } }
    uint64_t t[5]; u128 bw = 0;
    for (int i = 0; i < 5; i++) { u128 d = (u128)x[i] - r2[i] - (uint64_t)bw; t[i] = (uint64_t)d; bw = (d >> 64) & 1; }
    for (int it = 0; it < 3; it++) {
        if (t[4] == 0 && !sc_geq_l(t)) break;
        u128 b2 = 0; for (int i = 0; i < 5; i++) { u128 d = (u128)t[i] - (i < 4 ? SC_L[i] : 0) - (uint64_t)b2; t[i] = (uint64_t)d; b2 = (d >> 64) & 1; }
    }
    memcpy(r, t, 32);
}
static inline void sc_from_wide(uint8_t out[32], const uint8_t in[64]) { uint64_t x[8], r[4]; memcpy(x, in, 64); sc_reduce512(r, x); memcpy(out, r, 32); }
static inline void sc_muladd(uint8_t out[32], const uint8_t a[32], const uint8_t b[32], const uint8_t c[32]) { /* a*b + c mod l */
    uint64_t A[4], B[4], C[4], x[8] = {0}, r[4]; memcpy(A, a, 32); memcpy(B, b, 32); memcpy(C, c, 32);
    for (int i = 0; i < 4; i++) { u128 cy = 0; for (int j = 0; j < 4; j++) { cy += (u128)A[i] * B[j] + x[i + j]; x[i + j] = (uint64_t)cy; cy >>= 64; } x[i + 4] = (uint64_t)cy; }
    u128 cy = 0; for (int i = 0; i < 8; i++) { cy += (u128)x[i] + (i < 4 ? C[i] : 0); x[i] = (uint64_t)cy; cy >>= 64; }
    sc_reduce512(r, x); memcpy(out, r, 32);
}
static inline void sc_mul(uint8_t out[32], const uint8_t a[32], const uint8_t b[32]) { static const uint8_t z[32] = {0}; sc_muladd(out, a, b, z); }
static inline void sc_add(uint8_t out[32], const uint8_t a[32], const uint8_t b[32]) { static const uint8_t one[32] = {1}; sc_muladd(out, a, one, b); }
static inline void sc_neg(uint8_t out[32], const uint8_t a[32]) {
    uint64_t A[4], r[4]; memcpy(A, a, 32);
    if ((A[0] | A[1] | A[2] | A[3]) == 0) { memset(out, 0, 32); return; }
    u128 bw = 0; for (int i = 0; i < 4; i++) { u128 d = (u128)SC_L[i] - A[i] - (uint64_t)bw; r[i] = (uint64_t)d; bw = (d >> 64) & 1; }
    memcpy(out, r, 32);
}

/* ------------------------------------------------------------------ Keccak / STROBE / Merlin / SHAKE */
static inline uint64_t rol64(uint64_t x, int n) { return n ? (x << n) | (x >> (64 - n)) : x; }
static inline void keccak_f1600(uint64_t A[25]) {
    static const int rho[25] = {0, 1, 62, 28, 27, 36, 44, 6, 55, 20, 3, 10, 43, 25, 39, 41, 45, 15, 21, 8, 18, 2, 61, 56, 14};
    uint64_t rc = 1; /* round constants from the LFSR, computed once */
    static uint64_t RC[24]; static int init = 0;
    if (!init) {
        uint8_t lfsr = 1;
        for (int r = 0; r < 24; r++) { uint64_t c = 0; for (int j = 0; j < 7; j++) { if (lfsr & 1) c |= 1ULL << ((1 << j) - 1); lfsr = (uint8_t)((lfsr << 1) ^ ((lfsr >> 7) * 0x71)); } RC[r] = c; }
        __sync_synchronize(); init = 1;
    }
    (void)rc;
    for (int r = 0; r < 24; r++) {
        uint64_t C[5], D[5], B[25];
        for (int x = 0; x < 5; x++) C[x] = A[x] ^ A[x + 5] ^ A[x + 10] ^ A[x + 15] ^ A[x + 20];
        for (int x = 0; x < 5; x++) D[x] = C[(x + 4) % 5] ^ rol64(C[(x + 1) % 5], 1);
        for (int i = 0; i < 25; i++) A[i] ^= D[i % 5];
        for (int x = 0; x < 5; x++) for (int y = 0; y < 5; y++) B[y + 5 * ((2 * x + 3 * y) % 5)] = rol64(A[x + 5 * y], rho[x + 5 * y]);
        for (int y = 0; y < 5; y++) for (int x = 0; x < 5; x++) A[x + 5 * y] = B[x + 5 * y] ^ (~B[(x + 1) % 5 + 5 * y] & B[(x + 2) % 5 + 5 * y]);
        A[0] ^= RC[r];
    }
}
static inline void shake256(uint8_t* out, size_t outlen, const uint8_t* in, size_t inlen) { /* outlen <= 136 */
    uint64_t st[25] = {0}; uint8_t* b = (uint8_t*)st;
    while (inlen >= 136) { for (int i = 0; i < 136; i++) b[i] ^= in[i]; keccak_f1600(st); in += 136; inlen -= 136; }
    for (size_t i = 0; i < inlen; i++) b[i] ^= in[i];
    b[inlen] ^= 0x1f; b[135] ^= 0x80; keccak_f1600(st);
    memcpy(out, b, outlen);
}

typedef struct { uint64_t st[25]; uint8_t pos, pos_begin, cur_flags; uint32_t perms; } strobe;
#define STROBE_R 166
static inline void strobe_run_f(strobe* s) {
    uint8_t* b = (uint8_t*)s->st;
    b[s->pos] ^= s->pos_begin; b[s->pos + 1] ^= 0x04; b[STROBE_R + 1] ^= 0x80;
    keccak_f1600(s->st); s->perms++; s->pos = 0; s->pos_begin = 0;
}
static inline void strobe_absorb(strobe* s, const uint8_t* d, size_t n) {
    uint8_t* b = (uint8_t*)s->st;
    for (size_t i = 0; i < n; i++) { b[s->pos++] ^= d[i]; if (s->pos == STROBE_R) strobe_run_f(s); }
}
static inline void strobe_begin_op(strobe* s, uint8_t flags, int more) {
    if (more) return;
    uint8_t hdr[2] = {s->pos_begin, flags};
    s->pos_begin = s->pos + 1; s->cur_flags = flags;
    strobe_absorb(s, hdr, 2);
    if ((flags & (4 | 32)) && s->pos != 0) strobe_run_f(s);
}
static inline void strobe_meta_ad(strobe* s, const void* d, size_t n, int more) { strobe_begin_op(s, 16 | 2, more); strobe_absorb(s, (const uint8_t*)d, n); }
static inline void strobe_ad(strobe* s, const void* d, size_t n, int more) { strobe_begin_op(s, 2, more); strobe_absorb(s, (const uint8_t*)d, n); }
static inline void strobe_prf(strobe* s, uint8_t* out, size_t n) {
    strobe_begin_op(s, 1 | 2 | 4, 0);
    uint8_t* b = (uint8_t*)s->st;
    for (size_t i = 0; i < n; i++) { out[i] = b[s->pos]; b[s->pos++] = 0; if (s->pos == STROBE_R) strobe_run_f(s); }
}
static inline void strobe_init(strobe* s, const char* label) {
    memset(s, 0, sizeof(*s));
    uint8_t* b = (uint8_t*)s->st;
    static const uint8_t hdr[6] = {1, STROBE_R + 2, 1, 0, 1, 96};
    memcpy(b, hdr, 6); memcpy(b + 6, "STROBEv1.0.2", 12);
    keccak_f1600(s->st); s->perms++;
    strobe_meta_ad(s, label, strlen(label), 0);
}
static inline void merlin_append(strobe* s, const char* label, const void* msg, uint32_t len) {
    strobe_meta_ad(s, label, strlen(label), 0);
    uint8_t l4[4] = {(uint8_t)len, (uint8_t)(len >> 8), (uint8_t)(len >> 16), (uint8_t)(len >> 24)};
    strobe_meta_ad(s, l4, 4, 1);
    strobe_ad(s, msg, len, 0);
}
static inline void merlin_init(strobe* s, const char* label) { strobe_init(s, "Merlin v1.0"); merlin_append(s, "dom-sep", label, (uint32_t)strlen(label)); }
static inline void merlin_challenge(strobe* s, const char* label, uint8_t* out, uint32_t n) {
    strobe_meta_ad(s, label, strlen(label), 0);
    uint8_t l4[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
    strobe_meta_ad(s, l4, 4, 1);
    strobe_prf(s, out, n);
}

/* ------------------------------------------------------------------ SHA-512 */
static inline uint64_t ror64(uint64_t x, int n) { return (x >> n) | (x << (64 - n)); }
static inline void sha512(uint8_t out[64], const uint8_t* in, size_t len) {
    uint64_t H[8]; memcpy(H, SHA512_H0, 64);
    uint8_t buf[256]; size_t full = len / 128, rem = len % 128;
    size_t tail = (rem < 112) ? 128 : 256;
    memset(buf, 0, sizeof buf); memcpy(buf, in + full * 128, rem); buf[rem] = 0x80;
    uint64_t bits = (uint64_t)len * 8; for (int i = 0; i < 8; i++) buf[tail - 1 - i] = (uint8_t)(bits >> (8 * i));
    for (size_t blk = 0; blk < full + tail / 128; blk++) {
        const uint8_t* p = blk < full ? in + blk * 128 : buf + (blk - full) * 128;
        uint64_t W[80];
        for (int i = 0; i < 16; i++) { uint64_t v = 0; for (int j = 0; j < 8; j++) v = (v << 8) | p[8 * i + j]; W[i] = v; }
        for (int i = 16; i < 80; i++) {
            uint64_t s0 = ror64(W[i - 15], 1) ^ ror64(W[i - 15], 8) ^ (W[i - 15] >> 7);
            uint64_t s1 = ror64(W[i - 2], 19) ^ ror64(W[i - 2], 61) ^ (W[i - 2] >> 6);
            W[i] = W[i - 16] + s0 + W[i - 7] + s1;
        }
        uint64_t a = H[0], b = H[1], c = H[2], d = H[3], e = H[4], f = H[5], g = H[6], h = H[7];
        for (int i = 0; i < 80; i++) {
            uint64_t S1 = ror64(e, 14) ^ ror64(e, 18) ^ ror64(e, 41), ch = (e & f) ^ (~e & g);
            uint64_t t1 = h + S1 + ch + SHA512_K[i] + W[i];
            uint64_t S0 = ror64(a, 28) ^ ror64(a, 34) ^ ror64(a, 39), mj = (a & b) ^ (a & c) ^ (b & c);
            uint64_t t2 = S0 + mj;
            h = g; g = f; f = e; e = d + t1; d = c; c = b; b = a; a = t1 + t2;
        }
        H[0] += a; H[1] += b; H[2] += c; H[3] += d; H[4] += e; H[5] += f; H[6] += g; H[7] += h;
    }
    for (int i = 0; i < 8; i++) for (int j = 0; j < 8; j++) out[8 * i + j] = (uint8_t)(H[i] >> (56 - 8 * j));
}
