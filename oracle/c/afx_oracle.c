/* ORACLE (test infrastructure, NOT the product).
 *
 * C restatement of aeonflux's hot path in the reference's own CPU operation schedule
 * (paths relative to /root/reference/src):
 *   Issuer::verify            issuer.rs:141-147 -> nizk/presentation.rs:324-443
 *   ProofOfEncryption::verify nizk/encryption.rs:154-210
 *   Issuer::issue             issuer.rs:111-124 -> amacs.rs:276-294,256-272,224-244 -> nizk/issuance.rs:40-129
 *   CredentialIssuance::verify issuer.rs:48-57 -> nizk/issuance.rs:132-218
 *   AnonymousCredential::show credential.rs:37-46 -> presentation.rs:139-321, encryption.rs:58-142 (generator)
 * and of the zkp 0.7 toolbox it calls.  "Reference schedule" means: seven separate
 * constant-time scalar mults for the aMAC, compress() of every allocated point followed by
 * verify_compact's re-decompress, one vartime Straus NAF-5 MSM per constraint with no precomputed
 * tables.  It is the CPU baseline bench.py reports ("restated reference CPU path, not the Rust binary")
 * and the checker for GPU parity tests at sizes the Python oracle cannot reach; it is itself checked
 * against the Python big-int oracle byte for byte (tests/test_oracle_c.py).
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 */
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <time.h>
#include "prim.h"

#define MAXN 32
#define MAXPT 112
#define MAXSC 48
#define MAXC 40
#define MAXT 40

enum { K_PS = 0, K_SS = 1, K_PP = 2, K_SP = 3 };

typedef struct {
    uint32_t n, ny;
    ge G, G_w, G_wp, G_x0, G_x1, G_y[MAXN], G_m[MAXN], G_V, G_a, G_a0, G_a1;
    ge C_W, I;
    int has_secret;
    uint8_t w[32], wp[32], x0[32], x1[32], y[MAXN][32];
    ge W;
} afxo_issuer;

/* ------------------------------------------------------------------ zkp toolbox */
typedef struct {
    strobe t;
    int nsc, npt, ncons;
    uint8_t pt_enc[MAXPT][32];
    const char* pt_label[MAXPT];
    struct { int lhs, nterms, sc[MAXT], pt[MAXT]; } cons[MAXC];
    /* prover only */
    uint8_t sc_val[MAXSC][32];
    ge pt_val[MAXPT];
    /* trace */
    uint8_t commitments[MAXC][32];
    uint8_t challenge[32];
    int have_challenge;
} zk;

static void zk_init(zk* z, const char* tlabel, const char* plabel) {
    z->nsc = z->npt = z->ncons = 0; z->have_challenge = 0;
    merlin_init(&z->t, tlabel);
    merlin_append(&z->t, "dom-sep", "schnorrzkp/1.0/ristretto255", 27);
    merlin_append(&z->t, "dom-sep", plabel, (uint32_t)strlen(plabel));
}
static int zk_scalar(zk* z, const char* label) { merlin_append(&z->t, "scvar", label, (uint32_t)strlen(label)); return z->nsc++; }
static int zk_scalar_val(zk* z, const char* label, const uint8_t v[32]) { memcpy(z->sc_val[z->nsc], v, 32); return zk_scalar(z, label); }
/* verifier: validate_and_append_point_var; returns -1 on identity */
static int zk_point_v(zk* z, const char* label, const uint8_t enc[32]) {
    static const uint8_t zero[32] = {0};
    if (memcmp(enc, zero, 32) == 0) return -1;
    merlin_append(&z->t, "ptvar", label, (uint32_t)strlen(label));
    merlin_append(&z->t, "val", enc, 32);
    memcpy(z->pt_enc[z->npt], enc, 32); z->pt_label[z->npt] = label;
    return z->npt++;
}
/* verifier convenience: compress (reference schedule) then allocate */
static int zk_point_vc(zk* z, const char* label, const ge* p) { uint8_t e[32]; ge_compress(e, p); return zk_point_v(z, label, e); }
/* prover: append_point_var (no identity check) */
static int zk_point_p(zk* z, const char* label, const ge* p) {
    uint8_t e[32]; ge_compress(e, p);
    merlin_append(&z->t, "ptvar", label, (uint32_t)strlen(label));
    merlin_append(&z->t, "val", e, 32);
    z->pt_val[z->npt] = *p; z->pt_label[z->npt] = label;
    return z->npt++;
}
static int zk_constrain(zk* z, int lhs) { int c = z->ncons++; z->cons[c].lhs = lhs; z->cons[c].nterms = 0; return c; }
static void zk_term(zk* z, int c, int sc, int pt) { int k = z->cons[c].nterms++; z->cons[c].sc[k] = sc; z->cons[c].pt[k] = pt; }
static void zk_get_challenge(zk* z, uint8_t out[32]) { uint8_t b[64]; merlin_challenge(&z->t, "chal", b, 64); sc_from_wide(out, b); }

/* Verifier::verify_compact; 0 = ok, 1 = VerificationFailure */
static int zk_verify_compact(zk* z, const uint8_t c[32], int nresp, const uint8_t (*resp)[32]) {
    static __thread ge pts[MAXPT];
    if (nresp != z->nsc) return 1;
    for (int i = 0; i < z->npt; i++) if (!ge_decompress(&pts[i], z->pt_enc[i])) return 1;
    uint8_t minus_c[32]; sc_neg(minus_c, c);
    for (int k = 0; k < z->ncons; k++) {
        uint8_t s[MAXT + 1][32]; ge p[MAXT + 1]; int m = z->cons[k].nterms;
        for (int j = 0; j < m; j++) { memcpy(s[j], resp[z->cons[k].sc[j]], 32); p[j] = pts[z->cons[k].pt[j]]; }
        memcpy(s[m], minus_c, 32); p[m] = pts[z->cons[k].lhs];
        ge Rr; ge_msm_vartime(&Rr, m + 1, (const uint8_t(*)[32])s, p);
        ge_compress(z->commitments[k], &Rr);
        merlin_append(&z->t, "blindcom", z->pt_label[z->cons[k].lhs], (uint32_t)strlen(z->pt_label[z->cons[k].lhs]));
        merlin_append(&z->t, "val", z->commitments[k], 32);
    }
    zk_get_challenge(z, z->challenge); z->have_challenge = 1;
    return memcmp(z->challenge, c, 32) != 0;
}
/* Prover::prove_compact with supplied blindings */
static void zk_prove_compact(zk* z, const uint8_t (*blind)[32], uint8_t c[32], uint8_t (*resp)[32]) {
    for (int k = 0; k < z->ncons; k++) {
        uint8_t s[MAXT][32]; ge p[MAXT]; int m = z->cons[k].nterms;
        for (int j = 0; j < m; j++) { memcpy(s[j], blind[z->cons[k].sc[j]], 32); p[j] = z->pt_val[z->cons[k].pt[j]]; }
        ge Rr; ge_msm_ct(&Rr, m, (const uint8_t(*)[32])s, p);
        ge_compress(z->commitments[k], &Rr);
        merlin_append(&z->t, "blindcom", z->pt_label[z->cons[k].lhs], (uint32_t)strlen(z->pt_label[z->cons[k].lhs]));
        merlin_append(&z->t, "val", z->commitments[k], 32);
    }
    zk_get_challenge(z, c);
    for (int i = 0; i < z->nsc; i++) sc_muladd(resp[i], z->sc_val[i], c, blind[i]);
}

/* ------------------------------------------------------------------ in-memory structures (what the Rust types hold) */
typedef struct { uint8_t c[32], resp[6][32]; ge pk, E1, E2, C_y_1, C_y_2, C_y_3, C_y_2p; int index; } enc_proof;
typedef struct {
    int n; uint8_t kinds[MAXN];
    uint8_t c[32]; int nresp; uint8_t resp[3 + MAXN][32];
    ge C_x_0, C_x_1, C_V, C_y[MAXN];
    uint8_t rev_scalar[MAXN][32]; ge rev_point[MAXN];
    int nenc; enc_proof enc[MAXN];
} presentation;

typedef struct { uint8_t Z[32]; int have_Z; int ncommit; uint8_t commit[MAXC * (MAXN + 1)][32]; int nchal; uint8_t chal[MAXN + 1][32]; } vtrace;

static int presentation_words(int n, const uint8_t* kinds) {
    int hs = 0, r = 0, hp = 0;
    for (int i = 0; i < n; i++) { hs += kinds[i] == K_SS; r += (kinds[i] == K_PS || kinds[i] == K_PP); hp += kinds[i] == K_SP; }
    return 1 + 3 + hs + 3 + n + r + 14 * hp;
}
/* flat words -> presentation; 0 ok, 1 = undecodable point / non-canonical scalar (flat-wire rule) */
static int presentation_parse(presentation* p, int n, const uint8_t* kinds, const uint8_t (*w)[32]) {
    int k = 0, hs = 0;
    p->n = n; memcpy(p->kinds, kinds, n);
    for (int i = 0; i < n; i++) hs += kinds[i] == K_SS;
#define SC(dst) do { if (!sc_is_canonical(w[k])) return 1; memcpy(dst, w[k], 32); k++; } while (0)
#define PT(dst) do { if (!ge_decompress(dst, w[k])) return 1; k++; } while (0)
    SC(p->c); p->nresp = 3 + hs;
    for (int i = 0; i < p->nresp; i++) SC(p->resp[i]);
    PT(&p->C_x_0); PT(&p->C_x_1); PT(&p->C_V);
    for (int i = 0; i < n; i++) PT(&p->C_y[i]);
    for (int i = 0; i < n; i++) { if (kinds[i] == K_PS) SC(p->rev_scalar[i]); else if (kinds[i] == K_PP) PT(&p->rev_point[i]); }
    p->nenc = 0;
    for (int i = 0; i < n; i++) {
        if (kinds[i] != K_SP) continue;
        enc_proof* e = &p->enc[p->nenc++]; e->index = i;
        SC(e->c); for (int j = 0; j < 6; j++) SC(e->resp[j]);
        PT(&e->pk); PT(&e->E1); PT(&e->E2); PT(&e->C_y_1); PT(&e->C_y_2); PT(&e->C_y_3); PT(&e->C_y_2p);
    }
    return 0;
}

/* ProofOfEncryption::verify, encryption.rs:154-210 */
static int encryption_verify(const afxo_issuer* is, const enc_proof* e, vtrace* tr) {
    zk z; ge t;
    zk_init(&z, "2019/1416 anonymous credentials", "2019/1416 proof of encryption");
    int a = zk_scalar(&z, "a"), a0 = zk_scalar(&z, "a0"), a1 = zk_scalar(&z, "a1"), m3 = zk_scalar(&z, "m3"), zz = zk_scalar(&z, "z"), z1 = zk_scalar(&z, "z1");
    int pk, G_a, G_a_0, G_a_1, G_y_1, G_y_2, G_y_3, G_m_3, C_y_2, C_y_3, C_y_2p, CmE, E1, mE1;
#define AP(var, label, pt) do { var = zk_point_vc(&z, label, pt); if (var < 0) return 1; } while (0)
    AP(pk, "pk", &e->pk); AP(G_a, "G_a", &is->G_a); AP(G_a_0, "G_a_0", &is->G_a0); AP(G_a_1, "G_a_1", &is->G_a1);
    AP(G_y_1, "G_y_1", &is->G_y[0]); AP(G_y_2, "G_y_2", &is->G_y[1]); AP(G_y_3, "G_y_3", &is->G_y[2]);
    AP(G_m_3, "G_m_3", &is->G_m[e->index]);
    AP(C_y_2, "C_y_2", &e->C_y_2); AP(C_y_3, "C_y_3", &e->C_y_3); AP(C_y_2p, "C_y_2'", &e->C_y_2p);
    ge_sub(&t, &e->C_y_1, &e->E2); AP(CmE, "C_y_1-E2", &t);
    AP(E1, "E1", &e->E1);
    ge_neg(&t, &e->E1); AP(mE1, "-E1", &t);
    int c;
    c = zk_constrain(&z, pk); zk_term(&z, c, a, G_a); zk_term(&z, c, a0, G_a_0); zk_term(&z, c, a1, G_a_1);
    c = zk_constrain(&z, CmE); zk_term(&z, c, zz, G_y_1); zk_term(&z, c, a, mE1);
    c = zk_constrain(&z, C_y_2p); zk_term(&z, c, a1, C_y_2);
    c = zk_constrain(&z, E1); zk_term(&z, c, a0, C_y_2); zk_term(&z, c, m3, C_y_2p); zk_term(&z, c, z1, G_y_2);
    c = zk_constrain(&z, C_y_3); zk_term(&z, c, zz, G_y_3); zk_term(&z, c, m3, G_m_3);
    int rc = zk_verify_compact(&z, e->c, 6, e->resp);
    if (tr && z.have_challenge) {
        for (int k = 0; k < z.ncons; k++) memcpy(tr->commit[tr->ncommit++], z.commitments[k], 32);
        memcpy(tr->chal[tr->nchal++], z.challenge, 32);
    }
    return rc;
}

/* ProofOfValidCredential::verify, presentation.rs:324-443; 0 ok, 1 VerificationFailure, 2 structural (Rust panics) */
static int presentation_verify(const afxo_issuer* is, const presentation* p, vtrace* tr) {
    int n = p->n;
    if ((uint32_t)n != is->n) return 2;
    /* :342-352 -- seven separate constant-time scalar mults (n = 4) */
    ge Z, t, x;
    ge_sub(&Z, &p->C_V, &is->W);
    ge_scalarmult_ct(&t, &p->C_x_0, is->x0); ge_sub(&Z, &Z, &t);
    ge_scalarmult_ct(&t, &p->C_x_1, is->x1); ge_sub(&Z, &Z, &t);
    for (int i = 0; i < n; i++) {
        if (p->kinds[i] == K_PS) { ge_scalarmult_ct(&t, &is->G_m[i], p->rev_scalar[i]); ge_add(&x, &p->C_y[i], &t); }
        else if (p->kinds[i] == K_PP) ge_add(&x, &p->C_y[i], &p->rev_point[i]);
        else x = p->C_y[i];
        ge_scalarmult_ct(&t, &x, is->y[i]); ge_sub(&Z, &Z, &t);
    }
    if (tr) { ge_compress(tr->Z, &Z); tr->have_Z = 1; }
    zk z;
    zk_init(&z, "2019/1416 anonymous credential", "2019/1416 presentation proof");
    int zz = zk_scalar(&z, "z"), z_0 = zk_scalar(&z, "z_0"), tt = zk_scalar(&z, "t");
    int H_s[MAXN], G_m[MAXN];
    for (int i = 0; i < n; i++) { H_s[i] = -1; G_m[i] = -1; }
    for (int i = 0; i < n; i++) if (p->kinds[i] == K_SS) H_s[i] = zk_scalar(&z, "m");
    int I, C_x_1, C_x_0, G_x_0, G_x_1, C_y[MAXN], ncy = 0, G_y[MAXN], Zv;
    AP(I, "I", &is->I); AP(C_x_1, "C_x_1", &p->C_x_1); AP(C_x_0, "C_x_0", &p->C_x_0);
    AP(G_x_0, "G_x_0", &is->G_x0); AP(G_x_1, "G_x_1", &is->G_x1);
    for (int i = 0; i < n; i++) { if (p->kinds[i] == K_SP) continue; AP(C_y[ncy], "C_y", &p->C_y[i]); ncy++; }
    for (uint32_t i = 0; i < is->ny; i++) AP(G_y[i], "G_y", &is->G_y[i]);
    for (int i = 0; i < n; i++) if (p->kinds[i] == K_SS) AP(G_m[i], "G_m", &is->G_m[i]);
    AP(Zv, "Z", &Z);
    int c;
    c = zk_constrain(&z, Zv); zk_term(&z, c, zz, I);
    c = zk_constrain(&z, C_x_1); zk_term(&z, c, tt, C_x_0); zk_term(&z, c, z_0, G_x_0); zk_term(&z, c, zz, G_x_1);
    for (int i = 0; i < ncy; i++) { /* :427-433 compacted-index loop */
        if (p->kinds[i] == K_SP) continue;
        c = zk_constrain(&z, C_y[i]); zk_term(&z, c, zz, G_y[i]);
        if (p->kinds[i] == K_SS) zk_term(&z, c, H_s[i], G_m[i]);
    }
    int rc = zk_verify_compact(&z, p->c, p->nresp, p->resp);
    if (tr && z.have_challenge) {
        for (int k = 0; k < z.ncons; k++) memcpy(tr->commit[tr->ncommit++], z.commitments[k], 32);
        memcpy(tr->chal[tr->nchal++], z.challenge, 32);
    }
    if (rc) return 1;
    for (int j = 0; j < p->nenc; j++) if (encryption_verify(is, &p->enc[j], tr)) return 1; /* :438-440 */
    return 0;
}

/* ------------------------------------------------------------------ issuance */
typedef struct { int n; uint8_t kinds[MAXN]; uint8_t sc[MAXN][32]; ge pt[MAXN]; } attributes; /* kinds: K_PS scalar, K_PP point */

/* Messages::from_attributes, amacs.rs:224-244 */
static void messages(const afxo_issuer* is, const attributes* a, ge* M) {
    for (int i = 0; i < a->n; i++) { if (a->kinds[i] == K_PS) ge_scalarmult_ct(&M[i], &is->G_m[i], a->sc[i]); else M[i] = a->pt[i]; }
}
/* Amac::compute_V, amacs.rs:256-272 */
static void compute_V(const afxo_issuer* is, const attributes* a, const uint8_t t[32], const ge* U, ge* V) {
    ge M[MAXN], tmp; uint8_t x1t[32];
    messages(is, a, M);
    ge_scalarmult_ct(&tmp, U, is->x0); ge_add(V, &is->W, &tmp);
    sc_mul(x1t, is->x1, t); ge_scalarmult_ct(&tmp, U, x1t); ge_add(V, V, &tmp);
    ge_msm_ct(&tmp, a->n, (const uint8_t(*)[32])is->y, M); ge_add(V, V, &tmp);
}
static void issuance_statement(zk* z, const afxo_issuer* is, int prover, const ge* U, const ge* V, const ge* tU, const ge* M, int* ok) {
    int n = (int)is->n; static const uint8_t one[32] = {1};
    zk_init(z, "2019/1416 anonymous credential", "2019/1416 issuance proof");
    int w, wp, x0, x1, y[MAXN], o;
    if (prover) {
        w = zk_scalar_val(z, "w", is->w); wp = zk_scalar_val(z, "w'", is->wp); x0 = zk_scalar_val(z, "x_0", is->x0); x1 = zk_scalar_val(z, "x_1", is->x1);
        for (int i = 0; i < n; i++) y[i] = zk_scalar_val(z, "y", is->y[i]);
        o = zk_scalar_val(z, "1", one);
    } else {
        w = zk_scalar(z, "w"); wp = zk_scalar(z, "w'"); x0 = zk_scalar(z, "x_0"); x1 = zk_scalar(z, "x_1");
        for (int i = 0; i < n; i++) y[i] = zk_scalar(z, "y");
        o = zk_scalar(z, "1");
    }
    ge t; int G_V, G_w, G_wp, nGx0, nGx1, nGy[MAXN], C_W, I, Uv, Vv, tUv, Mv[MAXN];
    *ok = 0;
#define AQ(var, label, pt) do { var = prover ? zk_point_p(z, label, pt) : zk_point_vc(z, label, pt); if (var < 0) return; } while (0)
    AQ(G_V, "G_V", &is->G_V); AQ(G_w, "G_w", &is->G_w); AQ(G_wp, "G_w_prime", &is->G_wp);
    ge_neg(&t, &is->G_x0); AQ(nGx0, "-G_x_0", &t);
    ge_neg(&t, &is->G_x1); AQ(nGx1, "-G_x_1", &t);
    for (uint32_t i = 0; i < is->ny; i++) { ge_neg(&t, &is->G_y[i]); AQ(nGy[i], "-G_y", &t); }
    AQ(C_W, "C_W", &is->C_W); AQ(I, "I", &is->I); AQ(Uv, "U", U); AQ(Vv, "V", V); AQ(tUv, "tU", tU);
    for (int i = 0; i < n; i++) AQ(Mv[i], "M", &M[i]);
    int c;
    c = zk_constrain(z, C_W); zk_term(z, c, w, G_w); zk_term(z, c, wp, G_wp);
    c = zk_constrain(z, I); zk_term(z, c, o, G_V); zk_term(z, c, x0, nGx0); zk_term(z, c, x1, nGx1);
    for (int i = 0; i < n; i++) zk_term(z, c, y[i], nGy[i]);
    c = zk_constrain(z, Vv); zk_term(z, c, w, G_w); zk_term(z, c, x0, Uv); zk_term(z, c, x1, tUv);
    for (int i = 0; i < n; i++) zk_term(z, c, y[i], Mv[i]);
    *ok = 1;
}
/* Issuer::issue with supplied randomness: issuer.rs:111-124, issuance.rs:40-129 */
static void issue(const afxo_issuer* is, const attributes* a, const uint8_t t[32], const ge* U, const uint8_t (*blind)[32],
                  ge* V, uint8_t c[32], uint8_t (*resp)[32]) {
    static __thread zk z; ge tU, M[MAXN]; int ok;
    compute_V(is, a, t, U, V);
    ge_scalarmult_ct(&tU, U, t);          /* issuance.rs:91 */
    messages(is, a, M);                    /* issuance.rs:95 (recomputed) */
    issuance_statement(&z, is, 1, U, V, &tU, M, &ok);
    zk_prove_compact(&z, blind, c, resp);
}
/* CredentialIssuance::verify: issuer.rs:48-57, issuance.rs:132-218 */
static int issuance_verify(const afxo_issuer* is, const attributes* a, const uint8_t t[32], const ge* U, const ge* V,
                           const uint8_t c[32], const uint8_t (*resp)[32], vtrace* tr) {
    static __thread zk z; ge tU, M[MAXN]; int ok;
    ge_scalarmult_ct(&tU, U, t);          /* issuance.rs:180 */
    messages(is, a, M);                    /* issuance.rs:184 */
    issuance_statement(&z, is, 0, U, V, &tU, M, &ok);
    if (!ok) return 1;
    int rc = zk_verify_compact(&z, c, (int)is->n + 5, resp);
    if (tr && z.have_challenge) {
        for (int k = 0; k < z.ncons; k++) memcpy(tr->commit[tr->ncommit++], z.commitments[k], 32);
        memcpy(tr->chal[tr->nchal++], z.challenge, 32);
    }
    return rc;
}

/* ------------------------------------------------------------------ user side (generator only) */
typedef struct { ge M1, M2; uint8_t m3[32]; } plaintext;
typedef struct { uint8_t a[32], a0[32], a1[32]; ge pk; } sym_keypair;

static void encode_to_group(ge* out, const uint8_t data[30]) { /* encoding.rs:56-70 */
    uint8_t b[32] = {0}; memcpy(b + 1, data, 30);
    for (int j = 0; j < 64; j++) { b[31] = (uint8_t)j; for (int i = 0; i < 128; i++) { b[0] = (uint8_t)(2 * i); if (ge_decompress(out, b)) return; } }
    abort();
}
static void plaintext_from30(plaintext* p, const uint8_t src[30]) { /* symmetric.rs:135-143 */
    uint8_t h[64]; encode_to_group(&p->M1, src); sha512(h, src, 30); ge_from_uniform(&p->M2, h); sc_from_wide(p->m3, h);
}
static void keypair_derive(sym_keypair* k, const afxo_issuer* is, const uint8_t ms[64]) { /* symmetric.rs:197-215 */
    uint8_t h[64]; ge t;
    sha512(h, ms, 64); sc_from_wide(k->a, h);
    sha512(h, k->a, 32); sc_from_wide(k->a0, h);
    sha512(h, k->a0, 32); sc_from_wide(k->a1, h);
    ge_scalarmult_ct(&k->pk, &is->G_a, k->a);
    ge_scalarmult_ct(&t, &is->G_a0, k->a0); ge_add(&k->pk, &k->pk, &t);
    ge_scalarmult_ct(&t, &is->G_a1, k->a1); ge_add(&k->pk, &k->pk, &t);
}
/* ProofOfEncryption::prove, encryption.rs:58-142 */
static void encryption_prove(const afxo_issuer* is, const plaintext* pt, int index, const sym_keypair* kp, const uint8_t zs[32],
                             const uint8_t (*blind)[32], enc_proof* e) {
    static __thread zk z; ge t; uint8_t k[32], z1s[32];
    sc_muladd(k, kp->a1, pt->m3, kp->a0);                       /* a0 + a1*m3 */
    ge_scalarmult_ct(&e->E1, &pt->M2, k);                        /* symmetric.rs:257 */
    ge_scalarmult_ct(&t, &e->E1, kp->a); ge_add(&e->E2, &t, &pt->M1);
    ge_scalarmult_ct(&t, &is->G_y[0], zs); ge_add(&e->C_y_1, &t, &pt->M1);
    ge_scalarmult_ct(&t, &is->G_y[1], zs); ge_add(&e->C_y_2, &t, &pt->M2);
    ge_scalarmult_ct(&t, &is->G_y[2], zs); ge_scalarmult_ct(&e->C_y_3, &is->G_m[index], pt->m3); ge_add(&e->C_y_3, &t, &e->C_y_3);
    ge_scalarmult_ct(&e->C_y_2p, &e->C_y_2, kp->a1);
    sc_mul(z1s, zs, k); sc_neg(z1s, z1s);
    e->index = index;
    zk_init(&z, "2019/1416 anonymous credentials", "2019/1416 proof of encryption");
    int a = zk_scalar_val(&z, "a", kp->a), a0 = zk_scalar_val(&z, "a0", kp->a0), a1 = zk_scalar_val(&z, "a1", kp->a1);
    int m3 = zk_scalar_val(&z, "m3", pt->m3), zz = zk_scalar_val(&z, "z", zs), z1 = zk_scalar_val(&z, "z1", z1s);
    int pk = zk_point_p(&z, "pk", &kp->pk), G_a = zk_point_p(&z, "G_a", &is->G_a), G_a_0 = zk_point_p(&z, "G_a_0", &is->G_a0), G_a_1 = zk_point_p(&z, "G_a_1", &is->G_a1);
    int G_y_1 = zk_point_p(&z, "G_y_1", &is->G_y[0]), G_y_2 = zk_point_p(&z, "G_y_2", &is->G_y[1]), G_y_3 = zk_point_p(&z, "G_y_3", &is->G_y[2]);
    int G_m_3 = zk_point_p(&z, "G_m_3", &is->G_m[index]);
    int C_y_2 = zk_point_p(&z, "C_y_2", &e->C_y_2), C_y_3 = zk_point_p(&z, "C_y_3", &e->C_y_3), C_y_2p = zk_point_p(&z, "C_y_2'", &e->C_y_2p);
    ge_sub(&t, &e->C_y_1, &e->E2); int CmE = zk_point_p(&z, "C_y_1-E2", &t);
    int E1 = zk_point_p(&z, "E1", &e->E1);
    ge_neg(&t, &e->E1); int mE1 = zk_point_p(&z, "-E1", &t);
    int c;
    c = zk_constrain(&z, pk); zk_term(&z, c, a, G_a); zk_term(&z, c, a0, G_a_0); zk_term(&z, c, a1, G_a_1);
    c = zk_constrain(&z, CmE); zk_term(&z, c, zz, G_y_1); zk_term(&z, c, a, mE1);
    c = zk_constrain(&z, C_y_2p); zk_term(&z, c, a1, C_y_2);
    c = zk_constrain(&z, E1); zk_term(&z, c, a0, C_y_2); zk_term(&z, c, m3, C_y_2p); zk_term(&z, c, z1, G_y_2);
    c = zk_constrain(&z, C_y_3); zk_term(&z, c, zz, G_y_3); zk_term(&z, c, m3, G_m_3);
    e->pk = kp->pk;
    zk_prove_compact(&z, blind, e->c, e->resp);
}

/* credential attribute as the user holds it */
typedef struct { uint8_t kind; /* 'S' PublicScalar 's' SecretScalar 'P' PublicPoint 'E' EitherPoint 'H' SecretPoint */ uint8_t sc[32]; ge pt; plaintext pl; } cred_attr;

/* ProofOfValidCredential::prove, presentation.rs:139-321 */
static void presentation_prove(const afxo_issuer* is, const cred_attr* at, const uint8_t t[32], const ge* U, const ge* V,
                               const sym_keypair* kp, const uint8_t zs[32], const uint8_t (*blind)[32], const uint8_t (*enc_blind)[32],
                               presentation* p) {
    static __thread zk z; int n = (int)is->n; ge tmp, tmp2, Zp; uint8_t z0s[32];
    sc_mul(z0s, t, zs); sc_neg(z0s, z0s);
    p->n = n; p->nenc = 0;
    for (int i = 0; i < n; i++) {
        ge_scalarmult_ct(&tmp, &is->G_y[i], zs);
        switch (at[i].kind) {
        case 'P': case 'E': case 'S': p->C_y[i] = tmp; break;
        case 'H': ge_add(&p->C_y[i], &tmp, &at[i].pl.M1); break;
        default: ge_scalarmult_ct(&tmp2, &is->G_m[i], at[i].sc); ge_add(&p->C_y[i], &tmp, &tmp2); break;
        }
    }
    ge_scalarmult_ct(&tmp, &is->G_x0, zs); ge_add(&p->C_x_0, &tmp, U);
    ge_scalarmult_ct(&tmp, &is->G_x1, zs); ge_scalarmult_ct(&tmp2, U, t); ge_add(&p->C_x_1, &tmp, &tmp2);
    ge_scalarmult_ct(&tmp, &is->G_V, zs); ge_add(&p->C_V, &tmp, V);
    ge_scalarmult_ct(&Zp, &is->I, zs);
    zk_init(&z, "2019/1416 anonymous credential", "2019/1416 presentation proof");
    int zz = zk_scalar_val(&z, "z", zs), z_0 = zk_scalar_val(&z, "z_0", z0s), tt = zk_scalar_val(&z, "t", t);
    int H_s[MAXN], G_m[MAXN];
    for (int i = 0; i < n; i++) { H_s[i] = G_m[i] = -1; if (at[i].kind == 's') H_s[i] = zk_scalar_val(&z, "m", at[i].sc); }
    int I = zk_point_p(&z, "I", &is->I), C_x_1 = zk_point_p(&z, "C_x_1", &p->C_x_1), C_x_0 = zk_point_p(&z, "C_x_0", &p->C_x_0);
    int G_x_0 = zk_point_p(&z, "G_x_0", &is->G_x0), G_x_1 = zk_point_p(&z, "G_x_1", &is->G_x1);
    int C_y[MAXN], ncy = 0, G_y[MAXN];
    for (int i = 0; i < n; i++) { if (at[i].kind == 'H') continue; C_y[ncy++] = zk_point_p(&z, "C_y", &p->C_y[i]); }
    for (uint32_t i = 0; i < is->ny; i++) G_y[i] = zk_point_p(&z, "G_y", &is->G_y[i]);
    for (int i = 0; i < n; i++) if (at[i].kind == 's') G_m[i] = zk_point_p(&z, "G_m", &is->G_m[i]);
    int Zv = zk_point_p(&z, "Z", &Zp);
    int c;
    c = zk_constrain(&z, Zv); zk_term(&z, c, zz, I);
    c = zk_constrain(&z, C_x_1); zk_term(&z, c, tt, C_x_0); zk_term(&z, c, z_0, G_x_0); zk_term(&z, c, zz, G_x_1);
    for (int i = 0; i < ncy; i++) { /* :267-273 compacted-index loop */
        if (at[i].kind == 'H') continue;
        c = zk_constrain(&z, C_y[i]); zk_term(&z, c, zz, G_y[i]);
        if (at[i].kind == 's') { if (H_s[i] < 0) abort(); zk_term(&z, c, H_s[i], G_m[i]); }
    }
    p->nresp = z.nsc;
    zk_prove_compact(&z, blind, p->c, p->resp);
    int eb = 0;
    for (int i = 0; i < n; i++) {
        switch (at[i].kind) {
        case 'S': p->kinds[i] = K_PS; memcpy(p->rev_scalar[i], at[i].sc, 32); break;
        case 's': p->kinds[i] = K_SS; break;
        case 'P': p->kinds[i] = K_PP; p->rev_point[i] = at[i].pt; break;
        case 'E': p->kinds[i] = K_PP; p->rev_point[i] = at[i].pl.M1; break;
        default:
            p->kinds[i] = K_SP;
            encryption_prove(is, &at[i].pl, i, kp, zs, enc_blind + 6 * eb, &p->enc[p->nenc]);
            eb++; p->nenc++;
        }
    }
}
static void presentation_serialize(const presentation* p, uint8_t (*w)[32]) {
    int k = 0, n = p->n;
    memcpy(w[k++], p->c, 32);
    for (int i = 0; i < p->nresp; i++) memcpy(w[k++], p->resp[i], 32);
    ge_compress(w[k++], &p->C_x_0); ge_compress(w[k++], &p->C_x_1); ge_compress(w[k++], &p->C_V);
    for (int i = 0; i < n; i++) ge_compress(w[k++], &p->C_y[i]);
    for (int i = 0; i < n; i++) { if (p->kinds[i] == K_PS) memcpy(w[k++], p->rev_scalar[i], 32); else if (p->kinds[i] == K_PP) ge_compress(w[k++], &p->rev_point[i]); }
    for (int j = 0; j < p->nenc; j++) {
        const enc_proof* e = &p->enc[j];
        memcpy(w[k++], e->c, 32); for (int i = 0; i < 6; i++) memcpy(w[k++], e->resp[i], 32);
        ge_compress(w[k++], &e->pk); ge_compress(w[k++], &e->E1); ge_compress(w[k++], &e->E2);
        ge_compress(w[k++], &e->C_y_1); ge_compress(w[k++], &e->C_y_2); ge_compress(w[k++], &e->C_y_3); ge_compress(w[k++], &e->C_y_2p);
    }
}

/* ------------------------------------------------------------------ deterministic randomness (same draw order as pyoracle/synth.py) */
typedef struct { uint8_t seed[96]; size_t seedlen; uint32_t ctr; uint8_t buf[64]; size_t have; } shake_rng;
static void rng_init(shake_rng* r, const char* tag, const uint8_t* extra, size_t extralen) {
    size_t k = 0; memcpy(r->seed, "aeonflux-b200/", 14); k = 14;
    size_t tl = strlen(tag); memcpy(r->seed + k, tag, tl); k += tl;
    memcpy(r->seed + k, extra, extralen); k += extralen;
    r->seedlen = k; r->ctr = 0; r->have = 0;
}
static void rng_fill(shake_rng* r, uint8_t* out, size_t n) {
    while (n) {
        if (!r->have) {
            uint8_t in[100]; memcpy(in, r->seed, r->seedlen);
            in[r->seedlen] = (uint8_t)r->ctr; in[r->seedlen + 1] = (uint8_t)(r->ctr >> 8); in[r->seedlen + 2] = (uint8_t)(r->ctr >> 16); in[r->seedlen + 3] = (uint8_t)(r->ctr >> 24);
            shake256(r->buf, 64, in, r->seedlen + 4); r->ctr++; r->have = 64;
        }
        size_t take = n < r->have ? n : r->have;
        memcpy(out, r->buf + (64 - r->have), take); out += take; n -= take; r->have -= take;
    }
}
static void rng_scalar(shake_rng* r, uint8_t s[32]) { uint8_t b[64]; rng_fill(r, b, 64); sc_from_wide(s, b); }
static void rng_point(shake_rng* r, ge* p) { uint8_t b[64]; rng_fill(r, b, 64); ge_from_uniform(p, b); }

/* ================================================================== exported API (ctypes) */
#define API __attribute__((visibility("default")))

API int afxo_sysparams_size(uint32_t n) { return n < 3 ? 32 * (5 + 3 + (int)n + 4) + 4 : 32 * (5 + 2 * (int)n + 4) + 4; }
API int afxo_secret_size(uint32_t n) { return 32 * (5 + (int)n) + 4; }

/* parse SystemParameters::to_bytes (parameters.rs:155-184) || C_W||I || SecretKey::to_bytes (amacs.rs:110-125) */
API afxo_issuer* afxo_issuer_new(const uint8_t* sp, size_t sp_len, const uint8_t* ipub, const uint8_t* sk, size_t sk_len) {
    afxo_issuer* is = (afxo_issuer*)calloc(1, sizeof(afxo_issuer));
    uint32_t n; memcpy(&n, sp, 4);
    if (n == 0 || n > MAXN || sp_len != (size_t)afxo_sysparams_size(n)) { free(is); return NULL; }
    is->n = n; is->ny = n < 3 ? 3 : n;
    const uint8_t* q = sp + 4; int ok = 1;
#define RD(dst) do { ok &= ge_decompress(dst, q); q += 32; } while (0)
    RD(&is->G); RD(&is->G_w); RD(&is->G_wp); RD(&is->G_x0); RD(&is->G_x1);
    for (uint32_t i = 0; i < is->ny; i++) RD(&is->G_y[i]);
    for (uint32_t i = 0; i < n; i++) RD(&is->G_m[i]);
    RD(&is->G_V); RD(&is->G_a); RD(&is->G_a0); RD(&is->G_a1);
    ok &= ge_decompress(&is->C_W, ipub); ok &= ge_decompress(&is->I, ipub + 32);
    if (sk) {
        uint32_t m; memcpy(&m, sk, 4);
        if (m != n || sk_len != (size_t)afxo_secret_size(n)) { free(is); return NULL; }
        const uint8_t* s = sk + 4;
        memcpy(is->w, s, 32); memcpy(is->wp, s + 32, 32); memcpy(is->x0, s + 64, 32); memcpy(is->x1, s + 96, 32);
        for (uint32_t i = 0; i < n; i++) memcpy(is->y[i], s + 128 + 32 * i, 32);
        ok &= ge_decompress(&is->W, s + 128 + 32 * n);
        ok &= sc_is_canonical(is->w) & sc_is_canonical(is->wp) & sc_is_canonical(is->x0) & sc_is_canonical(is->x1);
        for (uint32_t i = 0; i < n; i++) ok &= sc_is_canonical(is->y[i]);
        is->has_secret = 1;
    }
    if (!ok) { free(is); return NULL; }
    return is;
}
API void afxo_issuer_free(afxo_issuer* is) { if (is) { memset(is, 0, sizeof *is); free(is); } }

/* SystemParameters::generate + Issuer::new from the deterministic rng (synth.make_issuer) */
API int afxo_make_issuer(uint32_t n, const char* tag, uint8_t* sp_out, uint8_t* ipub_out, uint8_t* sk_out) {
    if (n == 0 || n > MAXN) return -1;
    shake_rng r; uint8_t n4[4] = {(uint8_t)n, (uint8_t)(n >> 8), (uint8_t)(n >> 16), (uint8_t)(n >> 24)};
    rng_init(&r, tag, n4, 4);
    uint32_t ny = n < 3 ? 3 : n, npts = 4 + ny + n + 4;
    ge pts[4 + 2 * MAXN + 4 + 3];
    for (uint32_t i = 0; i < npts; i++) { uint8_t b[32]; do rng_fill(&r, b, 32); while (!ge_decompress(&pts[i], b)); }
    memcpy(sp_out, n4, 4); uint8_t* q = sp_out + 4;
    ge B; static const uint8_t Bc[32] = {0xe2, 0xf2, 0xae, 0x0a, 0x6a, 0xbc, 0x4e, 0x71, 0xa8, 0x84, 0xa9, 0x61, 0xc5, 0x00, 0x51, 0x5f, 0x58, 0xe3, 0x0b, 0x6a, 0xa5, 0x82, 0xdd, 0x8d, 0xb6, 0xa6, 0x59, 0x45, 0xe0, 0x8d, 0x2d, 0x76};
    ge_decompress(&B, Bc); ge_compress(q, &B); q += 32;
    for (uint32_t i = 0; i < npts; i++) { ge_compress(q, &pts[i]); q += 32; }
    /* order: G_w, G_w', G_x_0, G_x_1, G_y[ny], G_m[n], G_V, G_a, G_a0, G_a1 == wire order after G */
    uint8_t w[32], wp[32], x0[32], x1[32], y[MAXN][32];
    rng_scalar(&r, w); rng_scalar(&r, wp); rng_scalar(&r, x0); rng_scalar(&r, x1);
    for (uint32_t i = 0; i < n; i++) rng_scalar(&r, y[i]);
    ge W, C_W, I, t;
    ge_scalarmult_ct(&W, &pts[0], w);
    ge_scalarmult_ct(&C_W, &pts[0], w); ge_scalarmult_ct(&t, &pts[1], wp); ge_add(&C_W, &C_W, &t);
    I = pts[4 + ny + n];
    ge_scalarmult_ct(&t, &pts[2], x0); ge_sub(&I, &I, &t);
    ge_scalarmult_ct(&t, &pts[3], x1); ge_sub(&I, &I, &t);
    for (uint32_t i = 0; i < n; i++) { ge_scalarmult_ct(&t, &pts[4 + i], y[i]); ge_sub(&I, &I, &t); }
    ge_compress(ipub_out, &C_W); ge_compress(ipub_out + 32, &I);
    memcpy(sk_out, n4, 4); uint8_t* s = sk_out + 4;
    memcpy(s, w, 32); memcpy(s + 32, wp, 32); memcpy(s + 64, x0, 32); memcpy(s + 96, x1, 32);
    for (uint32_t i = 0; i < n; i++) memcpy(s + 128 + 32 * i, y[i], 32);
    ge_compress(s + 128 + 32 * n, &W);
    return 0;
}

/* ---- batch drivers ---- */
typedef struct {
    const afxo_issuer* is; int mode; int n; const uint8_t* kinds; const uint8_t* hide; const char* config;
    uint64_t start, count; int tid, nthreads;
    const uint8_t* in; uint8_t* out; uint8_t* out2; uint8_t* out3; uint8_t* verdicts; uint8_t* tz; uint8_t* tcommit; uint8_t* tchal; int ncommit_max, nchal_max;
    const uint8_t* randomness; double verify_seconds;
} job;

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

static void* worker(void* arg) {
    job* j = (job*)arg; const afxo_issuer* is = j->is; int n = j->n;
    uint64_t per = (j->count + j->nthreads - 1) / j->nthreads, lo = j->tid * per, hi = lo + per > j->count ? j->count : lo + per;
    static __thread presentation p; static __thread vtrace tr;
    if (j->mode == 0) { /* verify presentations: in = [count][W][32] */
        int W = presentation_words(n, j->kinds);
        for (uint64_t i = lo; i < hi; i++) {
            const uint8_t(*w)[32] = (const uint8_t(*)[32])(j->in + i * (uint64_t)W * 32);
            memset(&tr, 0, sizeof tr);
            int v = presentation_parse(&p, n, j->kinds, w);
            if (!v) { double t0 = now_s(); v = presentation_verify(is, &p, &tr); j->verify_seconds += now_s() - t0; }
            j->verdicts[i] = (uint8_t)v;
            if (j->tz) memcpy(j->tz + 32 * i, tr.Z, 32);
            if (j->tcommit) memcpy(j->tcommit + 32 * (uint64_t)j->ncommit_max * i, tr.commit, 32 * (size_t)(tr.ncommit < j->ncommit_max ? tr.ncommit : j->ncommit_max));
            if (j->tchal) memcpy(j->tchal + 32 * (uint64_t)j->nchal_max * i, tr.chal, 32 * (size_t)(tr.nchal < j->nchal_max ? tr.nchal : j->nchal_max));
        }
    } else if (j->mode == 1) { /* synth: out = presentations [count][W][32], out2 = issuances [count][Wi][32] (nullable) */
        int Wi = n + 3 + 1 + n + 5;
        for (uint64_t i = lo; i < hi; i++) {
            uint64_t item = j->start + i; uint8_t i8[8]; for (int k = 0; k < 8; k++) i8[k] = (uint8_t)(item >> (8 * k));
            shake_rng r; rng_init(&r, j->config, i8, 8);
            cred_attr at[MAXN]; attributes a; a.n = n;
            for (int k = 0; k < n; k++) {
                if (j->kinds[k] == 'S') { at[k].kind = 'S'; rng_scalar(&r, at[k].sc); a.kinds[k] = K_PS; memcpy(a.sc[k], at[k].sc, 32); }
                else if (j->kinds[k] == 'P') { at[k].kind = 'P'; rng_point(&r, &at[k].pt); a.kinds[k] = K_PP; a.pt[k] = at[k].pt; }
                else { uint8_t m[30]; at[k].kind = 'E'; rng_fill(&r, m, 30); plaintext_from30(&at[k].pl, m); a.kinds[k] = K_PP; a.pt[k] = at[k].pl.M1; }
            }
            uint8_t t[32], c[32], blind[MAXN + 5][32], resp[MAXN + 5][32]; ge U, V;
            rng_scalar(&r, t); rng_point(&r, &U);
            for (int k = 0; k < n + 5; k++) rng_scalar(&r, blind[k]);
            issue(is, &a, t, &U, (const uint8_t(*)[32])blind, &V, c, resp);
            if (j->out2) {
                uint8_t(*w)[32] = (uint8_t(*)[32])(j->out2 + i * (uint64_t)Wi * 32); int q = 0;
                for (int k = 0; k < n; k++) { if (a.kinds[k] == K_PS) memcpy(w[q++], a.sc[k], 32); else ge_compress(w[q++], &a.pt[k]); }
                memcpy(w[q++], t, 32); ge_compress(w[q++], &U); ge_compress(w[q++], &V); memcpy(w[q++], c, 32);
                for (int k = 0; k < n + 5; k++) memcpy(w[q++], resp[k], 32);
            }
            uint8_t ms[64]; rng_fill(&r, ms, 64);
            sym_keypair kp; int hp = 0, hs = 0;
            for (int k = 0; k < n; k++) if (j->hide[k]) { if (at[k].kind == 'S') { at[k].kind = 's'; hs++; } else if (at[k].kind == 'E') { at[k].kind = 'H'; hp++; } }
            keypair_derive(&kp, is, ms);
            uint8_t zs[32], pb[3 + MAXN][32], eb[6 * MAXN][32];
            uint8_t seeds[1 + 3 + MAXN + 6 * MAXN][64]; int ns = 0;      /* the rng output behind z and every blinding (Scalar::random = 64 bytes mod l) */
            rng_fill(&r, seeds[ns], 64); sc_from_wide(zs, seeds[ns++]);
            for (int k = 0; k < 3 + hs; k++) { rng_fill(&r, seeds[ns], 64); sc_from_wide(pb[k], seeds[ns++]); }
            for (int k = 0; k < 6 * hp; k++) { rng_fill(&r, seeds[ns], 64); sc_from_wide(eb[k], seeds[ns++]); }
            if (j->out3) { /* the flat input of a batch AnonymousCredential::show for this item (include/aeonflux_b200.h, afx_show) */
                int Ws = 3 + 2 * ns + (hp ? 4 : 0); for (int k = 0; k < n; k++) Ws += at[k].kind == 'H' ? 3 : 1;
                uint8_t(*w)[32] = (uint8_t(*)[32])(j->out3 + i * (uint64_t)Ws * 32); int q = 0;
                memcpy(w[q++], t, 32); ge_compress(w[q++], &U); ge_compress(w[q++], &V);
                for (int k = 0; k < n; k++) {
                    if (at[k].kind == 'S' || at[k].kind == 's') memcpy(w[q++], at[k].sc, 32);
                    else if (at[k].kind == 'P') ge_compress(w[q++], &at[k].pt);
                    else if (at[k].kind == 'E') ge_compress(w[q++], &at[k].pl.M1);
                    else { ge_compress(w[q++], &at[k].pl.M1); ge_compress(w[q++], &at[k].pl.M2); memcpy(w[q++], at[k].pl.m3, 32); }
                }
                if (hp) { memcpy(w[q++], kp.a, 32); memcpy(w[q++], kp.a0, 32); memcpy(w[q++], kp.a1, 32); ge_compress(w[q++], &kp.pk); }
                for (int k = 0; k < ns; k++) { memcpy(w[q++], seeds[k], 32); memcpy(w[q++], seeds[k] + 32, 32); }
            }
            if (j->out) {
                presentation_prove(is, at, t, &U, &V, &kp, zs, (const uint8_t(*)[32])pb, (const uint8_t(*)[32])eb, &p);
                int W = presentation_words(n, p.kinds);
                presentation_serialize(&p, (uint8_t(*)[32])(j->out + i * (uint64_t)W * 32));
            }
        }
    } else if (j->mode == 2) { /* verify issuances: in = [count][Wi][32], kinds = K_PS / K_PP */
        int Wi = n + 3 + 1 + n + 5;
        for (uint64_t i = lo; i < hi; i++) {
            const uint8_t(*w)[32] = (const uint8_t(*)[32])(j->in + i * (uint64_t)Wi * 32);
            attributes a; a.n = n; int bad = 0; ge U, V;
            memset(&tr, 0, sizeof tr);
            for (int k = 0; k < n; k++) { a.kinds[k] = j->kinds[k]; if (j->kinds[k] == K_PS) { bad |= !sc_is_canonical(w[k]); memcpy(a.sc[k], w[k], 32); } else bad |= !ge_decompress(&a.pt[k], w[k]); }
            bad |= !sc_is_canonical(w[n]); bad |= !ge_decompress(&U, w[n + 1]); bad |= !ge_decompress(&V, w[n + 2]);
            for (int k = 0; k < n + 6; k++) bad |= !sc_is_canonical(w[n + 3 + k]);
            int v = 1;
            if (!bad) { double t0 = now_s(); v = issuance_verify(is, &a, w[n], &U, &V, w[n + 3], w + n + 4, &tr); j->verify_seconds += now_s() - t0; }
            j->verdicts[i] = (uint8_t)v;
            if (j->tcommit) memcpy(j->tcommit + 32 * 3 * i, tr.commit, 32 * (size_t)(tr.ncommit < 3 ? tr.ncommit : 3));
            if (j->tchal && tr.nchal) memcpy(j->tchal + 32 * i, tr.chal, 32);
        }
    } else if (j->mode == 3) { /* issue with supplied randomness: in = [count][n][32] attrs; randomness = [count][(2+n+5)][64]; out = [count][(3+1+n+5)][32] */
        int Wr = 2 + n + 5, Wo = 3 + 1 + n + 5;
        for (uint64_t i = lo; i < hi; i++) {
            const uint8_t(*w)[32] = (const uint8_t(*)[32])(j->in + i * (uint64_t)n * 32);
            const uint8_t* rnd = j->randomness + i * (uint64_t)Wr * 64;
            uint8_t(*o)[32] = (uint8_t(*)[32])(j->out + i * (uint64_t)Wo * 32);
            attributes a; a.n = n; int bad = 0;
            for (int k = 0; k < n; k++) { a.kinds[k] = j->kinds[k]; if (j->kinds[k] == K_PS) { bad |= !sc_is_canonical(w[k]); memcpy(a.sc[k], w[k], 32); } else bad |= !ge_decompress(&a.pt[k], w[k]); }
            if (bad) { j->verdicts[i] = 1; memset(o, 0, (size_t)Wo * 32); continue; }
            uint8_t t[32], blind[MAXN + 5][32]; ge U, V;
            sc_from_wide(t, rnd); ge_from_uniform(&U, rnd + 64);
            for (int k = 0; k < n + 5; k++) sc_from_wide(blind[k], rnd + 128 + 64 * k);
            double t0 = now_s();
            issue(is, &a, t, &U, (const uint8_t(*)[32])blind, &V, o[3], o + 4);
            j->verify_seconds += now_s() - t0;
            memcpy(o[0], t, 32); ge_compress(o[1], &U); ge_compress(o[2], &V);
            j->verdicts[i] = 0;
        }
    }
    return NULL;
}

static double run_jobs(job* proto, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    job js[256]; pthread_t th[256];
    for (int t = 0; t < threads; t++) { js[t] = *proto; js[t].tid = t; js[t].nthreads = threads; js[t].verify_seconds = 0; }
    for (int t = 1; t < threads; t++) pthread_create(&th[t], NULL, worker, &js[t]);
    worker(&js[0]);
    for (int t = 1; t < threads; t++) pthread_join(th[t], NULL);
    double mx = 0; for (int t = 0; t < threads; t++) if (js[t].verify_seconds > mx) mx = js[t].verify_seconds;
    return mx;
}

/* Issuer::verify over a batch.  items = [count][W][32] (item-major).  Returns max-over-threads seconds spent inside
 * presentation_verify (parsing the flat words into points is excluded: the Rust types already hold points), or <0. */
API double afxo_verify_presentations(const afxo_issuer* is, const uint8_t* kinds, int n, const uint8_t* items, uint64_t count, int threads,
                                     uint8_t* verdicts, uint8_t* trace_Z, uint8_t* trace_commit, int ncommit_max, uint8_t* trace_chal, int nchal_max) {
    if (!is || !is->has_secret || (uint32_t)n != is->n) return -1;
    job j; memset(&j, 0, sizeof j);
    j.is = is; j.mode = 0; j.n = n; j.kinds = kinds; j.in = items; j.count = count; j.verdicts = verdicts;
    j.tz = trace_Z; j.tcommit = trace_commit; j.ncommit_max = ncommit_max; j.tchal = trace_chal; j.nchal_max = nchal_max;
    return run_jobs(&j, threads);
}
/* request_kinds: 'S' scalar, 'P' point, 'E' 30-byte plaintext; hide[k] != 0 => hidden at presentation */
API double afxo_synth(const afxo_issuer* is, const uint8_t* request_kinds, const uint8_t* hide, int n, const char* config, uint64_t start, uint64_t count,
                      int threads, uint8_t* presentations_out, uint8_t* issuances_out, uint8_t* show_inputs_out) {
    if (!is || !is->has_secret || (uint32_t)n != is->n) return -1;
    job j; memset(&j, 0, sizeof j);
    j.is = is; j.mode = 1; j.n = n; j.kinds = request_kinds; j.hide = hide; j.config = config; j.start = start; j.count = count;
    j.out = presentations_out; j.out2 = issuances_out; j.out3 = show_inputs_out;
    double t0 = now_s(); run_jobs(&j, threads); return now_s() - t0;
}
API double afxo_verify_issuances(const afxo_issuer* is, const uint8_t* kinds, int n, const uint8_t* items, uint64_t count, int threads,
                                 uint8_t* verdicts, uint8_t* trace_commit, uint8_t* trace_chal) {
    if (!is || (uint32_t)n != is->n) return -1;
    job j; memset(&j, 0, sizeof j);
    j.is = is; j.mode = 2; j.n = n; j.kinds = kinds; j.in = items; j.count = count; j.verdicts = verdicts; j.tcommit = trace_commit; j.tchal = trace_chal;
    return run_jobs(&j, threads);
}
API double afxo_issue(const afxo_issuer* is, const uint8_t* kinds, int n, const uint8_t* attrs, const uint8_t* randomness, uint64_t count, int threads,
                      uint8_t* out, uint8_t* status) {
    if (!is || !is->has_secret || (uint32_t)n != is->n) return -1;
    job j; memset(&j, 0, sizeof j);
    j.is = is; j.mode = 3; j.n = n; j.kinds = kinds; j.in = attrs; j.randomness = randomness; j.count = count; j.out = out; j.verdicts = status;
    return run_jobs(&j, threads);
}

/* ---- primitive hooks for cross-checking against the Python oracle / libsodium ---- */
API int afxo_decompress_compress(const uint8_t in[32], uint8_t out[32]) { ge p; if (!ge_decompress(&p, in)) return 0; ge_compress(out, &p); return 1; }
API void afxo_from_uniform(const uint8_t in[64], uint8_t out[32]) { ge p; ge_from_uniform(&p, in); ge_compress(out, &p); }
API int afxo_scalarmult(const uint8_t s[32], const uint8_t pt[32], uint8_t out[32], int vartime) {
    ge p, r; if (!ge_decompress(&p, pt)) return 0;
    if (vartime) ge_msm_vartime(&r, 1, (const uint8_t(*)[32])s, &p); else ge_scalarmult_ct(&r, &p, s);
    ge_compress(out, &r); return 1;
}
API void afxo_sc_from_wide(const uint8_t in[64], uint8_t out[32]) { sc_from_wide(out, in); }
API void afxo_sc_muladd(const uint8_t a[32], const uint8_t b[32], const uint8_t c[32], uint8_t out[32]) { sc_muladd(out, a, b, c); }
API void afxo_sha512(const uint8_t* in, size_t n, uint8_t out[64]) { sha512(out, in, n); }
API void afxo_shake256(const uint8_t* in, size_t n, uint8_t* out, size_t outlen) { shake256(out, outlen, in, n); }
API void afxo_merlin_kat(uint8_t out[32]) { strobe s; merlin_init(&s, "test protocol"); merlin_append(&s, "some label", "some data", 9); merlin_challenge(&s, "challenge", out, 32); }

__attribute__((constructor)) static void afxo_init(void) { uint64_t st[25] = {0}; keccak_f1600(st); }
