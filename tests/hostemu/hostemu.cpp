// TEST-ONLY host emulation of the engine: compiles the same __host__ __device__ stage bodies (engine.cuh) and the same
// C-ABI implementation (api_impl.inc) with plain loops in place of CUDA kernels, so the shape compiler, the marshaling
// and the stage logic can be debugged against the oracle on a machine without a GPU.  It is built by tests/ only, is
// never loaded by the aeonflux_b200 package, and exercises none of the PTX paths -- GPU parity is proven separately by
// the -m gpu tests through the real library.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "../../aeonflux_b200/csrc/engine.cuh"

using namespace afx;
typedef void* be_stream;
#define AFX_BACKEND_NAME "host-emulation (tests only)"

static int be_set_device(int) { return 0; }
static int be_malloc(void** p, size_t n) { *p = std::calloc(1, n ? n : 16); return *p == nullptr; }
static void be_free(void* p) { std::free(p); }
static int be_h2d(void* d, const void* h, size_t n, be_stream) { std::memcpy(d, h, n); return 0; }
static int be_d2h(void* h, const void* d, size_t n, be_stream) { std::memcpy(h, d, n); return 0; }
static int be_d2d(void* d, const void* s, size_t n, be_stream) { std::memcpy(d, s, n); return 0; }
static int be_os_random(void* p, size_t n) { FILE* f = std::fopen("/dev/urandom", "rb"); if (!f) return 1; size_t k = std::fread(p, 1, n, f); std::fclose(f); return k != n; }
static int be_bind_thread_to_device(int) { return 1; }
static size_t be_split_copy_min_items() { return 2; }
static int be_memset(void* d, int v, size_t n, be_stream) { std::memset(d, v, n); return 0; }
static int be_sync(be_stream) { return 0; }
static int be_check_launch() { return 0; }
typedef int be_event;
static void be_event_create(be_event*) {}
static void be_event_destroy(be_event) {}
static void be_event_record(be_event, be_stream) {}
static int be_event_sync(be_event) { return 0; }
static float be_event_elapsed(be_event, be_event) { return 0.f; }
static int be_stream_create(be_stream* s) { *s = nullptr; return 0; }
static void be_stream_destroy(be_stream) {}
static void be_stream_wait(be_stream, be_event) {}
static int be_host_alloc(void** p, size_t n) { *p = std::calloc(1, n ? n : 16); return *p == nullptr; }
static void be_host_free(void* p) { std::free(p); }

static void be_launch_scalar_check(const Workspace& ws, const u16* f, u32 nf, be_stream) {
    for (u32 k = 0; k < nf; k++) for (u32 i = ws.e_lo; i < ws.e_hi; i++) scalar_check_job(ws, f[k], i);
}
static void be_launch_points(const Workspace& ws, const PointJob* jobs, u32 nj, be_stream) {
    for (u32 k = 0; k < nj; k++) for (u32 i = ws.e_lo; i < ws.e_hi; i++) points_job(ws, jobs[k], i);
}
static u32 be_launch_ladders(const Workspace& ws, const AmacDesc* amac, u32 nps, const MsmDesc* msms, const u32* idx, u32 nidx, u32, u32 max_terms, u32,
                             int, u32*, u32, u32, u32*, be_stream) {
    std::vector<u32> scratch((size_t)(max_terms > nps ? max_terms : nps) * 8 + 8);
    if (amac) for (u32 i = 0; i < ws.count; i++) amac_job(ws, *amac, i, scratch.data(), 1);
    for (u32 k = 0; k < nidx; k++) {
        const MsmDesc& d = msms[idx[k]];
        CtabResolver r{ws.ctabs, &d};
        for (u32 i = 0; i < ws.count; i++) msm_job(ws, d, i, scratch.data(), 1, r);
    }
    return 1;
}
static void be_launch_ladders_parts(const Workspace& ws, const AmacDesc* amac, const AmacPart* parts, u32 n_parts, u32 part_terms, u32* parts_out,
                                    const MsmDesc* msms, const u32* idx, u32 nidx, u32 max_terms, be_stream) {
    std::vector<u32> scratch((size_t)(max_terms > part_terms ? max_terms : part_terms) * 8 + 8);
    for (u32 g = 0; g < n_parts; g++)
        for (u32 i = 0; i < ws.count; i++) amac_part_job(ws, *amac, parts[g], parts_out + (size_t)g * ws.count * 32, i, scratch.data(), 1);
    for (u32 k = 0; k < nidx; k++) {
        const MsmDesc& d = msms[idx[k]];
        CtabResolver r{ws.ctabs, &d};
        for (u32 i = 0; i < ws.count; i++) msm_job(ws, d, i, scratch.data(), 1, r);
    }
}
static void be_launch_amac_combine(const Workspace& ws, const AmacDesc* amac, const u32* parts_out, u32 n_parts, be_stream) {
    for (u32 i = 0; i < ws.count; i++) amac_combine_job(ws, *amac, parts_out, n_parts, i);
}
static void be_launch_msm_ct(const Workspace& ws, const MsmDesc* msms, const u32* idx, u32 nidx, u32 max_terms, be_stream) {
    std::vector<u32> scratch((size_t)max_terms * 8);
    for (u32 k = 0; k < nidx; k++) for (u32 i = 0; i < ws.count; i++) msm_ct_job(ws, msms[idx[k]], i, scratch.data(), 1);
}
static void be_launch_derive(const Workspace& ws, const DeriveOp* d, u32 nd, be_stream) {
    for (u32 i = 0; i < ws.count; i++) derive_program_job(ws, d, nd, i);
}
static void be_launch_out_words(const Workspace& ws, const OutWord* d, u32 nwords, u32* out, be_stream) {
    for (u32 k = 0; k < nwords; k++) for (u32 i = 0; i < ws.count; i++) out_word_job(ws, d[k], k, i, out);
}
static void be_launch_transcript(const Workspace& ws, const TxDesc* txs, u32 ntx, be_stream) {
    for (u32 k = 0; k < ntx; k++) for (u32 i = 0; i < ws.count; i++) transcript_job(ws, txs[k], i);
}
static void be_launch_ztable(const Workspace& ws, const AmacDesc* d, be_stream) { for (u32 i = 0; i < ws.count; i++) ztable_job(ws, *d, i); }
static void be_launch_commit_compare(const Workspace& ws, const CmpPair* pairs, u32 npairs, be_stream) {
    for (u32 k = 0; k < npairs; k++) for (u32 i = 0; i < ws.count; i++) commit_compare_job(ws, pairs[k], i);
}
static u32 be_launch_rlc(const Workspace& ws, const RlcDesc* d, u32 ncterms, const RlcBuffers& rb, be_stream, be_event* = nullptr) {
    std::memset(rb.hist, 0, (size_t)rb.nwin * (rb.nb + 1) * 4);
    for (u32 i = 0; i < rb.cnt; i++) rlc_scalars_job(ws, *d, rb, rb.lo + i);
    for (u32 t = 0; t < ncterms; t++) {
        u32 acc[9] = {0};
        for (u32 i = 0; i < rb.cnt; i++) rlc_add288(acc, rb.cterm + ((size_t)t * rb.cnt + i) * 8);
        sc r = rlc_reduce288(acc);
        for (int i = 0; i < 8; i++) rb.csum[8 * t + i] = r.v[i];
    }
    for (u32 n = 0; n < rb.N; n++) rlc_digits_job(*d, rb, n);
    for (u32 w = 0; w < rb.nwin; w++) rlc_scan_job(rb, w);
    for (u32 n = 0; n < rb.N; n++) rlc_scatter_job(rb, n);
    for (u32 w = 0; w < rb.nwin; w++) for (u32 b = 1; b <= rb.nb; b++) rlc_bucket_job(ws, *d, rb, w, b);
    for (u32 w = 0; w < rb.nwin; w++) store_ge(rb.wsum + (size_t)w * 32, rlc_segment_job(rb, w, 1, rb.nb));
    rlc_final_job(ws, *d, rb);
    return 8;
}
static void be_launch_verdict(const Workspace& ws, uint8_t* v, be_stream) { for (u32 i = 0; i < ws.count; i++) v[i] = ws.status[i] != 0; }
// wide (radix-2^16) tables only on request: building them on one core costs seconds per issuer
static size_t be_ctab16_budget() { const char* e = std::getenv("AFX_HOSTEMU_CTAB16"); return e && e[0] == '1' ? (size_t)72 << 20 : 0; }
static void be_launch_ctab_setup(const u32* enc, u32 ncp, u32* ctabs, u32* encneg, u32* bad, int bits, be_stream) {
    const int CTAB_N = 1 << bits;
    // same entries as ctab_entry_job (m * P in affine Niels form, m = 1..CTAB_ENTRIES), walked as running sums with one shared
    // inversion per base (Montgomery's trick) instead of a ladder and an inversion per entry: the emulation runs on one core
    std::vector<ge> mult(CTAB_N);
    std::vector<fe> pref(CTAB_N);
    for (u32 b = 0; b < ncp; b++) {
        ge p; u32 ok = ge_decompress(p, enc + 8 * b);
        if (!ok) *bad |= 1;
        pniels pn = ge_to_pniels(p);
        mult[0] = p;
        for (int m = 1; m < CTAB_N; m++) mult[m] = ge_add_pn(mult[m - 1], pn, true);
        pref[0] = mult[0].Z;
        for (int m = 1; m < CTAB_N; m++) pref[m] = fe_mul(pref[m - 1], mult[m].Z);
        fe z = pref[CTAB_N - 1];
        fe inv = fe_mul(fe_sqn(fe_pow_p58(z), 3), fe_mul(fe_sq(z), z));     // z^(p-2)
        for (int m = CTAB_N - 1; m >= 0; m--) {
            fe zinv = m ? fe_mul(inv, pref[m - 1]) : inv;
            if (m) inv = fe_mul(inv, mult[m].Z);
            fe x = fe_mul(mult[m].X, zinv), y = fe_mul(mult[m].Y, zinv);
            u32* out = ctabs + ((size_t)b * CTAB_N + m) * 24;
            store_fe(out, fe_add(y, x)); store_fe(out + 8, fe_sub(y, x)); store_fe(out + 16, fe_mul(fe_mul(x, y), FE_D2()));
        }
        ge_compress(encneg + 8 * b, ge_neg(p));
    }
}
static void be_launch_comb_setup(const u32* enc, u32 ncp, u32* comb, be_stream) {
    // same entries as comb_entry_job, walked window by window (16 * previous) instead of from scratch: the emulation runs on one core
    for (u32 b = 0; b < ncp; b++) {
        ge p; ge_decompress(p, enc + 8 * b);
        pniels pn = ge_to_pniels(p);
        ge mult[COMB_ENTRIES];
        mult[0] = p;
        for (int e = 1; e < COMB_ENTRIES; e++) mult[e] = ge_add_pn(mult[e - 1], pn, true);
        for (int i = 0; i < COMB_WINDOWS; i++) {
            for (int e = 0; e < COMB_ENTRIES; e++) {
                ge& a = mult[e];
                fe z = a.Z;
                fe t = fe_sqn(fe_pow_p58(z), 3);
                fe zinv = fe_mul(t, fe_mul(fe_sq(z), z));
                fe x = fe_mul(a.X, zinv), y = fe_mul(a.Y, zinv);
                u32* out = comb + (((size_t)b * COMB_WINDOWS + i) * COMB_ENTRIES + e) * 24;
                store_fe(out, fe_add(y, x)); store_fe(out + 8, fe_sub(y, x)); store_fe(out + 16, fe_mul(fe_mul(x, y), FE_D2()));
                for (int k = 0; k < 4; k++) a = ge_dbl(a, true);
            }
        }
    }
}
static void be_launch_link_setup(const u32* enc_gy, u32 ny, u32* out, be_stream) { for (u32 i = 1; i < ny; i++) link_entry_job(enc_gy, i, out + 8 * i); }
static void be_launch_primitive(u32 op, const u32* in, u32* out, u32* flags, u32 count, be_stream) {
    for (u32 i = 0; i < count; i++) primitive_job(op, in, out, flags, i);
}
static void be_launch_secret_setup(const u32* secsc, u32 nsec, u32* secdig, const u32* Wenc, u32* W, u32* bad, be_stream) {
    for (u32 t = 0; t < nsec; t++) { sc s = sc_from_words(secsc + 8 * t); if (!sc_is_canonical(s)) *bad |= 2; sc_recode16(secdig + 8 * t, s); }
    ge p; if (!ge_decompress(p, Wenc)) *bad |= 1;
    store_pniels(W, ge_to_pniels(p));
}

#include "../../aeonflux_b200/csrc/api_impl.inc"
