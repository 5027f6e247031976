// C++ host-side parity driver: the reference's README flow (README.md:44-117; the same sequence as the tests in
// src/nizk/presentation.rs:460-638) in batch form through include/aeonflux_b200.hpp:
//     Issuer::from_bytes -> issue_batch -> CredentialIssuance::verify_batch -> show_batch -> Issuer::verify_batch (+ wire form)
// against expectations computed by the CPU oracle and handed over in a fixture file written by tests/test_cpp_host.py.
// Linked against libaeonflux_b200.so (GPU, `-m gpu`) or the test-only host emulation of the same C ABI (CPU tests).
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>

#include "../../include/aeonflux_b200.hpp"

using namespace aeonflux;

static std::vector<uint8_t> rd(std::ifstream& f, size_t n) { std::vector<uint8_t> v(n); f.read((char*)v.data(), (std::streamsize)n); if (!f) { std::cerr << "short fixture\n"; std::exit(2); } return v; }
static uint32_t rd32(std::ifstream& f) { auto v = rd(f, 4); uint32_t x; std::memcpy(&x, v.data(), 4); return x; }
#define CHECK(cond) do { if (!(cond)) { std::fprintf(stderr, "CHECK failed at %s:%d: %s\n", __FILE__, __LINE__, #cond); return 1; } } while (0)

int main(int argc, char** argv) {
    if (argc < 2) { std::cerr << "usage: host_parity <fixture> [max_batch]\n"; return 2; }
    size_t max_batch = argc > 2 ? (size_t)std::atoi(argv[2]) : 64;
    std::ifstream f(argv[1], std::ios::binary);
    uint32_t n = rd32(f), count = rd32(f);
    auto issuer_bytes = rd(f, rd32(f));                       // Issuer::to_bytes layout
    auto request_kinds = rd(f, n);                            // 0 scalar / 2 point
    auto attrs = rd(f, (size_t)count * n * 32);               // [count][n][32]
    auto rnd = rd(f, (size_t)count * (n + 7) * 64);           // [count][n+7][64] rng bytes of Issuer::issue
    auto expect_issued = rd(f, (size_t)count * (n + 9) * 32); // [count][n+9][32]  t, U, V, c, responses
    auto pres_kinds = rd(f, n);
    uint32_t Ws = rd32(f), W = rd32(f);
    auto show_in = rd(f, (size_t)count * Ws * 32);            // [count][Ws][32]
    auto expect_pres = rd(f, (size_t)count * W * 32);         // the presentations show must produce
    auto corrupted = rd(f, (size_t)count * W * 32);           // the same with some items corrupted
    auto expect_verdicts = rd(f, count);

    Issuer issuer = Issuer::from_bytes(issuer_bytes, 0, max_batch);
    CHECK(issuer.to_bytes() == issuer_bytes);
    CHECK(issuer.number_of_attributes() == n);
    // the user only has the public parameters
    size_t a = issuer_bytes.size() - 64 - (32 * (size_t)(5 + n) + 4);
    Issuer user(std::vector<uint8_t>(issuer_bytes.begin(), issuer_bytes.begin() + a), std::vector<uint8_t>(issuer_bytes.begin() + a, issuer_bytes.begin() + a + 64), {}, 0, max_batch);

    // Issuer::issue
    std::vector<uint8_t> req_items((size_t)count * (3 * n + 14) * 32);
    for (uint32_t i = 0; i < count; i++) {
        std::memcpy(&req_items[(size_t)i * (3 * n + 14) * 32], &attrs[(size_t)i * n * 32], (size_t)n * 32);
        std::memcpy(&req_items[((size_t)i * (3 * n + 14) + n) * 32], &rnd[(size_t)i * (n + 7) * 64], (size_t)(n + 7) * 64);
    }
    RequestBatch req = Batch::from_items(request_kinds, req_items.data(), count, 3 * n + 14);
    auto issued = issuer.issue_batch(req);
    for (auto& r : issued.second) CHECK(r.is_ok());
    for (uint32_t i = 0; i < count; i++)
        for (uint32_t w = 0; w < n + 9; w++)
            CHECK(std::memcmp(&issued.first.fields[n + w][32 * i], &expect_issued[((size_t)i * (n + 9) + w) * 32], 32) == 0);
    // CredentialIssuance::verify on the user's side
    for (auto& r : CredentialIssuance::verify_batch(user, issued.first)) CHECK(r.is_ok());
    {   // a tampered response must fail with VerificationFailure (the only error this path can return, src/errors.rs:152-156)
        IssuanceBatch bad = issued.first; bad.fields[n + 4][0] ^= 1;
        auto r = CredentialIssuance::verify_batch(user, bad);
        CHECK(r[0].is_err() && r[0].error == CredentialError::VerificationFailure);
        for (uint32_t i = 1; i < count; i++) CHECK(r[i].is_ok());
    }
    {   // Amac::tag rejects a request whose attribute count is not the issuer's (src/amacs.rs:285-287 -> MacCreation)
        RequestBatch shorter; shorter.kinds.assign(n + 1, 0); shorter.fields.assign(3 * (n + 1) + 14, std::vector<uint8_t>(32));
        auto r = issuer.issue_batch(shorter);
        CHECK(r.second.size() == 1 && r.second[0].is_err() && r.second[0].error == CredentialError::MacCreation);
    }
    // AnonymousCredential::show on the user's side
    ShowBatch sb = Batch::from_items(pres_kinds, show_in.data(), count, Ws);
    auto shown = user.show_batch(sb);
    for (auto& r : shown.second) CHECK(r.is_ok());
    CHECK(shown.first.fields.size() == W);
    for (uint32_t i = 0; i < count; i++)
        for (uint32_t w = 0; w < W; w++) CHECK(std::memcmp(&shown.first.fields[w][32 * i], &expect_pres[((size_t)i * W + w) * 32], 32) == 0);
    // Issuer::verify
    for (auto& r : issuer.verify_batch(shown.first)) CHECK(r.is_ok());
    PresentationBatch cb = Batch::from_items(pres_kinds, corrupted.data(), count, W);
    auto v = issuer.verify_batch(cb);
    auto vw = issuer.verify_wire(pres_kinds, corrupted.data(), count);
    size_t rejected = 0;
    for (uint32_t i = 0; i < count; i++) {
        CHECK(v[i].is_ok() == (expect_verdicts[i] == 0));
        CHECK(vw[i].is_ok() == v[i].is_ok());
        if (v[i].is_err()) { CHECK(v[i].error == CredentialError::VerificationFailure); rejected++; }
    }
    // one process, two contexts, two host threads (the single-process form of the multi-GPU sharding)
    {
        MultiGpuIssuer multi(std::vector<uint8_t>(issuer_bytes.begin(), issuer_bytes.begin() + a), std::vector<uint8_t>(issuer_bytes.begin() + a, issuer_bytes.begin() + a + 64),
                             std::vector<uint8_t>(issuer_bytes.begin() + a + 64, issuer_bytes.end()), {0, 0}, max_batch);
        auto vm = multi.verify_batch(cb);
        CHECK(vm.size() == count && multi.devices() == 2);
        for (uint32_t i = 0; i < count; i++) CHECK(vm[i].is_ok() == v[i].is_ok());
    }
    try { user.verify_batch(cb); CHECK(!"a context without the issuer key must not verify presentations"); } catch (const Error& e) { CHECK(e.code == AFX_ERR_NO_SECRET); }
    std::printf("host_parity ok: %u items, %zu rejected, %s\n", count, rejected, afx_version());
    return 0;
}
