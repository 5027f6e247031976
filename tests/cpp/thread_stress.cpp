// Several host threads on ONE context through the C ABI (include/aeonflux_b200.h: "calls on one context serialise on an internal
// mutex"; a synchronous call while submissions are outstanding returns AFX_ERR_ARG).  Phase 1: four threads run synchronous
// item-major verify calls (presentations and issuances, different slices) concurrently -- every verdict must be the expected one.
// Phase 2: one thread streams submit / submit / wait / wait while two others keep issuing synchronous calls, each of which must
// either return the right verdicts or AFX_ERR_ARG -- never a wrong verdict.  Same fixture as host_parity.cpp.  Built by
// tests/test_cpp_host.py against the test-only emulation (with ThreadSanitizer when the runtime is there) and, under -m gpu,
// against the CUDA library.
#include <atomic>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iostream>
#include <thread>
#include <vector>

#include "../../include/aeonflux_b200.h"

#define CHECK(c) do { if (!(c)) { std::fprintf(stderr, "CHECK failed at line %d: %s\n", __LINE__, #c); std::exit(1); } } while (0)
static std::vector<uint8_t> rd(std::ifstream& f, size_t n) { std::vector<uint8_t> v(n); f.read((char*)v.data(), (std::streamsize)n); if (!f) { std::cerr << "short fixture\n"; std::exit(2); } return v; }
static uint32_t rd32(std::ifstream& f) { auto v = rd(f, 4); uint32_t x; std::memcpy(&x, v.data(), 4); return x; }

int main(int argc, char** argv) {
    if (argc < 2) { std::cerr << "usage: thread_stress <fixture> [max_batch] [rounds]\n"; return 2; }
    const size_t max_batch = argc > 2 ? (size_t)std::atoi(argv[2]) : 64;
    const int rounds = argc > 3 ? std::atoi(argv[3]) : 3;
    std::ifstream f(argv[1], std::ios::binary);
    const uint32_t n = rd32(f), count = rd32(f);
    auto issuer_bytes = rd(f, rd32(f));
    auto request_kinds = rd(f, n);
    auto attrs = rd(f, (size_t)count * n * 32);
    rd(f, (size_t)count * (n + 7) * 64);
    auto issued = rd(f, (size_t)count * (n + 9) * 32);
    auto pres_kinds = rd(f, n);
    const uint32_t Ws = rd32(f), W = rd32(f);
    rd(f, (size_t)count * Ws * 32);
    rd(f, (size_t)count * W * 32);
    auto corrupted = rd(f, (size_t)count * W * 32);
    auto expect = rd(f, count);
    const uint32_t Wi = 2 * n + 9;
    std::vector<uint8_t> issuances((size_t)count * Wi * 32);             // item-major issuances: attribute[n], t, U, V, challenge, responses
    for (uint32_t i = 0; i < count; i++) {
        std::memcpy(&issuances[(size_t)i * Wi * 32], &attrs[(size_t)i * n * 32], (size_t)n * 32);
        std::memcpy(&issuances[((size_t)i * Wi + n) * 32], &issued[(size_t)i * (n + 9) * 32], (size_t)(n + 9) * 32);
    }
    const size_t sk_len = 32 * (size_t)(5 + n) + 4, sp_len = issuer_bytes.size() - 64 - sk_len;
    afx_ctx* ctx = nullptr;
    CHECK(afx_ctx_create(issuer_bytes.data(), sp_len, issuer_bytes.data() + sp_len, issuer_bytes.data() + sp_len + 64, sk_len, 0, max_batch, &ctx) == AFX_OK);

    std::atomic<long> ok_calls{0}, refused{0};
    auto sync_worker = [&](int t, int T, bool may_be_refused) {
        const uint32_t lo = count * (uint32_t)t / (uint32_t)T, hi = count * (uint32_t)(t + 1) / (uint32_t)T;
        std::vector<uint8_t> v(hi - lo);
        for (int r = 0; r < rounds; r++) {
            std::fill(v.begin(), v.end(), 9);
            int rc = (r + t) % 2 == 0
                ? afx_verify_presentations_wire(ctx, (uint16_t)n, pres_kinds.data(), hi - lo, &corrupted[(size_t)lo * W * 32], v.data())
                : afx_verify_issuances_wire(ctx, (uint16_t)n, request_kinds.data(), hi - lo, &issuances[(size_t)lo * Wi * 32], v.data());
            if (rc == AFX_ERR_ARG && may_be_refused) { refused++; std::this_thread::yield(); continue; }
            CHECK(rc == AFX_OK);
            for (uint32_t i = lo; i < hi; i++) CHECK(v[i - lo] == ((r + t) % 2 == 0 ? expect[i] : 0));
            ok_calls++;
        }
    };
    {   // phase 1
        std::vector<std::thread> th;
        for (int t = 0; t < 4; t++) th.emplace_back(sync_worker, t, 4, false);
        for (auto& x : th) x.join();
        CHECK(ok_calls == 4 * rounds);
    }
    {   // phase 2
        const size_t half = count / 2 < max_batch ? count / 2 : max_batch;
        std::thread streamer([&] {
            std::vector<uint8_t> va(half), vb(half);
            for (int r = 0; r < 2 * rounds; r++) {
                uint64_t ta = 0, tb = 0;
                int rc;
                while ((rc = afx_verify_presentations_wire_submit(ctx, (uint16_t)n, pres_kinds.data(), half, corrupted.data(), va.data(), &ta)) == AFX_ERR_ARG) std::this_thread::yield();
                CHECK(rc == AFX_OK);
                while ((rc = afx_verify_presentations_wire_submit(ctx, (uint16_t)n, pres_kinds.data(), half, &corrupted[half * W * 32], vb.data(), &tb)) == AFX_ERR_ARG) std::this_thread::yield();
                CHECK(rc == AFX_OK);
                CHECK(afx_wait(ctx, ta) == AFX_OK && afx_wait(ctx, tb) == AFX_OK);
                for (size_t i = 0; i < half; i++) CHECK(va[i] == expect[i] && vb[i] == expect[half + i]);
            }
        });
        std::thread a(sync_worker, 0, 2, true), b(sync_worker, 1, 2, true);
        streamer.join(); a.join(); b.join();
    }
    afx_ctx_destroy(ctx);
    std::printf("thread_stress ok: %ld synchronous calls, %ld refused while submissions were outstanding, %s\n", ok_calls.load(), refused.load(), afx_version());
    return 0;
}
