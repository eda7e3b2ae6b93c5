// test_host_pool.cpp — CPU-only unit test of the host thread pool behind the *_host gathers (q-learning_b200/csrc/host_pool.cpp):
// plain widening (cached and non-temporal paths, unaligned destinations), the streamed form with a producer thread that raises the
// arrival flags in a shuffled order and with delays, pieces without a destination, and a producer that dies half way (the caller
// must come back and report the missing pieces instead of spinning forever).
//   g++ -std=c++17 -O2 -pthread tests/cpp/test_host_pool.cpp q-learning_b200/csrc/host_pool.cpp -o tests/cpp/test_host_pool
#include <algorithm>
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <functional>
#include <random>
#include <thread>
#include <vector>

#include "../../q-learning_b200/csrc/host_pool.h"

static int fails = 0;
#define CHECK(c) do { if (!(c)) { std::printf("FAIL %s:%d %s\n", __FILE__, __LINE__, #c); ++fails; } } while (0)

struct Producer { std::atomic<bool> alive{true}; };
static bool still_running(void* p) { return static_cast<Producer*>(p)->alive.load(); }

int main() {
    std::mt19937 rng(7);
    // 1. plain widening: small (single thread), pooled, and large enough for the non-temporal path, at shifted destinations
    for (size_t n : {size_t(1), size_t(1000), size_t(3) * 28224 * 8, (size_t(9) << 20) + 5}) {
        std::vector<uint8_t> s(n);
        for (auto& v : s) v = (uint8_t)rng();
        for (int off = 0; off < 3; ++off) {
            std::vector<float> d(n + 4, -1.0f);
            qlc_host::widen_u8_f32(s.data(), d.data() + off, n);
            bool ok = true;
            for (size_t i = 0; i < n; ++i) ok &= d[i + off] == (float)s[i];
            CHECK(ok);
            CHECK(d[n + off] == -1.0f);
        }
    }
    // 2. streamed: 96 pieces in groups of 2, flags raised by a producer in a shuffled order with pauses; every third group has no destination
    for (int round = 0; round < 20; ++round) {
        const size_t n_pieces = 96, per = 7056 * 2;
        std::vector<uint8_t> src(n_pieces * per);
        for (auto& v : src) v = (uint8_t)rng();
        std::vector<float> dst(n_pieces * per, -2.0f);
        std::vector<uint32_t> flags(n_pieces, (uint32_t)round);          // stale values of the previous call
        std::vector<qlc_host::StreamPiece> pieces(n_pieces);
        for (size_t i = 0; i < n_pieces; ++i) {
            const bool skip = (i / 2) % 3 == 2;
            pieces[i] = qlc_host::StreamPiece{src.data() + i * per, skip ? nullptr : dst.data() + i * per, (uint32_t)per};
        }
        const uint32_t value = (uint32_t)round + 1;
        Producer prod;
        std::vector<size_t> order(n_pieces);
        for (size_t i = 0; i < n_pieces; ++i) order[i] = i;
        std::shuffle(order.begin(), order.end(), rng);
        std::thread t([&] {
            for (size_t k = 0; k < n_pieces; ++k) {
                if (k % 7 == 0) std::this_thread::sleep_for(std::chrono::microseconds(50));
                __atomic_store_n(&flags[order[k]], value, __ATOMIC_RELEASE);
            }
            prod.alive.store(false);
        });
        const size_t missed = qlc_host::widen_stream(pieces.data(), n_pieces, 2, flags.data(), value, still_running, &prod);
        t.join();
        CHECK(missed == 0);
        bool ok = true;
        for (size_t i = 0; i < n_pieces; ++i)
            for (size_t j = 0; j < per; j += 97) ok &= dst[i * per + j] == (pieces[i].dst ? (float)src[i * per + j] : -2.0f);
        CHECK(ok);
    }
    // 3. a producer that stops after 10 of 40 pieces: the call returns and counts the rest as missing
    {
        const size_t n_pieces = 40, per = 4096;
        std::vector<uint8_t> src(n_pieces * per, 3);
        std::vector<float> dst(n_pieces * per, -2.0f);
        std::vector<uint32_t> flags(n_pieces, 0);
        std::vector<qlc_host::StreamPiece> pieces(n_pieces);
        for (size_t i = 0; i < n_pieces; ++i) pieces[i] = qlc_host::StreamPiece{src.data() + i * per, dst.data() + i * per, (uint32_t)per};
        Producer prod;
        std::thread t([&] {
            for (size_t k = 0; k < 10; ++k) __atomic_store_n(&flags[k], 5u, __ATOMIC_RELEASE);
            std::this_thread::sleep_for(std::chrono::milliseconds(2));
            prod.alive.store(false);
        });
        const auto t0 = std::chrono::steady_clock::now();
        const size_t missed = qlc_host::widen_stream(pieces.data(), n_pieces, 1, flags.data(), 5u, still_running, &prod);
        const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
        t.join();
        CHECK(missed == 30);
        CHECK(ms < 2000.0);
        CHECK(dst[0] == 3.0f && dst[9 * per] == 3.0f && dst[10 * per] == -2.0f);
        // the pool is usable afterwards
        std::vector<float> d2(n_pieces * per);
        qlc_host::widen_u8_f32(src.data(), d2.data(), src.size());
        CHECK(d2[123] == 3.0f && d2.back() == 3.0f);
    }
    // 4. a request of a stack or two is widened on the calling thread (no pool dispatch): arrival order, and a producer that stops early
    {
        const size_t n_pieces = 4, per = 7056;
        std::vector<uint8_t> src(n_pieces * per);
        for (auto& v : src) v = (uint8_t)rng();
        std::vector<float> dst(n_pieces * per, -2.0f);
        std::vector<uint32_t> flags(n_pieces, 0);
        std::vector<qlc_host::StreamPiece> pieces(n_pieces);
        for (size_t i = 0; i < n_pieces; ++i) pieces[i] = qlc_host::StreamPiece{src.data() + i * per, dst.data() + i * per, (uint32_t)per};
        Producer prod;
        std::thread t([&] {
            for (size_t k : {size_t(2), size_t(0), size_t(3), size_t(1)}) { std::this_thread::sleep_for(std::chrono::microseconds(30)); __atomic_store_n(&flags[k], 4u, __ATOMIC_RELEASE); }
            prod.alive.store(false);
        });
        CHECK(qlc_host::widen_stream(pieces.data(), n_pieces, 2, flags.data(), 4u, still_running, &prod) == 0);
        t.join();
        bool ok = true;
        for (size_t i = 0; i < src.size(); i += 7) ok &= dst[i] == (float)src[i];
        CHECK(ok);
        std::fill(flags.begin(), flags.end(), 0u); std::fill(dst.begin(), dst.end(), -2.0f);
        Producer prod2;
        std::thread t2([&] { __atomic_store_n(&flags[0], 6u, __ATOMIC_RELEASE); std::this_thread::sleep_for(std::chrono::milliseconds(1)); prod2.alive.store(false); });
        CHECK(qlc_host::widen_stream(pieces.data(), n_pieces, 2, flags.data(), 6u, still_running, &prod2) == 3);
        t2.join();
        CHECK(dst[0] == (float)src[0] && dst[per] == -2.0f);
    }
    // 5. two host threads (two env handles, e.g. one per GPU) use the pool at the same time
    {
        const size_t n = (size_t)40 * 28224;
        std::vector<uint8_t> s1(n), s2(n);
        for (size_t i = 0; i < n; ++i) { s1[i] = (uint8_t)rng(); s2[i] = (uint8_t)rng(); }
        std::vector<float> d1(n), d2(n);
        bool ok1 = true, ok2 = true;
        auto job = [&](const std::vector<uint8_t>& s, std::vector<float>& d, bool& ok) {
            for (int rep = 0; rep < 30; ++rep) {
                std::fill(d.begin(), d.end(), -1.0f);
                qlc_host::widen_u8_f32(s.data(), d.data(), n);
                for (size_t i = 0; i < n; i += 101) ok &= d[i] == (float)s[i];
            }
        };
        std::thread a(job, std::cref(s1), std::ref(d1), std::ref(ok1)), b(job, std::cref(s2), std::ref(d2), std::ref(ok2));
        a.join(); b.join();
        CHECK(ok1 && ok2);
    }
    std::printf(fails ? "host pool: %d check(s) FAILED\n" : "host pool ok (%d threads)\n", fails ? fails : qlc_host::pool_threads());
    return fails ? 1 : 0;
}
