// widen_rate.cpp — u8 -> f32 widening throughput of the host pool (csrc/host_pool.cpp) on this machine's cores.
//   g++ -O3 -std=c++17 -pthread tools/microbench/widen_rate.cpp q-learning_b200/csrc/host_pool.cpp -o tools/microbench/widen_rate
//   QLC_HOST_THREADS=8 tools/microbench/widen_rate
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "../../q-learning_b200/csrc/host_pool.h"

int main() {
    const size_t n = 32 * 28224;           // one stack of a 32-minibatch
    std::vector<uint8_t> s(n);
    std::vector<float> d(n), d2(n);
    for (size_t i = 0; i < n; ++i) s[i] = (uint8_t)(i * 2654435761u >> 24);
    for (int i = 0; i < 50; ++i) { qlc_host::widen_u8_f32(s.data(), d.data(), n); qlc_host::widen_u8_f32(s.data(), d2.data(), n); }
    const auto t0 = std::chrono::steady_clock::now();
    const int reps = 500;
    for (int i = 0; i < reps; ++i) { qlc_host::widen_u8_f32(s.data(), d.data(), n); qlc_host::widen_u8_f32(s.data(), d2.data(), n); }
    const double us = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t0).count() / reps;
    std::printf("threads %d: state + next of a 32-minibatch (7.2 MB written) in %.1f us = %.1f GB/s\n", qlc_host::pool_threads(), us, 2 * n * 4 / us / 1e3);
    // a 512-minibatch: 2 x 57.8 MB of f32 - far beyond the caches, the pool switches to streaming (non-temporal) stores
    const size_t nb = (size_t)512 * 28224;
    std::vector<uint8_t> sb(nb);
    std::vector<float> db(nb);
    for (size_t i = 0; i < nb; ++i) sb[i] = (uint8_t)(i * 2654435761u >> 24);
    qlc_host::widen_u8_f32(sb.data(), db.data(), nb);
    const auto t1 = std::chrono::steady_clock::now();
    for (int i = 0; i < 10; ++i) qlc_host::widen_u8_f32(sb.data(), db.data(), nb);
    const double usb = std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - t1).count() / 10;
    std::printf("threads %d: one stack of a 512-minibatch (57.8 MB written) in %.1f us = %.1f GB/s\n", qlc_host::pool_threads(), usb, nb * 4 / usb / 1e3);
    return 0;
}
