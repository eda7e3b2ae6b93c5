// host_pool.cpp — see host_pool.h. Plain C++ (compiled by the host compiler only).
#include "host_pool.h"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdlib>
#include <mutex>
#include <thread>
#include <vector>

namespace {

// one clone per ISA level, picked at load time by the ifunc resolver (the GPU box's host CPU is not the build container's)
__attribute__((target_clones("avx512f", "avx2", "default")))
void widen_range(const uint8_t* __restrict__ s, float* __restrict__ d, size_t n) {
    for (size_t i = 0; i < n; ++i) d[i] = (float)s[i];
}

constexpr size_t CHUNK = 32 * 1024;     // elements per work item: 32 KB read, 128 KB written

struct Pool {
    std::mutex m;
    std::condition_variable cv;
    std::vector<std::thread> workers;
    // the current job. `ticket` = job number << 32 | next chunk to hand out: a chunk is claimed by compare-and-swap on the whole
    // word, so a worker that is late for job g can never take (or skip) a chunk of job g+1 with job g's view of the fields.
    std::atomic<uint64_t> ticket{0};
    const uint8_t* src = nullptr; float* dst = nullptr; size_t n = 0, n_chunks = 0;
    std::atomic<size_t> done{0};
    int n_threads = 1;

    void run_chunks(uint64_t job) {
        for (;;) {
            uint64_t v = ticket.load(std::memory_order_acquire);
            if ((v >> 32) != job) return;                            // that job is over
            const size_t c = (size_t)(v & 0xFFFFFFFFu);
            if (c >= n_chunks) return;                               // fields belong to `job`: published before its ticket
            if (!ticket.compare_exchange_weak(v, v + 1, std::memory_order_acq_rel)) continue;
            const size_t at = c * CHUNK, len = n - at < CHUNK ? n - at : CHUNK;
            widen_range(src + at, dst + at, len);
            done.fetch_add(1, std::memory_order_release);
        }
    }

    void worker() {
        uint64_t seen = 0;
        for (;;) {
            // spin briefly (back-to-back jobs: state, then state_next), then sleep
            const auto t0 = std::chrono::steady_clock::now();
            while ((ticket.load(std::memory_order_acquire) >> 32) == seen) {
                if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(200)) {
                    std::unique_lock<std::mutex> lk(m);
                    cv.wait(lk, [&] { return (ticket.load(std::memory_order_acquire) >> 32) != seen; });
                    break;
                }
                std::this_thread::yield();
            }
            seen = ticket.load(std::memory_order_acquire) >> 32;
            run_chunks(seen);
        }
    }

    Pool() {
        int want = 0;
        if (const char* e = getenv("QLC_HOST_THREADS")) want = atoi(e);
        if (want <= 0) {
            unsigned hw = std::thread::hardware_concurrency();
            // one process per GPU shares the host: LOCAL_WORLD_SIZE (torchrun) ranks split the cores
            int ranks = 1;
            if (const char* e = getenv("LOCAL_WORLD_SIZE")) ranks = atoi(e) > 0 ? atoi(e) : 1;
            want = ((int)(hw ? hw : 2) - 1) / ranks;     // (one hardware thread is left to the driver's own threads) measured on the 16-thread GPU box (tools/microbench/widen_rate.cpp): 26 / 51 / 68 /
            if (want > 16) want = 16;                    // 106 / 183 GB/s with 1 / 2 / 4 / 8 / 16 threads - it keeps scaling to every hardware thread
        }
        if (want < 1) want = 1;
        n_threads = want;
        for (int i = 1; i < n_threads; ++i) workers.emplace_back([this] { worker(); });
        for (auto& t : workers) t.detach();      // parked on the condition variable for the life of the process
    }

    void widen(const uint8_t* s, float* d, size_t count) {
        if (count == 0) return;
        if (n_threads == 1 || count <= 2 * CHUNK) { widen_range(s, d, count); return; }
        uint64_t job;
        {
            std::lock_guard<std::mutex> lk(m);
            src = s; dst = d; n = count; n_chunks = (count + CHUNK - 1) / CHUNK;
            done.store(0, std::memory_order_relaxed);
            job = (ticket.load(std::memory_order_relaxed) >> 32) + 1;
            ticket.store(job << 32, std::memory_order_release);
        }
        cv.notify_all();
        run_chunks(job);
        while (done.load(std::memory_order_acquire) < n_chunks) std::this_thread::yield();   // every claimed chunk has been written
    }
};

Pool& pool() {
    static Pool* p = new Pool();     // never destroyed: detached workers may still be parked at exit
    return *p;
}

}  // namespace

namespace qlc_host {
int pool_threads() { return pool().n_threads; }
void widen_u8_f32(const uint8_t* src, float* dst, size_t n) { pool().widen(src, dst, n); }
}  // namespace qlc_host
