// host_pool.cpp — see host_pool.h. Plain C++ (compiled by the host compiler only).
#include "host_pool.h"

#include <atomic>
#include <chrono>
#include <condition_variable>
#include <cstdio>
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

// Non-temporal variant for outputs far larger than the caches (a 512-minibatch is 115 MB of f32): plain stores read every
// destination line before overwriting it (read-for-ownership), streaming stores do not - the host's DRAM sees 5 instead of 9
// bytes per pixel. Only used from NT_MIN_BYTES on: a minibatch of 32 (7.2 MB) is better left in the cache for its reader.
#if defined(__x86_64__)
#include <immintrin.h>
__attribute__((target("avx512f"))) void widen_nt_avx512(const uint8_t* __restrict__ s, float* __restrict__ d, size_t n) {
    size_t i = 0;
    for (; i < n && ((uintptr_t)(d + i) & 63u); ++i) d[i] = (float)s[i];
    for (; i + 16 <= n; i += 16)
        _mm512_stream_ps(d + i, _mm512_cvtepi32_ps(_mm512_cvtepu8_epi32(_mm_loadu_si128(reinterpret_cast<const __m128i*>(s + i)))));
    for (; i < n; ++i) d[i] = (float)s[i];
    _mm_sfence();
}
__attribute__((target("avx2"))) void widen_nt_avx2(const uint8_t* __restrict__ s, float* __restrict__ d, size_t n) {
    size_t i = 0;
    for (; i < n && ((uintptr_t)(d + i) & 31u); ++i) d[i] = (float)s[i];
    for (; i + 8 <= n; i += 8)
        _mm256_stream_ps(d + i, _mm256_cvtepi32_ps(_mm256_cvtepu8_epi32(_mm_loadl_epi64(reinterpret_cast<const __m128i*>(s + i)))));
    for (; i < n; ++i) d[i] = (float)s[i];
    _mm_sfence();
}
void widen_nt(const uint8_t* s, float* d, size_t n) {
    static const int level = __builtin_cpu_supports("avx512f") ? 2 : (__builtin_cpu_supports("avx2") ? 1 : 0);
    if (level == 2) widen_nt_avx512(s, d, n);
    else if (level == 1) widen_nt_avx2(s, d, n);
    else widen_range(s, d, n);
}
#else
void widen_nt(const uint8_t* s, float* d, size_t n) { widen_range(s, d, n); }
#endif
constexpr size_t NT_MIN_BYTES = (size_t)32 << 20;     // (sharing it among the ranks of a node - 4 MB each at 8 ranks - made a 32-minibatch slower: 0.365 vs 0.267 ms)

constexpr size_t CHUNK = 32 * 1024;     // elements per work item: 32 KB read, 128 KB written

inline void cpu_relax() {
#if defined(__x86_64__) || defined(__i386__)
    __builtin_ia32_pause();
#endif
}

struct Pool {
    std::mutex job_mutex;                    // one job at a time: env handles on different host threads (one per GPU) share this pool
    std::mutex m;
    std::condition_variable cv;
    std::vector<std::thread> workers;
    // the current job. `ticket` = job number << 32 | next item to hand out: an item is claimed by compare-and-swap on the whole
    // word, so a worker that is late for job g can never take (or skip) an item of job g+1 with job g's view of the fields.
    std::atomic<uint64_t> ticket{0};
    // plain job: item c = elements [c * CHUNK, ...) of src -> dst
    const uint8_t* src = nullptr; float* dst = nullptr; size_t n = 0;
    // streamed job (pieces != nullptr): item c = pieces[c], ready when flags[c] == flag_value
    const qlc_host::StreamPiece* pieces = nullptr; const volatile uint32_t* flags = nullptr; uint32_t flag_value = 0;
    std::atomic<bool> abandon{false};
    std::atomic<size_t> missed{0};
    size_t n_items = 0, n_pieces = 0, group = 1;
    bool nt = false;                         // this job's output is far larger than the caches: streaming stores
    std::atomic<size_t> done{0};
    int n_threads = 1;

    void run_item(size_t c, bool (*still_running)(void*) = nullptr, void* ctx = nullptr) {
        if (!pieces) {
            const size_t at = c * CHUNK, len = n - at < CHUNK ? n - at : CHUNK;
            if (nt) widen_nt(src + at, dst + at, len); else widen_range(src + at, dst + at, len);
            return;
        }
        for (size_t i = c * group; i < (c + 1) * group && i < n_pieces; ++i) {     // the pieces of one claim, in arrival order
            const qlc_host::StreamPiece& p = pieces[i];
            if (!p.dst) continue;
            bool landed = true;
            unsigned spins = 0;
            while (__atomic_load_n(const_cast<const uint32_t*>(&flags[i]), __ATOMIC_ACQUIRE) != flag_value) {
                // the calling thread keeps an eye on the producer: once the stream is over, flags that are still missing never come
                if (still_running && (++spins & 255u) == 0 && !still_running(ctx)) abandon.store(true, std::memory_order_release);
                if (abandon.load(std::memory_order_acquire)) {
                    landed = __atomic_load_n(const_cast<const uint32_t*>(&flags[i]), __ATOMIC_ACQUIRE) == flag_value;   // it may have landed meanwhile
                    break;
                }
                cpu_relax();
            }
            if (landed) { if (nt) widen_nt(p.src, p.dst, p.n); else widen_range(p.src, p.dst, p.n); }
            else missed.fetch_add(1, std::memory_order_relaxed);
        }
    }

    void run_items(uint64_t job) {
        for (;;) {
            uint64_t v = ticket.load(std::memory_order_acquire);
            if ((v >> 32) != job) return;                            // that job is over
            const size_t c = (size_t)(v & 0xFFFFFFFFu);
            if (c >= n_items) return;                                // fields belong to `job`: published before its ticket
            if (!ticket.compare_exchange_weak(v, v + 1, std::memory_order_acq_rel)) continue;
            run_item(c);
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
            run_items(seen);
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

    uint64_t publish() {                     // fields of the new job are set: hand it out
        uint64_t job;
        {
            std::lock_guard<std::mutex> lk(m);
            done.store(0, std::memory_order_relaxed);
            job = (ticket.load(std::memory_order_relaxed) >> 32) + 1;
            ticket.store(job << 32, std::memory_order_release);
        }
        cv.notify_all();
        return job;
    }

    // no item of this job can be claimed any more: a straggler that still holds the job's ticket value fails its compare-and-swap,
    // so the fields may be rewritten for the next job
    void close(uint64_t job) { ticket.store((job << 32) | 0xFFFFFFFFull, std::memory_order_release); }

    void widen(const uint8_t* s, float* d, size_t count) {
        if (count == 0) return;
        std::lock_guard<std::mutex> one_job(job_mutex);
        if (n_threads == 1 || count <= 2 * CHUNK) { if (count * 4 >= NT_MIN_BYTES) widen_nt(s, d, count); else widen_range(s, d, count); return; }
        pieces = nullptr; src = s; dst = d; n = count; n_items = (count + CHUNK - 1) / CHUNK; nt = count * 4 >= NT_MIN_BYTES;
        const uint64_t job = publish();
        run_items(job);
        while (done.load(std::memory_order_acquire) < n_items) std::this_thread::yield();   // every claimed item has been written
        close(job);
    }

    size_t stream(const qlc_host::StreamPiece* ps, size_t count, size_t grp, const volatile uint32_t* fl, uint32_t value, bool (*still_running)(void*), void* ctx) {
        if (count == 0) return 0;
        std::lock_guard<std::mutex> one_job(job_mutex);
        static const bool timing = getenv("QLC_HOST_TIMING") != nullptr;
        const auto t_pub = std::chrono::steady_clock::now();
        if (timing) {      // one thread, in order: when does the first / the last flag come up, how long does the widening take behind it
            double first = -1, last = 0, work = 0;
            for (size_t c = 0; c < count; ++c) {
                if (!ps[c].dst) continue;
                while (__atomic_load_n(const_cast<const uint32_t*>(&fl[c]), __ATOMIC_ACQUIRE) != value) cpu_relax();
                const auto a = std::chrono::steady_clock::now();
                if (first < 0) first = std::chrono::duration<double, std::micro>(a - t_pub).count();
                last = std::chrono::duration<double, std::micro>(a - t_pub).count();
                widen_range(ps[c].src, ps[c].dst, ps[c].n);
                work += std::chrono::duration<double, std::micro>(std::chrono::steady_clock::now() - a).count();
            }
            std::fprintf(stderr, "[qlc host stream] %zu pieces: first flag %.1f us after publish, last flag seen at %.1f us, widening %.1f us (1 thread)\n", count, first, last, work);
            return 0;
        }
        size_t out_bytes = 0;
        for (size_t i = 0; i < count; ++i) if (ps[i].dst) out_bytes += (size_t)ps[i].n * 4;
        if (n_threads == 1 || out_bytes <= 2 * CHUNK * 4) {
            // a stack or two (predict_action's single observation): waking the pool costs more than the widening
            size_t lost = 0;
            for (size_t i = 0; i < count; ++i) {
                if (!ps[i].dst) continue;
                unsigned spins = 0; bool landed = true;
                while (__atomic_load_n(const_cast<const uint32_t*>(&fl[i]), __ATOMIC_ACQUIRE) != value) {
                    if ((++spins & 255u) == 0 && !still_running(ctx)) { landed = __atomic_load_n(const_cast<const uint32_t*>(&fl[i]), __ATOMIC_ACQUIRE) == value; break; }
                    cpu_relax();
                }
                if (!landed) { ++lost; continue; }
                if (out_bytes >= NT_MIN_BYTES) widen_nt(ps[i].src, ps[i].dst, ps[i].n); else widen_range(ps[i].src, ps[i].dst, ps[i].n);
            }
            return lost;
        }
        pieces = ps; flags = fl; flag_value = value; n_pieces = count; group = grp ? grp : 1; n_items = (count + group - 1) / group;
        nt = out_bytes >= NT_MIN_BYTES;
        abandon.store(false, std::memory_order_relaxed); missed.store(0, std::memory_order_relaxed);
        const uint64_t job = publish();
        // the calling thread takes pieces like a worker, but looks after the producer between them: when the stream is over, flags
        // that are still missing will never come
        for (;;) {
            uint64_t v = ticket.load(std::memory_order_acquire);
            const size_t c = (size_t)(v & 0xFFFFFFFFu);
            if (c >= n_items) break;
            if (!ticket.compare_exchange_weak(v, v + 1, std::memory_order_acq_rel)) continue;
            run_item(c, still_running, ctx);
            done.fetch_add(1, std::memory_order_release);
        }
        unsigned spins = 0;
        while (done.load(std::memory_order_acquire) < n_items) {
            if ((++spins & 255u) == 0 && !still_running(ctx)) abandon.store(true, std::memory_order_release);
            cpu_relax();
        }
        close(job);
        pieces = nullptr;
        return missed.load(std::memory_order_relaxed);
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
size_t widen_stream(const StreamPiece* pieces, size_t n_pieces, size_t group, const volatile uint32_t* flags, uint32_t flag_value, bool (*still_running)(void*), void* ctx) {
    return pool().stream(pieces, n_pieces, group, flags, flag_value, still_running, ctx);
}
}  // namespace qlc_host
