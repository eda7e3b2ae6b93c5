// kernels.cuh — sm_100a kernels of the hot path: fused env step + rasterise + frame-ring append + replay record,
// Philox distinct-index sampling, and the frame-stack gather (TMA bulk copies through shared memory).
#pragma once
#include <stdint.h>
#include <cuda_runtime.h>
#include "physics.cuh"

namespace qlc {

constexpr int FRAME_W = 84, FRAME_H = 84, FRAME_BYTES = FRAME_W * FRAME_H;   // 7056 = 441 * 16
constexpr int FRAME_VEC16 = FRAME_BYTES / 16;
constexpr int ENVS_PER_CTA = 32;          // one physics warp = one "env batch"

// transition record, one u32 per (time slot, env) — ReplayBuffer::add (replay_buffer.rs:85-98) as a scalar write:
//   bits 0-1 action | bit 2 done | bits 3-4 k mod 4 | bits 5-7 min(k,4) | bits 8-15 reward | bit 16 truncated
// k = frames already in the episode's ring when the step was taken (episode_step before the step).
__host__ __device__ __forceinline__ uint32_t pack_record(uint32_t action, bool done, uint32_t k, uint32_t reward, bool truncated) {
    return (action & 3u) | (done ? 4u : 0u) | ((k & 3u) << 3) | ((k < 4u ? k : 4u) << 5) | ((reward & 0xFFu) << 8) | (truncated ? 0x10000u : 0u);
}

struct EnvArrays {
    float *ball_cx, *ball_cy, *ball_dx, *ball_dy, *pad_min_x, *pad_max_x, *pad_speed;
    uint64_t* bricks;
    uint32_t *score, *episode_step, *episode, *err;
    uint8_t* finished;
};

struct DeviceStats {                      // order-independent accumulators
    unsigned long long sum_return, episodes;
    unsigned int min_return, max_return;
};

struct RasterTables;

// Release builds cannot skip work: the ablation switch exists only in -DQLC_PROFILING builds (qlc_build_info() says which).
#ifdef QLC_PROFILING
#define QLC_DEBUG_SKIP_IS(p, v) ((p).debug_skip == (v))
#else
#define QLC_DEBUG_SKIP_IS(p, v) false
#endif

// Timeline of a launch (profiling builds only, QLC_TIMELINE_FILE): clock64 stamps of CTA b at fixed points, 16 slots per CTA.
#ifdef QLC_PROFILING
#define QLC_STAMP(p, slot) do { if ((p).timeline && blockIdx.x < 1024u) (p).timeline[blockIdx.x * 16u + (slot)] = (unsigned long long)clock64(); } while (0)
#define QLC_STAMP_SEQ(p) ((p).debug_skip >= 16u ? (p).debug_skip - 16u : 0u)     /* which step of the launch the per-step stamps are taken at */
#define QLC_STAMP_GT(p, slot) do { if ((p).timeline && blockIdx.x < 1024u) { unsigned long long t_; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t_)); (p).timeline[blockIdx.x * 16u + (slot)] = t_; } } while (0)
#else
#define QLC_STAMP(p, slot) do { } while (0)
#define QLC_STAMP_GT(p, slot) do { } while (0)
#define QLC_STAMP_SEQ(p) 0u
#endif

struct StepParams {
    uint32_t n_envs, env_id_base, time_slots, max_episode_steps, auto_reset, n_steps;
    unsigned long long* timeline;   // profiling builds: [1024][16] stamps of the LAST launch (NULL = off)
    uint32_t debug_skip;       // ablation aid, only honoured in -DQLC_PROFILING builds (QLC_DEBUG_SKIP): 1 = no physics, 2 = no frame stores
    const RasterTables* tables;   // built once per env handle by raster_tables_kernel
    unsigned int* work_counter; uint32_t work_base;   // work hand-out: first item = blockIdx.x, then gridDim.x + atomicAdd(counter, 1) - base
    // time chunking (0 = off): an item is (chunk c, batch b) = steps [c*chunk_len, (c+1)*chunk_len) of batch b; chunk c of a
    // batch may run on another CTA than chunk c-1 — the env state travels through HBM and `progress[b]` (launch serial << 32 |
    // steps done) is the release/acquire flag. Lets fast SMs / GPCs take more of a single-wave launch.
    uint32_t chunk_len; uint32_t launch_serial; unsigned long long* progress; unsigned int* spin_error;
    uint32_t epc;              // envs per CTA (<= R*NE), chosen by the host so that the grid fills all SMs evenly
    uint64_t t0, seed;
    uint32_t slot0;            // t0 mod time_slots (n_steps <= time_slots: the slot of step s is slot0 + s, minus time_slots if it wraps)
    uint8_t* frames; uint32_t* records; DeviceStats* stats;
    // statistics snapshot (NULL = off): the last CTA to finish copies the shard accumulators here, so that a reduction running on
    // a side stream reads the state "after this launch" while later launches already mutate `stats`
    DeviceStats* snap; unsigned int* exit_counter; uint32_t exit_base;
    const uint8_t* actions;    // [n_steps][n_envs]; NULL = uniform random policy drawn here (policy_action), written to actions_out if given
    uint8_t* actions_out;
    float* reward; uint8_t* done;
};

// ---------------------------------------------------------------------------------------------------------
// PTX helpers: mbarrier, bulk async copies (TMA 1-D), proxy fence
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "WAIT_%=:\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
        "@p bra DONE_%=;\n\t"
        "bra WAIT_%=;\n\t"
        "DONE_%=:\n\t}" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void fence_mbar_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// shared -> global, completion tracked by the issuing thread's bulk groups
__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait_read() { asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory"); }
template <int N> __device__ __forceinline__ void bulk_wait() { asm volatile("cp.async.bulk.wait_group %0;" ::"n"(N) : "memory"); }
// global -> shared, completion on an mbarrier
__device__ __forceinline__ void bulk_load(void* sdst, const void* gsrc, uint32_t bytes, uint64_t* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(sdst)), "l"(gsrc),
                 "r"(bytes), "r"(smem_u32(bar))
                 : "memory");
}

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
    unsigned long long v;
    asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void st_release_u64(unsigned long long* p, unsigned long long v) {
    asm volatile("st.release.gpu.global.u64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}

// ---------------------------------------------------------------------------------------------------------
// Raster tables (computed per CTA from the f32 raster spec, never hard-coded)
// ---------------------------------------------------------------------------------------------------------
struct RasterTables {
    uint4 word_bits[FRAME_W / 4];   // per 4-pixel word of a brick row: the brick bit (1<<k) covering each pixel, 0 = gap
    int grp_first[3], grp_count[3]; // pixel rows of brick row r
    int pad_first, pad_count;       // pixel rows of the paddle
    uint32_t col_bit[FRAME_W];
    int8_t row_grp[FRAME_H];
    int8_t row_pad[FRAME_H];
};

__device__ __forceinline__ float scale84(float pos) { return pos * 84.0f / 600.0f; }   // app_game_drawer.rs:21-36

__device__ __forceinline__ void build_raster_tables(RasterTables& T, int tid, int nthreads) {
    for (int i = tid; i < FRAME_W; i += nthreads) {
        const float p = (float)i + 0.5f;
        uint32_t bit = 0u;
        for (int k = 0; k < 20; ++k) {
            const float x0 = scale84(30.0f + 27.0f * (float)k), x1 = scale84(55.0f + 27.0f * (float)k);
            if (p >= x0 && p < x1) bit = 1u << k;
        }
        T.col_bit[i] = bit;
        int g = -1;
        for (int r = 0; r < 3; ++r) {
            const float y0 = scale84(35.0f + 27.0f * (float)r), y1 = scale84(60.0f + 27.0f * (float)r);
            if (p >= y0 && p < y1) g = r;
        }
        T.row_grp[i] = (int8_t)g;
        T.row_pad[i] = (int8_t)((p >= scale84(PAD_MIN_Y) && p < scale84(PAD_MAX_Y)) ? 1 : 0);
    }
    __syncthreads();
    for (int w = tid; w < FRAME_W / 4; w += nthreads)
        T.word_bits[w] = make_uint4(T.col_bit[4 * w], T.col_bit[4 * w + 1], T.col_bit[4 * w + 2], T.col_bit[4 * w + 3]);
    if (tid < 3) {
        int first = 0, count = 0;
        for (int j = 0; j < FRAME_H; ++j) if (T.row_grp[j] == tid) { if (!count) first = j; ++count; }
        T.grp_first[tid] = first; T.grp_count[tid] = count;
    }
    if (tid == 3) {
        int first = 0, count = 0;
        for (int j = 0; j < FRAME_H; ++j) if (T.row_pad[j]) { if (!count) first = j; ++count; }
        T.pad_first = first; T.pad_count = count;
    }
    __syncthreads();
}

// one-time table build (env creation): the per-launch prologue only copies the ~1 KB result into shared memory
__global__ void raster_tables_kernel(RasterTables* out) {
    __shared__ RasterTables T;
    build_raster_tables(T, threadIdx.x, blockDim.x);
    uint32_t* src = reinterpret_cast<uint32_t*>(&T); uint32_t* dst = reinterpret_cast<uint32_t*>(out);
    for (int i = threadIdx.x; i < (int)(sizeof(RasterTables) / 4); i += blockDim.x) dst[i] = src[i];
}

// What a render warp needs to draw one frame; the scaled coordinates are computed once per env by the physics lane
// (pos * 84 / 600, app_game_drawer.rs:21-36) instead of redundantly by all 32 lanes of a render warp.
struct __align__(16) RenderRec {
    float bx, by;        // ball centre, frame coordinates
    float x0, x1;        // paddle [min.x, max.x), frame coordinates
    uint64_t bricks;
    int32_t box;         // top-left pixel of the 6x6 ball box: (j0 << 16) | (i0 & 0xFFFF)
    uint32_t pad;
};

// ---------------------------------------------------------------------------------------------------------
// Kernel 1: env_advance — n_steps of {time_step, rasterise, frame-ring append, replay record} for R*NE envs per CTA.
//   warp 0     : physics, one lane per env, state in registers across all n_steps; publishes a 24-byte render
//                record per env and step into a D-deep shared-memory queue (mbarrier full/empty).
//   warps 1..R : each owns NE envs and ONE private 7,056-byte shared-memory frame per env that stays resident for
//                the whole launch. Per step only what changed is redrawn (old ball box cleared, paddle row
//                rewritten, ball drawn; the brick band only when a brick vanished or the ball left it), then the
//                frame goes to the HBM frame ring as one cp.async.bulk (UBLKCP) of 7,056 contiguous bytes.
// Raster spec: DESIGN.md "Raster spec" (bricks luma 96, ball ring luma 236, paddle luma 255; later overwrites
// earlier; pixel-centre coverage) — identical pixels to drawing every frame from scratch.
// ---------------------------------------------------------------------------------------------------------
template <int D, int EPC_MAX>
struct AdvanceSmem {
    RasterTables tables;
    RenderRec queue[D][EPC_MAX];
    uint32_t item_env0[D], item_n[D], item_step[D];   // which env batch / step a queue slot carries; item_n == 0 = no more work
    uint64_t full[D], empty[D];
};

template <int R, int NE, int D, int MINB>
__global__ void __launch_bounds__(32 * (R + 1), MINB) env_advance_kernel(EnvArrays st, StepParams p) {
    static_assert(R * NE <= ENVS_PER_CTA, "one physics lane per env");
    const uint32_t EPC = p.epc;                       // envs per CTA (<= R*NE)
    extern __shared__ __align__(128) uint8_t dyn_smem[];
    __shared__ AdvanceSmem<D, R * NE> S;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t n_batches = (p.n_envs + EPC - 1) / EPC;   // env batches, handed out to the CTAs dynamically
    if (tid == 0) { QLC_STAMP_GT(p, 0); QLC_STAMP(p, 1); }

    if (tid == 0) {
        for (int q = 0; q < D; ++q) { mbar_init(&S.full[q], 1); mbar_init(&S.empty[q], R); }
        fence_mbar_init();
    }
    {   // raster tables: ~1 KB copy from the per-handle global copy, requested first so that the zero fill hides its latency
        static_assert(sizeof(RasterTables) % 4 == 0 && sizeof(RasterTables) / 4 <= 2 * 32 * (R + 1), "word copy, at most two words per thread");
        const uint32_t* src = reinterpret_cast<const uint32_t*>(p.tables); uint32_t* dst = reinterpret_cast<uint32_t*>(&S.tables);
        constexpr int WORDS = (int)(sizeof(RasterTables) / 4), NT = 32 * (R + 1);
        const uint32_t w0 = tid < WORDS ? __ldg(src + tid) : 0u;
        const uint32_t w1 = tid + NT < WORDS ? __ldg(src + tid + NT) : 0u;
        if (warp != 0) {   // zero the resident frames
            uint4* z = reinterpret_cast<uint4*>(dyn_smem + (size_t)(warp - 1) * NE * FRAME_BYTES);
            for (int i = lane; i < NE * FRAME_VEC16; i += 32) z[i] = make_uint4(0, 0, 0, 0);
        }
        if (tid < WORDS) dst[tid] = w0;
        if (tid + NT < WORDS) dst[tid + NT] = w1;
    }
    __syncthreads();
    // Programmatic dependent launch (single-step launches, qlc_api.cu): everything above touches nothing an earlier kernel
    // wrote, so it may overlap the predecessor's tail; from here on its results (actions, env state, frame ring) are needed.
    if (tid == 0) QLC_STAMP(p, 2);
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (tid == 0) QLC_STAMP(p, 3);

    if (warp == 0) {
        // ------------------------------- physics warp -------------------------------
        uint32_t seq = 0;                                  // (batch, step) items published so far
        bool first_done = false;
        for (;;) {
        // env batches are handed out dynamically (SMs differ in their distance to L2/HBM; a static split would wait for the slowest)
        const uint32_t n_chunks = p.chunk_len ? (p.n_steps + p.chunk_len - 1) / p.chunk_len : 1u;
        const uint32_t n_items = n_batches * n_chunks;
        uint32_t item;
        if (seq == 0 && !first_done) { item = blockIdx.x; first_done = true; }       // first item is static: no atomic on the critical path
        else if (n_items <= gridDim.x) item = n_items;                               // one item per CTA: nothing left to hand out
        else {
            item = 0;
            if (lane == 0) item = gridDim.x + (atomicAdd(p.work_counter, 1u) - p.work_base);
            item = __shfl_sync(0xFFFFFFFFu, item, 0);
        }
        if (item >= n_items) {
            const int q = seq % D;
            if (seq >= (uint32_t)D) mbar_wait(&S.empty[q], ((seq / D) - 1) & 1);
            if (lane == 0) { S.item_n[q] = 0u; mbar_arrive(&S.full[q]); }
            break;
        }
        const uint32_t chunk = item / n_batches, batch = item - chunk * n_batches;     // chunk-major: predecessors are handed out first
        const uint32_t s_begin = p.chunk_len ? chunk * p.chunk_len : 0u;
        const uint32_t s_end = p.chunk_len ? min(s_begin + p.chunk_len, p.n_steps) : p.n_steps;
        if (chunk > 0) {
            // wait until chunk-1 of this batch has been finished (by whichever CTA took it) and its state is visible
            if (lane == 0) {
                const unsigned long long want = ((unsigned long long)p.launch_serial << 32) | s_begin;
                uint32_t polls = 0;
                while (ld_acquire_u64(&p.progress[batch]) != want) {
                    __nanosleep(100);
                    if (++polls > (1u << 24)) { atomicExch(p.spin_error, 1u); break; }     // never hang the GPU
                }
            }
            __syncwarp();
        }
        const uint32_t env0 = batch * EPC;
        const uint32_t n_here = min(EPC, p.n_envs - env0);
        const uint32_t e = env0 + lane;
        const bool active = lane < n_here;
        // Shapes with register headroom (<= 2 CTAs per SM: the learner-driven single-step launches) request the first action before
        // the state, so that its trip to HBM hides behind the state loads and the move cache instead of following them (timeline of
        // a single-step launch: 3,500 -> 2,600 cycles from the dependency wait to the physics). The 3-CTA shapes sit exactly at their
        // 72-register budget: one more live value there is a spill in the headline launch.
        uint32_t action = 0u;
        constexpr bool EARLY_ACTION = MINB <= 2;
        if (EARLY_ACTION && active && p.actions) action = p.actions[(size_t)s_begin * p.n_envs + e];
        Env env; uint32_t k = 0, episode = 0;
        if (active) {   // __ldcg: read at L2 — with chunking another SM may have written this state during the launch
            env.cx = __ldcg(&st.ball_cx[e]); env.cy = __ldcg(&st.ball_cy[e]); env.dx = __ldcg(&st.ball_dx[e]); env.dy = __ldcg(&st.ball_dy[e]);
            env.pmin = __ldcg(&st.pad_min_x[e]); env.pmax = __ldcg(&st.pad_max_x[e]); env.pspeed = __ldcg(&st.pad_speed[e]);
            env.bricks = __ldcg(&st.bricks[e]); env.score = __ldcg(&st.score[e]); env.err = __ldcg(&st.err[e]); env.finished = __ldcg(&st.finished[e]) != 0;
            k = __ldcg(&st.episode_step[e]); episode = __ldcg(&st.episode[e]);
        } else {
            env_init(env, -0.25f); env.err = 0;
        }
        MoveCache mc; move_cache_update(mc, env);
        if (active && !(EARLY_ACTION && p.actions)) action = p.actions ? p.actions[(size_t)s_begin * p.n_envs + e] : policy_action(p.seed, p.env_id_base + e, (uint32_t)(p.t0 + s_begin));
        for (uint32_t s = s_begin; s < s_end; ++s, ++seq) {
            const int q = seq % D;
            uint32_t next_action = 0u;
            if (active && s + 1 < p.n_steps) next_action = p.actions ? p.actions[(size_t)(s + 1) * p.n_envs + e] : policy_action(p.seed, p.env_id_base + e, (uint32_t)(p.t0 + s + 1));
            if (active && p.actions_out) p.actions_out[(size_t)s * p.n_envs + e] = (uint8_t)action;
            if (action >= 3u) { env.err |= ENVERR_ACTION; action = 0u; }
            const uint32_t score_before = env.score;
            if (lane == 0 && seq == QLC_STAMP_SEQ(p)) QLC_STAMP(p, 4);
            if (active && !QLC_DEBUG_SKIP_IS(p, 1u)) time_step(env, action, mc);
            __syncwarp();
            if (lane == 0 && seq == QLC_STAMP_SEQ(p)) QLC_STAMP(p, 5);
            if (seq >= (uint32_t)D) mbar_wait(&S.empty[q], ((seq / D) - 1) & 1);
            if (active) {
                RenderRec rr;
                rr.bx = scale84(env.cx); rr.by = scale84(env.cy); rr.x0 = scale84(env.pmin); rr.x1 = scale84(env.pmax); rr.bricks = env.bricks;
                const int i0 = (int)floorf(fminf(fmaxf(rr.bx, -100.0f), 200.0f) - 2.9f);   // ring radius 2.4 + half a pixel
                const int j0 = (int)floorf(fminf(fmaxf(rr.by, -100.0f), 200.0f) - 2.9f);
                rr.box = (j0 << 16) | (i0 & 0xFFFF); rr.pad = 0u;
                S.queue[q][lane] = rr;
                if (lane == 0) { S.item_env0[q] = env0; S.item_n[q] = n_here; S.item_step[q] = s; }
                const uint32_t reward = env.score - score_before;
                const bool done = env.finished;
                const uint32_t k_before = k;
                k += 1;
                const bool truncated = !done && p.max_episode_steps != 0u && k >= p.max_episode_steps;
                const uint32_t slot = (uint32_t)((p.t0 + s) % p.time_slots);
                p.records[(size_t)slot * p.n_envs + e] = pack_record(action, done, k_before, reward, truncated);
                if (p.reward) p.reward[(size_t)s * p.n_envs + e] = (float)reward;
                if (p.done) p.done[(size_t)s * p.n_envs + e] = done ? 1 : 0;
                if ((done || truncated) && p.auto_reset) {
                    // episode end: shard statistics, then restart on device (learn_episode :142,:220)
                    atomicAdd(&p.stats->sum_return, (unsigned long long)env.score);
                    atomicAdd(&p.stats->episodes, 1ull);
                    atomicMin(&p.stats->min_return, env.score);
                    atomicMax(&p.stats->max_return, env.score);
                    episode += 1;
                    env_init(env, reset_dir_x(p.seed, p.env_id_base + e, episode));
                    move_cache_update(mc, env);
                    k = 0;
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.full[q]);
            if (lane == 0 && seq == QLC_STAMP_SEQ(p)) QLC_STAMP(p, 6);
            action = next_action;
        }
        if (active) {
            st.ball_cx[e] = env.cx; st.ball_cy[e] = env.cy; st.ball_dx[e] = env.dx; st.ball_dy[e] = env.dy;
            st.pad_min_x[e] = env.pmin; st.pad_max_x[e] = env.pmax; st.pad_speed[e] = env.pspeed;
            st.bricks[e] = env.bricks; st.score[e] = env.score; st.err[e] = env.err; st.finished[e] = env.finished ? 1 : 0;
            st.episode_step[e] = k; st.episode[e] = episode;
        }
        if (p.chunk_len) {   // publish: state stores of all lanes, then the flag (release)
            __threadfence();
            __syncwarp();
            if (lane == 0) st_release_u64(&p.progress[batch], ((unsigned long long)p.launch_serial << 32) | s_end);
        }
        }   // items
        // Programmatic dependent launch, triggered LATE: the physics of every item of this CTA is done, what is left is the
        // render warps' last frames and the store drain. Once every CTA has got here the next kernel of the stream may start
        // its prologue (barriers, zeroed frames, index sampling) on the SMs that free up; it blocks in its own
        // griddepcontrol.wait until this grid has completed and flushed. (Triggering at the top of the kernel would let a small
        // dependent grid — a minibatch gather — pile onto the few SMs this grid leaves free.)
        asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
        if (lane == 0) QLC_STAMP(p, 7);
        if (p.snap) {    // every statistics atomic of this CTA is done: count it out, the last one takes the snapshot
            __threadfence();
            __syncwarp();
            if (lane == 0 && atomicAdd(p.exit_counter, 1u) - p.exit_base == gridDim.x - 1u) {
                __threadfence();
                DeviceStats v;
                v.sum_return = atomicAdd(&p.stats->sum_return, 0ull); v.episodes = atomicAdd(&p.stats->episodes, 0ull);
                v.min_return = atomicMin(&p.stats->min_return, 0xFFFFFFFFu); v.max_return = atomicMax(&p.stats->max_return, 0u);
                *p.snap = v;
            }
        }
    } else {
        // ------------------------------- render warps -------------------------------
        const int rw = warp - 1;
        const RasterTables& T = S.tables;
        uint8_t* my_bufs = dyn_smem + (size_t)rw * NE * FRAME_BYTES;
        // per-lane constants: which pixel(s) of a 6x6 ball box and which word(s) of the brick band this lane owns
        const int bdi0 = lane % 6, bdj0 = lane / 6, bdi1 = (lane + 32) % 6, bdj1 = (lane + 32) / 6;   // 2nd only for lanes 0..3
        constexpr int WPR = FRAME_W / 4;   // 21 words per pixel row
        const int bg0 = lane / WPR, bw0 = lane % WPR, bg1 = (lane + 32) / WPR, bw1 = (lane + 32) % WPR;   // 2nd only for lanes 0..30
        const int band_lo = T.grp_first[0], band_hi = T.grp_first[2] + T.grp_count[2] - 1;
        const int pad_first = T.pad_first, pad_count = T.pad_count;
        const float rs = scale84(BALL_R);
        const float r_out = rs + 1.0f, r_in = rs - 1.0f;
        const float out2 = r_out * r_out, in2 = r_in * r_in;
        uint64_t prev_bricks[NE]; int prev_box[NE];
        #pragma unroll
        for (int i = 0; i < NE; ++i) { prev_bricks[i] = 0ull; prev_box[i] = (int)0x80008000u; }   // box far outside the frame

        // bulk groups per step: one half of the warp's frames drains while the other is drawn. (One group - all frames drawn, one proxy
        // fence, all stores - is no faster for single-step launches, 10.91 vs 10.60 us: the SM's bulk-copy unit takes the 16 frames of a
        // CTA one after the other either way, ~0.8 us per 7 KB store issue when every warp issues at once.)
        constexpr int G = NE >= 2 ? 2 : 1;
        constexpr int FPG = NE / G;                 // frames per group
        static_assert(NE % G == 0, "NE must be 1 or even");
        // Learner-driven launches (a step or two per launch): while the physics warp is still loading the state and advancing the
        // ball, draw the brick band of this CTA's first (static) item from the brick masks as they are BEFORE the step - most steps
        // leave them alone, and the incremental redraw below rewrites the band whenever a brick did vanish. Takes the band (the bulk
        // of a fresh frame's phase 1) off the launch's critical path; same pixels by construction.
        if (p.n_steps < 4u && p.chunk_len == 0u && blockIdx.x < n_batches) {
            const uint32_t env0 = blockIdx.x * EPC, n_here = min(EPC, p.n_envs - env0);
            #pragma unroll
            for (int i = 0; i < NE; ++i) {
                const uint32_t j = (uint32_t)(i * R + rw);
                if (j < n_here) {
                    const uint64_t bricks = __ldcg(&st.bricks[env0 + j]);
                    uint32_t* wbuf = reinterpret_cast<uint32_t*>(my_bufs + (size_t)i * FRAME_BYTES);
                    {
                        const uint32_t m = (uint32_t)(bricks >> (20 * bg0)) & 0xFFFFFu;
                        const uint4 b = T.word_bits[bw0];
                        const uint32_t v = ((m & b.x) ? 96u : 0u) | ((m & b.y) ? 96u << 8 : 0u) | ((m & b.z) ? 96u << 16 : 0u) | ((m & b.w) ? 96u << 24 : 0u);
                        const int first = T.grp_first[bg0], count = T.grp_count[bg0];
                        for (int r = 0; r < count; ++r) wbuf[(first + r) * WPR + bw0] = v;
                    }
                    if (lane < 3 * WPR - 32) {
                        const uint32_t m = (uint32_t)(bricks >> (20 * bg1)) & 0xFFFFFu;
                        const uint4 b = T.word_bits[bw1];
                        const uint32_t v = ((m & b.x) ? 96u : 0u) | ((m & b.y) ? 96u << 8 : 0u) | ((m & b.z) ? 96u << 16 : 0u) | ((m & b.w) ? 96u << 24 : 0u);
                        const int first = T.grp_first[bg1], count = T.grp_count[bg1];
                        for (int r = 0; r < count; ++r) wbuf[(first + r) * WPR + bw1] = v;
                    }
                    prev_bricks[i] = bricks;
                }
            }
            __syncwarp();
        }
        for (uint32_t seq = 0;; ++seq) {
        {
            const int q = seq % D;
            mbar_wait(&S.full[q], (seq / D) & 1);
            const uint32_t env0 = S.item_env0[q], n_here = S.item_n[q], s = S.item_step[q];
            if (n_here == 0u) break;                       // the physics warp found no more env batches
            if (rw == 0 && lane == 0 && seq == QLC_STAMP_SEQ(p)) QLC_STAMP(p, 8);
            RenderRec rr[NE];
            #pragma unroll
            for (int i = 0; i < NE; ++i) {
                const uint32_t j = (uint32_t)(i * R + rw);
                if (j < n_here) rr[i] = S.queue[q][j];
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&S.empty[q]);      // records are in registers: hand the queue slot back early
            const uint32_t slot = p.slot0 + s >= p.time_slots ? p.slot0 + s - p.time_slots : p.slot0 + s;
            uint8_t* slot_base = p.frames + ((size_t)slot * p.n_envs + env0) * FRAME_BYTES;
            #pragma unroll
            for (int g = 0; g < G; ++g) {
                // this group's frames were last stored G groups ago: that store must have left shared memory
                if (seq > 0 && lane == 0) bulk_wait_read<G - 1>();
                __syncwarp();
                // ---- phase 1: clear the old ball boxes, rewrite the paddle rows, rewrite the brick band where needed ----
                #pragma unroll
                for (int f = 0; f < FPG; ++f) {
                    const int i = g * FPG + f;
                    const uint32_t j = (uint32_t)(i * R + rw);
                    if (j < n_here) {
                        uint8_t* buf = my_bufs + (size_t)i * FRAME_BYTES;
                        uint32_t* wbuf = reinterpret_cast<uint32_t*>(buf);
                        const int pi0 = (int)(short)(prev_box[i] & 0xFFFF), pj0 = prev_box[i] >> 16;
                        // old ball pixels outside the paddle rows (rewritten below) and outside the brick rows (the band
                        // is rewritten whenever the old box touched it)
                        {
                            int x = pi0 + bdi0, y = pj0 + bdj0;
                            if (x >= 0 && x < FRAME_W && y >= 0 && y < FRAME_H && !T.row_pad[y] && T.row_grp[y] < 0) buf[y * FRAME_W + x] = 0;
                            if (lane < 4) {
                                x = pi0 + bdi1; y = pj0 + bdj1;
                                if (x >= 0 && x < FRAME_W && y >= 0 && y < FRAME_H && !T.row_pad[y] && T.row_grp[y] < 0) buf[y * FRAME_W + x] = 0;
                            }
                        }
                        // paddle row(s): every word rewritten (255 inside [x0, x1), else 0)
                        if (lane < WPR) {
                            const float x0 = rr[i].x0, x1 = rr[i].x1;
                            uint32_t v = 0u;
                            #pragma unroll
                            for (int b = 0; b < 4; ++b) { const float pc = (float)(4 * lane + b) + 0.5f; if (pc >= x0 && pc < x1) v |= 255u << (8 * b); }
                            for (int r = 0; r < pad_count; ++r) wbuf[(pad_first + r) * WPR + lane] = v;
                        }
                        // brick band: 3 row patterns x 21 words, replicated over the pixel rows of each brick row
                        const uint64_t bricks = rr[i].bricks;
                        if ((bricks != prev_bricks[i]) || (pj0 + 5 >= band_lo && pj0 <= band_hi)) {
                            {
                                const uint32_t m = (uint32_t)(bricks >> (20 * bg0)) & 0xFFFFFu;
                                const uint4 b = T.word_bits[bw0];
                                const uint32_t v = ((m & b.x) ? 96u : 0u) | ((m & b.y) ? 96u << 8 : 0u) | ((m & b.z) ? 96u << 16 : 0u) | ((m & b.w) ? 96u << 24 : 0u);
                                const int first = T.grp_first[bg0], count = T.grp_count[bg0];
                                for (int r = 0; r < count; ++r) wbuf[(first + r) * WPR + bw0] = v;
                            }
                            if (lane < 3 * WPR - 32) {
                                const uint32_t m = (uint32_t)(bricks >> (20 * bg1)) & 0xFFFFFu;
                                const uint4 b = T.word_bits[bw1];
                                const uint32_t v = ((m & b.x) ? 96u : 0u) | ((m & b.y) ? 96u << 8 : 0u) | ((m & b.z) ? 96u << 16 : 0u) | ((m & b.w) ? 96u << 24 : 0u);
                                const int first = T.grp_first[bg1], count = T.grp_count[bg1];
                                for (int r = 0; r < count; ++r) wbuf[(first + r) * WPR + bw1] = v;
                            }
                        }
                        prev_bricks[i] = bricks; prev_box[i] = rr[i].box;
                    }
                }
                __syncwarp();
                if (rw == 0 && lane == 0 && seq == QLC_STAMP_SEQ(p) && g == 0) QLC_STAMP(p, 9);
                // ---- phase 2: ball ring on top of bricks, under the paddle ----
                #pragma unroll
                for (int f = 0; f < FPG; ++f) {
                    const int i = g * FPG + f;
                    const uint32_t j = (uint32_t)(i * R + rw);
                    if (j < n_here) {
                        uint8_t* buf = my_bufs + (size_t)i * FRAME_BYTES;
                        const float bx = rr[i].bx, by = rr[i].by, x0 = rr[i].x0, x1 = rr[i].x1;
                        const int i0 = (int)(short)(rr[i].box & 0xFFFF), j0 = rr[i].box >> 16;
                        int x = i0 + bdi0, y = j0 + bdj0;
                        if (x >= 0 && x < FRAME_W && y >= 0 && y < FRAME_H) {
                            const float px = (float)x + 0.5f, dx = px - bx, dy = ((float)y + 0.5f) - by;
                            const float d2 = dx * dx + dy * dy;
                            if (d2 <= out2 && d2 >= in2 && !(T.row_pad[y] && px >= x0 && px < x1)) buf[y * FRAME_W + x] = 236;
                        }
                        if (lane < 4) {
                            x = i0 + bdi1; y = j0 + bdj1;
                            if (x >= 0 && x < FRAME_W && y >= 0 && y < FRAME_H) {
                                const float px = (float)x + 0.5f, dx = px - bx, dy = ((float)y + 0.5f) - by;
                                const float d2 = dx * dx + dy * dy;
                                if (d2 <= out2 && d2 >= in2 && !(T.row_pad[y] && px >= x0 && px < x1)) buf[y * FRAME_W + x] = 236;
                            }
                        }
                    }
                }
                if (rw == 0 && lane == 0 && seq == QLC_STAMP_SEQ(p) && g == 0) QLC_STAMP(p, 10);
                fence_proxy_async_smem();
                __syncwarp();
                if (rw == 0 && lane == 0 && seq == QLC_STAMP_SEQ(p) && g == 0) QLC_STAMP(p, 14);
                // ---- the group's frames leave as one bulk group (committed even when empty, so the count stays uniform) ----
                if (lane == 0) {
                    if (!QLC_DEBUG_SKIP_IS(p, 2u)) {
                        #pragma unroll
                        for (int f = 0; f < FPG; ++f) {
                            const int i = g * FPG + f;
                            const uint32_t j = (uint32_t)(i * R + rw);
                            if (j < n_here) bulk_store(slot_base + (size_t)j * FRAME_BYTES, my_bufs + (size_t)i * FRAME_BYTES, FRAME_BYTES);
                        }
                    }
                    bulk_commit();
                    if (rw == 0 && seq == QLC_STAMP_SEQ(p) && g == 0) QLC_STAMP(p, 15);
                }
            }
        }
        }   // batches
        if (rw == 0 && lane == 0) QLC_STAMP(p, 11);
        if (lane == 0) bulk_wait<0>();
        if (rw == 0 && lane == 0) { QLC_STAMP(p, 12); QLC_STAMP_GT(p, 13); }
        // (waiting only for the reads, cp.async.bulk.wait_group.read, measured the same: 10.87 vs 10.80 us per single-step launch)
    }
}

// ---------------------------------------------------------------------------------------------------------
// Kernel 2: reset (Environment::reset): mechanics = default(dir_x), frame stack = zeros (episode_step = 0)
// ---------------------------------------------------------------------------------------------------------
__global__ void env_reset_kernel(EnvArrays st, uint32_t n_envs, uint32_t env_id_base, uint64_t seed, const uint8_t* mask,
                                 const float* dir_x, int first_time) {
    const uint32_t e = blockIdx.x * blockDim.x + threadIdx.x;
    if (e >= n_envs) return;
    if (mask && !mask[e]) return;
    uint32_t episode = first_time ? 0u : st.episode[e] + 1u;
    Env env; env.err = first_time ? 0u : st.err[e];
    env_init(env, dir_x ? dir_x[e] : reset_dir_x(seed, env_id_base + e, episode));
    st.ball_cx[e] = env.cx; st.ball_cy[e] = env.cy; st.ball_dx[e] = env.dx; st.ball_dy[e] = env.dy;
    st.pad_min_x[e] = env.pmin; st.pad_max_x[e] = env.pmax; st.pad_speed[e] = env.pspeed;
    st.bricks[e] = env.bricks; st.score[e] = 0u; st.err[e] = env.err; st.finished[e] = 0;
    st.episode_step[e] = 0u; st.episode[e] = episode;
}

// ---------------------------------------------------------------------------------------------------------
// Distinct uniform index sampling (generate_distinct_random_ids, self_driving_tf_q_learner.rs:276-296) as a WARP routine.
// The reference's sequential rejection loop keeps the FIRST OCCURRENCES of the accepted draws, in stream order. Draw p of
// minibatch `call` is word p & 3 of philox({p >> 2, call_lo, call_hi, 'SAMP'}), mapped to [0, len) by Lemire's multiply-shift
// with rejection. The warp walks the stream 128 positions per round — lane L owns positions 4L..4L+3 of the round, the four
// words of ONE Philox block — and finds first occurrences with a shared-memory hash table (value, stream position): every
// accepted draw claims the slot of its value with ONE atomicCAS (the four claims of a lane are issued back to back); the claim
// that creates the entry is a first occurrence and records its position with a plain store, a claim that finds the value
// there is a duplicate — of an earlier round (dropped) or, rarely, of a draw of the same round, in which case the two settle
// by atomicMin on the position which of them comes first in the stream. Ranks come from a warp prefix sum of the per-lane
// counts. The routine runs INSIDE the gather kernels, so that a sampled minibatch is ONE kernel launch: every CTA derives the
// index of its own item and stops as soon as it has it. (Tried and dropped, profiles/r02_notes.md: one sampler per
// thread-block cluster of 8 handing indices out through distributed shared memory, a producer CTA publishing through global
// memory, an atomic-free insert - all slower than every CTA drawing for itself.) The index-only entry point
// (qlc_replay_sample) runs the same routine, one warp per minibatch, storing all of them.
// ---------------------------------------------------------------------------------------------------------
constexpr int SAMPLE_MAX_BATCH = 1024;
constexpr uint32_t SAMPLE_EMPTY = 0xFFFFFFFFu;           // never a value: len < 2^32 - 1 is checked by the host
constexpr uint32_t SAMPLE_MAX_ROUNDS = 1u << 19;         // x 128 stream positions: the bound of the "loop" in the reference

constexpr uint32_t SAMPLE_BLOCK_MIN_BATCH = 129;    // from this batch size on a CTA's warps walk the stream side by side (sample_distinct_block)
__host__ __device__ __forceinline__ uint32_t sample_table_size(uint32_t batch) {   // slots: power of two >= 4 * (batch + 128)
    // CTA-wide sampler: any size works (multiply-shift slot, wrap-around probing). 3,528 slots x 8 B = the 28,224 bytes of the four
    // staged frames the [b][x][y][slot] gather borrows the table from - a 32 KB table cost that kernel its 7th CTA per SM, i.e. a second
    // partial wave for one minibatch of 512 (1,024 CTAs). batch + 8 * 128 entries at most: load <= 0.58
    if (batch >= SAMPLE_BLOCK_MIN_BATCH) return 3528u;
    uint32_t n = 1024;
    while (n < 4u * (batch + 128u) && n < 4096u) n <<= 1;
    return n;                                                                       // <= 4096 slots x (value, position) = 32 KB
}

__device__ __forceinline__ void sample_table_clear(uint32_t* table, uint32_t tsize, int tid, int nthreads) {
    for (uint32_t i = tid; i < 2u * tsize / 4u; i += nthreads) reinterpret_cast<uint4*>(table)[i] = make_uint4(SAMPLE_EMPTY, SAMPLE_EMPTY, SAMPLE_EMPTY, SAMPLE_EMPTY);
}

// Walks the stream of minibatch `call` and hands every first occurrence of rank in [j_lo, j_hi] to deliver(rank, value) (called
// by the lane that owns the draw); returns once rank j_hi has been delivered. `table` = 2 * tsize words of shared memory, all
// SAMPLE_EMPTY (values in the first half, stream positions in the second).
template <class Deliver>
__device__ __forceinline__ void sample_distinct_warp(uint32_t* table, uint32_t tsize, uint32_t len, uint64_t seed, uint64_t call,
                                                     uint32_t j_lo, uint32_t j_hi, int lane, Deliver deliver) {
    const uint32_t thresh = (uint32_t)((0x100000000ull - (uint64_t)len) % (uint64_t)len);   // Lemire rejection zone
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    const uint32_t tmask = tsize - 1u;
    uint32_t* tval = table; uint32_t* tpos = table + tsize;
    uint32_t kept = 0;
    for (uint32_t round = 0; round < SAMPLE_MAX_ROUNDS; ++round) {
        const uint32_t ctr = round * 32u + (uint32_t)lane;
        const uint4 r = philox4x32_10(make_uint4(ctr, (uint32_t)call, (uint32_t)(call >> 32), STREAM_SAMPLE), key);
        const uint32_t raw[4] = {r.x, r.y, r.z, r.w};
        uint32_t val[4], slot[4], old[4]; bool valid[4];
        #pragma unroll
        for (int w = 0; w < 4; ++w) {
            const uint64_t m = (uint64_t)raw[w] * (uint64_t)len;
            valid[w] = !((uint32_t)m < thresh);
            val[w] = (uint32_t)(m >> 32);
            slot[w] = ((val[w] * 0x9E3779B1u) >> 12) & tmask;
        }
        #pragma unroll
        for (int w = 0; w < 4; ++w) old[w] = valid[w] ? atomicCAS(&tval[slot[w]], SAMPLE_EMPTY, val[w]) : SAMPLE_EMPTY;   // four claims in flight
        bool first[4], again[4];                          // first: my claim created the entry; again: the value was there already
        #pragma unroll
        for (int w = 0; w < 4; ++w) {
            first[w] = false; again[w] = false;
            if (valid[w]) {
                while (old[w] != SAMPLE_EMPTY && old[w] != val[w]) {       // the slot belongs to another value: probe on
                    slot[w] = (slot[w] + 1u) & tmask;
                    old[w] = atomicCAS(&tval[slot[w]], SAMPLE_EMPTY, val[w]);
                }
                if (old[w] == SAMPLE_EMPTY) { first[w] = true; tpos[slot[w]] = ctr * 4u + (uint32_t)w; }   // only the creator writes the position
                else again[w] = true;
            }
        }
        __syncwarp();
        // a value that was there already is a duplicate of an earlier ROUND (the usual case: drop it) or of a draw of THIS round,
        // whose creator is whichever claim won the race, not necessarily the earlier stream position: settle those (rare: two
        // equal values among 128 draws) by atomicMin on the position
        bool clash = false;
        #pragma unroll
        for (int w = 0; w < 4; ++w) if (again[w] && tpos[slot[w]] >= round * 128u) clash = true;
        if (__any_sync(0xFFFFFFFFu, clash)) {
            #pragma unroll
            for (int w = 0; w < 4; ++w) if ((first[w] || again[w]) && tpos[slot[w]] >= round * 128u) atomicMin(&tpos[slot[w]], ctr * 4u + (uint32_t)w);
            __syncwarp();
            #pragma unroll
            for (int w = 0; w < 4; ++w) first[w] = (first[w] || again[w]) && tpos[slot[w]] == ctr * 4u + (uint32_t)w;
        }
        uint32_t cnt = 0;
        #pragma unroll
        for (int w = 0; w < 4; ++w) cnt += first[w] ? 1u : 0u;
        uint32_t incl = cnt;                              // ordered prefix sum over the lanes (= over the stream positions)
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += v; }
        const uint32_t total = __shfl_sync(0xFFFFFFFFu, incl, 31);
        uint32_t rank = kept + incl - cnt;
        #pragma unroll
        for (int w = 0; w < 4; ++w) if (first[w]) { if (rank >= j_lo && rank <= j_hi) deliver(rank, val[w]); ++rank; }
        kept += total;
        if (kept > j_hi) return;
        __syncwarp();
    }
}

// the index of item j of the minibatch (every lane gets it)
__device__ __forceinline__ uint32_t sample_one(uint32_t* table, uint32_t tsize, uint32_t len, uint64_t seed, uint64_t call, uint32_t j, int lane) {
    uint32_t mine = SAMPLE_EMPTY;
    sample_distinct_warp(table, tsize, len, seed, call, j, j, lane, [&](uint32_t, uint32_t v) { mine = v; });
    const uint32_t hit = __ballot_sync(0xFFFFFFFFu, mine != SAMPLE_EMPTY);
    return hit ? __shfl_sync(0xFFFFFFFFu, mine, __ffs(hit) - 1) : 0u;      // no hit: unreachable for len >= batch (the reference would loop forever)
}

// CTA-wide form for big minibatches (batch >= SAMPLE_BLOCK_MIN_BATCH): the NW warps of the CTA take NW consecutive rounds of the
// stream at once (warp w of pass s = round s * NW + w), so a minibatch of 512 is ONE pass of 8 warps instead of five rounds of
// one. Every accepted draw claims the slot of its value (atomicCAS) and bids for it with its stream position (atomicMin); after
// the CTA barrier a draw is a first occurrence iff the slot still shows its own position - entries of earlier passes always win.
// Ranks: warp prefix sum, then the warp totals in round order. Same deliver(rank, value) contract as sample_distinct_warp.
template <int NW, class Deliver>
__device__ __forceinline__ void sample_distinct_block(uint32_t* table, uint32_t tsize, uint32_t len, uint64_t seed, uint64_t call,
                                                      uint32_t j_lo, uint32_t j_hi, int tid, uint32_t* warp_totals /* [NW] shared */, Deliver deliver) {
    const int warp = tid >> 5, lane = tid & 31;
    const uint32_t thresh = (uint32_t)((0x100000000ull - (uint64_t)len) % (uint64_t)len);
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    uint32_t* tval = table; uint32_t* tpos = table + tsize;
    uint32_t kept = 0;
    for (uint32_t pass = 0; pass < SAMPLE_MAX_ROUNDS / NW; ++pass) {
        const uint32_t ctr = (pass * NW + (uint32_t)warp) * 32u + (uint32_t)lane;
        const uint4 r = philox4x32_10(make_uint4(ctr, (uint32_t)call, (uint32_t)(call >> 32), STREAM_SAMPLE), key);
        const uint32_t raw[4] = {r.x, r.y, r.z, r.w};
        uint32_t val[4], slot[4]; bool valid[4];
        #pragma unroll
        for (int w = 0; w < 4; ++w) {
            const uint64_t m = (uint64_t)raw[w] * (uint64_t)len;
            valid[w] = !((uint32_t)m < thresh);
            val[w] = (uint32_t)(m >> 32);
            slot[w] = __umulhi(val[w] * 0x9E3779B1u, tsize);           // multiply-shift into [0, tsize): the size need not be a power of two
        }
        #pragma unroll
        for (int w = 0; w < 4; ++w) {
            if (valid[w]) {
                uint32_t old = atomicCAS(&tval[slot[w]], SAMPLE_EMPTY, val[w]);
                while (old != SAMPLE_EMPTY && old != val[w]) { slot[w] = slot[w] + 1u == tsize ? 0u : slot[w] + 1u; old = atomicCAS(&tval[slot[w]], SAMPLE_EMPTY, val[w]); }
                atomicMin(&tpos[slot[w]], ctr * 4u + (uint32_t)w);          // positions start as SAMPLE_EMPTY = the largest value
            }
        }
        __syncthreads();
        bool first[4]; uint32_t cnt = 0;
        #pragma unroll
        for (int w = 0; w < 4; ++w) { first[w] = valid[w] && tpos[slot[w]] == ctr * 4u + (uint32_t)w; cnt += first[w] ? 1u : 0u; }
        uint32_t incl = cnt;
        #pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xFFFFFFFFu, incl, o); if (lane >= o) incl += v; }
        if (lane == 31) warp_totals[warp] = incl;
        __syncthreads();
        uint32_t before = 0, all = 0;
        #pragma unroll
        for (int w = 0; w < NW; ++w) { const uint32_t t = warp_totals[w]; all += t; if (w < warp) before += t; }
        uint32_t rank = kept + before + incl - cnt;
        #pragma unroll
        for (int w = 0; w < 4; ++w) if (first[w]) { if (rank >= j_lo && rank <= j_hi) deliver(rank, val[w]); ++rank; }
        kept += all;
        if (kept > j_hi) return;                                            // the same decision in every thread
        __syncthreads();                                                    // warp_totals is rewritten by the next pass
    }
}

// index-only form: one CTA of NW warps per minibatch (NW = 1: the warp routine; NW = 8 for batch >= SAMPLE_BLOCK_MIN_BATCH)
template <int NW>
__global__ void __launch_bounds__(32 * NW) replay_sample_kernel(uint32_t* out, uint32_t batch, uint32_t len, uint64_t seed, uint64_t call0) {
    __shared__ __align__(16) uint32_t table[2 * 4096];
    __shared__ uint32_t warp_totals[NW];
    const int tid = threadIdx.x;
    const uint32_t tsize = sample_table_size(batch);
    sample_table_clear(table, tsize, tid, 32 * NW);
    uint32_t* dst = out + (size_t)blockIdx.x * batch;
    if (NW == 1) {
        __syncwarp();
        sample_distinct_warp(table, tsize, len, seed, call0 + blockIdx.x, 0u, batch - 1u, tid, [&](uint32_t rank, uint32_t v) { dst[rank] = v; });
    } else {
        __syncthreads();
        sample_distinct_block<NW>(table, tsize, len, seed, call0 + blockIdx.x, 0u, batch - 1u, tid, warp_totals, [&](uint32_t rank, uint32_t v) { dst[rank] = v; });
    }
}

// ---------------------------------------------------------------------------------------------------------
// Frame-stack gather (ReplayBuffer::get_many + batch_to_multi_dim_array, and Environment::state).
// A transition at time T of env e with k frames already in its episode uses frames F_{T-d}: state d=1..4,
// next d=0..3, valid iff d <= k; F_{T-d} sits in ring slot (k-d) mod 4 of the reference's FrameRingBuffer.
// Where an item comes from (GatherParams::mode):
//   GATHER_INDICES  logical replay indices given by the caller (get_many)
//   GATHER_CURRENT  item b = env b at the current time (Environment::state for the whole shard)
//   GATHER_HANDLES  state handles {time, k, env} (what a cloned BreakoutState is on the host: two integers, not pixels)
//   GATHER_SAMPLE   the kernel draws the distinct indices itself (minibatch = item / batch): sample + gather in one launch
// ---------------------------------------------------------------------------------------------------------
enum { GATHER_INDICES = 0, GATHER_CURRENT = 1, GATHER_HANDLES = 2, GATHER_SAMPLE = 3 };

struct ObsHandle { unsigned long long time; uint32_t k, env; };   // = qlc_obs_handle

struct GatherParams {
    const uint8_t* frames; const uint32_t* records; const uint32_t* episode_step;
    const uint32_t* indices; const ObsHandle* handles;
    uint32_t mode, n_items, n_envs, time_slots;
    uint64_t t_now, t_oldest;  // replay holds transitions of times [t_oldest, t_now)
    uint32_t sample_batch, sample_len; uint64_t seed, call0; uint32_t* idx_out;   // GATHER_SAMPLE
    uint32_t slices;           // [b][x][y][slot] kernel: CTAs per (item, state | next), each writes 1/slices of the pixels
    void* out_state; void* out_next;
    float* reward; uint8_t* action; uint8_t* done;
    // streamed host gathers (gather_xyh_stream_kernel, outputs in page-locked host memory): piece i raises cta_flags[i] = flag_value
    // (system scope, after its bulk store has been performed) so that host threads can consume it while the pieces behind it are
    // still crossing PCIe
    uint32_t* cta_flags; uint32_t flag_value; uint32_t preload;   // preload: copy the indices / handles into shared memory first
};

// false = the item does not exist (index >= len, handle whose frames have left the ring): every slot is zero-filled
__device__ __forceinline__ bool locate(const GatherParams& g, uint32_t b, uint32_t idx, uint64_t& T, uint32_t& e, uint32_t& k, uint32_t& rec) {
    rec = 0; k = 0; e = 0; T = g.t_now;
    if (g.mode == GATHER_CURRENT) {
        e = b;
        const uint32_t ks = g.episode_step[e];
        k = ks < 4u ? ks : 4u + (ks & 3u);
        return true;
    }
    if (g.mode == GATHER_HANDLES) {
        const ObsHandle h = g.handles[b];
        const uint32_t need = h.k < 4u ? h.k : 4u;                      // frames F_{T-1} .. F_{T-need} must still be in the ring
        if (h.env >= g.n_envs || h.time > g.t_now || h.time < need || h.time - need + g.time_slots < g.t_now) return false;
        T = h.time; e = h.env; k = h.k < 4u ? h.k : 4u + (h.k & 3u);
        return true;
    }
    T = g.t_oldest + idx / g.n_envs; e = idx % g.n_envs;
    if (T >= g.t_now) { T = g.t_now; e = 0; return false; }
    rec = g.records[(size_t)(T % g.time_slots) * g.n_envs + e];
    const uint32_t kmin = (rec >> 5) & 7u, kmod = (rec >> 3) & 3u;
    k = kmin < 4u ? kmin : 4u + kmod;    // any k' with k' mod 4 and min(k',4) preserved
    return true;
}

// which frame of the ring (time-slot * n_envs + env) holds ring slot h of item b's stack; ~0u = zero-filled slot
__device__ __forceinline__ uint4 slot_frames(const GatherParams& g, uint32_t b, uint32_t which) {
    uint64_t T; uint32_t e, k, rec;
    const bool exists = locate(g, b, g.mode == GATHER_INDICES ? g.indices[b] : 0u, T, e, k, rec);
    uint32_t f[4];
    #pragma unroll
    for (uint32_t h = 0; h < 4; ++h) {
        const uint32_t d = which ? ((k - h) & 3u) : (((k - h - 1u) & 3u) + 1u);
        f[h] = (exists && d <= k) ? (uint32_t)((T - d) % g.time_slots) * g.n_envs + e : 0xFFFFFFFFu;
    }
    return make_uint4(f[0], f[1], f[2], f[3]);
}

__device__ __forceinline__ void write_scalars(const GatherParams& g, uint32_t b, uint32_t rec) {
    if (g.reward) g.reward[b] = (float)((rec >> 8) & 0xFFu);
    if (g.action) g.action[b] = (uint8_t)(rec & 3u);
    if (g.done) g.done[b] = (uint8_t)((rec >> 2) & 1u);
}

// u8 [b][slot][y][x]: one warp per item; <= 5 distinct frames in, 8 frames out, all as 7,056-byte bulk copies. NW > 1 (sampled
// minibatches of >= SAMPLE_BLOCK_MIN_BATCH): NW - 1 more warps help with the index draw and leave.
template <int NW>
__global__ void __launch_bounds__(32 * NW) gather_u8_kernel(GatherParams g) {
    extern __shared__ __align__(128) uint8_t sm[];     // 5 frames + 1 zero frame; the sampler's hash table borrows the first <= 32 KB
    __shared__ uint64_t bar;
    __shared__ uint32_t warp_totals[NW], s_idx;
    const int tid = threadIdx.x, lane = tid & 31;
    const uint32_t b = blockIdx.x;
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next kernel's prologue may overlap this grid (see env_advance_kernel)
    uint8_t* zero = sm + 5 * FRAME_BYTES;
    for (int i = tid; i < FRAME_VEC16; i += 32 * NW) reinterpret_cast<uint4*>(zero)[i] = make_uint4(0, 0, 0, 0);
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    uint32_t idx = 0;
    uint32_t* table = reinterpret_cast<uint32_t*>(sm);
    if (g.mode == GATHER_SAMPLE) {
        const uint32_t tsize = sample_table_size(g.sample_batch);
        sample_table_clear(table, tsize, tid, 32 * NW);
        const uint32_t mb = b / g.sample_batch, j = b - mb * g.sample_batch;
        if (NW == 1) {
            __syncwarp();
            idx = sample_one(table, tsize, g.sample_len, g.seed, g.call0 + mb, j, lane);
        } else {
            __syncthreads();
            sample_distinct_block<NW>(table, tsize, g.sample_len, g.seed, g.call0 + mb, j, j, tid, warp_totals, [&](uint32_t, uint32_t v) { s_idx = v; });
            fence_proxy_async_smem();                    // the helpers' share of the zero frame and of the table, before the bulk copies
            __syncthreads();
            idx = s_idx;
        }
    }
    if (NW > 1 && g.mode != GATHER_SAMPLE) { fence_proxy_async_smem(); __syncthreads(); }   // (not launched this way: keeps the zero frame whole)
    if (NW > 1 && tid >= 32) return;                     // the helper warps are done
    // launched with programmatic stream serialization: everything above neither reads nor writes anything an earlier kernel touches
    asm volatile("griddepcontrol.wait;" ::: "memory");
    if (g.mode == GATHER_SAMPLE && lane == 0 && g.idx_out) g.idx_out[b] = idx;
    if (g.mode == GATHER_INDICES) idx = g.indices[b];
    uint64_t T; uint32_t e, k, rec;
    const bool exists = locate(g, b, idx, T, e, k, rec);
    fence_proxy_async_smem();                            // zero frame / hash table (generic proxy) before the bulk copies (async proxy)
    __syncwarp();
    if (lane == 0) {
        write_scalars(g, b, rec);
        if (!g.out_state && !g.out_next) return;         // scalars only (get_many without tensorisation)
        const int d_lo = g.out_next ? 0 : 1, d_hi = g.out_state ? 4 : 3;
        uint32_t bytes = 0;
        for (int d = d_lo; d <= d_hi; ++d) if (exists && (uint32_t)d <= k) bytes += FRAME_BYTES;
        if (bytes) {
            mbar_expect_tx(&bar, bytes);
            for (int d = d_lo; d <= d_hi; ++d)
                if ((uint32_t)d <= k) {
                    const uint64_t Tf = T - (uint64_t)d;
                    bulk_load(sm + d * FRAME_BYTES, g.frames + ((size_t)(Tf % g.time_slots) * g.n_envs + e) * FRAME_BYTES, FRAME_BYTES, &bar);
                }
            mbar_wait(&bar, 0);
        }
        for (int h = 0; h < 4; ++h) {
            if (g.out_state) {
                const uint32_t d = ((k - h - 1u) & 3u) + 1u;
                bulk_store((uint8_t*)g.out_state + ((size_t)b * 4 + h) * FRAME_BYTES, (exists && d <= k) ? sm + d * FRAME_BYTES : zero, FRAME_BYTES);
            }
            if (g.out_next) {
                const uint32_t d = (k - h) & 3u;
                bulk_store((uint8_t*)g.out_next + ((size_t)b * 4 + h) * FRAME_BYTES, (exists && d <= k) ? sm + d * FRAME_BYTES : zero, FRAME_BYTES);
            }
        }
        bulk_commit();
        bulk_wait<0>();
    }
}

// [b][x][y][slot] (the reference's ToMultiDimArray layout), as f32 (value = u8 as f32) or as u8: `slices` CTAs per (item, state |
// next); 4 slot frames staged in shared memory by bulk copies, then a conflict-free transposing read (row stride 84 B = 21
// words) and one coalesced 16-byte (f32) / 4-byte (u8) store per pixel of the CTA's slice. More slices = more, smaller CTAs: a
// single minibatch of 32 spreads over every SM, and a big call is balanced by the block scheduler (the SMs' store pipes are the
// limit: 6 instead of 7 resident CTAs on an SM would leave it idle for the last seventh of the kernel).
constexpr int GATHER_XYH_THREADS = 256;
__device__ __forceinline__ void store_pixel(float4* o, int idx, uint8_t a, uint8_t b, uint8_t c, uint8_t d) { o[idx] = make_float4((float)a, (float)b, (float)c, (float)d); }
__device__ __forceinline__ void store_pixel(uchar4* o, int idx, uint8_t a, uint8_t b, uint8_t c, uint8_t d) { o[idx] = make_uchar4(a, b, c, d); }

template <class Px>
__global__ void __launch_bounds__(GATHER_XYH_THREADS) gather_xyh_kernel(GatherParams g) {
    extern __shared__ __align__(128) uint8_t sm[];     // 4 slot frames; the sampler's hash table borrows the first <= 32 KB
    __shared__ uint64_t bar;
    __shared__ uint32_t s_idx, warp_totals[GATHER_XYH_THREADS / 32];
    const int tid = threadIdx.x;
    const uint32_t unit = blockIdx.x / g.slices, slice = blockIdx.x - unit * g.slices;
    const uint32_t b = unit >> 1, which = unit & 1u;   // 0 = state, 1 = next
    Px* out = reinterpret_cast<Px*>(which ? g.out_next : g.out_state);
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");   // the next kernel's prologue may overlap this grid (see env_advance_kernel)
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    uint32_t idx = 0;
    uint32_t* table = reinterpret_cast<uint32_t*>(sm);
    if (g.mode == GATHER_SAMPLE) {
        const uint32_t tsize = sample_table_size(g.sample_batch);
        const uint32_t mb = b / g.sample_batch, j = b - mb * g.sample_batch;
        sample_table_clear(table, tsize, tid, GATHER_XYH_THREADS);
        __syncthreads();
        if (g.sample_batch >= SAMPLE_BLOCK_MIN_BATCH) {   // all 8 warps walk the stream side by side
            sample_distinct_block<GATHER_XYH_THREADS / 32>(table, tsize, g.sample_len, g.seed, g.call0 + mb, j, j, tid, warp_totals, [&](uint32_t, uint32_t v) { s_idx = v; });
        } else if (tid < 32) {
            const uint32_t v = sample_one(table, tsize, g.sample_len, g.seed, g.call0 + mb, j, tid);
            if (tid == 0) s_idx = v;
        }
        __syncthreads();
        idx = s_idx;
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");   // programmatic stream serialization: nothing above touches an earlier kernel's data
    if (g.mode == GATHER_SAMPLE && tid == 0 && slice == 0 && g.idx_out && (which == 0 || !g.out_state)) g.idx_out[b] = idx;
    if (g.mode == GATHER_INDICES) idx = g.indices[b];
    uint64_t T; uint32_t e, k, rec;
    const bool exists = locate(g, b, idx, T, e, k, rec);
    if (tid == 0 && slice == 0 && (which == 0 || !g.out_state)) write_scalars(g, b, rec);
    if (!out) return;
    // zero-fill the slots that have no frame yet
    uint32_t dsl[4];
    #pragma unroll
    for (int h = 0; h < 4; ++h) {
        dsl[h] = which ? ((k - h) & 3u) : (((k - h - 1u) & 3u) + 1u);
        if (!exists) dsl[h] = 0xFFFFu;
        if (dsl[h] > k)
            for (int i = tid; i < FRAME_VEC16; i += GATHER_XYH_THREADS) reinterpret_cast<uint4*>(sm + h * FRAME_BYTES)[i] = make_uint4(0, 0, 0, 0);
    }
    fence_proxy_async_smem();
    __syncthreads();
    uint32_t bytes = 0;
    #pragma unroll
    for (int h = 0; h < 4; ++h) if (dsl[h] <= k) bytes += FRAME_BYTES;
    if (bytes) {
        if (tid == 0) {
            mbar_expect_tx(&bar, bytes);
            for (int h = 0; h < 4; ++h)
                if (dsl[h] <= k) {
                    const uint64_t Tf = T - (uint64_t)dsl[h];
                    bulk_load(sm + h * FRAME_BYTES, g.frames + ((size_t)(Tf % g.time_slots) * g.n_envs + e) * FRAME_BYTES, FRAME_BYTES, &bar);
                }
        }
        mbar_wait(&bar, 0);
    }
    Px* o = out + (size_t)b * FRAME_BYTES;
    const int per_slice = FRAME_BYTES / (int)g.slices;     // slices in {1, 2, 4}: 7,056 = 4 * 1,764
    for (int i = (int)slice * per_slice + tid; i < ((int)slice + 1) * per_slice; i += GATHER_XYH_THREADS) {
        const int x = i / FRAME_H, y = i - x * FRAME_H;
        const int src = y * FRAME_W + x;
        store_pixel(o, i, sm[src], sm[FRAME_BYTES + src], sm[2 * FRAME_BYTES + src], sm[3 * FRAME_BYTES + src]);
    }
}

// Streamed host gather (f32 [b][x][y][slot] requests of the *_host entry points): the u8 stacks go straight into page-locked HOST
// memory, in order, and every piece (1/slices of a stack) raises an arrival flag at system scope, so that host threads widen
// piece i into the caller's tensor while the pieces behind it are still crossing PCIe. A small persistent grid walks the pieces
// with stride gridDim.x, two stores in flight per CTA. Stores of different CTAs interleave on the bus, so what is in flight completes
// together: one CTA per piece raised its first flag 55 us after the launch (the whole transfer), 48 CTAs raise it after ~30 us; an
// issue window that limits the pieces in flight makes the arrival order strict but starves the bus (12 pieces: 128 us instead of 98).
// Each piece is transposed in shared memory and leaves as ONE bulk copy.
__global__ void __launch_bounds__(GATHER_XYH_THREADS) gather_xyh_stream_kernel(GatherParams g) {
    extern __shared__ __align__(128) uint8_t sm[];     // 4 slot frames + the transposed piece
    __shared__ uint64_t bar;
    const int tid = threadIdx.x;
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); }
    __syncthreads();
    const int per_slice = FRAME_BYTES / (int)g.slices;
    const bool both = g.out_state && g.out_next;         // pieces are numbered over the stacks that are wanted: no gaps in the arrival order
    const uint32_t n_pieces = g.n_items * (both ? 2u : 1u) * g.slices;
    // The indices / handles sit in page-locked host memory: fetch them ONCE, before any store is in flight - a PCIe read issued
    // per piece queues up behind the outbound frame data (measured: ~13 us per piece, whatever its size).
    const uint32_t* my_indices = g.indices; const ObsHandle* my_handles = g.handles;
    if (g.preload) {
        uint8_t* pre = sm + 4 * FRAME_BYTES + 2 * (size_t)per_slice * 4u;
        if (g.mode == GATHER_INDICES) {
            for (uint32_t i = tid; i < g.n_items; i += GATHER_XYH_THREADS) reinterpret_cast<uint32_t*>(pre)[i] = g.indices[i];
            my_indices = reinterpret_cast<const uint32_t*>(pre);
        } else if (g.mode == GATHER_HANDLES) {
            for (uint32_t i = tid; i < g.n_items; i += GATHER_XYH_THREADS) reinterpret_cast<uint4*>(pre)[i] = reinterpret_cast<const uint4*>(g.handles)[i];
            my_handles = reinterpret_cast<const ObsHandle*>(pre);
        }
        __syncthreads();
    }
    GatherParams gl = g; gl.handles = my_handles;
    uint32_t phase = 0, it = 0;
    uint32_t prev_piece = 0xFFFFFFFFu;                   // thread 0: the piece whose store is in flight, its flag not raised yet
    for (uint32_t piece = blockIdx.x; piece < n_pieces; piece += gridDim.x) {
        uchar4* stg = reinterpret_cast<uchar4*>(sm + 4 * FRAME_BYTES + (size_t)(it & 1u) * (size_t)per_slice * 4u);   // two staging buffers
        ++it;
        const uint32_t unit = piece / g.slices, slice = piece - unit * g.slices;
        const uint32_t b = both ? unit >> 1 : unit, which = both ? unit & 1u : (g.out_next ? 1u : 0u);
        uchar4* out = reinterpret_cast<uchar4*>(which ? g.out_next : g.out_state);
        uint64_t T; uint32_t e, k, rec;
        const bool exists = locate(gl, b, g.mode == GATHER_INDICES ? my_indices[b] : 0u, T, e, k, rec);
        if (tid == 0 && slice == 0 && (which == 0 || !g.out_state)) write_scalars(g, b, rec);
        uint32_t dsl[4];
        #pragma unroll
        for (int h = 0; h < 4; ++h) {
            dsl[h] = which ? ((k - h) & 3u) : (((k - h - 1u) & 3u) + 1u);
            if (!exists) dsl[h] = 0xFFFFu;
            if (dsl[h] > k)
                for (int i = tid; i < FRAME_VEC16; i += GATHER_XYH_THREADS) reinterpret_cast<uint4*>(sm + h * FRAME_BYTES)[i] = make_uint4(0, 0, 0, 0);
        }
        fence_proxy_async_smem();
        __syncthreads();                                 // also: thread 0 is back from the previous piece's store
        uint32_t bytes = 0;
        #pragma unroll
        for (int h = 0; h < 4; ++h) if (dsl[h] <= k) bytes += FRAME_BYTES;
        if (bytes) {
            if (tid == 0) {
                mbar_expect_tx(&bar, bytes);
                for (int h = 0; h < 4; ++h)
                    if (dsl[h] <= k) {
                        const uint64_t Tf = T - (uint64_t)dsl[h];
                        bulk_load(sm + h * FRAME_BYTES, g.frames + ((size_t)(Tf % g.time_slots) * g.n_envs + e) * FRAME_BYTES, FRAME_BYTES, &bar);
                    }
            }
            mbar_wait(&bar, phase);
            phase ^= 1u;
        }
        for (int i = (int)slice * per_slice + tid, q = tid; q < per_slice; i += GATHER_XYH_THREADS, q += GATHER_XYH_THREADS) {
            const int x = i / FRAME_H, y = i - x * FRAME_H;
            const int src = y * FRAME_W + x;
            stg[q] = make_uchar4(sm[src], sm[FRAME_BYTES + src], sm[2 * FRAME_BYTES + src], sm[3 * FRAME_BYTES + src]);
        }
        fence_proxy_async_smem();
        __syncthreads();
        if (tid == 0) {
            bulk_store(out + (size_t)b * FRAME_BYTES + (size_t)slice * per_slice, stg, (uint32_t)per_slice * 4u);
            bulk_commit();
            // this piece's store stays in flight while the next piece is loaded and transposed (into the other staging buffer); the
            // piece BEFORE it has been performed by now ...
            bulk_wait<1>();
            if (prev_piece != 0xFFFFFFFFu) {
                __threadfence_system();
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(g.cta_flags + prev_piece), "r"(g.flag_value) : "memory");   // ... so its flag goes up
            }
            prev_piece = piece;
        }
    }
    if (tid == 0 && prev_piece != 0xFFFFFFFFu) {
        bulk_wait<0>();
        __threadfence_system();
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(g.cta_flags + prev_piece), "r"(g.flag_value) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------
// small utility kernels
// ---------------------------------------------------------------------------------------------------------
__global__ void action_histogram_kernel(const uint32_t* records, uint32_t n_envs, uint32_t time_slots, uint64_t t_oldest, uint64_t t_now,
                                        unsigned long long* counts) {
    const uint64_t total = (t_now - t_oldest) * n_envs;
    unsigned long long c0 = 0, c1 = 0, c2 = 0;
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t T = t_oldest + i / n_envs; const uint32_t e = (uint32_t)(i % n_envs);
        const uint32_t a = records[(size_t)(T % time_slots) * n_envs + e] & 3u;
        c0 += a == 0; c1 += a == 1; c2 += a == 2;
    }
    for (int o = 16; o; o >>= 1) { c0 += __shfl_down_sync(~0u, c0, o); c1 += __shfl_down_sync(~0u, c1, o); c2 += __shfl_down_sync(~0u, c2, o); }
    if ((threadIdx.x & 31) == 0) { atomicAdd(&counts[0], c0); atomicAdd(&counts[1], c1); atomicAdd(&counts[2], c2); }
}

// "lives" of the reference game: the episode ends the first time the ball passes the paddle (mechanics.rs:131-135), so an
// env has exactly one life while it is not finished
__global__ void lives_kernel(const uint8_t* finished, uint8_t* lives, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) lives[i] = finished[i] ? 0 : 1;
}

__global__ void stats_export_kernel(const DeviceStats* s, uint64_t steps, double* out) {
    out[0] = (double)s->sum_return; out[1] = (double)s->episodes; out[2] = (double)steps;
    out[3] = s->episodes ? -(double)s->min_return : -1.0e300; out[4] = s->episodes ? (double)s->max_return : -1.0e300;
}
// all-gathered per-rank vectors [world][5] -> {sum, sum, sum, max, max} (one collective instead of a sum- and a max-all-reduce)
__global__ void stats_combine_kernel(const double* gathered, int world, double* out) {
    double a = 0.0, b = 0.0, c = 0.0, d = -1.0e300, e = -1.0e300;
    for (int r = 0; r < world; ++r) {
        const double* v = gathered + 5 * r;
        a += v[0]; b += v[1]; c += v[2]; d = v[3] > d ? v[3] : d; e = v[4] > e ? v[4] : e;
    }
    out[0] = a; out[1] = b; out[2] = c; out[3] = d; out[4] = e;
}

__global__ void err_or_kernel(const uint32_t* err, uint32_t n, uint32_t* out) {
    uint32_t v = 0;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) v |= err[i];
    for (int o = 16; o; o >>= 1) v |= __shfl_down_sync(~0u, v, o);
    if ((threadIdx.x & 31) == 0 && v) atomicOr(out, v);
}

// known-answer entry points (device collision routines, one thread)
__global__ void debug_collision_kernel(int which, float cx, float cy, float r, float mvx, float mvy, float minx, float miny, float maxx, float maxy,
                                       float* out /* some, way, approx, nx, ny */, uint32_t* err_out) {
    uint32_t err = 0; Surface s; s.way = s.approx = s.nx = s.ny = 0.0f; bool some = false;
    if (which == 3) {
        const float len = length2(mvx, mvy);
        some = sweep_ball_box(cx, cy, r, mvx, mvy, len, (minx + maxx) / 2.0f, (miny + maxy) / 2.0f, (maxx - minx) / 2.0f, (maxy - miny) / 2.0f, s, err);
    } else {
        float d; bool hit; float f;
        if (which == 0)      { d = cx - r;          hit = !(d + mvx > 0.0f); f = d / fabsf(mvx); s.nx = 1.0f; }
        else if (which == 1) { d = GRID_X - cx - r; hit = !(mvx < d);        f = d / fabsf(mvx); s.nx = -1.0f; }
        else                 { d = cy - r - 0.0f;   hit = !(d + mvy > 0.0f); f = d / fabsf(mvy); s.ny = 1.0f; }
        if (!(d >= 0.0f)) err |= ENVERR_WALL_DISTANCE;
        some = hit;
        if (hit) s.way = length2(mvx * f, mvy * f);
    }
    out[0] = some ? 1.0f : 0.0f; out[1] = s.way; out[2] = s.approx; out[3] = s.nx; out[4] = s.ny; *err_out = err;
}

// batched form of the above for fuzzing the device collision code against the oracle: in [n][9] = cx, cy, r, mvx, mvy,
// minx, miny, maxx, maxy; out [n][6] = some, way, approx, nx, ny, err (err as raw bits)
__global__ void debug_collision_batch_kernel(const float* in, float* out, uint32_t n) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float* a = in + (size_t)i * 9;
    uint32_t err = 0; Surface s; s.way = s.approx = s.nx = s.ny = 0.0f;
    const float len = length2(a[3], a[4]);
    const bool some = sweep_ball_box(a[0], a[1], a[2], a[3], a[4], len, (a[5] + a[7]) / 2.0f, (a[6] + a[8]) / 2.0f, (a[7] - a[5]) / 2.0f, (a[8] - a[6]) / 2.0f, s, err);
    float* o = out + (size_t)i * 6;
    o[0] = some ? 1.0f : 0.0f; o[1] = s.way; o[2] = s.approx; o[3] = s.nx; o[4] = s.ny; o[5] = __uint_as_float(err);
}

}  // namespace qlc
