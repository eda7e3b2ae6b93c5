/* TEST INFRASTRUCTURE ONLY — CPU baseline timing legs built on the oracle ("restated reference CPU path").
 * Used by bench.py (cpu_baseline and --impl reference) and nothing else. */
#include "breakout_oracle.h"
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <omp.h>

typedef struct orc_vec orc_vec;
typedef struct orc_replay orc_replay;
orc_vec* orc_vec_new(uint32_t, uint64_t, uint32_t, uint32_t, size_t, size_t);
void orc_vec_free(orc_vec*);
void orc_vec_step_range(orc_vec*, uint32_t, uint32_t, const uint8_t*, float*, uint8_t*);
void orc_vec_step(orc_vec*, const uint8_t*, float*, uint8_t*);
orc_replay* orc_vec_replay(orc_vec*);
size_t orc_replay_len(const orc_replay*);
void orc_replay_get_many_f32(const orc_replay*, const uint32_t*, size_t, float*, float*, float*, uint8_t*, uint8_t*);
int orc_sample_distinct(uint64_t, uint64_t, uint32_t, uint32_t, uint32_t*);
uint8_t orc_synthetic_action(uint64_t, uint32_t, uint32_t);
orc_env* orc_vec_env(orc_vec*, uint32_t);

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec; }

/* env leg: n_envs envs x n_steps steps, random policy, step -> draw -> grayscale -> ring add -> per-step state
 * clone into an Rc (prelude.rs:52-58), auto-reset; envs split over `threads` host threads. Returns seconds. */
double orc_bench_env_steps(uint32_t n_envs, uint32_t n_steps, uint64_t seed, int threads, uint64_t* checksum) {
    orc_vec* v = orc_vec_new(n_envs, seed, 0, 0, 0, 0);
    uint8_t* actions = (uint8_t*)malloc((size_t)n_envs * n_steps);
    for (uint32_t t = 0; t < n_steps; ++t) for (uint32_t e = 0; e < n_envs; ++e) actions[(size_t)t * n_envs + e] = orc_synthetic_action(seed, e, t);
    float* reward = (float*)malloc(n_envs * sizeof(float)); uint8_t* done = (uint8_t*)malloc(n_envs);
    if (threads < 1) threads = 1;
    uint64_t sum = 0;
    double t0 = now_s();
    #pragma omp parallel num_threads(threads) reduction(+:sum)
    {
        int tid = omp_get_thread_num(), nt = omp_get_num_threads();
        uint32_t e0 = (uint32_t)((uint64_t)n_envs * tid / nt), e1 = (uint32_t)((uint64_t)n_envs * (tid + 1) / nt);
        orc_state* clone = NULL;
        for (uint32_t t = 0; t < n_steps; ++t) {
            orc_vec_step_range(v, e0, e1, actions + (size_t)t * n_envs, reward, done);
            for (uint32_t e = e0; e < e1; ++e) {       /* step_as_rc: clone the state (4 frames + mechanics) */
                free(clone);
                clone = (orc_state*)malloc(sizeof(orc_state));
                memcpy(clone, &orc_vec_env(v, e)->state, sizeof(orc_state));
                sum += clone->frame_buffer.buffer[0][40 * 84 + 40] + (uint64_t)reward[e];
            }
        }
        free(clone);
    }
    double t1 = now_s();
    if (checksum) *checksum = sum;
    free(actions); free(reward); free(done); orc_vec_free(v);
    return t1 - t0;
}

/* replay leg: fill a replay of `capacity` transitions from n_envs envs, then time `n_batches` x
 * (generate_distinct_random_ids + get_many + batch_to_multi_dim_array for state and state_next). */
double orc_bench_sample(uint32_t n_envs, size_t capacity, uint32_t batch, uint32_t n_batches, uint64_t seed, uint64_t* checksum) {
    orc_vec* v = orc_vec_new(n_envs, seed, 0, 0, capacity, 100);
    uint8_t* actions = (uint8_t*)malloc(n_envs);
    float* reward = (float*)malloc(n_envs * sizeof(float)); uint8_t* done = (uint8_t*)malloc(n_envs);
    uint32_t fill_steps = (uint32_t)((capacity + n_envs - 1) / n_envs);
    for (uint32_t t = 0; t < fill_steps; ++t) {
        for (uint32_t e = 0; e < n_envs; ++e) actions[e] = orc_synthetic_action(seed, e, t);
        orc_vec_step(v, actions, reward, done);
    }
    const size_t per = (size_t)84 * 84 * 4;
    float* s = (float*)malloc(batch * per * sizeof(float)); float* sn = (float*)malloc(batch * per * sizeof(float));
    float* r = (float*)malloc(batch * sizeof(float)); uint8_t* a = (uint8_t*)malloc(batch); uint8_t* d = (uint8_t*)malloc(batch);
    uint32_t* idx = (uint32_t*)malloc(batch * sizeof(uint32_t));
    uint64_t sum = 0;
    orc_replay* rp = orc_vec_replay(v);
    double t0 = now_s();
    for (uint32_t c = 0; c < n_batches; ++c) {
        orc_sample_distinct(seed, c, (uint32_t)orc_replay_len(rp), batch, idx);
        orc_replay_get_many_f32(rp, idx, batch, s, sn, r, a, d);
        sum += (uint64_t)s[per / 2] + (uint64_t)sn[per / 3] + idx[0];
    }
    double t1 = now_s();
    if (checksum) *checksum = sum;
    free(s); free(sn); free(r); free(a); free(d); free(idx); free(actions); free(reward); free(done); orc_vec_free(v);
    return t1 - t0;
}

/* full-size parity leg (tests only): envs never interact (mechanics.rs has no shared state) and these sub-shards carry no
 * replay, so a shard of independent envs is stepped as `n_parts` sub-shards with the same global env ids, one host thread
 * each. actions / reward / done are [n_steps][n_total]; part p owns columns [offsets[p], offsets[p] + its n_envs). */
void orc_parts_run(orc_vec** parts, const uint32_t* offsets, const uint32_t* sizes, int n_parts, uint32_t n_total, uint32_t n_steps,
                   const uint8_t* actions, float* reward, uint8_t* done) {
    #pragma omp parallel for schedule(dynamic, 1)
    for (int p = 0; p < n_parts; ++p)
        for (uint32_t t = 0; t < n_steps; ++t) {
            const size_t row = (size_t)t * n_total + offsets[p];
            orc_vec_step_range(parts[p], 0, sizes[p], actions + row, reward + row, done + row);
        }
}

/* [n_steps][n_envs] block of the synthetic random policy stream (same values as orc_synthetic_action one by one) */
void orc_synthetic_actions_fill(uint64_t seed, uint32_t env_id_base, uint32_t n_envs, uint32_t t0, uint32_t n_steps, uint8_t* out) {
    #pragma omp parallel for
    for (uint32_t t = 0; t < n_steps; ++t)
        for (uint32_t e = 0; e < n_envs; ++e) out[(size_t)t * n_envs + e] = orc_synthetic_action(seed, env_id_base + e, t0 + t);
}

/* actor-loop leg (BASELINE configs[4] without the model): the reference's learn_episode data path for ONE env on one thread
 * (self_driving_tf_q_learner.rs:141-233) — step_as_rc (step, draw, grayscale, ring add, state clone), ReplayBuffer::add, and on
 * every 4th step once len > batch (:181) generate_distinct_random_ids + get_many + batch_to_multi_dim_array for state and
 * state_next (f32 [b][x][y][slot]); random policy, auto-reset like the episode loop. Returns seconds for n_steps env-steps. */
double orc_bench_actor_loop(uint32_t n_steps, size_t capacity, uint32_t batch, uint64_t seed, uint64_t* checksum) {
    orc_vec* v = orc_vec_new(1, seed, 0, 0, capacity, 100);
    const size_t per = (size_t)84 * 84 * 4;
    float* s = (float*)malloc(batch * per * sizeof(float)); float* sn = (float*)malloc(batch * per * sizeof(float));
    float* r = (float*)malloc(batch * sizeof(float)); uint8_t* a = (uint8_t*)malloc(batch); uint8_t* d = (uint8_t*)malloc(batch);
    uint32_t* idx = (uint32_t*)malloc(batch * sizeof(uint32_t));
    orc_replay* rp = orc_vec_replay(v);
    uint64_t sum = 0, calls = 0;
    double t0 = now_s();
    for (uint32_t t = 0; t < n_steps; ++t) {
        uint8_t act = orc_synthetic_action(seed, 0, t); float rew; uint8_t dn;
        orc_vec_step(v, &act, &rew, &dn);
        if (t % 4 == 0 && orc_replay_len(rp) > batch) {
            orc_sample_distinct(seed, calls++, (uint32_t)orc_replay_len(rp), batch, idx);
            orc_replay_get_many_f32(rp, idx, batch, s, sn, r, a, d);
            sum += (uint64_t)s[per / 2] + (uint64_t)sn[per / 3] + idx[0];
        }
        sum += (uint64_t)rew;
    }
    double t1 = now_s();
    if (checksum) *checksum = sum + calls;
    free(s); free(sn); free(r); free(a); free(d); free(idx); orc_vec_free(v);
    return t1 - t0;
}
