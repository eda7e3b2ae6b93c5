/* c_abi_latency.c — what the drop-in calls cost from a COMPILED host (plain C against include/ql_cuda.h, the way a Rust `ql-cuda`
 * crate binds them): no interpreter, no ctypes. One env stepped one call at a time (the unchanged learner's pattern), then the
 * learner's replay calls on a 4,096-env shard: state handles -> f32 tensor (`batch_to_multi_dim_array`), and
 * get_many(indices) -> f32 state + state_next tensors in host memory.
 *   gcc -O2 -I include tools/cabi/c_abi_latency.c -L q-learning_b200 -lqlcuda -Wl,-rpath,'$ORIGIN/../../q-learning_b200' -o tools/cabi/c_abi_latency */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>

#include "ql_cuda.h"

#define CHECK(x) do { int32_t rc_ = (x); if (rc_) { fprintf(stderr, "%s failed (%d): %s\n", #x, rc_, qlc_last_error_string()); return 1; } } while (0)

static double now_us(void) { struct timespec t; clock_gettime(CLOCK_MONOTONIC, &t); return t.tv_sec * 1e6 + t.tv_nsec * 1e-3; }

static uint64_t rng_state = 88172645463325252ull;
static uint32_t rnd(void) { rng_state ^= rng_state << 13; rng_state ^= rng_state >> 7; rng_state ^= rng_state << 17; return (uint32_t)(rng_state >> 11); }

static int json_mode = 0, n_res = 0;
static struct { const char* key; double us; } results[16];
static void report(const char* key, const char* text, double us) {
    results[n_res].key = key; results[n_res].us = us; ++n_res;
    if (!json_mode) printf("%-78s %9.2f us\n", text, us);
}

int main(int argc, char** argv) {
    json_mode = argc > 1 && strcmp(argv[1], "--json") == 0;
    qlc_config cfg; memset(&cfg, 0, sizeof cfg);
    cfg.struct_size = sizeof cfg; cfg.device = 0; cfg.frame_w = 84; cfg.frame_h = 84; cfg.seed = 7; cfg.episode_window = 100;
    /* ---- one env, one step per call (reference: Environment::step, auto_reset off: the learner resets) ---- */
    {
        qlc_env* env = NULL;
        cfg.n_envs = 1; cfg.replay_capacity = 50000; cfg.auto_reset = 1;
        CHECK(qlc_env_create(&cfg, &env));
        uint8_t* a; float* r; uint8_t* d;
        CHECK(qlc_host_alloc(64, (void**)&a)); CHECK(qlc_host_alloc(64, (void**)&r)); CHECK(qlc_host_alloc(64, (void**)&d));
        for (int i = 0; i < 500; ++i) { a[0] = (uint8_t)(rnd() % 3); CHECK(qlc_env_step_host(env, a, 1, r, d)); }
        const int reps = 5000;
        double t0 = now_us(); float sum = 0;
        for (int i = 0; i < reps; ++i) { a[0] = (uint8_t)(rnd() % 3); CHECK(qlc_env_step_host(env, a, 1, r, d)); sum += r[0] + d[0]; }
        report("one_env_step_host_us", "1 env, qlc_env_step_host (page-locked buffers), per step:", (now_us() - t0) / reps);
        if (sum < 0) return 2;
        uint8_t a2[1]; float r2[1]; uint8_t d2[1];
        t0 = now_us();
        for (int i = 0; i < reps; ++i) { a2[0] = (uint8_t)(rnd() % 3); CHECK(qlc_env_step_host(env, a2, 1, r2, d2)); }
        report("one_env_step_host_pageable_us", "1 env, qlc_env_step_host (pageable buffers), per step:", (now_us() - t0) / reps);
        /* predict_action's input: the current state handle -> f32 [84][84][4] in host memory */
        float* obs; CHECK(qlc_host_alloc(4 * 84 * 84 * sizeof(float), (void**)&obs));
        uint64_t t = 0; CHECK(qlc_env_time(env, &t));
        qlc_obs_handle h; h.time = t; h.k = 4; h.env = 0;
        for (int i = 0; i < 200; ++i) CHECK(qlc_obs_gather_host(env, &h, 1, QLC_LAYOUT_F32_BXYH, obs));
        t0 = now_us();
        for (int i = 0; i < 2000; ++i) CHECK(qlc_obs_gather_host(env, &h, 1, QLC_LAYOUT_F32_BXYH, obs));
        report("one_env_obs_gather_host_f32_us", "1 env, qlc_obs_gather_host(1 handle, f32), per call:", (now_us() - t0) / 2000);
        qlc_host_free(a); qlc_host_free(r); qlc_host_free(d); qlc_host_free(obs);
        CHECK(qlc_env_destroy(env));
    }
    /* ---- the replay calls of a train step, 4,096-env shard with a 1 M-transition ring ---- */
    {
        qlc_env* env = NULL;
        cfg.n_envs = 4096; cfg.replay_capacity = 1u << 20; cfg.auto_reset = 1;
        CHECK(qlc_env_create(&cfg, &env));
        const uint32_t n = 4096, k = 64;
        uint8_t* a; CHECK(qlc_host_alloc((size_t)n * k, (void**)&a));
        for (size_t i = 0; i < (size_t)n * k; ++i) a[i] = (uint8_t)(rnd() % 3);
        for (int i = 0; i < 8; ++i) CHECK(qlc_env_step_host(env, a, k, NULL, NULL));
        uint64_t len = 0; CHECK(qlc_replay_len(env, &len));
        for (int bi = 0; bi < 2; ++bi) {
            const uint32_t batch = bi ? 512u : 32u;
            const size_t per = (size_t)4 * 84 * 84;
            float *s, *sn; CHECK(qlc_host_alloc(batch * per * sizeof(float), (void**)&s)); CHECK(qlc_host_alloc(batch * per * sizeof(float), (void**)&sn));
            uint32_t* idx = (uint32_t*)malloc(batch * 4); float* rw = (float*)malloc(batch * 4); uint8_t* ac = (uint8_t*)malloc(batch); uint8_t* dn = (uint8_t*)malloc(batch);
            const int reps = bi ? 100 : 1000;
            double total = 0, chk = 0;
            for (int it = -20; it < reps; ++it) {
                for (uint32_t i = 0; i < batch; ++i) {               /* the learner's own distinct draw (host side, not timed) */
                    for (;;) { uint32_t v = (uint32_t)(((uint64_t)rnd() * len) >> 21) % (uint32_t)len; uint32_t j = 0; while (j < i && idx[j] != v) ++j; if (j == i) { idx[i] = v; break; } }
                }
                const double t0 = now_us();
                CHECK(qlc_replay_gather_host(env, idx, batch, QLC_LAYOUT_F32_BXYH, s, sn, rw, ac, dn));
                if (it >= 0) total += now_us() - t0;
                chk += s[per * (batch - 1) + 17] + sn[5] + rw[0];
            }
            report(bi ? "gather_host_512_f32_us" : "gather_host_32_f32_us", bi ? "4,096 envs, qlc_replay_gather_host(512, f32 state + state_next), per minibatch:" : "4,096 envs, qlc_replay_gather_host( 32, f32 state + state_next), per minibatch:", total / reps);
            if (chk < 0) return 2;
            total = 0;
            for (int it = -20; it < reps; ++it) {
                const double t0 = now_us();
                CHECK(qlc_replay_gather_host(env, idx, batch, QLC_LAYOUT_U8_BHYX, s, sn, rw, ac, dn));
                if (it >= 0) total += now_us() - t0;
            }
            report(bi ? "gather_host_512_u8_us" : "gather_host_32_u8_us", bi ? "4,096 envs, qlc_replay_gather_host(512, u8 [b][slot][y][x]), per minibatch:" : "4,096 envs, qlc_replay_gather_host( 32, u8 [b][slot][y][x]), per minibatch:", total / reps);
            qlc_host_free(s); qlc_host_free(sn); free(idx); free(rw); free(ac); free(dn);
        }
        /* what the unchanged learner's model does twice per train step: batch_to_multi_dim_array of 32 state handles -> ONE f32 tensor */
        {
            const uint32_t batch = 32; const size_t per = (size_t)4 * 84 * 84;
            float* s; CHECK(qlc_host_alloc(batch * per * sizeof(float), (void**)&s));
            uint64_t t = 0; CHECK(qlc_env_time(env, &t));
            qlc_obs_handle h[32];
            double total = 0; const int reps = 1000;
            for (int it = -20; it < reps; ++it) {
                for (uint32_t i = 0; i < batch; ++i) { h[i].time = t - (rnd() % 200); h[i].k = 4 + (rnd() % 4); h[i].env = rnd() % n; }
                const double t0 = now_us();
                CHECK(qlc_obs_gather_host(env, h, batch, QLC_LAYOUT_F32_BXYH, s));
                if (it >= 0) total += now_us() - t0;
            }
            report("obs_gather_host_32_handles_f32_us", "4,096 envs, qlc_obs_gather_host(32 handles, f32), per call (one tensor):", total / reps);
            qlc_host_free(s);
        }
        qlc_host_free(a);
        CHECK(qlc_env_destroy(env));
    }
    if (json_mode) {
        printf("{");
        for (int i = 0; i < n_res; ++i) printf("%s\"%s\": %.3f", i ? ", " : "", results[i].key, results[i].us);
        printf("}\n");
    }
    return 0;
}
