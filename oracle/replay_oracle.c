/* TEST INFRASTRUCTURE ONLY — see breakout_oracle.h. Replay + sampler + vectorised driver half of the oracle.
 *
 * Follows:
 *   /root/reference/src/ql-with-tensorflow/src/learn/replay_buffer.rs   (Buffer :5-50, ReplayBuffer :53-137,
 *        BufferSample :140-146) — five parallel bounded FIFOs, logical index 0 = oldest, plus the episode-reward
 *        window (:100-124);
 *   /root/reference/src/ql-with-tensorflow/src/learn/self_driving_tf_q_learner.rs:276-296
 *        generate_distinct_random_ids — BATCH distinct uniform indices by sequential rejection;
 *   /root/reference/src/ql/src/prelude.rs:36,52-58 — state_as_rc / step_as_rc clone the whole state per step;
 *   /root/reference/src/_breakout-ml/src/breakout_environment.rs:56-77 — batch_to_multi_dim_array.
 * The reference drives ONE env (learn_episode, self_driving_tf_q_learner.rs:141-233). The vectorised driver here
 * is that loop's env/replay part run over N envs in env-index order per time step (N = 1 is exactly the
 * reference order). No reference test covers replay or gather: PARITY UNPINNED.
 */
#include "breakout_oracle.h"
#include "philox.h"
#include <stdlib.h>
#include <string.h>
#include <math.h>

/* ---- Rc<BreakoutState> ---- */
typedef struct orc_rc_state {
    int refcount;
    orc_state state;
} orc_rc_state;

static orc_rc_state* rc_new(const orc_state* s) {           /* Rc::new(state.clone()) */
    orc_rc_state* r = (orc_rc_state*)malloc(sizeof *r);
    r->refcount = 1; memcpy(&r->state, s, sizeof *s);
    return r;
}
static orc_rc_state* rc_clone(orc_rc_state* r) { r->refcount++; return r; }
static void rc_drop(orc_rc_state* r) { if (r && --r->refcount == 0) free(r); }

/* ---- ReplayBuffer<Rc<S>, A> (replay_buffer.rs:53-137) as bounded FIFOs ---- */
typedef struct {
    size_t max_len, len, head;         /* head = physical position of logical index 0 */
    uint8_t* action;
    orc_rc_state** state;
    orc_rc_state** state_next;
    float* reward;
    uint8_t* done;
    /* episode reward window (:100-124) */
    size_t ep_max, ep_len, ep_head;
    float* ep_reward;
} orc_replay;

orc_replay* orc_replay_new(size_t step_buffer_len, size_t episode_reward_buffer_len) {
    orc_replay* r = (orc_replay*)calloc(1, sizeof *r);
    r->max_len = step_buffer_len;
    r->action = (uint8_t*)malloc(step_buffer_len);
    r->state = (orc_rc_state**)calloc(step_buffer_len, sizeof(void*));
    r->state_next = (orc_rc_state**)calloc(step_buffer_len, sizeof(void*));
    r->reward = (float*)malloc(step_buffer_len * sizeof(float));
    r->done = (uint8_t*)malloc(step_buffer_len);
    r->ep_max = episode_reward_buffer_len;
    r->ep_reward = (float*)malloc(episode_reward_buffer_len * sizeof(float));
    return r;
}
void orc_replay_free(orc_replay* r) {
    if (!r) return;
    for (size_t i = 0; i < r->len; ++i) { size_t p = (r->head + i) % r->max_len; rc_drop(r->state[p]); rc_drop(r->state_next[p]); }
    free(r->action); free(r->state); free(r->state_next); free(r->reward); free(r->done); free(r->ep_reward); free(r);
}
size_t orc_replay_len(const orc_replay* r) { return r->len; }

static void replay_add(orc_replay* r, uint8_t action, orc_rc_state* s, orc_rc_state* sn, float reward, int done) { /* :85-98 */
    if (r->len >= r->max_len) {                       /* Buffer::add pops the front when full (:21-29) */
        rc_drop(r->state[r->head]); rc_drop(r->state_next[r->head]);
        r->head = (r->head + 1) % r->max_len; r->len -= 1;
    }
    size_t p = (r->head + r->len) % r->max_len;
    r->action[p] = action; r->state[p] = s; r->state_next[p] = sn; r->reward[p] = reward; r->done[p] = (uint8_t)(done != 0);
    r->len += 1;
}
void orc_replay_add_episode_reward(orc_replay* r, float v) {                  /* :100-105 */
    if (r->ep_len >= r->ep_max) { r->ep_head = (r->ep_head + 1) % r->ep_max; r->ep_len -= 1; }
    r->ep_reward[(r->ep_head + r->ep_len) % r->ep_max] = v; r->ep_len += 1;
}
size_t orc_replay_episode_rewards(const orc_replay* r, float* out) {          /* :124 */
    for (size_t i = 0; i < r->ep_len; ++i) out[i] = r->ep_reward[(r->ep_head + i) % r->ep_max];
    return r->ep_len;
}
float orc_replay_avg_episode_reward(const orc_replay* r) {                    /* :107-111 */
    float sum = 0.0f;
    for (size_t i = 0; i < r->ep_len; ++i) sum = sum + r->ep_reward[(r->ep_head + i) % r->ep_max];
    return sum / (float)r->ep_len;
}
float orc_replay_min_episode_reward(const orc_replay* r) {                    /* :113-120 */
    float mn = INFINITY;
    for (size_t i = 0; i < r->ep_len; ++i) { float v = r->ep_reward[(r->ep_head + i) % r->ep_max]; if (v < mn) mn = v; }
    return mn;
}

/* get_many (:126-137) followed by batch_to_multi_dim_array (breakout_environment.rs:56-77) for state and
 * state_next: out_state / out_next are [n][84][84][4] f32, value = u8 as f32, last axis = ring SLOT index. */
void orc_replay_get_many_f32(const orc_replay* r, const uint32_t* indices, size_t n,
                             float* out_state, float* out_next, float* reward, uint8_t* action, uint8_t* done) {
    const size_t per = (size_t)ORC_FRAME_W * ORC_FRAME_H * ORC_NUM_FRAMES;
    for (size_t b = 0; b < n; ++b) {
        size_t p = (r->head + indices[b]) % r->max_len;
        if (out_state) orc_state_to_f32_xyh(&r->state[p]->state, out_state + b * per);
        if (out_next)  orc_state_to_f32_xyh(&r->state_next[p]->state, out_next + b * per);
        reward[b] = r->reward[p]; action[b] = r->action[p]; done[b] = r->done[p];
    }
}
/* same gather, u8 frame-major layout [n][4][84][84] (slot-indexed) */
void orc_replay_get_many_u8(const orc_replay* r, const uint32_t* indices, size_t n,
                            uint8_t* out_state, uint8_t* out_next, float* reward, uint8_t* action, uint8_t* done) {
    const size_t per = (size_t)ORC_FRAME_BYTES * ORC_NUM_FRAMES;
    for (size_t b = 0; b < n; ++b) {
        size_t p = (r->head + indices[b]) % r->max_len;
        if (out_state) memcpy(out_state + b * per, r->state[p]->state.frame_buffer.buffer, per);
        if (out_next)  memcpy(out_next + b * per, r->state_next[p]->state.frame_buffer.buffer, per);
        reward[b] = r->reward[p]; action[b] = r->action[p]; done[b] = r->done[p];
    }
}
void orc_replay_action_histogram(const orc_replay* r, uint64_t counts[3]) {   /* self_driving_tf_q_learner.rs:242-245 */
    counts[0] = counts[1] = counts[2] = 0;
    for (size_t i = 0; i < r->len; ++i) counts[r->action[(r->head + i) % r->max_len] % 3] += 1;
}

/* ---- generate_distinct_random_ids (self_driving_tf_q_learner.rs:276-296) ----
 * Sequential rejection exactly as the reference; the uniform draws come from OUR Philox stream:
 * raw u32 number j of call c = word (j & 3) of philox(ctr = {j >> 2, c_lo, c_hi, 'SAMP'}, key = seed);
 * a raw draw maps to [0, len) by multiply-shift with Lemire's rejection (unbiased). */
int orc_sample_distinct(uint64_t seed, uint64_t call, uint32_t len, uint32_t batch, uint32_t* out) {
    if (len < batch) return -1;                       /* assert!(range.end - range.start >= BATCH_SIZE) */
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t thresh = (uint32_t)(((uint64_t)1 << 32) - len) % len;
    uint32_t words[4]; uint64_t j = 0;
    for (uint32_t i = 0; i < batch; ++i) {
        for (;;) {
            if ((j & 3) == 0) {
                uint32_t ctr[4] = { (uint32_t)(j >> 2), (uint32_t)call, (uint32_t)(call >> 32), ORC_STREAM_SAMPLE };
                orc_philox4x32_10(ctr, key, words);
            }
            uint32_t raw = words[j & 3]; ++j;
            uint64_t m = (uint64_t)raw * (uint64_t)len;
            if ((uint32_t)m < thresh) continue;       /* biased zone: redraw */
            uint32_t x = (uint32_t)(m >> 32);
            int contains = 0;
            for (uint32_t q = 0; q < i; ++q) if (out[q] == x) { contains = 1; break; }
            if (!contains) { out[i] = x; break; }
        }
    }
    return 0;
}

/* ---- vectorised driver: the env/replay part of learn_episode over N envs ---- */
typedef struct {
    uint32_t n_envs;
    uint64_t seed;
    uint32_t env_id_base;
    uint32_t max_episode_steps;        /* 0 = unlimited (learner default 10_000: self_driving_tf_q_learner.rs:57) */
    orc_env* envs;
    orc_rc_state** cur;                /* the learner's `state` variable per env (Rc) */
    orc_replay* replay;                /* may be NULL */
    /* shard episode statistics */
    double sum_return; uint64_t episodes; float min_return, max_return; uint64_t steps;
} orc_vec;

orc_vec* orc_vec_new(uint32_t n_envs, uint64_t seed, uint32_t env_id_base, uint32_t max_episode_steps,
                     size_t replay_capacity, size_t episode_window) {
    orc_vec* v = (orc_vec*)calloc(1, sizeof *v);
    v->n_envs = n_envs; v->seed = seed; v->env_id_base = env_id_base; v->max_episode_steps = max_episode_steps;
    v->envs = (orc_env*)calloc(n_envs, sizeof(orc_env));
    v->cur = (orc_rc_state**)calloc(n_envs, sizeof(void*));
    v->replay = replay_capacity ? orc_replay_new(replay_capacity, episode_window ? episode_window : 100) : NULL;
    v->min_return = INFINITY; v->max_return = -INFINITY;
    for (uint32_t e = 0; e < n_envs; ++e) {
        orc_env* env = &v->envs[e];
        env->seed = seed; env->env_global_id = env_id_base + e; env->episode = 0;
        orc_env_reset_with(env, orc_reset_dir_x(seed, env->env_global_id, 0));
        if (v->replay) v->cur[e] = rc_new(&env->state);            /* state_as_rc() after reset (:142-144) */
    }
    return v;
}
void orc_vec_free(orc_vec* v) {
    if (!v) return;
    for (uint32_t e = 0; e < v->n_envs; ++e) rc_drop(v->cur[e]);
    orc_replay_free(v->replay); free(v->cur); free(v->envs); free(v);
}
orc_replay* orc_vec_replay(orc_vec* v) { return v->replay; }
orc_env* orc_vec_env(orc_vec* v, uint32_t e) { return &v->envs[e]; }

/* explicit reset of one env with a caller-given direction (tests drive dir_x as an explicit input) */
void orc_vec_reset_env(orc_vec* v, uint32_t e, float dir_x) {
    orc_env* env = &v->envs[e];
    env->episode += 1;
    orc_env_reset_with(env, dir_x);
    if (v->replay) { rc_drop(v->cur[e]); v->cur[e] = rc_new(&env->state); }
}

/* one time step for envs [e0, e1): step, replay add, episode bookkeeping, auto-reset. */
void orc_vec_step_range(orc_vec* v, uint32_t e0, uint32_t e1, const uint8_t* actions, float* reward, uint8_t* done) {
    for (uint32_t e = e0; e < e1; ++e) {
        orc_env* env = &v->envs[e];
        float r; int d;
        orc_env_step(env, actions[e], &r, &d);
        reward[e] = r; done[e] = (uint8_t)d;
        if (v->replay) {
            orc_rc_state* next = rc_new(&env->state);              /* step_as_rc: Rc::new(state.clone()) */
            replay_add(v->replay, actions[e], v->cur[e], rc_clone(next), r, d);   /* :177 */
            v->cur[e] = next;                                       /* state = state_next :178 */
        }
        int truncated = (!d && v->max_episode_steps && env->episode_step >= v->max_episode_steps);
        if (d || truncated) {
            float ret = env->episode_return;
            v->sum_return += ret; v->episodes += 1;
            if (ret < v->min_return) v->min_return = ret;
            if (ret > v->max_return) v->max_return = ret;
            if (v->replay) orc_replay_add_episode_reward(v->replay, ret);           /* :220 */
            env->episode += 1;
            orc_env_reset_with(env, orc_reset_dir_x(v->seed, env->env_global_id, env->episode));
            if (v->replay) { rc_drop(v->cur[e]); v->cur[e] = rc_new(&env->state); }
        }
    }
    if (e0 == 0) v->steps += v->n_envs;
}
void orc_vec_step(orc_vec* v, const uint8_t* actions, float* reward, uint8_t* done) {
    orc_vec_step_range(v, 0, v->n_envs, actions, reward, done);
}
void orc_vec_stats(const orc_vec* v, double out[5]) {
    out[0] = v->sum_return; out[1] = (double)v->episodes; out[2] = v->min_return; out[3] = v->max_return; out[4] = (double)v->steps;
}

/* flat state read-back for comparisons */
void orc_vec_read_state(const orc_vec* v, float* ball_cx, float* ball_cy, float* ball_dx, float* ball_dy,
                        float* pad_min_x, float* pad_max_x, float* pad_speed, uint64_t* bricks, uint32_t* score,
                        uint8_t* finished, uint32_t* episode_step, uint32_t* err) {
    for (uint32_t e = 0; e < v->n_envs; ++e) {
        const orc_mechanics* m = &v->envs[e].state.mechanics;
        ball_cx[e] = m->ball_shape.center.x; ball_cy[e] = m->ball_shape.center.y;
        ball_dx[e] = m->ball_direction.x; ball_dy[e] = m->ball_direction.y;
        pad_min_x[e] = m->panel_shape.min.x; pad_max_x[e] = m->panel_shape.max.x; pad_speed[e] = m->panel_speed_per_sec;
        bricks[e] = orc_mechanics_brick_mask(m); score[e] = m->score; finished[e] = (uint8_t)m->finished;
        episode_step[e] = v->envs[e].episode_step; err[e] = m->err | v->envs[e].sticky_err;
    }
}
/* current observation stacks, u8 [n][4][84][84] slot-indexed */
void orc_vec_read_obs_u8(const orc_vec* v, uint8_t* out) {
    for (uint32_t e = 0; e < v->n_envs; ++e)
        memcpy(out + (size_t)e * ORC_NUM_FRAMES * ORC_FRAME_BYTES, v->envs[e].state.frame_buffer.buffer, (size_t)ORC_NUM_FRAMES * ORC_FRAME_BYTES);
}
void orc_vec_read_obs_f32(const orc_vec* v, float* out) {
    for (uint32_t e = 0; e < v->n_envs; ++e)
        orc_state_to_f32_xyh(&v->envs[e].state, out + (size_t)e * ORC_NUM_FRAMES * ORC_FRAME_BYTES);
}

/* synthetic random policy: action of env g at global step t (shared definition with bench.py) */
uint8_t orc_synthetic_action(uint64_t seed, uint32_t env_global_id, uint32_t t) {
    uint32_t ctr[4] = { env_global_id, t, 0u, ORC_STREAM_ACTION };
    uint32_t key[2] = { (uint32_t)seed, (uint32_t)(seed >> 32) };
    uint32_t out[4];
    orc_philox4x32_10(ctr, key, out);
    return (uint8_t)(((uint64_t)out[0] * 3u) >> 32);
}
