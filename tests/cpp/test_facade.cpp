// Drives the C++ mirror of the reference interfaces (include/ql_cuda.hpp) the way SelfDrivingQLearner::learn_episode
// does (self_driving_tf_q_learner.rs:141-233), minus the model: reset, state_as_rc, step_as_rc, replay add, the
// every-4th-step sample gate, generate_distinct_random_ids, get_many, batch_to_multi_dim_array, episode rewards.
// Exit code 0 = all checks passed. Needs a B200 (there is no CPU fallback).
#include <cstdio>
#include <cstdlib>
#include <set>
#include <vector>

#include "ql_cuda.hpp"

#define CHECK(c, ...) do { if (!(c)) { std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); return 1; } } while (0)

int main() {
    using namespace ql;
    constexpr size_t BATCH = 32;
    const uint64_t history_buffer_len = 4096;
    try {
        BreakoutEnvironment env(84, 84, history_buffer_len, /*seed=*/2024);
        ReplayBuffer replay(env, history_buffer_len, 100);
        CHECK(env.episode_reward_goal_mean() == 59.0f, "goal mean");
        // the model's inference half on the tensor cores (random weights of the reference architecture)
        std::vector<float> w1(8 * 8 * 4 * 32), b1(32, 0.01f), w2(4 * 4 * 32 * 64), b2(64, 0.01f), w3(3 * 3 * 64 * 64), b3(64, 0.01f), w4((size_t)3136 * 512), b4(512, 0.01f), w5(512 * 3), b5(3, 0.0f);
        {
            uint64_t r = 99;
            auto fill = [&](std::vector<float>& v, float scale) { for (float& x : v) { r = r * 6364136223846793005ULL + 1442695040888963407ULL; x = scale * ((float)((r >> 40) & 0xFFFF) / 32768.0f - 1.0f); } };
            fill(w1, 0.02f); fill(w2, 0.05f); fill(w3, 0.05f); fill(w4, 0.02f); fill(w5, 0.05f);
        }
        const qlc_qnet_weights weights{w1.data(), b1.data(), w2.data(), b2.data(), w3.data(), b3.data(), w4.data(), b4.data(), w5.data(), b5.data()};
        TensorCoreQModel model(env, weights);
        struct Row { uint8_t a; float r; bool d; };
        std::vector<Row> history;
        uint64_t lcg = 12345, calls = 0;
        size_t step_count = 0, trained = 0;
        for (int episode = 0; episode < 3; ++episode) {
            env.reset();
            auto state = env.state_as_rc();
            { Tensor t0 = state->to_multi_dim_array(); for (float v : t0.data) CHECK(v == 0.0f, "post-reset stack not empty"); CHECK(t0.dims.size() == 3, "dims"); }
            float episode_reward = 0.0f;
            for (int it = 0; it < 10000; ++it) {
                step_count += 1;
                lcg = lcg * 6364136223846793005ULL + 1442695040888963407ULL;
                const BreakoutAction action = BreakoutActionTrait::try_from_numeric((ModelActionType)((lcg >> 33) % 3));
                auto [state_next, reward, done] = env.step_as_rc(action);
                episode_reward += reward;
                replay.add(action, state, state_next, reward, done);
                history.push_back({BreakoutActionTrait::numeric(action), reward, done});
                state = state_next;
                CHECK(replay.len() == history.size(), "replay len %zu vs %zu", replay.len(), history.size());
                if (step_count % 4 == 0 && replay.len() > BATCH) {
                    auto indices = generate_distinct_random_ids<BATCH>(replay, calls++);
                    std::set<size_t> uniq(indices.begin(), indices.end());
                    CHECK(uniq.size() == BATCH, "indices not distinct");
                    auto sample = replay.get_many(indices);
                    std::array<const std::shared_ptr<BreakoutState>*, BATCH> sn, st;
                    for (size_t i = 0; i < BATCH; ++i) { sn[i] = &sample.state_next[i]; st[i] = &sample.state[i]; }
                    Tensor tn = BreakoutState::batch_to_multi_dim_array<BATCH>(sn);
                    Tensor ts = BreakoutState::batch_to_multi_dim_array<BATCH>(st);
                    CHECK(tn.dims.size() == 4 && tn.dims[0] == BATCH && tn.dims[1] == 84 && tn.dims[2] == 84 && tn.dims[3] == 4, "dims");
                    for (float v : tn.data) CHECK(v == 0.0f || v == 96.0f || v == 236.0f || v == 255.0f, "pixel value %f", v);
                    for (size_t i = 0; i < BATCH; ++i) {
                        CHECK(indices[i] < history.size(), "index range");
                        const Row& h = history[indices[i]];
                        CHECK(BreakoutActionTrait::numeric(sample.action[i]) == h.a && sample.reward[i] == h.r && sample.done[i] == h.d, "row %zu mismatch", indices[i]);
                    }
                    // s' of row i is s of row i+1 inside an episode
                    const size_t per = 84 * 84 * 4;
                    for (size_t i = 0; i < BATCH; ++i)
                        for (size_t k = 0; k < BATCH; ++k)
                            if (indices[k] == indices[i] + 1 && !history[indices[i]].d)
                                for (size_t p = 0; p < per; ++p) CHECK(tn.data[i * per + p] == ts.data[k * per + p], "s'(t) != s(t+1)");
                    // model inference on sample handles: max-Q of s'(t) equals max-Q of s(t+1) inside an episode, finite everywhere
                    const std::array<float, BATCH> mq_next = model.batch_predict_max_future_reward<BATCH>(sn), mq_state = model.batch_predict_max_future_reward<BATCH>(st);
                    for (size_t i = 0; i < BATCH; ++i) {
                        CHECK(mq_next[i] == mq_next[i] && mq_state[i] == mq_state[i], "max-Q is NaN");
                        for (size_t k = 0; k < BATCH; ++k)
                            if (indices[k] == indices[i] + 1 && !history[indices[i]].d) CHECK(mq_next[i] == mq_state[k], "maxQ(s'(t)) != maxQ(s(t+1))");
                    }
                    if (trained == 0) {
                        const std::array<float, 3> qv = model.q_values(*state);                    // live state == s' of the newest transition
                        const BreakoutAction greedy = model.predict_action(*state);
                        const uint8_t g = BreakoutActionTrait::numeric(greedy);
                        CHECK(qv[g] >= qv[0] && qv[g] >= qv[1] && qv[g] >= qv[2], "predict_action is not the arg max");
                        BreakoutState newest(env.handle(), BreakoutState::Kind::ReplayNext, (uint32_t)(replay.len() - 1), state->time());
                        const std::array<float, 3> qn = model.q_values(newest);
                        CHECK(qn[0] == qv[0] && qn[1] == qv[1] && qn[2] == qv[2], "live state and newest state_next differ");
                    }
                    trained += 1;
                }
                if (done) break;
            }
            replay.add_episode_reward(episode_reward);
            CHECK(replay.episode_rewards().size() == (size_t)episode + 1, "episode window");
        }
        CHECK(trained > 10, "sample gate never opened");
        CHECK(replay.avg_episode_reward() >= replay.min_episode_reward(), "avg/min");
        auto counts = replay.actions();
        CHECK(counts[0] + counts[1] + counts[2] == history.size(), "action histogram");
        bool threw = false;
        try { BreakoutActionTrait::try_from_numeric(3); } catch (const QlError& e) { threw = e.code == QLC_ERR_OUT_OF_RANGE; }
        CHECK(threw, "try_from_numeric(3) must fail");
        std::printf("facade ok: %zu steps, %zu sampled minibatches, %s\n", step_count, trained, env.state().one_line_info().c_str());
    } catch (const ql::QlError& e) {
        std::printf("QlError %d: %s\n", e.code, e.what());
        return 2;
    }
    return 0;
}
