// Drives the C++ mirror of the reference interfaces (include/ql_cuda.hpp) with the SAME call sites, argument shapes and types
// as SelfDrivingQLearner (self_driving_tf_q_learner.rs; line numbers on the right), minus the TensorFlow model:
//   replay_buffer: ReplayBuffer<Rc<E::S>, E::A>                                   :81
//   ReplayBuffer::new(param.history_buffer_len, param.episode_reward_history_buffer_len)   :100
//   environment.reset(); state = environment.state_as_rc()                        :142-144
//   (state_next, reward, done) = environment.step_as_rc(action)                   :171
//   replay_buffer.add(action, state, Rc::clone(&state_next), reward, done)        :177
//   indices: [usize; BATCH] = generate_distinct_random_ids(&mut rng, 0..len)      :183
//   replay_samples = replay_buffer.get_many(&indices)                             :185
//   batch_predict_max_future_reward(replay_samples.state_next)  -> [&Rc<S>; BATCH] :189
//   train(replay_samples.state, replay_samples.action, updated_q_values)          :201
//   replay_buffer.add_episode_reward / avg_episode_reward / min_episode_reward    :136,220-222
//   for &a in &replay_buffer.actions().buffer; .actions().buffer.len()            :242-247
// Exit code 0 = all checks passed. Needs a B200 (there is no CPU fallback).
#include <cstdio>
#include <cstdlib>
#include <map>
#include <random>
#include <set>
#include <vector>

#include "ql_cuda.hpp"

#define CHECK(c, ...) do { if (!(c)) { std::printf("FAIL %s:%d: ", __FILE__, __LINE__); std::printf(__VA_ARGS__); std::printf("\n"); return 1; } } while (0)

int main() {
    using namespace ql;
    constexpr size_t BATCH_SIZE = 32;
    using E = BreakoutEnvironment;
    using Rc = std::shared_ptr<E::S>;
    const size_t history_buffer_len = 4096, episode_reward_history_buffer_len = 100;
    try {
        E environment(84, 84, history_buffer_len, /*seed=*/2024);
        ReplayBuffer<Rc, E::A> replay_buffer(history_buffer_len, episode_reward_history_buffer_len);        // :81,:100
        CHECK(environment.episode_reward_goal_mean() == 59.0f, "goal mean");
        // the model's inference half on the tensor cores (random weights of the reference architecture)
        std::vector<float> w1(8 * 8 * 4 * 32), b1(32, 0.01f), w2(4 * 4 * 32 * 64), b2(64, 0.01f), w3(3 * 3 * 64 * 64), b3(64, 0.01f), w4((size_t)3136 * 512), b4(512, 0.01f), w5(512 * 3), b5(3, 0.0f);
        {
            uint64_t r = 99;
            auto fill = [&](std::vector<float>& v, float scale) { for (float& x : v) { r = r * 6364136223846793005ULL + 1442695040888963407ULL; x = scale * ((float)((r >> 40) & 0xFFFF) / 32768.0f - 1.0f); } };
            fill(w1, 0.02f); fill(w2, 0.05f); fill(w3, 0.05f); fill(w4, 0.02f); fill(w5, 0.05f);
        }
        const qlc_qnet_weights weights{w1.data(), b1.data(), w2.data(), b2.data(), w3.data(), b3.data(), w4.data(), b4.data(), w5.data(), b5.data()};
        TensorCoreQModel model(environment, weights);
        struct Row { uint8_t a; float r; bool d; };
        std::vector<Row> history;
        std::mt19937_64 rng(12345);
        size_t step_count = 0, trained = 0, dead_seen = 0;
        for (int episode = 0; episode < 3; ++episode) {
            environment.reset();                                                                             // :142
            CHECK(environment.lives() == 1, "a fresh episode has its one life");
            Rc state = environment.state_as_rc();                                                            // :144
            { Tensor t0 = state->to_multi_dim_array(); for (float v : t0.data) CHECK(v == 0.0f, "post-reset stack not empty"); CHECK(t0.dims.size() == 3, "dims"); }
            float episode_reward = 0.0f;
            for (int it = 0; it < 10000; ++it) {                                                             // :149
                step_count += 1;
                const E::A action = BreakoutActionTrait::try_from_numeric((ModelActionType)(rng() % BreakoutActionTrait::ACTION_SPACE));   // :156-157
                auto [state_next, reward, done] = environment.step_as_rc(action);                            // :171
                episode_reward += reward;
                replay_buffer.add(action, state, Rc(state_next), reward, done);                              // :177
                history.push_back({BreakoutActionTrait::numeric(action), reward, done});
                if (history.size() > history_buffer_len) history.erase(history.begin());
                state = state_next;                                                                          // :178
                CHECK(replay_buffer.len() == history.size(), "replay len %zu vs %zu", replay_buffer.len(), history.size());
                if (step_count % 4 == 0 && replay_buffer.len() > BATCH_SIZE) {                               // :181
                    const std::array<size_t, BATCH_SIZE> indices = generate_distinct_random_ids<BATCH_SIZE>(rng, 0, replay_buffer.len());   // :183
                    std::set<size_t> uniq(indices.begin(), indices.end());
                    CHECK(uniq.size() == BATCH_SIZE, "indices not distinct");
                    const BufferSample<BATCH_SIZE, Rc, E::A> replay_samples = replay_buffer.get_many(indices);   // :185
                    // what the model does with them (q_learning_model.rs:137,171): S::batch_to_multi_dim_array(&state_batch)
                    Tensor tn = E::S::batch_to_multi_dim_array<BATCH_SIZE>(replay_samples.state_next);
                    Tensor ts = E::S::batch_to_multi_dim_array<BATCH_SIZE>(replay_samples.state);
                    CHECK(tn.dims.size() == 4 && tn.dims[0] == BATCH_SIZE && tn.dims[1] == 84 && tn.dims[2] == 84 && tn.dims[3] == 4, "dims");
                    for (float v : tn.data) CHECK(v == 0.0f || v == 96.0f || v == 236.0f || v == 255.0f, "pixel value %f", v);
                    for (size_t i = 0; i < BATCH_SIZE; ++i) {
                        CHECK(indices[i] < history.size(), "index range");
                        const Row& h = history[indices[i]];
                        CHECK(BreakoutActionTrait::numeric(replay_samples.action[i]) == h.a && replay_samples.reward[i] == h.r && replay_samples.done[i] == h.d, "row %zu mismatch", indices[i]);
                    }
                    // s' of row i is s of row i+1 inside an episode: the SAME handle object, hence the same pixels
                    const size_t per = 84 * 84 * 4;
                    for (size_t i = 0; i < BATCH_SIZE; ++i)
                        for (size_t k = 0; k < BATCH_SIZE; ++k)
                            if (indices[k] == indices[i] + 1 && !history[indices[i]].d) {
                                CHECK(replay_samples.state_next[i]->get() == replay_samples.state[k]->get(), "s'(t) and s(t+1) should share one Rc");
                                for (size_t p = 0; p < per; ++p) CHECK(tn.data[i * per + p] == ts.data[k * per + p], "s'(t) != s(t+1)");
                            }
                    // TD target (:189-199) with the tensor-core model: finite, and consistent between s'(t) and s(t+1)
                    const std::array<float, BATCH_SIZE> max_future_rewards = model.batch_predict_max_future_reward<BATCH_SIZE>(replay_samples.state_next);   // :189
                    const std::array<float, BATCH_SIZE> mq_state = model.batch_predict_max_future_reward<BATCH_SIZE>(replay_samples.state);
                    std::array<float, BATCH_SIZE> updated_q_values;
                    for (size_t i = 0; i < BATCH_SIZE; ++i) {
                        updated_q_values[i] = replay_samples.done[i] ? replay_samples.reward[i] : replay_samples.reward[i] + 0.99f * max_future_rewards[i];
                        CHECK(updated_q_values[i] == updated_q_values[i] && mq_state[i] == mq_state[i], "max-Q is NaN");
                        for (size_t k = 0; k < BATCH_SIZE; ++k)
                            if (indices[k] == indices[i] + 1 && !history[indices[i]].d) CHECK(max_future_rewards[i] == mq_state[k], "maxQ(s'(t)) != maxQ(s(t+1))");
                    }
                    if (trained == 0) {
                        const std::array<float, 3> qv = model.q_values(*state);                              // the live state
                        const uint8_t g = BreakoutActionTrait::numeric(model.predict_action(*state));        // :160
                        CHECK(qv[g] >= qv[0] && qv[g] >= qv[1] && qv[g] >= qv[2], "predict_action is not the arg max");
                    }
                    trained += 1;
                }
                if (done) { CHECK(environment.lives() == 0, "a finished episode has no life left"); dead_seen += 1; break; }   // :214
            }
            replay_buffer.add_episode_reward(episode_reward);                                                // :220
            CHECK(replay_buffer.episode_rewards().size() == (size_t)episode + 1, "episode window");
        }
        CHECK(trained > 10 && dead_seen == 3, "sample gate never opened / episodes never ended");
        CHECK(replay_buffer.avg_episode_reward() >= replay_buffer.min_episode_reward(), "avg/min");          // :136,:222
        std::map<E::A, size_t> action_counts;                                                                // :242-247
        for (const E::A& a : replay_buffer.actions().buffer) action_counts[a] += 1;
        const size_t total_actions = replay_buffer.actions().buffer.size();
        CHECK(total_actions == history.size() && action_counts.size() == 3, "action histogram");
        bool threw = false;
        try { BreakoutActionTrait::try_from_numeric(3); } catch (const QlError& e) { threw = e.code == QLC_ERR_OUT_OF_RANGE; }
        CHECK(threw, "try_from_numeric(3) must fail");
        // a handle whose frames have left the ring is refused
        threw = false;
        {
            E small(84, 84, /*history_buffer_len=*/16, 7);
            small.reset();
            Rc old = small.state_as_rc();
            for (int i = 0; i < 40; ++i) { auto [s, r, d] = small.step_as_rc(BreakoutAction::None); old = i == 0 ? s : old; if (d) small.reset(); }
            try { E::S::batch_to_multi_dim_array<1>({&old}); } catch (const QlError& e) { threw = e.code == QLC_ERR_OUT_OF_RANGE; }
        }
        CHECK(threw, "a stale handle must be refused");
        std::printf("facade ok: %zu steps, %zu sampled minibatches, %s\n", step_count, trained, environment.state().one_line_info().c_str());
    } catch (const ql::QlError& e) {
        std::printf("QlError %d: %s\n", e.code, e.what());
        return 2;
    }
    return 0;
}
