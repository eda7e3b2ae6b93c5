// ql_cuda.hpp — C++ host-side mirror of the reference's Rust interfaces for the hot path, over the C ABI (ql_cuda.h).
//
// The reference is compiled Rust and this image has no cargo/rustc, so the host layer above the C ABI is written in
// C++ with the SAME names, argument meaning and error behaviour as the reference, so that the learner loop
// (self_driving_tf_q_learner.rs:141-233) reads the same against it:
//
//   ql::prelude::Action / BreakoutAction ....... prelude.rs:12-18, breakout_environment.rs:94-120  -> ql::BreakoutAction
//   ql::prelude::Environment ................... prelude.rs:21-63                                  -> ql::BreakoutEnvironment
//   BreakoutState + ToMultiDimArray ............ breakout_environment.rs:24-78, model.rs:12-26      -> ql::BreakoutState
//   ReplayBuffer / BufferSample ................ replay_buffer.rs:53-146                            -> ql::ReplayBuffer / ql::BufferSample
//   generate_distinct_random_ids ............... self_driving_tf_q_learner.rs:276-296               -> ql::generate_distinct_random_ids
//
// A Rust `ql-cuda` crate with the same shape is in bindings/rust/ql-cuda (source only).
// Header-only; link with libqlcuda.so. No CPU fallback: every call fails with QlError when CUDA is unavailable.
#pragma once
#include <array>
#include <cstdint>
#include <memory>
#include <stdexcept>
#include <string>
#include <tuple>
#include <vector>

#include "ql_cuda.h"

namespace ql {

using ModelActionType = uint8_t;                                   // prelude.rs:10

struct QlError : std::runtime_error {                              // prelude.rs:70-86
    int32_t code;
    QlError(const std::string& msg, int32_t c = QLC_ERR_INVALID_ARG) : std::runtime_error(msg), code(c) {}
};

inline void check(int32_t rc) {
    if (rc != QLC_OK) throw QlError(qlc_last_error_string(), rc);
}

// ---- Action (breakout_environment.rs:94-120) ----
enum class BreakoutAction : ModelActionType { None = 0, Left = 1, Right = 2 };
struct BreakoutActionTrait {
    static constexpr ModelActionType ACTION_SPACE = 3;
    static ModelActionType numeric(BreakoutAction a) { return static_cast<ModelActionType>(a); }
    static BreakoutAction try_from_numeric(ModelActionType v) {
        if (v >= ACTION_SPACE) throw QlError("value out of range", QLC_ERR_OUT_OF_RANGE);
        return static_cast<BreakoutAction>(v);
    }
};

// ---- Tensor<f32> stand-in: dims + row-major data (what ToMultiDimArray<Tensor<f32>> produces) ----
struct Tensor {
    std::vector<uint64_t> dims;
    std::vector<float> data;
};

class EnvHandle {   // owns the qlc_env; shared by the environment, its states and its replay buffer
public:
    explicit EnvHandle(const qlc_config& cfg) { check(qlc_env_create(&cfg, &h_)); n_envs_ = cfg.n_envs; }
    ~EnvHandle() { qlc_env_destroy(h_); }
    EnvHandle(const EnvHandle&) = delete;
    EnvHandle& operator=(const EnvHandle&) = delete;
    qlc_env* get() const { return h_; }
    uint32_t n_envs() const { return n_envs_; }
    uint64_t time() const { uint64_t t = 0; check(qlc_env_time(h_, &t)); return t; }
private:
    qlc_env* h_ = nullptr;
    uint32_t n_envs_ = 0;
};

// ---- BreakoutState (breakout_environment.rs:24-28): a cheap handle, Clone = copy of indices, pixels stay in HBM ----
class BreakoutState {
public:
    enum class Kind { Live, ReplayState, ReplayNext };
    BreakoutState(std::shared_ptr<EnvHandle> env, uint64_t time) : env_(std::move(env)), kind_(Kind::Live), time_(time) {}
    BreakoutState(std::shared_ptr<EnvHandle> env, Kind kind, uint32_t replay_index, uint64_t time)
        : env_(std::move(env)), kind_(kind), index_(replay_index), time_(time) {}

    std::array<uint64_t, 3> dims() const { return {QLC_FRAME_W, QLC_FRAME_H, QLC_NUM_FRAMES}; }   // model_dims :148

    // ToMultiDimArray::to_multi_dim_array (:42-54): [x][y][slot] f32, value = u8 as f32 (env 0 of the shard)
    Tensor to_multi_dim_array() const {
        const std::shared_ptr<BreakoutState> self = std::make_shared<BreakoutState>(*this);
        std::array<const std::shared_ptr<BreakoutState>*, 1> b{&self};
        Tensor t = batch_to_multi_dim_array<1>(b);
        t.dims.erase(t.dims.begin());
        return t;
    }

    // ToMultiDimArray::batch_to_multi_dim_array (:56-77): [b][x][y][slot] f32 — ONE gather kernel for the batch
    template <size_t N>
    static Tensor batch_to_multi_dim_array(const std::array<const std::shared_ptr<BreakoutState>*, N>& batch) {
        static_assert(N > 0, "empty batch");
        const BreakoutState& first = **batch[0];
        const size_t per = (size_t)QLC_FRAME_W * QLC_FRAME_H * QLC_NUM_FRAMES;
        Tensor t;
        t.dims = {N, QLC_FRAME_W, QLC_FRAME_H, QLC_NUM_FRAMES};
        t.data.resize(N * per);
        if (first.kind_ == Kind::Live) {
            // live handles: the current observation of env 0..n-1 (single-env drop-in: N == 1)
            if (first.time_ != first.env_->time()) throw QlError("stale BreakoutState handle (the env has stepped since)");
            std::vector<float> all((size_t)first.env_->n_envs() * per);
            check(qlc_env_obs_host(first.env_->get(), QLC_LAYOUT_F32_BXYH, all.data()));
            for (size_t b = 0; b < N; ++b) std::copy(all.begin(), all.begin() + per, t.data.begin() + b * per);
            return t;
        }
        std::array<uint32_t, N> idx;
        for (size_t b = 0; b < N; ++b) {
            const BreakoutState& s = **batch[b];
            if (s.kind_ != first.kind_) throw QlError("mixed state kinds in one batch");
            if (s.time_ != s.env_->time()) throw QlError("stale replay sample (the env has stepped since get_many)");
            idx[b] = s.index_;
        }
        const bool next = first.kind_ == Kind::ReplayNext;
        check(qlc_replay_gather_host(first.env_->get(), idx.data(), (uint32_t)N, QLC_LAYOUT_F32_BXYH, next ? nullptr : t.data.data(),
                                     next ? t.data.data() : nullptr, nullptr, nullptr, nullptr));
        return t;
    }

    std::string one_line_info() const {                            // DebugVisualizer :81-89
        std::vector<float> cx(env_->n_envs()), cy(env_->n_envs()), pmin(env_->n_envs()), pmax(env_->n_envs());
        std::vector<uint64_t> bricks(env_->n_envs());
        qlc_state_host sh{};
        sh.ball_cx = cx.data(); sh.ball_cy = cy.data(); sh.pad_min_x = pmin.data(); sh.pad_max_x = pmax.data(); sh.bricks = bricks.data();
        check(qlc_env_read_state(env_->get(), &sh));
        return "Breakout [" + std::to_string(__builtin_popcountll(bricks[0])) + " bricks, ball_pos: [" + std::to_string(cx[0]) + " " +
               std::to_string(cy[0]) + "], panel_pos: [" + std::to_string((pmin[0] + pmax[0]) / 2.0f) + " 570]]";
    }
    uint64_t time() const { return time_; }
    Kind kind() const { return kind_; }
    uint32_t replay_index() const { return index_; }

private:
    std::shared_ptr<EnvHandle> env_;
    Kind kind_;
    uint32_t index_ = 0;
    uint64_t time_;
};

// ---- Environment (prelude.rs:21-63) for ONE Breakout env: the drop-in the unchanged learner drives ----
class BreakoutEnvironment {
public:
    using S = BreakoutState;
    using A = BreakoutAction;

    // BreakoutEnvironment::new(frame_size_x, frame_size_y) (:139-153); replay_capacity sizes the HBM frame ring that
    // ReplayBuffer views (ReplayBuffer::new(step_buffer_len, ..)).
    BreakoutEnvironment(uint32_t frame_size_x, uint32_t frame_size_y, uint64_t replay_capacity = 0, uint64_t seed = 0, int32_t device = 0) {
        qlc_config cfg{};
        cfg.struct_size = sizeof cfg; cfg.device = device; cfg.n_envs = 1; cfg.env_id_base = 0;
        cfg.frame_w = frame_size_x; cfg.frame_h = frame_size_y; cfg.seed = seed; cfg.replay_capacity = replay_capacity;
        cfg.max_episode_steps = 0; cfg.episode_window = 100; cfg.auto_reset = 0;     // the learner resets (learn_episode :142)
        env_ = std::make_shared<EnvHandle>(cfg);
        state_ = std::make_unique<BreakoutState>(env_, env_->time());
    }

    void reset() {                                                  // :177-180
        check(qlc_env_reset(env_->get(), nullptr, nullptr));
        state_ = std::make_unique<BreakoutState>(env_, env_->time());
    }
    const S& state() const { return *state_; }                      // :182
    std::shared_ptr<S> state_as_rc() const { return std::make_shared<S>(*state_); }   // prelude.rs:36

    std::tuple<const S&, float, bool> step(A action) {              // :184-201
        const uint8_t a = BreakoutActionTrait::numeric(action);
        float reward = 0.0f; uint8_t done = 0;
        check(qlc_env_step_host(env_->get(), &a, 1, &reward, &done));
        state_ = std::make_unique<BreakoutState>(env_, env_->time());
        return {*state_, reward, done != 0};
    }
    std::tuple<std::shared_ptr<S>, float, bool> step_as_rc(A action) {   // prelude.rs:52-58
        auto [s, r, d] = step(action);
        return {std::make_shared<S>(s), r, d};
    }
    float episode_reward_goal_mean() const { return qlc_env_goal_mean(); }   // :203-206

    const std::shared_ptr<EnvHandle>& handle() const { return env_; }

private:
    std::shared_ptr<EnvHandle> env_;
    std::unique_ptr<BreakoutState> state_;
};

// ---- BufferSample / ReplayBuffer (replay_buffer.rs:53-146) ----
template <size_t N>
struct BufferSample {
    std::array<std::shared_ptr<BreakoutState>, N> state;
    std::array<std::shared_ptr<BreakoutState>, N> state_next;
    std::array<float, N> reward;
    std::array<BreakoutAction, N> action;
    std::array<bool, N> done;
};

class ReplayBuffer {
public:
    // ReplayBuffer::new(step_buffer_len, episode_reward_buffer_len): the step buffer is the frame/record ring the
    // environment was created with; it must be at least as long as asked for here.
    ReplayBuffer(const BreakoutEnvironment& env, uint64_t step_buffer_len, uint64_t /*episode_reward_buffer_len*/ = 100) : env_(env.handle()) {
        uint64_t cap = 0; check(qlc_replay_capacity(env_->get(), &cap));
        if (cap < step_buffer_len) throw QlError("the environment's replay ring is shorter than step_buffer_len");
    }
    size_t len() const { uint64_t n = 0; check(qlc_replay_len(env_->get(), &n)); return (size_t)n; }   // :83

    // add (:85-98). The step kernel already appended this transition (frame + 4-byte record) on the device; the call
    // is kept so the learner's call site is unchanged, and it checks that it is handed the transition just made.
    void add(BreakoutAction, const std::shared_ptr<BreakoutState>& state, const std::shared_ptr<BreakoutState>& state_next, float, bool) {
        if (state_next->time() != env_->time() || state->time() + 1 != state_next->time())
            throw QlError("ReplayBuffer::add: not the transition the environment just produced");
    }
    void add_episode_reward(float r) { check(qlc_stats_push(env_->get(), r)); }                        // :100-105
    float avg_episode_reward() const { float v = 0; check(qlc_stats_mean(env_->get(), &v)); return v; } // :107-111
    float min_episode_reward() const { float v = 0; check(qlc_stats_min(env_->get(), &v)); return v; }  // :113-120
    std::array<uint64_t, 3> actions() const { std::array<uint64_t, 3> c{}; check(qlc_replay_action_counts(env_->get(), c.data())); return c; }   // :122 (as counts)
    std::vector<float> episode_rewards() const {                                                       // :124
        uint32_t n = 0; check(qlc_stats_window(env_->get(), nullptr, 0, &n));
        std::vector<float> out(n); if (n) check(qlc_stats_window(env_->get(), out.data(), n, &n));
        return out;
    }
    template <size_t N>
    BufferSample<N> get_many(const std::array<size_t, N>& indices) const {                             // :126-137
        std::array<uint32_t, N> idx; std::array<uint8_t, N> action, done; BufferSample<N> s;
        for (size_t i = 0; i < N; ++i) idx[i] = (uint32_t)indices[i];
        check(qlc_replay_gather_host(env_->get(), idx.data(), (uint32_t)N, QLC_LAYOUT_U8_BHYX, nullptr, nullptr, s.reward.data(), action.data(), done.data()));
        const uint64_t now = env_->time();
        for (size_t i = 0; i < N; ++i) {
            s.state[i] = std::make_shared<BreakoutState>(env_, BreakoutState::Kind::ReplayState, idx[i], now);
            s.state_next[i] = std::make_shared<BreakoutState>(env_, BreakoutState::Kind::ReplayNext, idx[i], now);
            s.action[i] = BreakoutActionTrait::try_from_numeric(action[i]);
            s.done[i] = done[i] != 0;
        }
        return s;
    }
    const std::shared_ptr<EnvHandle>& handle() const { return env_; }

private:
    std::shared_ptr<EnvHandle> env_;
};

// generate_distinct_random_ids (self_driving_tf_q_learner.rs:276-296): BATCH distinct uniform ids in 0..len, drawn on the
// device from the Philox stream (seed, call_index).
template <size_t BATCH_SIZE>
std::array<size_t, BATCH_SIZE> generate_distinct_random_ids(const ReplayBuffer& rb, uint64_t call_index) {
    std::array<uint32_t, BATCH_SIZE> idx;
    check(qlc_replay_sample_host(rb.handle()->get(), (uint32_t)BATCH_SIZE, call_index, idx.data()));
    std::array<size_t, BATCH_SIZE> out;
    for (size_t i = 0; i < BATCH_SIZE; ++i) out[i] = idx[i];
    return out;
}

// ---- the inference half of DeepQLearningModel (ml_model/model.rs:29-77) on the library's tensor-core Q-network ----
// predict_action (:40-46) and batch_predict_max_future_reward (:48-57) for BreakoutState handles: the network reads the u8
// frames in the frame ring directly, no tensor is materialised. train() (:59-71) stays with the caller, who hands updated
// weights over with set_weights (ten f32 arrays in the Keras layouts, see qlc_qnet_weights).
class TensorCoreQModel {
public:
    TensorCoreQModel(const BreakoutEnvironment& env, const qlc_qnet_weights& weights) : env_(env.handle()) { check(qlc_qnet_create(env_->get(), &weights, &q_)); }
    ~TensorCoreQModel() { qlc_qnet_destroy(q_); }
    TensorCoreQModel(const TensorCoreQModel&) = delete;
    TensorCoreQModel& operator=(const TensorCoreQModel&) = delete;
    void set_weights(const qlc_qnet_weights& weights) { check(qlc_qnet_set_weights(q_, &weights)); }

    BreakoutAction predict_action(const BreakoutState& state) const {
        std::array<float, 3> qv = q_values(state);
        (void)qv;
        return BreakoutActionTrait::try_from_numeric(last_action_);
    }
    std::array<float, 3> q_values(const BreakoutState& state) const {
        std::array<float, 3> qv{};
        if (state.kind() == BreakoutState::Kind::Live) {
            if (state.time() != env_->time()) throw QlError("stale BreakoutState handle (the env has stepped since)");
            std::vector<float> q((size_t)env_->n_envs() * 3); std::vector<uint8_t> a(env_->n_envs());
            check(qlc_qnet_forward_host(q_, nullptr, env_->n_envs(), 0, q.data(), a.data(), nullptr));
            std::copy(q.begin(), q.begin() + 3, qv.begin()); last_action_ = a[0];
        } else {
            const uint32_t idx = state.replay_index();
            check(qlc_qnet_forward_host(q_, &idx, 1, state.kind() == BreakoutState::Kind::ReplayNext ? 1 : 0, qv.data(), &last_action_, nullptr));
        }
        return qv;
    }
    template <size_t N>
    std::array<float, N> batch_predict_max_future_reward(const std::array<const std::shared_ptr<BreakoutState>*, N>& batch) const {
        std::array<uint32_t, N> idx; std::array<float, N> out{};
        const BreakoutState::Kind kind = (**batch[0]).kind();
        if (kind == BreakoutState::Kind::Live) throw QlError("batch_predict_max_future_reward takes replay sample handles");
        for (size_t b = 0; b < N; ++b) {
            const BreakoutState& s = **batch[b];
            if (s.kind() != kind) throw QlError("mixed state kinds in one batch");
            if (s.time() != env_->time()) throw QlError("stale replay sample (the env has stepped since get_many)");
            idx[b] = s.replay_index();
        }
        check(qlc_qnet_forward_host(q_, idx.data(), (uint32_t)N, kind == BreakoutState::Kind::ReplayNext ? 1 : 0, nullptr, nullptr, out.data()));
        return out;
    }

private:
    std::shared_ptr<EnvHandle> env_;
    qlc_qnet* q_ = nullptr;
    mutable uint8_t last_action_ = 0;
};

}  // namespace ql
