"""Round-2 additions: the one-launch sample + gather, state handles, the 1:1 drop-in types (q-learning_b200/dropin.py) driven
by a literal restatement of SelfDrivingQLearner::learn_episode, `lives`, the statistics reduction behind the C ABI, the u8
[b][x][y][slot] layout and host-side widening, and the hygiene items (ablation switch compiled out, stale-library detection,
exact roofline bytes, long launches split at the ring length, Q-network / env lifetime)."""
import hashlib
import importlib
import os
import subprocess
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


# ---------------------------------------------------------------------------------------------------------------------
# CPU: host logic of the drop-in types
# ---------------------------------------------------------------------------------------------------------------------
def test_buffer_and_replay_buffer_are_the_reference_fifo(qlb):
    """Buffer<T> / ReplayBuffer<S, A> (replay_buffer.rs:5-146): bounded FIFO, index 0 = oldest, get_many by logical index,
    f32 running mean front to back, min."""
    D = importlib.import_module("q-learning_b200.dropin")
    b = D.Buffer(5)
    model = []
    for i in range(23):
        b.add(i)
        model.append(i)
        model = model[-5:]
        assert b.len() == len(model) and b.buffer == model
        assert b.get_many(list(range(len(model)))) == model
    with pytest.raises(qlb.QlError):
        b.get_many([5])
    rb = D.ReplayBuffer(7, 3)
    for t in range(20):
        rb.add(t % 3, ("s", t), ("s", t + 1), float(t), t % 5 == 0)
    assert rb.len() == 7
    s = rb.get_many([0, 6, 3])
    assert s.state == [("s", 13), ("s", 19), ("s", 16)] and s.state_next == [("s", 14), ("s", 20), ("s", 17)]
    assert s.reward == [13.0, 19.0, 16.0] and s.action == [1, 1, 1] and s.done == [False, False, False]
    assert rb.actions().buffer == [t % 3 for t in range(13, 20)] and rb.actions().max_buffer_len == 7
    for r in (0.1, 0.2, 0.7, 1.9):
        rb.add_episode_reward(r)
    w = [np.float32(0.2), np.float32(0.7), np.float32(1.9)]
    assert rb.episode_rewards() == w
    assert rb.avg_episode_reward() == np.float32(np.float32(np.float32(w[0] + w[1]) + w[2]) / np.float32(3))
    assert rb.min_episode_reward() == w[0]


def test_host_generate_distinct_random_ids_property(qlb):
    """the reference's own test of its private sampler (:346-361): 50 distinct ids in 0..100, 100 times"""
    D = importlib.import_module("q-learning_b200.dropin")
    rng = np.random.default_rng(3)
    for _ in range(100):
        ids = D.generate_distinct_random_ids(rng, (0, 100), 50)
        assert len(set(ids)) == 50 and all(0 <= i < 100 for i in ids)


def test_stale_library_is_detected_not_used(qlb, tmp_path):
    """build.py: a library built from other sources is stale (hash marker inside the file); a failed rebuild raises instead of
    silently returning yesterday's binary."""
    B = qlb._build
    info = B.library_info()
    assert info.get("src_hash") == B.source_hash() and info.get("profiling") == "0"
    assert qlb.build_info()["src_hash"] == B.source_hash()
    fake = tmp_path / "libqlcuda.so"
    data = open(B.SO_PATH, "rb").read()
    fake.write_bytes(data.replace(info["src_hash"].encode(), b"0" * 64))
    assert B.library_info(str(fake))["src_hash"] == "0" * 64 != B.source_hash()
    code = ("import importlib, sys; sys.path.insert(0, %r); b = importlib.import_module('q-learning_b200.build'); "
            "b.SO_PATH = %r; b._nvcc = lambda: '/bin/false'\n"
            "assert b.needs_build()\n"
            "try:\n    b.build()\nexcept RuntimeError as e:\n    print('RAISED', 'NOT used' in str(e))\nelse:\n    print('RETURNED')\n") % (ROOT, str(fake))
    res = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True)
    assert "RAISED True" in res.stdout, res.stdout + res.stderr


def test_release_library_has_no_ablation_switch(qlb):
    """QLC_DEBUG_SKIP is compiled out of release builds: the library says profiling=0 and never reads the variable."""
    assert qlb.build_info()["profiling"] == "0"
    data = open(qlb.library_path(), "rb").read()
    assert b"QLC_DEBUG_SKIP" not in data


def test_bench_bytes_formula():
    """roofline bytes of one step launch, exact: 7,066 B per env-step (frame 7,056 + action 1 + reward 4 + done 1 + record 4)
    plus 106 B per env and launch (the 53-byte state read once and written once)."""
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    assert bench.launch_bytes(4096, 64) == 7066 * 4096 * 64 + 106 * 4096 == 1852743680
    assert bench.launch_bytes(1, 1) == 7066 + 106
    assert bench.STATE_BYTES_PER_ENV == 4 * 4 + 3 * 4 + 8 + 4 * 4 + 1


def test_comm_entry_points_fail_loudly_without_a_gpu(qlb):
    if qlb.device_count() > 0:
        pytest.skip("a GPU is present")
    lib = qlb.load_library()
    assert lib.qlc_comm_init(None, 0, 1, None) == qlb.ERR_INVALID_ARG
    assert lib.qlc_stats_allreduce(None, None) == qlb.ERR_INVALID_ARG
    assert lib.qlc_replay_sample_gather(None, 32, 1, 0, 0, None, None, None, None, None, None, None) == qlb.ERR_INVALID_ARG
    assert lib.qlc_obs_gather_host(None, None, 0, 0, None) == qlb.ERR_INVALID_ARG


# ---------------------------------------------------------------------------------------------------------------------
# GPU
# ---------------------------------------------------------------------------------------------------------------------
def _run(env, ora, O, seed, n, steps, t0=0):
    acts = O.synthetic_actions(seed, 0, n, t0, steps)
    env.step_many(acts)
    for s in range(steps):
        ora.step(acts[s])


@pytest.mark.gpu
@pytest.mark.parametrize("n_envs,steps,batch,n_batches", [(64, 40, 32, 1), (64, 40, 32, 9), (32, 40, 512, 1), (32, 40, 512, 3), (32, 33, 1024, 2),
                                                          (4, 9, 32, 2), (1, 1, 1, 1), (16, 300, 7, 5), (16, 9, 129, 3), (16, 9, 143, 2), (32, 60, 300, 2)])
def test_sample_gather_in_one_launch(qlb, O, n_envs, steps, batch, n_batches):
    """qlc_replay_sample_gather: the gather kernels draw the indices themselves. Indices equal the oracle's sequential rejection
    loop (and qlc_replay_sample), stacks and scalars equal the oracle's get_many of those indices, all three layouts."""
    torch = pytest.importorskip("torch")
    seed = 31
    env = qlb.BreakoutEnvironment(n_envs=n_envs, seed=seed, replay_capacity=n_envs * 64)
    rb = qlb.ReplayBuffer(env)
    ora = O.VecEnv(n_envs, seed=seed, replay_capacity=n_envs * 64)
    _run(env, ora, O, seed, n_envs, steps)
    ln = rb.len()
    assert ln == ora.replay_len() and ln >= batch
    n = batch * n_batches
    per = 4 * 84 * 84
    for call in (0, 5, (1 << 40) + 3):
        want_idx = np.concatenate([O.sample_distinct(seed, call + i, ln, batch) for i in range(n_batches)])
        o8, o32 = ora.get_many(want_idx, "u8"), ora.get_many(want_idx, "f32")
        for layout, dt in ((qlb.LAYOUT_U8_BHYX, torch.uint8), (qlb.LAYOUT_F32_BXYH, torch.float32), (qlb.LAYOUT_U8_BXYH, torch.uint8)):
            idx = torch.full((n,), -1, dtype=torch.int32, device="cuda")
            st = torch.full((n, per), 7, dtype=dt, device="cuda"); nx = torch.full((n, per), 7, dtype=dt, device="cuda")
            r = torch.empty((n,), dtype=torch.float32, device="cuda"); a = torch.empty((n,), dtype=torch.uint8, device="cuda"); d = torch.empty((n,), dtype=torch.uint8, device="cuda")
            rb.sample_gather_device(batch, n_batches, call, layout, idx.data_ptr(), st.data_ptr(), nx.data_ptr(), r.data_ptr(), a.data_ptr(), d.data_ptr())
            torch.cuda.synchronize()
            assert np.array_equal(idx.cpu().numpy().view(np.uint32), want_idx), "indices differ (layout %d)" % layout
            o = o8 if layout == qlb.LAYOUT_U8_BHYX else o32
            want_s, want_n = o["state"].reshape(n, per), o["state_next"].reshape(n, per)
            assert np.array_equal(st.cpu().numpy(), want_s.astype(st.cpu().numpy().dtype)) and np.array_equal(nx.cpu().numpy(), want_n.astype(nx.cpu().numpy().dtype))
            assert np.array_equal(r.cpu().numpy(), o["reward"]) and np.array_equal(a.cpu().numpy(), o["action"]) and np.array_equal(d.cpu().numpy(), o["done"])
        # without an index output, state only / next only
        st = torch.zeros((n, per), dtype=torch.uint8, device="cuda")
        rb.sample_gather_device(batch, n_batches, call, qlb.LAYOUT_U8_BHYX, None, st.data_ptr(), None)
        torch.cuda.synchronize()
        assert np.array_equal(st.cpu().numpy(), o8["state"].reshape(n, per))
    # host form
    idx, smp = rb.sample(batch, qlb.LAYOUT_F32_BXYH, call_index=11)
    o = ora.get_many(idx, "f32")
    assert np.array_equal(idx, O.sample_distinct(seed, 11, ln, batch))
    assert np.array_equal(smp.state, o["state"]) and np.array_equal(smp.state_next, o["state_next"]) and np.array_equal(smp.reward, o["reward"])
    env.close(); ora.close()


@pytest.mark.gpu
def test_scalar_only_gather_and_out_of_range_indices(qlb, O):
    """get_many without tensorisation in BOTH layouts writes reward / action / done (the f32 kernel left them
    uninitialised); on the device path an index >= len gives all-zero stacks and zero scalars, never a stale frame."""
    torch = pytest.importorskip("torch")
    n, seed = 16, 23
    env = qlb.BreakoutEnvironment(n_envs=n, seed=seed, replay_capacity=n * 32)
    rb = qlb.ReplayBuffer(env)
    ora = O.VecEnv(n, seed=seed, replay_capacity=n * 32)
    _run(env, ora, O, seed, n, 50)
    idx = rb.generate_distinct_random_ids(40, 0)
    o = ora.get_many(idx, "u8")
    for layout in (qlb.LAYOUT_F32_BXYH, qlb.LAYOUT_U8_BHYX, qlb.LAYOUT_U8_BXYH):
        g = rb.get_many(idx, layout, want_state=False, want_next=False)
        assert g.state is None and g.state_next is None
        assert np.array_equal(g.reward, o["reward"]) and np.array_equal(g.action, o["action"]) and np.array_equal(g.done, o["done"])
    per = 4 * 84 * 84
    bad = np.array([rb.len(), rb.len() + 5, 0xFFFFFFF0, 3], dtype=np.uint32)
    want = ora.get_many(np.array([3], dtype=np.uint32), "u8")
    for layout, dt in ((qlb.LAYOUT_U8_BHYX, torch.uint8), (qlb.LAYOUT_F32_BXYH, torch.float32)):
        t_idx = torch.from_numpy(bad.view(np.int32)).cuda()
        st = torch.full((4, per), 9, dtype=dt, device="cuda"); nx = torch.full((4, per), 9, dtype=dt, device="cuda")
        r = torch.full((4,), 5.0, dtype=torch.float32, device="cuda"); a = torch.full((4,), 5, dtype=torch.uint8, device="cuda"); d = torch.full((4,), 5, dtype=torch.uint8, device="cuda")
        rb.gather_device(t_idx.data_ptr(), 4, layout, st.data_ptr(), nx.data_ptr(), r.data_ptr(), a.data_ptr(), d.data_ptr())
        torch.cuda.synchronize()
        assert float(st[:3].abs().sum()) == 0.0 and float(nx[:3].abs().sum()) == 0.0
        assert r[:3].tolist() == [0.0] * 3 and a[:3].tolist() == [0] * 3 and d[:3].tolist() == [0] * 3
        assert float(nx[3].sum()) > 0 and r[3].item() == want["reward"][0]
    env.close(); ora.close()


@pytest.mark.gpu
def test_long_launch_is_split_at_the_ring_length(qlb, O):
    """n_steps > time_slots: the frame ring wraps inside the call; it is cut into launches of at most one ring length, and the
    result equals the oracle (state, stacks, replay rows)."""
    n, seed, cap_steps = 300, 5, 6                     # ring of 6 + 4 = 10 time slots, 47 steps per call
    env = qlb.BreakoutEnvironment(n_envs=n, seed=seed, replay_capacity=n * cap_steps)
    rb = qlb.ReplayBuffer(env)
    ora = O.VecEnv(n, seed=seed, replay_capacity=n * cap_steps)
    t = 0
    for steps in (47, 10, 11, 64):
        _run(env, ora, O, seed, n, steps, t)
        t += steps
        gs, os_ = env.read_state(), ora.state()
        for k in ("ball_cx", "ball_cy", "pad_min_x", "bricks", "score", "episode_step"):
            assert np.array_equal(gs[k], os_[k]), k
        assert np.array_equal(env.obs(qlb.LAYOUT_U8_BHYX), ora.obs_u8())
        idx = np.arange(0, rb.len(), 7, dtype=np.uint32)
        g, o = rb.get_many(idx, qlb.LAYOUT_U8_BHYX), ora.get_many(idx, "u8")
        assert np.array_equal(g.state, o["state"]) and np.array_equal(g.state_next, o["state_next"]) and np.array_equal(g.action, o["action"])
    assert env.error_flags() & qlb.ENVERR_HANDOVER == 0
    env.close(); ora.close()


@pytest.mark.gpu
def test_debug_skip_variable_is_ignored_by_the_release_build(qlb, O, monkeypatch):
    monkeypatch.setenv("QLC_DEBUG_SKIP", "2")          # would drop the frame stores in an ablation build
    n, seed = 40, 77
    env = qlb.BreakoutEnvironment(n_envs=n, seed=seed)
    ora = O.VecEnv(n, seed=seed)
    _run(env, ora, O, seed, n, 30)
    assert np.array_equal(env.obs(qlb.LAYOUT_U8_BHYX), ora.obs_u8()) and env.obs().any()
    monkeypatch.setenv("QLC_DEBUG_SKIP", "1")          # would drop the physics
    env2 = qlb.BreakoutEnvironment(n_envs=n, seed=seed)
    env2.step_many(O.synthetic_actions(seed, 0, n, 0, 30))
    assert np.array_equal(env2.read_state()["ball_cy"], ora.state()["ball_cy"])
    env.close(); env2.close(); ora.close()


@pytest.mark.gpu
def test_random_policy_drawn_inside_the_step_kernel(qlb, O):
    """qlc_env_step_random: the uniform action stream of the pure-random phase is drawn by the step kernel; it equals the
    oracle's synthetic stream, and so do the trajectories, for single steps and long launches."""
    torch = pytest.importorskip("torch")
    n, seed = 300, 8
    env = qlb.BreakoutEnvironment(n_envs=n, seed=seed, env_id_base=1000, replay_capacity=n * 16)
    ora = O.VecEnv(n, seed=seed, env_id_base=1000, replay_capacity=n * 16)
    t = 0
    for steps in (1, 1, 3, 40, 1, 64):
        acts = torch.full((steps, n), 9, dtype=torch.uint8, device="cuda")
        rew = torch.empty((steps, n), dtype=torch.float32, device="cuda"); done = torch.empty((steps, n), dtype=torch.uint8, device="cuda")
        env.step_random_device(steps, acts.data_ptr(), rew.data_ptr(), done.data_ptr())
        torch.cuda.synchronize()
        want = O.synthetic_actions(seed, 1000, n, t, steps)
        assert np.array_equal(acts.cpu().numpy(), want)
        for s in range(steps):
            r, d = ora.step(want[s])
            assert np.array_equal(r, rew[s].cpu().numpy()) and np.array_equal(d, done[s].cpu().numpy())
        t += steps
    env.step_random_device(5)                              # no outputs at all
    for s in range(5):
        ora.step(O.synthetic_actions(seed, 1000, n, t + s, 1)[0])
    gs, os_ = env.read_state(), ora.state()
    for k in ("ball_cx", "ball_cy", "pad_min_x", "bricks", "score", "episode_step"):
        assert np.array_equal(gs[k], os_[k]), k
    rb = qlb.ReplayBuffer(env)
    idx = np.arange(0, rb.len(), 11, dtype=np.uint32)
    g, o = rb.get_many(idx, qlb.LAYOUT_U8_BHYX), ora.get_many(idx, "u8")
    assert np.array_equal(g.action, o["action"]) and np.array_equal(g.state_next, o["state_next"])
    env.close(); ora.close()


@pytest.mark.gpu
def test_lives(qlb):
    """lives = 1 while not finished, 0 after the miss (mechanics.rs:131-135: the first miss ends the game)."""
    n = 64
    env = qlb.BreakoutEnvironment(n_envs=n, seed=4, auto_reset=False)
    assert env.lives().tolist() == [1] * n and env.read_state()["lives"].tolist() == [1] * n
    seen_dead = False
    for _ in range(12):
        _, done = env.step_many(np.zeros((40, n), dtype=np.uint8))      # nobody moves the paddle: every ball is eventually missed
        st = env.read_state()
        assert np.array_equal(env.lives(), 1 - st["finished"]) and np.array_equal(st["lives"], 1 - st["finished"])
        seen_dead = seen_dead or bool((env.lives() == 0).any())
    assert seen_dead
    env.reset()
    assert env.lives().tolist() == [1] * n
    env.close()


class HashModel:
    """Deterministic stand-in for DeepQLearningModel over state OBJECTS (handles or pixel copies): it tensorises through the state
    type's own to_multi_dim_array / batch_to_multi_dim_array, like QLearningTensorflowModel does (q_learning_model.rs:111,137,171)."""

    def __init__(self, state_type):
        self.S = state_type
        self.trained = []

    def predict_action(self, state):
        s = state.to_multi_dim_array()
        assert s.shape == (84, 84, 4) and s.dtype == np.float32
        return int(s.sum(dtype=np.float64)) // 96 % 3

    def batch_predict_max_future_reward(self, states):
        s = self.S.batch_to_multi_dim_array(states).reshape(len(states), -1)
        return (s.sum(axis=1, dtype=np.float32) / np.float32(255.0 * 512)).astype(np.float32)

    def train(self, state_batch, action_batch, updated_q):
        s = self.S.batch_to_multi_dim_array(state_batch)
        self.trained.append((hashlib.sha256(np.ascontiguousarray(s).tobytes()).hexdigest(), list(action_batch), [float(x) for x in updated_q]))


def learn_episodes(env, replay, state_type, param, batch_size, draws, dirs, D):
    """SelfDrivingQLearner::learn_episode (self_driving_tf_q_learner.rs:141-233), statement by statement, against whatever
    Environment / state / ReplayBuffer types it is handed. Random draws (thread_rng in the reference) are explicit inputs."""
    model, target = HashModel(state_type), HashModel(state_type)
    rng = np.random.default_rng(99)
    step_count, episode_count, running_reward, epsilon = 0, 0, np.float32(0), float(param["epsilon_max"])
    actions, episode_log, lives = [], [], []
    it = iter(draws)
    for dir_x in dirs:
        env.reset(dir_x)                                                    # :142
        state = env.state_as_rc()                                           # :144
        episode_reward = np.float32(0)
        for _ in range(param["max_steps_per_episode"]):                     # :149
            step_count += 1
            u, a_rand = next(it)
            if step_count < param["epsilon_pure_random_steps"] or epsilon > u:   # :153
                action = int(a_rand)
            else:
                action = model.predict_action(state)
            epsilon = max(epsilon - (param["epsilon_max"] - param["epsilon_min"]) / param["epsilon_greedy_steps"], param["epsilon_min"])
            state_next, reward, done = env.step_as_rc(action)               # :171
            episode_reward = np.float32(episode_reward + np.float32(reward))
            replay.add(action, state, state_next, reward, done)             # :177
            state = state_next
            actions.append(action)
            lives.append(env.lives())
            if step_count % param["update_after_actions"] == 0 and replay.len() > batch_size:   # :181
                indices = D.generate_distinct_random_ids(rng, (0, replay.len()), batch_size)
                samples = replay.get_many(indices)
                mf = target.batch_predict_max_future_reward(samples.state_next)
                q = [np.float32(np.float32(r) + np.float32(m * np.float32(param["gamma"]))) for r, m in zip(samples.reward, mf)]
                for i in range(batch_size):
                    if samples.done[i]:
                        q[i] = np.float32(samples.reward[i])
                model.train(samples.state, samples.action, q)
            if done:
                break
        replay.add_episode_reward(episode_reward)                           # :220
        if episode_count >= param["episode_reward_history_buffer_len"]:
            running_reward = replay.avg_episode_reward()
        episode_count += 1
        episode_log.append((float(episode_reward), float(running_reward), float(replay.min_episode_reward())))
    return actions, model.trained, episode_log, lives


class OraclePixels:
    """the reference's BreakoutState: a full copy of the four frames (Clone = 112,896-byte copy here, f32 [x][y][slot])"""

    def __init__(self, a):
        self.a = a

    def to_multi_dim_array(self):
        return self.a

    @staticmethod
    def batch_to_multi_dim_array(batch):
        return np.stack([s.a for s in batch])


class OracleEnvironment:
    """Environment over the CPU oracle with the reference's semantics: step_as_rc returns a COPY of the pixels."""

    def __init__(self, O, seed):
        self.v = O.VecEnv(1, seed=seed, replay_capacity=8)
        self._state = None
        self._done = False

    def reset(self, dir_x):
        self.v.reset_env(0, dir_x)
        self._state = OraclePixels(self.v.obs_f32()[0].copy())
        self._done = False

    def state_as_rc(self):
        return OraclePixels(self._state.a.copy())

    def step_as_rc(self, action):
        r, d = self.v.step(np.array([action], dtype=np.uint8))
        nxt = self.v.get_many(np.array([self.v.replay_len() - 1], dtype=np.uint32), "f32")["state_next"][0].copy()   # the terminal stack too
        self._state = OraclePixels(nxt)
        self._done = bool(d[0])
        return OraclePixels(nxt.copy()), float(r[0]), bool(d[0])

    def lives(self):
        return 0 if self._done else 1


@pytest.mark.gpu
def test_unchanged_learner_loop_on_the_drop_in_types(qlb, O):
    """learn_episode, literally, once on the oracle (states are pixel copies, the reference's cost model) and once on
    CudaBreakoutEnvironment + the generic ReplayBuffer holding HANDLES: same actions (incl. the greedy ones computed from
    tensorised states), same training batches bit for bit, same TD targets, same episode log, same lives."""
    D = importlib.import_module("q-learning_b200.dropin")
    seed, batch = 21, 8
    param = dict(gamma=0.99, epsilon_max=1.0, epsilon_min=0.1, max_steps_per_episode=230, epsilon_pure_random_steps=60, epsilon_greedy_steps=300.0,
                 history_buffer_len=256, update_after_actions=4, episode_reward_history_buffer_len=3)      # a miss takes ~200 steps: some episodes end, some are cut
    rng = np.random.default_rng(8)
    n_episodes = 7
    draws = [(float(rng.random()), int(rng.integers(0, 3))) for _ in range(n_episodes * 230)]
    dirs = [float(np.float32(-0.35 + 0.2 * rng.random())) for _ in range(n_episodes)]
    ref = learn_episodes(OracleEnvironment(O, seed), D.ReplayBuffer(256, 3), OraclePixels, param, batch, draws, dirs, D)
    env = D.CudaBreakoutEnvironment(84, 84, history_buffer_len=256, seed=seed)
    assert env.episode_reward_goal_mean() == 59.0
    got = learn_episodes(env, D.ReplayBuffer(256, 3), D.CudaBreakoutState, param, batch, draws, dirs, D)
    assert got[0] == ref[0] and len(got[0]) > 600
    assert len(got[1]) == len(ref[1]) > 100
    for g, r in zip(got[1], ref[1]):
        assert g == r
    assert got[2] == ref[2] and got[3] == ref[3] and 0 in got[3]
    # a handle whose frames have left the ring is refused, not served from overwritten slots
    old = D.CudaBreakoutState(env.vector_env, 5, 5)
    assert env.vector_env.time() > 256 + 9
    with pytest.raises(qlb.QlError) as ei:
        D.CudaBreakoutState.batch_to_multi_dim_array([old])
    assert ei.value.code == qlb.ERR_OUT_OF_RANGE and "stale" in str(ei.value)
    env.close()


@pytest.mark.gpu
def test_state_handles_vector_env(qlb, O):
    """qlc_obs_gather(_host) for handles of many envs and times equals the oracle's observation at those times (all layouts)."""
    n, seed, cap = 24, 12, 24 * 64
    env = qlb.BreakoutEnvironment(n_envs=n, seed=seed, replay_capacity=cap, max_episode_steps=25)
    ora = O.VecEnv(n, seed=seed, max_episode_steps=25, replay_capacity=cap)
    snaps = {}
    t = 0
    for steps in (1, 2, 3, 30, 9):
        _run(env, ora, O, seed, n, steps, t)
        t += steps
        snaps[t] = (ora.obs_u8(), ora.obs_f32(), ora.state()["episode_step"].copy())
    hs, want8, want32 = [], [], []
    for tt, (o8, o32, k) in snaps.items():
        for e in (0, 5, n - 1):
            hs.append((tt, int(k[e]), e)); want8.append(o8[e]); want32.append(o32[e])
    h = np.array(hs, dtype=qlb.OBS_HANDLE_DTYPE)
    assert np.array_equal(env.obs_gather(h, qlb.LAYOUT_U8_BHYX), np.stack(want8))
    assert np.array_equal(env.obs_gather(h, qlb.LAYOUT_F32_BXYH), np.stack(want32))
    assert np.array_equal(env.obs_gather(h, qlb.LAYOUT_U8_BXYH), np.stack(want32).astype(np.uint8))
    for bad in ((t + 1, 1, 0), (3, 1, n)):                 # from the future / another env id
        with pytest.raises(qlb.QlError):
            env.obs_gather(np.array([bad], dtype=qlb.OBS_HANDLE_DTYPE))
    env.close(); ora.close()


@pytest.mark.gpu
def test_host_widening_equals_device_f32(qlb, O, monkeypatch):
    """f32 [b][x][y][slot] host gathers cross PCIe as u8 and are widened on the host: same bytes as the oracle for pageable and
    page-locked targets and for a batch large enough to use every pool thread."""
    n, seed = 64, 41
    env = qlb.BreakoutEnvironment(n_envs=n, seed=seed, replay_capacity=n * 32)
    rb = qlb.ReplayBuffer(env)
    ora = O.VecEnv(n, seed=seed, replay_capacity=n * 32)
    _run(env, ora, O, seed, n, 40)
    for batch in (1, 32, 600):
        idx = rb.generate_distinct_random_ids(batch, batch)
        o = ora.get_many(idx, "f32")
        for reuse in (False, True):
            g = rb.get_many(idx, qlb.LAYOUT_F32_BXYH, reuse=reuse)
            assert np.array_equal(g.state, o["state"]) and np.array_equal(g.state_next, o["state_next"])
    assert np.array_equal(env.obs(qlb.LAYOUT_F32_BXYH), ora.obs_f32())
    env.close(); ora.close()


_PIECEWISE = """
import importlib, sys
import numpy as np
sys.path.insert(0, %r)
q = importlib.import_module("q-learning_b200")
from oracle import oracle as O
n, seed = 16, 43
env = q.BreakoutEnvironment(n_envs=n, seed=seed, replay_capacity=n * 32)
rb = q.ReplayBuffer(env)
ora = O.VecEnv(n, seed=seed, replay_capacity=n * 32)
acts = O.synthetic_actions(seed, 0, n, 0, 30)
env.step_many(acts)
for a in acts:
    ora.step(a)
for batch in (1, 32, 200):
    idx = rb.generate_distinct_random_ids(batch, batch)
    o = ora.get_many(idx, "f32")
    for reuse in (False, True):
        g = rb.get_many(idx, q.LAYOUT_F32_BXYH, reuse=reuse)
        assert np.array_equal(g.state, o["state"]) and np.array_equal(g.state_next, o["state_next"]) and np.array_equal(g.reward, o["reward"])
print("piecewise ok")
"""


@pytest.mark.gpu
def test_host_gather_piecewise_copy_path():
    """QLC_HOST_STREAM=0 keeps the older transport of f32 host gathers (device staging, piecewise device->host copies widened as they
    arrive) for A/B measurements; the default is the streamed one (the kernel stores into page-locked memory and raises arrival flags,
    covered by every other host-gather test). Same bytes either way. The switch is read once per process, hence the subprocess."""
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-c", _PIECEWISE % root], env=dict(os.environ, QLC_HOST_STREAM="0"), capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "piecewise ok" in r.stdout, r.stdout + r.stderr


@pytest.mark.gpu
def test_stats_reduction_behind_the_c_abi(qlb):
    """qlc_comm_init / qlc_stats_allreduce / qlc_stats_global on one rank: without NCCL (world 1, no id) and through a real
    one-rank NCCL communicator; the reduced statistics are those of the end of the last launch even when later launches are
    already queued, and enqueueing a reduction after every step does not change the trajectories."""
    torch = pytest.importorskip("torch")
    n, seed = 256, 6
    acts = np.random.default_rng(0).integers(0, 3, size=(50, n), dtype=np.uint8)
    ref = qlb.BreakoutEnvironment(n_envs=n, seed=seed, max_episode_steps=40)
    for use_nccl in (False, True):
        env = qlb.BreakoutEnvironment(n_envs=n, seed=seed, max_episode_steps=40)
        env.comm_init(0, 1, qlb.comm_unique_id() if use_nccl else None)
        info = env.comm_info()
        assert info["world"] == 1 and (info["nccl_ranks"] == 1 and info["nccl_version"] > 20000 if use_nccl else info["nccl_ranks"] == 0)
        env.stats_allreduce()
        assert env.stats_global()["episodes"] == 0 and env.stats_global()["steps"] == 0
        a_dev = torch.from_numpy(acts).cuda()
        stream = torch.cuda.current_stream().cuda_stream
        for s in range(50):
            env.step_device(a_dev[s].data_ptr(), 1, None, None, stream)
            env.stats_allreduce(stream)
        torch.cuda.synchronize()
        g, local = env.stats_global(wait=True), env.stats()
        assert g == local and g["episodes"] >= n and g["steps"] == 50 * n
        # snapshot semantics: enqueue, then queue more launches before looking
        env.step_device(a_dev.data_ptr(), 10, None, None, stream)
        env.stats_allreduce(stream)
        before = None
        env.step_device(a_dev.data_ptr(), 50, None, None, stream)
        g2 = env.stats_global(wait=True)
        torch.cuda.synchronize()
        assert g2["steps"] == 60 * n and g2["episodes"] <= env.stats()["episodes"]
        if ref is not None and not use_nccl:
            ref.step_many(acts); ref.step_many(acts[:10]); ref.step_many(acts)
            assert np.array_equal(ref.read_state()["ball_cx"], env.read_state()["ball_cx"]) and ref.stats() == env.stats()
        env.close()
    ref.close()
    with pytest.raises(qlb.QlError):
        e = qlb.BreakoutEnvironment(n_envs=4)
        try:
            e.stats_allreduce()                            # no communicator
        finally:
            e.close()


@pytest.mark.gpu
def test_qnet_outliving_its_env_is_an_error_not_a_crash(qlb):
    sys.path.insert(0, ROOT)
    bench = importlib.import_module("bench")
    env = qlb.BreakoutEnvironment(n_envs=8, seed=1)
    net = qlb.QNetwork(env, bench._random_qnet_weights(qlb))
    q, a, m = net.forward()
    assert q.shape == (8, 3) and net.error() == 0
    env.close()                                            # the env goes first (Python: env.close() before QNetwork.__del__)
    with pytest.raises(qlb.QlError) as ei:
        net.forward()
    assert "destroyed" in str(ei.value)
    net.close()                                            # still fine
