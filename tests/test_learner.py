"""The vectorised learner adapter (q-learning_b200/learner.py) against a literal sequential restatement of
SelfDrivingQLearner::learn_episode (self_driving_tf_q_learner.rs:141-233) driving the CPU oracle env + FIFO replay, with
the same deterministic stand-in model and the same random draws."""
import hashlib
import importlib

import numpy as np
import pytest


class HashModel:
    """Deterministic stand-in for DeepQLearningModel (the Q-network is out of scope): values are functions of the pixels."""

    def __init__(self):
        self.trained = []

    def predict_action(self, states):
        s = np.asarray(states, dtype=np.float32).reshape(len(states), -1)
        return (s.sum(axis=1).astype(np.int64) // 96 % 3).astype(np.uint8)

    def batch_predict_max_future_reward(self, states):
        s = np.asarray(states, dtype=np.float32).reshape(len(states), -1)
        return (s.sum(axis=1, dtype=np.float32) / np.float32(255.0 * 512)).astype(np.float32)

    def train(self, state_batch, action_batch, updated_q):
        self.trained.append((hashlib.sha256(np.ascontiguousarray(state_batch).tobytes()).hexdigest(), action_batch.tolist(), updated_q.tolist()))


def reference_loop(O, param, draws, batch, seed, n_steps_total):
    """learn_episode, literally, for ONE env on the oracle (reset / step_as_rc / replay.add / gate / TD target / episode end)."""
    env = O.VecEnv(1, seed=seed, max_episode_steps=param.max_steps_per_episode, replay_capacity=param.history_buffer_len,
                   episode_window=param.episode_reward_history_buffer_len)
    model, target = HashModel(), HashModel()
    step_count, episode_count, running_reward, epsilon, calls = 0, 0, np.float32(0), float(param.epsilon_max), 0
    goal = np.float32(O.lib().orc_env_goal_mean())
    actions, log = [], []
    it = 0
    while it < n_steps_total:
        episode_reward = np.float32(0)
        for _ in range(param.max_steps_per_episode):
            if it >= n_steps_total:
                break
            u, a_rand = draws[it]
            it += 1
            step_count += 1
            if step_count < param.epsilon_pure_random_steps or epsilon > u[0]:
                a = int(a_rand[0])
            else:
                a = int(model.predict_action(env.obs_f32())[0])
            epsilon = max(epsilon - param.epsilon_interval() / param.epsilon_greedy_steps, param.epsilon_min)
            r, d = env.step(np.array([a], dtype=np.uint8))
            actions.append(a)
            episode_reward = np.float32(episode_reward + r[0])
            if step_count % param.update_after_actions == 0 and env.replay_len() > batch:
                idx = O.sample_distinct(seed, calls, env.replay_len(), batch)
                calls += 1
                smp = env.get_many(idx, "f32")
                mf = target.batch_predict_max_future_reward(smp["state_next"])
                q = smp["reward"] + mf * np.float32(param.gamma)
                q = np.where(smp["done"] != 0, smp["reward"], q).astype(np.float32)
                model.train(smp["state"], smp["action"], q)
            if d[0]:
                break
        else:
            pass
        if it >= n_steps_total and not d[0] and env.state()["episode_step"][0] != 0:
            break      # ran out of budget mid-episode: no episode bookkeeping
        log.append(float(episode_reward))
        if episode_count >= param.episode_reward_history_buffer_len:
            running_reward = np.float32(env.avg_episode_reward())
        episode_count += 1
    return actions, model.trained, log, episode_count, running_reward, epsilon


@pytest.mark.gpu
def test_single_env_matches_the_reference_loop(qlb, O):
    L = importlib.import_module("q-learning_b200.learner")
    seed, batch, total = 13, 8, 1500
    param = L.Parameter(max_steps_per_episode=120, epsilon_pure_random_steps=60, epsilon_greedy_steps=300.0, history_buffer_len=256,
                        update_after_actions=4, episode_reward_history_buffer_len=5, stats_after_steps=500)
    rng = np.random.default_rng(5)
    draws = [(rng.random(1), rng.integers(0, 3, size=1, dtype=np.uint8)) for _ in range(total)]
    ref_actions, ref_trained, ref_log, ref_eps, ref_running, ref_epsilon = reference_loop(O, param, draws, batch, seed, total)

    env = qlb.BreakoutEnvironment(n_envs=1, seed=seed, replay_capacity=param.history_buffer_len, max_episode_steps=param.max_steps_per_episode,
                                  episode_window=param.episode_reward_history_buffer_len)
    model, target = HashModel(), HashModel()
    learner = L.SelfDrivingQLearner(env, param, model, target, batch_size=batch, seed=0)
    got_actions = []
    for it in range(total):
        a, r, d, trained = learner.learn_iteration(draws[it])
        got_actions.append(int(a[0]))
    assert got_actions == ref_actions
    assert len(model.trained) == len(ref_trained) > 100
    for g, r in zip(model.trained, ref_trained):
        assert g == r                                     # state batch bits, actions, TD targets
    assert learner.replay_buffer.episode_rewards().tolist()[-len(ref_log):] == ref_log[-5:][-len(learner.replay_buffer.episode_rewards()):]
    assert learner.epsilon == ref_epsilon and learner.step_count == total
    assert len(learner.log) == total // 500 and learner.log[-1]["steps"] == 1500
    assert not learner.solved()
    assert any(a != b for a, b in zip(got_actions[100:], [int(d[1][0]) for d in draws[100:]])), "greedy path never taken"
    env.close()


@pytest.mark.gpu
def test_vector_loop_runs_and_keeps_the_train_ratio(qlb):
    L = importlib.import_module("q-learning_b200.learner")
    n, batch, iters = 64, 32, 200
    param = L.Parameter(max_steps_per_episode=300, epsilon_pure_random_steps=2000, epsilon_greedy_steps=200000.0, history_buffer_len=n * 64,
                        stats_after_steps=4096)
    env = qlb.BreakoutEnvironment(n_envs=n, seed=3, replay_capacity=param.history_buffer_len, max_episode_steps=300)
    model, target = HashModel(), HashModel()
    learner = L.SelfDrivingQLearner(env, param, model, target, batch_size=batch, seed=1)
    for _ in range(iters):
        learner.learn_iteration()
    assert learner.step_count == iters * n
    assert len(model.trained) == iters * n // 4            # one minibatch per 4 env-steps, like the reference
    assert learner.episode_count == env.stats()["episodes"] > 0
    assert learner.epsilon == pytest.approx(1.0 - iters * n * 0.9 / 200000.0)
    assert len(learner.log) == iters * n // 4096 and sum(learner.log[-1]["action_distribution"].values()) == pytest.approx(100.0)
    assert len(learner.replay_buffer.episode_rewards()) == min(100, learner.episode_count)
    env.close()


def test_parameter_defaults(qlb):
    L = importlib.import_module("q-learning_b200.learner")
    p = L.Parameter()
    assert (float(p.gamma), p.epsilon_max, p.epsilon_min, p.max_steps_per_episode, p.epsilon_pure_random_steps, p.epsilon_greedy_steps,
            p.history_buffer_len, p.update_after_actions, p.episode_reward_history_buffer_len, p.stats_after_steps) == \
        (pytest.approx(0.99), 1.0, 0.1, 10_000, 50_000, 1_000_000.0, 1_000_000, 4, 100, 25_000)      # :50-67
    with pytest.raises(qlb.QlError):
        L.Parameter(nope=1)


@pytest.mark.gpu
def test_zero_copy_tensors_equal_the_host_path(qlb, O):
    """§8f-2: the torch hand-off (gather kernels writing into CUDA tensors) gives the bytes of the host-buffer ABI."""
    torch = pytest.importorskip("torch")
    tio = importlib.import_module("q-learning_b200.torch_io")
    n = 96
    env = qlb.BreakoutEnvironment(n_envs=n, seed=8, replay_capacity=n * 32)
    rb = qlb.ReplayBuffer(env)
    acts = torch.randint(0, 3, (50, n), dtype=torch.uint8, device="cuda")
    reward, done = tio.step(env, acts)
    torch.cuda.synchronize()
    assert float(reward.sum()) >= 0 and rb.len() == n * 32
    for layout in (qlb.LAYOUT_F32_BXYH, qlb.LAYOUT_U8_BHYX):
        smp = tio.DeviceSampler(rb, 32, 4, layout).sample(11)
        torch.cuda.synchronize()
        idx = smp.indices.cpu().numpy().view(np.uint32)
        for b in range(4):
            assert np.array_equal(idx[b], O.sample_distinct(8, 11 + b, rb.len(), 32))
            host = rb.get_many(idx[b], layout)
            assert np.array_equal(smp.state[b].cpu().numpy(), host.state) and np.array_equal(smp.state_next[b].cpu().numpy(), host.state_next)
            assert np.array_equal(smp.reward[b].cpu().numpy(), host.reward) and np.array_equal(smp.action[b].cpu().numpy(), host.action)
            assert np.array_equal(smp.done[b].cpu().numpy(), host.done)
        assert np.array_equal(tio.observe(env, layout).cpu().numpy(), env.obs(layout))
    env.close()


@pytest.mark.gpu
def test_torch_dqn_example_runs_on_device(qlb):
    """examples/dqn_breakout_torch.py: env + replay (this repo) feeding a torch Q-network with no host round trip."""
    pytest.importorskip("torch")
    import importlib.util
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    spec = importlib.util.spec_from_file_location("dqn_example", os.path.join(root, "examples", "dqn_breakout_torch.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    out = mod.run(n_envs=256, iterations=60, batch=32, quiet=True)
    assert out["env_steps"] == 256 * 60 and out["train_calls"] >= 50 and np.isfinite(out["last_loss"])
    assert out["epsilon"] < 1.0 and out["error_flags"] & ~qlb.ENVERR_DEGENERATE == 0
    out = mod.run(n_envs=256, iterations=60, batch=32, quiet=True, tensor_core_actor=True)     # greedy actions from the tcgen05 Q-network
    assert out["env_steps"] == 256 * 60 and out["train_calls"] >= 50 and np.isfinite(out["last_loss"])


@pytest.mark.gpu
def test_tensor_core_actor_follows_a_torch_network(qlb, O):
    """torch_io.TensorCoreActor: weights of a torch network with the reference architecture, copied into the library's
    tcgen05 Q-network, give the torch network's greedy actions (wherever its top-2 gap is clear) and max-Q of sampled
    next states — read from the frame ring by index, never gathered. Tolerance 4e-2 * max|Q| (bf16 operands vs fp32)."""
    torch = pytest.importorskip("torch")
    import importlib
    tio = importlib.import_module("q-learning_b200.torch_io")
    n = 256
    env = qlb.BreakoutEnvironment(n_envs=n, seed=3, replay_capacity=n * 16)
    rb = qlb.ReplayBuffer(env)
    env.step_many(np.random.default_rng(0).integers(0, 3, size=(9, n), dtype=np.uint8))
    torch.manual_seed(0)
    nn = torch.nn
    net = nn.Sequential(nn.Conv2d(4, 32, 8, stride=4), nn.ReLU(), nn.Conv2d(32, 64, 4, stride=2), nn.ReLU(), nn.Conv2d(64, 64, 3, stride=1), nn.ReLU(),
                        nn.Flatten(), nn.Linear(3136, 512), nn.ReLU(), nn.Linear(512, 3)).cuda()
    actor = tio.TensorCoreActor(env, net)
    with torch.no_grad():
        ref = net(tio.observe(env, qlb.LAYOUT_F32_BXYH).permute(0, 3, 1, 2))
    act = actor.predict_action(want_q=True)
    torch.cuda.synchronize()
    scale = float(ref.abs().max())
    assert float((actor.q - ref).abs().max()) <= 4e-2 * scale
    top2 = ref.topk(2, dim=1).values
    clear = (top2[:, 0] - top2[:, 1]) > 8e-2 * scale
    assert bool(clear.any()) and torch.equal(act[0][clear], ref.argmax(dim=1).to(torch.uint8)[clear])
    smp = tio.DeviceSampler(rb, 64).sample(0)
    with torch.no_grad():
        ref_next = net(smp.state_next.view(64, 84, 84, 4).permute(0, 3, 1, 2)).max(dim=1).values
    got = actor.max_future_reward(smp.indices.view(-1))
    torch.cuda.synchronize()
    assert float((got - ref_next).abs().max()) <= 4e-2 * scale
    with torch.no_grad():                                                   # new weights take effect after sync()
        for p in net.parameters():
            p.mul_(0.5)
        ref2 = net(tio.observe(env, qlb.LAYOUT_F32_BXYH).permute(0, 3, 1, 2))
    actor.sync(net)
    actor.predict_action(want_q=True)
    torch.cuda.synchronize()
    assert float((actor.q - ref2).abs().max()) <= 4e-2 * float(ref2.abs().max())
    actor.close(); env.close()
