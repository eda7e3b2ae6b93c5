"""Vectorised twin of the reference's DQN loop — the CALLER of the hot path, kept model-agnostic.

Mirrors `SelfDrivingQLearner` (ql-with-tensorflow/src/learn/self_driving_tf_q_learner.rs:69-233): same `Parameter` defaults
(:50-67), same ε-greedy rule and decay (:153-167), same replay gate (:181), same TD target (:189-199), same episode
bookkeeping (:220-224) and `solved()` (:134-139) — but one *iteration* advances all N envs of a `BreakoutEnvironment` by one
step (N env-steps), treated as N consecutive steps of the reference loop in env order. With N = 1 the sequence of calls and
values is exactly the reference's; for N > 1 the ε decay uses the closed form `max(ε₀ - j·δ, ε_min)` for the j-th env of the
iteration and the minibatches due in an iteration (one per 4 env-steps) are sampled together after the N inserts.

The Q-network is out of scope: `model` / `stabilized_model` are any objects with the three methods of
`DeepQLearningModel` (ml_model/model.rs:29-77) on batched arrays:
    predict_action(states_f32[n,84,84,4]) -> actions u8[n]
    batch_predict_max_future_reward(states_f32[B,84,84,4]) -> f32[B]
    train(state_batch_f32[B,84,84,4], action_batch_u8[B], updated_q_values_f32[B]) -> None
    write_checkpoint(path) -> str            (optional)
Random draws (`rng.gen_range`, thread_rng in the reference) come from a seeded numpy Generator: explicit inputs.
"""
import numpy as np

from . import LAYOUT_F32_BXYH, BreakoutAction, QlError, ReplayBuffer


class Parameter:
    """self_driving_tf_q_learner.rs:20-67"""

    def __init__(self, **kw):
        self.gamma = np.float32(0.99)
        self.epsilon_max = 1.0
        self.epsilon_min = 0.1
        self.max_steps_per_episode = 10_000
        self.epsilon_pure_random_steps = 50_000
        self.epsilon_greedy_steps = 1_000_000.0
        self.history_buffer_len = 1_000_000
        self.update_after_actions = 4
        self.update_target_network_after_num_steps = 10_000      # unused in the reference too (:38)
        self.episode_reward_history_buffer_len = 100
        self.stats_after_steps = 25_000
        self.lowest_episode_reward_goal_threshold_pct = np.float32(0.9)
        for k, v in kw.items():
            if not hasattr(self, k):
                raise QlError("unknown Parameter field %s" % k)
            setattr(self, k, v)

    def epsilon_interval(self):
        return self.epsilon_max - self.epsilon_min


class SelfDrivingQLearner:
    def __init__(self, environment, param, model, stabilized_model, batch_size=32, seed=0, checkpoint_file=None):
        if environment.max_episode_steps != param.max_steps_per_episode or not environment.auto_reset:
            raise QlError("create the BreakoutEnvironment with auto_reset=True and max_episode_steps=param.max_steps_per_episode")
        self.environment = environment
        self.param = param
        self.model = model
        self.stabilized_model = stabilized_model
        self.batch_size = batch_size
        self.checkpoint_file = checkpoint_file
        self.replay_buffer = ReplayBuffer(environment)
        # Parameter::history_buffer_len is the FIFO length the learner trains from (:32-34,:100); the ring holds whole time steps
        # of all envs, so the effective length is history_buffer_len rounded down to a multiple of n_envs (at least one step)
        want = max(param.history_buffer_len // environment.n_envs, 1) * environment.n_envs
        if self.replay_buffer.capacity() != want:
            raise QlError("the environment's replay ring holds %d transitions but Parameter.history_buffer_len asks for %d (%d rounded down to whole "
                          "time steps of %d envs): create the BreakoutEnvironment with replay_capacity=param.history_buffer_len"
                          % (self.replay_buffer.capacity(), want, param.history_buffer_len, environment.n_envs))
        self.effective_history_buffer_len = want
        self.rng = np.random.default_rng(seed)
        self.step_count = 0
        self.episode_count = 0
        self.running_reward = np.float32(0.0)
        self.epsilon = float(param.epsilon_max)
        n = environment.n_envs
        self._episode_reward = np.zeros(n, dtype=np.float32)
        self._episode_step = np.zeros(n, dtype=np.int64)
        self._sample_calls = 0
        self.log = []                      # learning_update_log records (:235-273), as dicts

    # :134-139
    def solved(self):
        goal = np.float32(self.environment.episode_reward_goal_mean())
        if len(self.replay_buffer.episode_rewards()) == 0:
            return False
        return bool(self.running_reward >= goal and
                    np.float32(self.replay_buffer.min_episode_reward()) >= goal * self.param.lowest_episode_reward_goal_threshold_pct)

    def draw(self, n):
        """The iteration's random inputs: u[n] for the ε test (:153), a_rand[n] for the random action (:156)."""
        return self.rng.random(n), self.rng.integers(0, BreakoutAction.ACTION_SPACE, size=n, dtype=np.uint8)

    def learn_iteration(self, draws=None):
        """One step of every env: N consecutive iterations of the reference's inner loop (:149-217)."""
        p, env, n = self.param, self.environment, self.environment.n_envs
        u, a_rand = self.draw(n) if draws is None else draws
        delta = p.epsilon_interval() / p.epsilon_greedy_steps
        step_counts = self.step_count + 1 + np.arange(n)
        if n == 1:
            eps = np.array([self.epsilon])
        else:
            eps = np.maximum(self.epsilon - np.arange(n) * delta, p.epsilon_min)
        random_mask = (step_counts < p.epsilon_pure_random_steps) | (eps > u)
        actions = a_rand.copy()
        if not random_mask.all():
            greedy = np.asarray(self.model.predict_action(env.obs(LAYOUT_F32_BXYH)), dtype=np.uint8)
            actions = np.where(random_mask, a_rand, greedy).astype(np.uint8)
        # ε decay (:164-167), once per env-step
        self.epsilon = max(self.epsilon - delta, p.epsilon_min) if n == 1 else max(self.epsilon - n * delta, p.epsilon_min)
        # step + replay insert (the step kernel appends frame + record) (:171-178)
        state_next, reward, done = env.step(actions)
        self.replay_buffer.add(state_next=state_next)
        self._episode_reward += reward
        self._episode_step += 1
        self.step_count += n
        # training gate (:181): one minibatch per env-step whose running count is a multiple of update_after_actions
        due = int(np.count_nonzero(step_counts % p.update_after_actions == 0))
        trained = []
        if due and self.replay_buffer.len() > self.batch_size:
            for _ in range(due):
                indices = self.replay_buffer.generate_distinct_random_ids(self.batch_size, self._sample_calls)
                self._sample_calls += 1
                sample = self.replay_buffer.get_many(indices, LAYOUT_F32_BXYH, reuse=True)    # page-locked buffers, overwritten by the next minibatch
                max_future = np.asarray(self.stabilized_model.batch_predict_max_future_reward(sample.state_next), dtype=np.float32)
                updated_q = sample.reward + max_future * np.float32(p.gamma)              # add_arrays / array_mul in f32 (:192,:298-315)
                updated_q = np.where(sample.done != 0, sample.reward, updated_q).astype(np.float32)   # terminal steps (:195-199)
                self.model.train(sample.state, sample.action, updated_q)
                trained.append((indices, updated_q))
        # stats / checkpoint every stats_after_steps (:204-212)
        if (self.step_count // p.stats_after_steps) > ((self.step_count - n) // p.stats_after_steps):
            self._checkpoint()
            self.learning_update_log()
        # episode ends (:214-231), in env order
        truncated = (done == 0) & (self._episode_step >= p.max_steps_per_episode)
        for e in np.nonzero((done != 0) | truncated)[0]:
            self.replay_buffer.add_episode_reward(float(self._episode_reward[e]))
            if self.episode_count >= p.episode_reward_history_buffer_len:
                self.running_reward = np.float32(self.replay_buffer.avg_episode_reward())
            self.episode_count += 1
            self._episode_reward[e] = 0.0
            self._episode_step[e] = 0
            if self.solved():
                self._checkpoint()
                self.learning_update_log()
        return actions, reward, done, trained

    def learn_till_mastered(self, max_iterations=None):                                   # :127-132
        it = 0
        while not self.solved():
            self.learn_iteration()
            it += 1
            if max_iterations is not None and it >= max_iterations:
                break
        return it

    def _checkpoint(self):
        if self.checkpoint_file and hasattr(self.model, "write_checkpoint"):
            self.model.write_checkpoint(self.checkpoint_file)

    def learning_update_log(self):                                                        # :235-273 (DBSCAN cosmetics left out)
        rewards = self.replay_buffer.episode_rewards()
        counts = self.replay_buffer.actions()
        total = int(counts.sum())
        rec = {
            "episode": self.episode_count, "steps": self.step_count, "gamma": float(self.param.gamma), "epsilon": self.epsilon,
            "reward_goal_mean": self.environment.episode_reward_goal_mean(),
            "reward_goal_low": float(np.float32(self.environment.episode_reward_goal_mean()) * self.param.lowest_episode_reward_goal_threshold_pct),
            "current_mean": float(self.replay_buffer.avg_episode_reward()) if len(rewards) else None,
            "current_low": float(self.replay_buffer.min_episode_reward()) if len(rewards) else None,
            "action_distribution": {a.name: (100.0 * int(c) / total if total else 0.0) for a, c in zip(BreakoutAction, counts)},
            "actions_in_replay": total,
        }
        self.log.append(rec)
        return rec
