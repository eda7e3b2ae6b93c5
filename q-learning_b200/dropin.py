"""The single-env drop-in, 1:1 with the reference's types — what the UNCHANGED learner (SelfDrivingQLearner::learn_episode,
ql-with-tensorflow/src/learn/self_driving_tf_q_learner.rs:141-233) holds and calls, host side in Python over the C ABI.
(The Rust crate of the same shape is bindings/rust/ql-cuda; the C++ one include/ql_cuda.hpp.)

    reference                                                         here
    ----------------------------------------------------------------  ---------------------------------------
    BreakoutEnvironment: Environment   breakout_environment.rs:131-207  CudaBreakoutEnvironment
    BreakoutState: Clone + ToMultiDimArray          :24-78               CudaBreakoutState  (a HANDLE: env, time, k)
    Buffer<T>, ReplayBuffer<S, A>, BufferSample     replay_buffer.rs     Buffer, ReplayBuffer, BufferSample  (generic, host)
    generate_distinct_random_ids(rng, range)        learner :276-296     generate_distinct_random_ids  (the learner's own fn)

Why the reference's generic ReplayBuffer<Rc<E::S>, E::A> can stay exactly what it is: `E::S` is a handle, so `Rc<S>` / `S::clone`
copy two integers instead of 4 x 7,056 pixels, the five deques hold handles and scalars (the reference's own cost model:
pointers and scalars), and the pixels of a minibatch move exactly once — on the GPU, from the HBM frame ring, when the model
calls `S::batch_to_multi_dim_array(batch)` (q_learning_model.rs:137,171), which is ONE gather kernel (qlc_obs_gather_host).
A handle stays valid for `replay_capacity` further steps of its env; the ring is created with
`replay_capacity >= Parameter::history_buffer_len`, so every handle the FIFO still holds is alive.
"""
import numpy as np

from . import (ERR_OUT_OF_RANGE, FRAME_H, FRAME_W, LAYOUT_F32_BXYH, NUM_FRAMES, OBS_HANDLE_DTYPE, BreakoutAction, BreakoutEnvironment, QlError, _check,
               _np_ptr)

DEFAULT_HISTORY_BUFFER_LEN = 1_000_000      # Parameter::default().history_buffer_len (self_driving_tf_q_learner.rs:59)


class CudaBreakoutState:
    """BreakoutState (breakout_environment.rs:24-28) as a handle on frames in the HBM ring: `time` env-steps taken, `k` of
    them in the current episode. Clone / Rc::new(state.clone()) (prelude.rs:36,57) copies the handle."""
    __slots__ = ("_env", "time", "k")

    def __init__(self, env, time, k):
        self._env, self.time, self.k = env, int(time), int(k)

    def clone(self):
        return CudaBreakoutState(self._env, self.time, self.k)

    def dims(self):
        return [FRAME_W, FRAME_H, NUM_FRAMES]                                   # model_dims, breakout_environment.rs:148

    def to_multi_dim_array(self):
        """[x][y][slot] f32 (breakout_environment.rs:42-54)"""
        return CudaBreakoutState.batch_to_multi_dim_array([self])[0]

    @staticmethod
    def batch_to_multi_dim_array(batch):
        """[b][x][y][slot] f32, value = u8 as f32 (breakout_environment.rs:56-77) — one gather kernel for the whole batch."""
        if len(batch) == 0:
            raise QlError("empty batch")
        env = batch[0]._env
        h = np.empty(len(batch), dtype=OBS_HANDLE_DTYPE)
        for i, s in enumerate(batch):
            if s._env is not env:
                raise QlError("states of different environments in one batch")
            h[i] = (s.time, s.k, 0)
        out = np.empty((len(batch), FRAME_W, FRAME_H, NUM_FRAMES), dtype=np.float32)
        _check(env._L.qlc_obs_gather_host(env._h, _np_ptr(h), len(batch), LAYOUT_F32_BXYH, _np_ptr(out)))
        return out

    def one_line_info(self):                                                    # DebugVisualizer, breakout_environment.rs:81-89
        s = self._env.read_state()
        return "Breakout [%d bricks, ball_pos: [%.1f %.1f], panel_pos: [%.1f 570.0]]" % (
            bin(int(s["bricks"][0])).count("1"), s["ball_cx"][0], s["ball_cy"][0], (s["pad_min_x"][0] + s["pad_max_x"][0]) / 2)


class CudaBreakoutEnvironment:
    """ONE Breakout env on the GPU behind ql::prelude::Environment (prelude.rs:21-63; breakout_environment.rs:131-207)."""

    def __init__(self, frame_size_x=FRAME_W, frame_size_y=FRAME_H, history_buffer_len=DEFAULT_HISTORY_BUFFER_LEN, seed=0, device=0):
        # auto_reset off, no step limit: the learner resets (learn_episode :142) and counts the steps of an episode itself (:149)
        self._env = BreakoutEnvironment(n_envs=1, frame_size_x=frame_size_x, frame_size_y=frame_size_y, device=device, seed=seed,
                                        replay_capacity=history_buffer_len, max_episode_steps=0, auto_reset=False)
        self._state = CudaBreakoutState(self._env, 0, 0)
        self._a = np.zeros((1, 1), dtype=np.uint8)
        self._r = np.zeros((1, 1), dtype=np.float32)
        self._d = np.zeros((1, 1), dtype=np.uint8)

    def close(self):
        self._env.close()

    def reset(self, dir_x=None):
        """:177-180. `dir_x`: the one random draw of the mechanics (mechanics.rs:103) as an explicit input (tests)."""
        self._env.reset(dir_x=None if dir_x is None else np.array([dir_x], dtype=np.float32))
        self._state = CudaBreakoutState(self._env, self._env.time(), 0)

    def state(self):                                                            # :182
        return self._state

    def state_as_rc(self):                                                      # prelude.rs:36
        return self._state.clone()

    def step(self, action):                                                     # :184-201
        self._a[0, 0] = BreakoutAction.try_from_numeric(int(action)).numeric()
        self._env.step_many(self._a, out=(self._r, self._d))
        self._state = CudaBreakoutState(self._env, self._state.time + 1, self._state.k + 1)
        return self._state, float(self._r[0, 0]), bool(self._d[0, 0])

    def step_as_rc(self, action):                                               # prelude.rs:52-58
        s, r, d = self.step(action)
        return s.clone(), r, d

    def episode_reward_goal_mean(self):                                         # :203-206
        return self._env.episode_reward_goal_mean()

    def lives(self):
        """1 while the episode runs, 0 once it is over: the reference game ends with the first miss (mechanics.rs:131-135)."""
        return int(self._env.lives()[0])

    @property
    def vector_env(self):
        return self._env


class Buffer:
    """Buffer<T> (replay_buffer.rs:5-50): bounded FIFO, index 0 = oldest. A ring over a Python list, so that `get_many` is O(1)
    per index at any length (a collections.deque walks from an end)."""

    def __init__(self, max_buffer_len):
        assert max_buffer_len > 0
        self.max_buffer_len = max_buffer_len
        self._items = []
        self._head = 0                                  # position of the oldest element once the ring is full

    def len(self):
        return len(self._items)

    __len__ = len

    @property
    def buffer(self):
        """the elements, oldest first (the reference's public `buffer: VecDeque<T>`)"""
        return self._items[self._head:] + self._items[:self._head]

    def add(self, element):
        if len(self._items) >= self.max_buffer_len:     # pop_front + push_back
            self._items[self._head] = element
            self._head = (self._head + 1) % self.max_buffer_len
        else:
            self._items.append(element)

    def get_many(self, indices):
        n = len(self._items)
        for i in indices:
            if not 0 <= i < n:
                raise QlError("index out of range", ERR_OUT_OF_RANGE)
        return [self._items[(self._head + i) % n] for i in indices]

    get_many_as_val = get_many


class BufferSample:
    """BufferSample (replay_buffer.rs:140-146)"""
    __slots__ = ("state", "state_next", "reward", "action", "done")

    def __init__(self, state, state_next, reward, action, done):
        self.state, self.state_next, self.reward, self.action, self.done = state, state_next, reward, action, done


class ReplayBuffer:
    """ReplayBuffer<S, A> (replay_buffer.rs:53-137), generic like the reference: S is whatever the environment hands out —
    with CudaBreakoutState that is a handle, and this struct is all the host ever keeps of a transition."""

    def __init__(self, step_buffer_len, episode_reward_buffer_len):
        self.action_history = Buffer(step_buffer_len)
        self.state_history = Buffer(step_buffer_len)
        self.state_next_history = Buffer(step_buffer_len)
        self.reward_history = Buffer(step_buffer_len)
        self.done_history = Buffer(step_buffer_len)
        self.episode_reward_history = Buffer(episode_reward_buffer_len)

    def len(self):
        return self.done_history.len()

    __len__ = len

    def add(self, action, state, state_next, reward, done):
        self.action_history.add(action)
        self.state_history.add(state)
        self.state_next_history.add(state_next)
        self.reward_history.add(reward)
        self.done_history.add(done)

    def add_episode_reward(self, episode_reward):
        self.episode_reward_history.add(np.float32(episode_reward))

    def avg_episode_reward(self):
        assert self.episode_reward_history.len() > 0
        s = np.float32(0)
        for v in self.episode_reward_history.buffer:          # f32 running sum, front to back, like iter().sum::<f32>()
            s = np.float32(s + v)
        return np.float32(s / np.float32(self.episode_reward_history.len()))

    def min_episode_reward(self):
        assert self.episode_reward_history.len() > 0
        return min(self.episode_reward_history.buffer)

    def actions(self):
        return self.action_history

    def episode_rewards(self):
        return list(self.episode_reward_history.buffer)

    def get_many(self, indices):
        return BufferSample(self.state_history.get_many(indices), self.state_next_history.get_many(indices),
                            self.reward_history.get_many_as_val(indices), self.action_history.get_many_as_val(indices),
                            self.done_history.get_many_as_val(indices))


def generate_distinct_random_ids(rng, range_, batch_size):
    """The learner's own private function (:276-296): BATCH_SIZE distinct uniform ids by rejection, on the host — it stays with
    the learner; only the pixel gather behind batch_to_multi_dim_array goes to the device. `rng`: numpy Generator."""
    start, end = range_
    assert end - start >= batch_size
    result = []
    for _ in range(batch_size):
        while True:
            x = int(rng.integers(start, end))
            if x not in result:
                result.append(x)
                break
    return result
