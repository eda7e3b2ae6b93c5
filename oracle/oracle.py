"""TEST INFRASTRUCTURE ONLY — ctypes loader for the CPU oracle (oracle/liboracle.so).

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference leg may import this
module. The product package (q-learning_b200/) never does; its ops fail loudly without the CUDA library.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

FRAME_W = 84
FRAME_H = 84
FRAME_BYTES = FRAME_W * FRAME_H
NUM_FRAMES = 4

ERR_WALL_DISTANCE = 1
ERR_APPROX_RANGE = 2
ERR_RECURSION = 4
ERR_BISECTION = 8
ERR_DEGENERATE = 16


def build(force=False):
    """Compile liboracle.so + selftest with the Makefile in oracle/ (gcc, seconds)."""
    so = os.path.join(_HERE, "liboracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".c", ".h")) or f == "Makefile"]
    if force or not os.path.exists(so) or not os.path.exists(os.path.join(_HERE, "selftest")) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        try:
            subprocess.run(["make", "-C", _HERE, "all"], check=True, stdout=subprocess.DEVNULL)
        except (subprocess.CalledProcessError, OSError):
            if force or not os.path.exists(so):
                raise
    return so


class _V2(C.Structure):
    _fields_ = [("x", C.c_float), ("y", C.c_float)]


class _Circle(C.Structure):
    _fields_ = [("center", _V2), ("radius", C.c_float)]


class _Aabb(C.Structure):
    _fields_ = [("min", _V2), ("max", _V2)]


class _Surface(C.Structure):
    _fields_ = [("way", C.c_float), ("approximation", C.c_float), ("surface_normal", _V2)]


def lib():
    global _LIB
    if _LIB is None:
        so = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(so):
            build()
        L = C.CDLL(so)
        vp, u8p, f32p, u32p, u64p = C.c_void_p, C.POINTER(C.c_uint8), C.POINTER(C.c_float), C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)
        L.orc_vec_new.restype = vp
        L.orc_vec_new.argtypes = [C.c_uint32, C.c_uint64, C.c_uint32, C.c_uint32, C.c_size_t, C.c_size_t]
        L.orc_vec_free.argtypes = [vp]
        L.orc_vec_replay.restype = vp
        L.orc_vec_replay.argtypes = [vp]
        L.orc_vec_reset_env.argtypes = [vp, C.c_uint32, C.c_float]
        L.orc_vec_step.argtypes = [vp, vp, vp, vp]
        L.orc_vec_stats.argtypes = [vp, vp]
        L.orc_vec_read_state.argtypes = [vp] + [vp] * 12
        L.orc_vec_read_obs_u8.argtypes = [vp, vp]
        L.orc_vec_read_obs_f32.argtypes = [vp, vp]
        L.orc_replay_len.restype = C.c_size_t
        L.orc_replay_len.argtypes = [vp]
        L.orc_replay_get_many_f32.argtypes = [vp, vp, C.c_size_t, vp, vp, vp, vp, vp]
        L.orc_replay_get_many_u8.argtypes = [vp, vp, C.c_size_t, vp, vp, vp, vp, vp]
        L.orc_replay_action_histogram.argtypes = [vp, vp]
        L.orc_replay_episode_rewards.restype = C.c_size_t
        L.orc_replay_episode_rewards.argtypes = [vp, vp]
        L.orc_replay_avg_episode_reward.restype = C.c_float
        L.orc_replay_avg_episode_reward.argtypes = [vp]
        L.orc_replay_min_episode_reward.restype = C.c_float
        L.orc_replay_min_episode_reward.argtypes = [vp]
        L.orc_sample_distinct.restype = C.c_int
        L.orc_sample_distinct.argtypes = [C.c_uint64, C.c_uint64, C.c_uint32, C.c_uint32, vp]
        L.orc_synthetic_action.restype = C.c_uint8
        L.orc_synthetic_action.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
        L.orc_reset_dir_x.restype = C.c_float
        L.orc_reset_dir_x.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32]
        L.orc_dir_x_from_bits.restype = C.c_float
        L.orc_dir_x_from_bits.argtypes = [C.c_uint32]
        L.orc_env_goal_mean.restype = C.c_float
        L.orc_bench_env_steps.restype = C.c_double
        L.orc_bench_env_steps.argtypes = [C.c_uint32, C.c_uint32, C.c_uint64, C.c_int, vp]
        L.orc_parts_run.argtypes = [vp, vp, vp, C.c_int, C.c_uint32, C.c_uint32, vp, vp, vp]
        L.orc_synthetic_actions_fill.argtypes = [C.c_uint64, C.c_uint32, C.c_uint32, C.c_uint32, C.c_uint32, vp]
        L.orc_bench_actor_loop.restype = C.c_double
        L.orc_bench_actor_loop.argtypes = [C.c_uint32, C.c_size_t, C.c_uint32, C.c_uint64, vp]
        L.orc_bench_sample.restype = C.c_double
        L.orc_bench_sample.argtypes = [C.c_uint32, C.c_size_t, C.c_uint32, C.c_uint32, C.c_uint64, vp]
        for name in ("orc_collision_test_left_wall", "orc_collision_test_right_wall", "orc_collision_test_top_wall"):
            f = getattr(L, name)
            f.restype = C.c_int
            f.argtypes = [C.POINTER(_Circle), _V2, C.POINTER(_Surface), u32p]
        L.orc_collision_check_with_rectangle.restype = C.c_int
        L.orc_collision_check_with_rectangle.argtypes = [C.POINTER(_Circle), _V2, C.POINTER(_Aabb), C.POINTER(_Surface), u32p]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def synthetic_actions(seed, env_id_base, n_envs, t0, n_steps):
    """[n_steps][n_envs] u8 synthetic random policy (Philox stream 'ACTI'), the bench's action stream."""
    out = np.empty((n_steps, n_envs), dtype=np.uint8)
    lib().orc_synthetic_actions_fill(seed, env_id_base, n_envs, t0, n_steps, _p(out))
    return out


def collision_wall(which, center, radius, mv):
    """(some, way, approximation, nx, ny, err) of Ball::collision_test_{left,right,top}_wall."""
    L = lib()
    ball = _Circle(_V2(*center), radius)
    out = _Surface()
    err = C.c_uint32(0)
    f = getattr(L, "orc_collision_test_%s_wall" % which)
    some = f(C.byref(ball), _V2(*mv), C.byref(out), C.byref(err))
    return some, out.way, out.approximation, out.surface_normal.x, out.surface_normal.y, err.value


def collision_rect(center, radius, mv, rmin, rmax):
    """(some, way, approximation, nx, ny, err) of Ball::collision_check_with_rectangle."""
    L = lib()
    ball = _Circle(_V2(*center), radius)
    box = _Aabb(_V2(*rmin), _V2(*rmax))
    out = _Surface()
    err = C.c_uint32(0)
    some = L.orc_collision_check_with_rectangle(C.byref(ball), _V2(*mv), C.byref(box), C.byref(out), C.byref(err))
    return some, out.way, out.approximation, out.surface_normal.x, out.surface_normal.y, err.value


def collision_rect_batch(cases):
    a = np.ascontiguousarray(cases, dtype=np.float32)
    out = np.empty((a.shape[0], 6), dtype=np.float32)
    L = lib()
    L.orc_collision_rect_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_uint32]
    L.orc_collision_rect_batch(_p(a), _p(out), a.shape[0])
    return (out[:, 0] != 0).astype(np.uint8), out[:, 1:5].copy(), out[:, 5].copy().view(np.uint32)


def sample_distinct(seed, call, length, batch):
    out = np.empty(batch, dtype=np.uint32)
    rc = lib().orc_sample_distinct(seed, call, length, batch, _p(out))
    if rc != 0:
        raise ValueError("range smaller than batch")
    return out


class VecEnv:
    """N Breakout envs + optional FIFO replay, driven in env-index order per time step."""

    def __init__(self, n_envs, seed=0, env_id_base=0, max_episode_steps=0, replay_capacity=0, episode_window=100):
        self.L = lib()
        self.n = n_envs
        self.h = self.L.orc_vec_new(n_envs, seed, env_id_base, max_episode_steps, replay_capacity, episode_window)
        self.replay = self.L.orc_vec_replay(self.h) if replay_capacity else None

    def close(self):
        if self.h:
            self.L.orc_vec_free(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def reset_env(self, e, dir_x):
        self.L.orc_vec_reset_env(self.h, e, dir_x)

    def step(self, actions):
        a = np.ascontiguousarray(actions, dtype=np.uint8)
        assert a.shape == (self.n,)
        reward = np.empty(self.n, dtype=np.float32)
        done = np.empty(self.n, dtype=np.uint8)
        self.L.orc_vec_step(self.h, _p(a), _p(reward), _p(done))
        return reward, done

    def state(self):
        n = self.n
        f = {k: np.empty(n, dtype=np.float32) for k in ("ball_cx", "ball_cy", "ball_dx", "ball_dy", "pad_min_x", "pad_max_x", "pad_speed")}
        f["bricks"] = np.empty(n, dtype=np.uint64)
        f["score"] = np.empty(n, dtype=np.uint32)
        f["finished"] = np.empty(n, dtype=np.uint8)
        f["episode_step"] = np.empty(n, dtype=np.uint32)
        f["err"] = np.empty(n, dtype=np.uint32)
        order = ("ball_cx", "ball_cy", "ball_dx", "ball_dy", "pad_min_x", "pad_max_x", "pad_speed", "bricks", "score", "finished", "episode_step", "err")
        self.L.orc_vec_read_state(self.h, *[_p(f[k]) for k in order])
        return f

    def obs_u8(self):
        out = np.empty((self.n, NUM_FRAMES, FRAME_H, FRAME_W), dtype=np.uint8)
        self.L.orc_vec_read_obs_u8(self.h, _p(out))
        return out

    def obs_f32(self):
        out = np.empty((self.n, FRAME_W, FRAME_H, NUM_FRAMES), dtype=np.float32)
        self.L.orc_vec_read_obs_f32(self.h, _p(out))
        return out

    def stats(self):
        out = np.empty(5, dtype=np.float64)
        self.L.orc_vec_stats(self.h, _p(out))
        return dict(sum_return=out[0], episodes=int(out[1]), min_return=out[2], max_return=out[3], steps=int(out[4]))

    # replay
    def replay_len(self):
        return self.L.orc_replay_len(self.replay)

    def get_many(self, indices, layout="f32"):
        idx = np.ascontiguousarray(indices, dtype=np.uint32)
        n = idx.size
        reward = np.empty(n, dtype=np.float32)
        action = np.empty(n, dtype=np.uint8)
        done = np.empty(n, dtype=np.uint8)
        if layout == "f32":
            s = np.empty((n, FRAME_W, FRAME_H, NUM_FRAMES), dtype=np.float32)
            sn = np.empty_like(s)
            self.L.orc_replay_get_many_f32(self.replay, _p(idx), n, _p(s), _p(sn), _p(reward), _p(action), _p(done))
        else:
            s = np.empty((n, NUM_FRAMES, FRAME_H, FRAME_W), dtype=np.uint8)
            sn = np.empty_like(s)
            self.L.orc_replay_get_many_u8(self.replay, _p(idx), n, _p(s), _p(sn), _p(reward), _p(action), _p(done))
        return dict(state=s, state_next=sn, reward=reward, action=action, done=done)

    def action_histogram(self):
        out = np.zeros(3, dtype=np.uint64)
        self.L.orc_replay_action_histogram(self.replay, _p(out))
        return out

    def episode_rewards(self):
        out = np.empty(4096, dtype=np.float32)
        n = self.L.orc_replay_episode_rewards(self.replay, _p(out))
        return out[:n].copy()

    def avg_episode_reward(self):
        return self.L.orc_replay_avg_episode_reward(self.replay)

    def min_episode_reward(self):
        return self.L.orc_replay_min_episode_reward(self.replay)


class ShardedVecEnv:
    """A shard of independent envs (no replay) split into sub-shards that keep their global env ids, stepped on one host
    thread each — the same trajectories as one VecEnv, for full-size parity runs (4,096 envs x 10,000 steps)."""

    def __init__(self, n_envs, seed=0, env_id_base=0, parts=None):
        parts = max(1, min(n_envs, parts or 4 * (os.cpu_count() or 1)))
        self.n = n_envs
        bounds = [n_envs * p // parts for p in range(parts + 1)]
        self.offsets = np.array(bounds[:-1], dtype=np.uint32)
        self.sizes = np.array([bounds[p + 1] - bounds[p] for p in range(parts)], dtype=np.uint32)
        self.parts = [VecEnv(int(self.sizes[p]), seed=seed, env_id_base=env_id_base + int(self.offsets[p])) for p in range(parts)]
        self.handles = (C.c_void_p * parts)(*[v.h for v in self.parts])

    def run(self, actions):
        """actions [T][n] -> reward [T][n] f32, done [T][n] u8"""
        a = np.ascontiguousarray(actions, dtype=np.uint8)
        assert a.ndim == 2 and a.shape[1] == self.n
        reward = np.empty(a.shape, dtype=np.float32)
        done = np.empty(a.shape, dtype=np.uint8)
        lib().orc_parts_run(self.handles, _p(self.offsets), _p(self.sizes), len(self.parts), self.n, a.shape[0], _p(a), _p(reward), _p(done))
        return reward, done

    def state(self):
        st = [v.state() for v in self.parts]
        return {k: np.concatenate([s[k] for s in st]) for k in st[0]}

    def obs_u8(self):
        return np.concatenate([v.obs_u8() for v in self.parts])

    def close(self):
        for v in self.parts:
            v.close()


def bench_env_steps(n_envs, n_steps, seed=1, threads=1):
    chk = C.c_uint64(0)
    secs = lib().orc_bench_env_steps(n_envs, n_steps, seed, threads, C.byref(chk))
    return secs


def bench_sample(n_envs, capacity, batch, n_batches, seed=1):
    chk = C.c_uint64(0)
    return lib().orc_bench_sample(n_envs, capacity, batch, n_batches, seed, C.byref(chk))


def bench_actor_loop(n_steps, capacity=4096, batch=32, seed=1):
    """seconds for n_steps env-steps of the reference's learn_episode data path (one env, one thread, no model)"""
    chk = C.c_uint64(0)
    return lib().orc_bench_actor_loop(n_steps, capacity, batch, seed, C.byref(chk))
