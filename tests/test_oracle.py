"""CPU tests of the oracle (test infrastructure): the reference's own known-answer vectors, the committed golden
traces, and the semantics of the restated frame ring / replay FIFO / sampler."""
import hashlib
import json
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
GOLD = json.load(open(os.path.join(HERE, "golden", "oracle_trace_v1.json")))
STATE_KEYS = ("ball_cx", "ball_cy", "ball_dx", "ball_dy", "pad_min_x", "pad_max_x", "pad_speed", "bricks", "score", "episode_step")


def test_selftest_binary(O):
    """oracle/selftest.c: 13 rstest vectors (mechanics.rs:659-752), Random123 Philox KATs, brick layout, dir_x range."""
    exe = os.path.join(os.path.dirname(HERE), "oracle", "selftest")
    assert os.path.exists(exe)
    res = subprocess.run([exe], capture_output=True, text=True)
    assert res.returncode == 0, res.stdout


def test_reference_wall_vectors_exact(O):
    # mechanics.rs:659-693: assert_eq! on the whole ContactSurface
    assert O.collision_wall("left", (10.0, 10.0), 5.0, (-2.0, 2.0))[0] == 0
    assert O.collision_wall("left", (5.0, 10.0), 5.0, (-5.0, 0.0))[:5] == (1, 0.0, 0.0, 1.0, 0.0)
    assert O.collision_wall("left", (7.0, 7.0), 5.0, (-5.0, 0.0))[:5] == (1, 2.0, 0.0, 1.0, 0.0)
    assert O.collision_wall("right", (590.0, 10.0), 5.0, (2.0, 2.0))[0] == 0
    assert O.collision_wall("right", (595.0, 10.0), 5.0, (5.0, 0.0))[:5] == (1, 0.0, 0.0, -1.0, 0.0)
    assert O.collision_wall("right", (593.0, 7.0), 5.0, (5.0, 0.0))[:5] == (1, 2.0, 0.0, -1.0, 0.0)


@pytest.mark.parametrize("mv,rmin,rmax,exp", [
    ((10.0, 0.0), (150.0, 90.0), (170.0, 110.0), None),
    ((5.0, 0.0), (110.0, 90.0), (130.0, 110.0), (5.0, -1.0, 0.0)),
    ((3.0, -3.0), (100.0, 70.0), (120.0, 93.0), (2.83, 0.0, 1.0)),
    ((-8.0, -8.0), (70.0, 80.0), (90.0, 100.0), (7.07, 1.0, 0.0)),
    ((-1.46, -1.46), (80.0, 80.0), (95.0, 95.0), (2.07, 0.7071, 0.7071)),
    ((-5.0, -5.0), (80.0, 80.0), (95.0, 95.0), (2.07, 0.7071, 0.7071)),
    ((-4.2, -4.2), (80.0, 80.0), (90.0, 90.0), None),
])
def test_reference_rectangle_vectors(O, mv, rmin, rmax, exp):
    # mechanics.rs:708-752 with its tolerances
    some, way, approx, nx, ny, err = O.collision_rect((100.0, 100.0), 5.0, mv, rmin, rmax)
    assert bool(some) == (exp is not None)
    if exp:
        assert abs(nx - exp[1]) <= 0.01 and abs(ny - exp[2]) <= 0.01 and abs(way - exp[0]) <= 0.1 and 0.0 <= approx < 0.8


def _trace(O, n_envs, seed, steps, replay_capacity=0):
    env = O.VecEnv(n_envs, seed=seed, replay_capacity=replay_capacity)
    acts = O.synthetic_actions(seed, 0, n_envs, 0, steps)
    hs, hf, hr = hashlib.sha256(), hashlib.sha256(), hashlib.sha256()
    for t in range(steps):
        r, d = env.step(acts[t])
        hr.update(r.tobytes()); hr.update(d.tobytes())
        st = env.state()
        for k in STATE_KEYS:
            hs.update(st[k].tobytes())
        hf.update(env.obs_u8().tobytes())
    return env, hs.hexdigest(), hf.hexdigest(), hr.hexdigest()


def test_golden_single_env_trace(O):
    """BASELINE configs[0]: one env, fixed seed, random policy — first 2,500 of the 10k golden steps are re-hashed
    step by step here (the full 10k run is checked by the slow test below and by the GPU suite)."""
    g = GOLD["single_env_10k"]
    env = O.VecEnv(1, seed=g["seed"])
    acts = O.synthetic_actions(g["seed"], 0, 1, 0, 1000)
    for t in range(1000):
        env.step(acts[t])
        if str(t + 1) in g["checkpoints"]:
            st = env.state()
            for k in STATE_KEYS:
                bits = st[k].view(np.uint32 if st[k].dtype == np.float32 else st[k].dtype)
                assert [int(x) for x in bits] == g["checkpoints"][str(t + 1)][k], (t + 1, k)
    env.close()


@pytest.mark.slow
def test_golden_single_env_trace_full(O):
    g = GOLD["single_env_10k"]
    env, hs, hf, hr = _trace(O, 1, g["seed"], g["steps"])
    assert hs == g["sha256_state_bits_every_step"] and hf == g["sha256_obs_u8_every_step"] and hr == g["sha256_reward_done_every_step"]
    assert {k: float(v) for k, v in env.stats().items()} == g["stats"]


def test_golden_multi_env_replay(O):
    g = GOLD["multi_env_replay"]
    env, hs, hf, hr = _trace(O, g["n_envs"], g["seed"], g["steps"], g["replay_capacity"])
    assert hs == g["sha256_state_bits_every_step"] and hf == g["sha256_obs_u8_every_step"] and hr == g["sha256_reward_done_every_step"]
    r = g["replay"]
    assert env.replay_len() == r["len"]
    idx = O.sample_distinct(g["seed"], 0, r["len"], 32)
    assert [int(x) for x in idx] == r["indices_call0_batch32"]
    got = env.get_many(idx, "u8")
    assert hashlib.sha256(got["state"].tobytes()).hexdigest() == r["sha256_state_u8"]
    assert hashlib.sha256(got["state_next"].tobytes()).hexdigest() == r["sha256_next_u8"]
    assert [float(x) for x in got["reward"]] == r["reward"] and [int(x) for x in got["action"]] == r["action"] and [int(x) for x in got["done"]] == r["done"]
    assert hashlib.sha256(env.get_many(idx, "f32")["state"].tobytes()).hexdigest() == r["sha256_state_f32"]


def test_golden_small_vectors(O):
    assert float(O.lib().orc_dir_x_from_bits(0)) == GOLD["dir_x"]["bits_0"]
    assert float(O.lib().orc_reset_dir_x(9, 3, 5)) == GOLD["dir_x"]["env3_ep5_seed9"]
    assert [int(x) for x in O.sample_distinct(1, 0, 100, 50)] == GOLD["sample_distinct"]["seed1_call0_len100_b50"]
    assert [int(x) for x in O.sample_distinct(1, 5, 33, 32)] == GOLD["sample_distinct"]["seed1_call5_len33_b32"]


def test_sampler_properties(O):
    """The reference's own sampler test (self_driving_tf_q_learner.rs:346-361): 50 distinct ids in 0..100, x100."""
    for call in range(100):
        r = O.sample_distinct(12345, call, 100, 50)
        assert len(set(r.tolist())) == 50 and r.max() < 100
    with pytest.raises(ValueError):
        O.sample_distinct(1, 0, 10, 11)            # assert!(range.end - range.start >= BATCH_SIZE)
    full = O.sample_distinct(3, 0, 64, 64)         # len == batch: a permutation
    assert sorted(full.tolist()) == list(range(64))


def test_frame_ring_and_state_layout(O):
    """FrameRingBuffer (frame_ring_buffer.rs:17-63): zeroed at reset, slot k mod 4 written at step k; tensor layout
    [x][y][slot] with value = u8 as f32 (breakout_environment.rs:42-54)."""
    env = O.VecEnv(1, seed=5)
    assert not env.obs_u8().any()
    frames = []
    for k in range(6):
        env.step(np.array([k % 3], dtype=np.uint8))
        obs = env.obs_u8()[0]
        frames.append(obs[k % 4].copy())
        assert obs[k % 4].any()
        for j in range(max(0, k - 3), k + 1):
            assert np.array_equal(obs[j % 4], frames[j])
        if k < 3:
            assert not obs[k + 1:].any()           # slots not yet written stay zero
    f32 = env.obs_f32()[0]
    u8 = env.obs_u8()[0]
    assert f32.shape == (84, 84, 4)
    assert np.array_equal(f32, np.transpose(u8, (2, 1, 0)).astype(np.float32))
    assert set(np.unique(u8).tolist()) <= {0, 96, 236, 255}
    env.close()


def test_replay_fifo_semantics(O):
    """ReplayBuffer (replay_buffer.rs:21-29,85-98): FIFO with eviction, index 0 = oldest; state_next of a row is the
    state of the next row of the same env inside an episode; episode starts carry an all-zero state."""
    n, cap = 2, 10
    env = O.VecEnv(n, seed=3, replay_capacity=cap)
    acts = O.synthetic_actions(3, 0, n, 0, 12)
    for t in range(12):
        env.step(acts[t])
        assert env.replay_len() == min(cap, n * (t + 1))
    g = env.get_many(np.arange(cap), "u8")
    # after 12 steps x 2 envs = 24 inserts, rows 14..23 remain: (t, e) = (7,0), (7,1), ..., (11,1)
    assert [int(a) for a in g["action"]] == [int(acts[7 + i // 2, i % 2]) for i in range(cap)]
    for i in range(cap - n):
        assert np.array_equal(g["state_next"][i], g["state"][i + n])
    first = env.get_many([0], "u8") if False else None
    env.close()
    env = O.VecEnv(1, seed=3, replay_capacity=4)
    env.step(np.array([0], dtype=np.uint8))
    g = env.get_many([0], "u8")
    assert not g["state"].any() and g["state_next"][0, 0].any() and not g["state_next"][0, 1:].any()
    env.close()


def test_episode_window(O):
    """episode reward window (replay_buffer.rs:100-124)."""
    env = O.VecEnv(4, seed=8, replay_capacity=64, episode_window=5)
    acts = O.synthetic_actions(8, 0, 4, 0, 600)
    for t in range(600):
        env.step(acts[t])
    w = env.episode_rewards()
    assert len(w) == 5
    assert env.avg_episode_reward() == pytest.approx(float(np.float32(w.sum()) / 5))
    assert env.min_episode_reward() == w.min()
    env.close()


def test_sharded_vec_env_equals_one_vec_env(O):
    """oracle.ShardedVecEnv (the threaded full-size parity leg) is the single VecEnv bit for bit; the block action fill equals
    the one-by-one stream."""
    n, T, seed = 203, 260, 17
    acts = O.synthetic_actions(seed, 40, n, 3, T)
    L = O.lib()
    for t, e in ((0, 0), (5, 77), (T - 1, n - 1)):
        assert acts[t, e] == L.orc_synthetic_action(seed, 40 + e, 3 + t)
    one, many = O.VecEnv(n, seed=seed, env_id_base=40), O.ShardedVecEnv(n, seed=seed, env_id_base=40, parts=6)
    r2, d2 = many.run(acts)
    for t in range(T):
        r, d = one.step(acts[t])
        assert np.array_equal(r, r2[t]) and np.array_equal(d, d2[t])
    a, b = one.state(), many.state()
    for k in a:
        assert np.array_equal(a[k], b[k]), k
    assert np.array_equal(one.obs_u8(), many.obs_u8()) and d2.sum() > 0
    one.close(); many.close()


def test_contact_query_golden(O):
    """The restated parry2d contact query answers the committed input list (tests/golden/contact_inputs_v1.txt, the list a
    `cargo run -p trace-dumper -- --contacts` run of the real reference is compared against) exactly as recorded, and the
    answers make geometric sense: dist = distance(ball centre, box) - radius within float tolerance when the centre is outside,
    normals are opposite unit vectors."""
    import os
    import sys
    import numpy as np
    here = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    sys.path.insert(0, here)
    import make_contact_inputs as M
    cs = M.read_inputs(os.path.join(here, "contact_inputs_v1.txt"))
    lines = M.oracle_lines(cs)
    want = [l.strip() for l in open(os.path.join(here, "contact_oracle_v1.txt")) if l.strip()]
    assert len(cs) == len(want) > 1000 and lines == want
    n_some = 0
    for (cx, cy, r, x0, y0, x1, y1), line in zip(cs, lines):
        dx = max(x0 - cx, 0.0, cx - x1); dy = max(y0 - cy, 0.0, cy - y1)
        gap = float(np.hypot(np.float64(dx), np.float64(dy))) - float(r)
        t = line.split()
        if t[0] == "0":
            assert gap > 0.8 - 1e-4
            continue
        n_some += 1
        dist, n1x, n1y, n2x, n2y = [float(np.uint32(int(v, 16)).view(np.float32)) for v in t[1:]]
        assert dist <= 0.8 and abs(np.hypot(n1x, n1y) - 1.0) < 1e-5 and n1x == -n2x and n1y == -n2y
        if dx > 1e-3 or dy > 1e-3:                       # centre outside the box: the plain closest-point distance
            assert abs(dist - gap) < 1e-4
    assert n_some > 800
