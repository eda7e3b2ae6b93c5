"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on identical seeded inputs.
Bar: bit-exact for bricks / score / done / reward / episode counters / pixels / gathered batches / sampled
indices; ball and paddle f32 state is compared bit-exactly too (north_star allows 1e-6 relative; we hold 0)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

STATE_F32 = ("ball_cx", "ball_cy", "ball_dx", "ball_dy", "pad_min_x", "pad_max_x", "pad_speed")
STATE_INT = ("bricks", "score", "episode_step")


def _assert_state_equal(gs, os_, ctx="", nan_ok=False):
    for k in STATE_INT:
        assert np.array_equal(gs[k], os_[k]), "%s %s differs at envs %s" % (ctx, k, np.nonzero(gs[k] != os_[k])[0][:8])
    assert np.array_equal(gs["finished"] != 0, os_["finished"] != 0), ctx + " finished differs"
    for k in STATE_F32:
        a, b = gs[k], os_[k]
        if nan_ok:      # envs past a sticky error flag (the reference would have panicked) may hold NaNs: NaN on both sides is agreement
            both = np.isnan(a) & np.isnan(b)
            a, b = np.where(both, np.float32(0), a), np.where(both, np.float32(0), b)
        rel = np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
        assert np.all((a == b) | (rel <= 1e-6)), "%s %s beyond 1e-6 relative" % (ctx, k)   # north_star tolerance
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), "%s %s not bit-identical (max rel %.3g)" % (ctx, k, rel.max())


# ---- the reference's own known-answer vectors, replayed through the DEVICE collision code ----
def test_device_collision_known_answers(qlb):
    """mechanics.rs:659-693 (walls, exact) and :708-752 (rectangle, normal +-0.01, way +-0.1, 0 <= approx < 0.8)."""
    W = 600.0
    assert qlb.debug_collision_wall("left", (10.0, 10.0), 5.0, (-2.0, 2.0))[0] == 0
    assert qlb.debug_collision_wall("left", (5.0, 10.0), 5.0, (-5.0, 0.0))[:5] == (1, 0.0, 0.0, 1.0, 0.0)
    assert qlb.debug_collision_wall("left", (7.0, 7.0), 5.0, (-5.0, 0.0))[:5] == (1, 2.0, 0.0, 1.0, 0.0)
    assert qlb.debug_collision_wall("right", (W - 10.0, 10.0), 5.0, (2.0, 2.0))[0] == 0
    assert qlb.debug_collision_wall("right", (W - 5.0, 10.0), 5.0, (5.0, 0.0))[:5] == (1, 0.0, 0.0, -1.0, 0.0)
    assert qlb.debug_collision_wall("right", (W - 7.0, 7.0), 5.0, (5.0, 0.0))[:5] == (1, 2.0, 0.0, -1.0, 0.0)
    s = 0.5 ** 0.5
    cases = [
        ((10.0, 0.0), (150.0, 90.0), (170.0, 110.0), None),
        ((5.0, 0.0), (110.0, 90.0), (130.0, 110.0), (5.0, -1.0, 0.0)),
        ((3.0, -3.0), (100.0, 70.0), (120.0, 93.0), (2.83, 0.0, 1.0)),
        ((-8.0, -8.0), (70.0, 80.0), (90.0, 100.0), (7.07, 1.0, 0.0)),
        ((-1.46, -1.46), (80.0, 80.0), (95.0, 95.0), (2.07, s, s)),
        ((-5.0, -5.0), (80.0, 80.0), (95.0, 95.0), (2.07, s, s)),
        ((-4.2, -4.2), (80.0, 80.0), (90.0, 90.0), None),
    ]
    for mv, rmin, rmax, exp in cases:
        some, way, approx, nx, ny, err = qlb.debug_collision_rect((100.0, 100.0), 5.0, mv, rmin, rmax)
        assert bool(some) == (exp is not None), (mv, some)
        if exp:
            assert abs(nx - exp[1]) <= 0.01 and abs(ny - exp[2]) <= 0.01 and abs(way - exp[0]) <= 0.1
            assert 0.0 <= approx < 0.8


def test_device_collision_matches_oracle_bitwise(qlb, O):
    rng = np.random.default_rng(5)
    for _ in range(300):
        c = (float(np.float32(rng.uniform(20, 580))), float(np.float32(rng.uniform(20, 580))))
        ang = rng.uniform(0, 2 * np.pi)
        mv = (float(np.float32(4 * np.cos(ang))), float(np.float32(4 * np.sin(ang))))
        off = rng.uniform(-16, 16, size=2)
        rmin = (float(np.float32(c[0] + off[0])), float(np.float32(c[1] + off[1])))
        rmax = (float(np.float32(rmin[0] + 25)), float(np.float32(rmin[1] + 25)))
        g = qlb.debug_collision_rect(c, 10.0, mv, rmin, rmax)
        o = O.collision_rect(c, 10.0, mv, rmin, rmax)
        assert g[0] == o[0]
        if g[0]:
            assert np.array_equal(np.float32(g[1:5]).view(np.uint32), np.float32(o[1:5]).view(np.uint32)), (c, mv, rmin, g, o)
        assert g[5] == o[5]


# ---- trajectories ----
@pytest.mark.parametrize("n_envs,chunks,seed", [(256, [1, 3, 64, 200, 500, 32], 11), (37, [5, 1, 1, 90, 300], 3), (1, [400, 400], 9)])
def test_trajectory_parity(qlb, O, n_envs, chunks, seed):
    """step+render+frame-stack+auto-reset for N envs, several launch sizes, vs the oracle step by step."""
    env = qlb.BreakoutEnvironment(n_envs=n_envs, seed=seed, replay_capacity=0)
    ora = O.VecEnv(n_envs, seed=seed)
    _assert_state_equal(env.read_state(), ora.state(), "initial")
    t = 0
    for k in chunks:
        acts = O.synthetic_actions(seed, 0, n_envs, t, k)
        reward, done = env.step_many(acts)
        for s in range(k):
            r, d = ora.step(acts[s])
            assert np.array_equal(r, reward[s]), "reward differs at step %d" % (t + s)
            assert np.array_equal(d, done[s]), "done differs at step %d" % (t + s)
        t += k
        _assert_state_equal(env.read_state(), ora.state(), "t=%d" % t)
        assert np.array_equal(env.obs(qlb.LAYOUT_U8_BHYX), ora.obs_u8()), "frame stacks differ at t=%d" % t
    assert np.array_equal(env.obs(qlb.LAYOUT_F32_BXYH), ora.obs_f32())
    gs, os_ = env.stats(), ora.stats()
    assert gs["episodes"] == os_["episodes"] and gs["sum_return"] == int(os_["sum_return"]) and gs["steps"] == os_["steps"]
    if gs["episodes"]:
        assert gs["min_return"] == int(os_["min_return"]) and gs["max_return"] == int(os_["max_return"])
    assert env.error_flags() == 0 and not ora.state()["err"].any()
    env.close()


@pytest.mark.parametrize("cfg,epc,chunk", [(1, 0, -1), (2, 0, -1), (3, 0, -1), (4, 0, -1), (5, 0, -1), (6, 0, -1), (1, 28, 0), (1, 5, 4), (2, 13, 1),
                                           (6, 31, 3), (5, 1, 7), (1, 3, 2), (2, 2, 5), (5, 4, 4)])
def test_all_kernel_shapes_agree(qlb, O, cfg, epc, chunk, monkeypatch):
    """Every instantiated CTA shape of the fused kernel (render warps x resident frames per warp), any envs-per-batch
    split and any time-chunk length (env state handed from CTA to CTA through HBM) gives the same bits."""
    monkeypatch.setenv("QLC_ADVANCE_CFG", str(cfg))
    monkeypatch.setenv("QLC_EPC", str(epc))
    if chunk >= 0:
        monkeypatch.setenv("QLC_CHUNK", str(chunk))
    n, seed = 77, 31
    env = qlb.BreakoutEnvironment(n_envs=n, seed=seed, replay_capacity=n * 8)
    ora = O.VecEnv(n, seed=seed)
    t = 0
    for k in (1, 1, 2, 37, 150):
        acts = O.synthetic_actions(seed, 0, n, t, k)
        reward, done = env.step_many(acts)
        for s in range(k):
            r, d = ora.step(acts[s])
            assert np.array_equal(r, reward[s]) and np.array_equal(d, done[s])
        t += k
        assert np.array_equal(env.obs(qlb.LAYOUT_U8_BHYX), ora.obs_u8()), "frame stacks differ at t=%d (cfg %d)" % (t, cfg)
    _assert_state_equal(env.read_state(), ora.state(), "cfg %d" % cfg)
    env.close()


def test_skilled_play_hits_many_bricks(qlb, O):
    """A paddle that tracks the ball keeps episodes alive for thousands of steps: bounces off paddle, walls and
    bricks (multi-contact, bisection paths) must stay bit-identical."""
    n = 64
    seed = 21
    env = qlb.BreakoutEnvironment(n_envs=n, seed=seed)
    ora = O.VecEnv(n, seed=seed)
    rng = np.random.default_rng(0)
    for it in range(1500):
        st = ora.state()
        centre = (st["pad_min_x"] + st["pad_max_x"]) / 2
        target = st["ball_cx"] + rng.uniform(-25, 25, size=n).astype(np.float32)
        a = np.where(target < centre - 4, 1, np.where(target > centre + 4, 2, 0)).astype(np.uint8)
        r, d = ora.step(a)
        _, gr, gd = env.step(a)
        assert np.array_equal(r, gr) and np.array_equal(d, gd), "step %d" % it
        if it % 100 == 99:
            _assert_state_equal(env.read_state(), ora.state(), "it=%d" % it)
    _assert_state_equal(env.read_state(), ora.state(), "final")
    assert np.array_equal(env.obs(), ora.obs_u8())
    assert ora.state()["score"].max() >= 3 or ora.stats()["max_return"] >= 3, "controller never hit bricks; test too weak"
    assert env.error_flags() & ~qlb.ENVERR_DEGENERATE == 0
    env.close()


def test_explicit_reset_and_no_auto_reset(qlb, O):
    """Reference behaviour: no auto reset, the caller resets (learn_episode :142); dir_x is an explicit input."""
    n = 8
    env = qlb.BreakoutEnvironment(n_envs=n, seed=1, auto_reset=False)
    dirs = np.linspace(-0.35, -0.15, n, endpoint=False).astype(np.float32)
    env.reset(dir_x=dirs)
    st = env.read_state()
    assert np.array_equal(st["ball_dx"], dirs) and np.all(st["ball_dy"] == -1.0) and np.all(st["episode_step"] == 0)
    assert np.all(st["bricks"] == (1 << 60) - 1) and np.all(st["pad_min_x"] == 270) and np.all(st["pad_max_x"] == 330)
    # oracle: single envs with the same explicit directions, stepped past the end of the episode
    ora = O.VecEnv(n, seed=1, max_episode_steps=0)
    # the oracle driver auto-resets, so compare only until the first done
    for e in range(n):
        ora.reset_env(e, float(dirs[e]))
    acts = O.synthetic_actions(1, 0, n, 0, 400)
    reward, done = env.step_many(acts)
    alive = np.ones(n, dtype=bool)
    for s in range(400):
        r, d = ora.step(acts[s])
        assert np.array_equal(r[alive], reward[s][alive]) and np.array_equal(d[alive], done[s][alive])
        alive &= d == 0
        # once finished, the GPU env stays finished (sticky) and keeps reporting done
        assert np.all(done[s][~alive] == 1)
    assert not alive.all(), "no episode ended in 400 steps"
    # masked reset restarts only the selected envs
    mask = np.zeros(n, dtype=np.uint8); mask[::2] = 1
    before = env.read_state()
    env.reset(mask=mask)
    after = env.read_state()
    assert np.all(after["episode_step"][::2] == 0) and np.all(after["finished"][::2] == 0)
    assert np.array_equal(after["ball_cx"][1::2], before["ball_cx"][1::2])
    assert np.array_equal(after["episode"][::2], before["episode"][::2] + 1)
    for e in range(0, n, 2):   # Philox reset direction of (env, episode)
        assert after["ball_dx"][e] == np.float32(O.lib().orc_reset_dir_x(1, e, int(after["episode"][e])))
    env.close()


def test_truncation(qlb, O):
    n = 16
    env = qlb.BreakoutEnvironment(n_envs=n, seed=4, max_episode_steps=50)
    ora = O.VecEnv(n, seed=4, max_episode_steps=50)
    acts = O.synthetic_actions(4, 0, n, 0, 260)
    reward, done = env.step_many(acts)
    for s in range(260):
        r, d = ora.step(acts[s])
        assert np.array_equal(r, reward[s]) and np.array_equal(d, done[s])
    _assert_state_equal(env.read_state(), ora.state())
    assert env.read_state()["episode_step"].max() < 50
    assert env.stats()["episodes"] == ora.stats()["episodes"] >= n * 5
    env.close()


def test_invalid_action_is_an_error(qlb):
    env = qlb.BreakoutEnvironment(n_envs=4)
    with pytest.raises(qlb.QlError) as ei:
        env.step(np.array([0, 1, 3, 2], dtype=np.uint8))
    assert ei.value.code == qlb.ERR_OUT_OF_RANGE and "out of range" in str(ei.value)
    with pytest.raises(qlb.QlError):
        qlb.BreakoutAction.try_from_numeric(3)
    env.close()


# ---- replay ----
@pytest.mark.parametrize("n_envs,capacity,steps", [(1, 300, 700), (8, 8 * 40, 333), (64, 64 * 16, 100)])
def test_replay_parity(qlb, O, n_envs, capacity, steps):
    """FIFO semantics (index 0 = oldest, eviction when full), get_many + batch_to_multi_dim_array in both layouts."""
    seed = 17
    env = qlb.BreakoutEnvironment(n_envs=n_envs, seed=seed, replay_capacity=capacity)
    rb = qlb.ReplayBuffer(env)
    ora = O.VecEnv(n_envs, seed=seed, replay_capacity=capacity)
    rng = np.random.default_rng(1)
    t = 0
    for chunk in (1, 2, 5, 17, steps):
        acts = O.synthetic_actions(seed, 0, n_envs, t, chunk)
        env.step_many(acts)
        for s in range(chunk):
            ora.step(acts[s])
        t += chunk
        assert rb.len() == ora.replay_len()
        ln = rb.len()
        idx = rng.integers(0, ln, size=min(48, ln)).astype(np.uint32)
        idx[0] = 0; idx[-1] = ln - 1
        for layout, name in ((qlb.LAYOUT_U8_BHYX, "u8"), (qlb.LAYOUT_F32_BXYH, "f32")):
            g = rb.get_many(idx, layout)
            o = ora.get_many(idx, name)
            assert np.array_equal(g.reward, o["reward"]) and np.array_equal(g.action, o["action"]) and np.array_equal(g.done, o["done"])
            assert np.array_equal(g.state, o["state"]), "state differs (%s) at t=%d" % (name, t)
            assert np.array_equal(g.state_next, o["state_next"]), "state_next differs (%s) at t=%d" % (name, t)
    assert rb.capacity() == capacity
    assert np.array_equal(rb.actions(), ora.action_histogram())
    with pytest.raises(qlb.QlError) as ei:
        rb.get_many(np.array([rb.len()], dtype=np.uint32))
    assert ei.value.code == qlb.ERR_OUT_OF_RANGE
    env.close()



def test_host_gather_into_page_locked_buffers(qlb, O):
    """get_many(reuse=True): the stacks are copied straight into page-locked host arrays owned by the ReplayBuffer — same bytes as
    the pageable path and as the oracle, both layouts, changing batch sizes."""
    n, seed, steps = 48, 19, 70
    env = qlb.BreakoutEnvironment(n_envs=n, seed=seed, replay_capacity=n * 32)
    rb = qlb.ReplayBuffer(env)
    ora = O.VecEnv(n, seed=seed, replay_capacity=n * 32)
    acts = O.synthetic_actions(seed, 0, n, 0, steps)
    env.step_many(acts)
    for t in range(steps):
        ora.step(acts[t])
    for batch, call in ((32, 0), (7, 1), (32, 2), (200, 3)):
        idx = rb.generate_distinct_random_ids(batch, call)
        for layout, name in ((qlb.LAYOUT_U8_BHYX, "u8"), (qlb.LAYOUT_F32_BXYH, "f32")):
            g = rb.get_many(idx, layout, reuse=True)
            h = rb.get_many(idx, layout)
            o = ora.get_many(idx, name)
            assert np.array_equal(g.state, o["state"]) and np.array_equal(g.state_next, o["state_next"])
            assert np.array_equal(g.state, h.state) and np.array_equal(g.state_next, h.state_next)
            assert np.array_equal(g.reward, o["reward"]) and np.array_equal(g.action, o["action"]) and np.array_equal(g.done, o["done"])
    env.close(); ora.close()


@pytest.mark.parametrize("length_steps,batch", [(40, 32), (5, 32), (200, 512), (33, 1024), (5, 129), (5, 160), (9, 128)])
def test_sampler_parity(qlb, O, length_steps, batch):
    """generate_distinct_random_ids: distinct, in range (the reference's own property test :346-361) and equal to
    the oracle's sequential rejection on the same Philox stream — including ranges barely larger than the batch."""
    n_envs = 32
    env = qlb.BreakoutEnvironment(n_envs=n_envs, seed=99, replay_capacity=n_envs * 256)
    rb = qlb.ReplayBuffer(env)
    env.step_many(np.zeros((length_steps, n_envs), dtype=np.uint8))
    ln = rb.len()
    assert ln == length_steps * n_envs
    if ln < batch:
        with pytest.raises(qlb.QlError) as ei:
            rb.generate_distinct_random_ids(batch, 0)
        assert ei.value.code == qlb.ERR_NOT_ENOUGH
        env.close()
        return
    for call in (0, 1, 2, 1 << 33):
        g = rb.generate_distinct_random_ids(batch, call)
        assert len(set(g.tolist())) == batch and g.max() < ln
        assert np.array_equal(g, O.sample_distinct(99, call, ln, batch))
    env.close()


def test_sampler_many_batches_device(qlb, O):
    torch = pytest.importorskip("torch")
    n_envs, batch, nb = 64, 32, 100
    env = qlb.BreakoutEnvironment(n_envs=n_envs, seed=5, replay_capacity=n_envs * 64)
    rb = qlb.ReplayBuffer(env)
    env.step_many(np.zeros((2, n_envs), dtype=np.uint8))      # len = 128: collisions are frequent
    idx = torch.empty((nb, batch), dtype=torch.int32, device="cuda")
    rb.sample_device(batch, nb, 7, idx.data_ptr())
    torch.cuda.synchronize()
    got = idx.cpu().numpy().view(np.uint32)
    for i in range(nb):
        assert np.array_equal(got[i], O.sample_distinct(5, 7 + i, rb.len(), batch))
    env.close()


def test_full_size_properties(qlb, O):
    """BASELINE configs[1] size (4,096 envs) with a 64-step replay: size-independent checks — a spot-checked subset
    of envs against the oracle, frame-stack consistency between the env observation and the replay gather,
    pixel value set, statistics identities."""
    n, seed, T = 4096, 123, 96
    env = qlb.BreakoutEnvironment(n_envs=n, seed=seed, replay_capacity=n * 64)
    rb = qlb.ReplayBuffer(env)
    sub = np.arange(0, n, 97)
    acts = np.empty((T, n), dtype=np.uint8)
    lib = O.lib()
    # actions: oracle stream on the spot-checked envs, cheap numpy stream elsewhere
    acts[:] = np.random.default_rng(0).integers(0, 3, size=(T, n), dtype=np.uint8)
    for e in sub:
        for t in range(T):
            acts[t, e] = lib.orc_synthetic_action(seed, int(e), t)
    reward, done = env.step_many(acts)
    # spot check vs oracle: env e of the GPU run == oracle env with global id e
    for e in sub[:12]:
        o = O.VecEnv(1, seed=seed, env_id_base=int(e))
        for t in range(T):
            r, d = o.step(acts[t, e:e + 1])
            assert r[0] == reward[t, e] and d[0] == done[t, e]
        so = o.state(); sg = env.read_state()
        for k in STATE_F32 + STATE_INT:
            assert so[k][0] == sg[k][e], (k, e)
    obs = env.obs()
    assert set(np.unique(obs).tolist()) <= {0, 96, 236, 255}
    # the newest transition of every env: its state_next stack must equal the env's current observation unless the
    # episode ended on that step (then the env has been reset and shows an empty stack)
    ln = rb.len()
    assert ln == n * 64
    newest = np.arange(ln - n, ln, dtype=np.uint32)
    g = rb.get_many(newest[:512], qlb.LAYOUT_U8_BHYX)
    ended = (g.done != 0)
    assert np.array_equal(g.state_next[~ended], obs[:512][~ended])
    assert not obs[:512][ended].any()
    assert np.array_equal(g.reward, reward[-1, :512]) and np.array_equal(g.done, done[-1, :512]) and np.array_equal(g.action, acts[-1, :512])
    st = env.stats()
    assert st["steps"] == n * T and st["episodes"] == int(done.sum())
    assert st["sum_return"] + int(env.read_state()["score"].sum()) == int(reward.sum())
    assert env.error_flags() == 0
    env.close()


# ---- committed golden fixtures (tests/golden/oracle_trace_v1.json) straight against the CUDA path ----
import hashlib
import json
import os

_GOLD = json.load(open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_trace_v1.json")))
_GKEYS = ("ball_cx", "ball_cy", "ball_dx", "ball_dy", "pad_min_x", "pad_max_x", "pad_speed", "bricks", "score", "episode_step")


def test_golden_single_env_10k(qlb, O):
    """BASELINE configs[0]: single env, fixed seed, random policy, 10k steps — per-step state bits, frame stacks and
    reward/done hashed exactly like tests/golden/make_golden.py did with the oracle."""
    g = _GOLD["single_env_10k"]
    env = qlb.BreakoutEnvironment(n_envs=1, seed=g["seed"])
    acts = O.synthetic_actions(g["seed"], 0, 1, 0, g["steps"])
    hs, hf, hr = hashlib.sha256(), hashlib.sha256(), hashlib.sha256()
    for t in range(g["steps"]):
        _, r, d = env.step(acts[t])
        hr.update(r.tobytes()); hr.update(d.tobytes())
        st = env.read_state()
        for k in _GKEYS:
            hs.update(st[k].tobytes())
        hf.update(env.obs(qlb.LAYOUT_U8_BHYX).tobytes())
    assert hr.hexdigest() == g["sha256_reward_done_every_step"]
    assert hs.hexdigest() == g["sha256_state_bits_every_step"]
    assert hf.hexdigest() == g["sha256_obs_u8_every_step"]
    s = env.stats()
    assert float(s["episodes"]) == g["stats"]["episodes"] and float(s["sum_return"]) == g["stats"]["sum_return"]
    env.close()


def test_golden_multi_env_replay(qlb, O):
    g = _GOLD["multi_env_replay"]
    env = qlb.BreakoutEnvironment(n_envs=g["n_envs"], seed=g["seed"], replay_capacity=g["replay_capacity"])
    rb = qlb.ReplayBuffer(env)
    acts = O.synthetic_actions(g["seed"], 0, g["n_envs"], 0, g["steps"])
    env.step_many(acts)
    r = g["replay"]
    assert rb.len() == r["len"]
    idx = rb.generate_distinct_random_ids(32, 0)
    assert [int(x) for x in idx] == r["indices_call0_batch32"]
    got = rb.get_many(idx, qlb.LAYOUT_U8_BHYX)
    assert hashlib.sha256(got.state.tobytes()).hexdigest() == r["sha256_state_u8"]
    assert hashlib.sha256(got.state_next.tobytes()).hexdigest() == r["sha256_next_u8"]
    assert [float(x) for x in got.reward] == r["reward"] and [int(x) for x in got.action] == r["action"] and [int(x) for x in got.done] == r["done"]
    assert hashlib.sha256(rb.get_many(idx, qlb.LAYOUT_F32_BXYH).state.tobytes()).hexdigest() == r["sha256_state_f32"]
    st = env.read_state()
    for k in _GKEYS:
        bits = st[k].view(np.uint32 if st[k].dtype == np.float32 else st[k].dtype)
        assert [int(x) for x in bits] == g["checkpoints"][str(g["steps"])][k], k
    env.close()


def test_golden_tracking_policy(qlb):
    """Skilled play (returns up to 20+ per episode) against the golden hashes — no oracle involved at run time: the
    policy reads the GPU state, so any divergence changes the action stream hash as well."""
    g = _GOLD["tracking_policy"]
    n = g["n_envs"]
    env = qlb.BreakoutEnvironment(n_envs=n, seed=g["seed"])
    hs, ha, hf = hashlib.sha256(), hashlib.sha256(), hashlib.sha256()
    st = env.read_state()
    for t in range(g["steps"]):
        centre = (st["pad_min_x"] + st["pad_max_x"]) / 2
        off = (((np.arange(n) * 37 + t * 11) % 51) - 25).astype(np.float32)
        target = st["ball_cx"] + off
        a = np.where(target < centre - 4, 1, np.where(target > centre + 4, 2, 0)).astype(np.uint8)
        ha.update(a.tobytes())
        env.step(a)
        st = env.read_state()
        for k in _GKEYS:
            hs.update(st[k].tobytes())
        if t % 16 == 15:
            hf.update(env.obs(qlb.LAYOUT_U8_BHYX).tobytes())
    assert ha.hexdigest() == g["sha256_actions"]
    assert hs.hexdigest() == g["sha256_state_bits_every_step"]
    assert hf.hexdigest() == g["sha256_obs_u8_every_16_steps"]
    assert [int(x) for x in st["score"]] == g["final_score"] and [int(x) for x in st["bricks"]] == g["final_bricks"]
    s = env.stats()
    assert float(s["episodes"]) == g["stats"]["episodes"] and float(s["sum_return"]) == g["stats"]["sum_return"]
    assert float(s["min_return"]) == g["stats"]["min_return"] and float(s["max_return"]) == g["stats"]["max_return"]
    env.close()


def test_rare_events_parity(qlb, O):
    """~3e8 env-steps of random play on the GPU; every env whose sticky error flags fired (situations in which the
    reference would panic or recurse without bound: mechanics.rs:265,284,303,361-389) is replayed from t = 0 on the
    oracle with the same action history and must agree bit for bit, flags included. Also replays unflagged envs."""
    n, chunk, n_chunks, seed = 16384, 500, 40, 4242
    env = qlb.BreakoutEnvironment(n_envs=n, seed=seed)
    rng = np.random.default_rng(99)
    history = []
    for _ in range(n_chunks):
        a = rng.integers(0, 3, size=(chunk, n), dtype=np.uint8)
        # a sideways-sweeping paddle makes the rare paddle/ball side hits much more likely
        a[:, : n // 2] = np.where(rng.random((chunk, n // 2)) < 0.85, (np.arange(chunk)[:, None] // 40 % 2 + 1), a[:, : n // 2]).astype(np.uint8)
        history.append(a)
        env.step_many(a)
    st = env.read_state()
    flagged = np.nonzero(st["err"])[0]
    print("flagged envs: %d of %d after %d steps; flags OR = %d" % (flagged.size, n, chunk * n_chunks, int(np.bitwise_or.reduce(st["err"]))))
    pick = list(flagged[:6]) + [0, 1, n // 2, n - 1]
    acts = np.concatenate(history, axis=0)
    for e in pick:
        o = O.VecEnv(1, seed=seed, env_id_base=int(e))
        for t in range(acts.shape[0]):
            o.step(acts[t, e:e + 1])
        so = o.state()
        for k in STATE_F32 + STATE_INT + ("err",):
            assert so[k][0] == st[k][e] or (so[k][0] != so[k][0] and st[k][e] != st[k][e]), (k, int(e), so[k][0], st[k][e])
        o.close()
    env.close()


def test_replay_config_1m_transitions(qlb, O):
    """BASELINE configs[2] size: 4,096 envs, 1,048,576-transition replay (7.5 GB frame ring), filled past wrap-around;
    uniform samples of 32 and 512 with frame-stack gather. Size-independent checks: host-recorded (action, reward, done)
    history per sampled row, s'(t) == s(t+1) chaining across rows of the same env, and exact equality with the oracle's
    FIFO replay for whole envs replayed on the CPU."""
    n, cap, seed = 4096, 1 << 20, 77
    t_cap = cap // n                       # 256 time steps
    T = t_cap + 45                         # wrap the ring
    env = qlb.BreakoutEnvironment(n_envs=n, seed=seed, replay_capacity=cap)
    rb = qlb.ReplayBuffer(env)
    rng = np.random.default_rng(3)
    acts = rng.integers(0, 3, size=(T, n), dtype=np.uint8)
    reward = np.empty((T, n), dtype=np.float32); done = np.empty((T, n), dtype=np.uint8)
    for t0 in range(0, T, 43):
        k = min(43, T - t0)
        r, d = env.step_many(acts[t0:t0 + k])
        reward[t0:t0 + k] = r; done[t0:t0 + k] = d
    assert rb.len() == cap == rb.capacity()
    t_old = T - t_cap
    for batch in (32, 512):
        idx = rb.generate_distinct_random_ids(batch, 5)
        assert len(set(idx.tolist())) == batch and idx.max() < cap
        assert np.array_equal(idx, O.sample_distinct(seed, 5, cap, batch))
        g = rb.get_many(idx, qlb.LAYOUT_U8_BHYX)
        tt, ee = t_old + idx // n, idx % n
        assert np.array_equal(g.action, acts[tt, ee]) and np.array_equal(g.reward, reward[tt, ee]) and np.array_equal(g.done, done[tt, ee])
        # chaining: row (t, e) and row (t+1, e)
        nxt = idx + n
        ok = (nxt < cap) & (g.done == 0)
        g2 = rb.get_many(nxt[ok], qlb.LAYOUT_U8_BHYX)
        assert np.array_equal(g.state_next[ok], g2.state)
        f = rb.get_many(idx[:16], qlb.LAYOUT_F32_BXYH)
        assert np.array_equal(f.state, np.transpose(g.state[:16], (0, 3, 2, 1)).astype(np.float32))
        assert np.array_equal(f.state_next, np.transpose(g.state_next[:16], (0, 3, 2, 1)).astype(np.float32))
    # EVERY row of a sampled minibatch of 512 against the oracle: the env of each sampled row is replayed on the CPU with a
    # FIFO of the same time window (capacity t_cap for a single env), its row j = idx // n gathered there
    idx = rb.generate_distinct_random_ids(512, 9)
    g = rb.get_many(idx, qlb.LAYOUT_U8_BHYX)
    for e in np.unique(idx % n):
        rows = np.nonzero(idx % n == e)[0]
        o = O.VecEnv(1, seed=seed, env_id_base=int(e), replay_capacity=t_cap)
        for t in range(T):
            o.step(acts[t, e:e + 1])
        og = o.get_many((idx[rows] // n).astype(np.uint32), "u8")
        assert np.array_equal(g.state[rows], og["state"]) and np.array_equal(g.state_next[rows], og["state_next"]), "env %d" % e
        assert np.array_equal(g.reward[rows], og["reward"]) and np.array_equal(g.action[rows], og["action"]) and np.array_equal(g.done[rows], og["done"])
        o.close()
    # whole envs against the oracle FIFO
    for e in (0, 1777, n - 1):
        o = O.VecEnv(1, seed=seed, env_id_base=e, replay_capacity=t_cap)
        for t in range(T):
            o.step(acts[t, e:e + 1])
        j = np.array([0, 1, 2, 3, 4, 5, t_cap // 2, t_cap - 2, t_cap - 1], dtype=np.uint32)
        og = o.get_many(j, "u8")
        gg = rb.get_many(j * n + e, qlb.LAYOUT_U8_BHYX)
        assert np.array_equal(gg.state, og["state"]) and np.array_equal(gg.state_next, og["state_next"])
        assert np.array_equal(gg.reward, og["reward"]) and np.array_equal(gg.action, og["action"]) and np.array_equal(gg.done, og["done"])
        o.close()
    env.close()


def test_device_sweep_fuzz_vs_oracle(qlb, O):
    """400k random and adversarial ball-vs-box sweeps through the DEVICE routine (sweep_ball_box: contact query, back-off
    estimate, bisection, acos-free acceptance test) against the oracle's collision_check_with_rectangle: the Some/None
    decision, way, approximation, normal and flags must agree bit for bit."""
    rng = np.random.default_rng(2026)
    n = 400_000
    cases = np.empty((n, 9), dtype=np.float32)
    # game-like: radius 10, |mv| about 4 (also after partial moves), bricks 25x25 / paddle 60x10 near the ball
    c = rng.uniform(20, 580, size=(n, 2))
    ang = rng.uniform(0, 2 * np.pi, size=n)
    ln = np.where(rng.random(n) < 0.7, 4.0, rng.uniform(0.002, 4.0, size=n))
    w = np.where(rng.random(n) < 0.8, 25.0, 60.0); h = np.where(w == 25.0, 25.0, 10.0)
    off = rng.uniform(-18, 18, size=(n, 2))
    cases[:, 0:2] = c; cases[:, 2] = 10.0
    cases[:, 3] = ln * np.cos(ang); cases[:, 4] = ln * np.sin(ang)
    # box placed around the ball so that touching / grazing / penetrating configurations are frequent
    cases[:, 5] = c[:, 0] + off[:, 0] - np.where(off[:, 0] < 0, w, 0); cases[:, 6] = c[:, 1] + off[:, 1] - np.where(off[:, 1] < 0, h, 0)
    cases[:, 7] = cases[:, 5] + w; cases[:, 8] = cases[:, 6] + h
    # adversarial quarter: axis-aligned motion, exact touching distances, corners on the diagonal, integer coordinates
    k = n // 4
    cases[:k, 0:2] = np.round(cases[:k, 0:2]); cases[:k, 5:9] = np.round(cases[:k, 5:9])
    cases[: k // 2, 3] = np.where(rng.random(k // 2) < 0.5, 0.0, cases[: k // 2, 3]); cases[k // 2: k, 3] = cases[k // 2: k, 4]
    g_some, g_surf, g_err = qlb.debug_collision_rect_batch(cases)
    o_some, o_surf, o_err = O.collision_rect_batch(cases)
    assert np.array_equal(g_some, o_some), "Some/None differs in %d cases" % int((g_some != o_some).sum())
    hit = g_some != 0
    assert 0.02 < hit.mean() < 0.9, "fuzz distribution degenerate: %.3f hits" % hit.mean()
    assert np.array_equal(g_surf[hit].view(np.uint32), o_surf[hit].view(np.uint32)), "surface bits differ"
    assert np.array_equal(g_err, o_err)
    print("fuzz: %d cases, %.1f%% contacts, %d with flags" % (n, 100 * hit.mean(), int((g_err != 0).sum())))


def test_checkpoint_resume_is_bit_identical(qlb, O, tmp_path):
    """save the env shard + replay ring, resume in a fresh env, continue — identical to never stopping."""
    n, seed, cap = 160, 66, 160 * 24
    acts = O.synthetic_actions(seed, 0, n, 0, 90)
    a = qlb.BreakoutEnvironment(n_envs=n, seed=seed, replay_capacity=cap, max_episode_steps=200)
    ra = qlb.ReplayBuffer(a)
    a.step_many(acts[:60])
    ra.add_episode_reward(3.0); ra.add_episode_reward(7.0)
    path = str(tmp_path / "shard.qlc")
    a.save(path)
    a.step_many(acts[60:])
    b = qlb.BreakoutEnvironment(n_envs=n, seed=seed, replay_capacity=cap, max_episode_steps=200)
    rb = qlb.ReplayBuffer(b)
    b.load(path)
    assert b.time() == 60 and rb.len() == ra.len() and rb.episode_rewards().tolist() == [3.0, 7.0]
    b.step_many(acts[60:])
    sa, sb = a.read_state(), b.read_state()
    for k in sa:
        assert np.array_equal(sa[k].view(np.uint8), sb[k].view(np.uint8)), k
    assert np.array_equal(a.obs(), b.obs()) and a.stats() == b.stats()
    idx = ra.generate_distinct_random_ids(64, 3)
    assert np.array_equal(idx, rb.generate_distinct_random_ids(64, 3))
    ga, gb = ra.get_many(idx, qlb.LAYOUT_U8_BHYX), rb.get_many(idx, qlb.LAYOUT_U8_BHYX)
    assert np.array_equal(ga.state, gb.state) and np.array_equal(ga.state_next, gb.state_next) and np.array_equal(ga.reward, gb.reward)
    # a differently configured env refuses the file
    c = qlb.BreakoutEnvironment(n_envs=n, seed=seed + 1, replay_capacity=cap, max_episode_steps=200)
    with pytest.raises(qlb.QlError):
        c.load(path)
    for e in (a, b, c):
        e.close()


def test_shard_of_65536_envs(qlb, O):
    """BASELINE configs[3] shard size: 65,536 envs on one GPU (multi-wave launches, a single-step launch in between), EVERY env
    against the oracle (threaded sub-shards with the same global env ids): reward / done of all 3.1e6 env-steps, the complete
    final state and the 4-frame stacks of all envs; then statistics identities and replay chaining on the 1 M-transition ring."""
    n, seed, k, base = 65536, 909, 48, 1_000_000
    env = qlb.BreakoutEnvironment(n_envs=n, seed=seed, env_id_base=base, replay_capacity=n * 16)
    rb = qlb.ReplayBuffer(env)
    ora = O.ShardedVecEnv(n, seed=seed, env_id_base=base)
    rng = np.random.default_rng(12)
    acts = rng.integers(0, 3, size=(k, n), dtype=np.uint8)
    reward, done = env.step_many(acts[:16])
    r2, d2 = env.step_many(acts[16:17])          # a single-step launch in between
    r3, d3 = env.step_many(acts[17:])
    reward = np.concatenate([reward, r2, r3]); done = np.concatenate([done, d2, d3])
    ro, do = ora.run(acts)
    bad = np.nonzero((reward != ro) | (done != do))
    assert bad[0].size == 0, "reward/done differ first at step %d env %d" % (bad[0][0], bad[1][0])
    st = env.read_state()
    _assert_state_equal(st, ora.state(), "65536 envs", nan_ok=True)
    obs = env.obs(qlb.LAYOUT_U8_BHYX)
    for part, off, size in zip(ora.parts, ora.offsets, ora.sizes):           # part by part: the stacks are 1.85 GB per side
        assert np.array_equal(obs[int(off):int(off) + int(size)], part.obs_u8()), "frame stacks differ in envs %d.." % off
    del obs
    ora.close()
    s = env.stats()
    assert s["steps"] == n * k and s["episodes"] == int(done.sum())
    assert s["sum_return"] + int(st["score"].sum()) == int(reward.sum())
    assert rb.len() == n * 16
    idx = rb.generate_distinct_random_ids(512, 1)
    g = rb.get_many(idx, qlb.LAYOUT_U8_BHYX)
    assert set(np.unique(g.state).tolist()) <= {0, 96, 236, 255}
    tt, ee = (k - 16) + idx // n, idx % n
    assert np.array_equal(g.action, acts[tt, ee]) and np.array_equal(g.reward, reward[tt, ee]) and np.array_equal(g.done, done[tt, ee])
    nxt = idx + n
    ok = (nxt < rb.len()) & (g.done == 0)
    assert np.array_equal(g.state_next[ok], rb.get_many(nxt[ok], qlb.LAYOUT_U8_BHYX).state)
    assert env.error_flags() & ~qlb.ENVERR_DEGENERATE == 0
    env.close()


def test_config1_every_env_every_step(qlb, O):
    """BASELINE configs[1] in full: 4,096 envs x 10,000 env-steps (the configs[0] trace length) on the bench's launch shape
    (64 env-steps per launch and odd tails), EVERY env and EVERY step against the oracle: reward and done of all 4.1e7
    env-steps, and at every chunk boundary the complete state (f32 bit patterns, brick masks, scores, episode counters,
    sticky error flags) and the 4-frame stacks of all envs. The oracle runs as per-thread sub-shards with the same global
    env ids (oracle.ShardedVecEnv). QLC_FULL_TRACE_STEPS shortens it."""
    import os
    n, seed = 4096, 20261018
    total = int(os.environ.get("QLC_FULL_TRACE_STEPS", "10000"))
    env = qlb.BreakoutEnvironment(n_envs=n, seed=seed, replay_capacity=n * 8)
    ora = O.ShardedVecEnv(n, seed=seed)
    t, chunk_i, n_done, n_reward = 0, 0, 0, 0.0
    while t < total:
        k = min(total - t, (640, 64, 1, 999, 1280)[chunk_i % 5]); chunk_i += 1
        acts = O.synthetic_actions(seed, 0, n, t, k)
        reward, done = env.step_many(acts)
        r, d = ora.run(acts)
        bad = np.nonzero((reward != r) | (done != d))
        assert bad[0].size == 0, "reward/done differ first at step %d env %d" % (t + bad[0][0], bad[1][0])
        t += k
        gs, os_ = env.read_state(), ora.state()
        _assert_state_equal(gs, os_, "t=%d" % t, nan_ok=True)
        assert np.array_equal(gs["err"], os_["err"]), "sticky error flags differ at t=%d" % t
        assert np.array_equal(env.obs(qlb.LAYOUT_U8_BHYX), ora.obs_u8()), "frame stacks differ at t=%d" % t
        n_done += int(d.sum()); n_reward += float(r.sum())
    st = env.stats()
    assert st["steps"] == n * total and st["episodes"] == n_done
    assert st["sum_return"] + int(env.read_state()["score"].sum()) == int(n_reward)
    assert n_done > n * total // 400          # random play loses the ball every ~200 steps: thousands of device-side restarts per env batch
    env.close(); ora.close()


def test_pipelined_host_steps_equal_synchronous_ones(qlb, O):
    """qlc_env_step_host_submit / _wait (several host-buffer steps in flight, page-locked buffers) produce the bytes of the
    synchronous qlc_env_step_host and of the oracle; pageable buffers are refused."""
    n, k, rounds, seed = 96, 7, 6, 31
    acts = np.random.default_rng(seed).integers(0, 3, size=(rounds, k, n), dtype=np.uint8)
    env_a = qlb.BreakoutEnvironment(n_envs=n, seed=seed, replay_capacity=n * 64)
    env_s = qlb.BreakoutEnvironment(n_envs=n, seed=seed, replay_capacity=n * 64)
    ora = O.VecEnv(n, seed=seed, replay_capacity=n * 64)
    pa = [qlb.PinnedArray((k, n), np.uint8) for _ in range(rounds)]
    pr = [qlb.PinnedArray((k, n), np.float32) for _ in range(rounds)]
    pd = [qlb.PinnedArray((k, n), np.uint8) for _ in range(rounds)]
    for i in range(rounds):
        pa[i].array[:] = acts[i]
        env_a.step_many_submit(pa[i].array, pr[i].array, pd[i].array)      # all rounds queued before the first wait
    env_a.step_many_wait()
    for i in range(rounds):
        r, d = env_s.step_many(acts[i])
        assert np.array_equal(pr[i].array, r) and np.array_equal(pd[i].array, d), "round %d" % i
        for s in range(k):
            ro, do = ora.step(acts[i, s])
            assert np.array_equal(ro, r[s]) and np.array_equal(do, d[s])
    _assert_state_equal(env_a.read_state(), env_s.read_state(), "pipelined vs synchronous")
    assert np.array_equal(env_a.obs(qlb.LAYOUT_U8_BHYX), ora.obs_u8())
    with pytest.raises(qlb.QlError):
        env_a.step_many_submit(acts[0].copy(), np.empty((k, n), np.float32), np.empty((k, n), np.uint8))   # pageable
    bad = qlb.PinnedArray((k, n), np.uint8); bad.array[:] = 0; bad.array[2, 5] = 3
    with pytest.raises(qlb.QlError, match="value out of range"):
        env_a.step_many_submit(bad.array, pr[0].array, pd[0].array)
    env_a.close(); env_s.close()
