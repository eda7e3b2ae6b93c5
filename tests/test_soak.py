"""Soak (GPU): thousands of back-to-back chunked launches — every launch hands env state from CTA to CTA through HBM — then
(1) no hand-over time-out / unexpected error flag, (2) statistics identities, (3) 32 envs replayed from t = 0 on the CPU oracle
over their WHOLE trajectory: state, sticky flags and frame stacks bit for bit. Default 3,000 launches (7.9e8 env-steps, 192,000
steps per env); QLC_SOAK_LAUNCHES=20000 is the 5.2e9-env-step run quoted in profiles/r01_notes.md."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_soak_chunked_launches_and_full_trajectory_replay(qlb, O):
    import torch
    n, k, seed = 4096, 64, 99
    launches = int(os.environ.get("QLC_SOAK_LAUNCHES", "3000"))
    env = qlb.BreakoutEnvironment(n_envs=n, seed=seed, replay_capacity=n * 64)
    acts = O.synthetic_actions(seed, 0, n, 0, k)                     # the same 64-step action block every launch
    a_dev = torch.from_numpy(acts).cuda()
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(launches):
        env.step_device(a_dev.data_ptr(), k, None, None, s)
    torch.cuda.synchronize()
    flags = env.error_flags()
    st, stats = env.read_state(), env.stats()
    assert stats["steps"] == n * k * launches
    assert not (flags & qlb.ENVERR_HANDOVER) and not (flags & qlb.ENVERR_ACTION)
    assert stats["sum_return"] >= stats["episodes"] * 0 and stats["episodes"] > n
    sub = np.unique(np.concatenate([np.arange(0, n, 137), np.nonzero(st["err"])[0][:4]])).astype(np.int64)[:32]
    obs = env.obs(qlb.LAYOUT_U8_BHYX)
    ora = [O.VecEnv(1, seed=seed, env_id_base=int(e)) for e in sub]
    handles = (O.C.c_void_p * len(ora))(*[v.h for v in ora])
    a_sub = np.ascontiguousarray(np.tile(acts[:, sub], (launches, 1)))   # [k * launches][len(sub)]
    offs = np.arange(len(sub), dtype=np.uint32); sizes = np.ones(len(sub), dtype=np.uint32)
    r = np.empty(a_sub.shape, dtype=np.float32); d = np.empty(a_sub.shape, dtype=np.uint8)
    O.lib().orc_parts_run(handles, O._p(offs), O._p(sizes), len(sub), len(sub), a_sub.shape[0], O._p(a_sub), O._p(r), O._p(d))
    for j, e in enumerate(sub):
        so = ora[j].state()
        for key in ("ball_cx", "ball_cy", "ball_dx", "ball_dy", "pad_min_x", "pad_max_x", "pad_speed"):
            a, b = st[key][e], so[key][0]
            assert a.view(np.uint32) == b.view(np.uint32) or (a != a and b != b), (int(e), key, a, b)
        for key in ("bricks", "score", "episode_step", "err"):
            assert st[key][e] == so[key][0], (int(e), key)
        assert np.array_equal(obs[e], ora[j].obs_u8()[0]), "frames of env %d" % e
        ora[j].close()
    assert d.sum() > len(sub) * launches // 8          # hundreds of episodes per replayed env
    env.close()


def test_soak_single_step_launches_every_env(qlb, O):
    """The learner-driven mode at BASELINE configs[1] size: 4,096 envs advanced by 2,400 launches of 1, 2 or 3 steps (the shapes whose render
    warps pre-draw the brick band and whose physics lanes request the action first), uniform random actions, every reward and done of
    every step plus the final state and frame stacks of EVERY env against the oracle; then the same on a 65,536-env shard for 60 launches
    (the persistent form: later items reuse the CTAs' frames)."""
    import torch
    for n, launches in ((4096, 2400), (65536, 60)):
        seed = 123 + n
        env = qlb.BreakoutEnvironment(n_envs=n, seed=seed, replay_capacity=n * 8)
        ora = O.ShardedVecEnv(n, seed=seed)
        lens = [1 + (i % 7 == 3) + 2 * (i % 11 == 5) for i in range(launches)]          # mostly 1, some 2, 3, a few 4 (the multi-step shape in between)
        total = sum(lens)
        acts = O.synthetic_actions(seed, 0, n, 0, total)
        a_dev = torch.from_numpy(acts).cuda()
        rew = torch.empty((total, n), dtype=torch.float32, device="cuda"); dn = torch.empty((total, n), dtype=torch.uint8, device="cuda")
        s = torch.cuda.current_stream().cuda_stream
        at = 0
        for k in lens:
            env.step_device(a_dev[at:].data_ptr(), k, rew[at:].data_ptr(), dn[at:].data_ptr(), s)
            at += k
        torch.cuda.synchronize()
        r_o, d_o = ora.run(acts)
        assert np.array_equal(rew.cpu().numpy(), r_o) and np.array_equal(dn.cpu().numpy(), d_o), "reward / done differ (n=%d)" % n
        st, so = env.read_state(), ora.state()
        for key in ("ball_cx", "ball_cy", "ball_dx", "ball_dy", "pad_min_x", "pad_max_x", "pad_speed"):
            assert np.array_equal(st[key].view(np.uint32), so[key].view(np.uint32)), (n, key)
        for key in ("bricks", "score", "episode_step", "err"):
            assert np.array_equal(st[key], so[key]), (n, key)
        assert np.array_equal(env.obs(qlb.LAYOUT_U8_BHYX), ora.obs_u8()), "frame stacks differ (n=%d)" % n
        if n == 4096:
            assert d_o.sum() > 4 * n    # several finished episodes per env: resets and fresh frames are covered
        env.close(); ora.close()
