"""Target for compute-sanitizer (memcheck / racecheck / synccheck): every hot kernel once or twice on a small shard, checked against
the oracle like smoke().   compute-sanitizer --tool memcheck python tests/sanitize_target.py"""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
q = importlib.import_module("q-learning_b200")
from oracle import oracle as O

n, seed, cap = 48, 77, 48 * 16
env = q.BreakoutEnvironment(n_envs=n, seed=seed, replay_capacity=cap, device=0)
rb = q.ReplayBuffer(env)
ora = O.VecEnv(n, seed=seed, replay_capacity=cap)
acts = O.synthetic_actions(seed, 0, n, 0, 24)
r, d = env.step_many(acts[:20])                       # one multi-step launch (wraps the 20-slot ring)
for s in range(20, 24):                               # single-step launches
    env.step_many(acts[s:s + 1])
for a in acts:
    ora.step(a)
assert np.array_equal(env.obs(q.LAYOUT_U8_BHYX), ora.obs_u8())
assert np.array_equal(env.obs(q.LAYOUT_F32_BXYH), ora.obs_f32())          # streamed host gather (GATHER_CURRENT)
for batch in (32, 200):                                                   # warp sampler / CTA-wide sampler
    idx = rb.generate_distinct_random_ids(batch, 3)
    assert np.array_equal(idx, O.sample_distinct(seed, 3, rb.len(), batch))
    for layout, kind in ((q.LAYOUT_U8_BHYX, "u8"), (q.LAYOUT_F32_BXYH, "f32")):
        g, o = rb.get_many(idx, layout), ora.get_many(idx, kind)
        assert np.array_equal(g.state, o["state"]) and np.array_equal(g.state_next, o["state_next"]) and np.array_equal(g.reward, o["reward"])
        i2, s2 = rb.sample(batch, layout, call_index=5)                   # one-launch sample + gather (host form)
        assert np.array_equal(i2, O.sample_distinct(seed, 5, rb.len(), batch))
        assert np.array_equal(s2.state, ora.get_many(i2, kind)["state"])
print("error flags", env.error_flags(), "lives", int(env.lives().sum()), "stats", env.stats())
env.close(); ora.close()
print("sanitize target ok")
