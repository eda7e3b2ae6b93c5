"""Scratch: latency of the literal drop-in use - ONE env, one step per host call (what the unchanged learner does)."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
q = importlib.import_module("q-learning_b200")
for n in (1, 32):
    env = q.BreakoutEnvironment(n_envs=n, seed=1, replay_capacity=1 << 16)
    rb = q.ReplayBuffer(env)
    pa, pr, pd = q.PinnedArray((1, n), np.uint8), q.PinnedArray((1, n), np.float32), q.PinnedArray((1, n), np.uint8)
    pa.array[:] = 0
    a_pageable = np.zeros((1, n), dtype=np.uint8)
    for name, fn in (("step_many pinned", lambda: env.step_many(pa.array, out=(pr.array, pd.array))),
                     ("step_many pageable", lambda: env.step_many(a_pageable))):
        for _ in range(200): fn()
        t0 = time.perf_counter()
        for _ in range(2000): fn()
        print("n=%2d %-20s %7.2f us per call" % (n, name, (time.perf_counter() - t0) / 2000 * 1e6), flush=True)
    for _ in range(100): env.step_many(pa.array, out=(pr.array, pd.array))
    t0 = time.perf_counter()
    for c in range(500):
        idx = rb.generate_distinct_random_ids(32, c)
        s = rb.get_many(idx, q.LAYOUT_F32_BXYH)
    print("n=%2d sample+get_many(32, f32) host %7.2f us per call" % (n, (time.perf_counter() - t0) / 500 * 1e6), flush=True)
    t0 = time.perf_counter()
    for c in range(500):
        idx = rb.generate_distinct_random_ids(32, c)
        s = rb.get_many(idx, q.LAYOUT_F32_BXYH, reuse=True)
    print("n=%2d sample+get_many(32, f32, pinned reuse) %7.2f us per call" % (n, (time.perf_counter() - t0) / 500 * 1e6), flush=True)
    ref = rb.get_many(idx, q.LAYOUT_F32_BXYH)
    assert np.array_equal(ref.state, s.state) and np.array_equal(ref.state_next, s.state_next)
    t0 = time.perf_counter()
    for c in range(500):
        o = env.obs(q.LAYOUT_F32_BXYH)
    print("n=%2d obs f32 host %7.2f us per call" % (n, (time.perf_counter() - t0) / 500 * 1e6), flush=True)
    env.close()
