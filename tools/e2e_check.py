import importlib, os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
q = importlib.import_module("q-learning_b200")
for zc in ("2", "1", "0", "2", "1"):
    os.environ["QLC_ZERO_COPY"] = zc
    env = q.BreakoutEnvironment(n_envs=4096, seed=1, replay_capacity=1 << 20)
    pa, pr, pd = q.PinnedArray((64, 4096), np.uint8), q.PinnedArray((64, 4096), np.float32), q.PinnedArray((64, 4096), np.uint8)
    pa.array[:] = np.random.default_rng(0).integers(0, 3, size=(64, 4096), dtype=np.uint8)
    for _ in range(5):
        env.step_many(pa.array, out=(pr.array, pd.array))
    t0 = time.perf_counter()
    for _ in range(100):
        env.step_many(pa.array, out=(pr.array, pd.array))
    dt = (time.perf_counter() - t0) / 100
    print("zero_copy", zc, "%.3f ms/step  %.3e env-steps/s" % (dt * 1e3, 4096 * 64 / dt), "reward sum", float(pr.array.sum()), "done", int(pd.array.sum()))
    # parity of outputs between the two modes on a fresh env
    env.close()
a = {}
for zc in ("2", "1", "0"):
    os.environ["QLC_ZERO_COPY"] = zc
    env = q.BreakoutEnvironment(n_envs=300, seed=5)
    pa, pr, pd = q.PinnedArray((200, 300), np.uint8), q.PinnedArray((200, 300), np.float32), q.PinnedArray((200, 300), np.uint8)
    pa.array[:] = np.random.default_rng(1).integers(0, 3, size=(200, 300), dtype=np.uint8)
    env.step_many(pa.array, out=(pr.array, pd.array))
    a[zc] = (pr.array.copy(), pd.array.copy())
    env.close()
print("modes agree:", all(np.array_equal(a[m][0], a["0"][0]) and np.array_equal(a[m][1], a["0"][1]) for m in ("1", "2")), a["1"][1].sum())
