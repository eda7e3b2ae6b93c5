"""Scratch: latency of single-step launches and of the closed actor loop (QLC_STEP_PDL=0/1 A/B)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
q = importlib.import_module("q-learning_b200")
for n in (4096, 65536):
    env = q.BreakoutEnvironment(n_envs=n, seed=1, replay_capacity=n * 16)
    s = torch.cuda.current_stream().cuda_stream
    acts = torch.randint(0, 3, (1, n), dtype=torch.uint8, device="cuda")
    rng = np.random.default_rng(0)
    net = q.QNetwork(env, {k: (rng.standard_normal(sh) * 0.02).astype(np.float32) for k, sh in q.QNET_SHAPES.items()})
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for name, fn in (("step x1", lambda: env.step_device(acts.data_ptr(), 1, None, None, s)),
                     ("qnet -> step", lambda: (net.forward_device(None, n, 0, None, acts.data_ptr(), None, s), env.step_device(acts.data_ptr(), 1, None, None, s)))):
        for _ in range(20): fn()
        torch.cuda.synchronize(); e0.record()
        for _ in range(200): fn()
        e1.record(); torch.cuda.synchronize()
        print("n=%6d %-14s %8.2f us" % (n, name, e0.elapsed_time(e1) / 200 * 1e3), flush=True)
    net.close(); env.close()
