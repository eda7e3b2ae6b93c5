"""Scratch: the closed actor loop (Q-network -> greedy action -> env step) run twice from the same seed gives the same bits."""
import hashlib, importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
q = importlib.import_module("q-learning_b200")


def run(n, iters, seed):
    env = q.BreakoutEnvironment(n_envs=n, seed=seed, replay_capacity=n * 64)
    rng = np.random.default_rng(5)
    w = {}
    for name, shape in q.QNET_SHAPES.items():
        lim = np.sqrt(6.0 / (int(np.prod(shape[:-1])) + shape[-1])) if name.endswith("kernel") else 0.05
        w[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
    net = q.QNetwork(env, w)
    s = torch.cuda.current_stream().cuda_stream
    acts = torch.empty((1, n), dtype=torch.uint8, device="cuda")
    rew = torch.empty((1, n), dtype=torch.float32, device="cuda"); done = torch.empty((1, n), dtype=torch.uint8, device="cuda")
    total, hist = torch.zeros((), device="cuda"), torch.zeros(3, device="cuda")
    for _ in range(iters):
        net.forward_device(None, n, 0, None, acts.data_ptr(), None, s)
        env.step_device(acts.data_ptr(), 1, rew.data_ptr(), done.data_ptr(), s)
        total += rew.sum(); hist += torch.bincount(acts[0].long(), minlength=3).float()
    torch.cuda.synchronize()
    st = env.read_state()
    h = hashlib.sha256()
    for k in sorted(st): h.update(np.ascontiguousarray(st[k]).tobytes())
    h.update(env.obs(q.LAYOUT_U8_BHYX).tobytes())
    out = (h.hexdigest(), float(total), hist.tolist(), env.error_flags())
    net.close(); env.close()
    return out


a = run(2048, 600, 9); b = run(2048, 600, 9)
print(a); print(b); print("reproducible:", a == b)
