import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
q = importlib.import_module("q-learning_b200")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
env = q.BreakoutEnvironment(n_envs=n, seed=1, replay_capacity=n * 8)
acts = torch.randint(0, 3, (8, n), dtype=torch.uint8, device="cuda")
s = torch.cuda.current_stream().cuda_stream
env.step_device(acts.data_ptr(), 8, None, None, s)
rng = np.random.default_rng(0)
net = q.QNetwork(env, {k: (rng.standard_normal(sh) * 0.02).astype(np.float32) for k, sh in q.QNET_SHAPES.items()})
qv = torch.empty((n, 3), dtype=torch.float32, device="cuda")
for _ in range(4):
    net.forward_device(None, n, 0, qv.data_ptr(), None, None, s)
torch.cuda.synchronize()
print("ok", float(qv.abs().sum()))
