"""One forward for ncu: the step kernel on a 65,536-env shard (BASELINE configs[3], 16 env-steps per launch) and as a single-step
launch of 4,096 envs (the learner-driven mode). Launch order: 4 x (65,536 x 16), then 6 x (4,096 x 1)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
q = importlib.import_module("q-learning_b200")
s = torch.cuda.current_stream().cuda_stream
big = q.BreakoutEnvironment(n_envs=65536, seed=3, replay_capacity=65536 * 32)
acts = torch.randint(0, 3, (16, 65536), dtype=torch.uint8, device="cuda")
for _ in range(4):
    big.step_device(acts.data_ptr(), 16, None, None, s)
torch.cuda.synchronize()
big.close()
env = q.BreakoutEnvironment(n_envs=4096, seed=3, replay_capacity=1 << 20)
a1 = torch.randint(0, 3, (1, 4096), dtype=torch.uint8, device="cuda")
for _ in range(6):
    env.step_device(a1.data_ptr(), 1, None, None, s)
torch.cuda.synchronize()
env.close()
