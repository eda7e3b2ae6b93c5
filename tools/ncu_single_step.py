"""Target for an ncu source-level capture of the learner-driven launch: 4,096 envs with a lived-in state, then single-step launches."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
q = importlib.import_module("q-learning_b200")
s = torch.cuda.current_stream().cuda_stream
env = q.BreakoutEnvironment(n_envs=4096, seed=3, replay_capacity=1 << 20)
gen = torch.Generator(device="cuda"); gen.manual_seed(1)
acts = torch.randint(0, 3, (64, 4096), dtype=torch.uint8, device="cuda", generator=gen)
for _ in range(20):
    env.step_device(acts.data_ptr(), 64, None, None, s)
for i in range(8):
    env.step_device(acts[i:].data_ptr(), 1, None, None, s)
torch.cuda.synchronize()
env.close()
