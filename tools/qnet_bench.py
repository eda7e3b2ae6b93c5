"""Scratch timing of the tensor-core Q-network forward (not the contract bench)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
q = importlib.import_module("q-learning_b200")

FLOP_PER_ITEM = 2 * (400 * 32 * 256 + 81 * 64 * 512 + 49 * 64 * 576 + 512 * 3136 + 3 * 512)


def weights(seed=0):
    rng = np.random.default_rng(seed)
    return {k: (rng.standard_normal(s) * 0.02).astype(np.float32) for k, s in q.QNET_SHAPES.items()}


def run(n_envs, reps=20):
    env = q.BreakoutEnvironment(n_envs=n_envs, seed=1, replay_capacity=n_envs * 8)
    acts = torch.randint(0, 3, (8, n_envs), dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    env.step_device(acts.data_ptr(), 8, None, None, s)
    net = q.QNetwork(env, weights())
    qv = torch.empty((n_envs, 3), dtype=torch.float32, device="cuda"); act = torch.empty((n_envs,), dtype=torch.uint8, device="cuda")
    for _ in range(3):
        net.forward_device(None, n_envs, 0, qv.data_ptr(), act.data_ptr(), None, s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        net.forward_device(None, n_envs, 0, qv.data_ptr(), act.data_ptr(), None, s)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    print("qnet forward N=%6d: %8.3f ms  %.3e obs/s  %.1f TFLOP/s" % (n_envs, ms, n_envs / (ms * 1e-3), FLOP_PER_ITEM * n_envs / (ms * 1e-3) / 1e12), flush=True)
    net.close(); env.close()


if __name__ == "__main__":
    for n in (32, 512, 4096, 16384):
        run(n)
