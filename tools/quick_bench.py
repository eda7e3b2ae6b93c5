"""Scratch timing of the hot-path kernels on one GPU (not the contract bench; see bench.py)."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch

q = importlib.import_module("q-learning_b200")


def time_advance(n_envs, k_steps, launches, cfg, cap_steps=64):
    os.environ["QLC_ADVANCE_CFG"] = str(cfg)
    env = q.BreakoutEnvironment(n_envs=n_envs, seed=1, replay_capacity=n_envs * cap_steps)
    acts = torch.randint(0, 3, (k_steps, n_envs), dtype=torch.uint8, device="cuda")
    rew = torch.empty((k_steps, n_envs), dtype=torch.float32, device="cuda")
    done = torch.empty((k_steps, n_envs), dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(3):
        env.step_device(acts.data_ptr(), k_steps, rew.data_ptr(), done.data_ptr(), s)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(launches):
        env.step_device(acts.data_ptr(), k_steps, rew.data_ptr(), done.data_ptr(), s)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    steps = n_envs * k_steps * launches
    rate = steps / (ms * 1e-3)
    print("advance cfg=%d N=%6d K=%3d: %8.3f ms/launch  %.3e env-steps/s  %.1f GB/s (7154 B/step)" % (
        cfg, n_envs, k_steps, ms / launches, rate, rate * 7154 / 1e9), flush=True)
    env.close()


def time_gather(n_envs, batch, n_batches, layout, reps=20):
    env = q.BreakoutEnvironment(n_envs=n_envs, seed=1, replay_capacity=n_envs * 256)
    rb = q.ReplayBuffer(env)
    acts = torch.randint(0, 3, (64, n_envs), dtype=torch.uint8, device="cuda")
    s = torch.cuda.current_stream().cuda_stream
    for _ in range(4):
        env.step_device(acts.data_ptr(), 64, None, None, s)
    n = batch * n_batches
    idx = torch.empty((n,), dtype=torch.int32, device="cuda")
    per = 4 * 84 * 84
    dt = torch.uint8 if layout == q.LAYOUT_U8_BHYX else torch.float32
    st = torch.empty((n, per), dtype=dt, device="cuda")
    nx = torch.empty((n, per), dtype=dt, device="cuda")
    r = torch.empty((n,), dtype=torch.float32, device="cuda")
    a = torch.empty((n,), dtype=torch.uint8, device="cuda")
    d = torch.empty((n,), dtype=torch.uint8, device="cuda")
    def once(c):
        rb.sample_device(batch, n_batches, c, idx.data_ptr(), s)
        rb.gather_device(idx.data_ptr(), n, layout, st.data_ptr(), nx.data_ptr(), r.data_ptr(), a.data_ptr(), d.data_ptr(), s)
    for c in range(3):
        once(c)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for c in range(reps):
        once(100 + c)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / reps
    bytes_per = 91744 if layout == q.LAYOUT_U8_BHYX else 261088
    rate = n / (ms * 1e-3)
    print("sample+gather layout=%d B=%d x %d: %.3f ms  %.3e transitions/s  %.1f GB/s" % (layout, batch, n_batches, ms, rate, rate * bytes_per / 1e9), flush=True)
    env.close()


if __name__ == "__main__":
    print(torch.cuda.get_device_name(0))
    for n, k in ((4096, 64), (4096, 1), (4096, 2), (1024, 1), (65536, 1), (65536, 16), (8192, 64)):
        time_advance(n, k, max(5, 640 // k), 0, cap_steps=64)
