"""Launch sequence for `ncu --set full` (round 2): every hot kernel a few times, in a fixed order, nothing else.
  3 x env_advance (4,096 envs x 64 steps, the bench launch)   3 x env_advance single step (4,096 envs)   3 x single step with the
  random policy drawn in the kernel   then, per layout (u8 [b][slot][y][x], f32 [b][x][y][slot]): 3 x one-launch sample+gather of
  one minibatch of 32, of 512, of 256 minibatches of 32; 3 x gather with given indices (8,192 transitions); 3 x index-only sample of 512 and of 32; 3 x the host call
  get_many(f32) for 32 and 512 (the streamed gather kernel)."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
q = importlib.import_module("q-learning_b200")
s = torch.cuda.current_stream().cuda_stream
n = 4096
env = q.BreakoutEnvironment(n_envs=n, seed=3, replay_capacity=1 << 20)
rb = q.ReplayBuffer(env)
acts = torch.randint(0, 3, (64, n), dtype=torch.uint8, device="cuda")
for _ in range(3):
    env.step_device(acts.data_ptr(), 64, None, None, s)
for _ in range(3):
    env.step_device(acts.data_ptr(), 1, None, None, s)
for _ in range(3):
    env.step_random_device(1, None, None, None, s)
torch.cuda.synchronize()
per = 4 * 84 * 84
for layout, dt in ((q.LAYOUT_U8_BHYX, torch.uint8), (q.LAYOUT_F32_BXYH, torch.float32)):
    for batch, nb in ((32, 1), (512, 1), (32, 256)):
        m = batch * nb
        idx = torch.empty((m,), dtype=torch.int32, device="cuda")
        st = torch.empty((m, per), dtype=dt, device="cuda"); nx = torch.empty((m, per), dtype=dt, device="cuda")
        r = torch.empty((m,), dtype=torch.float32, device="cuda"); a = torch.empty((m,), dtype=torch.uint8, device="cuda"); d = torch.empty((m,), dtype=torch.uint8, device="cuda")
        for c in range(3):
            rb.sample_gather_device(batch, nb, c, layout, idx.data_ptr(), st.data_ptr(), nx.data_ptr(), r.data_ptr(), a.data_ptr(), d.data_ptr(), s)
        if nb == 256:
            for c in range(3):
                rb.gather_device(idx.data_ptr(), m, layout, st.data_ptr(), nx.data_ptr(), r.data_ptr(), a.data_ptr(), d.data_ptr(), s)
        torch.cuda.synchronize()
idx = torch.empty((512,), dtype=torch.int32, device="cuda")
for c in range(3):
    rb.sample_device(512, 1, c, idx.data_ptr(), s)
    rb.sample_device(32, 1, c, idx.data_ptr(), s)
torch.cuda.synchronize()
# the reference-shaped host call: f32 tensors in host memory (gather_xyh_stream_kernel: bulk stores into page-locked memory + arrival flags)
import numpy as np
rng = np.random.default_rng(0)
for batch in (32, 512):
    for c in range(3):
        rb.get_many(rng.choice(rb.len(), size=batch, replace=False).astype(np.uint32), q.LAYOUT_F32_BXYH, reuse=True)
env.close()
