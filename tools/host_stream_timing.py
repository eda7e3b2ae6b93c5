"""Scratch: QLC_HOST_TIMING=1 timeline of a few streamed f32 host gathers (stderr lines from the library)."""
import importlib, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
q = importlib.import_module("q-learning_b200")
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
env = q.BreakoutEnvironment(n_envs=4096, seed=1, replay_capacity=1 << 20)
rb = q.ReplayBuffer(env)
env.step_many(np.random.default_rng(0).integers(0, 3, size=(64, 4096), dtype=np.uint8))
rng = np.random.default_rng(1)
for i in range(6):
    ids = rng.choice(rb.len(), size=batch, replace=False).astype(np.uint32)
    rb.get_many(ids, q.LAYOUT_F32_BXYH, reuse=True)
env.close()
