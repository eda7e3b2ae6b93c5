"""Where a host-buffer minibatch gather (qlc_replay_gather_host, f32 [b][x][y][slot]) spends its time: run on the GPU box.
python tools/host_gather_breakdown.py [batch]"""
import importlib
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
q = importlib.import_module("q-learning_b200")


def timeit(fn, reps=200):
    for _ in range(10):
        fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps * 1e6


def main():
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    env = q.BreakoutEnvironment(n_envs=4096, seed=1, replay_capacity=1 << 20)
    rb = q.ReplayBuffer(env)
    env.step_many(np.random.default_rng(0).integers(0, 3, size=(64, 4096), dtype=np.uint8))
    rng = np.random.default_rng(1)
    ids = rng.choice(rb.len(), size=batch, replace=False).astype(np.uint32)
    print("host threads (QLC_HOST_THREADS=%s), batch %d" % (os.environ.get("QLC_HOST_THREADS", "auto"), batch))
    print("ids: rng.choice                       %7.1f us" % timeit(lambda: rng.choice(rb.len(), size=batch, replace=False).astype(np.uint32)))
    print("scalars only                          %7.1f us" % timeit(lambda: rb.get_many(ids, q.LAYOUT_U8_BHYX, want_state=False, want_next=False)))
    for layout, name in ((q.LAYOUT_U8_BHYX, "u8 bhyx"), (q.LAYOUT_U8_BXYH, "u8 bxyh"), (q.LAYOUT_F32_BXYH, "f32 bxyh")):
        print("%-9s state only  (pinned)        %7.1f us" % (name, timeit(lambda: rb.get_many(ids, layout, want_next=False, reuse=True))))
        print("%-9s state + next (pinned)       %7.1f us" % (name, timeit(lambda: rb.get_many(ids, layout, reuse=True))))
        print("%-9s state + next (pageable new) %7.1f us" % (name, timeit(lambda: rb.get_many(ids, layout), reps=50)))
    env.close()


if __name__ == "__main__":
    main()
