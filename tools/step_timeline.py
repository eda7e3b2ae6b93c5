"""Scratch (profiling build, QLC_LIB=.../libqlcuda_prof.so QLC_TIMELINE_FILE=...): where a single-step launch of n envs spends its
time, from clock64 stamps taken by every CTA of the LAST launch. python tools/step_timeline.py run N | python tools/step_timeline.py show FILE"""
import importlib, os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
NAMES = ["gt entry", "entry", "prologue done", "dep wait done", "state+action loaded", "physics done", "record published", "physics warp end",
         "render got record", "frame0 phase 1 done", "frame0 phase 2 done", "render loop end", "store drained", "gt exit", "frame0 proxy fence done", "frame0 store committed"]
if sys.argv[1] == "run":
    sys.path.insert(0, ROOT)
    import torch
    q = importlib.import_module("q-learning_b200")
    n = int(sys.argv[2])
    env = q.BreakoutEnvironment(n_envs=n, seed=1, replay_capacity=n * 16)
    s = torch.cuda.current_stream().cuda_stream
    gen = torch.Generator(device="cuda"); gen.manual_seed(1)
    acts = torch.randint(0, 3, (64, n), dtype=torch.uint8, device="cuda", generator=gen)
    for _ in range(20): env.step_device(acts.data_ptr(), 64, None, None, s)          # a lived-in state: random policy for 1,280 steps
    k = int(sys.argv[3]) if len(sys.argv) > 3 else 1                                      # steps per launch (QLC_DEBUG_SKIP=16+j: stamps at step j)
    for i in range(200): env.step_device(acts[i % 32:].data_ptr(), k, None, None, s)     # back-to-back launches; the last one is recorded
    torch.cuda.synchronize()
    env.close()
else:
    t = np.fromfile(sys.argv[2], dtype=np.uint64).reshape(1024, 16).astype(np.int64)
    used = (t[:, 1] != 0) & (t[:, 0] > t[:, 0].max() - 200_000)      # the CTAs of the last launch (older launches used more CTAs)
    t = t[used]
    print("%d CTAs recorded; cycles relative to each CTA's own entry (median / p90 / max), 1.965 GHz" % len(t))
    for i in (2, 3, 4, 5, 6, 7, 8, 9, 10, 14, 15, 11, 12):
        d = (t[:, i] - t[:, 1])[t[:, i] != 0]
        if len(d): print("  %-22s %7.0f %7.0f %7.0f cyc = %5.2f / %5.2f / %5.2f us" % (NAMES[i], np.median(d), np.percentile(d, 90), d.max(), np.median(d) / 1965, np.percentile(d, 90) / 1965, d.max() / 1965))
    gt0, gt1 = t[:, 0], t[:, 13]
    print("  globaltimer: first entry -> last exit %.2f us; entries spread over %.2f us; per-CTA lifetime median %.2f us" % (
        (gt1.max() - gt0.min()) / 1e3, (gt0.max() - gt0.min()) / 1e3, np.median(gt1 - gt0) / 1e3))
