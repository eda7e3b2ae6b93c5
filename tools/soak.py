"""Soak: 20,000 chunked 64-step launches of 4,096 envs (5.2e9 env-steps, 1.28 M steps per env; every launch hands env state from
CTA to CTA through HBM), then (1) no hand-over time-out / unexpected error flag, (2) statistics identities, (3) 32 envs replayed
from t = 0 on the CPU oracle over the whole 1.28 M-step trajectory - state, sticky flags and frame stacks bit for bit."""
import importlib, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
q = importlib.import_module("q-learning_b200")
from oracle import oracle as O
O.build()
n, k, launches, seed = 4096, 64, int(os.environ.get("SOAK_LAUNCHES", "20000")), 99
env = q.BreakoutEnvironment(n_envs=n, seed=seed, replay_capacity=n * 64)
acts = O.synthetic_actions(seed, 0, n, 0, k)                     # the same 64-step action block every launch
a_dev = torch.from_numpy(acts).cuda()
s = torch.cuda.current_stream().cuda_stream
t0 = time.time()
for i in range(launches):
    env.step_device(a_dev.data_ptr(), k, None, None, s)
torch.cuda.synchronize()
dt = time.time() - t0
flags = env.error_flags()
st = env.read_state(); stats = env.stats()
print("%d launches, %.3e env-steps in %.1f s (%.3e /s); error flags OR = %d (hand-over bit %s)" % (launches, n * k * launches, dt, n * k * launches / dt, flags, bool(flags & q.ENVERR_HANDOVER) if hasattr(q, "ENVERR_HANDOVER") else "n/a"), flush=True)
assert stats["steps"] == n * k * launches
assert not (flags & getattr(q, "ENVERR_HANDOVER", 64))
sub = np.unique(np.concatenate([np.arange(0, n, 137), np.nonzero(st["err"])[0][:4]])).astype(np.int64)[:32]
obs = env.obs(q.LAYOUT_U8_BHYX)
t0 = time.time()
parts = [O.ShardedVecEnv(1, seed=seed, env_id_base=int(e), parts=1) for e in sub]
handles = (O.C.c_void_p * len(parts))(*[p.parts[0].h for p in parts])
a_sub = np.ascontiguousarray(np.tile(acts[:, sub], (launches, 1)))   # [k * launches][len(sub)]
offs = np.arange(len(sub), dtype=np.uint32); sizes = np.ones(len(sub), dtype=np.uint32)
r = np.empty(a_sub.shape, dtype=np.float32); d = np.empty(a_sub.shape, dtype=np.uint8)
O.lib().orc_parts_run(handles, O._p(offs), O._p(sizes), len(sub), len(sub), a_sub.shape[0], O._p(a_sub), O._p(r), O._p(d))
bad = 0
for j, e in enumerate(sub):
    so = parts[j].state()
    for key in ("ball_cx", "ball_cy", "ball_dx", "ball_dy", "pad_min_x", "pad_max_x", "pad_speed"):
        a, b = st[key][e], so[key][0]
        if not (a.view(np.uint32) == b.view(np.uint32) or (a != a and b != b)): bad += 1; print("MISMATCH", e, key, a, b)
    for key in ("bricks", "score", "episode_step", "err"):
        if st[key][e] != so[key][0]: bad += 1; print("MISMATCH", e, key, st[key][e], so[key][0])
    if not np.array_equal(obs[e], parts[j].obs_u8()[0]): bad += 1; print("MISMATCH frames", e)
print("oracle replay of %d envs x %d steps: %.1f s, mismatches %d, episodes per env ~%d, flagged among them %d" % (
    len(sub), k * launches, time.time() - t0, bad, int(d.sum() / len(sub)), int(sum(st["err"][e] != 0 for e in sub))), flush=True)
assert bad == 0
print("soak ok")
