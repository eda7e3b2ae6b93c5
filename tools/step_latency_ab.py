"""Scratch: where a single-step launch (the learner-driven mode) spends its time. One subprocess per library switch
(QLC_ADVANCE_CFG / QLC_EPC / QLC_DEBUG_SKIP / QLC_STEP_PDL are read at env creation)."""
import importlib, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

def child():
    sys.path.insert(0, ROOT)
    import torch
    q = importlib.import_module("q-learning_b200")
    n, cap = int(sys.argv[2]), int(sys.argv[3])
    env = q.BreakoutEnvironment(n_envs=n, seed=1, replay_capacity=cap)
    s = torch.cuda.current_stream().cuda_stream
    acts = torch.randint(0, 3, (64, n), dtype=torch.uint8, device="cuda")
    rew = torch.empty((1, n), dtype=torch.float32, device="cuda"); dn = torch.empty((1, n), dtype=torch.uint8, device="cuda")
    env.step_device(acts.data_ptr(), 64, None, None, s)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    out = []
    for outs in (False, True):
        fn = (lambda: env.step_device(acts.data_ptr(), 1, rew.data_ptr(), dn.data_ptr(), s)) if outs else (lambda: env.step_device(acts.data_ptr(), 1, None, None, s))
        for _ in range(50): fn()
        torch.cuda.synchronize(); e0.record()
        for _ in range(1000): fn()
        e1.record(); torch.cuda.synchronize()
        out.append(e0.elapsed_time(e1))
    print("%7.2f us (no outputs) %7.2f us (reward+done)" % (out[0], out[1]), flush=True)

if len(sys.argv) > 1 and sys.argv[1] == "child":
    child(); sys.exit(0)
cases = [({"QLC_ADVANCE_CFG": str(c)}, n, n * 16) for n in (1024, 4096, 16384, 65536) for c in (1, 2, 5)]
if len(sys.argv) > 1 and sys.argv[1] == "breakdown":
    cases = [({}, 4096, 4096 * 16), ({}, 4096, 1 << 20), ({"QLC_STEP_PDL": "0"}, 4096, 1 << 20),
             ({"QLC_DEBUG_SKIP": "1"}, 4096, 1 << 20), ({"QLC_DEBUG_SKIP": "2"}, 4096, 1 << 20),
             ({"QLC_EPC": "28"}, 4096, 1 << 20), ({}, 1024, 1 << 18), ({}, 148, 148 * 64), ({}, 1, 4096), ({}, 65536, 65536 * 16)]
for envv, n, cap in cases:
    r = subprocess.run([sys.executable, __file__, "child", str(n), str(cap)], env=dict(os.environ, **envv), capture_output=True, text=True)
    print("n=%6d cap=%8d %-44s %s" % (n, cap, envv, (r.stdout.strip() or r.stderr.strip()[-300:])), flush=True)
