"""Scratch: latency of the learner-driven calls on one GPU — one-launch sample+gather at the reference's batch sizes, the gather
with given indices, single-step launches, and the actor-loop iteration. CUDA events around back-to-back calls, after warm-up."""
import importlib, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
q = importlib.import_module("q-learning_b200")
PEAK = 6537.3


def timed(fn, reps=200, warm=20):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    for i in range(warm): fn(i)
    torch.cuda.synchronize(); e0.record()
    for i in range(reps): fn(warm + i)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    n = 4096
    env = q.BreakoutEnvironment(n_envs=n, seed=1, replay_capacity=1 << 20)
    rb = q.ReplayBuffer(env)
    s = torch.cuda.current_stream().cuda_stream
    acts = torch.randint(0, 3, (64, n), dtype=torch.uint8, device="cuda")
    for _ in range(5): env.step_device(acts.data_ptr(), 64, None, None, s)
    per = 4 * 84 * 84
    for batch in (32, 512):
        for layout, name, dt, bps in ((q.LAYOUT_U8_BHYX, "u8", torch.uint8, 91744), (q.LAYOUT_F32_BXYH, "f32", torch.float32, 261088)):
            idx = torch.empty((batch,), dtype=torch.int32, device="cuda")
            st = torch.empty((batch, per), dtype=dt, device="cuda"); nx = torch.empty((batch, per), dtype=dt, device="cuda")
            r = torch.empty((batch,), dtype=torch.float32, device="cuda"); a = torch.empty((batch,), dtype=torch.uint8, device="cuda"); d = torch.empty((batch,), dtype=torch.uint8, device="cuda")
            fused = timed(lambda c: rb.sample_gather_device(batch, 1, c, layout, idx.data_ptr(), st.data_ptr(), nx.data_ptr(), r.data_ptr(), a.data_ptr(), d.data_ptr(), s))
            given = timed(lambda c: rb.gather_device(idx.data_ptr(), batch, layout, st.data_ptr(), nx.data_ptr(), r.data_ptr(), a.data_ptr(), d.data_ptr(), s))
            only = timed(lambda c: rb.sample_device(batch, 1, c, idx.data_ptr(), s))
            print("B=%3d %-3s sample+gather 1 launch %6.2f us (%.3f of peak) | gather given ids %6.2f us (%.3f) | sample only %5.2f us" % (
                batch, name, fused, batch * bps / fused / 1e3 / PEAK, given, batch * bps / given / 1e3 / PEAK, only), flush=True)
    a1 = torch.randint(0, 3, (1, n), dtype=torch.uint8, device="cuda")
    rew = torch.empty((1, n), dtype=torch.float32, device="cuda"); done = torch.empty((1, n), dtype=torch.uint8, device="cuda")
    print("single step, 4096 envs: %.2f us (no outputs) %.2f us (reward+done)" % (
        timed(lambda c: env.step_device(a1.data_ptr(), 1, None, None, s), 500), timed(lambda c: env.step_device(a1.data_ptr(), 1, rew.data_ptr(), done.data_ptr(), s), 500)), flush=True)
    idx = torch.empty((32,), dtype=torch.int32, device="cuda")
    st = torch.empty((32, per), dtype=torch.float32, device="cuda"); nx = torch.empty((32, per), dtype=torch.float32, device="cuda")
    r = torch.empty((32,), dtype=torch.float32, device="cuda"); a = torch.empty((32,), dtype=torch.uint8, device="cuda"); d = torch.empty((32,), dtype=torch.uint8, device="cuda")

    def it(i):
        ra = torch.randint(0, 3, (1, n), dtype=torch.uint8, device="cuda")
        env.step_device(ra.data_ptr(), 1, rew.data_ptr(), done.data_ptr(), s)
        if i % 4 == 0:
            rb.sample_gather_device(32, 1, i, q.LAYOUT_F32_BXYH, idx.data_ptr(), st.data_ptr(), nx.data_ptr(), r.data_ptr(), a.data_ptr(), d.data_ptr(), s)
    print("actor loop iteration (randint + step + 1/4 sample+gather): %.2f us" % timed(it, 400), flush=True)

    def it2(i):
        env.step_device(a1.data_ptr(), 1, rew.data_ptr(), done.data_ptr(), s)
        if i % 4 == 0:
            rb.sample_gather_device(32, 1, i, q.LAYOUT_F32_BXYH, idx.data_ptr(), st.data_ptr(), nx.data_ptr(), r.data_ptr(), a.data_ptr(), d.data_ptr(), s)
    print("same without the policy kernel: %.2f us" % timed(it2, 400), flush=True)
    def it3(i):
        env.step_random_device(1, None, rew.data_ptr(), done.data_ptr(), s)
        if i % 4 == 0:
            rb.sample_gather_device(32, 1, i, q.LAYOUT_F32_BXYH, idx.data_ptr(), st.data_ptr(), nx.data_ptr(), r.data_ptr(), a.data_ptr(), d.data_ptr(), s)
    print("random policy drawn inside the step kernel: %.2f us; step_random alone %.2f us" % (
        timed(it3, 400), timed(lambda c: env.step_random_device(1, None, rew.data_ptr(), done.data_ptr(), s), 400)), flush=True)
    print("torch.randint alone: %.2f us" % timed(lambda i: torch.randint(0, 3, (1, n), dtype=torch.uint8, device="cuda"), 400), flush=True)
    env.close()
    big = q.BreakoutEnvironment(n_envs=65536, seed=2, replay_capacity=65536 * 16)
    ab = torch.randint(0, 3, (1, 65536), dtype=torch.uint8, device="cuda")
    print("single step, 65536 envs: %.2f us" % timed(lambda c: big.step_device(ab.data_ptr(), 1, None, None, s), 200), flush=True)
    big.close()


if __name__ == "__main__":
    main()
