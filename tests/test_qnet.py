"""Q-network forward on the tensor cores. Floating point: compared against an fp32 reference of the same
op on bf16-rounded operands; tolerance stated per test (bf16 output rounding = 2^-9 relative)."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _bf16_round(x):
    """round-to-nearest-even to bfloat16, returned as f32"""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


@pytest.mark.parametrize("m,n,k,relu", [(128, 32, 64, False), (256, 64, 256, True), (300, 64, 576, True), (128, 128, 128, False),
                                        (77, 256, 192, True), (200, 512, 3136, True), (1000, 32, 256, True)])
def test_tcgen05_gemm_matches_fp32_reference(qlb, m, n, k, relu):
    rng = np.random.default_rng(m * 7 + n)
    a = rng.standard_normal((m, k)).astype(np.float32)
    w = (rng.standard_normal((n, k)) / np.sqrt(k)).astype(np.float32)
    bias = rng.standard_normal(n).astype(np.float32)
    got = qlb.debug_gemm_bf16(a, w, bias, relu)
    ref = _bf16_round(a).astype(np.float64) @ _bf16_round(w).astype(np.float64).T + bias
    if relu:
        ref = np.maximum(ref, 0)
    # tolerance: bf16 output (rel 2^-8) + f32 accumulation order; absolute floor for values near zero
    err = np.abs(got - ref)
    assert np.all(err <= 1e-2 * np.abs(ref) + 2e-2), "max abs err %.4g at %s" % (err.max(), np.unravel_index(err.argmax(), err.shape))


def _random_weights(qlb, seed):
    rng = np.random.default_rng(seed)
    w = {}
    for name, shape in qlb.QNET_SHAPES.items():
        if name.endswith("kernel"):
            fan_in = int(np.prod(shape[:-1])); fan_out = shape[-1]
            lim = np.sqrt(6.0 / (fan_in + fan_out))                      # Keras default glorot_uniform
            w[name] = rng.uniform(-lim, lim, size=shape).astype(np.float32)
        else:
            w[name] = rng.uniform(-0.1, 0.1, size=shape).astype(np.float32)
    return w


def _torch_reference(torch, w, obs_bxyh, quantise):
    """fp32 forward of the Keras model on [B, 84(x), 84(y), 4(slot)]; quantise=True rounds weights and the activations
    between layers to bf16 at the same points as the tensor-core path (accumulation stays fp32)."""
    F = torch.nn.functional
    bf = (lambda t: t.to(torch.bfloat16).to(torch.float32)) if quantise else (lambda t: t)
    x = torch.from_numpy(obs_bxyh).permute(0, 3, 1, 2).contiguous()       # NCHW with H = x, W = y
    for i in (1, 2, 3):
        k = bf(torch.from_numpy(w["conv%d_kernel" % i]).permute(3, 2, 0, 1).contiguous())
        x = bf(torch.relu(F.conv2d(x, k, torch.from_numpy(w["conv%d_bias" % i]), stride={1: 4, 2: 2, 3: 1}[i])))
    x = x.permute(0, 2, 3, 1).reshape(x.shape[0], -1)                        # Flatten of [7][7][64]
    x = bf(torch.relu(x @ bf(torch.from_numpy(w["dense1_kernel"])) + torch.from_numpy(w["dense1_bias"])))
    q = x @ torch.from_numpy(w["dense2_kernel"]) + torch.from_numpy(w["dense2_bias"])
    return q.numpy()


def test_qnet_forward_matches_torch_fp32(qlb, O):
    """predict_action for every env and batch_predict_max_future_reward for replay samples vs an fp32 torch restatement of
    the Keras model. Tolerances: vs the reference quantised at the same points 1e-2 * max|Q| (accumulation order + bf16
    ties); vs pure fp32 4e-2 * max|Q| (bf16 operands through five layers)."""
    torch = pytest.importorskip("torch")
    n, seed = 300, 5
    env = qlb.BreakoutEnvironment(n_envs=n, seed=seed, replay_capacity=n * 16)
    rb = qlb.ReplayBuffer(env)
    rng = np.random.default_rng(1)
    env.step_many(rng.integers(0, 3, size=(37, n), dtype=np.uint8))
    env.reset(mask=(np.arange(n) % 50 == 0).astype(np.uint8))              # some envs with an empty stack
    env.step_many(rng.integers(0, 3, size=(2, n), dtype=np.uint8))         # ... and partially filled stacks
    w = _random_weights(qlb, 3)
    net = qlb.QNetwork(env, w)
    q, action, max_q = net.forward()
    obs = env.obs(qlb.LAYOUT_F32_BXYH)
    ref_q = _torch_reference(torch, w, obs, quantise=True)
    ref_f = _torch_reference(torch, w, obs, quantise=False)
    scale = float(np.abs(ref_f).max())
    assert scale > 1.0, "degenerate test: Q-values too small (%g)" % scale
    assert np.abs(q - ref_q).max() <= 1e-2 * scale, np.abs(q - ref_q).max() / scale
    assert np.abs(q - ref_f).max() <= 4e-2 * scale, np.abs(q - ref_f).max() / scale
    assert np.array_equal(max_q, q.max(axis=1)) and np.array_equal(action, q.argmax(axis=1).astype(np.uint8))
    srt = np.sort(ref_f, axis=1)
    clear = (srt[:, 2] - srt[:, 1]) > 8e-2 * scale                           # envs whose best action is unambiguous
    assert clear.mean() > 0.3 and np.array_equal(action[clear], ref_f.argmax(axis=1)[clear].astype(np.uint8))
    # replay samples: state_next of sampled transitions (batch_predict_max_future_reward) and state
    idx = rb.generate_distinct_random_ids(64, 0)
    smp = rb.get_many(idx, qlb.LAYOUT_F32_BXYH)
    for which, batch in ((1, smp.state_next), (0, smp.state)):
        qs, _, ms = net.forward(idx, which)
        rs = _torch_reference(torch, w, batch, quantise=True)
        assert np.abs(qs - rs).max() <= 1e-2 * scale
        assert np.array_equal(ms, qs.max(axis=1))
    # new weights take effect
    w2 = _random_weights(qlb, 4)
    net.set_weights(w2)
    q2, _, _ = net.forward()
    assert np.abs(q2 - _torch_reference(torch, w2, obs, quantise=True)).max() <= 1e-2 * scale
    net.close(); env.close()


@pytest.mark.parametrize("n", [1, 7, 2000])
def test_qnet_forward_batch_sizes(qlb, n):
    """Edge batch sizes of the shifted-window pipeline: fewer items than one conv batch (conv2 packs 2, conv3 3, the dense
    layer 128 items per tile) and enough items that every persistent CTA runs many iterations (mbarrier phases wrap, the
    TMEM accumulator sets and the raw / plane stage rings are reused). Same tolerance as above: 1e-2 * max|Q| against the
    fp32 reference quantised to bf16 at the same points."""
    torch = pytest.importorskip("torch")
    env = qlb.BreakoutEnvironment(n_envs=n, seed=11, replay_capacity=n * 8)
    rng = np.random.default_rng(n)
    env.step_many(rng.integers(0, 3, size=(5, n), dtype=np.uint8))
    w = _random_weights(qlb, 9)
    net = qlb.QNetwork(env, w)
    q, action, max_q = net.forward()
    obs = env.obs(qlb.LAYOUT_F32_BXYH)
    ref = _torch_reference(torch, w, obs, quantise=True)
    scale = float(np.abs(ref).max())
    assert np.abs(q - ref).max() <= 1e-2 * scale, np.abs(q - ref).max() / scale
    assert np.array_equal(max_q, q.max(axis=1)) and np.array_equal(action, q.argmax(axis=1).astype(np.uint8))
    q2, _, _ = net.forward()                                               # a second pass over reused buffers gives the same bits
    assert np.array_equal(q, q2)
    net.close(); env.close()


def test_closed_actor_loop_is_bitwise_reproducible(qlb):
    """Q-network -> greedy action -> env step, all on the device (no host in the loop): two runs from the same seed and
    weights end in the same bits (state, pixels, reward sum, action histogram). The fused head adds its per-tile partials
    in tile order, so nothing in the loop depends on scheduling."""
    torch = pytest.importorskip("torch")

    def run():
        n, iters = 300, 120
        env = qlb.BreakoutEnvironment(n_envs=n, seed=9, replay_capacity=n * 32)
        net = qlb.QNetwork(env, _random_weights(qlb, 5))
        s = torch.cuda.current_stream().cuda_stream
        acts = torch.empty((1, n), dtype=torch.uint8, device="cuda")
        rew = torch.empty((1, n), dtype=torch.float32, device="cuda")
        total, hist = torch.zeros((), device="cuda"), torch.zeros(3, device="cuda")
        for _ in range(iters):
            net.forward_device(None, n, 0, None, acts.data_ptr(), None, s)
            env.step_device(acts.data_ptr(), 1, rew.data_ptr(), None, s)
            total += rew.sum(); hist += torch.bincount(acts[0].long(), minlength=3).float()
        torch.cuda.synchronize()
        st = env.read_state()
        out = ({k: st[k].copy() for k in st}, env.obs(qlb.LAYOUT_U8_BHYX), float(total), hist.tolist())
        net.close(); env.close()
        return out
    a, b = run(), run()
    assert a[2] == b[2] and a[3] == b[3] and np.array_equal(a[1], b[1])
    for k in a[0]:
        assert np.array_equal(a[0][k], b[0][k]), k
    assert sum(a[3]) == 300 * 120 and max(a[3]) < 300 * 120      # the policy is not constant
