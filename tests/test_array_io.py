"""Framework-neutral zero-copy hand-off (q-learning_b200/array_io.py): CUDA Array Interface and DLPack, out of the library (views on the
frame ring and the SoA state) and into arrays a framework owns (ArraySampler, observe_into). CPU: protocol handling on fake
pointers; GPU: torch as the consumer / producer on both protocols, against the oracle and the torch_io path."""
import importlib

import numpy as np
import pytest


@pytest.fixture(scope="module")
def aio(qlb):
    return importlib.import_module("q-learning_b200.array_io")


class _FakeCai:
    def __init__(self, ptr, shape, typestr, strides=None):
        self.__cuda_array_interface__ = {"shape": shape, "typestr": typestr, "data": (ptr, False), "version": 3, "strides": strides}


def test_protocol_parsing_on_fake_pointers(qlb, aio):
    v = aio.DeviceView(0x7F0000001000, (5, 3, 84, 84), np.uint8, 2, owner=None)
    assert v.__cuda_array_interface__["shape"] == (5, 3, 84, 84) and v.__cuda_array_interface__["typestr"] == "|u1" and v.nbytes == 5 * 3 * 7056
    assert v.__dlpack_device__() == (2, 2)
    # DLPack capsule round trip through our own reader (a framework would consume it the same way)
    ptr, shape, dt, dev = aio.describe(v.__dlpack__())
    assert (ptr, shape, dt, dev) == (0x7F0000001000, (5, 3, 84, 84), np.dtype(np.uint8), 2)
    assert aio.describe(v)[:3] == (0x7F0000001000, (5, 3, 84, 84), np.dtype(np.uint8))
    f32 = aio.DeviceView(0x7F0000002000, (32, 84, 84, 4), np.float32, 0, owner=None)
    assert aio.device_pointer(f32, (32 * 84 * 84 * 4,), np.float32, 0, "state", 16) == 0x7F0000002000
    with pytest.raises(qlb.QlError, match="dtype"):
        aio.device_pointer(f32, None, np.uint8)
    with pytest.raises(qlb.QlError, match="shape"):
        aio.device_pointer(f32, (31, 84, 84, 4), np.float32)
    with pytest.raises(qlb.QlError, match="cuda:0"):
        aio.device_pointer(aio.DeviceView(0x1000, (4,), np.float32, 0, None).__dlpack__(), (4,), np.float32, device=1)
    with pytest.raises(qlb.QlError, match="aligned"):
        aio.device_pointer(_FakeCai(0x1004, (8,), "<f4"), (8,), np.float32, align=16)
    with pytest.raises(qlb.QlError, match="contiguous"):
        aio.describe(_FakeCai(0x1000, (4, 4), "<f4", strides=(32, 4)))
    assert aio.describe(_FakeCai(0x1000, (4, 4), "<f4", strides=(16, 4)))[1] == (4, 4)
    with pytest.raises(qlb.QlError, match="not in CUDA memory"):
        aio.describe(np.zeros(4, dtype=np.float32))              # numpy speaks DLPack, but it is host memory
    with pytest.raises(qlb.QlError, match="neither"):
        aio.describe(object())


@pytest.mark.gpu
def test_views_and_foreign_arrays_on_the_gpu(qlb, O, aio):
    torch = pytest.importorskip("torch")
    tio = importlib.import_module("q-learning_b200.torch_io")
    n, seed, steps = 48, 17, 30
    env = qlb.BreakoutEnvironment(n_envs=n, seed=seed, replay_capacity=n * 32)
    rb = qlb.ReplayBuffer(env)
    ora = O.VecEnv(n, seed=seed, replay_capacity=n * 32)
    acts = O.synthetic_actions(seed, 0, n, 0, steps)
    env.step_many(acts)
    for a in acts:
        ora.step(a)
    views = aio.device_views(env)
    assert views["_time"] == steps
    # OUT, DLPack: the newest frame of every env, read where the step kernel wrote it, equals slot (k-1) mod 4 of the oracle's stack
    frames = torch.from_dlpack(views["frames"])
    assert frames.shape == (views["_time_slots"], n, 84, 84) and frames.dtype == torch.uint8 and frames.data_ptr() == views["frames"].ptr
    newest = frames[(steps - 1) % views["_time_slots"]].cpu().numpy()
    stack = ora.obs_u8()                                           # [n][slot][y][x]
    k = ora.state()["episode_step"]
    for e in range(n):
        if k[e] > 0:
            assert np.array_equal(newest[e], stack[e, (k[e] - 1) % 4]), e
    # OUT, CUDA Array Interface: the SoA state
    st = ora.state()
    assert np.array_equal(torch.as_tensor(views["ball_cx"], device="cuda").cpu().numpy().view(np.uint32), st["ball_cx"].view(np.uint32))
    assert np.array_equal(torch.as_tensor(views["score"], device="cuda").cpu().numpy().view(np.uint32), st["score"])
    # INTO foreign arrays: torch tensors handed over as CUDA-array-interface objects and as DLPack capsules
    per = 4 * 84 * 84
    batch, nb = 32, 3
    want = tio.DeviceSampler(rb, batch, nb, qlb.LAYOUT_F32_BXYH).sample(7)
    torch.cuda.synchronize()
    s1 = torch.zeros((batch * nb, 84, 84, 4), dtype=torch.float32, device="cuda"); s2 = torch.zeros_like(s1)
    idx = torch.zeros((batch * nb,), dtype=torch.int32, device="cuda"); r = torch.zeros((batch * nb,), dtype=torch.float32, device="cuda")
    a = torch.zeros((batch * nb,), dtype=torch.uint8, device="cuda"); d = torch.zeros_like(a)
    cap = torch.utils.dlpack.to_dlpack(s2)                         # what tf.experimental.dlpack.to_dlpack hands out
    aio.ArraySampler(rb, batch, nb, s1, cap, qlb.LAYOUT_F32_BXYH, indices=idx, reward=r, action=a, done=d).sample(7)
    torch.cuda.synchronize()
    assert torch.equal(s1.view(nb, batch, 84, 84, 4), want.state) and torch.equal(s2.view(nb, batch, 84, 84, 4), want.state_next)
    assert torch.equal(idx.view(nb, batch), want.indices) and torch.equal(r.view(nb, batch), want.reward) and torch.equal(a.view(nb, batch), want.action)
    obs = torch.zeros((n, 4, 84, 84), dtype=torch.uint8, device="cuda")
    aio.observe_into(env, obs, qlb.LAYOUT_U8_BHYX)
    torch.cuda.synchronize()
    assert np.array_equal(obs.cpu().numpy(), ora.obs_u8())
    with pytest.raises(qlb.QlError, match="dtype"):
        aio.observe_into(env, obs, qlb.LAYOUT_F32_BXYH)
    with pytest.raises(qlb.QlError, match="contiguous"):
        aio.observe_into(env, torch.zeros((n, 4, 84, 168), dtype=torch.uint8, device="cuda")[..., ::2], qlb.LAYOUT_U8_BHYX)
    del frames
    env.close(); ora.close()
