"""Zero-copy hand-off of the hot path's outputs to a torch consumer: the gather kernels write straight
into CUDA tensors that the model reads — no host round trip. torch is plumbing here (device memory + streams); every byte
is produced by the kernels in csrc/."""
import torch

from . import FRAME_H, FRAME_W, LAYOUT_F32_BXYH, LAYOUT_U8_BHYX, NUM_FRAMES, QlError


def _stream():
    return torch.cuda.current_stream().cuda_stream


class DeviceSampler:
    """Reusable device buffers for `n_batches` minibatches of `batch` transitions; `sample()` = one
    `qlc_replay_sample` + one `qlc_replay_gather` on torch's current stream, returns views on the buffers."""

    def __init__(self, replay, batch, n_batches=1, layout=LAYOUT_F32_BXYH, device=None):
        self.replay, self.batch, self.n_batches, self.layout = replay, batch, n_batches, layout
        dev = torch.device("cuda", replay._env.device if device is None else device)
        n = batch * n_batches
        if layout == LAYOUT_F32_BXYH:
            shape, dt = (n_batches, batch, FRAME_W, FRAME_H, NUM_FRAMES), torch.float32
        elif layout == LAYOUT_U8_BHYX:
            shape, dt = (n_batches, batch, NUM_FRAMES, FRAME_H, FRAME_W), torch.uint8
        else:
            raise QlError("unknown layout")
        self.indices = torch.empty((n_batches, batch), dtype=torch.int32, device=dev)
        self.state = torch.empty(shape, dtype=dt, device=dev)
        self.state_next = torch.empty(shape, dtype=dt, device=dev)
        self.reward = torch.empty((n_batches, batch), dtype=torch.float32, device=dev)
        self.action = torch.empty((n_batches, batch), dtype=torch.uint8, device=dev)
        self.done = torch.empty((n_batches, batch), dtype=torch.uint8, device=dev)

    def sample(self, call_index):
        s = _stream()
        self.replay.sample_device(self.batch, self.n_batches, call_index, self.indices.data_ptr(), s)
        self.replay.gather_device(self.indices.data_ptr(), self.batch * self.n_batches, self.layout, self.state.data_ptr(),
                                  self.state_next.data_ptr(), self.reward.data_ptr(), self.action.data_ptr(), self.done.data_ptr(), s)
        return self


def observe(env, layout=LAYOUT_F32_BXYH, out=None):
    """Current observation stacks of all envs as a CUDA tensor (Environment::state + ToMultiDimArray, no host copy)."""
    dev = torch.device("cuda", env.device)
    if layout == LAYOUT_F32_BXYH:
        shape, dt = (env.n_envs, FRAME_W, FRAME_H, NUM_FRAMES), torch.float32
    else:
        shape, dt = (env.n_envs, NUM_FRAMES, FRAME_H, FRAME_W), torch.uint8
    if out is None:
        out = torch.empty(shape, dtype=dt, device=dev)
    env.obs_device(layout, out.data_ptr(), _stream())
    return out


def step(env, actions, reward=None, done=None):
    """One launch advancing all envs by actions.shape[0] steps; actions / reward / done are CUDA tensors [K][N]."""
    if actions.dtype != torch.uint8 or actions.dim() != 2 or actions.shape[1] != env.n_envs or not actions.is_contiguous():
        raise QlError("actions must be a contiguous u8 CUDA tensor [n_steps][n_envs]")
    k = actions.shape[0]
    if reward is None:
        reward = torch.empty((k, env.n_envs), dtype=torch.float32, device=actions.device)
    if done is None:
        done = torch.empty((k, env.n_envs), dtype=torch.uint8, device=actions.device)
    env.step_device(actions.data_ptr(), k, reward.data_ptr(), done.data_ptr(), _stream())
    return reward, done
