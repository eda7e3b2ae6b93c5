"""Framework-neutral zero-copy hand-off — no torch needed.

Two directions, two protocols:

* OUT of the library: `DeviceView` wraps device memory the library owns (the HBM frame ring, the transition records, the SoA
  env state: `device_views(env)`) and exposes it through the CUDA Array Interface v3 (`__cuda_array_interface__`: CuPy, Numba,
  PyTorch `torch.as_tensor`, JAX) and through DLPack (`__dlpack__` / `__dlpack_device__`: `torch.from_dlpack`,
  `tf.experimental.dlpack.from_dlpack(view.__dlpack__())`, `jax.dlpack.from_dlpack`, `cupy.from_dlpack`). A consumer reads the frames
  where the step kernel wrote them. The views are valid while the env lives and are read-only BY CONTRACT (the flag itself stays off: PyTorch refuses flagged interfaces).
* INTO buffers a framework owns: `device_pointer(obj)` accepts anything that speaks either protocol (or is a DLPack capsule, e.g.
  `tf.experimental.dlpack.to_dlpack(t)`), checks dtype / shape / C-contiguity / device and returns the raw pointer the C ABI takes;
  `ArraySampler` and `observe_into` are the torch-free twins of `torch_io.DeviceSampler` / `torch_io.observe`: the gather kernels
  write straight into the framework's arrays (the reference's `ToMultiDimArray` tensors, `breakout_environment.rs:42-77`).

Synchronisation is the caller's: pass the `cudaStream_t` the consumer works on (`stream=`), or synchronise that stream yourself.
"""
import ctypes as C

import numpy as np

from . import FRAME_H, FRAME_W, LAYOUT_F32_BXYH, LAYOUT_U8_BHYX, LAYOUT_U8_BXYH, NUM_FRAMES, QlError

# ---- DLPack structures (dlpack.h, v0.8 ABI: the unversioned DLManagedTensor every framework accepts) ----
_kDLCUDA, _kDLCUDAHost = 2, 3
_kDLInt, _kDLUInt, _kDLFloat = 0, 1, 2


class _DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int32), ("device_id", C.c_int32)]


class _DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class _DLTensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("device", _DLDevice), ("ndim", C.c_int32), ("dtype", _DLDataType),
                ("shape", C.POINTER(C.c_int64)), ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]


class _DLManagedTensor(C.Structure):
    pass


_DELETER = C.CFUNCTYPE(None, C.POINTER(_DLManagedTensor))
_DLManagedTensor._fields_ = [("dl_tensor", _DLTensor), ("manager_ctx", C.c_void_p), ("deleter", _DELETER)]

_api = C.pythonapi
_api.PyCapsule_New.restype = C.py_object
_api.PyCapsule_New.argtypes = [C.c_void_p, C.c_char_p, C.c_void_p]
_api.PyCapsule_GetPointer.restype = C.c_void_p
_api.PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
_api.PyCapsule_IsValid.restype = C.c_int
_api.PyCapsule_IsValid.argtypes = [C.py_object, C.c_char_p]

_DTYPE_TO_DL = {"u1": (_kDLUInt, 8), "u4": (_kDLUInt, 32), "u8": (_kDLUInt, 64), "i4": (_kDLInt, 32), "f4": (_kDLFloat, 32), "f8": (_kDLFloat, 64)}
_DL_TO_DTYPE = {v: k for k, v in _DTYPE_TO_DL.items()}
_LIVE = {}       # id -> (managed tensor, shape array, owner): kept alive until the consumer calls the deleter


@_DELETER
def _deleter(ptr):
    _LIVE.pop(C.addressof(ptr.contents), None)


class DeviceView:
    """A typed, shaped, C-contiguous window on device memory owned by the library. `owner` (the env) is kept alive by the view."""

    def __init__(self, ptr, shape, dtype, device, owner, readonly=False):      # (PyTorch refuses interfaces flagged read-only)
        self.ptr, self.shape, self.dtype, self.device, self._owner, self.readonly = int(ptr), tuple(int(s) for s in shape), np.dtype(dtype), int(device), owner, readonly

    @property
    def nbytes(self):
        return int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize

    @property
    def __cuda_array_interface__(self):
        return {"shape": self.shape, "typestr": self.dtype.str, "data": (self.ptr, self.readonly), "version": 3, "strides": None, "stream": None}

    def __dlpack_device__(self):
        return (_kDLCUDA, self.device)

    def __dlpack__(self, stream=None, **_):
        code_bits = _DTYPE_TO_DL.get(self.dtype.str[1:])
        if code_bits is None:
            raise QlError("dtype %s has no DLPack code here" % self.dtype)
        shape = (C.c_int64 * len(self.shape))(*self.shape)
        m = _DLManagedTensor()
        m.dl_tensor.data = self.ptr
        m.dl_tensor.device = _DLDevice(_kDLCUDA, self.device)
        m.dl_tensor.ndim = len(self.shape)
        m.dl_tensor.dtype = _DLDataType(code_bits[0], code_bits[1], 1)
        m.dl_tensor.shape = C.cast(shape, C.POINTER(C.c_int64))
        m.dl_tensor.strides = None                      # NULL = compact row-major
        m.dl_tensor.byte_offset = 0
        m.manager_ctx = None
        m.deleter = _deleter
        _LIVE[C.addressof(m)] = (m, shape, self)
        return _api.PyCapsule_New(C.addressof(m), b"dltensor", None)


def device_views(env):
    """dict of DeviceView on everything the env keeps in HBM: 'frames' u8 [time_slots][n_envs][84][84] (the replay
    frame ring, time-major; slot of time t = t mod time_slots), 'records' u32 [time_slots][n_envs] (packed transitions), and the
    SoA state arrays [n_envs]. Plus 'time' (env-steps taken) and 'time_slots' as plain ints under the same names with a leading '_'."""
    v = env.state_view()
    n, ts, dev = int(v.n_envs), int(v.time_slots), env.device
    out = {"frames": DeviceView(v.frames, (ts, n, FRAME_H, FRAME_W), np.uint8, dev, env),
           "records": DeviceView(v.records, (ts, n), np.uint32, dev, env)}
    for name in ("ball_cx", "ball_cy", "ball_dx", "ball_dy", "pad_min_x", "pad_max_x", "pad_speed"):
        out[name] = DeviceView(getattr(v, name), (n,), np.float32, dev, env)
    out["bricks"] = DeviceView(v.bricks, (n,), np.uint64, dev, env)
    for name in ("score", "episode_step", "episode", "err"):
        out[name] = DeviceView(getattr(v, name), (n,), np.uint32, dev, env)
    out["finished"] = DeviceView(v.finished, (n,), np.uint8, dev, env)
    out["_time"], out["_time_slots"] = int(v.time), ts
    return out


def _from_dl(t, what):
    key = (t.dtype.code, t.dtype.bits)
    if t.dtype.lanes != 1 or key not in _DL_TO_DTYPE:
        raise QlError("%s: unsupported DLPack dtype (code %d, %d bits, %d lanes)" % (what, t.dtype.code, t.dtype.bits, t.dtype.lanes))
    shape = tuple(int(t.shape[i]) for i in range(t.ndim))
    if t.strides:
        stride, want = 1, []
        for s in reversed(shape):
            want.append(stride); stride *= s
        got = [int(t.strides[i]) for i in range(t.ndim)]
        if any(g != w for g, w, s in zip(got, reversed(want), shape) if s > 1):
            raise QlError("%s must be C-contiguous (strides %s)" % (what, got))
    if t.device.device_type not in (_kDLCUDA, _kDLCUDAHost):
        raise QlError("%s lives on DLPack device type %d, not in CUDA memory" % (what, t.device.device_type))
    return (int(t.data or 0) + int(t.byte_offset), shape, np.dtype(_DL_TO_DTYPE[key]), int(t.device.device_id))


def describe(obj, what="array"):
    """(pointer, shape, dtype, device ordinal or None) of a CUDA array given by any of: DLPack capsule, `__cuda_array_interface__`,
    `__dlpack__`. Raises QlError for non-contiguous or non-CUDA arrays."""
    if _api.PyCapsule_IsValid(obj, b"dltensor") if type(obj).__name__ == "PyCapsule" else False:
        m = C.cast(_api.PyCapsule_GetPointer(obj, b"dltensor"), C.POINTER(_DLManagedTensor)).contents
        return _from_dl(m.dl_tensor, what)      # the capsule stays unconsumed: its producer keeps the memory alive
    cai = getattr(obj, "__cuda_array_interface__", None)
    if cai is not None:
        shape, dt = tuple(int(s) for s in cai["shape"]), np.dtype(cai["typestr"])
        strides = cai.get("strides")
        if strides is not None:
            stride = dt.itemsize
            for s, st in zip(reversed(shape), reversed(tuple(strides))):
                if s > 1 and int(st) != stride:
                    raise QlError("%s must be C-contiguous (strides %s)" % (what, tuple(strides)))
                stride *= s
        return (int(cai["data"][0]), shape, dt, None)
    if hasattr(obj, "__dlpack__"):
        cap = obj.__dlpack__()
        m = C.cast(_api.PyCapsule_GetPointer(cap, b"dltensor"), C.POINTER(_DLManagedTensor)).contents
        res = _from_dl(m.dl_tensor, what)
        if m.deleter:
            m.deleter(C.pointer(m))             # we only looked: hand the tensor back; `obj` itself keeps the memory alive
        return res
    raise QlError("%s speaks neither the CUDA Array Interface nor DLPack" % what)


def device_pointer(obj, shape=None, dtype=None, device=None, what="array", align=1):
    """Raw device pointer of `obj` for the C ABI after checking dtype, element count (`shape` may be any shape with the same number
    of elements), contiguity and, where the protocol tells, the device ordinal."""
    ptr, shp, dt, dev = describe(obj, what)
    if dtype is not None and dt != np.dtype(dtype):
        raise QlError("%s has dtype %s, expected %s" % (what, dt, np.dtype(dtype)))
    if shape is not None and int(np.prod(shp, dtype=np.int64)) != int(np.prod(shape, dtype=np.int64)):
        raise QlError("%s has shape %s, expected %s" % (what, shp, tuple(shape)))
    if device is not None and dev is not None and dev != device:
        raise QlError("%s is on cuda:%d, the env on cuda:%d" % (what, dev, device))
    if ptr % align:
        raise QlError("%s must be %d-byte aligned" % (what, align))
    return ptr


def _stack(layout):
    if layout == LAYOUT_F32_BXYH:
        return (FRAME_W, FRAME_H, NUM_FRAMES), np.float32
    if layout == LAYOUT_U8_BXYH:
        return (FRAME_W, FRAME_H, NUM_FRAMES), np.uint8
    if layout == LAYOUT_U8_BHYX:
        return (NUM_FRAMES, FRAME_H, FRAME_W), np.uint8
    raise QlError("unknown layout")


def observe_into(env, out, layout=LAYOUT_F32_BXYH, stream=None):
    """Environment::state + ToMultiDimArray of all envs, written by the gather kernel straight into `out` (any framework's CUDA array)."""
    per, dt = _stack(layout)
    env.obs_device(layout, device_pointer(out, (env.n_envs,) + per, dt, env.device, "out", 16), stream)
    return out


class ArraySampler:
    """Torch-free twin of torch_io.DeviceSampler: ONE kernel launch (`qlc_replay_sample_gather`) draws `n_batches` minibatches of
    `batch` distinct indices and gathers state / state_next / reward / action / done into arrays the caller's framework owns.
    Pass arrays with `batch * n_batches` leading elements; `indices`, `reward`, `action`, `done` may be None."""

    def __init__(self, replay, batch, n_batches, state, state_next, layout=LAYOUT_F32_BXYH, indices=None, reward=None, action=None, done=None):
        self.replay, self.batch, self.n_batches, self.layout = replay, batch, n_batches, layout
        n, dev = batch * n_batches, replay._env.device
        per, dt = _stack(layout)
        self._keep = (state, state_next, indices, reward, action, done)
        self._state = device_pointer(state, (n,) + per, dt, dev, "state", 16) if state is not None else None
        self._next = device_pointer(state_next, (n,) + per, dt, dev, "state_next", 16) if state_next is not None else None
        if self._state is None and self._next is None:
            raise QlError("state or state_next is needed")
        idx_dt = None if indices is None else describe(indices, "indices")[2]
        if idx_dt is not None and idx_dt not in (np.dtype(np.uint32), np.dtype(np.int32)):
            raise QlError("indices must be 32-bit integers")
        self._idx = device_pointer(indices, (n,), idx_dt, dev, "indices") if indices is not None else None
        self._reward = device_pointer(reward, (n,), np.float32, dev, "reward") if reward is not None else None
        self._action = device_pointer(action, (n,), np.uint8, dev, "action") if action is not None else None
        self._done = device_pointer(done, (n,), np.uint8, dev, "done") if done is not None else None

    def sample(self, call_index, stream=None):
        self.replay.sample_gather_device(self.batch, self.n_batches, call_index, self.layout, self._idx, self._state, self._next,
                                         self._reward, self._action, self._done, stream)
        return self
