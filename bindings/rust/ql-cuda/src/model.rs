//! The inference half of `DeepQLearningModel` (ql-with-tensorflow/src/ml_model/model.rs:29-77) on the library's tensor-core
//! Q-network: `predict_action` (:39-42) and `batch_predict_max_future_reward` (:44-47) take the state handles the
//! environment hands out (live or taken from the replay FIFO); the network reads the u8 frames in the HBM frame ring directly.
//! `train` (:60-65) stays with the caller's model (e.g. `QLearningTensorflowModel`), which passes updated weights to
//! `set_weights` — ten f32 slices in the Keras layouts of
//! python_model/create_ql_model_breakout_84x84x4_3_32.py:17-33. Source only (no Rust toolchain in the build image).
use std::rc::Rc;

use anyhow::Result;
use ql::prelude::{Action, QlError};

use crate::env::{BreakoutAction, CudaBreakoutEnvironment, CudaBreakoutState, Handle};
use crate::{check, ffi};

/// Weights in the Keras layouts: conv kernels `[kx][ky][cin][cout]` (kx runs along frame x), dense kernels `[in][out]`.
pub struct QNetWeights<'a> {
    pub conv1_kernel: &'a [f32], pub conv1_bias: &'a [f32],     // 8*8*4*32, 32
    pub conv2_kernel: &'a [f32], pub conv2_bias: &'a [f32],     // 4*4*32*64, 64
    pub conv3_kernel: &'a [f32], pub conv3_bias: &'a [f32],     // 3*3*64*64, 64
    pub dense1_kernel: &'a [f32], pub dense1_bias: &'a [f32],   // 3136*512, 512
    pub dense2_kernel: &'a [f32], pub dense2_bias: &'a [f32],   // 512*3, 3
}

impl<'a> QNetWeights<'a> {
    fn raw(&self) -> Result<ffi::qlc_qnet_weights> {
        let expect = [8 * 8 * 4 * 32, 32, 4 * 4 * 32 * 64, 64, 3 * 3 * 64 * 64, 64, 3136 * 512, 512, 512 * 3, 3];
        let got = [self.conv1_kernel.len(), self.conv1_bias.len(), self.conv2_kernel.len(), self.conv2_bias.len(), self.conv3_kernel.len(),
                   self.conv3_bias.len(), self.dense1_kernel.len(), self.dense1_bias.len(), self.dense2_kernel.len(), self.dense2_bias.len()];
        if expect != got {
            Err(QlError("weight slice lengths do not match the 84x84x4 -> 3 model".to_string()))?
        }
        Ok(ffi::qlc_qnet_weights {
            conv1_kernel: self.conv1_kernel.as_ptr(), conv1_bias: self.conv1_bias.as_ptr(),
            conv2_kernel: self.conv2_kernel.as_ptr(), conv2_bias: self.conv2_bias.as_ptr(),
            conv3_kernel: self.conv3_kernel.as_ptr(), conv3_bias: self.conv3_bias.as_ptr(),
            dense1_kernel: self.dense1_kernel.as_ptr(), dense1_bias: self.dense1_bias.as_ptr(),
            dense2_kernel: self.dense2_kernel.as_ptr(), dense2_bias: self.dense2_bias.as_ptr(),
        })
    }
}

pub struct TensorCoreQModel {
    env: Rc<Handle>,
    qnet: *mut ffi::qlc_qnet,
}

impl TensorCoreQModel {
    pub fn new(env: &CudaBreakoutEnvironment, weights: &QNetWeights) -> Result<Self> {
        let h = env.handle();
        let raw = weights.raw()?;
        let mut qnet: *mut ffi::qlc_qnet = std::ptr::null_mut();
        check(unsafe { ffi::qlc_qnet_create(h.0, &raw, &mut qnet) })?;
        Ok(Self { env: h, qnet })
    }

    pub fn set_weights(&self, weights: &QNetWeights) -> Result<()> {
        let raw = weights.raw()?;
        check(unsafe { ffi::qlc_qnet_set_weights(self.qnet, &raw) })
    }

    /// A handle names the observation after `time` steps. The network addresses replay rows: that observation is `state_next`
    /// of the transition taken at `time - 1` (k >= 1), or `state` of the one taken at `time` (k = 0: right after a reset).
    /// Returns (logical replay index, which) — `None` for the live observation, which has no transition after it yet.
    fn locate(&self, s: &CudaBreakoutState) -> Result<Option<(u32, i32)>> {
        if !Rc::ptr_eq(&s.env, &self.env) {
            Err(QlError("state of another environment".to_string()))?
        }
        let now = self.env.time();
        if s.time == now && s.k == 0 {
            return Ok(None);
        }
        let (mut cap, mut len) = (0u64, 0u64);
        check(unsafe { ffi::qlc_replay_capacity(self.env.0, &mut cap) })?;
        check(unsafe { ffi::qlc_replay_len(self.env.0, &mut len) })?;
        let oldest = now.saturating_sub(cap);
        let (t, which) = if s.k >= 1 { (s.time - 1, 1) } else { (s.time, 0) };
        if t < oldest || t - oldest >= len {
            Err(QlError("stale state handle: its frames have left the frame ring".to_string()))?
        }
        Ok(Some(((t - oldest) as u32, which)))
    }

    /// Q-values and greedy action of one state handle.
    pub fn q_values(&self, state: &CudaBreakoutState) -> Result<([f32; 3], BreakoutAction)> {
        let (mut q, mut a) = ([0f32; 3], 0u8);
        let rc = match self.locate(state)? {
            None => unsafe { ffi::qlc_qnet_forward_host(self.qnet, std::ptr::null(), 1, 0, q.as_mut_ptr(), &mut a, std::ptr::null_mut()) },
            Some((i, which)) => unsafe { ffi::qlc_qnet_forward_host(self.qnet, &i, 1, which, q.as_mut_ptr(), &mut a, std::ptr::null_mut()) },
        };
        check(rc)?;
        Ok((q, BreakoutAction::try_from_numeric(a)?))
    }

    /// `DeepQLearningModel::predict_action` (model.rs:39-42)
    pub fn predict_action(&self, state: &CudaBreakoutState) -> BreakoutAction { self.q_values(state).expect("qlc_qnet_forward_host").1 }

    /// `DeepQLearningModel::batch_predict_max_future_reward` (model.rs:44-47) for handles out of the replay FIFO.
    pub fn batch_predict_max_future_reward<const N: usize>(&self, states: [&Rc<CudaBreakoutState>; N]) -> [f32; N] {
        let mut out = [0f32; N];
        // one forward pass per kind of row (state_next rows: the usual case; state rows: handles taken right after a reset)
        for which in [1i32, 0i32] {
            let mut idx = Vec::with_capacity(N);
            let mut pos = Vec::with_capacity(N);
            for (p, s) in states.iter().enumerate() {
                match self.locate(s).expect("state handle") {
                    Some((i, w)) if w == which => { idx.push(i); pos.push(p); }
                    Some(_) => {}
                    None => {
                        if which == 1 {
                            out[p] = self.q_values(s).expect("qlc_qnet_forward_host").0.iter().cloned().fold(f32::MIN, f32::max);
                        }
                    }
                }
            }
            if idx.is_empty() {
                continue;
            }
            let mut max_q = vec![0f32; idx.len()];
            check(unsafe {
                ffi::qlc_qnet_forward_host(self.qnet, idx.as_ptr(), idx.len() as u32, which, std::ptr::null_mut(), std::ptr::null_mut(), max_q.as_mut_ptr())
            })
            .expect("qlc_qnet_forward_host");
            for (p, v) in pos.into_iter().zip(max_q) {
                out[p] = v;
            }
        }
        out
    }
}

impl Drop for TensorCoreQModel {
    fn drop(&mut self) {
        unsafe { ffi::qlc_qnet_destroy(self.qnet) };
    }
}
