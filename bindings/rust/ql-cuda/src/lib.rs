//! ql-cuda: B200-native drop-in for the Breakout environment, its state and the replay memory of
//! bitmagier/q-learning. `CudaBreakoutEnvironment: ql::prelude::Environment`, `CudaBreakoutState` (cheap handle,
//! `ToMultiDimArray` behind the `tensor` feature) and `ReplayBuffer` with the method set of
//! ql-with-tensorflow/src/learn/replay_buffer.rs, so `SelfDrivingQLearner` runs unchanged.
pub mod ffi;

mod env;
mod model;
mod replay;

pub use env::{BreakoutAction, CudaBreakoutEnvironment, CudaBreakoutState, StateKind};
pub use model::{QNetWeights, TensorCoreQModel};
pub use replay::{generate_distinct_random_ids, BufferSample, ReplayBuffer};

use ql::prelude::QlError;
use std::ffi::CStr;

pub(crate) fn check(rc: i32) -> anyhow::Result<()> {
    if rc == ffi::QLC_OK {
        return Ok(());
    }
    let msg = unsafe { CStr::from_ptr(ffi::qlc_last_error_string()) }.to_string_lossy().into_owned();
    Err(QlError(msg).into())
}
