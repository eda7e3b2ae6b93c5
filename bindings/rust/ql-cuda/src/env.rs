//! Environment / Action / state (ql/src/prelude.rs:12-68, _breakout-ml/src/breakout_environment.rs:24-207).
use std::fmt::{Debug, Display, Formatter};
use std::rc::Rc;

use anyhow::Result;
use console_engine::screen::Screen;
use ql::prelude::{Action, DebugVisualizer, Environment, ModelActionType, QlError};

use crate::{check, ffi};

/// Owns the `qlc_env` (N = 1 env + frame ring + replay shard in HBM). Shared by the environment, its states and the replay buffer.
pub(crate) struct Handle(pub(crate) *mut ffi::qlc_env);

impl Handle {
    pub(crate) fn time(&self) -> u64 {
        let mut t = 0u64;
        unsafe { ffi::qlc_env_time(self.0, &mut t) };
        t
    }
}

impl Drop for Handle {
    fn drop(&mut self) {
        unsafe { ffi::qlc_env_destroy(self.0) };
    }
}

#[derive(Debug, Clone, Copy, Hash, PartialEq, Eq)]
pub enum BreakoutAction {
    None,
    Left,
    Right,
}

impl Action for BreakoutAction {
    const ACTION_SPACE: ModelActionType = 3;

    fn numeric(&self) -> ModelActionType {
        match self {
            BreakoutAction::None => 0,
            BreakoutAction::Left => 1,
            BreakoutAction::Right => 2,
        }
    }

    fn try_from_numeric(value: ModelActionType) -> Result<Self> {
        match value {
            0 => Ok(BreakoutAction::None),
            1 => Ok(BreakoutAction::Left),
            2 => Ok(BreakoutAction::Right),
            _ => Err(QlError("value out of range".to_string()))?,
        }
    }
}

impl Display for BreakoutAction {
    fn fmt(&self, f: &mut Formatter<'_>) -> std::fmt::Result { write!(f, "{:?}", self) }
}

#[derive(Clone, Copy, Debug, PartialEq, Eq)]
pub enum StateKind {
    /// the environment's observation at `time`
    Live,
    /// `state` / `state_next` of replay row `index`, valid until the env steps again
    ReplayState(u32),
    ReplayNext(u32),
}

/// BreakoutState as a cheap handle: `Clone` copies indices, the pixels stay in the HBM frame ring until tensorised.
#[derive(Clone)]
pub struct CudaBreakoutState {
    pub(crate) env: Rc<Handle>,
    pub(crate) kind: StateKind,
    pub(crate) time: u64,
    model_dims: [u64; 3],
}

impl CudaBreakoutState {
    pub(crate) fn new(env: Rc<Handle>, kind: StateKind, time: u64) -> Self {
        Self { env, kind, time, model_dims: [ffi::QLC_FRAME_W as u64, ffi::QLC_FRAME_H as u64, ffi::QLC_NUM_FRAMES as u64] }
    }
    pub fn dims(&self) -> &[u64] { &self.model_dims }

    /// `[b][x][y][slot]` f32, value = u8 as f32 (breakout_environment.rs:56-77) — one gather kernel for the whole batch.
    pub fn batch_to_f32<const N: usize>(batch: &[&Rc<Self>; N]) -> Result<Vec<f32>> {
        let per = ffi::QLC_FRAME_W * ffi::QLC_FRAME_H * ffi::QLC_NUM_FRAMES;
        let mut out = vec![0f32; N * per];
        let first = batch[0];
        match first.kind {
            StateKind::Live => {
                if first.time != first.env.time() {
                    Err(QlError("stale state handle".to_string()))?
                }
                let mut one = vec![0f32; per];
                check(unsafe { ffi::qlc_env_obs_host(first.env.0, ffi::QLC_LAYOUT_F32_BXYH, one.as_mut_ptr() as *mut _) })?;
                for b in 0..N {
                    out[b * per..(b + 1) * per].copy_from_slice(&one);
                }
            }
            StateKind::ReplayState(_) | StateKind::ReplayNext(_) => {
                let next = matches!(first.kind, StateKind::ReplayNext(_));
                let mut idx = [0u32; N];
                for (b, s) in batch.iter().enumerate() {
                    idx[b] = match (s.kind, next) {
                        (StateKind::ReplayState(i), false) | (StateKind::ReplayNext(i), true) => i,
                        _ => Err(QlError("mixed state kinds in one batch".to_string()))?,
                    };
                    if s.time != s.env.time() {
                        Err(QlError("stale replay sample".to_string()))?
                    }
                }
                let p = out.as_mut_ptr() as *mut std::os::raw::c_void;
                let (sp, np) = if next { (std::ptr::null_mut(), p) } else { (p, std::ptr::null_mut()) };
                check(unsafe {
                    ffi::qlc_replay_gather_host(first.env.0, idx.as_ptr(), N as u32, ffi::QLC_LAYOUT_F32_BXYH, sp, np,
                                                std::ptr::null_mut(), std::ptr::null_mut(), std::ptr::null_mut())
                })?;
            }
        }
        Ok(out)
    }
}

impl Debug for CudaBreakoutState {
    fn fmt(&self, f: &mut Formatter<'_>) -> std::fmt::Result { write!(f, "CudaBreakoutState {{ {:?} @ t={} }}", self.kind, self.time) }
}

impl DebugVisualizer for CudaBreakoutState {
    fn one_line_info(&self) -> String {
        let (mut cx, mut cy, mut pmin, mut pmax, mut bricks) = (0f32, 0f32, 0f32, 0f32, 0u64);
        let sh = ffi::qlc_state_host {
            ball_cx: &mut cx, ball_cy: &mut cy, ball_dx: std::ptr::null_mut(), ball_dy: std::ptr::null_mut(),
            pad_min_x: &mut pmin, pad_max_x: &mut pmax, pad_speed: std::ptr::null_mut(), bricks: &mut bricks,
            score: std::ptr::null_mut(), episode_step: std::ptr::null_mut(), episode: std::ptr::null_mut(),
            err: std::ptr::null_mut(), finished: std::ptr::null_mut(),
        };
        unsafe { ffi::qlc_env_read_state(self.env.0, &sh) };
        format!("Breakout [{} bricks, ball_pos: [{cx} {cy}], panel_pos: [{} 570]]", bricks.count_ones(), (pmin + pmax) / 2.0)
    }
    fn render_to_console(&self) -> Screen { todo!() } // as in the reference (breakout_environment.rs:91)
}

// `ToMultiDimArray<Tensor<f32>>` (ql-with-tensorflow/src/ml_model/model.rs:12-26) is implemented for
// `CudaBreakoutState` inside ql-with-tensorflow (the trait's home crate), see INTEGRATION.md section 3:
//   dims()                      -> self.dims()
//   to_multi_dim_array()        -> Tensor::new(&[84, 84, 4]).with_values(&CudaBreakoutState::batch_to_f32(&[&Rc::new(self.clone())])?)
//   batch_to_multi_dim_array()  -> Tensor::new(&[N, 84, 84, 4]).with_values(&CudaBreakoutState::batch_to_f32(batch)?)

/// One Breakout env on the GPU behind `ql::prelude::Environment` (breakout_environment.rs:131-207).
pub struct CudaBreakoutEnvironment {
    env: Rc<Handle>,
    state: CudaBreakoutState,
}

impl CudaBreakoutEnvironment {
    /// `BreakoutEnvironment::new(frame_size_x, frame_size_y)` plus the replay ring length (`Parameter::history_buffer_len`).
    pub fn new(frame_size_x: usize, frame_size_y: usize, history_buffer_len: usize, seed: u64, device: i32) -> Result<Self> {
        let cfg = ffi::qlc_config {
            struct_size: std::mem::size_of::<ffi::qlc_config>() as u32,
            device,
            n_envs: 1,
            env_id_base: 0,
            frame_w: frame_size_x as u32,
            frame_h: frame_size_y as u32,
            seed,
            replay_capacity: history_buffer_len as u64,
            max_episode_steps: 0,
            episode_window: 100,
            auto_reset: 0, // the learner resets: learn_episode :142
            reserved: 0,
        };
        let mut h: *mut ffi::qlc_env = std::ptr::null_mut();
        check(unsafe { ffi::qlc_env_create(&cfg, &mut h) })?;
        let env = Rc::new(Handle(h));
        let state = CudaBreakoutState::new(Rc::clone(&env), StateKind::Live, 0);
        Ok(Self { env, state })
    }
    pub(crate) fn handle(&self) -> Rc<Handle> { Rc::clone(&self.env) }
}

impl Environment for CudaBreakoutEnvironment {
    type S = CudaBreakoutState;
    type A = BreakoutAction;

    fn reset(&mut self) {
        check(unsafe { ffi::qlc_env_reset(self.env.0, std::ptr::null(), std::ptr::null()) }).expect("qlc_env_reset");
        self.state = CudaBreakoutState::new(Rc::clone(&self.env), StateKind::Live, self.env.time());
    }

    fn state(&self) -> &Self::S { &self.state }

    fn step(&mut self, action: BreakoutAction) -> (&CudaBreakoutState, f32, bool) {
        let a = action.numeric();
        let (mut reward, mut done) = (0f32, 0u8);
        check(unsafe { ffi::qlc_env_step_host(self.env.0, &a, 1, &mut reward, &mut done) }).expect("qlc_env_step_host");
        self.state = CudaBreakoutState::new(Rc::clone(&self.env), StateKind::Live, self.env.time());
        (&self.state, reward, done != 0)
    }

    fn episode_reward_goal_mean(&self) -> f32 { unsafe { ffi::qlc_env_goal_mean() } }
}
